set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$? | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python - <<'PY' > gpurun_out/multi_call.log 2>&1
import time, numpy as np
import icikendalltau_b200 as ik
from icikendalltau_b200 import synth, _lib
nd = _lib.load().icikt_device_count()
x, _ = synth.make("target", C=1200)
for devs in ([0], list(range(nd))):
    ik.run_pairs(x, (), devices=devs)
    t0 = time.perf_counter(); r = ik.run_pairs(x, (), devices=devs); dt = time.perf_counter() - t0
    print("devices", devs, "pairs", r["raw"].size, "seconds %.3f" % dt, "pairs/s %.4g" % (r["raw"].size / dt), r["timings"])
PY
cat gpurun_out/multi_call.log
