"""One-shot icikt_all_pairs on a BASELINE workload with and without the pipelined call: wall time and
the library's own event timings (pinned and pageable input)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from icikendalltau_b200 import _lib, synth

name = sys.argv[1] if len(sys.argv) > 1 else "target"
x, persp = synth.make(name)
xp = torch.from_numpy(np.ascontiguousarray(x.T)).pin_memory().numpy().T
for mode in ("plain", "pipe"):
    if mode == "plain":
        os.environ["ICIKT_NO_PIPELINE"] = "1"
    else:
        os.environ.pop("ICIKT_NO_PIPELINE", None)
    _lib.release_workspace()
    for label, data in (("pinned", xp), ("pageable", x)):
        best, tm = 1e9, None
        for _ in range(5):
            t0 = time.perf_counter()
            r = _lib.run_pairs(data, (), perspective=persp)
            dt = time.perf_counter() - t0
            if dt < best:
                best, tm = dt, r["timings"]
        print(f"{name} {mode:5s} {label:8s} wall {best * 1e3:8.2f} ms", {k: round(v, 2) for k, v in tm.items()})
_lib.release_workspace()
