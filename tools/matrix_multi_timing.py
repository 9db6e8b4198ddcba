"""ici_kendalltau(return_matrix = TRUE) at the config-5 shape (2 000 features x 5 000 samples, 12.5 M
pairs, five 5 000 x 5 000 matrices = 1 GB) on 1 .. N GPUs of the box: wall time of the whole call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import icikendalltau_b200 as ik
from icikendalltau_b200 import _lib, synth

C = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
x, persp = synth.make("config5", C=C)
names = [f"s{i}" for i in range(x.shape[1])]
ndev = _lib.load().icikt_device_count()
ref = None
for g in [1, 2, 4, 8]:
    if g > ndev:
        break
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        r = ik.ici_kendalltau(x, global_na=(np.nan,), perspective=persp, colnames=names, n_gpus=g)
        ts.append(time.perf_counter() - t0)
    if ref is None:
        ref = r
    else:
        for k in _lib.MATRIX_NAMES:
            assert np.array_equal(ref[k], r[k], equal_nan=True), k
    print(f"ici_kendalltau {x.shape[0]} x {x.shape[1]} on {g} GPU(s): best {min(ts[1:]):.3f} s, first call {ts[0]:.3f} s"
          f" (run_time inside: {r['run_time']:.3f} s)", flush=True)
