"""Times ONE icikt_all_pairs_multi call on the north-star target (2 000 samples x 20 000 features)
for 1 and all visible GPUs, host memory to host memory, and checks the results agree."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from icikendalltau_b200 import _lib, synth

name = sys.argv[1] if len(sys.argv) > 1 else "target"
x, persp = synth.make(name)
ng = torch.cuda.device_count()
ref = None
for devs in ([0], list(range(ng))):
    best = None
    for rep in range(3):  # first call builds the cached plans
        t0 = time.perf_counter()
        out = _lib.run_pairs(x, perspective=persp, devices=devs)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    P = out["raw"].shape[0]
    print("devices", len(devs), "pairs", P, "seconds %.3f" % best, "pairs/s %.3e" % (P / best), out.get("timings"))
    if ref is None:
        ref = out
    else:
        assert np.array_equal(ref["raw"], out["raw"], equal_nan=True)
        assert np.array_equal(ref["pvalue"], out["pvalue"], equal_nan=True)
        print("identical to the single-GPU result")
