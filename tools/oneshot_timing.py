"""Event timings of the one-shot call (host memory in, host memory out) on a small workload,
with and without the split run.  Usage: python tools/oneshot_timing.py [workload]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from icikendalltau_b200 import _lib, synth

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
x, persp = synth.make(name)
xp = torch.from_numpy(np.ascontiguousarray(x.T)).pin_memory().numpy().T  # pinned, column-major
gna = (np.nan, np.inf, 0.0) if name == "config1" else ()
for rep in range(30):
    _lib.run_pairs(xp, gna, perspective=persp)
ts, walls = [], []
for rep in range(200):
    t0 = time.perf_counter()
    r = _lib.run_pairs(xp, gna, perspective=persp)
    walls.append(time.perf_counter() - t0)
    ts.append(r["timings"])
med = {k: float(np.median([t[k] for t in ts])) for k in ts[0]}
print(name, "split" if not os.environ.get("ICIKT_NO_SPLIT") else "whole", "wall median %.1f us" % (1e6 * np.median(walls)),
      {k: round(v * 1000, 1) if k.endswith("_ms") else v for k, v in med.items()})
