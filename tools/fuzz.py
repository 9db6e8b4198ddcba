"""Randomised parity campaign on the GPU: lengths around every shape boundary of the pair kernel
(warp/run counts, the 8 192-row limit of the fused column kernel, the shared-memory / in-place /
global-scratch variants), tie structures from none to count data, missing fractions from 0 to
~100 %, both perspectives, against the CPU oracle.  Runs until the time budget is spent.
Usage: python tools/fuzz.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import icikendalltau_b200 as ik
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_gpu_parity import assert_parity, oracle_pairs

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
rng = np.random.default_rng(seed)
SIZES = [2, 3, 31, 32, 33, 255, 256, 257, 511, 1000, 1791, 1792, 1793, 2047, 2048, 2049, 4095, 4096, 5000,
         5376, 5377, 8191, 8192, 8193, 10000, 16383, 16384, 20000, 22527, 22528, 22529, 24576, 28672, 32768, 32769, 40000, 50000,
         57344, 64511, 64512, 64513, 65535]
if os.environ.get("ICIKT_FUZZ_SIZES"):  # e.g. only the lengths of the in-place / global-scratch variants
    SIZES = [int(v) for v in os.environ["ICIKT_FUZZ_SIZES"].split(",")]
t0 = time.time()
cases = 0
fails = 0
while time.time() - t0 < budget:
    n = int(rng.choice(SIZES))
    C = int(rng.integers(2, 5)) if n > 20000 else int(rng.integers(3, 8))
    kind = str(rng.choice(["normal", "round", "counts", "mixed", "few"]))
    x = rng.normal(size=(n, 1)) * float(rng.choice([0.0, 1.0])) + rng.normal(size=(n, C))
    if kind == "round":
        x = np.round(x * float(rng.choice([0.5, 3, 30, 300])))
    elif kind == "counts":
        x = np.floor(np.exp(x * float(rng.choice([0.5, 1.5, 3.0]))))
    elif kind == "mixed":
        x[:, ::2] = np.round(x[:, ::2] * float(rng.choice([1, 10, 100])))
    elif kind == "few":
        x = np.round(x * 0.7)
    gna = ()
    if kind == "counts" and rng.random() < 0.5:
        gna = (np.nan, np.inf, 0.0)
    for c in range(C):
        frac = float(rng.choice([0.0, 0.0, 0.01, 0.3, 0.7, 0.97]))
        if frac > 0:
            if rng.random() < 0.5:  # left-censored like a detection limit
                x[x[:, c] <= np.quantile(x[:, c], frac), c] = np.nan
            else:
                x[rng.random(n) < frac, c] = np.nan
    if rng.random() < 0.15:
        x[:, int(rng.integers(0, C))] = 7.0
    if rng.random() < 0.1:
        x[:, int(rng.integers(0, C))] = np.nan
    x = np.asfortranarray(x)
    persp = str(rng.choice(["global", "local"]))
    what = f"case {cases} seed {seed} n={n} C={C} {kind} {persp} gna={len(gna)}"
    got = ik.run_pairs(x, gna, perspective=persp, include_diag=bool(rng.random() < 0.3), want_counts=True)
    P = C * (C - 1) // 2
    ref = oracle_pairs(x, include_diag=got["raw"].size > P, global_na=gna, perspective=persp)
    try:
        assert_parity(got, ref, what)
    except AssertionError as e:
        fails += 1
        print("FAIL", what, "\n", str(e)[:1500], flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        np.save(f"gpurun_out/fuzz_fail_{seed}_{cases}.npy", x)
        # is it reproducible?  run the same call again a few times
        for rep in range(5):
            again = ik.run_pairs(x, gna, perspective=persp, include_diag=got["raw"].size > P, want_counts=True)
            same = all(np.array_equal(again[k], got[k], equal_nan=True) for k in ("counts", "status", "raw", "pvalue", "taumax"))
            try:
                assert_parity(again, ref, what)
                ok = True
            except AssertionError:
                ok = False
            print(f"  rerun {rep}: identical to the failing result: {same}; parity: {ok}", flush=True)
    cases += 1
print(f"failures: {fails}")
print(f"fuzz: {cases} random matrices passed in {time.time() - t0:.0f} s (seed {seed})")
