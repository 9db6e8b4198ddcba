"""Times the host-facing calls around the pair kernel at the config-5 shape (2 000 features x
5 000 samples): per-pair arrays + host scatter against the device-side matrix fill, and the
pairwise_completeness kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icikendalltau_b200 import _lib, api, synth

name = sys.argv[1] if len(sys.argv) > 1 else "config5"
x, persp = synth.make(name)
n, C = x.shape
gna = (np.nan, np.inf, 0.0)
names = [f"s{i}" for i in range(C)]


def best(f, reps=3):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = f()
        t.append(time.perf_counter() - t0)
    return min(t), r


def host_path():
    r = _lib.run_pairs(x, gna, perspective=persp)
    pi, pj = np.triu_indices(C, k=1)
    cols = dict(cor=r["raw"] / r["max_taumax"], raw=r["raw"], pvalue=r["pvalue"], taumax=r["taumax"],
                completeness=r["completeness"])
    return api._reshape(names, pi, pj, cols)


t_host, m_host = best(host_path, 2)
t_dev, m_dev = best(lambda: _lib.run_matrices(x, gna, True, True, None, perspective=persp))
print(f"{name} n={n} C={C}: per-pair arrays + host scatter {t_host:.3f} s; device matrix fill {t_dev:.3f} s", m_dev["timings"])
iu = np.triu_indices(C, k=1)
for k in ("raw", "pvalue", "taumax", "completeness", "cor"):
    assert np.array_equal(m_host[k][iu], m_dev[k][iu], equal_nan=True), k
t_api, _ = best(lambda: api.ici_kendalltau(x, colnames=names, perspective=persp), 2)
print(f"ici_kendalltau() whole call {t_api:.3f} s")
t_pc, pc = best(lambda: _lib.pairwise_completeness(x, gna, want_matrix=True))
print(f"pairwise_completeness: {pc['missing'].size} pairs + matrix in {t_pc:.3f} s")
excl = np.isnan(x) | np.isinf(x) | (x == 0)
k = np.random.default_rng(0).integers(0, pc["missing"].size - C, 200)
assert all(pc["missing"][q] == (excl[:, iu[0][q]] | excl[:, iu[1][q]]).sum() for q in k)
