"""GPU diagnostic sweep: compares libicikt_b200 with the CPU oracle case by case and prints
which integer count (if any) differs.  Run on the GPU box: python tools/gpu_diag.py [--tiny]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import icikendalltau_b200 as ik  # noqa: E402
from icikendalltau_b200 import _lib  # noqa: E402
from oracle import oracle as O  # noqa: E402

NAMES = ["dis", "ntie", "xtie", "ytie", "tot", "n_entry", "b"]


def gen(n, C, kind, na, seed):
    rng = np.random.default_rng(seed)
    base = rng.normal(size=(n, 1))
    x = base + rng.normal(size=(n, C)) * 0.7
    if kind == "ties":
        x = np.round(x * 2)
    elif kind == "heavy":
        x = np.floor(np.exp(x))
    elif kind == "mixed":
        x[:, ::2] = np.round(x[:, ::2] * 3)
    if na > 0:
        thr = np.quantile(x, na)
        x = np.where(x <= thr, np.nan, x) if kind != "normal" else np.where(x < thr, np.nan, x)
    return np.asfortranarray(x)


def compare(tag, x, persp="global", alt="two.sided", cont=False, kernel=_lib.KERNEL_TILED, pi=None,
            pj=None, include_diag=False, global_na=()):
    t0 = time.time()
    got = ik.run_pairs(x, global_na, pi=pi, pj=pj, want_counts=True, perspective=persp,
                       alternative=alt, continuity=cont, kernel=kernel, include_diag=include_diag)
    t1 = time.time()
    ex = np.array(x, copy=True)
    if len(global_na):
        ex[O.setup_missing_matrix(ex, global_na)] = np.nan
    if pi is None:
        opi, opj = O.setup_comparisons(x.shape[1], None, not include_diag)
    else:
        opi, opj = np.asarray(pi, np.int32), np.asarray(pj, np.int32)
    ref = O.pair_loop(ex, opi, opj, perspective=persp, alternative=alt, continuity=cont, ncore=8,
                      want_counts=True)
    ok = True
    msgs = []
    if not np.array_equal(got["status"], ref["status"]):
        bad = np.nonzero(got["status"] != ref["status"])[0]
        msgs.append(f"status differs at {bad[:5]} got {got['status'][bad[:5]]} ref {ref['status'][bad[:5]]}")
        ok = False
    good = (ref["status"] == 0) & (got["status"] == 0)
    for k, nm in enumerate(NAMES):
        d = got["counts"][good, k] != ref["counts"][good, k]
        if d.any():
            idx = np.nonzero(good)[0][d][:4]
            msgs.append(f"{nm}: {int(d.sum())}/{int(good.sum())} differ, e.g. pair {idx} got "
                        f"{got['counts'][idx, k]} ref {ref['counts'][idx, k]}")
            ok = False
    for nm, tol in (("raw", 1e-12), ("taumax", 1e-12), ("completeness", 1e-15), ("pvalue", 1e-9)):
        a, b = got[nm][good], ref[nm][good]
        both_nan = np.isnan(a) & np.isnan(b)
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        rel = np.where(both_nan | (a == b), 0.0, rel)
        if np.isnan(rel).any() or (rel.size and rel.max() > tol):
            msgs.append(f"{nm}: max rel err {np.nanmax(rel) if rel.size else 0:.3e} (nan mismatch "
                        f"{int(np.isnan(rel).sum())})")
            ok = False
    mx = np.nanmax(ref["taumax"]) if good.any() else np.nan
    if good.any() and not (got["max_taumax"] == mx or abs(got["max_taumax"] - mx) <= 1e-12 * mx):
        msgs.append(f"max_taumax got {got['max_taumax']} ref {mx}")
        ok = False
    print(f"[{'OK' if ok else 'FAIL'}] {tag}: n={x.shape[0]} C={x.shape[1]} P={got['raw'].size} "
          f"gpu {t1 - t0:.3f}s timings={got['timings']}", flush=True)
    for m in msgs:
        print("      ", m, flush=True)
    return ok


def main():
    tiny = "--tiny" in sys.argv
    L = ik.load()
    print("devices:", L.icikt_device_count(), "max_n:", L.icikt_max_n(), flush=True)
    allok = True
    cases = [
        ("normal-noNA", 100, 6, "normal", 0.0),
        ("normal-NA25", 100, 6, "normal", 0.25),
        ("ties-noNA", 100, 6, "ties", 0.0),
        ("ties-NA25", 100, 6, "ties", 0.25),
        ("heavy-NA", 333, 5, "heavy", 0.3),
        ("mixed-NA", 1000, 8, "mixed", 0.2),
        ("n33", 33, 4, "ties", 0.2),
        ("n2", 2, 3, "normal", 0.0),
        ("n3", 3, 3, "ties", 0.0),
    ]
    if not tiny:
        cases += [
            ("n2049-normal", 2049, 6, "normal", 0.25),
            ("n2049-ties", 2049, 6, "ties", 0.25),
            ("n5000-normal", 5000, 12, "normal", 0.2),
            ("n5000-heavy", 5000, 6, "heavy", 0.2),
            ("n9000-mixed", 9000, 6, "mixed", 0.25),
            ("n20000-normal", 20000, 6, "normal", 0.25),
            ("n20000-heavy", 20000, 4, "heavy", 0.25),
            ("n30000-mixed", 30000, 4, "mixed", 0.25),
        ]
    for tag, n, C, kind, na in cases:
        x = gen(n, C, kind, na, seed=n + C)
        allok &= compare(tag, x)
        allok &= compare(tag + "/local", x, persp="local")
    x = gen(500, 7, "mixed", 0.3, 5)
    allok &= compare("naive", x, kernel=_lib.KERNEL_NAIVE)
    allok &= compare("naive/local", x, persp="local", kernel=_lib.KERNEL_NAIVE)
    allok &= compare("less+cont", x, alt="less", cont=True)
    allok &= compare("greater", x, alt="greater")
    allok &= compare("diag", x, include_diag=True, persp="local")
    allok &= compare("pairlist", x, pi=[0, 0, 3, 6, 2, 2], pj=[1, 5, 3, 0, 4, 2])
    xz = x.copy()
    xz[::5, 1] = 0.0
    xz[::9, 2] = np.inf
    xz[:, 4] = np.nan
    xz[:, 5] = 3.0
    allok &= compare("global_na+degenerate", xz, global_na=(np.nan, np.inf, 0.0))
    allok &= compare("global_na+degenerate/local", xz, persp="local", global_na=(np.nan, np.inf, 0.0))
    print("ALL OK" if allok else "SOME FAILED", flush=True)
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
