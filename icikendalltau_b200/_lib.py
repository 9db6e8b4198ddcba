"""ctypes binding of libicikt_b200.so (the C ABI in include/icikt_b200.h).

The library is the product's only compute path.  If it is missing, or no CUDA device is
usable, every call raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ICIKT_LIB_PATH: another build of the same library (A/B runs of tuning variants); default: the in-tree build
LIB_PATH = os.environ.get("ICIKT_LIB_PATH") or os.path.join(_HERE, "libicikt_b200.so")

OK = 0
ERR_NO_DEVICE, ERR_BAD_ARG, ERR_TOO_LONG, ERR_CUDA, ERR_ALLOC = -1, -2, -3, -4, -5
NCOUNTS = 7
KERNEL_TILED, KERNEL_NAIVE = 0, 1
PERSPECTIVE = {"global": 0, "local": 1, "complete": 2}
ALTERNATIVE = {"two.sided": 0, "less": 1, "greater": 2}

# every symbol include/icikt_b200.h declares (checked by tests/test_abi_cpu.py)
EXPORTS = [
    "icikt_default_opts", "icikt_abi_version", "icikt_device_count", "icikt_max_n",
    "icikt_last_error", "icikt_all_pairs", "icikt_pair_list", "icikt_plan_create",
    "icikt_plan_num_pairs", "icikt_plan_upload", "icikt_plan_set_device_matrix",
    "icikt_plan_columns", "icikt_plan_pairs", "icikt_plan_sync", "icikt_plan_download",
    "icikt_plan_column_info", "icikt_plan_stream", "icikt_plan_timings", "icikt_plan_destroy",
    "icikt_pnorm_device", "icikt_release_workspace", "icikt_measure_smem_bandwidth",
    "icikt_pair_from_index", "icikt_all_pairs_multi", "icikt_matrices", "icikt_plan_download_matrices",
    "icikt_pairwise_completeness", "icikt_plan_upload_columns", "icikt_plan_columns_range",
    "icikt_plan_tables", "icikt_plan_columns_finish", "icikt_measure_issue_rate",
    "icikt_matrices_multi", "icikt_stage_table", "icikt_launch_shape",
]
NSTATUS = 10


class IciktError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libicikt_b200 error {code}: {msg}")
        self.code = code


class Opts(ctypes.Structure):
    _fields_ = [("perspective", ctypes.c_int32), ("alternative", ctypes.c_int32),
                ("continuity", ctypes.c_int32), ("include_diag", ctypes.c_int32),
                ("na_inf", ctypes.c_int32), ("device", ctypes.c_int32),
                ("kernel", ctypes.c_int32), ("want_counts", ctypes.c_int32),
                ("pair_lo", ctypes.c_int64), ("pair_hi", ctypes.c_int64)]


class Timings(ctypes.Structure):
    _fields_ = [("h2d_ms", ctypes.c_float), ("columns_ms", ctypes.c_float),
                ("pairs_ms", ctypes.c_float), ("epilogue_ms", ctypes.c_float), ("d2h_ms", ctypes.c_float),
                ("total_ms", ctypes.c_float), ("n_launches", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]

    def as_dict(self):
        return dict(h2d_ms=self.h2d_ms, columns_ms=self.columns_ms, pairs_ms=self.pairs_ms,
                    epilogue_ms=self.epilogue_ms, d2h_ms=self.d2h_ms, total_ms=self.total_ms,
                    n_launches=self.n_launches)


class Table(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("bytes_per_column", ctypes.c_int64)]


MAX_TABLES = 16

_lib = None
# array arguments travel as plain addresses (void*): building typed ctypes pointers from NumPy
# arrays costs microseconds each, which shows in the one-shot call on small matrices
_dp = _ip = _lp = ctypes.c_void_p


def load():
    """Load the shared library; raises if it has not been built (python -m icikendalltau_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IciktError(ERR_NO_DEVICE, f"{LIB_PATH} is missing: build it with "
                         "`python -m icikendalltau_b200.build` (nvcc, sm_100a); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    L.icikt_default_opts.argtypes = [ctypes.POINTER(Opts)]
    L.icikt_default_opts.restype = None
    L.icikt_abi_version.restype = ctypes.c_int
    L.icikt_device_count.restype = ctypes.c_int
    L.icikt_max_n.restype = ctypes.c_int64
    L.icikt_last_error.restype = ctypes.c_char_p
    common_out = [_dp, _dp, _dp, _dp, _ip, _lp, _dp, ctypes.POINTER(Timings)]
    L.icikt_all_pairs.argtypes = [_dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _dp,
                                  ctypes.c_int32, ctypes.POINTER(Opts)] + common_out
    L.icikt_all_pairs.restype = ctypes.c_int
    L.icikt_all_pairs_multi.argtypes = [_dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _dp,
                                        ctypes.c_int32, ctypes.POINTER(Opts), _ip, ctypes.c_int32] + common_out
    L.icikt_all_pairs_multi.restype = ctypes.c_int
    L.icikt_pair_list.argtypes = [_dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _dp,
                                  ctypes.c_int32, _ip, _ip, ctypes.c_int64,
                                  ctypes.POINTER(Opts)] + common_out
    L.icikt_pair_list.restype = ctypes.c_int
    L.icikt_plan_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int64, ctypes.c_int64, _ip, _ip,
                                    ctypes.c_int64, ctypes.POINTER(Opts)]
    L.icikt_plan_create.restype = ctypes.c_int
    L.icikt_plan_num_pairs.argtypes = [vp]
    L.icikt_plan_num_pairs.restype = ctypes.c_int64
    L.icikt_plan_upload.argtypes = [vp, _dp, ctypes.c_int64]
    L.icikt_plan_upload.restype = ctypes.c_int
    L.icikt_plan_set_device_matrix.argtypes = [vp, vp, ctypes.c_int64]
    L.icikt_plan_set_device_matrix.restype = ctypes.c_int
    L.icikt_plan_columns.argtypes = [vp, _dp, ctypes.c_int32]
    L.icikt_plan_columns.restype = ctypes.c_int
    L.icikt_plan_upload_columns.argtypes = [vp, _dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64]
    L.icikt_plan_upload_columns.restype = ctypes.c_int
    L.icikt_plan_columns_range.argtypes = [vp, _dp, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64]
    L.icikt_plan_columns_range.restype = ctypes.c_int
    L.icikt_plan_tables.argtypes = [vp, ctypes.POINTER(Table), ctypes.c_int32]
    L.icikt_plan_tables.restype = ctypes.c_int
    L.icikt_plan_columns_finish.argtypes = [vp]
    L.icikt_plan_columns_finish.restype = ctypes.c_int
    L.icikt_plan_pairs.argtypes = [vp]
    L.icikt_plan_pairs.restype = ctypes.c_int
    L.icikt_plan_sync.argtypes = [vp]
    L.icikt_plan_sync.restype = ctypes.c_int
    L.icikt_plan_download.argtypes = [vp, _dp, _dp, _dp, _dp, _ip, _lp, _dp]
    L.icikt_plan_download.restype = ctypes.c_int
    L.icikt_plan_download_matrices.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, _ip, _dp, _dp, _dp, _dp, _dp,
                                               _lp, _dp]
    L.icikt_plan_download_matrices.restype = ctypes.c_int
    L.icikt_matrices.argtypes = [_dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _dp, ctypes.c_int32,
                                 _ip, _ip, ctypes.c_int64, ctypes.POINTER(Opts), ctypes.c_int32, ctypes.c_int32,
                                 _ip, _dp, _dp, _dp, _dp, _dp, _lp, _dp, ctypes.POINTER(Timings)]
    L.icikt_matrices.restype = ctypes.c_int
    L.icikt_matrices_multi.argtypes = [_dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _dp, ctypes.c_int32,
                                       ctypes.POINTER(Opts), _ip, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                       _ip, _dp, _dp, _dp, _dp, _dp, _lp, _dp, ctypes.POINTER(Timings)]
    L.icikt_matrices_multi.restype = ctypes.c_int
    L.icikt_pairwise_completeness.argtypes = [_dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _dp,
                                              ctypes.c_int32, ctypes.c_int32, _ip, _ip, ctypes.c_int64, _ip,
                                              _dp, _dp]
    L.icikt_pairwise_completeness.restype = ctypes.c_int
    L.icikt_plan_column_info.argtypes = [vp, _ip]
    L.icikt_plan_column_info.restype = ctypes.c_int
    L.icikt_plan_stream.argtypes = [vp]
    L.icikt_plan_stream.restype = vp
    L.icikt_plan_timings.argtypes = [vp, ctypes.POINTER(Timings)]
    L.icikt_plan_timings.restype = ctypes.c_int
    L.icikt_plan_destroy.argtypes = [vp]
    L.icikt_plan_destroy.restype = None
    L.icikt_pnorm_device.argtypes = [_dp, ctypes.c_int64, ctypes.c_int32, _dp, ctypes.c_int32]
    L.icikt_pnorm_device.restype = ctypes.c_int
    L.icikt_pair_from_index.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, _ip, _ip]
    L.icikt_pair_from_index.restype = ctypes.c_int
    L.icikt_release_workspace.argtypes = []
    L.icikt_release_workspace.restype = None
    L.icikt_measure_smem_bandwidth.argtypes = [ctypes.c_int32, _dp, _dp]
    L.icikt_measure_smem_bandwidth.restype = ctypes.c_int
    L.icikt_measure_issue_rate.argtypes = [ctypes.c_int32, _dp, _dp, _dp]
    L.icikt_measure_issue_rate.restype = ctypes.c_int
    L.icikt_stage_table.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                    _lp, ctypes.c_int64, _lp, _lp]
    L.icikt_stage_table.restype = ctypes.c_int64
    L.icikt_launch_shape.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _ip]
    L.icikt_launch_shape.restype = ctypes.c_int
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise IciktError(rc, load().icikt_last_error().decode(errors="replace"))


def make_opts(perspective="global", alternative="two.sided", continuity=False, include_diag=False,
              na_inf=False, device=0, kernel=KERNEL_TILED, want_counts=False, pair_lo=0, pair_hi=0):
    # same values as icikt_default_opts() for everything not given (all zero)
    return Opts(PERSPECTIVE.get(perspective, 0),   # any other string behaves as global
                ALTERNATIVE.get(alternative, 3),   # unknown alternative: p-value stays 0
                int(bool(continuity)), int(bool(include_diag)), int(bool(na_inf)), int(device), int(kernel),
                int(bool(want_counts)), int(pair_lo), int(pair_hi))


def _global_na_array(global_na):
    if len(global_na) == 0:
        return None, None, 0
    g = np.ascontiguousarray(np.asarray(list(global_na), dtype=np.float64))
    return g, _ptr(g), int(g.size)


def _ptr(a, t=None):
    """address of a NumPy array's buffer (None -> NULL); `t` names the element type for the reader"""
    return a.__array_interface__["data"][0] if a is not None else None


def run_pairs(data, global_na=(), pi=None, pj=None, want_counts=False, devices=None, **opt_kw):
    """One-shot call through icikt_all_pairs / icikt_pair_list with host buffers.

    data: (n, C) array (copied to Fortran order if needed).  Returns a dict of NumPy arrays in
    pair order: raw, pvalue, taumax, completeness, status, [counts], max_taumax, timings.
    devices: list of CUDA ordinals -> icikt_all_pairs_multi (all pairs only): every device takes a
    contiguous slice of the pair order.
    """
    L = load()
    data = np.asfortranarray(data, dtype=np.float64)
    if data.ndim != 2:
        raise ValueError("data must be a 2-D array (features x samples)")
    n, C = data.shape
    o = make_opts(want_counts=want_counts, **opt_kw)
    if pi is None:
        ptot = C * (C - 1) // 2 + (C if o.include_diag else 0)
        lo, hi = o.pair_lo, o.pair_hi
        P = ptot if (lo == 0 and hi == 0) else hi - lo
    else:
        pi = np.ascontiguousarray(pi, dtype=np.int32)
        pj = np.ascontiguousarray(pj, dtype=np.int32)
        P = int(pi.size)
    # the five result arrays are views of one allocation: one address lookup instead of five
    P = max(P, 0)
    buf = np.empty(4 * P + (P + 1) // 2, dtype=np.float64)
    raw, pv, tm, comp = buf[:P], buf[P:2 * P], buf[2 * P:3 * P], buf[3 * P:4 * P]
    status = buf[4 * P:].view(np.int32)[:P]
    counts = np.empty((P, NCOUNTS), dtype=np.int64) if want_counts else None
    a_raw = _ptr(buf)
    a_pv, a_tm, a_comp, a_status = a_raw + 8 * P, a_raw + 16 * P, a_raw + 24 * P, a_raw + 32 * P
    mx = ctypes.c_double(float("nan"))
    t = Timings()
    g, gp, ng = _global_na_array(global_na)
    if pi is None and devices is not None:
        dev = np.ascontiguousarray(list(devices), dtype=np.int32)
        rc = L.icikt_all_pairs_multi(_ptr(data, _dp), n, C, n, gp, ng, ctypes.byref(o), _ptr(dev, _ip),
                                     int(dev.size), a_raw, a_pv, a_tm, a_comp, a_status, _ptr(counts, _lp),
                                     ctypes.byref(mx), ctypes.byref(t))
    elif pi is None:
        rc = L.icikt_all_pairs(_ptr(data, _dp), n, C, n, gp, ng, ctypes.byref(o), a_raw, a_pv, a_tm, a_comp,
                               a_status, _ptr(counts, _lp), ctypes.byref(mx), ctypes.byref(t))
    else:
        rc = L.icikt_pair_list(_ptr(data, _dp), n, C, n, gp, ng, _ptr(pi, _ip), _ptr(pj, _ip), P,
                               ctypes.byref(o), a_raw, a_pv, a_tm, a_comp, a_status, _ptr(counts, _lp),
                               ctypes.byref(mx), ctypes.byref(t))
    check(rc)
    out = dict(raw=raw, pvalue=pv, taumax=tm, completeness=comp, status=status,
               max_taumax=mx.value, timings=t.as_dict())
    if want_counts:
        out["counts"] = counts
    return out


MATRIX_NAMES = ("cor", "raw", "pvalue", "taumax", "completeness")


def run_matrices(data, global_na=(), scale_max=True, diag_good=True, n_good=None, pi=None, pj=None,
                 want=MATRIX_NAMES, devices=None, **opt_kw):
    """icikt_matrices: the pair results as symmetric C x C matrices filled on the device
    (scale_and_reshape, R/kendalltau.R:357-421).  Returns the requested matrices plus
    status_counts (pairs per status class), max_taumax and timings.
    devices: list of CUDA ordinals -> icikt_matrices_multi (all pairs only): the pair order is sliced
    over the devices and every device returns its own block of columns of each matrix."""
    L = load()
    data = np.asfortranarray(data, dtype=np.float64)
    if data.ndim != 2:
        raise ValueError("data must be a 2-D array (features x samples)")
    n, C = data.shape
    o = make_opts(**opt_kw)
    mats = {k: (np.empty((C, C), dtype=np.float64, order="F") if k in want else None) for k in MATRIX_NAMES}
    hist = np.zeros(NSTATUS, dtype=np.int64)
    mx = ctypes.c_double(float("nan"))
    t = Timings()
    g, gp, ng = _global_na_array(global_na)
    if pi is not None:
        pi = np.ascontiguousarray(pi, dtype=np.int32)
        pj = np.ascontiguousarray(pj, dtype=np.int32)
    ngood = None if n_good is None else np.ascontiguousarray(n_good, dtype=np.int32)
    if devices is not None:
        if pi is not None:
            raise ValueError("several devices: all pairs only")
        dev = np.ascontiguousarray(list(devices), dtype=np.int32)
        check(L.icikt_matrices_multi(_ptr(data, _dp), n, C, n, gp, ng, ctypes.byref(o), _ptr(dev, _ip), int(dev.size),
                                     int(bool(scale_max)), int(bool(diag_good)), _ptr(ngood, _ip),
                                     *(_ptr(mats[k], _dp) for k in MATRIX_NAMES), _ptr(hist, _lp), ctypes.byref(mx),
                                     ctypes.byref(t)))
        out = {k: v for k, v in mats.items() if v is not None}
        out.update(status_counts=hist, max_taumax=mx.value, timings=t.as_dict())
        return out
    check(L.icikt_matrices(_ptr(data, _dp), n, C, n, gp, ng, _ptr(pi, _ip), _ptr(pj, _ip),
                           0 if pi is None else int(pi.size), ctypes.byref(o), int(bool(scale_max)),
                           int(bool(diag_good)), _ptr(ngood, _ip), *(_ptr(mats[k], _dp) for k in MATRIX_NAMES),
                           _ptr(hist, _lp), ctypes.byref(mx), ctypes.byref(t)))
    out = {k: v for k, v in mats.items() if v is not None}
    out.update(status_counts=hist, max_taumax=mx.value, timings=t.as_dict())
    return out


def pairwise_completeness(data, global_na=(), pi=None, pj=None, want_matrix=False, want_pairs=True, device=0):
    """icikt_pairwise_completeness: rows missing in either column and 1 - missing/n per pair
    (all pairs in combn order followed by the diagonal when pi is None)."""
    L = load()
    data = np.asfortranarray(data, dtype=np.float64)
    n, C = data.shape
    if pi is not None:
        pi = np.ascontiguousarray(pi, dtype=np.int32)
        pj = np.ascontiguousarray(pj, dtype=np.int32)
        P = int(pi.size)
    else:
        P = C * (C - 1) // 2 + C
    missing = np.empty(P, dtype=np.int32) if want_pairs else None
    comp = np.empty(P, dtype=np.float64) if want_pairs else None
    mat = np.empty((C, C), dtype=np.float64, order="F") if want_matrix else None
    g, gp, ng = _global_na_array(global_na)
    check(L.icikt_pairwise_completeness(_ptr(data, _dp), n, C, n, gp, ng, int(device), _ptr(pi, _ip), _ptr(pj, _ip),
                                        P, _ptr(missing, _ip), _ptr(comp, _dp), _ptr(mat, _dp)))
    out = dict(missing=missing, completeness=comp) if want_pairs else {}
    if want_matrix:
        out["matrix"] = mat
    return out


class Plan:
    """Thin RAII wrapper of the plan API (device-resident matrix, tables and results)."""

    def __init__(self, n, C, pi=None, pj=None, **opt_kw):
        self._L = load()
        self._h = ctypes.c_void_p()
        self.opts = make_opts(**opt_kw)
        if pi is not None:
            self._pi = np.ascontiguousarray(pi, dtype=np.int32)
            self._pj = np.ascontiguousarray(pj, dtype=np.int32)
            check(self._L.icikt_plan_create(ctypes.byref(self._h), n, C, _ptr(self._pi, _ip),
                                            _ptr(self._pj, _ip), self._pi.size, ctypes.byref(self.opts)))
        else:
            check(self._L.icikt_plan_create(ctypes.byref(self._h), n, C, None, None, 0,
                                            ctypes.byref(self.opts)))
        self.n, self.C = n, C
        self.P = int(self._L.icikt_plan_num_pairs(self._h))

    def upload(self, data):
        data = np.asfortranarray(data, dtype=np.float64)
        assert data.shape == (self.n, self.C)
        check(self._L.icikt_plan_upload(self._h, _ptr(data, _dp), self.n))
        self.sync()  # the host array may go away

    def set_device_matrix(self, data_ptr, ld):
        check(self._L.icikt_plan_set_device_matrix(self._h, ctypes.c_void_p(int(data_ptr)), int(ld)))

    def columns(self, global_na=()):
        g, gp, ng = _global_na_array(global_na)
        check(self._L.icikt_plan_columns(self._h, gp, ng))

    # ---- sharded K1 (one rank per GPU): upload + preprocess a slice of the columns, exchange the
    # table slices (sharding.exchange_tables), then columns_finish() releases pairs()
    def upload_columns(self, data, col_lo, col_hi):
        """`data` is the whole (n, C) Fortran-ordered host matrix; only the slice is copied."""
        assert data.flags["F_CONTIGUOUS"] and data.dtype == np.float64 and data.shape == (self.n, self.C)
        check(self._L.icikt_plan_upload_columns(self._h, _ptr(data, _dp), self.n, int(col_lo), int(col_hi)))

    def columns_range(self, global_na, col_lo, col_hi):
        g, gp, ng = _global_na_array(global_na)
        check(self._L.icikt_plan_columns_range(self._h, gp, ng, int(col_lo), int(col_hi)))

    def tables(self):
        """[(device pointer of column 0, bytes per column)] of the per-column tables to exchange."""
        arr = (Table * MAX_TABLES)()
        n = self._L.icikt_plan_tables(self._h, arr, MAX_TABLES)
        if n < 0:
            check(n)
        return [(int(arr[k].ptr), int(arr[k].bytes_per_column)) for k in range(n)]

    def columns_finish(self):
        check(self._L.icikt_plan_columns_finish(self._h))

    def pairs(self):
        check(self._L.icikt_plan_pairs(self._h))

    def sync(self):
        check(self._L.icikt_plan_sync(self._h))

    def stream(self):
        return int(self._L.icikt_plan_stream(self._h) or 0)

    def download(self, want_counts=False):
        P = self.P
        raw, pv, tm, comp = (np.empty(P, dtype=np.float64) for _ in range(4))
        status = np.empty(P, dtype=np.int32)
        counts = np.empty((P, NCOUNTS), dtype=np.int64) if want_counts else None
        mx = ctypes.c_double(float("nan"))
        check(self._L.icikt_plan_download(self._h, _ptr(raw, _dp), _ptr(pv, _dp), _ptr(tm, _dp),
                                          _ptr(comp, _dp), _ptr(status, _ip), _ptr(counts, _lp),
                                          ctypes.byref(mx)))
        out = dict(raw=raw, pvalue=pv, taumax=tm, completeness=comp, status=status, max_taumax=mx.value)
        if want_counts:
            out["counts"] = counts
        return out

    def column_n_na(self):
        a = np.empty(self.C, dtype=np.int32)
        check(self._L.icikt_plan_column_info(self._h, _ptr(a, _ip)))
        return a

    def timings(self):
        t = Timings()
        check(self._L.icikt_plan_timings(self._h, ctypes.byref(t)))
        return t.as_dict()

    def close(self):
        if self._h:
            self._L.icikt_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pnorm_device(z, lower_tail=True, device=0):
    z = np.ascontiguousarray(z, dtype=np.float64)
    out = np.empty_like(z)
    check(load().icikt_pnorm_device(_ptr(z, _dp), z.size, int(bool(lower_tail)), _ptr(out, _dp), device))
    return out


def release_workspace():
    load().icikt_release_workspace()


def measure_smem_bandwidth(device=0):
    """(GB/s with 32-bit accesses, GB/s with 128-bit accesses) of a conflict-free load+store sweep."""
    a, b = ctypes.c_double(0), ctypes.c_double(0)
    check(load().icikt_measure_smem_bandwidth(device, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def measure_issue_rate(device=0):
    """(LOP3 only, IMAD only, interleaved) sustained issue rate in G warp-instructions/s over the GPU."""
    a, f, m = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
    check(load().icikt_measure_issue_rate(device, ctypes.byref(a), ctypes.byref(f), ctypes.byref(m)))
    return a.value, f.value, m.value


def stage_table(C, include_diag=False, cta_slots=296, n_blocks=8):
    """The launches of the pipelined one-shot call for C columns (host only, no device needed):
    (units [U, 4] = slot, first column, other column of the first pair, pairs;
     launches [L, 6] = col_lo, col_hi, unit_lo, unit_hi, slot_lo, slot_hi)."""
    L = load()
    nl = ctypes.c_int64(0)
    nu = L.icikt_stage_table(C, int(bool(include_diag)), cta_slots, n_blocks, 0, None, 0, None, ctypes.byref(nl))
    if nu < 0:
        check(int(nu))
    units = np.zeros((max(nu, 1), 4), dtype=np.int64)
    launches = np.zeros((max(nl.value, 1), 6), dtype=np.int64)
    L.icikt_stage_table(C, int(bool(include_diag)), cta_slots, n_blocks, nu, units.ctypes.data, nl.value,
                        launches.ctypes.data, ctypes.byref(nl))
    return units[:nu], launches[:nl.value]


def launch_shape(n, tier=0, n_sm=148, complete_obs=False):
    """The pair kernel's launch shape for n rows (host only): dict(warps, runs, region_bytes, variant, cap, stage_rows);
    variant: 'smem' two sequence buffers in shared memory, 'inplace', 'gmem' global scratch."""
    out = np.zeros(8, dtype=np.int32)
    check(load().icikt_launch_shape(int(n), int(tier), int(n_sm), int(bool(complete_obs)), out.ctypes.data))
    return dict(warps=int(out[0]), runs=int(out[1]), region_bytes=int(out[2]),
                variant=("smem", "inplace", "gmem")[int(out[3])], cap=int(out[4]), stage_rows=int(out[5]))


def pair_from_index(C, index, include_diag=False):
    i, j = ctypes.c_int32(), ctypes.c_int32()
    check(load().icikt_pair_from_index(C, int(bool(include_diag)), int(index), ctypes.byref(i), ctypes.byref(j)))
    return i.value, j.value
