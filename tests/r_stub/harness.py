"""TEST INFRASTRUCTURE: builds r_package/src/icikt_shim.c against the stand-in R C API of this
directory (no R in the image) and drives its REGISTERED .Call routines from Python.

    sh = harness.load()                # gcc the shim + stub, run R_init_ICIKendallTauB200
    res = sh.dot_call("C_icikt_all_pairs", sh.real_matrix(x), sh.real([]), sh.string("global"), ...)
    res["raw"], sh.protect_depth(), sh.last_error()
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SHIM_SRC = os.path.join(ROOT, "r_package", "src", "icikt_shim.c")
LIB_DIR = os.path.join(ROOT, "icikendalltau_b200")
OUT = os.path.join(HERE, "_build", "libicikt_rshim_test.so")

NILSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP = 0, 10, 13, 14, 16, 19


def build(force=False):
    srcs = [SHIM_SRC, os.path.join(HERE, "r_stub.c")]
    deps = srcs + [os.path.join(HERE, h) for h in ("Rinternals.h", "R.h", os.path.join("R_ext", "Rdynload.h"))] + \
        [os.path.join(ROOT, "include", "icikt_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # the same flags r_package/src/Makevars hands to R CMD SHLIB, with the stub headers standing in for R's
    cmd = ["gcc", "-std=gnu11", "-O1", "-g", "-Wall", "-Wextra", "-Werror",
           "-Wno-cast-function-type",  # (DL_FUNC)&fn is R's own registration idiom (src/RcppExports.cpp:114-121)
           "-fPIC", "-shared", "-I", HERE,
           "-I", os.path.join(ROOT, "include")] + srcs + \
          ["-L", LIB_DIR, "-licikt_b200", f"-Wl,-rpath,{LIB_DIR}", "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


class Shim:
    def __init__(self, path):
        L = self.L = ctypes.CDLL(path)
        vp = ctypes.c_void_p
        for name, res, args in (
                ("rstub_load_package", ctypes.c_int, []), ("rstub_use_dynamic_symbols", ctypes.c_int, []),
                ("rstub_n_routines", ctypes.c_int, []), ("rstub_routine_name", ctypes.c_char_p, [ctypes.c_int]),
                ("rstub_routine_nargs", ctypes.c_int, [ctypes.c_int]),
                ("rstub_dot_call", vp, [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp)]),
                ("rstub_last_error", ctypes.c_char_p, []), ("rstub_protect_depth", ctypes.c_int, []),
                ("rstub_protect_max", ctypes.c_int, []), ("rstub_protect_underflow", ctypes.c_int, []),
                ("rstub_reset", None, []), ("rstub_nil", vp, []),
                ("rstub_real", vp, [vp, ctypes.c_ssize_t]), ("rstub_real_matrix", vp, [vp, ctypes.c_int, ctypes.c_int]),
                ("rstub_int", vp, [vp, ctypes.c_ssize_t]), ("rstub_int_matrix", vp, [vp, ctypes.c_int, ctypes.c_int]),
                ("rstub_logical", vp, [ctypes.c_int]), ("rstub_string", vp, [ctypes.c_char_p]),
                ("rstub_type", ctypes.c_int, [vp]), ("rstub_length", ctypes.c_ssize_t, [vp]),
                ("rstub_nrow", ctypes.c_int, [vp]), ("rstub_ncol", ctypes.c_int, [vp]), ("rstub_data", vp, [vp]),
                ("rstub_list_get", vp, [vp, ctypes.c_char_p]), ("rstub_list_name", ctypes.c_char_p, [vp, ctypes.c_ssize_t]),
                ("rstub_na_real", ctypes.c_double, []), ("R_IsNA", ctypes.c_int, [ctypes.c_double])):
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        assert L.rstub_load_package() == 1, "R_init_ICIKendallTauB200 registered no routines"

    # ---- the registration table (src/RcppExports.cpp:113-128 discipline)
    def routines(self):
        return {self.L.rstub_routine_name(k).decode(): self.L.rstub_routine_nargs(k)
                for k in range(self.L.rstub_n_routines())}

    def use_dynamic_symbols(self):
        return bool(self.L.rstub_use_dynamic_symbols())

    # ---- argument constructors (copies, like R vectors)
    def nil(self):
        return self.L.rstub_nil()

    def real(self, v):
        a = np.ascontiguousarray(v, dtype=np.float64).ravel()
        return self.L.rstub_real(a.ctypes.data, a.size)

    def real_matrix(self, m):
        a = np.asfortranarray(m, dtype=np.float64)
        return self.L.rstub_real_matrix(a.ctypes.data, a.shape[0], a.shape[1])

    def integer(self, v):
        a = np.ascontiguousarray(v, dtype=np.int32).ravel()
        return self.L.rstub_int(a.ctypes.data, a.size)

    def int_matrix(self, m):
        a = np.asfortranarray(m, dtype=np.int32)
        return self.L.rstub_int_matrix(a.ctypes.data, a.shape[0], a.shape[1])

    def logical(self, v):
        return self.L.rstub_logical(int(bool(v)))

    def string(self, s):
        return self.L.rstub_string(s.encode())

    # ---- .Call
    def dot_call(self, name, *args):
        """Returns the result converted to Python (named list -> dict of NumPy arrays), or raises
        RuntimeError with R's error message."""
        arr = (ctypes.c_void_p * max(len(args), 1))(*args)
        res = self.L.rstub_dot_call(name.encode(), len(args), arr)
        if not res:
            raise RuntimeError(self.L.rstub_last_error().decode())
        return self.to_python(res)

    def to_python(self, s):
        t, n = self.L.rstub_type(s), self.L.rstub_length(s)
        if t == NILSXP:
            return None
        if t in (REALSXP, INTSXP, LGLSXP):
            ct = ctypes.c_double if t == REALSXP else ctypes.c_int32
            a = np.ctypeslib.as_array(ctypes.cast(self.L.rstub_data(s), ctypes.POINTER(ct)), shape=(max(n, 1),))[:n].copy()
            nc = self.L.rstub_ncol(s)
            return a.reshape((self.L.rstub_nrow(s), nc), order="F") if nc >= 0 else a
        if t == VECSXP:
            ptrs = ctypes.cast(self.L.rstub_data(s), ctypes.POINTER(ctypes.c_void_p))
            return {self.L.rstub_list_name(s, k).decode(): self.to_python(ptrs[k]) for k in range(n)}
        raise TypeError(f"unsupported SEXP type {t}")

    def protect_depth(self):
        return self.L.rstub_protect_depth()

    def protect_max(self):
        return self.L.rstub_protect_max()

    def protect_underflow(self):
        return bool(self.L.rstub_protect_underflow())

    def reset(self):
        self.L.rstub_reset()

    def is_na(self, a):
        """R's is.na() restricted to NA_real_ proper (payload 1954), element-wise."""
        a = np.asarray(a, dtype=np.float64)
        return np.isnan(a) & ((a.view(np.uint64) & np.uint64(0xFFFFFFFF)) == np.uint64(1954))


_shim = None


def load():
    global _shim
    if _shim is None:
        _shim = Shim(build())
    return _shim
