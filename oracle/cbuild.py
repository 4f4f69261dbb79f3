"""Build / load the C restatement (oracle/cost_volume.c).  ORACLE: test infrastructure only."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "cost_volume.c")
_OUT = os.path.join(_HERE, "_build", "liboracle_cv.so")


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        # -ffp-contract=off: every product/sum is rounded separately unless the
        # source calls fmaf() explicitly (fma_mode=1).
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", _OUT, _SRC, "-lm"])
    return _OUT


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i64, p = ctypes.c_int64, ctypes.c_void_p
        for suf in ("f32", "f64"):
            f = getattr(_lib, f"oracle_cost_volume_fwd_{suf}")
            f.argtypes = [p, p, p, p, i64, i64, i64, i64, i64, i64, ctypes.c_int]
            f.restype = None
            b = getattr(_lib, f"oracle_cost_volume_bwd_{suf}")
            b.argtypes = [p, p, p, p, i64, i64, i64, i64, i64, i64]
            b.restype = None
        _lib.oracle_cost_volume_xlow_f32.argtypes = [p, p, i64, i64, i64, i64]
        _lib.oracle_cost_volume_xlow_f32.restype = None
    return _lib
