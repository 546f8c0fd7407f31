"""ctypes binding of the C ABI declared in ``include/abn_b200.h``.

The shared library ``abn_b200/csrc/libabn_b200.so`` is the product; there is no CPU
fallback.  ``lib()`` raises if it is missing instead of degrading to something else.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ABN_B200_LIB lets tuning sweeps point at a differently built copy of the same library
LIB_PATH = os.environ.get("ABN_B200_LIB") or os.path.join(_HERE, "csrc", "libabn_b200.so")

BINOMIAL, GAUSSIAN, POISSON, MULTINOMIAL = 1, 2, 3, 4
METHOD_MLE, METHOD_BAYES = 0, 1
SOLVER_NEWTON, SOLVER_HYBRID = 0, 1
FLAG_EXACT_PATH = 1  # abn_control.flags: follow the reference's IRLS pass for pass (include/abn_b200.h)
FLAG_RESULT_ON_DEVICE = 4  # abn_b200_score: the result pointers are device pointers (include/abn_b200.h)
FLAG_NO_BIAS_REDUCTION = 2  # binomial mle without the bias reduction (irls_binomial_cpp_fast, fitAbn's first choice)
FIT_OK, FIT_NOT_CONVERGED, FIT_NON_FINITE = 0, 1, 2
MAX_K = 32   # widest design
FAST_K = 8   # widest design of the register-resident kernels
MAX_KERNELS = 64

FAMILY_CODE = {"binomial": BINOMIAL, "gaussian": GAUSSIAN, "poisson": POISSON, "multinomial": MULTINOMIAL}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class Dataset(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("n_nodes", C.c_int32),
        ("n_cols", C.c_int32),
        ("X", c_double_p),
        ("node_family", c_int32_p),
        ("node_col0", c_int32_p),
    ]


class Candidates(C.Structure):
    _fields_ = [
        ("rows", C.c_int64),
        ("max_parents", C.c_int32),
        ("child", c_int32_p),
        ("parents", c_int32_p),
    ]


class Control(C.Structure):
    _fields_ = [
        ("method", C.c_int32),
        ("max_iters", C.c_int32),
        ("tol", C.c_double),
        ("epsabs", C.c_double),
        ("prior_mean", C.c_double),
        ("prior_sd", C.c_double),
        ("gam_shape", C.c_double),
        ("gam_scale", C.c_double),
        ("solver", C.c_int32),
        ("device", C.c_int32),
        ("nvars", C.c_int32),
        ("flags", C.c_int32),
        ("ngpus", C.c_int32),
        ("block", C.c_int32),
        ("devices", C.c_int32 * 16),
    ]


class Result(C.Structure):
    _fields_ = [
        ("mlik", c_double_p),
        ("aic", c_double_p),
        ("bic", c_double_p),
        ("mdl", c_double_p),
        ("status", c_int32_p),
        ("iters", c_int32_p),
        ("coef", c_double_p),
        ("coef_stride", C.c_int32),
        ("reserved", C.c_int32),
        ("se", c_double_p),
        ("sse", c_double_p),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_double),
        ("plan_ms", C.c_double),
        ("fit_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("launches", C.c_int64),
        ("algo_bytes", C.c_int64),
        ("n_kernels", C.c_int32),
        ("reserved", C.c_int32),
        ("kernel_id", C.c_int32 * MAX_KERNELS),
        ("kernel_fits", C.c_int64 * MAX_KERNELS),
        ("kernel_bytes", C.c_int64 * MAX_KERNELS),
        ("kernel_ms", C.c_double * MAX_KERNELS),
        ("kernel_fp64", C.c_int64 * MAX_KERNELS),
    ]


# symbols include/abn_b200.h declares; tests check every one is exported
EXPORTS = (
    "abn_b200_default_control",
    "abn_b200_version",
    "abn_b200_device_count",
    "abn_b200_enumerate",
    "abn_b200_node_defn",
    "abn_b200_data_create",
    "abn_b200_data_destroy",
    "abn_b200_score",
    "abn_b200_score_cache",
    "abn_b200_score_enumerated",
    "abn_b200_last_timing",
)

_lib = None


class AbnB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"abn_b200 error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Load libabn_b200.so (built by ``__graft_entry__.build()`` / ``make -C abn_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AbnB200Error(
            -3, f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    errargs = [C.c_char_p, C.c_size_t]
    L.abn_b200_default_control.argtypes = [C.c_int32, C.POINTER(Control)]
    L.abn_b200_default_control.restype = None
    L.abn_b200_version.restype = C.c_int32
    L.abn_b200_device_count.restype = C.c_int32
    L.abn_b200_enumerate.argtypes = [C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_int32_p, C.c_int32,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int32), c_int32_p, c_int32_p] + errargs
    L.abn_b200_enumerate.restype = C.c_int
    L.abn_b200_node_defn.argtypes = [C.POINTER(Candidates), C.c_int32, c_int32_p] + errargs
    L.abn_b200_node_defn.restype = C.c_int
    L.abn_b200_data_create.argtypes = [C.POINTER(Dataset), C.c_int32, C.POINTER(C.c_void_p)] + errargs
    L.abn_b200_data_create.restype = C.c_int
    L.abn_b200_data_destroy.argtypes = [C.c_void_p]
    L.abn_b200_data_destroy.restype = None
    L.abn_b200_score.argtypes = [C.c_void_p, C.POINTER(Candidates), C.POINTER(Control), C.POINTER(Result)] + errargs
    L.abn_b200_score.restype = C.c_int
    L.abn_b200_score_cache.argtypes = [C.POINTER(Dataset), C.POINTER(Candidates), C.POINTER(Control),
                                       C.POINTER(Result)] + errargs
    L.abn_b200_score_cache.restype = C.c_int
    L.abn_b200_score_enumerated.argtypes = [C.c_void_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p, C.c_int32,
                                            C.POINTER(Control), C.POINTER(Result), C.c_int64, C.c_int32,
                                            c_int32_p, c_int32_p] + errargs
    L.abn_b200_score_enumerated.restype = C.c_int
    L.abn_b200_last_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    L.abn_b200_last_timing.restype = C.c_int
    _lib = L
    return L


# ---- numpy -> struct helpers (the returned tuple keeps the arrays alive) -------------------------

def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr_d(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def ptr_i(a):
    return a.ctypes.data_as(c_int32_p) if a is not None else None


def make_dataset(X: np.ndarray, node_family, node_col0):
    """X: (n, n_cols) array in Fortran (column-major) order, float64."""
    Xf = np.asfortranarray(X, dtype=np.float64)
    fam = _i32(node_family)
    col0 = _i32(node_col0)
    ds = Dataset(Xf.shape[0], len(fam), Xf.shape[1], ptr_d(Xf), ptr_i(fam), ptr_i(col0))
    return ds, (Xf, fam, col0)


def make_candidates(child, parents):
    ch = _i32(child)
    pa = _i32(parents).reshape(len(ch), -1) if len(ch) else _i32(parents).reshape(0, max(1, np.size(parents)))
    pa = np.ascontiguousarray(pa)
    cd = Candidates(len(ch), pa.shape[1], ptr_i(ch), ptr_i(pa))
    return cd, (ch, pa)


def make_result(rows: int, method: int, coef_stride: int = 0, se: bool = False):
    """Result struct over fresh numpy arrays; coef_stride > 0 also asks for the coefficients
    (rows x coef_stride, layout in include/abn_b200.h)."""
    arrs = {
        "mlik": np.full(rows, np.nan),
        "aic": np.full(rows, np.nan),
        "bic": np.full(rows, np.nan),
        "mdl": np.full(rows, np.nan),
        "status": np.zeros(rows, dtype=np.int32),
        "iters": np.zeros(rows, dtype=np.int32),
    }
    res = Result(ptr_d(arrs["mlik"]), ptr_d(arrs["aic"]), ptr_d(arrs["bic"]), ptr_d(arrs["mdl"]),
                 ptr_i(arrs["status"]), ptr_i(arrs["iters"]))
    if coef_stride > 0:
        arrs["coef"] = np.full((rows, coef_stride), np.nan)
        res.coef = ptr_d(arrs["coef"])
        res.coef_stride = coef_stride
        if se:
            arrs["se"] = np.full((rows, coef_stride), np.nan)
            res.se = ptr_d(arrs["se"])
            arrs["sse"] = np.full(rows, np.nan)
            res.sse = ptr_d(arrs["sse"])
    return res, arrs


def check(code: int, errbuf) -> None:
    if code != 0:
        raise AbnB200Error(code, errbuf.value.decode("utf-8", "replace"))
