"""ctypes wrapper of the CPU oracle (``oracle/libabn_oracle.so``).

TEST INFRASTRUCTURE: importable only from ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py`` (cpu_baseline / ``--impl reference``).  Nothing under ``abn_b200/`` imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from abn_b200 import _capi as capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libabn_oracle.so")
OR_MAXCOEF = 40


class Fit(C.Structure):
    _fields_ = [
        ("loglik", C.c_double),
        ("aic", C.c_double),
        ("bic", C.c_double),
        ("iters", C.c_int),
        ("passes", C.c_int),
        ("status", C.c_int),
        ("rank", C.c_int),
        ("coef", C.c_double * OR_MAXCOEF),
        ("se", C.c_double * OR_MAXCOEF),
    ]


_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "abn_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.or_enumerate_bayes.argtypes = [C.c_int, ip, ip, ip, ip, C.c_int, ip, ip]
    L.or_enumerate_bayes.restype = C.c_int64
    L.or_enumerate_mle.argtypes = [C.c_int, ip, ip, C.c_int, ip, ip]
    L.or_enumerate_mle.restype = C.c_int64
    fp = C.POINTER(Fit)
    L.or_mle_gaussian.argtypes = [dp, dp, C.c_long, C.c_int, fp]
    L.or_mle_binomial_br.argtypes = [dp, dp, C.c_long, C.c_int, C.c_int, C.c_double, fp]
    L.or_mle_binomial.argtypes = [dp, dp, C.c_long, C.c_int, C.c_int, C.c_double, fp]
    L.or_mle_poisson.argtypes = [dp, dp, C.c_long, C.c_int, C.c_int, C.c_double, fp]
    L.or_mle_multinomial.argtypes = [dp, ip, C.c_long, C.c_int, C.c_int, C.c_int, fp]
    L.or_rank_cpp.argtypes = [dp, C.c_long, C.c_int]
    L.or_rank_cpp.restype = C.c_int
    L.or_bayes_binomial.argtypes = [dp, dp, C.c_long, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, fp]
    L.or_bayes_poisson.argtypes = [dp, dp, C.c_long, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, fp]
    L.or_bayes_gaussian.argtypes = [dp, dp, C.c_long, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_int, C.c_double, C.c_int, fp]
    L.or_score_cache.argtypes = [C.POINTER(capi.Dataset), C.POINTER(capi.Candidates), C.POINTER(capi.Control),
                                 C.POINTER(capi.Result), C.c_int]
    L.or_score_cache.restype = C.c_int
    L.or_num_threads.restype = C.c_int
    L.or_set_sum_order.argtypes = [C.c_int]
    L.or_set_sum_order.restype = None
    _lib = L
    return L


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def enumerate_bayes(d, banned, retained, max_parents, which_nodes=None):
    """banned / retained: (d, d) arrays indexed [child, parent]. Returns (children 0-based, node_defn rows x d)."""
    L = lib()
    ban = np.asfortranarray(banned, dtype=np.int32).ravel(order="F").astype(np.intc)
    ret = np.asfortranarray(retained, dtype=np.int32).ravel(order="F").astype(np.intc)
    mp = np.ascontiguousarray(np.broadcast_to(np.asarray(max_parents), (d,)), dtype=np.intc)
    wn = np.arange(d, dtype=np.intc) if which_nodes is None else np.ascontiguousarray(which_nodes, dtype=np.intc)
    rows = L.or_enumerate_bayes(d, _ip(ban), _ip(ret), _ip(mp), _ip(wn), len(wn), None, None)
    ch = np.zeros(rows, dtype=np.intc)
    nd = np.zeros((rows, d), dtype=np.intc)
    L.or_enumerate_bayes(d, _ip(ban), _ip(ret), _ip(mp), _ip(wn), len(wn), _ip(ch), _ip(nd))
    return ch, nd


def enumerate_mle(d, banned, retained, max_parents):
    L = lib()
    ban = np.asfortranarray(banned, dtype=np.int32).ravel(order="F").astype(np.intc)
    ret = np.asfortranarray(retained, dtype=np.int32).ravel(order="F").astype(np.intc)
    rows = L.or_enumerate_mle(d, _ip(ban), _ip(ret), int(max_parents), None, None)
    ch = np.zeros(rows, dtype=np.intc)
    nd = np.zeros((rows, d), dtype=np.intc)
    L.or_enumerate_mle(d, _ip(ban), _ip(ret), int(max_parents), _ip(ch), _ip(nd))
    return ch, nd


def _xy(X, y):
    Xf = np.asfortranarray(X, dtype=np.float64)
    yy = np.ascontiguousarray(y, dtype=np.float64)
    return Xf, yy


def _fit_dict(f: Fit, ncoef: int):
    return {"loglik": f.loglik, "aic": f.aic, "bic": f.bic, "iters": f.iters, "passes": f.passes,
            "status": f.status, "coef": np.array(f.coef[:ncoef]), "se": np.array(f.se[:ncoef])}


def mle_gaussian(X, y):
    Xf, yy = _xy(X, y)
    f = Fit()
    lib().or_mle_gaussian(_dp(Xf), _dp(yy), Xf.shape[0], Xf.shape[1], C.byref(f))
    return _fit_dict(f, Xf.shape[1])


def mle_binomial_br(X, y, maxit=100, tol=1e-8):
    Xf, yy = _xy(X, y)
    f = Fit()
    lib().or_mle_binomial_br(_dp(Xf), _dp(yy), Xf.shape[0], Xf.shape[1], maxit, tol, C.byref(f))
    return _fit_dict(f, Xf.shape[1])


def mle_binomial(X, y, maxit=100, tol=1e-8):
    """irls_binomial_cpp_fast: no bias reduction (fitAbn's first choice)."""
    Xf, yy = _xy(X, y)
    f = Fit()
    lib().or_mle_binomial(_dp(Xf), _dp(yy), Xf.shape[0], Xf.shape[1], maxit, tol, C.byref(f))
    return _fit_dict(f, Xf.shape[1])


def mle_poisson(X, y, maxit=100, tol=1e-8):
    Xf, yy = _xy(X, y)
    f = Fit()
    lib().or_mle_poisson(_dp(Xf), _dp(yy), Xf.shape[0], Xf.shape[1], maxit, tol, C.byref(f))
    return _fit_dict(f, Xf.shape[1])


def mle_multinomial(X, ycode, L, R=None):
    Xf = np.asfortranarray(X, dtype=np.float64)
    yc = np.ascontiguousarray(ycode, dtype=np.intc)
    if R is None:
        R = rank(Xf)
    f = Fit()
    lib().or_mle_multinomial(_dp(Xf), _ip(yc), Xf.shape[0], Xf.shape[1], L, R, C.byref(f))
    return _fit_dict(f, min(OR_MAXCOEF, (L - 1) * Xf.shape[1]))


def rank(X):
    Xf = np.asfortranarray(X, dtype=np.float64)
    return lib().or_rank_cpp(_dp(Xf), Xf.shape[0], Xf.shape[1])


def bayes_fit(family, X, y, prior_mean=0.0, prior_sd=1 / np.sqrt(0.001), gam_shape=1.0, gam_scale=1 / 5e-5,
              max_iters=100, epsabs=1e-7, solver=capi.SOLVER_NEWTON):
    Xf, yy = _xy(X, y)
    f = Fit()
    n, m = Xf.shape
    L = lib()
    if family == capi.BINOMIAL:
        L.or_bayes_binomial(_dp(Xf), _dp(yy), n, m, prior_mean, prior_sd, max_iters, epsabs, solver, C.byref(f))
    elif family == capi.POISSON:
        L.or_bayes_poisson(_dp(Xf), _dp(yy), n, m, prior_mean, prior_sd, max_iters, epsabs, solver, C.byref(f))
    else:
        L.or_bayes_gaussian(_dp(Xf), _dp(yy), n, m, prior_mean, prior_sd, gam_shape, gam_scale, max_iters, epsabs,
                            solver, C.byref(f))
    return _fit_dict(f, m + (1 if family == capi.GAUSSIAN else 0))


def score_cache(X, node_family, node_col0, child, parents, control: capi.Control, threads: int = 0):
    """Same inputs as abn_b200_score_cache. Returns dict of numpy arrays."""
    ds, keep1 = capi.make_dataset(X, node_family, node_col0)
    cd, keep2 = capi.make_candidates(child, parents)
    res, arrs = capi.make_result(cd.rows, control.method)
    rc = lib().or_score_cache(C.byref(ds), C.byref(cd), C.byref(control), C.byref(res), threads)
    if rc != 0:
        raise RuntimeError(f"or_score_cache failed: {rc}")
    del keep1, keep2
    return arrs


def set_sum_order(order: int) -> None:
    """Test knob: row order of the Gram sums (0 ascending, 1 descending) - the reference leaves it to its BLAS."""
    lib().or_set_sum_order(int(order))


def num_threads() -> int:
    return lib().or_num_threads()
