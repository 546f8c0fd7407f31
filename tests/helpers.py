"""Shared test helpers: golden fixture loading, oracle / host-sim / C-ABI scoring wrappers.

TEST INFRASTRUCTURE. The oracle (oracle/) and the host simulation (tests/host_sim.cpp) are used
here as checkers only.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from abn_b200 import _capi as capi
from abn_b200 import cache as abncache
from abn_b200 import prep

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def df_from_npz(z, prefix=""):
    names = [str(x) for x in z[f"{prefix}names"]]
    kinds = [str(x) for x in z[f"{prefix}kinds"]]
    data = {}
    for nm, kind in zip(names, kinds):
        v = z[f"{prefix}v_{nm}"]
        if kind == "factor":
            data[nm] = (v.astype(np.int32), [str(x) for x in z[f"{prefix}lv_{nm}"]])
        else:
            data[nm] = v.astype(np.float64) if kind == "double" else v.astype(np.int64)
    return names, data


def dists_from_npz(z):
    return dict(str(x).split("=", 1) for x in z["dists"])


def parents_from_node_defn(nd, width=None):
    width = width or max(1, int(nd.sum(axis=1).max()))
    par = np.full((nd.shape[0], width), -1, dtype=np.int32)
    for r in range(nd.shape[0]):
        idx = np.nonzero(nd[r])[0]
        par[r, :len(idx)] = idx
    return par


def control(method, d, **kw):
    ctrl = abncache.build_control(method)
    ctrl.update(kw)
    return abncache.to_c_control(ctrl, method, d)


_sim = {}


def sim_lib(wide=False):
    key = "wide" if wide else "narrow"
    if key not in _sim:
        path = os.path.join(HERE, "libhost_sim_wide.so" if wide else "libhost_sim.so")
        subprocess.run(["make", "-s", "-C", HERE], check=True, capture_output=True)
        lib = C.CDLL(path)
        lib.sim_score_cache.argtypes = [C.POINTER(capi.Dataset), C.POINTER(capi.Candidates), C.POINTER(capi.Control),
                                        C.POINTER(capi.Result), C.POINTER(C.c_int)]
        lib.sim_score_cache.restype = C.c_int
        _sim[key] = lib
    return _sim[key]


def sim_score(P: prep.Prepared, children0, parents, ctrl, wide=False, coef_stride=0, se=False):
    """The product's solver state machines driven on the CPU (tests/host_sim.cpp); wide=True uses the
    MAXK = 32 build of the same headers (the limits of abn_b200/csrc/wide.cu)."""
    ds, k1 = capi.make_dataset(P.X, P.node_family, P.node_col0)
    cd, k2 = capi.make_candidates(children0, parents)
    res, arrs = capi.make_result(cd.rows, ctrl.method, coef_stride, se)
    buckets = np.zeros(max(1, cd.rows), dtype=np.intc)
    rc = sim_lib(wide).sim_score_cache(C.byref(ds), C.byref(cd), C.byref(ctrl), C.byref(res),
                                   buckets.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    arrs["bucket"] = buckets[:cd.rows]
    return arrs


def oracle_score(P: prep.Prepared, children0, parents, ctrl, threads=0):
    from oracle import pyoracle
    return pyoracle.score_cache(P.X, P.node_family, P.node_col0, children0, parents, ctrl, threads)


def oracle_order_sensitive(P: prep.Prepared, children0, parents, ctrl, ref=None, tol=1e-9):
    """Rows whose ORACLE result depends on the row order of its Gram sums - the reference leaves that order to
    the BLAS behind Armadillo (oracle/mle.c, or_set_sum_order). They are rounding-noise-driven fits (binomial
    near separation, linear convergence at the iteration cap): the reference does not reproduce its own flag
    and value on them across platforms, so parity is defined on all other rows. Returns (mask, ref, ref_rev)."""
    from oracle import pyoracle
    if ref is None:
        ref = oracle_score(P, children0, parents, ctrl)
    pyoracle.set_sum_order(1)
    try:
        rev = oracle_score(P, children0, parents, ctrl)
    finally:
        pyoracle.set_sum_order(0)
    mask = ref["status"] != rev["status"]
    for k in ("mlik",):
        mask |= rel_err(rev[k], ref[k]) > tol
    return mask, ref, rev


def gpu_score(P: prep.Prepared, children0, parents, ctrl, device=0, coef_stride=0, se=False):
    """Through the C ABI: abn_b200_score_cache (one shot, host buffers in and out)."""
    L = capi.lib()
    ds, k1 = capi.make_dataset(P.X, P.node_family, P.node_col0)
    cd, k2 = capi.make_candidates(children0, parents)
    res, arrs = capi.make_result(cd.rows, ctrl.method, coef_stride, se)
    err = C.create_string_buffer(512)
    ctrl.device = device
    capi.check(L.abn_b200_score_cache(C.byref(ds), C.byref(cd), C.byref(ctrl), C.byref(res), err, len(err)), err)
    return arrs


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.abs(a - b) / np.abs(b)
    r[np.isnan(a) & np.isnan(b)] = 0.0
    r[(a == b)] = 0.0
    return r


def has_gpu():
    try:
        return capi.lib().abn_b200_device_count() > 0
    except Exception:
        return False


def oracle_coef(P: prep.Prepared, child, parents_row, method="mle"):
    """Coefficients of ONE fit from the oracle's single-fit entry points, in the slot layout of abn_result.coef
    (intercept, then the parents' stored columns in candidate order). Binomial / Poisson / Gaussian children,
    designs with an intercept."""
    from oracle import pyoracle
    X = np.asarray(P.X)
    cols = [c for p in parents_row if p >= 0 for c in range(P.node_col0[p], P.node_col0[p + 1])]
    D = np.column_stack([np.ones(X.shape[0])] + [X[:, c] for c in cols])
    y = X[:, P.node_col0[child]]
    fam = P.node_family[child]
    if method == "mle":
        f = {capi.GAUSSIAN: pyoracle.mle_gaussian, capi.BINOMIAL: pyoracle.mle_binomial_br,
             capi.POISSON: pyoracle.mle_poisson}[fam](D, y)
    else:
        f = pyoracle.bayes_fit(fam, D, y, solver=capi.SOLVER_HYBRID)  # the reference's GSL hybrid path
    return f
