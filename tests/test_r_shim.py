"""integration/abn_cuda_shim.c - the .Call shim a maintainer adds to the reference's src/  -
compiled against the R API stand-in of tests/r_stub/ and called the way .Call calls it: R-style 1-based integer
vectors in, a named list of fresh vectors out, error() for failures. Compared bit for bit with the ctypes path
over the same C ABI. What it must fit into: the R_CallMethodDef table of src/RcppExports.cpp:122-143."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from abn_b200 import _capi as capi
from abn_b200 import cache as abncache
from abn_b200 import prep, synth

import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
INTSXP, REALSXP, VECSXP = 13, 14, 19


class Sexp(C.Structure):
    _fields_ = [("type", C.c_int), ("length", C.c_ssize_t), ("nrow", C.c_int), ("ncol", C.c_int),
                ("data", C.c_void_p), ("names", C.POINTER(C.c_char_p))]


SEXP = C.POINTER(Sexp)


@pytest.fixture(scope="module")
def shim():
    subprocess.run(["make", "-s", "-C", HERE, "libabn_shim_test.so"], check=True, capture_output=True)
    capi.lib()  # libabn_b200.so first: the shim library links against it
    L = C.CDLL(os.path.join(HERE, "libabn_shim_test.so"))
    L.stub_int_vector.argtypes = [C.POINTER(C.c_int), C.c_ssize_t]
    L.stub_int_vector.restype = SEXP
    L.stub_real_vector.argtypes = [C.POINTER(C.c_double), C.c_ssize_t]
    L.stub_real_vector.restype = SEXP
    L.stub_int_matrix.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int]
    L.stub_int_matrix.restype = SEXP
    L.stub_real_matrix_view.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int]
    L.stub_real_matrix_view.restype = SEXP
    L.stub_call.argtypes = [C.c_void_p, C.c_int, C.POINTER(SEXP)]
    L.stub_call.restype = SEXP
    L.stub_last_error.restype = C.c_char_p
    L.stub_protect_depth.restype = C.c_int
    L.stub_name.argtypes = [SEXP, C.c_int]
    L.stub_name.restype = C.c_char_p
    yield L
    L.stub_free_all()


def _ivec(L, a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return L.stub_int_vector(a.ctypes.data_as(C.POINTER(C.c_int)), a.size)


def _imat(L, a):  # a: (nrow, ncol) -> column-major like R
    a = np.asfortranarray(a, dtype=np.int32)
    return L.stub_int_matrix(a.ctypes.data_as(C.POINTER(C.c_int)), a.shape[0], a.shape[1])


def _call(L, name, args):
    fn = C.cast(getattr(L, name), C.c_void_p)
    arr = (SEXP * len(args))(*args)
    out = L.stub_call(fn, len(args), arr)
    assert L.stub_protect_depth() == 0          # PROTECT / UNPROTECT balanced on every path
    return out


def _list(L, sx):
    s = sx.contents
    assert s.type == VECSXP
    elts = C.cast(s.data, C.POINTER(SEXP))
    out = {}
    for i in range(s.length):
        e = elts[i].contents
        buf = C.cast(e.data, C.POINTER(C.c_int if e.type == INTSXP else C.c_double))
        arr = np.ctypeslib.as_array(buf, shape=(e.length,)).copy()
        if e.nrow:
            arr = arr.reshape((e.ncol, e.nrow)).T  # column-major
        name = s.names[i].decode() if s.names else ""
        out[name or i] = arr
    return out


def _enumerate_via_shim(L, d, banned, retained, mp, which1):
    out = _call(L, "abn_cuda_enumerate", [_ivec(L, [d]), _imat(L, banned), _imat(L, retained), _ivec(L, mp), _ivec(L, which1)])
    assert out, L.stub_last_error()
    lst = _list(L, out)
    return lst[0], lst[1]          # children (1-based), parents (width x rows, 1-based, 0 = none)


def test_enumerate_through_the_shim_is_the_ctypes_enumeration(shim):
    """No GPU needed: abn_b200_enumerate is host code. Same rows as the ctypes path, R's 1-based convention."""
    d = 7
    rng = np.random.default_rng(5)
    banned = (rng.random((d, d)) < 0.15).astype(np.int32)
    np.fill_diagonal(banned, 0)
    retained = np.zeros((d, d), np.int32)
    retained[2, 4] = 1
    banned[2, 4] = 0
    mp = np.array([3, 2, 3, 1, 0, 3, 2], np.int32)
    ch1, pa = _enumerate_via_shim(shim, d, banned, retained, mp, np.arange(1, d + 1))
    ch0, par = abncache.enumerate_candidates(d, banned, retained, mp)
    assert np.array_equal(ch1, ch0 + 1)
    assert np.array_equal(pa.T, par + 1)
    # a subset of nodes, as buildScoreCache(which.nodes = ...) passes it
    ch1s, pas = _enumerate_via_shim(shim, d, banned, retained, mp, np.array([2, 6]))
    ch0s, pars = abncache.enumerate_candidates(d, banned, retained, mp, np.array([1, 5], np.int32))
    assert np.array_equal(ch1s, ch0s + 1) and np.array_equal(pas.T[:, :pars.shape[1]], pars + 1)


def test_shim_reports_library_errors_through_error(shim):
    d = 3
    zero = np.zeros((d, d), np.int32)
    out = _call(shim, "abn_cuda_enumerate", [_ivec(shim, [d]), _imat(shim, zero), _imat(shim, zero),
                                              _ivec(shim, [1, 1, 1]), _ivec(shim, [9])])   # which.nodes out of range
    assert not out and shim.stub_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["mle", "bayes"])
def test_score_cache_through_the_shim_is_bit_identical_to_ctypes(shim, method):
    L = shim
    names, data, dists, mp, _ = synth.make_config(4, n=5000, n_nodes=8)
    d = len(names)
    P = prep.prepare(data, dists, method)
    zero = np.zeros((d, d), np.int32)
    ch1, pa = _enumerate_via_shim(L, d, zero, zero, np.full(d, 3, np.int32), np.arange(1, d + 1))
    X = np.asfortranarray(P.X, dtype=np.float64)
    R_X = L.stub_real_matrix_view(X.ctypes.data_as(C.POINTER(C.c_double)), X.shape[0], X.shape[1])
    b = abncache.build_control(method)
    ctrl = np.array([b["max.iters"], b.get("tol", 1e-8), b.get("epsabs", 1e-7), b.get("mean", 0.0),
                     1 / np.sqrt(b.get("prec", 0.001)), b.get("loggam.shape", 1.0), 1 / b.get("loggam.inv.scale", 5e-5)])
    args = [R_X, _ivec(L, P.node_family), _ivec(L, P.node_col0), _ivec(L, ch1), _imat(L, pa),
            _ivec(L, [0 if method == "mle" else 1]), L.stub_real_vector(ctrl.ctypes.data_as(C.POINTER(C.c_double)), len(ctrl)),
            _ivec(L, [d]), _ivec(L, [0])]
    out = _call(L, "abn_cuda_score_cache", args)
    assert out, L.stub_last_error()
    got = _list(L, out)
    assert list(got) == ["mlik", "aic", "bic", "mdl", "error.code", "passes"]
    ref = H.gpu_score(P, ch1 - 1, pa.T - 1, H.control(method, d))
    for k in (("mlik", "aic", "bic", "mdl") if method == "mle" else ("mlik",)):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    assert np.array_equal(got["error.code"], ref["status"]) and np.array_equal(got["passes"], ref["iters"])
    # a bad device ordinal comes back as an R error, not a crash
    args[-1] = _ivec(L, [99])
    assert not _call(L, "abn_cuda_score_cache", args) and b"device" in L.stub_last_error()
