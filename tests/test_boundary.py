"""The C-ABI library: loads without a GPU, exports every declared symbol, enumerates bit-exactly, and
fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

from abn_b200 import _capi as capi
from abn_b200 import cache as abncache
from abn_b200 import prep, synth
from oracle import pyoracle

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "abn_b200.h")).read()
    declared = set(re.findall(r"\b(abn_b200_[a-z_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS)
    lib = capi.lib()
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.abn_b200_version() == 200


def test_default_control_matches_build_control():
    c = capi.Control()
    capi.lib().abn_b200_default_control(capi.METHOD_BAYES, C.byref(c))
    ref = abncache.to_c_control(abncache.build_control("bayes"), "bayes", 0)
    for f, _ in capi.Control._fields_:
        a, b = getattr(c, f), getattr(ref, f)
        assert (list(a) == list(b)) if f == "devices" else (a == b), f
    assert abs(c.prior_sd - 1 / np.sqrt(0.001)) < 1e-12 and c.gam_scale == 20000.0 and c.max_iters == 100


def test_no_gpu_is_an_error_not_a_fallback():
    if H.has_gpu():
        pytest.skip("a GPU is present")
    names, data, dists = synth.make_columns(50, ["gaussian", "binomial"], seed=1)
    with pytest.raises(capi.AbnB200Error) as e:
        abncache.buildScoreCache(data, dists, method="mle", max_parents=1)
    assert e.value.code == -3


def _enum_cases():
    rng = np.random.default_rng(0)
    d = 7
    zero = np.zeros((d, d), int)
    ban = (rng.random((d, d)) < 0.15).astype(int)
    np.fill_diagonal(ban, 0)
    ret = np.zeros((d, d), int)
    ret[2, 5] = 1
    ret[4, 0] = 1
    ban[2, 5] = 0
    ban[4, 0] = 0
    return [
        (d, zero, zero, 3, None),
        (d, ban, zero, 4, None),
        (d, ban, ret, 3, None),
        (d, zero, ret, np.array([0, 1, 2, 3, 4, 5, 6]), None),     # per-node limits (bayes only)
        (d, ban, ret, 6, None),
        (d, ban, ret, 2, np.array([4, 1, 2], dtype=np.int32)),     # which.nodes
        (1, np.zeros((1, 1), int), np.zeros((1, 1), int), 0, None),
    ]


@pytest.mark.parametrize("case", range(7))
def test_enumerate_bit_exact_with_reference_algorithm(case):
    d, ban, ret, mp, wn = _enum_cases()[case]
    ch, par = abncache.enumerate_candidates(d, ban, ret, mp, wn)
    ch_o, nd_o = pyoracle.enumerate_bayes(d, ban, ret, mp, wn)
    assert np.array_equal(ch, ch_o)
    nd = abncache.node_defn(ch, par, d)
    assert np.array_equal(nd, nd_o)
    assert (np.diff(par, axis=1)[(par[:, 1:] >= 0)] > 0).all() if par.shape[1] > 1 else True
    if np.ndim(mp) == 0 and wn is None and d > 1:
        ch_m, nd_m = pyoracle.enumerate_mle(d, ban, ret, mp)
        assert np.array_equal(ch, ch_m) and np.array_equal(nd, nd_m)


def test_enumerate_ex1_golden_and_counts():
    z = H.load_npz("ex1_bayes.npz")
    ch, par = abncache.enumerate_candidates(10, np.zeros((10, 10), int), np.zeros((10, 10), int), 4)
    assert np.array_equal(ch + 1, z["children"]) and np.array_equal(par, z["parents"])
    nd = abncache.node_defn(ch, par, 10)
    assert hashlib.sha1(np.ascontiguousarray(nd.astype(np.int32)).tobytes()).hexdigest() == str(z["node_defn_sha1"])
    # row counts of BASELINE.md section 4
    from math import comb
    for d, mp, per_node in ((10, 3, 130), (33, 3, 5489), (50, 4, 231526), (100, 3, 161800), (40, 4, 92171)):
        assert sum(comb(d - 1, i) for i in range(mp + 1)) == per_node
    ch, par = abncache.enumerate_candidates(33, np.zeros((33, 33), int), np.zeros((33, 33), int), 3)
    assert len(ch) == 181137


def test_enumerate_rejects_bad_arguments():
    with pytest.raises(capi.AbnB200Error):
        abncache.enumerate_candidates(4, np.zeros((4, 4), int), np.zeros((4, 4), int), 2, np.array([7], np.int32))
    with pytest.raises(capi.AbnB200Error):
        abncache.enumerate_candidates(4, np.zeros((4, 4), int), np.zeros((4, 4), int), -1)


def test_one_hot_expansion_matches_reference_fixture():
    """R/build_score_cache_mle.R:568-599 on tests/testthat/testdata/forLoopContent_data.Rdata."""
    z = H.load_npz("forloop_onehot.npz")
    names, data = H.df_from_npz(z, "lvl_")
    mnames, mdata = H.df_from_npz(z, "multi_")
    dists = H.dists_from_npz(z)
    P = prep.prepare(data, dists, "mle")
    assert P.col_names == mnames
    # Outdoor and the Sex one-hot block are not touched by centring
    for j, nm in enumerate(mnames[:5]):
        assert np.array_equal(P.X[:, j], np.asarray(mdata[nm], dtype=float)), nm
    assert list(P.node_col0) == [0, 1, 5, 6, 7]
    nd, ndm = z["node_defn"].astype(int), z["node_defn_multi"].astype(int)
    rep = np.repeat(np.arange(4), np.diff(P.node_col0))
    assert np.array_equal(nd[:, rep], ndm)
