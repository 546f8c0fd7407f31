"""Pins the CPU oracle against the reference's own golden caches."""
import numpy as np
import pytest

from abn_b200 import _capi as capi
from abn_b200 import prep
from oracle import pyoracle

import helpers as H


def _case(name):
    z = H.load_npz(name)
    names, data = H.df_from_npz(z)
    return z, names, data, H.dists_from_npz(z)


def test_prep_matches_golden_prepared_frame():
    """mycache$data.df of buildscorecache_ex1.Rdata is the prepped frame (0/1, centred with sd n-1)."""
    z, names, data, dists = _case("ex1_bayes.npz")
    P = prep.prepare(data, dists, "bayes")
    assert np.abs(P.X[:64] - z["prepped_head"]).max() < 1e-15
    assert abs(P.X[0, names.index("g1")] - 1.501633731415002) < 1e-15


def test_enumeration_bit_exact_both_paths():
    z, names, data, dists = _case("ex1_bayes.npz")
    d = len(names)
    zero = np.zeros((d, d), int)
    ch_b, nd_b = pyoracle.enumerate_bayes(d, zero, zero, 4)
    ch_m, nd_m = pyoracle.enumerate_mle(d, zero, zero, 4)
    assert np.array_equal(ch_b + 1, z["children"])
    assert np.array_equal(H.parents_from_node_defn(nd_b, 4), z["parents"])
    # tests/testthat/test-build_score_cache.R:174-184: mle and bayes enumerations are identical
    assert np.array_equal(ch_b, ch_m) and np.array_equal(nd_b, nd_m)


@pytest.mark.parametrize("solver,tol_max,tol_agg", [(capi.SOLVER_HYBRID, 1e-10, 1e-12), (capi.SOLVER_NEWTON, 2e-8, 1e-10)])
def test_ex1_bayes_enforced_golden(solver, tol_max, tol_agg):
    """tests/testthat/test-build_score_cache_bayes.R:126-135 (2 560 rows, max.par = 4)."""
    z, names, data, dists = _case("ex1_bayes.npz")
    P = prep.prepare(data, dists, "bayes")
    ctrl = H.control("bayes", len(names), solver=solver)
    res = H.oracle_score(P, z["children"].astype(np.int32) - 1, z["parents"].astype(np.int32), ctrl)
    rel = H.rel_err(res["mlik"], z["mlik"])
    assert rel.max() < tol_max
    assert np.abs(res["mlik"] - z["mlik"]).mean() / np.abs(z["mlik"]).mean() < tol_agg
    assert (res["status"] == 0).all()


@pytest.mark.parametrize("tag", ["c3", "c4", "c2"])
def test_secondary_goldens(tag):
    """inst/extdata/mycache.c*.rds: values where error.code == 0 and the error.code == 1 flags, which
    only an emulation of GSL's hybridsj/hybridj iterate sequence reproduces."""
    z, names, data, dists = _case(f"{tag}_bayes.npz")
    d = len(names)
    P = prep.prepare(data, dists, "bayes")
    ch, nd = pyoracle.enumerate_bayes(d, z["banned"], z["retained"], z["max_par"])
    assert np.array_equal(ch + 1, z["children"])
    import hashlib
    assert hashlib.sha1(np.ascontiguousarray(nd.astype(np.int32)).tobytes()).hexdigest() == str(z["node_defn_sha1"])
    par = H.parents_from_node_defn(nd)
    res = H.oracle_score(P, ch.astype(np.int32), par, H.control("bayes", d))
    ec = z["error_code"]
    assert (res["status"] == ec).all()           # c2: 363 flagged rows, c3: 7, c4: 0
    ok = ec == 0
    rel = H.rel_err(res["mlik"], z["mlik"])
    if tag == "c2":
        assert (rel[ok] > 1e-8).sum() <= 25 and rel[ok].max() < 1e-5
    else:
        assert rel[ok].max() < 1e-10
    if tag == "c3":
        assert rel[~ok].max() < 1e-12             # the reference's non-modes, reproduced


def test_mle_golden_rows():
    """tests/testthat/testdata/mycache.Rdata: Gaussian rows 5-8 and separated-binomial rows 3,4,10,12
    come from the current code path; mdl = bic + (1 + p) log(3)."""
    z, names, data, dists = _case("mycache_mle.npz")
    d = len(names)
    P = prep.prepare(data, dists, "mle", centre=False)  # the fixture was built on the raw Gaussian column
    ch0 = z["children"].astype(np.int32) - 1
    par = z["parents"].astype(np.int32)
    res = H.oracle_score(P, ch0, par, H.control("mle", d))
    rows = np.array([3, 4, 5, 6, 7, 8, 10, 12]) - 1
    for k in ("mlik", "aic", "bic", "mdl"):
        assert H.rel_err(res[k][rows], z[k][rows]).max() < 1e-12, k
    p = (par >= 0).sum(axis=1)
    assert np.allclose(z["mdl"] - z["bic"], (1 + p) * np.log(3), rtol=0, atol=1e-9)


def test_gaussian_mle_is_ols():
    """tests/testthat/test-build_score_cache_mle.R:61-82: equals logLik / AIC / BIC of lm()."""
    rng = np.random.default_rng(1)
    n = 500
    X = np.column_stack([np.ones(n), rng.normal(size=(n, 3))])
    y = X @ np.array([0.3, 1.0, -2.0, 0.5]) + rng.normal(size=n)
    f = pyoracle.mle_gaussian(X, y)
    beta, *_ = np.linalg.lstsq(X, y, rcond=None)
    rss = float(((y - X @ beta) ** 2).sum())
    ll = -0.5 * n * (np.log(2 * np.pi) + 1 - np.log(n) + np.log(rss))
    assert abs(f["loglik"] - ll) < 1e-9
    assert abs(f["aic"] - (-2 * ll + 2 * 5)) < 1e-9 and abs(f["bic"] - (-2 * ll + np.log(n) * 5)) < 1e-9


def test_binomial_and_poisson_mle_close_to_unpenalised_glm():
    """test-build_score_cache_mle.R:88-123,163-197 compare with glm() at 5e-5: reproduce with a plain
    Newton GLM fit."""
    rng = np.random.default_rng(2)
    n = 4000
    X = np.column_stack([np.ones(n), rng.normal(size=(n, 2))])
    eta = X @ np.array([-0.2, 0.8, -0.5])
    yb = (rng.random(n) < 1 / (1 + np.exp(-eta))).astype(float)
    yp = rng.poisson(np.exp(0.3 + 0.4 * X[:, 1])).astype(float)

    def newton(loglik_grad_hess):
        b = np.zeros(3)
        for _ in range(50):
            ll, g, Hm = loglik_grad_hess(b)
            b = b + np.linalg.solve(Hm, g)
        return loglik_grad_hess(b)[0]

    def logit(b):
        e = X @ b
        p = 1 / (1 + np.exp(-e))
        return (yb * e - np.log1p(np.exp(e))).sum(), X.T @ (yb - p), (X * (p * (1 - p))[:, None]).T @ X

    def pois(b):
        e = X @ b
        mu = np.exp(e)
        from math import lgamma
        return (yp * e - mu).sum() - sum(lgamma(v + 1) for v in yp), X.T @ (yp - mu), (X * mu[:, None]).T @ X

    fb = pyoracle.mle_binomial_br(X, yb)
    fp = pyoracle.mle_poisson(X, yp)
    assert abs(fb["loglik"] - newton(logit)) / abs(fb["loglik"]) < 5e-5
    assert abs(fp["loglik"] - newton(pois)) / abs(fp["loglik"]) < 1e-9


def test_multinomial_restatement_reaches_the_mle():
    rng = np.random.default_rng(3)
    n, L = 3000, 3
    X = np.column_stack([np.ones(n), rng.normal(size=(n, 2))])
    B = np.array([[0, 0, 0], [0.5, 1.0, -0.5], [-0.3, 0.2, 0.8]])
    logits = X @ B.T
    p = np.exp(logits - logits.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    y = np.array([rng.choice(L, p=pi) for pi in p])
    f = pyoracle.mle_multinomial(X, y, L)
    # Newton reference for the softmax MLE
    W = np.zeros((L - 1, 3))
    for _ in range(30):
        lg = np.column_stack([np.zeros(n), X @ W.T])
        pr = np.exp(lg - lg.max(1, keepdims=True))
        pr /= pr.sum(1, keepdims=True)
        g = np.concatenate([X.T @ ((y == l).astype(float) - pr[:, l]) for l in range(1, L)])
        Hm = np.zeros((2 * 3, 2 * 3))
        for a in range(1, L):
            for b in range(1, L):
                w = pr[:, a] * ((a == b) - pr[:, b])
                Hm[(a - 1) * 3:a * 3, (b - 1) * 3:b * 3] = (X * w[:, None]).T @ X
        W = W + np.linalg.solve(Hm, g).reshape(L - 1, 3)
    lg = np.column_stack([np.zeros(n), X @ W.T])
    ll = (lg[np.arange(n), y] - np.log(np.exp(lg).sum(1))).sum()
    assert f["status"] == 0
    assert abs(f["loglik"] - ll) / abs(ll) < 1e-6    # BFGS stops on reltol 1e-8, not at the optimum
    assert f["loglik"] <= ll + 1e-9
