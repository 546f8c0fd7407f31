"""abn_b200.fitAbn host logic (DAG -> candidate rows -> result assembly, binomial fallback) without a GPU: the
device object is replaced by the host simulation of the kernels' state machines (tests/host_sim.cpp). The same
checks run through the real C ABI in tests/test_gpu_parity.py::test_fit_abn_mle_on_a_dag."""
import numpy as np
import pytest

import abn_b200
from abn_b200 import _capi as capi
from abn_b200 import cache as abncache
from abn_b200 import prep
from oracle import pyoracle

import helpers as H


class _SimData:
    def __init__(self, prepared, device=0):
        self.prepared = prepared

    def score(self, children0, parents, control, coef_stride=0, se=False):
        return H.sim_score(self.prepared, np.asarray(children0, np.int32), np.asarray(parents, np.int32), control,
                           coef_stride=coef_stride, se=se)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _ex1():
    z = H.load_npz("ex1_bayes.npz")
    names, data = H.df_from_npz(z)
    return names, data, H.dists_from_npz(z)


def test_fit_abn_mle_matches_the_oracle_node_by_node(monkeypatch):
    monkeypatch.setattr(abncache, "DeviceData", _SimData)
    names, data, dists = _ex1()
    d = len(names)
    rng = np.random.default_rng(21)
    dag = np.zeros((d, d), dtype=int)
    for i in range(1, d):
        for p in rng.choice(np.arange(i), size=min(i, int(rng.integers(0, 4))), replace=False):
            if not (dists[names[p]] == "poisson" and dists[names[i]] == "binomial"):
                dag[i, p] = 1
    fit = abn_b200.fitAbn(dag, data, dists, method="mle")
    P = prep.prepare(data, dists, "mle")
    X = np.asarray(P.X)
    total = 0.0
    for i, nm in enumerate(names):
        ps = np.flatnonzero(dag[i])
        D = np.column_stack([np.ones(X.shape[0])] + [X[:, P.node_col0[p]] for p in ps])
        ref = {capi.GAUSSIAN: pyoracle.mle_gaussian, capi.BINOMIAL: pyoracle.mle_binomial,
               capi.POISSON: pyoracle.mle_poisson}[P.node_family[i]](D, X[:, P.node_col0[i]])
        assert ref["status"] == 0 and fit["error.code"][nm] == 0
        assert abs(fit["mliknode"][nm] - ref["loglik"]) < 1e-10 * abs(ref["loglik"])
        assert abs(fit["bicnode"][nm] - ref["bic"]) < 1e-10 * abs(ref["bic"])
        assert list(fit["coef"][nm]) == [f"{nm}|intercept"] + [names[p] for p in ps]
        coef = np.array(list(fit["coef"][nm].values()))
        se = np.array(list(fit["Stderror"][nm].values()))
        assert np.abs(coef - ref["coef"]).max() < 1e-7 * (1 + np.abs(ref["coef"]).max())
        assert (np.abs(se - ref["se"]) < 1e-6 * ref["se"]).all()
        assert fit["df"][nm] == X.shape[0] - (len(ps) + 1)
        sse = float(((X[:, P.node_col0[i]] - D @ ref["coef"]) ** 2).sum())   # e = b - A x, irls_*_fast.cpp
        assert abs(fit["sse"][nm] - sse) < 1e-8 * sse and abs(fit["mse"][nm] - sse / fit["df"][nm]) < 1e-8 * sse
        total += ref["loglik"]
    assert abs(fit["mlik"] - total) < 1e-10 * abs(total) and fit["method"] == "mle"
    # the dict form of the DAG gives the same fit
    as_dict = {names[i]: [names[p] for p in np.flatnonzero(dag[i])] for i in range(d)}
    again = abn_b200.fitAbn(as_dict, data, dists)
    assert again["mliknode"] == fit["mliknode"]


def test_fit_abn_binomial_falls_back_to_the_bias_reduced_fit(monkeypatch):
    """R/fitabn_mle.R:585-594: a separated binomial child has no finite plain mle; the reference then calls
    irls_binomial_cpp_fast_br."""
    monkeypatch.setattr(abncache, "DeviceData", _SimData)
    rng = np.random.default_rng(2)
    n = 300
    g = rng.normal(size=n)
    data = {"b": (g > 0).astype(float), "g": g, "h": rng.normal(size=n)}    # b is separated by g
    dists = {"b": "binomial", "g": "gaussian", "h": "gaussian"}
    fit = abn_b200.fitAbn({"b": ["g"], "g": [], "h": ["g"]}, data, dists)
    P = prep.prepare(data, dists, "mle")
    X = np.asarray(P.X)
    D = np.column_stack([np.ones(n), X[:, 1]])
    assert pyoracle.mle_binomial(D, X[:, 0])["status"] != 0              # the plain fit diverges
    ref = pyoracle.mle_binomial_br(D, X[:, 0])
    assert ref["status"] == fit["error.code"]["b"]
    if ref["status"] == 0:
        assert abs(fit["mliknode"]["b"] - ref["loglik"]) < 1e-8 * abs(ref["loglik"])


def test_fit_abn_bayes_modes_and_mlik(monkeypatch):
    """fitAbn(method = "bayes", compute.fixed = FALSE): modes and Laplace mlik per node vs the oracle's GSL-hybrid path."""
    monkeypatch.setattr(abncache, "DeviceData", _SimData)
    names, data, dists = _ex1()
    d = len(names)
    rng = np.random.default_rng(4)
    dag = np.zeros((d, d), dtype=int)
    for i in range(1, d):
        for p in rng.choice(np.arange(i), size=min(i, int(rng.integers(0, 3))), replace=False):
            dag[i, p] = 1
    fit = abn_b200.fitAbn(dag, data, dists, method="bayes")
    P = prep.prepare(data, dists, "bayes")
    X = np.asarray(P.X)
    for i, nm in enumerate(names):
        ps = np.flatnonzero(dag[i])
        D = np.column_stack([np.ones(X.shape[0])] + [X[:, P.node_col0[p]] for p in ps])
        ref = pyoracle.bayes_fit(P.node_family[i], D, X[:, P.node_col0[i]], solver=capi.SOLVER_HYBRID)
        assert fit["error.code"][nm] == ref["status"] == 0 and fit["error.code.desc"][nm] == "success"
        assert abs(fit["mliknode"][nm] - ref["loglik"]) < 1e-9 * abs(ref["loglik"])
        modes = fit["modes"][nm]
        want = [f"{nm}|(Intercept)"] + [f"{nm}|{names[p]}" for p in ps] + ([f"{nm}|precision"] if dists[nm] == "gaussian" else [])
        assert list(modes) == want
        got = np.array(list(modes.values()))
        assert np.abs(got - ref["coef"]).max() < 1e-7 * (1 + np.abs(ref["coef"]).max())
    assert abs(fit["mlik"] - sum(fit["mliknode"].values())) < 1e-9 and fit["method"] == "bayes"


def test_fit_abn_argument_checks():
    names, data, dists = _ex1()
    with pytest.raises(ValueError):
        abn_b200.fitAbn(np.zeros((10, 10), int), data, dists, method="inla")
    with pytest.raises(ValueError):
        abn_b200.fitAbn(np.eye(10, dtype=int), data, dists)
    with pytest.raises(ValueError):
        abn_b200.fitAbn(np.zeros((3, 3), int), data, dists)


def test_fit_abn_with_a_multinomial_parent_has_no_intercept(monkeypatch):
    """R/fitabn_mle.R:546-548: a multinomial parent is expanded into its level columns and the design has no
    intercept; the coefficient names are the level columns (:806-808)."""
    monkeypatch.setattr(abncache, "DeviceData", _SimData)
    rng = np.random.default_rng(8)
    n = 600
    codes = rng.integers(1, 4, size=n)
    g = rng.normal(size=n)
    y = (rng.random(n) < 1 / (1 + np.exp(-(0.4 * g + 0.5 * (codes == 2) - 0.3)))).astype(float)
    data = {"b": y, "g": g, "m": (codes, ["x", "y", "z"])}
    dists = {"b": "binomial", "g": "gaussian", "m": "multinomial"}
    fit = abn_b200.fitAbn({"b": ["g", "m"], "g": [], "m": []}, data, dists)
    P = prep.prepare(data, dists, "mle")
    X = np.asarray(P.X)
    cols = [P.node_col0[1]] + list(range(P.node_col0[2], P.node_col0[3]))
    D = np.column_stack([X[:, c] for c in cols])                       # no intercept
    ref = pyoracle.mle_binomial(D, X[:, P.node_col0[0]])
    assert ref["status"] == 0 and fit["error.code"]["b"] == 0
    assert list(fit["coef"]["b"]) == [P.col_names[c] for c in cols]
    assert abs(fit["mliknode"]["b"] - ref["loglik"]) < 1e-10 * abs(ref["loglik"])
    assert np.abs(np.array(list(fit["coef"]["b"].values())) - ref["coef"]).max() < 1e-7
    assert (np.abs(np.array(list(fit["Stderror"]["b"].values())) - ref["se"]) < 1e-6 * ref["se"]).all()
    assert fit["coef"]["m"] is None and np.isfinite(fit["mliknode"]["m"])   # multinomial child: scores only
