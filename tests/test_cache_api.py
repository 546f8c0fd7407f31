"""Argument handling of the Python mirror of buildScoreCache / build.control (no GPU needed)."""
import numpy as np
import pytest

from abn_b200 import cache as abncache
from abn_b200 import synth


def _data():
    return synth.make_columns(40, ["gaussian", "binomial", "poisson", "binomial"], seed=11)


def test_build_control_defaults():
    b = abncache.build_control("bayes")
    assert (b["mean"], b["prec"], b["loggam.shape"], b["loggam.inv.scale"], b["max.iters"], b["epsabs"]) == \
        (0, 0.001, 1, 5e-05, 100, 1e-07)
    m = abncache.build_control("mle", max_iters=50)
    assert m["max.iters"] == 50 and m["tol"] == 1e-8 and "prec" not in m
    with pytest.raises(ValueError):
        abncache.build_control("foo")
    with pytest.raises(TypeError):
        abncache.build_control("mle", nonsense=1)


def test_argument_validation():
    names, data, dists = _data()
    with pytest.raises(ValueError):
        abncache.buildScoreCache(None, dists)
    with pytest.raises(ValueError):
        abncache.buildScoreCache(data, dists, method="xyz")
    with pytest.raises(ValueError):
        abncache.buildScoreCache(data, {k: v for k, v in list(dists.items())[:2]}, method="mle", max_parents=1)
    ban = np.zeros((4, 4), int)
    ret = np.zeros((4, 4), int)
    ban[0, 1] = ret[0, 1] = 1
    with pytest.raises(ValueError, match="inconsistent"):
        abncache.buildScoreCache(data, dists, method="mle", max_parents=2, dag_banned=ban, dag_retained=ret)
    ret2 = np.zeros((4, 4), int)
    ret2[0, 1:] = 1
    with pytest.raises(ValueError, match="max.parents"):
        abncache.buildScoreCache(data, dists, method="mle", max_parents=1, dag_retained=ret2)
    with pytest.raises(ValueError, match="own parent"):
        abncache.buildScoreCache(data, dists, method="mle", max_parents=1, dag_banned=np.eye(4, dtype=int))
    with pytest.raises(ValueError, match="ISSUE"):
        abncache.buildScoreCache(data, dists, method="mle", max_parents={n: i % 2 + 1 for i, n in enumerate(names)})
    with pytest.raises(NotImplementedError):
        abncache.buildScoreCache(data, dists, method="mle", group_var="g1")
    md = dict(dists)
    md[names[1]] = "multinomial"
    with pytest.raises(ValueError, match="Multinomial"):
        abncache.buildScoreCache(data, md, method="bayes", max_parents=1)


def test_dry_run_returns_node_definitions_without_a_gpu():
    names, data, dists = _data()
    with pytest.warns(UserWarning):
        out = abncache.buildScoreCache(data, dists, method="bayes", max_parents=9, dry_run=True)
    assert out["node.defn"].shape == (4 * 8, 4) and list(out["children"][:8]) == [1] * 8
    out = abncache.buildScoreCache(data, dists, method="mle", max_parents=2, dry_run=True,
                                   dag_banned="~g1|b1", dag_retained={"p1": ["g1"]})
    nd, ch = out["node.defn"], out["children"]
    assert not nd[(ch == 1), 1].any() and nd[(ch == 3), 0].all()


def test_formula_parser():
    m = abncache._parse_formula("~a|b:c+d|.", ["a", "b", "c", "d"])
    assert m.tolist() == [[0, 1, 1, 0], [0, 0, 0, 0], [0, 0, 0, 0], [1, 1, 1, 0]]


def test_binomial_columns_are_validated_and_recoded_like_the_reference():
    """R/build_score_cache_mle.R:408-410 (anything not of class "numeric" is recoded through factor()) and
    check.valid.data (R/abn-internal.R:277-290): binary, both cases observed."""
    import numpy as np
    import pytest
    from abn_b200 import prep
    rng = np.random.default_rng(0)
    g = rng.standard_normal(40)
    ones_twos = rng.integers(1, 3, 40)                       # integer column coded 1/2
    P = prep.prepare({"b": ones_twos, "g": g}, {"b": "binomial", "g": "gaussian"}, "mle")
    assert set(np.unique(P.X[:, 0])) == {0.0, 1.0} and np.array_equal(P.X[:, 0], ones_twos - 1.0)
    with pytest.raises(ValueError, match="coded 0/1"):
        prep.prepare({"b": ones_twos.astype(float), "g": g}, {"b": "binomial", "g": "gaussian"}, "mle")
    with pytest.raises(ValueError, match="must be binary"):
        prep.prepare({"b": rng.integers(0, 3, 40), "g": g}, {"b": "binomial", "g": "gaussian"}, "mle")
    with pytest.raises(ValueError, match="both cases"):
        prep.prepare({"b": np.zeros(40), "g": g}, {"b": "binomial", "g": "gaussian"}, "bayes")


def test_kernel_names_cover_every_bucket_of_the_plan():
    """abn_timing.kernel_id -> name (plan.cuh: wide, Gaussian, 2 x (families x K x intercept), then the table kernels
    CELL_BUCKET0 + 3 * (Poisson child) + class)."""
    from abn_b200 import _capi as capi
    cell0 = 2 + 6 * (capi.FAST_K + 1) * 2
    assert abncache.kernel_name(0) == "fit_wide_kernel" and abncache.kernel_name(1) == "gauss_kernel"
    assert abncache.kernel_name(2 + 2 * 4 + 1) == "fit_kernel<binomial,K=4,icpt=1>"
    assert abncache.kernel_name(2 + 2 * ((capi.FAST_K + 1) + 4) + 1) == "fit_kernel<poisson,K=4,icpt=1>"
    names = [abncache.kernel_name(cell0 + i) for i in range(6)]
    assert names == ["fit_cell_kernel<binomial,small>", "fit_cell_kernel<binomial,large>", "fit_cell_kernel<binomial,planes>",
                     "fit_cell_kernel<poisson,small>", "fit_cell_kernel<poisson,large>", "fit_cell_kernel<poisson,planes>"]


def test_gather_slab_on_the_host_ignores_the_pinned_switch():
    """shard.gather_slab(pinned=True) stages through page-locked buffers for CUDA slabs only; a host slab (the gloo
    tests, callers without a device) takes the plain path and returns the same arrays."""
    import torch
    from abn_b200 import shard
    rows, block = 5000, 512
    rng = np.random.default_rng(3)
    local = {"mlik": rng.normal(size=rows), "aic": rng.normal(size=rows), "bic": rng.normal(size=rows), "mdl": rng.normal(size=rows),
             "status": rng.integers(0, 3, rows).astype(np.int32), "iters": rng.integers(0, 99, rows).astype(np.int32)}
    slab = shard.slab_from_dict(local, rows, 1, block)
    assert isinstance(slab, torch.Tensor) and not slab.is_cuda
    for pinned in (False, True):
        out = shard.gather_slab(slab, rows, 1, block, pinned=pinned)
        for k, v in local.items():
            assert np.array_equal(out[k], v), (pinned, k)
