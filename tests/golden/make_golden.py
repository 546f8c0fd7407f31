"""Generates the committed golden fixtures under tests/golden/ from the reference tree.

Run in the build container (``/root/reference`` present):  python tests/golden/make_golden.py
The GPU box has no ``/root/reference``; tests only ever read the ``.npz`` files written here.

Sources (paths relative to the reference tree):
  tests/testthat/testdata/buildscorecache_ex1.Rdata   enforced Bayes golden (test-build_score_cache_bayes.R:126-135)
  inst/extdata/mycache.c2.rds / c3.rds / c4.rds       archived-vignette Bayes caches from the internal C path
  tests/testthat/testdata/mycache.Rdata               mle cache (Gaussian + separated-binomial rows usable)
  tests/testthat/testdata/forLoopContent_data.Rdata   one-hot expansion fixture
  tests/testthat/testdata/n_250_k_2_..._graph1.Rdata  four-family robustness data (expect_no_error only)
  data/var33.RData, data/FCV.RData                    bundled data sets (config 2; the vignette benchmark)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from abn_b200 import rdata  # noqa: E402

REF = os.environ.get("ABN_REFERENCE", "/root/reference")


def pack_df(prefix, obj, out):
    names, cols = rdata.data_frame(obj)
    out[f"{prefix}names"] = np.array(names)
    kinds = []
    for nm, c in zip(names, cols):
        if isinstance(c, tuple):
            kinds.append("factor")
            out[f"{prefix}v_{nm}"] = np.asarray(c[0], dtype=np.uint8 if max(c[0]) < 256 else np.int32)
            out[f"{prefix}lv_{nm}"] = np.array([str(x) for x in c[1]])
        else:
            a = np.asarray(c)
            if a.dtype.kind == "i":
                kinds.append("int")
                out[f"{prefix}v_{nm}"] = a.astype(np.int32)
            else:
                kinds.append("double")
                out[f"{prefix}v_{nm}"] = a.astype(np.float64)
    out[f"{prefix}kinds"] = np.array(kinds)


def dists_of(obj):
    return np.array([f"{k}={obj[k][0]}" for k in obj.names])


def compact_parents(nd):
    w = max(1, int(nd.sum(axis=1).max()))
    par = np.full((nd.shape[0], w), -1, dtype=np.int8)
    for r in range(nd.shape[0]):
        idx = np.nonzero(nd[r])[0]
        par[r, :len(idx)] = idx
    return par


def cache_fixture(cache, out, with_parents=True):
    nd = rdata.matrix(cache["node.defn"]).astype(np.int32)
    out["children"] = np.asarray(cache["children"], dtype=np.int16)
    if with_parents:
        out["parents"] = compact_parents(nd)
    out["node_defn_sha1"] = np.array(hashlib.sha1(np.ascontiguousarray(nd).tobytes()).hexdigest())
    out["mlik"] = np.asarray(cache["mlik"], dtype=np.float64)
    if cache.get("error.code") is not None and len(cache["error.code"]):
        out["error_code"] = np.asarray(cache["error.code"], dtype=np.int8)


def main():
    # ---- enforced Bayes golden --------------------------------------------------------------------
    d = rdata.read_rdata(f"{REF}/tests/testthat/testdata/buildscorecache_ex1.Rdata")
    out = {}
    pack_df("", d["mydat"], out)
    out["dists"] = dists_of(d["mydists"])
    out["max_par"] = np.array([int(x[0]) for x in d["max.par"]])
    cache_fixture(d["mycache"], out)
    gn, gc = rdata.data_frame(d["mycache"]["data.df"])
    out["prepped_head"] = np.column_stack([np.asarray(c, dtype=np.float64)[:64] for c in gc])
    np.savez_compressed(os.path.join(HERE, "ex1_bayes.npz"), **out)

    # ---- secondary Bayes goldens -------------------------------------------------------------------
    for tag in ("c2", "c3", "c4"):
        o = rdata.read_rds(f"{REF}/inst/extdata/mycache.{tag}.rds")
        out = {}
        pack_df("", o["data.df"], out)
        out["dists"] = dists_of(o["data.dists"])
        out["max_par"] = np.asarray(o["max.parents"]).astype(np.int32)
        out["banned"] = rdata.matrix(o["dag.banned"]).astype(np.int8)
        out["retained"] = rdata.matrix(o["dag.retained"]).astype(np.int8)
        cache_fixture(o, out, with_parents=(tag != "c2"))
        np.savez_compressed(os.path.join(HERE, f"{tag}_bayes.npz"), **out)

    # ---- mle golden (Gaussian rows 5-8, separated binomial rows 3,4,10,12 are current-code) -------
    d = rdata.read_rdata(f"{REF}/tests/testthat/testdata/mycache.Rdata")["mycache"]
    out = {}
    pack_df("", d["data.df"], out)
    out["dists"] = dists_of(d["data.dists"])
    out["max_par"] = np.asarray(d["max.parents"]).astype(np.int32).ravel()
    cache_fixture(d, out)
    for k in ("aic", "bic", "mdl"):
        out[k] = np.asarray(d[k], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "mycache_mle.npz"), **out)

    # ---- one-hot expansion fixture ------------------------------------------------------------------
    d = rdata.read_rdata(f"{REF}/tests/testthat/testdata/forLoopContent_data.Rdata")
    out = {}
    pack_df("lvl_", d["data.df.lvl"], out)
    pack_df("multi_", d["data.df.multi"], out)
    out["dists"] = dists_of(d["data.dists"])
    out["node_defn"] = rdata.matrix(d["mycache"]["node.defn"]).astype(np.int8)
    out["node_defn_multi"] = rdata.matrix(d["mycache"]["node.defn.multi"]).astype(np.int8)
    out["children"] = np.asarray(d["mycache"]["children"], dtype=np.int16)
    np.savez_compressed(os.path.join(HERE, "forloop_onehot.npz"), **out)

    # ---- four-family robustness data ---------------------------------------------------------------
    d = rdata.read_rdata(f"{REF}/tests/testthat/testdata/n_250_k_2_groups_Even_mp_7_nodedists_Balanced_s_04_graph1.Rdata")["data"]
    out = {}
    pack_df("", d["data"], out)
    out["dists"] = dists_of(d["dists"])
    np.savez_compressed(os.path.join(HERE, "four_family_250.npz"), **out)

    # ---- bundled data sets -----------------------------------------------------------------------------
    for fn, key in (("var33.RData", "var33"), ("FCV.RData", "FCV")):
        d = rdata.read_rdata(f"{REF}/data/{fn}")
        out = {}
        pack_df("", d[key], out)
        np.savez_compressed(os.path.join(HERE, f"{key}.npz"), **out)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
