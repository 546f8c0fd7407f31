#!/usr/bin/env python
"""The one timing the reference publishes for this path: the microbenchmark of vignettes/multiprocessing.Rmd:53-128,
buildScoreCache on FCV[, -13] (300 rows x 14 nodes: 12 binomial, GroupSize Poisson, Age Gaussian), max.parents = 5
=> 33 320 fits; read off vignettes/benchmarkBoxPlot-1.png: mle ~20.4 s, bayes ~12.9 s on one core of unstated
hardware. Runs it through abn_b200.buildScoreCache and checks every row against the CPU
oracle. Needs a GPU. TEST INFRASTRUCTURE (uses oracle/).

    python tests/run_fcv.py [--out profiles/fcv_r1.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import helpers as H
    import abn_b200
    from abn_b200 import cache as abncache
    from abn_b200 import prep

    z = H.load_npz("FCV.npz")
    names, data = H.df_from_npz(z)
    names = [n for n in names if n != "Sex"]
    data = {n: data[n] for n in names}
    dists = {n: "binomial" for n in names}
    dists["GroupSize"], dists["Age"] = "poisson", "gaussian"
    d = len(names)
    out = {"workload": "FCV[, -13]: 300 rows x 14 nodes, max.parents = 5 (vignettes/multiprocessing.Rmd:53-128)",
           "reference_published_s": {"mle_1core": 20.4, "bayes_1core": 12.9}}
    for method in ("mle", "bayes"):
        abn_b200.buildScoreCache(data, dists, method=method, max_parents=5)     # warm-up (library load, pool)
        t = time.perf_counter()
        cache = abn_b200.buildScoreCache(data, dists, method=method, max_parents=5)
        dt = time.perf_counter() - t
        rows = len(cache["children"])
        P = prep.prepare(data, dists, method)
        zero = np.zeros((d, d), dtype=np.int32)
        ch0, par = abncache.enumerate_candidates(d, zero, zero, 5)
        ctrl = H.control(method, d)
        t = time.perf_counter()
        ref = H.oracle_score(P, ch0, par, ctrl)
        dt_ref = time.perf_counter() - t
        mlik = np.asarray(cache["mlik"], dtype=float)
        ok = (ref["status"] == 0) & np.isfinite(ref["mlik"])
        code = np.asarray(cache["error.code"]) if "error.code" in cache else None
        if code is not None and code.shape != ref["status"].shape:
            code = None                                   # mle caches carry no error.code
        out[method] = {"fits": int(rows), "seconds": dt, "fits_per_s": rows / dt,
                       "oracle_seconds_all_cores": dt_ref,
                       "max_rel_diff_vs_oracle_ok_rows": float(H.rel_err(mlik[ok], ref["mlik"][ok]).max()),
                       "p999_rel_diff": float(np.quantile(H.rel_err(mlik[ok], ref["mlik"][ok]), 0.999)),
                       "rows_ok": int(ok.sum()),
                       "flags_equal": None if code is None else float((code == ref["status"]).mean())}
    line = json.dumps(out)
    print(line)
    if a.out:
        open(a.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
