#!/usr/bin/env python
"""Time abn_b200.fitAbn on BASELINE config 4's data (100 mixed nodes x 1 M rows, a random DAG with <= 3 parents per
node) and check it against the CPU oracle on a few nodes. Needs a GPU. TEST INFRASTRUCTURE (uses oracle/).

    python tests/run_fitabn.py [--rows 1000000] [--nodes 100] [--check 6] [--out profiles/fitabn_r1.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--nodes", type=int, default=100)
    ap.add_argument("--check", type=int, default=6)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import abn_b200
    from abn_b200 import _capi as capi
    from abn_b200 import prep, synth
    from oracle import pyoracle

    names, data, dists, mp, _ = synth.make_config(4, n=a.rows, n_nodes=a.nodes)
    d = len(names)
    rng = np.random.default_rng(1)
    dag = np.zeros((d, d), dtype=int)
    for i in range(1, d):
        for p in rng.choice(np.arange(i), size=min(i, 3), replace=False):
            dag[i, p] = 1
    out = {"workload": f"fitAbn on {d} nodes x {a.rows} rows, random DAG with <= 3 parents per node"}
    for method in ("mle", "bayes"):
        abn_b200.fitAbn(dag, data, dists, method=method)                 # warm-up (library load, pool)
        t = time.perf_counter()
        fit = abn_b200.fitAbn(dag, data, dists, method=method)
        out[f"{method}_seconds_incl_prep_and_upload"] = time.perf_counter() - t
        P = prep.prepare(data, dists, method)
        X = np.asarray(P.X)
        worst, cpu = 0.0, 0.0
        nodes = rng.choice(np.arange(1, d), size=min(a.check, d - 1), replace=False)
        for i in nodes:
            ps = np.flatnonzero(dag[i])
            D = np.column_stack([np.ones(X.shape[0])] + [X[:, P.node_col0[p]] for p in ps])
            y = X[:, P.node_col0[i]]
            t = time.perf_counter()
            if method == "mle":
                ref = {capi.GAUSSIAN: pyoracle.mle_gaussian, capi.BINOMIAL: pyoracle.mle_binomial,
                       capi.POISSON: pyoracle.mle_poisson}[P.node_family[i]](D, y)
            else:
                ref = pyoracle.bayes_fit(P.node_family[i], D, y, solver=capi.SOLVER_HYBRID)
            cpu += time.perf_counter() - t
            worst = max(worst, abs(fit["mliknode"][names[i]] - ref["loglik"]) / abs(ref["loglik"]))
        out[f"{method}_max_rel_diff_mliknode_vs_oracle"] = worst
        out[f"{method}_oracle_seconds_per_node_1_core"] = cpu / len(nodes)
        out[f"{method}_error_codes"] = {str(k): int(v) for k, v in zip(*np.unique(list(fit["error.code"].values()), return_counts=True))}
    line = json.dumps(out)
    print(line)
    if a.out:
        open(a.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
