"""Multinomial children of BASELINE.json configs[4] (40 nodes incl. one 4-level multinomial, 500 000 rows,
max.parents = 4) on ONE GPU: a strided slice of the 92 171 parent sets of the multinomial node, default path
(lean pass: design columns only, X'y_l from the Gram matrix) against ABN_FLAG_EXACT_PATH (round-1 row arithmetic,
response streamed), and a seeded sample against the oracle. usage: python tests/run_mn_slice.py [--fits N] [--out F]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import helpers as H  # noqa: E402
from abn_b200 import cache as abncache  # noqa: E402
from abn_b200 import prep, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fits", type=int, default=2368)
    ap.add_argument("--rows", type=int, default=500_000)
    ap.add_argument("--sample", type=int, default=12)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    names, data, dists, mp, _ = synth.make_config(5, n=a.rows)
    d = len(names)
    P = prep.prepare(data, dists, "mle")
    zero = np.zeros((d, d), dtype=np.int32)
    node = names.index("m1")
    ch0, par = abncache.enumerate_candidates(d, zero, zero, mp, which_nodes=np.array([node], dtype=np.int32))
    assert (ch0 == node).all()
    sel = np.linspace(0, len(ch0) - 1, a.fits).astype(np.int64)
    ch0, par = ch0[sel], par[sel]
    out = {"config": 5, "rows": a.rows, "fits": int(len(ch0)), "candidates_of_node": int(len(sel))}
    with abncache.DeviceData(P) as dev:
        res = {}
        for tag, kw in (("default", {}), ("exact_path", {"exact.path": True})):
            ctrl = H.control("mle", d, **kw)
            dev.score(ch0[:64], par[:64], ctrl)  # warm-up (attributes, pools)
            t = time.perf_counter()
            r = dev.score(ch0, par, ctrl)
            el = time.perf_counter() - t
            res[tag] = r
            out[tag] = {"seconds": el, "fits_per_s": len(ch0) / el, "mean_passes": float(r["iters"].mean()),
                        "not_converged": int((r["status"] == 1).sum()),
                        "kernels": [{k: v for k, v in kk.items() if k in ("name", "fits", "ms")} for kk in dev.timing()["kernels"]]}
    out["default_vs_exact"] = {"max_rel_mlik": float(H.rel_err(res["default"]["mlik"], res["exact_path"]["mlik"]).max()),
                               "status_equal": bool(np.array_equal(res["default"]["status"], res["exact_path"]["status"])),
                               "passes_equal": int((res["default"]["iters"] == res["exact_path"]["iters"]).sum())}
    if a.sample:
        rng = np.random.default_rng(5)
        rows = np.sort(rng.choice(len(ch0), size=min(a.sample, len(ch0)), replace=False))
        t = time.perf_counter()
        ref = H.oracle_score(P, ch0[rows], par[rows], H.control("mle", d))
        out["oracle"] = {"rows": int(len(rows)), "seconds": time.perf_counter() - t,
                         "max_rel_mlik": float(H.rel_err(res["default"]["mlik"][rows], ref["mlik"]).max()),
                         "status_equal": bool(np.array_equal(res["default"]["status"][rows], ref["status"]))}
    txt = json.dumps(out)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
