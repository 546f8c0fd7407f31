"""GPU-side probe (not a test): where exactly the CUDA path, the oracle and the reference goldens differ.
Writes gpurun_out/probe_parity.json; the -m gpu tests assert what this prints."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from abn_b200 import cache as abncache  # noqa: E402
from abn_b200 import prep  # noqa: E402

out = {}


def report(tag, res, ref, golden_ec=None, golden_mlik=None, chaotic=None):
    rel = H.rel_err(res["mlik"], ref["mlik"])
    diff = np.nonzero(res["status"] != ref["status"])[0]
    ok = (ref["status"] == 0) & (res["status"] == 0)
    big = np.nonzero(ok & (rel > 1e-8))[0]
    r = {"rows": int(len(rel)), "status_diff_vs_oracle": diff.tolist()[:60], "n_status_diff": int(len(diff)),
         "rows_rel_gt_1e-8_vs_oracle": big.tolist()[:60], "n_rel_gt_1e-8": int(len(big)),
         "max_rel_ok_rows": float(np.nanmax(rel[ok])) if ok.any() else None}
    if chaotic is not None:
        r["n_chaotic"] = int(chaotic.sum())
        r["status_diff_not_chaotic"] = [int(i) for i in diff if not chaotic[i]]
        r["rel_gt_1e-8_not_chaotic"] = [int(i) for i in big if not chaotic[i]]
        okc = ok & ~chaotic
        r["max_rel_ok_not_chaotic"] = float(np.nanmax(rel[okc])) if okc.any() else None
    if golden_ec is not None:
        dg = np.nonzero(res["status"] != golden_ec)[0]
        r["status_diff_vs_golden"] = dg.tolist()[:60]
        r["oracle_status_diff_vs_golden"] = np.nonzero(ref["status"] != golden_ec)[0].tolist()[:60]
        relg = H.rel_err(res["mlik"], golden_mlik)
        okg = (golden_ec == 0) & (res["status"] == 0)
        r["max_rel_vs_golden_ok"] = float(np.nanmax(relg[okg]))
        r["rows_rel_gt_1e-8_vs_golden"] = np.nonzero(okg & (relg > 1e-8))[0].tolist()[:80]
        if chaotic is not None:
            r["status_diff_vs_golden_not_chaotic"] = [int(i) for i in dg if not chaotic[i]]
            r["rows_rel_gt_1e-8_vs_golden_not_chaotic"] = [int(i) for i in np.nonzero(okg & (relg > 1e-8))[0] if not chaotic[i]]
    out[tag] = r
    print(tag, json.dumps(r)[:1500], flush=True)


for tag in ("c4", "c3", "c2"):
    z = H.load_npz(f"{tag}_bayes.npz")
    names, data = H.df_from_npz(z)
    dists = H.dists_from_npz(z)
    d = len(names)
    P = prep.prepare(data, dists, "bayes")
    ch0, par = abncache.enumerate_candidates(d, z["banned"], z["retained"], z["max_par"])
    ctrl = H.control("bayes", d)
    res = H.gpu_score(P, ch0, par, ctrl)
    chaotic, ref, _ = H.oracle_order_sensitive(P, ch0, par, ctrl)
    report(tag, res, ref, z["error_code"], z["mlik"], chaotic)

z = H.load_npz("var33.npz")
names, data = H.df_from_npz(z)
dists = {nm: ("binomial" if isinstance(v, tuple) else "gaussian") for nm, v in data.items()}
d = len(names)
P = prep.prepare(data, dists, "bayes")
ch0, par = abncache.enumerate_candidates(d, np.zeros((d, d), int), np.zeros((d, d), int), 3)
ctrl = H.control("bayes", d)
res = H.gpu_score(P, ch0, par, ctrl)
chaotic, ref, _ = H.oracle_order_sensitive(P, ch0, par, ctrl)
report("var33", res, ref, chaotic=chaotic)

# the round-1 fuzz case: small n, strong effects (tests/test_host_sim.py)
rng = np.random.default_rng(7)
fams = ["binomial", "poisson", "gaussian", "binomial", "poisson", "gaussian"]
from oracle import pyoracle  # noqa: E402
for n, scale in [(40, 2.5), (80, 1.0), (150, 2.5), (400, 0.3), (1500, 1.0), (6000, 2.5)]:
    zz = rng.standard_normal((n, 3))
    data, dists = {}, {}
    for j, f in enumerate(fams):
        s = rng.uniform(-1, 1) + zz @ rng.normal(0, scale, 3)
        if f == "binomial":
            data[f"v{j}"] = (rng.random(n) < 1 / (1 + np.exp(-s))).astype(float)
        elif f == "poisson":
            data[f"v{j}"] = rng.poisson(np.exp(np.clip(0.5 + 0.5 * s, -5, 5))).astype(float)
        else:
            data[f"v{j}"] = s + rng.standard_normal(n)
        dists[f"v{j}"] = f
    if any(np.unique(v).size < 2 for v in data.values()):
        continue
    d = len(fams)
    P = prep.prepare(data, dists, "mle")
    ch, nd = pyoracle.enumerate_mle(d, np.zeros((d, d), int), np.zeros((d, d), int), 3)
    par = H.parents_from_node_defn(nd, 3)
    ctrl = H.control("mle", d)
    res = H.gpu_score(P, ch.astype(np.int32), par, ctrl)
    chaotic, ref, _ = H.oracle_order_sensitive(P, ch.astype(np.int32), par, ctrl)
    report(f"fuzz_n{n}", res, ref, chaotic=chaotic)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_parity.json"), "w"), indent=1)
