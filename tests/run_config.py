#!/usr/bin/env python
"""Run one BASELINE.json config at its FULL size through the public API and report wall time,
fits/s, flags and parity against the CPU oracle (full for the small real-data configs, a seeded
random sample of candidates for the synthetic ones).

    python tests/run_config.py --config 1|2|3|4|5 [--method mle|bayes] [--sample 1500] [--limit N]

Writes one JSON line (also to --out). Needs a GPU. TEST INFRASTRUCTURE (lives under tests/ because it
loads the CPU oracle as the checker; it is not collected by pytest).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def load_real(cfg):
    import helpers as H
    if cfg == 1:
        z = H.load_npz("ex1_bayes.npz")
        names, data = H.df_from_npz(z)
        return names, data, H.dists_from_npz(z), 3, "mle"
    z = H.load_npz("var33.npz")
    names, data = H.df_from_npz(z)
    dists = {nm: ("binomial" if isinstance(v, tuple) else "gaussian") for nm, v in data.items()}
    return names, data, dists, 3, "bayes"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--method", default=None)
    ap.add_argument("--sample", type=int, default=1500, help="candidates checked against the oracle (synthetic configs)")
    ap.add_argument("--limit", type=int, default=0, help="score only the first N candidates of an evenly strided subset")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    import abn_b200
    from abn_b200 import cache as abncache
    from abn_b200 import prep, synth
    from oracle import pyoracle

    t0 = time.time()
    if args.config in (1, 2):
        names, data, dists, mp, method = load_real(args.config)
    else:
        names, data, dists, mp, method = synth.make_config(args.config)
    method = args.method or method
    d = len(names)
    P = prep.prepare(data, dists, method)
    zero = np.zeros((d, d), dtype=np.int32)
    ch0, par = abncache.enumerate_candidates(d, zero, zero, mp)
    rows_total = len(ch0)
    if args.limit and args.limit < rows_total:
        blk = 512
        nblk = max(1, args.limit // blk)
        starts = (np.arange(nblk, dtype=np.int64) * (rows_total // nblk))
        sel = (starts[:, None] + np.arange(blk)[None, :]).ravel()
        sel = sel[sel < rows_total]
        ch0, par = ch0[sel], par[sel]
    setup_s = time.time() - t0
    ctrl = abncache.to_c_control(abncache.build_control(method), method, d)

    # one process per GPU under torchrun: every rank holds the data set, scores its interleaved share
    # of the candidate list, and the score vectors are all-gathered over NCCL (abn_b200/shard.py)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        from abn_b200 import shard
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        ctrl.device = local_rank
        idx = shard.partition(len(ch0), world, rank, block=1024)
    else:
        idx = np.arange(len(ch0))
    t1 = time.time()
    with abncache.DeviceData(P, device=local_rank) as dev:
        create_s = time.time() - t1
        if world > 1:
            shard._pinned_staging(len(ch0), ring=1)  # page-locked landing buffers of the gather: once per process, not per call
            dist.barrier()
        t2 = time.time()
        if world == 1 and not args.limit:
            # the product path of buildScoreCache: candidates generated on the device, nothing uploaded
            res = dev.score_enumerated(zero, zero, mp, ctrl)
            score_local_s = time.time() - t2
            assert np.array_equal(res["children"], ch0) and np.array_equal(res["parents"], par)
        elif world > 1:
            # the rank's scores stay on its device (40-byte rows of a result slab) for the all-gather below
            slab = shard.new_slab(len(ch0), world, 1024, device=f"cuda:{local_rank}")
            dev.score_to_slab(ch0[idx], par[idx], ctrl, slab)
            torch.cuda.synchronize()
            score_local_s = time.time() - t2
        else:
            res = dev.score(ch0[idx], par[idx], ctrl)
            score_local_s = time.time() - t2
        timing = dev.timing()
    gather_s = 0.0
    if world > 1:
        tt = torch.tensor([score_local_s], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        score_s = float(tt.item())
        dist.barrier()
        torch.cuda.synchronize()
        t4 = time.time()
        res = shard.gather_slab(slab, len(ch0), world, 1024, pinned=True)   # one all-gather, one permute, one D2H
        torch.cuda.synchronize()
        gather_s = time.time() - t4
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
            return
    else:
        score_s = score_local_s
    fits = len(ch0)
    fam = P.node_family[ch0]

    rng = np.random.default_rng(20261019 + args.config)
    if args.config in (1, 2) and not args.limit:
        sample = np.arange(fits)
    else:
        sample = np.sort(rng.choice(fits, size=min(args.sample, fits), replace=False))
    t3 = time.time()
    ref = pyoracle.score_cache(P.X, P.node_family, P.node_col0, ch0[sample], par[sample], ctrl,
                               max(1, len(os.sched_getaffinity(0))))
    oracle_s = time.time() - t3
    keys = ("mlik", "aic", "bic", "mdl") if method == "mle" else ("mlik",)
    parity = {}
    for k in keys:
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = np.abs(res[k][sample] - ref[k]) / np.abs(ref[k])
        rel[np.isnan(res[k][sample]) & np.isnan(ref[k])] = 0.0
        parity[k] = {"max_rel": float(np.nanmax(rel)), "p99_rel": float(np.nanquantile(rel, 0.99)),
                     "n_gt_1e-8": int(np.sum(rel > 1e-8)), "nan_agree": bool((np.isnan(res[k][sample]) == np.isnan(ref[k])).all())}
    by_family = {}
    for code, nm in ((1, "binomial"), (2, "gaussian"), (3, "poisson"), (4, "multinomial")):
        m = fam[sample] == code
        if m.any():
            with np.errstate(invalid="ignore", divide="ignore"):
                rel = np.abs(res["mlik"][sample][m] - ref["mlik"][m]) / np.abs(ref["mlik"][m])
            by_family[nm] = {"rows_checked": int(m.sum()), "max_rel_mlik": float(np.nanmax(rel)),
                             "mean_passes_gpu": float(res["iters"][sample][m].mean())}
    line = {
        "config": args.config, "method": method, "n": int(P.n), "d": d, "max_parents": int(mp),
        "candidates_total": int(rows_total), "candidates_scored": int(fits),
        "n_gpus": world, "allgather_s": gather_s,
        "setup_s": setup_s, "data_create_s": create_s, "score_s": score_s,
        "fits_per_s": fits / score_s, "fits_per_s_incl_upload": fits / (score_s + create_s),
        "device_ms": {k: timing[k] for k in ("h2d_ms", "plan_ms", "fit_ms", "d2h_ms")},
        "algo_GBps": timing["algo_bytes"] / (timing["fit_ms"] / 1e3) / 1e9 if timing["fit_ms"] > 0 else None,
        "status_counts": {str(i): int(c) for i, c in enumerate(np.bincount(res["status"], minlength=3))},
        "oracle": {"rows_checked": int(len(sample)), "seconds": oracle_s, "cores": len(os.sched_getaffinity(0)),
                   "fits_per_s": len(sample) / oracle_s,
                   "status_agree": float((res["status"][sample] == ref["status"]).mean())},
        "parity": parity, "by_family": by_family,
        "kernels": sorted(timing["kernels"], key=lambda k: -k["ms"])[:6],
    }
    s = json.dumps(line)
    print(s)
    if args.out:
        with open(args.out, "w") as f:
            f.write(s + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
