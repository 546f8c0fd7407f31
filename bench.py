#!/usr/bin/env python
"""bench.py - score-cache GLM fits/sec on BASELINE.json configs[3]:
synthetic 100 mixed Gaussian / binomial / Poisson nodes x 1 000 000 rows, max.parents = 3.

    python bench.py --gpus N --steps K --warmup W [--method mle|bayes] [--impl reference]

A *step* is one pass of the hot path over one batch of the candidate list: `--batch` fits per GPU
(default 262 144), taken as evenly spaced blocks of `--block` consecutive cache rows (default 512) out
of the 16 180 000-row cache so that every step sees the cache's own mix of children / parent counts
and its own neighbour structure (adjacent rows share the child and the leading parents). Every step
uses different blocks; the data matrix (800 MB) is far larger than L2 (126 MB).

value  = fits/s with the data set resident in HBM (abn_b200_score on a DeviceData handle), all
         N GPUs, timed over exactly K steps between barrier + synchronize, max over ranks.
e2e    = fits/s through the one-shot C-ABI call abn_b200_score_cache with HOST buffers: upload of
         the 800 MB matrix from pinned host memory, Gram, fits and result download all inside the
         timed region, every step.
roofline = algorithmic bytes (passes x n x (k_data + 1) x 8 per fit) of the dominant fit
         kernel / its CUDA-event duration on the library's stream / measured HBM peak.
cpu_baseline = the CPU oracle (a restatement of the reference's C/C++ numerics; the reference itself
         needs R + GSL + Armadillo, absent here) on a bounded sample of the same batch, all host cores.
--impl reference = that CPU oracle as the timed arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = 4
TOTAL_ROWS = 16_180_000


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.lines, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Host cores the CPU arm may use (torchrun exports OMP_NUM_THREADS=1, which would starve it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def make_data(n, n_nodes):
    from abn_b200 import prep, synth
    names, data, dists, mp, _ = synth.make_config(CFG, n=n, n_nodes=n_nodes)
    return names, data, dists, mp, prep


def batch_rows(total_rows, batch, block, step, rank, world):
    """Evenly spaced blocks of `block` consecutive cache rows; different for every (step, rank)."""
    nblk = max(1, batch // block)
    stride = total_rows // nblk
    salt = (step * world + rank) * 7919 * block
    starts = (np.arange(nblk, dtype=np.int64) * stride + salt) % (total_rows - block)
    return (starts[:, None] + np.arange(block, dtype=np.int64)[None, :]).ravel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default="mle", choices=["mle", "bayes"])
    ap.add_argument("--batch", type=int, default=262144, help="fits per GPU per step (1.6 % of config 4's cache)")
    ap.add_argument("--block", type=int, default=512)
    ap.add_argument("--rows", type=int, default=1_000_000, help="observations (BASELINE config: 1e6)")
    ap.add_argument("--nodes", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=192, help="fits in the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0

    from abn_b200 import _capi as capi
    from abn_b200 import cache as abncache

    t0 = time.time()
    names, data, dists, mp, prep = make_data(args.rows, args.nodes)
    d = len(names)
    P = prep.prepare(data, dists, args.method)
    zero = np.zeros((d, d), dtype=np.int32)
    children, parents = abncache.enumerate_candidates(d, zero, zero, mp)
    total_rows = len(children)
    ctrl = abncache.to_c_control(abncache.build_control(args.method), args.method, d)
    setup_s = time.time() - t0
    workload = (f"synthetic {d} mixed gaussian/binomial/poisson nodes x {args.rows} rows, method={args.method}, "
                f"max.parents={mp} ({total_rows} candidates); step = {args.batch} fits/GPU in blocks of "
                f"{args.block} consecutive cache rows")

    # ------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        from oracle import pyoracle
        threads = host_threads()
        sample = max(threads, min(args.cpu_sample, args.batch))
        times = []
        for step in range(args.warmup + args.steps):
            if step < args.warmup and step > 0:
                continue  # one warm-up pass is enough for a CPU loop; keeps the arm within minutes
            rows = batch_rows(total_rows, args.batch, args.block, step, 0, 1)
            rows = rows[:: max(1, len(rows) // sample)][:sample]
            t = time.perf_counter()
            pyoracle.score_cache(P.X, P.node_family, P.node_col0, children[rows], parents[rows], ctrl, threads)
            if step >= args.warmup:
                times.append(time.perf_counter() - t)
        tot = sum(times)
        v = sample * len(times) / tot
        line = {
            "impl": "reference", "metric": "score-cache GLM fits/sec", "value": v, "unit": "fits/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload},
            "cpu_baseline": {"value": v, "unit": "fits/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} fits per step, every {max(1, args.batch // sample)}-th row of the step's batch, "
                                       "CPU oracle (C restatement of the reference; the reference needs R+GSL+Armadillo)"},
            "e2e": {"value": v, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        _emit(line)
        return 0

    # ------------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    ctrl.device = local_rank
    from abn_b200 import shard

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dev = abncache.DeviceData(P, device=local_rank)
    launches = 0
    kernels = {}
    algo_bytes = 0

    def run_step(step, collect):
        nonlocal launches, algo_bytes
        rows = batch_rows(total_rows, args.batch, args.block, step, rank, world)
        res = dev.score(children[rows], parents[rows], ctrl)
        if collect:
            t = dev.timing()
            launches += t["launches"]
            algo_bytes += t["algo_bytes"]
            for k in t["kernels"]:
                e = kernels.setdefault(k["name"], {"ms": 0.0, "bytes": 0, "fits": 0, "launches": 0})
                e["ms"] += k["ms"]; e["bytes"] += k["bytes"]; e["fits"] += k["fits"]; e["launches"] += 1
        return res

    for w in range(args.warmup):
        run_step(w, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t1 = time.perf_counter()
    status_counts = np.zeros(3, dtype=np.int64)
    last = None
    for s in range(args.steps):
        last = run_step(args.warmup + s, True)
        status_counts += np.bincount(last["status"], minlength=3)[:3]
    barrier()
    elapsed = time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([elapsed], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed = float(tt.item())
        # the one collective of the path: gather the score vectors of the last step over NVLink
        rows_all = args.batch * world
        gathered = shard.gather({"mlik": last["mlik"]}, rows_all, world, rank, block=args.block,
                                device=f"cuda:{local_rank}")
        assert len(gathered["mlik"]) == rows_all
    fits = args.batch * args.steps * world
    value = fits / elapsed

    # ---- end to end through the one-shot C-ABI call with host buffers -------------------------------
    e2e = None
    if not args.no_e2e:
        Xpin = torch.empty((P.X.shape[1], P.X.shape[0]), dtype=torch.float64, pin_memory=True)  # column-major
        Xpin.numpy()[:] = P.X.T
        Xh = Xpin.numpy().T  # (n, cols) Fortran-ordered view on pinned memory
        P2 = prep.Prepared(P.names, P.dists, Xh, P.node_family, P.node_col0, P.col_names, P.centre)
        import ctypes as C
        L = capi.lib()

        def one_shot(step):
            rows = batch_rows(total_rows, args.batch, args.block, step, rank, world)
            ds, k1 = capi.make_dataset(P2.X, P2.node_family, P2.node_col0)
            assert k1[0].ctypes.data == Xh.ctypes.data  # no hidden copy: the call reads the pinned buffer
            cd, k2 = capi.make_candidates(children[rows], parents[rows])
            res, arrs = capi.make_result(cd.rows, ctrl.method)
            err = C.create_string_buffer(512)
            capi.check(L.abn_b200_score_cache(C.byref(ds), C.byref(cd), C.byref(ctrl), C.byref(res), err, len(err)), err)
            return arrs, cd.rows * (1 + parents.shape[1]) * 4, cd.rows * (4 * 8 + 2 * 4)

        one_shot(0)
        barrier()
        t2 = time.perf_counter()
        e_steps = max(1, min(args.steps, 3))
        for s in range(e_steps):
            _, cand_bytes, res_bytes = one_shot(args.warmup + s)
        barrier()
        e_elapsed = time.perf_counter() - t2
        if world > 1:
            tt = torch.tensor([e_elapsed], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_elapsed = float(tt.item())
        e2e = {"value": args.batch * e_steps * world / e_elapsed, "unit": "fits/s",
               "h2d_bytes_per_step": int(P.X.nbytes + cand_bytes), "d2h_bytes_per_step": int(res_bytes),
               "steps": e_steps, "ms_per_step": 1e3 * e_elapsed / e_steps,
               "call": "abn_b200_score_cache (data upload + Gram + fits + download per step)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    peak, peak_src = measured_peaks()
    top_name, top = max(kernels.items(), key=lambda kv: kv[1]["ms"]) if kernels else (None, None)
    roofline = None
    if top and top["ms"] > 0:
        per_launch_bytes = top["bytes"] / top["launches"]
        per_launch_s = top["ms"] / top["launches"] / 1e3
        ach = per_launch_bytes / per_launch_s / 1e9
        traffic, traffic_src = None, None
        try:  # recorded ncu measurement of the same kernel (never a live number: ncu is not run inside bench.py)
            rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(top_name)
            if rec:
                traffic = rec["dram_bytes_per_launch"] / rec["fits_per_launch"] * (top["fits"] / top["launches"])
                traffic_src = (f"{rec['source']}: {rec['dram_bytes_per_launch']} B for a {rec['fits_per_launch']}-fit "
                               "launch under ncu --set full, scaled per fit to this launch size")
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "kernel": top_name, "peak_source": peak_src,
                    "bytes_per_launch": per_launch_bytes, "ms_per_launch": 1e3 * per_launch_s,
                    "fits_per_launch": top["fits"] / top["launches"],
                    "kernel_share_of_fit_time": top["ms"] / sum(k["ms"] for k in kernels.values()),
                    "all_kernels": {k: {"ms": v["ms"], "GBps": (v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["ms"] > 0 else None),
                                        "fits": v["fits"]} for k, v in sorted(kernels.items())},
                    "whole_step_GBps": algo_bytes / elapsed / 1e9, "whole_step_frac": algo_bytes / elapsed / 1e9 / peak}

    # ---- CPU baseline on a bounded sample of the same batch ----------------------------------------------
    cpu = None
    try:
        from oracle import pyoracle
        threads = host_threads()
        rows = batch_rows(total_rows, args.batch, args.block, args.warmup, 0, world)
        sample = max(threads, min(args.cpu_sample, len(rows)))
        rows = rows[:: max(1, len(rows) // sample)][:sample]
        t3 = time.perf_counter()
        ref = pyoracle.score_cache(P.X, P.node_family, P.node_col0, children[rows], parents[rows], ctrl, threads)
        cpu_s = time.perf_counter() - t3
        gpu = dev.score(children[rows], parents[rows], ctrl)
        keys = ("mlik", "aic", "bic", "mdl") if args.method == "mle" else ("mlik",)
        maxrel = 0.0
        for k in keys:
            with np.errstate(invalid="ignore"):
                r = np.abs(gpu[k] - ref[k]) / np.abs(ref[k])
            maxrel = max(maxrel, float(np.nanmax(r)))
        cpu = {"value": sample / cpu_s, "unit": "fits/s", "cores": threads, "kind": "port",
               "sample": f"{sample} fits: every {max(1, args.batch // sample)}-th row of one step's batch; CPU oracle "
                         f"(C restatement of the reference, OpenMP over candidates), {cpu_s:.1f} s",
               "max_rel_diff_gpu_vs_oracle_on_sample": maxrel,
               "status_agree": bool((gpu["status"] == ref["status"]).all())}
    except Exception as e:  # pragma: no cover
        cpu = {"value": None, "unit": "fits/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    line = {
        "metric": "score-cache GLM fits/sec", "value": value, "unit": "fits/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "l2_policy": "inputs larger than L2 (800 MB matrix vs 126 MB L2); every step "
                   "reads different candidate blocks", "setup_s": setup_s, "method": args.method,
                   "status_counts": {"ok": int(status_counts[0]), "not_converged": int(status_counts[1]),
                                     "non_finite": int(status_counts[2])}},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    _emit(line)
    dev.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def _emit(line: dict) -> None:
    """The one JSON line goes to the process's ORIGINAL stdout; everything else that libraries write to
    fd 1 (NCCL prints its version there) is sent to stderr, see _quiet_stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def _quiet_stdout() -> int:
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


if __name__ == "__main__":
    _REAL_STDOUT = _quiet_stdout()
    sys.exit(main())
