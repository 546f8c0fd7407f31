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


def batch_rows(total_rows, batch, block, step):
    """One GLOBAL batch: evenly spaced blocks of `block` consecutive cache rows; different for every step."""
    nblk = max(1, batch // block)
    stride = total_rows // nblk
    salt = step * 7919 * block
    starts = (np.arange(nblk, dtype=np.int64) * stride + salt) % (total_rows - block)
    return (starts[:, None] + np.arange(block, dtype=np.int64)[None, :]).ravel()


def fp64_peak():
    """Measured DFMA ceiling of this pool's B200 (tools/fp64_peak.cu, profiles/fp64_peak_r2.json)."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r2.json")))
        return float(rec["dfma_T_lane_fma_per_s"]), "measured (profiles/fp64_peak_r2.json: tools/fp64_peak.cu, DFMA issue rate)"
    except Exception:
        return 18.6, "nominal (148 SMs x 64 DFMA lanes x 1.965 GHz)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default="mle", choices=["mle", "bayes"])
    ap.add_argument("--batch", type=int, default=2_097_152,
                    help="fits per step over ALL GPUs (13 % of config 4's cache); strong scaling: each rank scores batch / N")
    ap.add_argument("--block", type=int, default=512)
    ap.add_argument("--rows", type=int, default=1_000_000, help="observations (BASELINE config: 1e6)")
    ap.add_argument("--nodes", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=192, help="fits in the cpu_baseline sample")
    ap.add_argument("--side-batch", type=int, default=262_144, help="fits per step of the exact-path and bayes sub-records")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the exact-path and bayes sub-records")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0

    t0 = time.time()
    names, data, dists, mp, prep = make_data(args.rows, args.nodes)
    d = len(names)
    P = prep.prepare(data, dists, args.method)
    zero = np.zeros((d, d), dtype=np.int32)
    if args.impl == "reference":
        # the reference arm touches nothing of the product: enumeration by the oracle's own restatement of buildcachematrix
        from oracle import pyoracle
        children1, node_defn = pyoracle.enumerate_bayes(d, zero, zero, np.full(d, mp, dtype=np.int32))
        children = np.asarray(children1, dtype=np.int32) - (1 if np.min(children1) >= 1 else 0)
        parents = np.full((len(children), mp), -1, dtype=np.int32)
        rr, cc = np.nonzero(node_defn)
        slot = np.zeros(len(children), dtype=np.int64)
        for r, c in zip(rr, cc):  # rows are visited in order, columns ascending within a row
            parents[r, slot[r]] = c
            slot[r] += 1
    else:
        from abn_b200 import cache as abncache
        children, parents = abncache.enumerate_candidates(d, zero, zero, mp)
    total_rows = len(children)
    setup_s = time.time() - t0
    workload = (f"synthetic {d} mixed gaussian/binomial/poisson nodes x {args.rows} rows, method={args.method}, "
                f"max.parents={mp} ({total_rows} candidates); step = {args.batch} fits over all GPUs in blocks of "
                f"{args.block} consecutive cache rows")
    config = {"workload": workload, "method": args.method, "batch_fits_per_step": args.batch, "block": args.block,
              "l2_policy": "inputs larger than L2 (800 MB matrix vs 126 MB L2); every step reads different candidate blocks"}

    # ------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        from abn_b200 import cache as abncache  # control struct only (ctypes layout of include/abn_b200.h)
        ctrl = abncache.to_c_control(abncache.build_control(args.method), args.method, d)
        threads = host_threads()
        sample = max(threads, min(args.cpu_sample, args.batch))
        times = []
        for step in range(args.warmup + args.steps):
            if step < args.warmup and step > 0:
                continue  # one warm-up pass is enough for a CPU loop; keeps the arm within minutes
            rows = batch_rows(total_rows, args.batch, args.block, step)
            rows = rows[:: max(1, len(rows) // sample)][:sample]
            t = time.perf_counter()
            pyoracle.score_cache(P.X, P.node_family, P.node_col0, children[rows], parents[rows], ctrl, threads)
            if step >= args.warmup:
                times.append(time.perf_counter() - t)
        tot = sum(times)
        v = sample * len(times) / tot
        sample_txt = (f"{sample} fits per step (every {max(1, args.batch // sample)}-th row of the step's {args.batch}-fit "
                      "batch), CPU oracle (C restatement of the reference; the reference needs R+GSL+Armadillo)")
        config["reference_sample"] = sample_txt
        line = {
            "impl": "reference", "metric": "score-cache GLM fits/sec", "value": v, "unit": "fits/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": v, "unit": "fits/s", "cores": threads, "kind": "port", "sample": sample_txt},
            "e2e": {"value": v, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        _emit(line)
        return 0

    # ------------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist

    from abn_b200 import _capi as capi
    from abn_b200 import shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    devname = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(devname))
    ctrl = abncache.to_c_control(abncache.build_control(args.method), args.method, d)
    ctrl.device = local_rank

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class DeviceTimer:
        """Elapsed time of a region on the device: CUDA events recorded on an idle device (barrier() has
        synchronised every stream, the library's included) at both ends; the host clock runs beside it as a
        cross-check (`wall` - the two agree to the microseconds an event record costs)."""

        def __init__(self):
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            barrier()
            self.e0.record()
            self.t0 = time.perf_counter()

        def stop(self):
            barrier()
            self.e1.record()
            self.e1.synchronize()
            self.wall = max_over_ranks(time.perf_counter() - self.t0)
            return max_over_ranks(self.e0.elapsed_time(self.e1) * 1e-3)

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=devname)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    dev = abncache.DeviceData(P, device=local_rank)
    acc = {"launches": 0, "algo_bytes": 0, "kernels": {}}

    def run_step(step, collect, batch=args.batch, control=ctrl, device_data=dev):
        """One step of the strong-scaling job: this rank scores its blocks of the global batch into a device
        slab, ONE all_gather puts every rank's scores in cache order on the device, one copy brings them home."""
        rows = batch_rows(total_rows, batch, args.block, step)
        mine = rows[shard.partition(len(rows), world, rank, args.block)]
        slab = shard.new_slab(len(rows), world, args.block, device=devname)
        device_data.score_to_slab(children[mine], parents[mine], control, slab)
        if collect:
            t = device_data.timing()
            acc["launches"] += t["launches"]
            acc["algo_bytes"] += t["algo_bytes"]
            for k in t["kernels"]:
                e = acc["kernels"].setdefault(k["name"], {"ms": 0.0, "bytes": 0, "fits": 0, "launches": 0, "fp64": 0})
                e["ms"] += k["ms"]; e["bytes"] += k["bytes"]; e["fits"] += k["fits"]; e["launches"] += 1; e["fp64"] += k["fp64"]
        return shard.gather_slab(slab, len(rows), world, args.block, pinned=True)

    for w in range(args.warmup):
        run_step(w, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    timer = DeviceTimer()
    status_counts = np.zeros(3, dtype=np.int64)
    for s in range(args.steps):
        last = run_step(args.warmup + s, True)
        status_counts += np.bincount(last["status"], minlength=3)[:3]
    elapsed = timer.stop()
    clocks = sampler.stop() if rank == 0 else None
    assert len(last["mlik"]) == args.batch
    value = args.batch * args.steps / elapsed
    launches = int(acc["launches"]) + 2 * args.steps * (1 if world > 1 else 0)  # + the gather's permute kernels

    # ---- the same job with the reference's own IRLS iteration (ABN_FLAG_EXACT_PATH) and with method = bayes --------
    extras = {}
    if not args.no_extras:
        def timed(batch, control, device_data, steps=2):
            run_step(10_000, False, batch, control, device_data)
            tm = DeviceTimer()
            passes = 0
            for s in range(steps):
                r = run_step(10_001 + s, False, batch, control, device_data)
                passes += int(r["iters"].sum())
            el = tm.stop()
            return {"value": batch * steps / el, "unit": "fits/s", "ms_per_step": 1e3 * el / steps, "steps": steps,
                    "batch_fits_per_step": batch, "mean_passes_per_fit": passes / (batch * steps)}
        if args.method == "mle":
            cx = abncache.to_c_control(abncache.build_control("mle", **{"exact.path": True}), "mle", d)
            cx.device = local_rank
            extras["exact_path"] = timed(args.side_batch, cx, dev)
            extras["exact_path"]["note"] = ("ABN_FLAG_EXACT_PATH: the reference's start values and stop rule pass for pass; "
                                            "default-path figure on the same batch size beside it")
            extras["default_path_same_batch"] = timed(args.side_batch, ctrl, dev)
            Pb = prep.prepare(data, dists, "bayes")
            cb = abncache.to_c_control(abncache.build_control("bayes"), "bayes", d)
            cb.device = local_rank
            with abncache.DeviceData(Pb, device=local_rank) as devb:
                extras["bayes"] = timed(args.side_batch, cb, devb)
            extras["bayes"]["note"] = "method='bayes' (Laplace mlik, GSL-hybrid iterate sequence) on the same data and cache"

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
            rows = batch_rows(total_rows, args.batch, args.block, step)
            mine = rows[shard.partition(len(rows), world, rank, args.block)]
            ds, k1 = capi.make_dataset(P2.X, P2.node_family, P2.node_col0)
            assert k1[0].ctypes.data == Xh.ctypes.data  # no hidden copy: the call reads the pinned buffer
            cd, k2 = capi.make_candidates(children[mine], parents[mine])
            res, arrs = capi.make_result(cd.rows, ctrl.method)
            err = C.create_string_buffer(512)
            capi.check(L.abn_b200_score_cache(C.byref(ds), C.byref(cd), C.byref(ctrl), C.byref(res), err, len(err)), err)
            full = shard.gather(arrs, len(rows), world, rank, args.block, device=devname)
            return full, cd.rows * (1 + parents.shape[1]) * 4, cd.rows * (4 * 8 + 2 * 4)

        one_shot(0)
        tm2 = DeviceTimer()
        e_steps = max(1, min(args.steps, 3))
        for s in range(e_steps):
            _, cand_bytes, res_bytes = one_shot(args.warmup + s)
        e_elapsed = tm2.stop()
        e2e = {"value": args.batch * e_steps / e_elapsed, "unit": "fits/s",
               "h2d_bytes_per_step": int(P.X.nbytes + cand_bytes), "d2h_bytes_per_step": int(res_bytes),
               "steps": e_steps, "ms_per_step": 1e3 * e_elapsed / e_steps,
               "call": "abn_b200_score_cache per rank (data upload + Gram + fits + download per step) + all_gather of the host results"}

    if rank != 0:
        dev.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- rooflines of the dominant kernel ---------------------------------------------------------------
    kernels = acc["kernels"]
    peak, peak_src = measured_peaks()
    fpeak, fpeak_src = fp64_peak()
    top_name, top = max(kernels.items(), key=lambda kv: kv[1]["ms"]) if kernels else (None, None)
    roofline = roofline_fp64 = None
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
        main_ms = sum(v["ms"] for k, v in kernels.items() if v["fits"] > 4736)  # kernels on the main stream (large buckets)
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "kernel": top_name, "peak_source": peak_src,
                    "bytes_per_launch": per_launch_bytes, "ms_per_launch": 1e3 * per_launch_s,
                    "fits_per_launch": top["fits"] / top["launches"],
                    "kernel_share_of_main_stream_time": top["ms"] / main_ms if main_ms else None,
                    "note": "SURVEY 8d's no-reuse streaming model; the columns are served from L2 (traffic << bytes), so this "
                            "fraction exceeds 1 and says nothing about headroom - roofline_fp64 is the bound that binds",
                    "all_kernels": {k: {"ms": v["ms"], "GBps": (v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["ms"] > 0 else None),
                                        "fits": v["fits"],
                                        "fp64_T_inst_per_s": (v["fp64"] / (v["ms"] / 1e3) / 1e12 if v["ms"] > 0 and v["fp64"] else None)}
                                    for k, v in sorted(kernels.items())},
                    "whole_step_GBps": acc["algo_bytes"] * world / elapsed / 1e9}
        if top["fp64"]:
            f_ach = top["fp64"] / (top["ms"] / 1e3) / 1e12
            step_fp64 = sum(v["fp64"] for v in kernels.values())
            roofline_fp64 = {"bound": "fp64 pipe", "achieved": f_ach, "peak": fpeak, "unit": "T FP64 lane-instructions/s",
                             "frac": f_ach / fpeak, "kernel": top_name, "peak_source": fpeak_src,
                             "fp64_inst_per_launch": top["fp64"] / top["launches"],
                             "whole_step_frac": step_fp64 / (elapsed / world) / 1e12 / fpeak / world if world == 1 else
                             step_fp64 / elapsed / 1e12 / fpeak,
                             "how": "per pass mode tallied on the device x rows x FP64 instructions per row counted from the "
                                    "SASS (abn_timing.kernel_fp64), / CUDA-event kernel time / measured DFMA rate; "
                                    "cross-check against smsp__inst_executed_pipe_fp64 in profiles/"}

    # ---- CPU baseline on a bounded sample of the same batch ----------------------------------------------
    cpu = None
    try:
        from oracle import pyoracle
        threads = host_threads()
        rows = batch_rows(total_rows, args.batch, args.block, args.warmup)
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

    config.update({"setup_s": setup_s,
                   "status_counts": {"ok": int(status_counts[0]), "not_converged": int(status_counts[1]),
                                     "non_finite": int(status_counts[2])},
                   "collective": "one all_gather of the 40-byte-per-fit result slabs inside the timed region" if world > 1
                                 else "single GPU: no collective"})
    line = {
        "metric": "score-cache GLM fits/sec", "value": value, "unit": "fits/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "timing": "CUDA events around the K steps (device idle at both ends), max over ranks",
        "wall_ms_per_step": 1e3 * timer.wall / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config, "roofline": roofline, "roofline_fp64": roofline_fp64, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": launches, "clocks": clocks,
    }
    line.update(extras)
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
