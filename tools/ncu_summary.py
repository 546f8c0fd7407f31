#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the text form kept under profiles/: one block per launch with the
metrics the DESIGN / bench.py arguments rest on (durations, pipe utilisation, DRAM / L2 / shared-memory traffic,
stall reasons). usage: tools/ncu_summary.py <file.ncu-rep> > profiles/ncu_<tag>_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "Kernel Name", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        print("----")
        for w in WANT:
            if w in idx:
                print(f"{w:72s} {r[idx[w]]}  [{units[idx[w]]}]")
        st = sorted(((float(r[idx[h]]), h) for h in stalls), reverse=True)[:8]
        for v, h in st:
            print(f"{h:72s} {v:.3f}  [warps per issue]")


if __name__ == "__main__":
    main()
