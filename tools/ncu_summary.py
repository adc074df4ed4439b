#!/usr/bin/env python
"""Prints the handful of ncu metrics this repo tracks from a .ncu-rep (run where ncu is installed).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_xu.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.per_cycle_active",
]


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        print(f"== {path}: {len(data)} launch(es)")
        name_i = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
        if name_i is not None:
            print("kernel:", data[0][name_i][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:72s} {units[i]:10s} {[r[i] for r in data]}")
        stall = [(h, hdr.index(h)) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        vals = sorted(((float(data[0][i] or 0), h) for h, i in stall), reverse=True)
        print("stall (warps per issue-active cycle):")
        for v, h in vals[:9]:
            print(f"   {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")


if __name__ == "__main__":
    main()
