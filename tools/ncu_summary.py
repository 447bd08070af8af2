#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline metrics + top stalled SASS lines."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__cluster_size" ]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main(path, top=25):
    rows = list(csv.reader(run([path, "--page", "raw", "--csv"]).splitlines()))
    hdr, vals = rows[0], rows[2]
    units = rows[1]
    d = dict(zip(hdr, zip(units, vals)))
    for k in KEYS:
        if k in d:
            print(f"{k:75s} {d[k][1]:>16s} {d[k][0]}")
    stalls = {h: float(v[1]) for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    tot = sum(stalls.values())
    print("stall reasons (warps per issue slot, share):")
    for h, v in sorted(stalls.items(), key=lambda x: -x[1])[:10]:
        print(f"   {h.split('stalled_')[1].replace('_per_issue_active.ratio',''):28s} {v:6.2f}  {100*v/tot:5.1f}%")
    src = list(csv.reader(run([path, "--page", "source", "--csv"]).splitlines()))
    h2 = src[1]
    data = src[2:]
    iS = h2.index("Warp Stall Sampling (All Samples)")
    iE = h2.index("Instructions Executed")
    total = sum(int(r[iS]) for r in data)
    print("total stall samples", total, " static instrs", len(data))
    cols = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
    for idx, r in sorted(sorted(enumerate(data), key=lambda x: -int(x[1][iS]))[:top]):
        reasons = sorted(((h2[i], int(r[i])) for i in cols if r[i] not in ("", "0")), key=lambda x: -x[1])[:2]
        print(f"{idx:5d} {r[1].strip()[:64]:64s} {int(r[iS]):6d} {100*int(r[iS])/total:5.1f}% exec={r[iE]:>9s} {reasons}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
