#!/usr/bin/env python
"""tools/ncu_summary.py -- compact text summary of an .ncu-rep (run where ncu is installed; no GPU needed).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex] > profiles/x.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp16.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp16_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 else None
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in raw[2:]:
        name = r[idx["Kernel Name"]]
        if rx and not re.search(rx, name):
            continue
        print(f"== kernel: {name}   (launch id {r[0]})")
        for k in KEYS:
            if k in idx:
                print(f"   {k:70s} {r[idx[k]]:>16s} {units[idx[k]]}")
        stalls = [(h, float(r[i])) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        stalls.sort(key=lambda x: -x[1])
        print("   top stall reasons (warps stalled per issue-active cycle):")
        for h, v in stalls[:8]:
            print(f"      {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:6.2f}")
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + rx] if rx else [])))))
    hi = next((i for i, r in enumerate(src) if "Source" in r and "Instructions Executed" in r), None)
    if hi is None:
        return
    h = {n: i for i, n in enumerate(src[hi])}
    ops, samples = collections.Counter(), collections.Counter()
    total = 0
    for r in src[hi + 1:]:
        if len(r) <= h["Instructions Executed"] or not r[h["Instructions Executed"]].isdigit():
            continue
        t = r[h["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";")
        op = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("IMAD", "MOV", "LDS", "STG", "LDG", "STS") else op.split(".")[0]
        n = int(r[h["Instructions Executed"]])
        ops[op] += n
        total += n
        if "# Samples" in h and r[h["# Samples"]].isdigit():
            samples[op] += int(r[h["# Samples"]])
    print(f"   dynamic warp-instruction mix (source page, first matching launch): total {total}")
    ssum = sum(samples.values()) or 1
    for op, n in ops.most_common(28):
        print(f"      {op:14s} {n:12d} {100.0 * n / total:5.1f}%   stall samples {100.0 * samples[op] / ssum:5.1f}%")


if __name__ == "__main__":
    main()
