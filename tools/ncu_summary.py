#!/usr/bin/env python
"""Condense an Nsight Compute report (.ncu-rep, brought back in gpurun_out/) into the small
tracked files under profiles/: key raw metrics and the stall-sample split per barrier-delimited
segment of the kernel.  Usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/NAME"""
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active"]


def ncu(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst):
    raw = ncu(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = dict(zip(hdr, zip(units, vals)))
    lines = [f"# ncu summary of `{rep.split('/')[-1]}`", "", f"kernel: `{d['Kernel Name'][1]}`", "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in d:
            lines.append(f"| {k} | {d[k][1]} | {d[k][0]} |")
    lines += ["", "Warp-issue stall reasons (`smsp__average_warps_issue_stalled_*_per_issue_active`):", "",
              "| reason | warps stalled per issue |", "|---|---|"]
    for h in hdr:
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
        if m and float(d[h][1] or 0) > 0.01:
            lines.append(f"| {m.group(1)} | {float(d[h][1]):.3f} |")
    src = ncu(rep, "source")
    sh = src[1]
    rows = [dict(zip(sh, r)) for r in src[2:] if len(r) == len(sh)]
    tot = sum(int(r["# Samples"]) for r in rows) or 1
    bars = [i for i, r in enumerate(rows) if "BAR.SYNC" in r["Source"]]
    lines += ["", "Stall samples per barrier-delimited SASS segment (segments under 1 % omitted):", "",
              "| SASS instr range | share of samples | warp instructions executed | dominant opcodes |", "|---|---|---|---|"]
    prev = 0
    for b in bars + [len(rows) - 1]:
        seg = rows[prev:b + 1]
        s = sum(int(r["# Samples"]) for r in seg)
        if s / tot >= 0.01:
            ie = sum(int(r["Instructions Executed"]) for r in seg)
            ops = {}
            for r in seg:
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r["Source"])
                if m:
                    ops[m.group(2)] = ops.get(m.group(2), 0) + int(r["Instructions Executed"])
            top = ", ".join(k for k, _ in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
            lines.append(f"| {prev}..{b} | {100 * s / tot:.1f} % | {ie / 1e9:.2f} G | {top} |")
        prev = b + 1
    open(dst + ".md", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
