#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): key metrics per launch and the
hottest source lines.  Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--src N]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nsrc = int(sys.argv[sys.argv.index("--src") + 1]) if "--src" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        name = d[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("==", name[:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:72s} {d[i]:>16s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(d[i]), h.replace("smsp__average_warps_issue_stalled_", "")
                                   .replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
    if nsrc <= 0:
        return
    src = run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    rows = list(csv.reader(io.StringIO(src)))
    cur_file, hdr, lines = "?", None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] in ("Function Name", "Kernel Name") or not r[0].strip().isdigit():
            continue
        try:
            i_s, i_n = hdr.index("# Samples"), hdr.index("Instructions Executed")
            lines.append((int(float(r[i_s] or 0)), int(float(r[i_n] or 0)), cur_file, r[0], r[1]))
        except (ValueError, IndexError):
            continue
    tot = sum(l[0] for l in lines) or 1
    toti = sum(l[1] for l in lines) or 1
    print(f"== hottest source lines (of {tot} samples, {toti} warp-instructions)")
    for smp, n, f, ln, text in sorted(lines, reverse=True)[:nsrc]:
        print(f"  {100.0 * smp / tot:5.1f}% smp {100.0 * n / toti:5.1f}% inst  {f}:{ln:>4s}: {text.strip()[:100]}")


if __name__ == "__main__":
    main()
