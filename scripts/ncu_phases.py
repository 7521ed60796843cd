#!/usr/bin/env python
"""Aggregate the source page of an .ncu-rep of tag_continuous_kernel by kernel phase (line
ranges of wdb_tag_continuous.cu): share of warp-state samples, share of executed
warp-instructions and the dominant stall reasons of each phase, then the lines where warps
wait at barriers.  Usage: python scripts/ncu_phases.py gpurun_out/x.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

PHASES_V2 = [(0, 60, "sample_row_regs (register CDF + mask search)"), (60, 232, "prologue + sampling"),
             (232, 280, "kinematics"), (280, 530, "selection (history scan, sort, verify, exact)"),
             (530, 560, "kk / ids store + bookkeeping loads"), (560, 600, "rewards"),
             (600, 650, "finalize rewards / done / pushes"), (650, 700, "feature chunks"),
             (700, 770, "chunk copy-out"), (770, 100000, "reset")]
PHASES = [(0, 53, "file-level helpers"), (53, 176, "setup + probability TMA requests"),
          (176, 290, "prologue loads (state, ids, tables)"),
          (290, 365, "sampling + kinematics"),
          (365, 427, "obs setup / key planes"),
          (427, 546, "history scan + extract + sort (+ network fallback)"),
          (546, 660, "verification / exact path"), (660, 760, "features"),
          (760, 810, "barrier + TMA store issue + bookkeeping loads"), (810, 852, "rewards"),
          (852, 948, "barrier + push / bookkeeping"), (948, 100000, "reset")]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv",
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    hdr, cur, idx, data = None, "?", {}, []
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr, idx = r, {h: i for i, h in enumerate(r)}
        elif hdr is not None and r[0].strip().isdigit():
            data.append((cur, int(r[0]), r))

    def val(r, k):
        try:
            return float(r[idx[k]] or 0)
        except (ValueError, KeyError, IndexError):
            return 0.0

    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(collections.Counter)
    tot_s = tot_i = 0.0
    waits = []
    for cur, ln, r in data:
        s, n = val(r, "# Samples"), val(r, "Instructions Executed")
        tot_s += s
        tot_i += n
        if cur == "wdb_tag_continuous.cu":
            key = next(name for a, b, name in PHASES if a <= ln < b)
        elif cur == "wdb_tc_small_v2.cu":
            key = "v2: " + next(name for a, b, name in PHASES_V2 if a <= ln < b)
        else:
            key = cur
        agg[key]["smp"] += s
        agg[key]["inst"] += n
        for st in stalls:
            agg[key][st] += val(r, st)
        w = val(r, "stall_barrier") + val(r, "stall_sleep")
        if w >= 20:
            waits.append((w, cur, ln, r[1].strip()[:70]))
    print(f"total: {tot_s:.0f} warp-state samples, {tot_i / 1e6:.1f} M warp-instructions")
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["smp"]):
        if c["smp"] < 0.004 * tot_s:
            continue
        top = sorted(((c[s], s) for s in stalls), reverse=True)[:4]
        print(f"  {key:52s} samples {100 * c['smp'] / tot_s:5.1f} %   instructions "
              f"{100 * c['inst'] / tot_i:5.1f} %   " +
              ", ".join(f"{s[6:]} {100 * v / max(c['smp'], 1):.0f}%" for v, s in top))
    print("lines where warps wait for other warps (barrier + sleep samples):")
    for w, cur, ln, text in sorted(waits, reverse=True):
        print(f"  {100 * w / tot_s:5.1f} %  {cur}:{ln}  {text}")


if __name__ == "__main__":
    main()
