#!/usr/bin/env python
"""Instruction/sample share of source-line ranges of wdb_tag_continuous.cu in an .ncu-rep.
Usage: python scripts/ncu_sections.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
hdr, infile, out = None, None, []
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        infile = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].strip().isdigit():
        continue
    def num(name):
        try:
            return int(float(r[hdr.index(name)]))
        except (ValueError, IndexError):
            return 0
    out.append((infile, int(r[0]), num("Instructions Executed"), num("# Samples")))
tot_i = sum(o[2] for o in out) or 1
tot_s = sum(o[3] for o in out) or 1
src = open("warp_drive_b200/csrc/wdb_tag_continuous.cu").read().split("\n")

def find(s0, start=0):
    for i in range(start, len(src)):
        if s0 in src[i]:
            return i + 1
    raise KeyError(s0)

marks = [
    ("exact_select_warp", "__device__ __noinline__ int exact_select_warp", "// ---------------------------------------------------------------- sorting networks"),
    ("sampling prologue", "// block offsets inside the tile (floats)", "  int alive = 0;"),
    ("kinematics+stage", "  int alive = 0;", "  // tagger id list in id order"),
    ("history path", "        if (P.use_history) {", "        if (!have) {"),
    ("network path", "        if (!have) {", "        }   // !have"),
    ("verification", "        // ---- verification on EXACT", "    // exact path: the warp resolves"),
    ("exact loop", "    // exact path: the warp resolves", "      float *orow = P.stage_obs"),
    ("obs features", "      float *orow = P.stage_obs", "    // full observation (:55-113)"),
    ("rewards+push+copy-out+reset", "  // ------------------------------------------------------------------ rewards / tags", "struct LaunchPlan"),
]
print(f"total: {tot_i} warp-instructions, {tot_s} samples")
for label, a, b in marks:
    la, lb = find(a), find(b)
    n = sum(o[2] for o in out if o[0] == "wdb_tag_continuous.cu" and la <= o[1] < lb)
    s = sum(o[3] for o in out if o[0] == "wdb_tag_continuous.cu" and la <= o[1] < lb)
    print(f"  {label:30s} L{la:4d}-{lb:4d}  inst {100 * n / tot_i:5.1f}%  samples {100 * s / tot_s:5.1f}%")
other = {}
for f, ln, n, s in out:
    if f != "wdb_tag_continuous.cu":
        o = other.setdefault(f, [0, 0]); o[0] += n; o[1] += s
for f, (n, s) in sorted(other.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"  [{f:28s}]            inst {100 * n / tot_i:5.1f}%  samples {100 * s / tot_s:5.1f}%")
