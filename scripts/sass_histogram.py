#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libwdb200.so -> profiles/sass_opcodes.txt.

    python scripts/sass_histogram.py [out.txt]

Counts the mnemonics that prove which hardware paths a kernel uses (B200_PROFILING.md):
tcgen05 (UTCHMMA/UTCQMMA..., LDTM/STTM, UTCBAR, UTCCP), TMA (UBLKCP = 1-D bulk copy,
UTMALDG/UTMASTG = tensor-map copies), mbarrier (SYNCS), packed f32x2 math
(FFMA2/FADD2/FMUL2), cluster / DSMEM (UCGABAR*, LDS/STS/ATOMS with .CLUSTER? are not
distinguishable by mnemonic: mapa / ld.shared::cluster appear as MAPA? not emitted -> we
list ACQBULK / UCGABAR_ARV / UCGABAR_WAIT and MEMBAR), plus the instruction total."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "warp_drive_b200", "libwdb200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTCCP",
         "UTCATOMSWS", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "FFMA2",
         "FADD2", "FMUL2", "UCGABAR_ARV", "UCGABAR_WAIT", "CCTL", "REDUX", "SHFL", "MUFU",
         "DFMA", "DMUL", "DADD", "BAR", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED",
         "LDL", "STL"]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles",
                                                                  "sass_opcodes.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True,
                          check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    op_re = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = op_re.match(line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["__total__"] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True)
    names = demangle.stdout.splitlines() if demangle.returncode == 0 else list(kernels)
    lines = ["# cuobjdump -sass warp_drive_b200/libwdb200.so: per-kernel SASS opcode counts",
             "# (static instruction counts; nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)",
             ""]
    for (mangled, cnt), name in zip(kernels.items(), names):
        short = re.sub(r"\(.*", "", name)
        short = short.replace("(anonymous namespace)::", "").replace("void ", "")
        hits = [f"{op} x{cnt[op]}" for op in WATCH if cnt.get(op)]
        lines.append(f"{short}  [{cnt['__total__']} instructions]")
        lines.append("    " + (", ".join(hits) if hits else "-"))
    with open(out_path, "w") as fp:
        fp.write("\n".join(lines) + "\n")
    print(out_path)


if __name__ == "__main__":
    main()
