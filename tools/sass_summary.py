#!/usr/bin/env python
"""profiles/sass_rNN.md: per-kernel SASS opcode summary of libsdvar_b200.so (cuobjdump -sass), so a reader can check without a
disassembler that the hot kernels are tcgen05 / TMA / packed-fp32 code.    python tools/sass_summary.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sdvar_b200", "lib", "libsdvar_b200.so")
KEY = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "FFMA2", "FADD2", "FMUL2", "MUFU.EX2",
       "ATOMS", "REDUX", "LDG.E.128", "STG.E.128", "HMMA", "IMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_r02.md")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    dm = demangle(list(kernels))
    lines = ["# SASS opcode summary of libsdvar_b200.so (sm_100a)", "",
             "`cuobjdump -sass sdvar_b200/lib/libsdvar_b200.so`, static instruction counts per kernel (tools/sass_summary.py).",
             "tcgen05 shows up as `UTCHMMA` (`.2CTA` = cta_group::2), TMA as `UTMALDG` / `UTMASTG` / `UBLKCP`, tensor-memory access as",
             "`LDTM` / `STTM`, packed fp32 as `FFMA2` / `FADD2` / `FMUL2`.  No `HMMA` / `IMMA` (mma.sync) anywhere.", ""]
    tot = collections.Counter()
    for k, c in kernels.items():
        for op, n in c.items():
            tot[op] += n
    agg = lambda c, key: sum(n for op, n in c.items() if op == key or op.startswith(key + "."))
    lines += ["## Whole library", "", "| mnemonic | count |", "|---|---|"]
    for key in KEY:
        lines.append(f"| {key} | {agg(tot, key)} |")
    lines += ["", "## Per kernel (kernels with at least one of the mnemonics above, or > 500 instructions)", "",
              "| kernel | instr | UTCHMMA (.2CTA) | UTMALDG | UTMASTG | UBLKCP | LDTM | STTM | FFMA2 | MUFU.EX2 | ATOMS | REDUX | top opcodes |",
              "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for k, c in kernels.items():
        n = sum(c.values())
        hits = sum(agg(c, key) for key in KEY[:12])
        if hits == 0 and n < 500:
            continue
        name = re.sub(r"\(.*", "", dm.get(k, k)).replace("void ", "").replace("sdvar::", "")
        top = ", ".join(f"{op} {v}" for op, v in c.most_common(5))
        lines.append(f"| `{name}` | {n} | {agg(c, 'UTCHMMA')} ({agg(c, 'UTCHMMA.2CTA')}) | {agg(c, 'UTMALDG')} | {agg(c, 'UTMASTG')} | {agg(c, 'UBLKCP')} | "
                     f"{agg(c, 'LDTM')} | {agg(c, 'STTM')} | {agg(c, 'FFMA2')} | {agg(c, 'MUFU.EX2')} | {agg(c, 'ATOMS')} | {agg(c, 'REDUX')} | {top} |")
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("wrote", out_path, len(kernels), "kernels")


if __name__ == "__main__":
    main()
