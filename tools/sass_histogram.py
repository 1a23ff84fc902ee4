#!/usr/bin/env python3
"""Opcode histogram of the shipped kernels from `cuobjdump -sass` (runs on the CPU box, no GPU needed):

    python tools/sass_histogram.py [kernel-name regex] > profiles/rNN_sass_histogram.txt

For every kernel whose demangled name matches the regex (default: the fused step k_tile<1,1,0,0>) prints the instruction
count, registers (from `cuobjdump -res-usage`), the opcode histogram with the issue pipe each opcode was measured on
(tools/ubench/pipes.cu) and the memory instructions with their cache modifiers -- the static counterpart of the
`smsp__inst_executed` figures in the ncu summaries beside it."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "livevideo_b200", "liblivevideo.so")
# pipe per opcode as measured by tools/ubench/pipes.cu (profiles/ubench_pipes.txt); anything else is listed as "-"
PIPE = {"PRMT": "alu", "SHF": "alu", "LOP3": "alu", "VIADDMNMX": "alu", "VIMNMX": "alu", "VABSDIFF4": "alu", "IADD3": "alu",
        "IADD": "alu", "ISETP": "alu", "SEL": "alu", "LEA": "alu", "POPC": "xu", "FLO": "xu", "BREV": "xu", "MOV": "alu",
        "IMAD": "fma", "IDP": "fma", "LDG": "lsu", "STG": "lsu", "LDS": "lsu", "STS": "lsu", "ATOMG": "lsu", "RED": "lsu",
        "LDC": "lsu/const", "LDCU": "uniform", "S2R": "xu", "S2UR": "uniform", "SHFL": "lsu", "REDUX": "alu", "BAR": "cbu",
        "BRA": "cbu", "EXIT": "cbu", "BSSY": "cbu", "BSYNC": "cbu"}


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else r"k_tile<true, true, false, false>")
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", ln)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    funcs = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            funcs[cur].append(m.group(2))
    names = demangle(list(funcs))
    print(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)}  (sm_100a), kernels matching /{pat.pattern}/")
    for f, ops in funcs.items():
        d = names.get(f, f)
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        if not pat.search(d):
            continue
        r = regs.get(f, ("?", "?"))
        print(f"\n== {d.split('(')[0]}: {len(ops)} SASS instructions, {r[0]} registers, {r[1]} B static shared memory")
        base = collections.Counter(o.split(".")[0] for o in ops)
        pipes = collections.Counter()
        for o, n in base.items():
            pipes[PIPE.get(o, "-")] += n
        print("   by pipe: " + ", ".join(f"{p} {n}" for p, n in pipes.most_common()))
        print("   by opcode: " + ", ".join(f"{o} {n}" for o, n in base.most_common()))
        mem = collections.Counter(o for o in ops if o.split(".")[0] in ("LDG", "STG", "LDS", "STS", "ATOMG", "RED", "UTMALDG", "UTMASTG", "SYNCS", "SHFL", "REDUX"))
        print("   memory / exchange: " + ", ".join(f"{o} x{n}" for o, n in mem.most_common()))
        simd = collections.Counter(o for o in ops if o.split(".")[0] in ("VIADDMNMX", "VIMNMX", "VABSDIFF4", "IDP", "PRMT"))
        print("   integer SIMD: " + ", ".join(f"{o} x{n}" for o, n in simd.most_common()))


if __name__ == "__main__":
    main()
