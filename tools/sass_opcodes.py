#!/usr/bin/env python3
"""Opcode counts per kernel of the built library: python tools/sass_opcodes.py > profiles/<round>_sass_opcodes.txt
(cuobjdump -sass + c++filt; runs without a GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mir_prefer_b200", "libmirfold.so")
COLS = [("S16x2", r"^VIADDMNMX\.S16x2"), ("VIADDMNMX", r"^VIADDMNMX(?!\.S16x2)"), ("VIMNMX", r"^VIMNMX"), ("CREDUX", r"^CREDUX|^REDUX"),
        ("SHFL", r"^SHFL"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDG", r"^LDG"), ("STG", r"^STG"), ("ATOM*", r"^ATOM|^RED"),
        ("UBLKCP", r"^UBLKCP"), ("FENCE.A", r"^FENCE\.VIEW\.ASYNC|^FENCE.*ASYNC"), ("UTMA*", r"^UTMA"), ("LDGSTS", r"^LDGSTS"), ("BAR", r"^BAR"),
        ("PRMT", r"^PRMT")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m and cur is not None:
            cur["instr"] += 1
            for name, pat in COLS:
                if re.match(pat, m.group(1)):
                    cur[name] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass %s (sm_100a): opcode counts per kernel (tools/sass_opcodes.py)" % os.path.relpath(LIB, ROOT))
    print("# S16x2 = VIADDMNMX.S16x2 (two min-plus terms per instruction); CREDUX = redux.sync; UBLKCP = cp.async.bulk (TMA engine, 1-D bulk copy);")
    print("# FENCE.A = fence.proxy.async; UTMA* = UTMACMDFLUSH, the bulk-group commit (no tensor-map UTMALDG / UTMASTG: no boxes to describe, DESIGN.md section 3);")
    print("# no LDGSTS, no tcgen05 (integer min-plus).  k_fill_s16<.., true> = the instantiation with the int32-strip switch, launched for spans >= 400")
    print("%-62s %8s" % ("kernel", "instr") + "".join(" %9s" % c for c, _ in COLS))
    for (mangled, cnt), name in zip(kernels.items(), names):
        name = re.sub(r"\(.*$", "", name)
        print("%-62s %8d" % (name[:62], cnt["instr"]) + "".join(" %9d" % cnt[c] for c, _ in COLS))


if __name__ == "__main__":
    main()
