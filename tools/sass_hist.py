#!/usr/bin/env python
"""Opcode histogram of the hot kernels' SASS (cuobjdump -sass of the built library; no GPU needed).

usage: python tools/sass_hist.py [out.txt]

For every kernel whose name matches HOT it prints the instruction count, the share of each opcode family and the
mnemonics that prove what the kernel is made of (POPC for the distance, UBLKCP / SYNCS for TMA bulk copies and mbarriers,
MATCH / REDUX / VOTE for the warp-aggregated queues, BAR for CTA barriers, LDL / STL for spills).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "enclone_ranger_b200", "libenclone_join_b200.so")
HOT = re.compile(r"k_stage1|k_stage2a|k_stage2b|k_pack_rows|k_forest_round|k_truncate|k_p1_|k_quot_table|k_sr_table")
WATCH = ["POPC", "LOP3", "UBLKCP", "SYNCS", "MATCH", "REDUX", "VOTE", "SHFL", "BAR", "LDL", "STL", "ATOM", "ATOMS", "RED",
         "LDG", "STG", "LDS", "STS", "DADD", "DMUL", "DFMA", "CCTL"]


def main():
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kern, ops = None, collections.defaultdict(collections.Counter)
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and kern:
            ops[kern][m.group(1)] += 1
    names = sorted(ops)
    dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for a, b in zip(names, dm):
        demangle[a] = re.sub(r"\(.*", "", b)
    print("SASS opcode histogram, sm_100a, %s" % os.path.relpath(LIB, ROOT), file=out)
    for k in names:
        if not HOT.search(demangle[k]):
            continue
        c = ops[k]
        tot = sum(c.values())
        print("\n%s: %d instructions" % (demangle[k], tot), file=out)
        print("  watched: " + "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w]), file=out)
        print("  top:     " + "  ".join("%s %d (%.0f%%)" % (o, n, 100.0 * n / tot) for o, n in c.most_common(12)), file=out)


if __name__ == "__main__":
    main()
