#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show a Blackwell-native kernel (B200_PROFILING.md "What proves a Blackwell-native
kernel"): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG / UBLKCP = TMA tensor loads / stores /
reduce-add stores / 1-D bulk copies, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be absent).  No GPU needed:

    python tools/sass_evidence.py > profiles/r01_sass_evidence.csv
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "slsforasvspoof-2021-df_b200", "libslsb200.so")
COLS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "UTCBAR", "MUFU.TANH", "MUFU.EX2", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    pat = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|SYNCS|UTCBAR|MUFU\.TANH|MUFU\.EX2|HMMA)\b")
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
        elif cur:
            for t in pat.findall(line):
                counts[cur]["UTCHMMA" if t.startswith("UTC") and t.endswith("MMA") else t] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("kernel," + ",".join(COLS))
    for name, c in zip(names, counts.values()):
        name = name.replace("(anonymous namespace)::", "").replace("slsb::", "")
        name = re.sub(r"^void\s+", "", name)
        depth, cut = 0, len(name)
        for i, ch in enumerate(name):                      # cut the argument list, keep template arguments
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = i
                break
        print('"%s",' % name[:cut] + ",".join(str(c.get(k, 0)) for k in COLS))


if __name__ == "__main__":
    sys.exit(main())
