#!/usr/bin/env python
"""Static SASS opcode histogram per kernel of the built library.
usage: cuobjdump -sass matchy_b200/lib/libmatchy_b200.so | python profiles/sass_hist.py > profiles/sass_r2.txt
Per kernel: instruction count, the 14 most frequent opcodes, and the counts of the opcodes that show which hardware path a
kernel uses (bulk async copies UBLKCP, mbarrier SYNCS, warp reductions REDUX, votes VOTE, shuffles SHFL, named barriers BAR)."""
import collections
import re
import subprocess
import sys

WATCH = ["UBLKCP", "SYNCS", "REDUX", "VOTE", "VOTEU", "SHFL", "BAR", "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "ATOMS", "MATCH", "IMAD", "LOP3", "PRMT", "SHF", "IADD3", "ISETP", "BRA", "BSSY", "BSYNC", "WARPSYNC"]
cur = None
hist = collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        try:
            name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        except OSError:
            pass
        cur = hist.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
for name, c in hist.items():
    total = sum(c.values())
    print("%s: %d instructions" % (name, total))
    print("  top: " + ", ".join("%s %d" % kv for kv in c.most_common(14)))
    print("  watch: " + ", ".join("%s %d" % (k, c[k]) for k in WATCH if c[k]))
