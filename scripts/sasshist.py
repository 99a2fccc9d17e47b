#!/usr/bin/env python
"""Per-source-line SASS instruction histogram of one kernel section from `nvdisasm -g -c` output (dev tool)."""
import re, sys, collections
path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "decode_persistent.cuh"
cur = None
hist = collections.Counter()
total = 0
ops = collections.defaultdict(collections.Counter)
for ln in open(path):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
    if m and cur:
        hist[cur] += 1; total += 1
        ops[cur][m.group(3).split(".")[0]] += 1
print("total", total)
byfile = collections.Counter()
for (f, l), c in hist.items(): byfile[f] += c
print(byfile.most_common())
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
acc = 0
for (f, l), c in sorted(hist.items()):
    if f == want and lo <= l <= hi:
        acc += c
        print(f"{l:5d} {c:5d}  " + " ".join(f"{k}:{v}" for k, v in ops[(f, l)].most_common(6)))
print("sum in range", acc)
