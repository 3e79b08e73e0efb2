"""Summarise an `ncu --page source --csv` export: instructions executed and stall samples by opcode."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
S, IE, SM = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ops = collections.Counter()
smp = collections.Counter()
tot = 0
body = rows[2:]
for r in body:
    if len(r) <= IE:
        continue
    op = r[S].split()[0]
    if op.startswith("@"):
        op = r[S].split()[1]
    op = op.split(".")[0]
    n = int(r[IE] or 0)
    ops[op] += n
    smp[op] += int(r[SM] or 0)
    tot += n
print("total warp-instructions", tot)
for op, n in ops.most_common(25):
    print(f"{op:10s} {n:14d} {100.0 * n / tot:6.2f}%   samples {smp[op]}")
if len(sys.argv) > 2:
    win = int(sys.argv[2])
    best = sorted(range(len(body)), key=lambda i: -int(body[i][SM] or 0))[:win]
    for i in sorted(best):
        print(i, body[i][S][:70], body[i][IE], body[i][SM])
