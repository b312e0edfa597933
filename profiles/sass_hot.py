"""Executed-instruction breakdown of one kernel from `ncu --page source --csv` (SASS view):
opcode mix and the hottest address ranges.  usage: sass_hot.py source.csv [kernel_index] [bucket]"""
import collections
import csv
import sys

csv.field_size_limit(10**9)
path = sys.argv[1]
want = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 100
kern, ops, rows, name = 0, collections.Counter(), [], ""
for row in csv.reader(open(path)):
    if len(row) >= 2 and row[0] == "Kernel Name":
        kern += 1
        if kern == want:
            name = row[1]
        if kern > want:
            break
        continue
    if kern != want or len(row) < 6 or row[0] == "Address":
        continue
    sass, ex, samp = row[1].strip(), int(row[5]), int(row[4])
    tok = sass.split()
    op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
    ops[op] += ex
    rows.append((ex, samp, sass))
total = sum(ops.values())
print(name)
print("warp-instructions executed", total, "| SASS lines", len(rows))
print(" ".join(f"{op}:{c / total * 100:.1f}%" for op, c in ops.most_common(22)))
for i in range(0, len(rows), bucket):
    chunk = rows[i:i + bucket]
    e, s = sum(r[0] for r in chunk), sum(r[1] for r in chunk)
    if e / total > 0.01:
        print(f"{i:6d} {e / total * 100:5.1f}% exec {s:7d} samples   {chunk[0][2][:50]}")
