"""Opcode histogram of one kernel from an ncu SASS source page (executed warp-instructions per opcode, and per thread).
Usage:  ncu -i X.ncu-rep --page source --csv > src.csv ;  python profiles/tools/op_hist.py src.csv <threads>
`threads` = number of threads that did work in the launch (instances for one-thread-per-instance kernels)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
threads = float(sys.argv[2])
ops = collections.Counter()
tot = 0
isrc = ie = None
for r in rows:
    if r and r[0] == "Address":
        isrc, ie = r.index("Source"), r.index("Instructions Executed")
        continue
    if ie is None or len(r) <= ie:
        continue
    try:
        n = int(r[ie])
    except ValueError:
        continue
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ops[op.split(".")[0]] += n
    tot += n
print(f"total warp-instructions {tot}  ({tot * 32 / threads:.0f} per thread)")
for k, v in ops.most_common(30):
    print(f"{k:10s} {v:12d} {100 * v / tot:5.1f}%  per-thread {v * 32 / threads:7.1f}")
f64 = {k: ops[k] * 32 / threads for k in ("DFMA", "DMUL", "DADD", "DSETP")}
print("fp64-pipe instructions per thread:", {k: round(v, 1) for k, v in f64.items()}, "sum", round(sum(f64.values()), 1))
print("fp64 flops per thread (DFMA = 2):", round(2 * f64["DFMA"] + f64["DMUL"] + f64["DADD"], 1))
