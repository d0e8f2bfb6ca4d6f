"""Dynamic instructions and stall samples per CUDA source line from `ncu -i X --page source --csv --print-source sass,cuda`.
Usage: python profiles/tools/lines.py src_cuda.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, hdr = None, None
per, samples, text = collections.Counter(), collections.Counter(), {}
warps = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iex, ist = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr and r[0].isdigit():
        try:
            key = (cur, int(r[0]))
            per[key] += int(r[iex])
            samples[key] += int(r[ist])
            text[key] = r[1].strip()[:100]
        except ValueError:
            pass
    elif hdr and r[0] == "" and warps is None and len(r) > iex:
        try:
            warps = int(r[iex]) or None  # first SASS row: executed once per warp
        except ValueError:
            pass
warps = warps or 1
tot, ts = sum(per.values()), max(1, sum(samples.values()))
byfile = collections.Counter()
for k, v in per.items():
    byfile[k[0]] += v
print(f"instructions per warp: {tot / warps:.0f}; by file: " + ", ".join(f"{k} {v / warps:.0f}" for k, v in byfile.most_common()))
for k, v in per.most_common(top):
    print(f"{k[0]}:{k[1]:<4d} {v / warps:7.1f} instr {100 * samples[k] / ts:5.1f} % of stall samples   {text[k]}")
