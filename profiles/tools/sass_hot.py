"""Top stalled SASS instructions of a capture: ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv;
python profiles/tools/sass_hot.py sass.csv [top].  Prints stall samples, executed count, address offset, instruction."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Address")
start = rows.index(hdr)
isrc, ist, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
out, tot, base = [], 0, None
for n, r in enumerate(rows[start + 1:]):
    if len(r) <= iex:
        continue
    try:
        s, e = int(r[ist]), int(r[iex])
    except ValueError:
        continue
    addr = int(r[0], 16)
    base = addr if base is None else base
    tot += s
    out.append((s, e, addr - base, r[isrc].strip()[:100]))
print("total samples", tot, "instructions executed", sum(o[1] for o in out))
for s, e, a, src in sorted(out, reverse=True)[:top]:
    print(f"{s:6d} {100 * s / max(tot, 1):5.1f}%  ex={e:9d}  +{a:05x}  {src}")
