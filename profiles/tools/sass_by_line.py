"""Join an ncu SASS source page (per-address executed-instruction counts) with nvdisasm -g line info and print the
dynamic instruction count per CUDA source line / opcode.  Usage:
    ncu -i X.ncu-rep --page source --csv > src.csv ; nvdisasm -g -c uav.sm_100a.cubin > dis.txt
    python profiles/tools/sass_by_line.py src.csv dis.txt <mangled-substring> <demangled-substring> [top]
"""
import collections
import csv
import re
import sys

src_csv, dis_txt, kern, ncu_kern = sys.argv[1:5]  # mangled substring (nvdisasm), demangled substring (ncu)
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

# address -> (file, line) for the kernel's section
addr_line = {}
cur = None
insec = False
for ln in open(dis_txt, errors="replace"):
    if ln.startswith("//---") and ".text." in ln:
        insec = kern in ln
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m:
        addr_line[int(m.group(1), 16)] = cur

rows = list(csv.reader(open(src_csv)))
sec = None
counts = collections.Counter()
ops_by_line = collections.defaultdict(collections.Counter)
tot = 0
first = True
for r in rows:
    if r and r[0] == "Kernel Name":
        sec = r[1]
        take = ncu_kern in sec and first
        if ncu_kern in sec:
            first = False
        base_addr = None
        continue
    if r and r[0] == "Address":
        hdr = r
        ia, isrc, ie = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
        continue
    if not sec or not take or len(r) <= ie:
        continue
    try:
        n = int(r[ie])
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except ValueError:
        continue
    if base_addr is None:
        base_addr = a
    key = addr_line.get(a - base_addr) or ("?", 0)
    toks = r[isrc].split()
    op = (toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")).split(".")[0]
    counts[key] += n
    ops_by_line[key][op] += n
    tot += n
print("total warp-instructions", tot)
for key, n in counts.most_common(top):
    print(f"{key[0]}:{key[1]:<5d} {n:>12d} {100.0*n/tot:5.1f}%  ", dict(ops_by_line[key].most_common(5)))
