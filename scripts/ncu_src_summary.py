"""Summarise an `ncu --page source --csv` export: opcode mix (executed instructions, stall samples) and stall reasons.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> > src.csv ; python scripts/ncu_src_summary.py src.csv"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
iS, iI, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot = sum(int(r[iS]) for r in data)
totI = sum(int(r[iI]) for r in data)
print("stall samples", tot, "warp instructions", totI, "SASS lines", len(data))
ci, cs = Counter(), Counter()
for r in data:
    t = r[isrc].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    ci[op] += int(r[iI])
    cs[op] += int(r[iS])
for op, v in ci.most_common(18):
    print(f"{op:10s} instr {v / totI * 100:5.1f}%  samples {cs[op] / tot * 100:5.1f}%")
for h in hdr:
    if h.startswith("stall_") and "Not Issued" not in h:
        s = sum(int(r[hdr.index(h)]) for r in data)
        if s > tot * 0.02:
            print(h, round(s / tot * 100, 1))
