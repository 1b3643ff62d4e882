"""Per-CUDA-line stall samples from `ncu -i rep --page source --csv --print-source cuda,sass > f.csv`.
usage: python scripts/ncu_lines.py f.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Line No")
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
lines = [(int(r[0]), r[1], int(r[iS]), int(r[iI])) for r in rows if r and r[0].isdigit() and len(r) > iI and r[iS].isdigit()]
tot = sum(x[2] for x in lines)
print("total samples", tot)
acc = 0
for ln, src, smp, ins in sorted(lines, key=lambda x: -x[2])[:top]:
    print("%5d %6.2f%% %10d  %s" % (ln, 100.0 * smp / tot, ins, src.strip()[:130]))
