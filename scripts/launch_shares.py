"""Share of the GPU time by kernel from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file f.csv ...`).
usage: python scripts/launch_shares.py f.csv [first_launch last_launch]   (launch numbers = position in the list)"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
rows = rows[lo:hi]
t, n = defaultdict(float), defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    t[name] += float(r[-1])
    n[name] += 1
tot = sum(t.values())
print("launches %d, total %.3f ms" % (len(rows), tot * 1e-6))
for k in sorted(t, key=lambda k: -t[k]):
    print("%6.2f %%  %9.3f ms  %6d x %8.1f us  %s" % (100 * t[k] / tot, t[k] * 1e-6, n[k], t[k] / n[k] * 1e-3, k))
