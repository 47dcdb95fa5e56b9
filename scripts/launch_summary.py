"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python scripts/launch_summary.py list.csv ["header comment"]"""
import csv
import sys
from collections import defaultdict

tot, cnt = defaultdict(float), defaultdict(int)
for row in csv.reader(open(sys.argv[1], errors="replace")):
    if len(row) > 10 and row[0].isdigit():
        name = row[4].split("(")[0][:70]
        tot[name] += float(row[-1]) / 1e6
        cnt[name] += 1
allms = sum(tot.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("kernel,launches,total_ms,share_percent")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f'"{k}",{cnt[k]},{tot[k]:.4f},{100 * tot[k] / allms:.2f}')
