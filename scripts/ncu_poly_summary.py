"""From an .ncu-rep of polynomial-engine launches: one line per kernel name (its longest launch) with duration, DRAM bytes and
throughput, pipe utilisation: python scripts/ncu_poly_summary.py file.ncu-rep > profiles/r02_ncu_poly_engine_summary.csv"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
best = {}
for row in rows[2:]:
    name = row[ix['Kernel Name']].split('(')[0]
    t = float(row[ix['gpu__time_duration.sum']].replace(',', ''))
    if name not in best or t > best[name][0]:
        best[name] = (t, row)
print('kernel,' + ','.join(f'{w} [{units[ix[w]]}]' for w in WANT) + ',grid,block')
for name, (t, row) in sorted(best.items(), key=lambda kv: -kv[1][0]):
    print(name + ',' + ','.join(row[ix[w]].replace(',', '') for w in WANT) + ',' + row[ix['Grid Size']].replace(',', ' ') + ',' + row[ix['Block Size']].replace(',', ' '))
