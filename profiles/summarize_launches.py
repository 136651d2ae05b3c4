"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: count, total time, share."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[iv].replace(',', ''))
    except ValueError:
        continue
    u = r[iu]
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(u, 1e-6)
    name = re.sub(r'\(.*', '', r[ik]).split('::')[-1]
    tot[name][0] += 1
    tot[name][1] += v
T = sum(v[1] for v in tot.values())
print(f'{"kernel":34s} {"launches":>8s} {"total ms":>10s} {"share":>7s}')
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:34s} {n:8d} {ms:10.3f} {100 * ms / T:6.1f}%')
print(f'{"TOTAL":34s} {sum(v[0] for v in tot.values()):8d} {T:10.3f}')
