"""Mean device time per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_times.py launches.csv"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = defaultdict(list)
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get('Metric Name') == 'gpu__time_duration.sum':
            u = d['Metric Unit']
            scale = 1e-6 if u in ('ns', 'nsecond') else 1e-3 if u in ('us', 'usecond') else 1.0
            agg[d['Kernel Name'][:80]].append(float(d['Metric Value'].replace(',', '')) * scale)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print('%-82s n=%3d  mean %8.3f ms  share %5.1f%%' % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
