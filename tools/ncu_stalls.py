"""Stall-reason breakdown for a range of source lines: python tools/ncu_stalls.py rep file.cuh lo hi"""
import csv, subprocess, sys
from collections import defaultdict
rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None
tot = defaultdict(int)
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or cur != fname or len(r) < 8: continue
    if r[2] == '-' and r[0].isdigit() and lo <= int(r[0]) <= hi:
        for i, h in enumerate(hdr):
            if h.startswith('stall_') and '(Not Issued)' not in h:
                try: tot[h] += int(r[i])
                except ValueError: pass
        tot['# Samples'] += int(r[hdr.index('# Samples')])
        tot['inst'] += int(r[hdr.index('Instructions Executed')])
s = tot['# Samples']
print('samples', s, 'inst', tot['inst'])
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if k.startswith('stall_') and v: print('  %-28s %5.1f%%' % (k, 100.0 * v / max(s, 1)))
