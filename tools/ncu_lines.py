"""Aggregate an ncu report's warp-stall samples per CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
agg = defaultdict(lambda: [0, 0, ''])
hdr = None
total = 0
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        i_s = hdr.index('# Samples')
        i_x = hdr.index('Instructions Executed')
        continue
    if hdr is None or r[0] in ('Function Name',):
        continue
    if r[2] == '-' and r[0].isdigit():          # a CUDA source line row (aggregated over its SASS)
        try:
            s = int(r[i_s]); x = int(r[i_x])
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        agg[key][0] += s
        agg[key][1] += x
        agg[key][2] = r[1].strip()[:110]
        total += s
print('total samples', total)
for (f, ln), (s, x, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.1f%% %9d inst  %s:%d  %s' % (100.0 * s / max(total, 1), x, f, ln, src))

# ---- aggregate per enclosing function (by line ranges of "SB_HD ... name(" definitions) ----
import os, re
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'inbed_pose_estimation_b200', 'csrc')
funcs = {}
for fn in os.listdir(root):
    if not fn.endswith(('.cuh', '.cu', '.h')):
        continue
    starts = []
    for i, line in enumerate(open(os.path.join(root, fn)), 1):
        m = re.match(r'^(?:SB_HD|__global__|static|template|inline|cudaError_t)?.*?\b([A-Za-z_0-9]+)\s*\((?:const|float|int|bool|ModelView|FitParams|PoseParams)', line)
        if m and not line.startswith((' ', '\t', '//', '#')) and '(' in line and ';' not in line:
            starts.append((i, m.group(1)))
    funcs[fn] = starts
per = defaultdict(int)
for (f, ln), (s, x, src) in agg.items():
    name = '?'
    for st, nm in funcs.get(f, []):
        if st <= ln:
            name = nm
    per[(f, name)] += s
print('\nper function:')
for (f, name), s in sorted(per.items(), key=lambda kv: -kv[1])[:25]:
    print('%5.1f%%  %s:%s' % (100.0 * s / max(total, 1), f, name))
