"""Aggregate an ncu report's warp-stall samples per CUDA source line (with the top stall reasons of each line).
usage: python tools/ncu_lines.py report.ncu-rep [top_n] [kernel-name-regex [nth-launch]]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cmd = ['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv']
if len(sys.argv) > 4:                       # Nth launch of the kernels matching the regex
    cmd += ['--kernel-id', '::regex:%s:%s' % (sys.argv[3], sys.argv[4])]
elif len(sys.argv) > 3:
    cmd += ['--kernel-name', 'regex:' + sys.argv[3]]
out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
agg = defaultdict(lambda: [0, 0, '', defaultdict(int)])
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
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith('stall_') and '(Not Issued)' not in h]
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
        agg[key][2] = r[1].strip()[:90]
        for i, name in stall_cols:
            try:
                agg[key][3][name] += int(r[i])
            except ValueError:
                pass
        total += s
print('total samples', total)
for (f, ln), (s, x, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ' '.join('%s=%d%%' % (n, 100 * v // max(s, 1)) for n, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print('%5.1f%% %10d inst  %s:%d  %s   [%s]' % (100.0 * s / max(total, 1), x, f, ln, src, tops))
