"""DRAM traffic of the kernels in an .ncu-rep as the small JSON bench.py reads (profiles/fit_kernel_rN.json,
profiles/lbs_tc_kernels_rN.json).
usage: python tools/ncu_json.py fit report.ncu-rep BATCH > profiles/fit_kernel_r2.json
       python tools/ncu_json.py lbs report.ncu-rep BATCH > profiles/lbs_tc_kernels_r2.json"""
import csv
import json
import subprocess
import sys

kind, rep, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                     universal_newlines=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]


def to_bytes(v, u):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)


kernels = []
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    rd = to_bytes(d['dram__bytes_read.sum'], u['dram__bytes_read.sum'])
    wr = to_bytes(d['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
    kernels.append({'kernel': d.get('Kernel Name', '?')[:80], 'dram_bytes': rd + wr,
                    'duration_ms': float(d['gpu__time_duration.sum'].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}.get(u['gpu__time_duration.sum'], 1)})
if kind == 'fit':
    fit = [k for k in kernels if 'smplify_fit' in k['kernel']]
    print(json.dumps({'batch': batch, 'dram_bytes_per_launch': fit[0]['dram_bytes'] if fit else None, 'kernels': kernels,
                      'source': 'ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum'}, indent=1))
else:
    tot = sum(k['dram_bytes'] for k in kernels)
    print(json.dumps({'batch': batch, 'dram_bytes_per_sample': tot / batch, 'kernels': kernels,
                      'source': 'ncu --set full --clock-control none over one LBS forward + backward (pose + tcgen05 kernels), '
                                'sum of dram__bytes_read.sum + dram__bytes_write.sum / batch'}, indent=1))
