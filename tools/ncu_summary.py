"""Key metrics of every kernel in an .ncu-rep as a small markdown table (for profiles/).
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.md"""
import csv
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'regs'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma pipe %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
    ('sm__inst_executed_pipe_tensor.sum', 'tensor inst'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
    ('lts__t_sectors.sum', 'L2 sectors'), ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'),
]
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                     universal_newlines=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print('# ncu summary of `%s`\n' % rep.split('/')[-1])
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print('## %s\n' % d.get('Kernel Name', '?'))
    print('| metric | value | unit |\n|---|---|---|')
    for k, label in KEYS:
        if k in d and d[k] != '':
            print('| %s (`%s`) | %s | %s |' % (label, k, d[k], u.get(k, '')))
    print()
