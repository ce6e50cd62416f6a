"""Turn an .ncu-rep (from gpurun_out/) into the text summary committed under profiles/."""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[-1]))
units = dict(zip(rows[0], rows[1]))
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg']
with open(out, 'w') as f:
    f.write('# %s\n# raw metrics (ncu --set full --clock-control none)\n' % rep)
    for k in keys:
        if k in d:
            f.write('%-72s %s %s\n' % (k, d[k], units.get(k, '')))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    tmp = out + '.src.csv'
    open(tmp, 'w').write(src)
    summ = subprocess.run([sys.executable, 'tools/ncu_summary.py', tmp, '20'], capture_output=True, text=True).stdout
    f.write('\n# source-page summary (tools/ncu_summary.py)\n' + summ)
import os
os.remove(out + '.src.csv')
print(open(out).read()[:1500])
