#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu46.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu46.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_n1_r2_f.json 2> gpurun_out/bench_n1_r2_f.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1_r2_f.json'))
print('value %.4g ms %.4f frac %.4f steps %d launches %d'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['steps'],d['gpu_launches'])); print(d['e2e']); print(d['cpu_baseline']); print(d['clocks'])
PY
timeout 400 python bench.py --impl reference 2>/dev/null | cut -c1-300
