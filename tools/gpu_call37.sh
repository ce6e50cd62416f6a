#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "predict or host_path or pipelined" 2>&1 | tail -2
timeout 400 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_n1_r2_e.json 2> gpurun_out/bench_n1_r2_e.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1_r2_e.json'))
print('value %.4g ms %.4f frac %.4f'%(d['value'],d['ms_per_step'],d['roofline']['frac'])); print(d['e2e'])
PY
