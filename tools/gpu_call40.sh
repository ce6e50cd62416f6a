#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2_r2_e.json 2> gpurun_out/bench_n2_r2_e.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2_r2_e.json'))
print('N', d['n_gpus'], 'value %.4g ms %.4f frac %.4f'%(d['value'],d['ms_per_step'],d['roofline']['frac'])); print(d['e2e']); print(d['config'].get('workload'))
PY
tail -3 gpurun_out/bench_n2_r2_e.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_n2_ref_r2_e.json 2>/dev/null; cut -c1-400 gpurun_out/bench_n2_ref_r2_e.json
