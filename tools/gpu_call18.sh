#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 100 --warmup 5 --config cfg2 > gpurun_out/r2_bench_n8_cfg2.json 2> gpurun_out/r2_bench_n8_cfg2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_n8_ref.json 2> gpurun_out/r2_bench_n8_ref.err
cut -c1-400 gpurun_out/r2_bench_n8.json; tail -3 gpurun_out/r2_bench_n8.err; cut -c1-300 gpurun_out/r2_bench_n8_cfg2.json; tail -2 gpurun_out/r2_bench_n8_cfg2.err; cut -c1-300 gpurun_out/r2_bench_n8_ref.json
