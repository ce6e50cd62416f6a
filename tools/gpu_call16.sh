#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --config cfg5 > gpurun_out/r2_bench_n2_cfg5.json 2> gpurun_out/r2_bench_n2_cfg5.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_n2_ref.json 2> gpurun_out/r2_bench_n2_ref.err
cut -c1-700 gpurun_out/r2_bench_n2.json; tail -2 gpurun_out/r2_bench_n2.err; cut -c1-600 gpurun_out/r2_bench_n2_cfg5.json; tail -2 gpurun_out/r2_bench_n2_cfg5.err; cut -c1-400 gpurun_out/r2_bench_n2_ref.json
