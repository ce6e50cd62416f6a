#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu4.log
for nw in 0 4 1 2; do KAPRE_B200_NW=$nw timeout 120 python tools/small_batch_latency.py 2>&1 | tail -1 | sed "s/^/NW=$nw /"; done > gpurun_out/r2_small_batch.log 2>&1
KAPRE_B200_NW=0 timeout 300 python tools/config_table.py > gpurun_out/r2_config_table.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err
tail -3 gpurun_out/r2_pytest_gpu4.log; cat gpurun_out/r2_small_batch.log; tail -12 gpurun_out/r2_config_table.log | cut -c1-220; cut -c1-300 gpurun_out/r2_bench_n1_c.json
