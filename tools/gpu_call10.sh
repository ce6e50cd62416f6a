#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu2.log
timeout 600 python tools/nfft_table.py > gpurun_out/r2_nfft_table2.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err
timeout 300 python bench.py --config cfg5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg5_n1.json 2> gpurun_out/r2_bench_cfg5_n1.err
tail -4 gpurun_out/r2_pytest_gpu2.log; grep -E "n_fft" gpurun_out/r2_nfft_table2.log | cut -c1-110; cat gpurun_out/r2_bench_n1_b.json | cut -c1-1500; tail -2 gpurun_out/r2_bench_n1_b.err; cat gpurun_out/r2_bench_cfg5_n1.json | cut -c1-800; tail -2 gpurun_out/r2_bench_cfg5_n1.err
