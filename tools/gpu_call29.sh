#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu29.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu29.log
KAPRE_B200_PAIRED=0 timeout 600 python -m pytest tests -m gpu -x -q -k "parity" > gpurun_out/r2_pytest_gpu29_nat.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu29_nat.log
for v in 0 1; do KAPRE_B200_PAIRED=$v timeout 300 python tools/mode_ab.py > gpurun_out/r2_mode_ab_paired$v.json 2>&1; grep -o '"n_fft": [0-9]*\|"mode": "[^"]*"\|"kernel_ms": [0-9.]*' gpurun_out/r2_mode_ab_paired$v.json | paste - - -; done
bash tools/ab_bench.sh KAPRE_B200_PAIRED 0 1 2>&1 | tail -2
