#!/bin/bash
mkdir -p gpurun_out
for v in 2 6; do KAPRE_B200_VARIANT=$v timeout 900 python -m pytest tests -m gpu -x -q -k "mel or fused or log_freq or filterbank" > gpurun_out/r2_pytest_gpu27_v$v.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu27_v$v.log; done
bash tools/ab_bench.sh KAPRE_B200_VARIANT 0 2 4 6 2>&1 | tee gpurun_out/r2_ab_variants.txt
