#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu28.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu28.log
bash tools/ab_bench.sh KAPRE_B200_PAIRED 0 1 2>&1 | tee gpurun_out/r2_ab_paired.txt
for v in 0 1; do echo "PAIRED=$v"; KAPRE_B200_PAIRED=$v timeout 300 python tools/fused_config_survey.py 2>/dev/null | grep -o "\"cfg\": \"[^\"]*\"\|\"frames_per_s\": [0-9]*\|\"ms_per_call\": [0-9.]*" | paste - - - ; done 2>&1 | tee gpurun_out/r2_ab_paired_survey.txt
