#!/bin/bash
( bash tools/ab_bench.sh KAPRE_B200_FBBANDS 0 1 ) > gpurun_out/r2_ab_fbbands.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu3.log
cat gpurun_out/r2_ab_fbbands.log; tail -5 gpurun_out/r2_pytest_gpu3.log
