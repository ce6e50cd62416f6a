#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "mel or fused or log_freq or filterbank" 2>&1 | tail -2
bash tools/ab_bench.sh KAPRE_B200_PAIRED 1 2>&1 | tail -2
