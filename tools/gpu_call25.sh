#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu25.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu25.log
timeout 300 python tools/istft_scale.py > gpurun_out/r2_istft_scale_v5.json 2>&1; grep -o '"B": [0-9]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*\|"launch": "[^"]*"' gpurun_out/r2_istft_scale_v5.json | paste - - - -
for m in 3 4 8; do echo "M=$m"; KAPRE_B200_ISTFT2_M=$m timeout 300 python tools/istft_scale.py 2>&1 | grep -o '"B": [0-9]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*' | paste - - -; done
timeout 300 python tools/config_table.py > gpurun_out/r2_config_table.log 2>&1; tail -12 gpurun_out/r2_config_table.log | cut -c1-200
MODE=istft_big N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_istft2 -s 2 -c 1 -o gpurun_out/r2_prof_istft2_e -f python tools/prof_cfg2.py > gpurun_out/r2_prof_istft2_e.log 2>&1; tail -1 gpurun_out/r2_prof_istft2_e.log
