#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "n_fft or generic or istft or inverse or roundtrip or reconstruct or mel" > gpurun_out/r2_pytest_gpu23.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu23.log
timeout 300 python tools/istft_scale.py > gpurun_out/r2_istft_scale_v3.json 2>&1; grep -o '"B": [0-9]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*\|"launch": "[^"]*"' gpurun_out/r2_istft_scale_v3.json | paste - - - -
for m in 2 4 6; do echo "M=$m"; KAPRE_B200_ISTFT2_M=$m timeout 300 python tools/istft_scale.py 2>&1 | grep -o '"B": [0-9]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*' | paste - - -; done
MODE=istft_big N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_istft2 -s 2 -c 1 -o gpurun_out/r2_prof_istft2_c -f python tools/prof_cfg2.py > gpurun_out/r2_prof_istft2_c.log 2>&1; tail -1 gpurun_out/r2_prof_istft2_c.log
