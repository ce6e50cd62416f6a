#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "n_fft or generic or istft or inverse or roundtrip or reconstruct or mel" > gpurun_out/r2_pytest_gpu22.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu22.log
timeout 300 python tools/istft_scale.py > gpurun_out/r2_istft_scale_v2.json 2>&1; grep -o '"B": [0-9]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*\|"launch": "[^"]*"' gpurun_out/r2_istft_scale_v2.json | paste - - - -
timeout 300 python tools/speech400.py > gpurun_out/r2_speech400.log 2>&1; grep -o "'n_fft': [0-9]*\|'frames_per_s': [0-9.e+]*\|'frac_of_hbm': [0-9.e+-]*\|'launches_per_call': [0-9.]*\|'launch': '[^']*'" gpurun_out/r2_speech400.log | paste - - - - -
MODE=istft_big N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_istft2 -s 2 -c 1 -o gpurun_out/r2_prof_istft2_b -f python tools/prof_cfg2.py > gpurun_out/r2_prof_istft2_b.log 2>&1; tail -1 gpurun_out/r2_prof_istft2_b.log
