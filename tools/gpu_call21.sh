#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "n_fft or generic or istft or inverse or roundtrip or reconstruct or mel" > gpurun_out/r2_pytest_gpu21.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu21.log
timeout 300 python tools/speech400.py > gpurun_out/r2_speech400.log 2>&1; grep -o "'n_fft': [0-9]*\|'frames_per_s': [0-9.e+]*\|'frac_of_hbm': [0-9.e+-]*\|'launches_per_call': [0-9.]*\|'launch': '[^']*'" gpurun_out/r2_speech400.log | paste - - - - -
KAPRE_B200_NOMR=1 timeout 300 python tools/speech400.py > gpurun_out/r2_speech400_nomr.log 2>&1; grep -o "'n_fft': [0-9]*\|'frames_per_s': [0-9.e+]*\|'launches_per_call': [0-9.]*" gpurun_out/r2_speech400_nomr.log | paste - - -
MODE=istft_big N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_istft2 -s 2 -c 1 -o gpurun_out/r2_prof_istft2 -f python tools/prof_cfg2.py > gpurun_out/r2_prof_istft2.log 2>&1; tail -2 gpurun_out/r2_prof_istft2.log
