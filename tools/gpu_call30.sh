#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do KAPRE_B200_PAIRED=$v timeout 300 python tools/mode_ab.py > gpurun_out/r2_mode_ab_paired$v.json 2>&1; grep -o '"n_fft": [0-9]*\|"mode": "[^"]*"\|"kernel_ms": [0-9.]*' gpurun_out/r2_mode_ab_paired$v.json | paste - - -; done
