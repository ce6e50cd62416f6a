#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu35.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu35.log
timeout 300 python tools/stereo_mel_layouts.py > gpurun_out/stereo_mel_layouts_r2.json 2>&1; grep -o '"C": [0-9]*\|"channels_[a-z]*_ms": [0-9.]*' gpurun_out/stereo_mel_layouts_r2.json | paste - - -
for v in 0 1; do KAPRE_B200_PAIRED=$v timeout 300 python tools/mode_ab.py 2>/dev/null | grep -o '"n_fft": [0-9]*\|"mode": "[^"]*"\|"kernel_ms": [0-9.]*' | paste - - - | grep "1024.*phase"; done
