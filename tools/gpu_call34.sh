#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "phase or mag_phase" > gpurun_out/r2_pytest_gpu34.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu34.log
timeout 300 python tools/mode_ab.py > gpurun_out/r2_mode_ab_atan.json 2>&1; grep -o '"n_fft": [0-9]*\|"mode": "[^"]*"\|"kernel_ms": [0-9.]*' gpurun_out/r2_mode_ab_atan.json | paste - - - | grep phase
