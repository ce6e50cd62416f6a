#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "istft or inverse or roundtrip or reconstruct" > gpurun_out/r2_pytest_gpu39.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu39.log
timeout 300 python tools/istft_scale.py > gpurun_out/r2_istft_scale_v6.json 2>&1; grep -o '"B": [0-9]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2_istft_scale_v6.json | paste - - -
