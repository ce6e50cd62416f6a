#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "n_fft or spec_augment or concatenate or generic or istft or inverse or roundtrip or reconstruct" > gpurun_out/r2_pytest_gpu20.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu20.log
timeout 300 python tools/nfft_table.py > gpurun_out/r2_nfft_v3.log 2>&1; cp gpurun_out/r2_nfft_table.json gpurun_out/r2_nfft_table_v3.json
grep -o "'n_fft': [0-9]*\|'samples_per_s': [0-9.e+]*\|'launch': '[^']*'" gpurun_out/r2_nfft_v3.log | paste - - -
KAPRE_B200_MR_WARPS=16 timeout 300 python tools/nfft_table.py > gpurun_out/r2_nfft_v3_w16.log 2>&1
grep -o "'n_fft': [0-9]*\|'samples_per_s': [0-9.e+]*\|'launch': '[^']*'" gpurun_out/r2_nfft_v3_w16.log | paste - - - | grep MR
for v in 0 1; do echo "ISTFT2=$v"; KAPRE_B200_ISTFT2=$v timeout 300 python tools/istft_scale.py > gpurun_out/r2_istft_scale_v$v.json 2>&1; grep -o '"B": [0-9]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*\|"launch": "[^"]*"' gpurun_out/r2_istft_scale_v$v.json | paste - - - -; done
for m in 2 4 8 16 40; do echo "M=$m"; KAPRE_B200_ISTFT2_M=$m timeout 300 python tools/istft_scale.py 2>&1 | grep -o '"B": [0-9]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*' | paste - - -; done
