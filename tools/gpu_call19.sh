#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "n_fft or spec_augment or concatenate or generic" > gpurun_out/r2_pytest_gpu19.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu19.log
timeout 300 python tools/nfft_table.py > gpurun_out/r2_nfft_v3.log 2>&1; cp gpurun_out/r2_nfft_table.json gpurun_out/r2_nfft_table_v3.json
grep -o "'n_fft': [0-9]*\|'samples_per_s': [0-9.e+]*\|'launch': '[^']*'" gpurun_out/r2_nfft_v3.log | paste - - -
KAPRE_B200_MR_WARPS=16 timeout 300 python tools/nfft_table.py > gpurun_out/r2_nfft_v3_w16.log 2>&1
grep -o "'n_fft': [0-9]*\|'samples_per_s': [0-9.e+]*\|'launch': '[^']*'" gpurun_out/r2_nfft_v3_w16.log | paste - - - | grep MR
