#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nfft or mixed or mr or generic or any_n_fft or 1000 or 400 or 4096" > gpurun_out/r2_pytest_gpu32.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu32.log
timeout 300 python tools/nfft_table.py 2>&1 | grep -o "'n_fft': [0-9]*\|'ms': [0-9.]*\|'samples_per_s': [0-9.e+]*" | paste - - -
cp gpurun_out/r2_nfft_table.json gpurun_out/r2_nfft_table_v4.json
timeout 300 python tools/speech400.py > gpurun_out/r2_speech400_v2.json 2>&1; grep -o '"n_fft": [0-9]*\|"frames_per_s": [0-9.e+]*' gpurun_out/r2_speech400_v2.json | paste - -

