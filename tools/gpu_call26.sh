#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu26.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu26.log
timeout 300 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_n1_r2_c.json 2> gpurun_out/bench_n1_r2_c.err; cut -c1-600 gpurun_out/bench_n1_r2_c.json
timeout 300 python tools/nfft_table.py > gpurun_out/r2_nfft_table_v3.json 2>&1; tail -c 600 gpurun_out/r2_nfft_table_v3.json
timeout 300 python tools/speech400.py > gpurun_out/r2_speech400.json 2>&1; tail -c 400 gpurun_out/r2_speech400.json
timeout 300 python tools/config_table.py > gpurun_out/config_table_r2b.json 2>&1; tail -c 300 gpurun_out/config_table_r2b.json
timeout 300 python tools/istft_scale.py > gpurun_out/r2_istft_scale_v5.json 2>&1; grep -o '"B": [0-9]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2_istft_scale_v5.json | paste - - -
MODE=meldb N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_stft_kernel -s 2 -c 1 -o gpurun_out/r2_prof_meldb_final -f python tools/prof_cfg2.py > gpurun_out/r2_prof_meldb_final.log 2>&1; tail -1 gpurun_out/r2_prof_meldb_final.log
