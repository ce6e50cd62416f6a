#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/tc_dft_debug.py > gpurun_out/r2_tc_debug2.log 2>&1
timeout 300 python tools/tc_dft_probe.py > gpurun_out/r2_tc_probe2.json 2> gpurun_out/r2_tc_probe2.err
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "any_n_fft or n1000 or apply_filterbank or complex_vs_oracle" > gpurun_out/r2_pytest_mr.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_mr.log
timeout 600 python tools/nfft_table.py > gpurun_out/r2_nfft_table.log 2>&1
head -12 gpurun_out/r2_tc_debug2.log; cat gpurun_out/r2_tc_probe2.json; tail -3 gpurun_out/r2_tc_probe2.err; tail -5 gpurun_out/r2_pytest_mr.log; cat gpurun_out/r2_nfft_table.log
