#!/bin/bash
# sanitizer passes on the hot kernels (subset of the GPU tier) + smoke
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "melspectrogram_fused_vs_oracle or stft_any_n_fft or decibel_clamp_long or istft_vs_oracle or concatenate or stft_mag_phase" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log
KAPRE_B200_TC=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_tc_dft.py -x -q -m gpu > gpurun_out/r2_sanitizer_memcheck_tc.log 2>&1; echo "memcheck-tc rc=$?" >> gpurun_out/r2_sanitizer_memcheck_tc.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "melspectrogram_fused_vs_oracle and 1024 or stft_any_n_fft and 1000" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
tail -4 gpurun_out/r2_sanitizer_memcheck.log; tail -4 gpurun_out/r2_sanitizer_memcheck_tc.log; tail -6 gpurun_out/r2_sanitizer_racecheck.log; tail -3 gpurun_out/r2_smoke.log
