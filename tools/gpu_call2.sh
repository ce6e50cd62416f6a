#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_dft.py -x -q -m gpu > gpurun_out/r2_pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_tc.log
timeout 300 python tools/tc_dft_probe.py > gpurun_out/r2_tc_probe.json 2> gpurun_out/r2_tc_probe.err
MODE=prof timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_tc_dft -s 2 -c 1 -o gpurun_out/r2_prof_tc_stage1 -f python tools/tc_dft_probe.py > gpurun_out/r2_prof_tc.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "melspectrogram or log_frequency" > gpurun_out/r2_pytest_mel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_mel.log
( bash tools/ab_bench.sh KAPRE_B200_FBMMA 0 ) > gpurun_out/r2_ab_plan.log 2>&1
tail -5 gpurun_out/r2_pytest_tc.log; cat gpurun_out/r2_tc_probe.json; tail -3 gpurun_out/r2_tc_probe.err; tail -3 gpurun_out/r2_pytest_mel.log; cat gpurun_out/r2_ab_plan.log
