#!/bin/bash
# GPU call 1 (round 2): parity of the new paths, mma.sync issue rates, A/B of the filterbank phase and PDL.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu1.log
timeout 120 tools/_bin/mma_bench > gpurun_out/r2_mma_bench.json 2>&1
( bash tools/ab_bench.sh KAPRE_B200_FBMMA 0 1 ) > gpurun_out/r2_ab_fbmma.log 2>&1
( bash tools/ab_bench.sh KAPRE_B200_PDL_FWD 0 1 ) > gpurun_out/r2_ab_pdl.log 2>&1
KAPRE_B200_FBMMA=1 MODE=meldb N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_stft_kernel -s 2 -c 1 -o gpurun_out/r2_prof_meldb_fbmma -f python tools/prof_cfg2.py > gpurun_out/r2_prof_fbmma.log 2>&1
KAPRE_B200_FBMMA=0 MODE=meldb N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_stft_kernel -s 2 -c 1 -o gpurun_out/r2_prof_meldb_base -f python tools/prof_cfg2.py > gpurun_out/r2_prof_base.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
tail -3 gpurun_out/r2_pytest_gpu1.log; cat gpurun_out/r2_ab_fbmma.log gpurun_out/r2_ab_pdl.log; cat gpurun_out/r2_bench_n1_a.json
