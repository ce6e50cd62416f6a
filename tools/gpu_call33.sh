#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu33.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu33.log
timeout 400 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_n1_r2_d.json 2> gpurun_out/bench_n1_r2_d.err; cut -c1-300 gpurun_out/bench_n1_r2_d.json
timeout 300 python tools/config_table.py > gpurun_out/config_table_r2c.json 2>&1; grep -o '"config": "[^"]*"\|"kernel_ms": [0-9.]*\|"frac_of_measured_hbm": [0-9.]*' gpurun_out/config_table_r2c.json | paste - - -
timeout 300 python tools/fused_config_survey.py > gpurun_out/fused_config_survey_r2.json 2>&1; grep -o '"cfg": "[^"]*"\|"frames_per_s": [0-9]*\|"frac": [0-9.]*' gpurun_out/fused_config_survey_r2.json | paste - - -
MODE=meldb N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_stft_kernel -s 2 -c 1 -o gpurun_out/r2_prof_meldb_paired -f python tools/prof_cfg2.py > gpurun_out/r2_prof_meldb_paired.log 2>&1; tail -1 gpurun_out/r2_prof_meldb_paired.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launch_list_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launch_list_bench.log 2>&1; tail -c 300 gpurun_out/r2_launch_list_bench.log
