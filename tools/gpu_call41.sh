#!/bin/bash
mkdir -p gpurun_out
MODE=meldb N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_stft_kernel -s 2 -c 1 -o gpurun_out/r2_prof_meldb_g -f python tools/prof_cfg2.py > gpurun_out/r2_prof_meldb_g.log 2>&1; tail -1 gpurun_out/r2_prof_meldb_g.log
