#!/bin/bash
mkdir -p gpurun_out
MODE=istft_big N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_istft2 -s 2 -c 1 -o gpurun_out/r2_prof_istft2_f -f python tools/prof_cfg2.py > gpurun_out/r2_prof_istft2_f.log 2>&1; tail -1 gpurun_out/r2_prof_istft2_f.log
