#!/bin/bash
mkdir -p gpurun_out
MODE=cfg3 N=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_stft_mc -s 1 -c 1 -o gpurun_out/r2_prof_cfg3 -f python tools/prof_cfg2.py > gpurun_out/r2_prof_cfg3.log 2>&1; tail -1 gpurun_out/r2_prof_cfg3.log
