#!/bin/bash
mkdir -p gpurun_out
MODE=music N=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_stft_kernel -s 1 -c 1 -o gpurun_out/r2_prof_music -f python tools/prof_cfg2.py > gpurun_out/r2_prof_music.log 2>&1; tail -1 gpurun_out/r2_prof_music.log
