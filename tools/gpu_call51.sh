#!/bin/bash
mkdir -p gpurun_out
NFFT=400 timeout 200 ncu --set full --clock-control none --import-source on -k regex:kb_mr -s 2 -c 1 -o gpurun_out/r2_prof_mr400_c -f python tools/prof_mr.py > gpurun_out/r2_prof_mr400_c.log 2>&1; tail -1 gpurun_out/r2_prof_mr400_c.log
