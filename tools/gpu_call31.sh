#!/bin/bash
mkdir -p gpurun_out
for n in 400 1000; do NFFT=$n timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_mr -s 2 -c 1 -o gpurun_out/r2_prof_mr$n -f python tools/prof_mr.py > gpurun_out/r2_prof_mr$n.log 2>&1; tail -1 gpurun_out/r2_prof_mr$n.log; done
