#!/bin/bash
timeout 200 python tools/tc_mel_debug.py > gpurun_out/r2_tcmel_debug4.log 2>&1
KAPRE_B200_TC=1 MODE=meldb N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_tc_mel -s 2 -c 1 -o gpurun_out/r2_prof_tcmel_v4 -f python tools/prof_cfg2.py > gpurun_out/r2_prof_tcmel_v4.log 2>&1
tail -12 gpurun_out/r2_tcmel_debug4.log; tail -2 gpurun_out/r2_prof_tcmel_v4.log
