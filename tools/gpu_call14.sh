#!/bin/bash
KAPRE_B200_TC=1 KAPRE_B200_TC_ABLATE=6 MODE=meldb N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_tc_mel -s 2 -c 1 -o gpurun_out/r2_prof_tcmel_abl6 -f python tools/prof_cfg2.py > gpurun_out/r2_prof_tcmel_abl6.log 2>&1
tail -2 gpurun_out/r2_prof_tcmel_abl6.log
