#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "2048 or fused or cfg3 or music" 2>&1 | tail -2
for v in 0 1; do echo "LEAN=$v"; KAPRE_B200_LEAN=$v timeout 300 python tools/fused_config_survey.py 2>/dev/null | grep -o "\"cfg\": \"[^\"]*\"\|\"frames_per_s\": [0-9]*\|\"ms_per_call\": [0-9.]*\|\"launch\": \"[^\"]*\"" | paste - - - - | grep 2048; done
