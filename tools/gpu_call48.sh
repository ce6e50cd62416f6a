#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cfg3 or mixed_formats or channels or stereo or mc or magnitude_db or full_tensor or mag_phase" 2>&1 | tail -2
timeout 300 python tools/config_table.py > gpurun_out/config_table_r2d.json 2>&1; grep -o '"config": "[^"]*"\|"kernel_ms": [0-9.]*\|"frac_of_measured_hbm": [0-9.]*' gpurun_out/config_table_r2d.json | paste - - - | grep cfg3
timeout 300 python tools/stereo_mel_layouts.py 2>/dev/null | grep -o '"C": [0-9]*\|"channels_[a-z]*_ms": [0-9.]*' | paste - - -
