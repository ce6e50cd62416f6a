#!/bin/bash
timeout 200 python tools/tc_mel_debug.py > gpurun_out/r2_tcmel_debug5.log 2>&1
timeout 120 python tools/tc_ablate.py > gpurun_out/r2_tc_ablate5.log 2>&1
tail -12 gpurun_out/r2_tcmel_debug5.log | cut -c1-200; cat gpurun_out/r2_tc_ablate5.log
