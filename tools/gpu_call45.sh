#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "any_n_fft or n1000 or mixed" 2>&1 | tail -2


timeout 300 python tools/speech400.py > /dev/null 2>&1; python - <<'PY'
import json
for r in json.load(open('gpurun_out/r2_speech400.json')):
    print(r['n_fft'],r['hop'],r['n_mels'],'%.4f ms'%r['ms'],'%.3g fr/s'%r['frames_per_s'])
PY
