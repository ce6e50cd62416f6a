#!/bin/bash
# usage: tools/ab_bench.sh ENVVAR v1 v2 ...   -> runs bench.py for each value (twice) and prints value / kernel_ms
var=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    env $var=$v python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null > /tmp/ab_line.json
    python - "$var" "$v" <<'PY'
import json, sys
d = json.load(open('/tmp/ab_line.json'))
print(sys.argv[1], sys.argv[2], 'frames/s %.4g' % d['value'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'step_ms %.4f' % d['ms_per_step'])
PY
  done
done
