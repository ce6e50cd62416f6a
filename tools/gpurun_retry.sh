#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script> <outfile>   -- retries while the pod answers "transient" / busy
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" > "$3" 2>&1
  if grep -q "status=transient\|rc=3\|no box" "$3"; then sleep 150; continue; fi
  break
done
