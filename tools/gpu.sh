#!/bin/bash
# Build first (abort on failure, so a stale .so is never measured), then hand over to gpurun.
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > /tmp/kb_build.log 2>&1 || { tail -5 /tmp/kb_build.log; echo "BUILD FAILED"; exit 1; }
exec /usr/local/graft/bin/gpurun "$@"
