#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu49.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu49.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
