#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 300 python scripts/chain_trace.py 8 2>&1 | tee gpurun_out/chain_trace.log | tail -60
