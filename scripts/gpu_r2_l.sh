#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
for pf in 0 32 96; do echo "== prefetch $pf"; B200_CHAIN_PREFETCH=$pf timeout 300 python scripts/chain_trace.py 8 2>&1 | head -23 | tail -22; done | tee gpurun_out/chain_trace_pf.log
