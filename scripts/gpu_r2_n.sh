#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
for d in 0 1; do echo "== dbg $d"; B200_CHAIN_DBG=$d B200_CHAIN_PREFETCH=32 timeout 300 python scripts/chain_trace.py 8 2>&1 | head -23 | grep -E "first MMA|last MMA|segments out|barrier|done|CTA end|traced"; done | tee gpurun_out/chain_trace_dbg.log
