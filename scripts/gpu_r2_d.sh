#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
echo "== norm prologue variants"; timeout 300 python scripts/gemm3_trace.py norm > gpurun_out/g3_trace_norm.log 2>&1; cat gpurun_out/g3_trace_norm.log | tail -24
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -s > gpurun_out/fullsize.log 2>&1; echo "fullsize exit $?"
grep -E "passed|failed|max \|dlogit|Error|assert" gpurun_out/fullsize.log | tail -12
