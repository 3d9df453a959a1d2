#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_gpu.py tests/test_gemm3_gpu.py tests/test_ops_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/eng.log 2>&1; echo "tests exit $?"; grep -E "passed|failed|Error|assert" gpurun_out/eng.log | tail -5
timeout 300 python scripts/chain_trace.py 8 2>&1 | tee gpurun_out/chain_trace.log | head -27
echo "== bench 200/30 chain"; timeout 600 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_200_chain.json 2> gpurun_out/bench_200_chain.err; echo "exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_200_chain.json"))
print(d["value"], d["e2e"]["value"], d["ttft_p50_ms"], d["ttft_p99_ms"]); print(" mix", d["step_mix"]); print(" dec", d["kernel_us_per_decode_step"]); print(" roof", d["roofline"]["frac"], d["step_roofline"]["decode_steps"])
PY
tail -3 gpurun_out/bench_200_chain.err
