#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
for st in 4 6 8; do B200_GEMM3_STAGES=$st timeout 200 python scripts/gemm3_trace.py stages 2>&1 | tail -7; done | tee gpurun_out/g3_stages.log
echo "== norm"; timeout 200 python scripts/gemm3_trace.py 2>&1 | grep -E "norm|gate_up|qkv" | tee gpurun_out/g3_trace2.log
timeout 600 python -m pytest tests/test_gemm3_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/g3.log 2>&1; echo "g3 exit $?"; tail -3 gpurun_out/g3.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gemm3_gpu.py -m gpu -q -k norm_prologue --timeout 120 -p no:cacheprovider 2>&1 | tail -1; done
echo "== bench 200/30 fused"; timeout 600 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_200_fused.json 2> gpurun_out/bench_200_fused.err; echo "exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_200_fused.json"))
print(d["value"], d["e2e"]["value"], d["ttft_p50_ms"], d["ttft_p99_ms"]); print(" mix", d["step_mix"]); print(" dec", d["kernel_us_per_decode_step"])
PY
