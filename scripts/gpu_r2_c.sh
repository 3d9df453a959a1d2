#!/bin/bash
# Round-2 GPU call C: gemm3 after the transform fix + traces, tensor-core prefill attention, full suite, bench A/B.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
run() { local name=$1 to=$2; shift 2
  timeout $to python -m pytest "$@" -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit $?" | tee -a gpurun_out/summary_c.txt
  grep -E "passed|failed|skipped|Error|error|assert|differ|tolerance|max \|dlogit" gpurun_out/$name.log | tail -14
}
: > gpurun_out/summary_c.txt
run g3 600 tests/test_gemm3_gpu.py
run attn_tc 400 tests/test_ops_gpu.py -k "tensor_core"
run fullsize 900 tests/test_fullsize_gpu.py -s
B200_SKIP_VLLM=1 run rest 900 tests --deselect tests/test_gemm3_gpu.py --deselect tests/test_fullsize_gpu.py
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary_c.txt; tail -2 gpurun_out/smoke.log
echo "== gemm3 trace"; timeout 300 python scripts/gemm3_trace.py > gpurun_out/g3_trace.log 2>&1; cat gpurun_out/g3_trace.log | tail -20
echo "== bench 200/30 fused"; timeout 600 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_200_fused.json 2> gpurun_out/bench_200_fused.err; echo "exit $?"; python - <<'PY'
import json
for n in ("fused",):
    try:
        d=json.load(open(f"gpurun_out/bench_200_{n}.json"))
        print(n, d["value"], d["e2e"]["value"], d["ttft_p50_ms"], d["ttft_p99_ms"]); print(" mix", d["step_mix"]); print(" dec", d["kernel_us_per_decode_step"]); print(" all", d["kernel_us_per_step"])
    except Exception as e: print(n, "failed", e)
PY
tail -3 gpurun_out/bench_200_fused.err
echo "== bench 200/30 fused, old prefill attention"; B200_ATTN_TC=0 timeout 600 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_200_fused_oldattn.json 2> gpurun_out/bench_200_fused_oldattn.err; echo "exit $?"; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_200_fused_oldattn.json")); print(d["value"], d["step_mix"]); print(" all", d["kernel_us_per_step"])
except Exception as e: print("failed", e)
PY
