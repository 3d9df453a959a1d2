#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
echo "== umma layout probe"; timeout 60 scripts/exp/umma_layout_probe > gpurun_out/umma_probe.log 2>&1; echo "exit $?"; head -80 gpurun_out/umma_probe.log
timeout 600 python -m pytest tests/test_gemm3_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/g3.log 2>&1; echo "g3 exit $?"; tail -3 gpurun_out/g3.log
echo "== gemm3 trace"; timeout 300 python scripts/gemm3_trace.py > gpurun_out/g3_trace.log 2>&1; tail -14 gpurun_out/g3_trace.log
timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/eng.log 2>&1; echo "engine+fullsize exit $?"; tail -3 gpurun_out/eng.log
echo "== bench 200/30 fused"; timeout 600 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_200_fused.json 2> gpurun_out/bench_200_fused.err; echo "exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_200_fused.json"))
print(d["value"], d["e2e"]["value"], d["ttft_p50_ms"], d["ttft_p99_ms"]); print(" mix", d["step_mix"]); print(" dec", d["kernel_us_per_decode_step"]); print(" roof", d["roofline"]["frac"], d["step_roofline"])
PY
