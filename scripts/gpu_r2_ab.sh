#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests/test_gemm3_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/ab_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/ab_tests.log
for w in 1 0 1 0; do
B200_GEMM_WIDE_EPI=$w timeout 600 python bench.py --no-cpu-baseline > gpurun_out/ab_bench_w$w.json 2> gpurun_out/ab_bench.err; echo "wide=$w exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/ab_bench_w$w.json'))
print(d['value'], d['ms_per_step'], d['step_mix'].get('T>1024'), d['step_mix'].get('T<=128'), d['ttft_p50_ms'], d['harness_output_tok_s'], d['roofline_prefill']['frac'])
print({k:round(v) for k,v in d['kernel_us_per_step'].items()})
PY
done
