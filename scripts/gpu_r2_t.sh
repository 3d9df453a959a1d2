#!/bin/bash
# 1 GPU: fused epilogues on steps of more than 128 tokens — kernel tests, engine tests, bench A/B
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gemm3_gpu.py -x -q -m gpu -k "large_step" -p no:cacheprovider > gpurun_out/t_g3.log 2>&1; echo "gemm3 large exit $?"; tail -15 gpurun_out/t_g3.log
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_eng.log 2>&1; echo "engine exit $?"; tail -15 gpurun_out/t_eng.log
for v in 0 1; do
B200_FUSED_PREFILL=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/t_bench_fp$v.json 2> gpurun_out/t_bench.err; echo "fused_prefill=$v exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/t_bench_fp$v.json'))
print(d['value'], d['ms_per_step'], d['step_mix'].get('T>1024'), d['step_mix'].get('T<=128'), d['ttft_p50_ms'], d['harness_output_tok_s'])
print({k:round(v) for k,v in d['kernel_us_per_step'].items()})
PY
done
