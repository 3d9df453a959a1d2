#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests/test_gemm3_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/z_tests.log 2>&1; echo "tests exit $?"; tail -8 gpurun_out/z_tests.log
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/z_bench_$i.json 2> gpurun_out/z_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/z_bench_$i.json'))
print(d['value'], d['ms_per_step'], d['step_mix'].get('T>1024'), d['step_mix'].get('T<=128'), d['ttft_p50_ms'], d['harness_output_tok_s'])
print({k:round(v) for k,v in d['kernel_us_per_step'].items()})
PY
done
B200_FUSED_PREFILL=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/z_bench_seg.json 2> gpurun_out/z_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/z_bench_seg.json'))
print('segments', d['value'], d['ms_per_step'], d['step_mix'].get('T>1024'), d['step_mix'].get('T<=128'))
PY
