#!/bin/bash
# 1 GPU: per-projection choice between fused epilogues and fp32 segments on large steps — tests, then bench A/B
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/v_eng.log 2>&1; echo "engine exit $?"; tail -15 gpurun_out/v_eng.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline $EXTRA > gpurun_out/v_bench_$name.json 2> gpurun_out/v_bench.err; echo "$name exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/v_bench_$name.json'))
print(d['value'], d['ms_per_step'], d['step_mix'].get('T>1024'), d['step_mix'].get('T<=128'), d['ttft_p50_ms'], d['harness_output_tok_s'])
print({k:round(v) for k,v in d['kernel_us_per_step'].items()})
PY
}
EXTRA=""
run segments B200_FUSED_PREFILL=0
run auto B200_FUSED_PREFILL=1
run bn256 B200_F2_BN=256
run bn512 B200_F2_BN=512
EXTRA="--max-batched-tokens 512"
run b512_segments B200_FUSED_PREFILL=0
run b512_auto B200_FUSED_PREFILL=1
