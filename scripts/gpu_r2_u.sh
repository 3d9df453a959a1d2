#!/bin/bash
# 1 GPU: token-budget sweep with the fused large-step path
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
for b in 256 384 512 768 1024 1536; do
timeout 600 python bench.py --no-cpu-baseline --max-batched-tokens $b > gpurun_out/u_bench_b$b.json 2> gpurun_out/u_bench.err; echo "budget=$b exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/u_bench_b$b.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['ttft_p50_ms'], d['ttft_p99_ms'], d['harness_output_tok_s'], d['config']['step_tokens_mean'])
print(d['step_mix'])
PY
done
