#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
for b in 1536 1024 2048 3072 1536; do
timeout 600 python bench.py --no-cpu-baseline --max-batched-tokens $b > gpurun_out/bud_bench_b$b.json 2> gpurun_out/bud_bench.err; echo "budget=$b exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/bud_bench_b$b.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['ttft_p50_ms'], d['ttft_p99_ms'], d['harness_output_tok_s'], d['config']['step_tokens_mean'], d['roofline_prefill']['frac'])
print(d['step_mix'])
PY
done
