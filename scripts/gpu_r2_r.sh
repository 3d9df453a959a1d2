#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
for v in 256 1073741824; do
B200_DEFER_MAX_T=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r_bench_defer$v.json 2> gpurun_out/r_bench.err; echo "defer_max_t=$v exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/r_bench_defer$v.json'))
print(d['value'], d['ms_per_step'], d['step_mix'].get('T>1024'), d['step_mix'].get('T<=128'))
print({k:round(v) for k,v in d['kernel_us_per_step'].items()})
PY
done
