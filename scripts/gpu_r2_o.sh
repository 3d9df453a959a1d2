#!/bin/bash
# same-box A/B of the three decode paths
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_ab_$name.json 2> gpurun_out/bench_ab_$name.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab_$name.json"))
dec=d["step_mix"].get("T<=128",{})
pre=d["step_mix"].get("T>1024",{})
print("$name", "value", d["value"], "e2e", d["e2e"]["value"], "decode step ms", dec.get("mean_ms"), "prefill step ms", pre.get("mean_ms"), pre.get("mean_tokens"), "ttft", d["ttft_p50_ms"], d["ttft_p99_ms"], "clk", d["clocks"]["sm_mhz"])
PY
}
run unfused B200_FUSED_DECODE=0
run per_gemm B200_CHAIN=0
run chain B200_CHAIN=1
run chain_pf96 B200_CHAIN=1 B200_CHAIN_PREFETCH=96
run unfused2 B200_FUSED_DECODE=0
