#!/bin/bash
# Round-2 GPU call B: the decode-shape fused GEMM (gemm3) group by group, then the engine on the fused path, then timings.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
run() { # name, timeout, pytest args...
  local name=$1 to=$2; shift 2
  timeout $to python -m pytest "$@" -m gpu -q --timeout 120 -p no:cacheprovider -x > gpurun_out/$name.log 2>&1
  echo "$name exit $?" | tee -a gpurun_out/summary_b.txt
  grep -E "passed|failed|skipped|Error|error|assert|differ|tolerance" gpurun_out/$name.log | tail -12
}
: > gpurun_out/summary_b.txt
run g3_plain_s1 240 tests/test_gemm3_gpu.py -k "plain_product_matches_oracle_for_every and 1-"
run g3_plain_auto 300 tests/test_gemm3_gpu.py -k "plain_product_matches_oracle_for_every"
run g3_streamk 300 tests/test_gemm3_gpu.py -k "stream_k"
run g3_norm 300 tests/test_gemm3_gpu.py -k "norm_prologue"
run g3_resadd 300 tests/test_gemm3_gpu.py -k "residual_add"
run g3_silu 300 tests/test_gemm3_gpu.py -k "silu"
run g3_rope 300 tests/test_gemm3_gpu.py -k "rope"
run g3_argmax 300 tests/test_gemm3_gpu.py -k "argmax"
run engine 600 tests/test_engine_gpu.py
run fullsize 600 tests/test_fullsize_gpu.py -s
run server 400 tests/test_server_gpu.py
run loader 300 tests/test_loader.py
run ops 400 tests/test_ops_gpu.py
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary_b.txt; tail -2 gpurun_out/smoke.log
cat gpurun_out/summary_b.txt
echo "== microbench gemm3"; timeout 300 python scripts/microbench.py gemm3 > gpurun_out/micro_g3.log 2>&1; tail -n 40 gpurun_out/micro_g3.log
echo "== bench 200/30 fused"; timeout 600 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_200_fused.json 2> gpurun_out/bench_200_fused.err; echo "exit $?"; head -c 2500 gpurun_out/bench_200_fused.json; echo; tail -3 gpurun_out/bench_200_fused.err
echo "== bench 200/30 unfused"; B200_FUSED_DECODE=0 timeout 600 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_200_unfused.json 2> gpurun_out/bench_200_unfused.err; echo "exit $?"; head -c 900 gpurun_out/bench_200_unfused.json; echo
