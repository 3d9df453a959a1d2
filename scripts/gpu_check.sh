#!/bin/bash
# One gpurun call: per-kernel parity first (each group under its own timeout so a hung kernel
# cannot eat the box), then the engine tests, then optional microbenchmarks.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, args...
  local name=$1 to=$2; shift 2
  timeout $to python -m pytest "$@" -m gpu -q --timeout 150 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$name.log
}
: > gpurun_out/summary.txt
# make sure the .so matches the sources that travelled (digest-checked; rebuilds only if stale)
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
run gemm 400 tests/test_ops_gpu.py -k gemm
run ops 300 tests/test_ops_gpu.py -k "not gemm and not attention"
run attn 400 tests/test_ops_gpu.py -k attention
run engine 600 tests/test_engine_gpu.py
run server 400 tests/test_server_gpu.py
run loader 300 tests/test_loader.py
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
cat gpurun_out/summary.txt
if [ -n "$MICRO" ]; then
  timeout 900 python scripts/microbench.py $MICRO > gpurun_out/micro.log 2>&1
  tail -n 120 gpurun_out/micro.log
fi
if [ -n "$TRACE" ]; then
  timeout 300 python scripts/gemm_trace.py > gpurun_out/trace.log 2>&1; cat gpurun_out/trace.log
fi
if [ -n "$BENCH2" ]; then
  # A/B: same bench with an environment override (e.g. B200_DEFER_MAX_T=128)
  env $BENCH2_ENV timeout 1200 python bench.py $BENCH2 > gpurun_out/bench2.log 2> gpurun_out/bench2.err
  echo "bench2 ($BENCH2_ENV) exit $?"; tail -c 3000 gpurun_out/bench2.log
fi
if [ -n "$BENCH" ]; then
  timeout 1200 python bench.py $BENCH > gpurun_out/bench.log 2> gpurun_out/bench.err
  echo "bench exit $?"; tail -c 6000 gpurun_out/bench.log; tail -n 15 gpurun_out/bench.err
fi
