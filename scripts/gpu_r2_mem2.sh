#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 1800 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 120 python -m pytest -x -q -m gpu -p no:cacheprovider \
  tests/test_gemm3_gpu.py -k "not large_step" tests/test_ops_gpu.py -k "not large_step" \
  > gpurun_out/mem_r2b.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/mem_r2b.log | tail -8
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 120 python -m pytest -x -q -m gpu -p no:cacheprovider \
  tests/test_engine_gpu.py -k "decode_paths or forward_logits or greedy or continuous" > gpurun_out/mem_r2c.log 2>&1; echo "memcheck engine exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/mem_r2c.log | tail -8
