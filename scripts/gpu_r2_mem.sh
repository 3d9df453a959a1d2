#!/bin/bash
# compute-sanitizer memcheck over the kernels new in round 2 (small shapes: the tool slows kernels 10-100x)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 120 python -m pytest -x -q -m gpu -p no:cacheprovider \
  "tests/test_gemm3_gpu.py::test_large_step_plain_product_matches_oracle" \
  "tests/test_gemm3_gpu.py::test_large_step_residual_add_epilogue" \
  "tests/test_gemm3_gpu.py::test_large_step_rope_kv_epilogue" \
  "tests/test_gemm3_gpu.py::test_large_step_silu_epilogue" \
  "tests/test_fullsize_gpu.py::test_split_kv_decode_attention_matches_oracle_and_the_unsplit_kernel" \
  "tests/test_engine_gpu.py::test_few_long_sequences_decode_through_split_kv_attention" \
  "tests/test_engine_gpu.py::test_large_step_paths_agree_token_for_token" \
  "tests/test_engine_gpu.py::test_large_prefill_steps_use_multi_tile_gemm" \
  > gpurun_out/mem_r2.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/mem_r2.log | tail -8
