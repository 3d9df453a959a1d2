#!/bin/bash
# 1 GPU: vLLM parity with the ulp metric, default bench (20 and 200 steps), reference arm, tensor-core attention capture
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests/test_vllm_parity_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/q_vllm.log 2>&1; echo "vllm parity exit $?"; tail -5 gpurun_out/q_vllm.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/q_bench20.json 2> gpurun_out/q_bench20.err; echo "bench20 exit $?"; cut -c1-600 gpurun_out/q_bench20.json
timeout 600 python bench.py > gpurun_out/q_bench200.json 2> gpurun_out/q_bench200.err; echo "bench200 exit $?"; cut -c1-300 gpurun_out/q_bench200.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/q_ref.json 2> gpurun_out/q_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/q_ref.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:prefill_attn_tc -c 1 -o gpurun_out/r02_attn_tc -f \
   python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k "tensor_core" -p no:cacheprovider > gpurun_out/ncu_attn_tc.log 2>&1; echo "ncu attn_tc exit $?"
ls -la gpurun_out/*.ncu-rep
