#!/bin/bash
# Round-2 GPU call A: full GPU suite (vLLM parity apart), bench at the driver's and at the long setting, token-budget
# points below 1024, per-kernel bars, cluster/DSMEM probe.  Every part under its own timeout; logs in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/vllm_parity.json
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
echo "== probe"; timeout 120 scripts/exp/cluster_probe > gpurun_out/cluster_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/cluster_probe.log
echo "== tests (no vllm)"
B200_SKIP_VLLM=1 timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -s > gpurun_out/tests.log 2>&1; echo "tests exit $?"
grep -E "passed|failed|error|max \|dlogit\||FAILED|Error" gpurun_out/tests.log | tail -40
echo "== bench 20/5"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_20.json 2> gpurun_out/bench_20.err; echo "exit $?"; head -c 1500 gpurun_out/bench_20.json; echo
echo "== bench 200/30"; timeout 600 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_200.json 2> gpurun_out/bench_200.err; echo "exit $?"; head -c 1200 gpurun_out/bench_200.json; echo
for B in 384 512 768; do
  echo "== bench 200/30 budget $B"; timeout 600 python bench.py --steps 200 --warmup 30 --no-cpu-baseline --max-batched-tokens $B > gpurun_out/bench_200_b$B.json 2> gpurun_out/bench_200_b$B.err; echo "exit $?"; head -c 700 gpurun_out/bench_200_b$B.json; echo
done
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2>&1; tail -c 900 gpurun_out/bench_ref.json
echo "== microbench"; timeout 600 python scripts/microbench.py gemm attn > gpurun_out/micro.log 2>&1; echo "exit $?"; tail -n 60 gpurun_out/micro.log
echo "== vllm parity"; timeout 1100 python -m pytest tests/test_vllm_parity_gpu.py -m gpu -q --timeout 1000 -p no:cacheprovider -s > gpurun_out/vllm_tests.log 2>&1; echo "vllm exit $?"
grep -E "passed|failed|logprobs_compared|Error|assert" gpurun_out/vllm_tests.log | tail -12
