#!/bin/bash
# Round-2 ncu evidence (1 GPU), all under gpurun_out/:
#  (1) launch list (gpu__time_duration) of two steady decode steps of the engine at the bench's shape (default fused path)
#  (2) --set full captures: gemm3 gate_up (decode dominant kernel), gemm3 down (cluster split-K + fused norm), tensor-core prefill attention
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
K='regex:gemm|attn|rmsnorm|rope|silu|embed|argmax|reduce|chain'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k "$K" -c 400 --csv \
   --log-file gpurun_out/r02_launches.csv python scripts/ncu_step.py 400 2 > gpurun_out/ncu_step.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/r02_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm3_kernel -s 2 -c 1 -o gpurun_out/r02_gemm3_gate_up_silu -f \
   python scripts/ncu_gemm.py 128 28672 4096 gemm3_silu > gpurun_out/ncu_g3_gu.log 2>&1; echo "ncu gemm3 gate_up exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm3_kernel -s 2 -c 1 -o gpurun_out/r02_gemm3_down_resadd -f \
   python scripts/ncu_gemm.py 128 4096 14336 gemm3_resadd > gpurun_out/ncu_g3_down.log 2>&1; echo "ncu gemm3 down exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:prefill_attn_tc -s 1 -c 1 -o gpurun_out/r02_attn_tc -f \
   python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k "tensor_core" -p no:cacheprovider > gpurun_out/ncu_attn_tc.log 2>&1; echo "ncu attn_tc exit $?"
ls -la gpurun_out/*.ncu-rep
