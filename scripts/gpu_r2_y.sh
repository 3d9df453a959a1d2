#!/bin/bash
# 1 GPU: large-step kernel bars vs cuBLAS, ncu launch list of one prefill-burst step, tensor-core attention capture summary
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 600 python scripts/microbench.py large > gpurun_out/y_micro_large.log 2>&1; echo "micro exit $?"; tail -20 gpurun_out/y_micro_large.log
K='regex:gemm|attn|rmsnorm|rope|silu|embed|argmax|reduce|chain'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k "$K" -c 600 --csv \
   --log-file gpurun_out/r02_launches_burst.csv python scripts/ncu_burst.py > gpurun_out/y_ncu_burst.log 2>&1
echo "ncu burst exit $?"; tail -2 gpurun_out/y_ncu_burst.log; wc -l gpurun_out/r02_launches_burst.csv
