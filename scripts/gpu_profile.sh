#!/bin/bash
# ncu evidence (1 GPU): (1) launch list of two decode steps of the engine, (2) full capture of the pair GEMM.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|attn|rmsnorm|rope|silu|embed|argmax|reduce" \
   -s 590 -c 584 --csv --log-file gpurun_out/launches.csv python scripts/ncu_step.py 16 > gpurun_out/ncu_step.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/launches.csv
