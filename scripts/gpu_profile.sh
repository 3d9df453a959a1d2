#!/bin/bash
# ncu evidence (1 GPU): (1) launch list of two decode steps of the engine, (2) full capture of the pair GEMM.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|attn|rmsnorm|rope|silu|embed|argmax|reduce" \
   -s 700 -c 600 --csv --log-file gpurun_out/launches.csv python scripts/ncu_step.py 16 > gpurun_out/ncu_step.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/launches.csv
for shape in "128 4096 4096" "128 28672 4096"; do
  tag=$(echo $shape | tr ' ' 'x')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 2 -c 1 -f -o gpurun_out/gemm2_$tag python scripts/ncu_gemm.py $shape deferred > gpurun_out/ncu_gemm2_$tag.log 2>&1
  echo "ncu gemm2 $tag exit $?"
done
