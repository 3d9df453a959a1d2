#!/bin/bash
# ncu evidence (1 GPU), all under gpurun_out/:
#  (1) launch list (gpu__time_duration) of decode steps of the engine at the bench's shape (128 seqs, ctx ~440)
#  (2) --set full captures of the dominant kernel (pair GEMM) at a decode and a prefill shape, and of decode attention
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
K='regex:gemm|attn|rmsnorm|rope|silu|embed|argmax|reduce'
# ncu_step.py brackets two steady decode steps (292 launches each) with cudaProfilerStart/Stop
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k "$K" -c 584 --csv \
   --log-file gpurun_out/launches.csv python scripts/ncu_step.py 400 2 > gpurun_out/ncu_step.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/launches.csv
for shp in "128 28672 4096" "2048 28672 4096"; do
  n=$(echo $shp | tr " " x)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 2 -c 1 -o gpurun_out/gemm2_$n -f \
     python scripts/ncu_gemm.py $shp deferred > gpurun_out/ncu_gemm2_$n.log 2>&1
  echo "ncu gemm2 $n exit $?"
done
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:paged_attn_kernel -s 3 -c 1 \
   -o gpurun_out/attn_decode_ctx440 -f python scripts/ncu_step.py 400 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attention exit $?"
ls -la gpurun_out/*.ncu-rep
