#!/bin/bash
# ncu evidence: (1) full capture of the dominant kernel at a decode shape, (2) launch list of a short bench run.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
for shape in "128 4096 4096" "128 28672 4096"; do
  tag=$(echo $shape | tr ' ' 'x')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 2 -f -o gpurun_out/gemm_$tag python scripts/ncu_gemm.py $shape > gpurun_out/ncu_gemm_$tag.log 2>&1
  echo "ncu gemm $tag exit $?"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 30 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launches exit $?"; tail -3 gpurun_out/launches.csv | cut -c1-300
