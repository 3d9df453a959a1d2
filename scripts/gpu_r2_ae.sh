#!/bin/bash
# ncu launch list of the bench command's own timed region (the driver's flags), then the same command without ncu for the shares
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
K='regex:gemm|attn|rmsnorm|rope|silu|embed|argmax|reduce|chain'
B200_BENCH_CUDA_PROFILER=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k "$K" -c 6000 --csv \
   --log-file gpurun_out/r02_launches_bench20.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ae_bench_under_ncu.json 2> gpurun_out/ae_bench_under_ncu.err
echo "ncu bench exit $?"; wc -l gpurun_out/r02_launches_bench20.csv
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ae_bench20.json 2> gpurun_out/ae_bench20.err; echo "bench exit $?"; cut -c1-300 gpurun_out/ae_bench20.json
