#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
B200_SKIP_VLLM=1 timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/last_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/last_tests.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/last_bench20.json 2> gpurun_out/last_bench20.err; echo "bench20 exit $?"; cut -c1-200 gpurun_out/last_bench20.json
