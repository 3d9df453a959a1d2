#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_loader.py tests/test_server_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/last2_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/last2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/last2_bench20.json 2> gpurun_out/last2_bench20.err; echo "bench20 exit $?"; cut -c1-220 gpurun_out/last2_bench20.json
