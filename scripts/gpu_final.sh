#!/bin/bash
# the round-end sequence on one GPU: full GPU suite (incl. the vLLM parity test), smoke, the driver's bench command, the default bench, the reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 2400 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/final_tests.log 2>&1; echo "pytest -m gpu exit $?"; tail -6 gpurun_out/final_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench20.json 2> gpurun_out/final_bench20.err; echo "bench20 exit $?"; cut -c1-420 gpurun_out/final_bench20.json
timeout 600 python bench.py > gpurun_out/final_bench200.json 2> gpurun_out/final_bench200.err; echo "bench200 exit $?"; cut -c1-300 gpurun_out/final_bench200.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref exit $?"; cut -c1-260 gpurun_out/final_ref.json
