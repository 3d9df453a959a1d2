#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 200 --warmup 30 > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err; echo "bench4 exit $?"; cut -c1-300 gpurun_out/n4_bench.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/n4_bench20.json 2> gpurun_out/n4_bench20.err; echo "bench4/20 exit $?"; cut -c1-300 gpurun_out/n4_bench20.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29516 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > gpurun_out/n4_ref.json 2> gpurun_out/n4_ref.err; echo "ref4 exit $?"; cut -c1-200 gpurun_out/n4_ref.json
