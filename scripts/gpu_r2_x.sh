#!/bin/bash
# 8 GPUs: live routing in one process (LeastLoad vs PrefixHash over 8 replicas, 1024 concurrent sessions), then torchrun bench at N=8
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
nvidia-smi -L | wc -l; nproc
timeout 1200 python scripts/routing_run.py --gpus 8 --out gpurun_out/x_routing8.json 2> gpurun_out/x_routing8.err | cut -c1-1200; echo "routing exit $?"; tail -3 gpurun_out/x_routing8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 200 --warmup 30 > gpurun_out/x_bench8.json 2> gpurun_out/x_bench8.err; echo "bench8 exit $?"; cut -c1-400 gpurun_out/x_bench8.json
