#!/bin/bash
# 2 GPUs: live routing in one process (LeastLoad vs PrefixHash), then the torchrun bench at N=2 for the aggregate comparison
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
nvidia-smi -L
timeout 900 python scripts/routing_run.py --gpus 2 --out gpurun_out/s_routing2.json 2> gpurun_out/s_routing2.err | cut -c1-900; echo "routing exit $?"; tail -3 gpurun_out/s_routing2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 30 > gpurun_out/s_bench2.json 2> gpurun_out/s_bench2.err; echo "bench2 exit $?"; cut -c1-400 gpurun_out/s_bench2.json
