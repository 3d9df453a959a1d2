"""Engine steps at the bench's shape (128 sequences, Llama-3-8B, ragged contexts around `ctx`) for ncu: the prefill phase
runs unprofiled, then cudaProfilerStart() brackets `n` steady decode steps (run ncu with --profile-from-start off).
usage: python scripts/ncu_step.py [ctx=400] [decode_steps=2]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from kubeai_b200.engine import Engine, default_config
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 400
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
e = Engine(default_config(manual_step=1, max_batched_tokens=2048, max_num_seqs=128, max_model_len=2048, kv_fraction=0.3))
rng = np.random.default_rng(0)
for i in range(128):  # ragged contexts like the bench's sessions
    e.submit(rng.integers(0, 128000, size=ctx + (i * 7) % 90).tolist(), max_tokens=64)
warm = 0
while warm < 4:
    ran, info = e.step()
    if info.prefill_seqs == 0 and info.decode_seqs == 128:
        warm += 1
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(nsteps):
    ran, info = e.step()
    print(i, info.tokens, info.decode_seqs, info.prefill_seqs, round(info.device_us))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
e.close()
