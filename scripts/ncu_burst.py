"""One prefill-burst step at the bench's shape for ncu (run with --profile-from-start off): 118 sequences decoding at
context ~450 plus 10 requests whose ~130 new tokens each follow a ~320-token cached prefix (a new turn of a conversation)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from kubeai_b200.engine import Engine, default_config
e = Engine(default_config(manual_step=1, max_batched_tokens=1536, max_num_seqs=128, max_model_len=2048, kv_fraction=0.3))
rng = np.random.default_rng(0)
convs = [rng.integers(0, 128000, size=320 + (i * 7) % 60).tolist() for i in range(10)]
for c in convs:                                  # first turns: fill the prefix cache
    e.submit(c, max_tokens=2)
for i in range(118):
    e.submit(rng.integers(0, 128000, size=400 + (i * 7) % 90).tolist(), max_tokens=200)
for _ in range(60):
    ran, info = e.step()
    if info.prefill_seqs == 0 and info.decode_seqs == 118:
        break
for c in convs:                                  # second turns: cached prefix + ~130 new tokens
    e.submit(c + rng.integers(0, 128000, size=130).tolist(), max_tokens=8)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ran, info = e.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("burst step: tokens", info.tokens, "decode", info.decode_seqs, "prefill", info.prefill_seqs, "device us", round(info.device_us),
      "cached prompt tokens", e.stats().cached_prompt_tokens)
e.close()
