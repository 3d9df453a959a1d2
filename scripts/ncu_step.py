"""A few engine steps at the bench's shape (128 sequences, Llama-3-8B) for the ncu launch list:
one prefill step, then decode steps.  `ncu -k regex:... -s 600 -c 600` captures two decode steps."""
import sys
import numpy as np
sys.path.insert(0, ".")
from kubeai_b200.engine import Engine, default_config
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 16
e = Engine(default_config(manual_step=1, max_batched_tokens=2048, max_num_seqs=128, max_model_len=2048, kv_fraction=0.3))
rng = np.random.default_rng(0)
for _ in range(128):  # ragged contexts like the bench's sessions
    e.submit(rng.integers(0, 128000, size=ctx + (_ * 7) % 90).tolist(), max_tokens=int(sys.argv[2]) if len(sys.argv) > 2 else 16)
for i in range(5 + (128 * ctx + 2047) // 2048):
    ran, info = e.step()
    print(i, info.tokens, info.decode_seqs, info.prefill_seqs, round(info.device_us))
e.close()
