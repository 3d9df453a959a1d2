"""Timeline of the projection-chain launches of one decode step (Llama-3-8B shape, 128 sequences): per phase, when the pairs
finish their segments, how long the grid barriers and the elementwise phases take, when the next phase's first MMA starts."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kubeai_b200 import lib  # noqa: E402
from kubeai_b200.engine import Engine, default_config  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 8
e = Engine(default_config(manual_step=1, num_layers=layers, max_num_seqs=128, max_batched_tokens=1536, max_model_len=2048, kv_fraction=0.3))
rng = np.random.default_rng(0)
rids = [e.submit(rng.integers(0, 128000, size=int(rng.integers(380, 480))).tolist(), max_tokens=400) for _ in range(128)]
for i in range(120):
    ran, info = e.step()
    if info.decode_seqs == 128 and info.prefill_seqs == 0:
        break
buf = torch.zeros(40 * 148 * 32, dtype=torch.int64, device="cuda")
lib().b200_op_gemm_trace(C.c_void_p(buf.data_ptr()))
ran, info = e.step()
torch.cuda.synchronize()
lib().b200_op_gemm_trace(C.c_void_p(0))
print("traced step: T=%d decode=%d device %.3f ms (%d layers)" % (info.tokens, info.decode_seqs, info.device_us / 1e3, layers))
t = buf.cpu().numpy().reshape(40, 148, 32).astype(np.float64)
names = {30: "CTA start", 24: "O first token tile", 16: "O first MMA", 17: "O last MMA", 0: "O segments out", 1: "barrier 1 passed", 2: "resadd+norm done",
         25: "GU first token tile", 18: "GU first MMA", 19: "GU last MMA", 4: "GU segments out", 26: "DOWN first token tile", 20: "DOWN first MMA",
         21: "DOWN last MMA", 8: "DOWN segments out", 9: "barrier 4 passed", 10: "resadd+norm done", 27: "QKV first token tile", 22: "QKV first MMA",
         23: "QKV last MMA", 12: "QKV segments out", 13: "barrier 6 passed", 14: "rope+kv done", 31: "CTA end"}
order = [30, 24, 16, 17, 0, 1, 2, 25, 18, 19, 4, 26, 20, 21, 8, 9, 10, 27, 22, 23, 12, 13, 14, 31]
for slot in (1, 2):   # second and third chain launch of the step
    x = t[slot]
    t0 = x[:, 30][x[:, 30] > 0].min()
    print(f"-- chain launch {slot}: median (min .. max) over CTAs, us after the first CTA start")
    for k in order:
        v = x[:, k]
        v = v[v > 0]
        if len(v) == 0:
            continue
        v = (v - t0) / 1e3
        print(f"   {names[k]:24s} {np.median(v):7.1f}  ({v.min():6.1f} .. {v.max():6.1f})")
e.close()
