import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kubeai_b200 import ops
T, I, K = 1500, 14336, 4096
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(T, K, generator=g, device="cuda").bfloat16()
w = (torch.randn(2 * I, K, generator=g, device="cuda") / math.sqrt(K)).bfloat16()   # physical (interleaved) layout
plain, _ = ops.gemm3(x, w)
act, _ = ops.gemm3(x, w, epi=ops.EPI_SILU)
p = plain.reshape(T, I // 64, 2, 64)
gu = torch.cat([p[:, :, 0].reshape(T, I), p[:, :, 1].reshape(T, I)], dim=1).contiguous()
act2 = ops.silu_mul(gu)
torch.cuda.synchronize()
d = act.float() != act2.float()
print("silu epilogue vs silu kernel: differing", int(d.sum()), "of", d.numel())
if d.any():
    i = d.nonzero()[0]
    print(i.tolist(), float(act[i[0], i[1]]), float(act2[i[0], i[1]]), float(gu[i[0], i[1]]), float(gu[i[0], I + i[1]]))

# engine level: the burst step, each mode twice
from kubeai_b200.engine import Engine, default_config
SHAPE = dict(num_layers=1, hidden=4096, q_heads=32, kv_heads=8, intermediate=14336, vocab=128256, max_model_len=2048)
rng = np.random.default_rng(77)
lens = [700, 450, 300, 50]
prompts = [rng.integers(0, 128256, size=n).tolist() for n in lens]
res = {}
for mode in ("1", "0", "1", "0"):
    os.environ["B200_FUSED_PREFILL"] = mode
    with Engine(default_config(manual_step=1, max_num_seqs=8, max_batched_tokens=1536, num_kv_blocks=256, **SHAPE)) as e:
        e.set_keep_logits(True)
        rids = [e.submit(p, max_tokens=2) for p in prompts]
        ran, info = e.step()
        got = e.read_logits(len(lens)).copy()
        for r in rids:
            e.release(r)
    res.setdefault(mode, []).append(got)
for m in ("1", "0"):
    print("mode", m, "run-to-run identical:", np.array_equal(res[m][0], res[m][1]))
dd = res["1"][0] != res["0"][0]
print("fused vs segments differing logits:", int(dd.sum()), "of", dd.size, "per row", dd.sum(-1).tolist(), "max abs", float(np.abs(res["1"][0] - res["0"][0]).max()))
