"""Where do the fused large-step form (gemm2 mode 2) and the fp32-segment form differ?  Bitwise comparison of the plain
products at the gate_up shape."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kubeai_b200 import ops
T, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(T, K, generator=g, device="cuda").bfloat16()
w = (torch.randn(N, K, generator=g, device="cuda") / math.sqrt(K)).bfloat16()
a = ops.gemm_deferred(x, w)
b, sch = ops.gemm3(x, w)
torch.cuda.synchronize()
ref = (x.float() @ w.float().t())
d = (a.float() != b.float())
print("schedule", sch, "differing elements", int(d.sum()), "of", d.numel())
print("max |a-ref|", float((a.float() - ref).abs().max()), "max |b-ref|", float((b.float() - ref).abs().max()))
if d.any():
    idx = d.nonzero()
    tt = (idx[:, 0] // 512).tolist(); tile = (idx[:, 1] // 256).tolist()
    from collections import Counter
    c = Counter(zip(tile, tt))
    print("tiles (weight tile, token tile) with differences:", len(c), "of", (N // 256) * ((T + 511) // 512))
    print(sorted(c.items())[:40])
    i = idx[0]
    print("first diff at", i.tolist(), float(a[i[0], i[1]]), float(b[i[0], i[1]]), float(ref[i[0], i[1]]))
    # per-tile: fraction differing and max ulp
    ea = (a.float() - ref).abs(); eb = (b.float() - ref).abs()
    print("mean |a-ref|", float(ea.mean()), "mean |b-ref|", float(eb.mean()))
