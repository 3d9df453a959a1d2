"""One GEMM shape, a few launches — the target of `ncu --set full` (run under gpurun)."""
import math, sys, torch
sys.path.insert(0, ".")
from kubeai_b200 import ops
T, N, K = (int(a) for a in sys.argv[1:4])
x = torch.randn(T, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
fn = ops.gemm_deferred if len(sys.argv) > 4 and sys.argv[4] == "deferred" else ops.gemm
for _ in range(4):
    flush.zero_()
    fn(x, w)
torch.cuda.synchronize()
