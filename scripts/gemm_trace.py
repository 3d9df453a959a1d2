"""Phase timeline of the pair GEMM (clock64 stamps per CTA), averaged over CTAs."""
import ctypes as C, math, sys, torch
sys.path.insert(0, ".")
from kubeai_b200 import lib, ops
names = ["entry->prologue", "prologue->first tile done", "first tile->main loop+partials done", "->peers visible (spin)",
         "->bulk pull landed", "->reduce+store done", "->cluster sync+dealloc"]
DEF = len(sys.argv) > 1 and sys.argv[1] == "deferred"
fn = ops.gemm_deferred if DEF else ops.gemm
shapes = [(128, 4096, 4096), (128, 6144, 4096), (128, 28672, 4096), (128, 4096, 14336), (384, 28672, 4096)]
if DEF:
    shapes = [(128, 28672, 4096), (128, 4096, 14336), (2048, 4096, 4096), (2048, 6144, 4096), (2048, 28672, 4096), (2048, 4096, 14336)]
for T, N, K in shapes:
    x = torch.randn(T, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    tr = torch.zeros(148, 16, dtype=torch.int64, device="cuda")
    for _ in range(3):
        fn(x, w)
    torch.cuda.synchronize()
    lib().b200_op_gemm_trace(C.c_void_p(tr.data_ptr()))
    fn(x, w)
    torch.cuda.synchronize()
    lib().b200_op_gemm_trace(None)
    full = tr.cpu().double()
    full = full[full[:, 0] > 0]
    rel = lambda i: ((full[:, i] - full[:, 0])[full[:, i] > 0] / 1.9e3)
    ex = {k: rel(i) for k, i in dict(producer_dep_wait_done=12, producer_last_tma_issued=13, mma_first_data=10, mma_last_issue=11,
                                      reduce_loops_done=8, finish_atomics_done=9).items()}
    print("   since entry (us, mean/max): " + "  ".join(f"{k} {v.mean():.2f}/{v.max():.2f}" for k, v in ex.items() if len(v)))
    t = full[:, :8].clone()
    for i in range(1, 8):                      # missing stamps (no fix-up on that CTA): carry forward
        t[:, i] = torch.where(t[:, i] > 0, t[:, i], t[:, i - 1])
    d = (t[:, 1:] - t[:, :-1]) / 1.9e3        # us at ~1.9 GHz
    print(f"T={T} N={N} K={K}: total {((t[:,7]-t[:,0]).mean()/1.9e3):.1f} us (max {((t[:,7]-t[:,0]).max()/1.9e3):.1f}) over {len(t)} CTAs")
    for n, v, m in zip(names, d.mean(0).tolist(), d.max(0).values.tolist()):
        print(f"    {n:40s} mean {v:6.2f} us   max {m:6.2f} us")
