"""Phase timeline of the pair GEMM: %globaltimer stamps per CTA (ns, GPU-wide clock), relative to the earliest CTA entry.
usage: python scripts/gemm_trace.py [deferred]"""
import ctypes as C, math, sys, torch
sys.path.insert(0, ".")
from kubeai_b200 import lib, ops

DEF = len(sys.argv) > 1 and sys.argv[1] == "deferred"
fn = ops.gemm_deferred if DEF else ops.gemm
MARKS = {0: "cta entry", 1: "prologue done (barriers, TMEM)", 12: "producer: dependency wait passed", 10: "mma: first stage landed",
         13: "producer: last TMA issued", 11: "mma: last MMA issued", 2: "epilogue: first accumulator ready",
         3: "epilogue: all tiles stored", 6: "before cluster sync", 7: "exit"}
shapes = [(128, 6144, 4096), (128, 4096, 4096), (128, 28672, 4096), (128, 4096, 14336), (128, 128256, 4096),
          (2048, 28672, 4096), (2048, 4096, 14336)]
for T, N, K in shapes:
    x = torch.randn(T, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    tr = torch.zeros(148, 16, dtype=torch.int64, device="cuda")
    for _ in range(3):
        fn(x, w)
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib().b200_op_gemm_trace(C.c_void_p(tr.data_ptr()))
    e0.record()
    fn(x, w)
    e1.record()
    torch.cuda.synchronize()
    lib().b200_op_gemm_trace(None)
    full = tr.cpu().double()
    full = full[full[:, 0] > 0]
    t0 = full[:, 0].min()
    ideal = (N * K * 2 + T * K * 2 + T * N * 2) / 6.49e12 * 1e6
    print(f"T={T} N={N} K={K}: {len(full)} CTAs, event time {e0.elapsed_time(e1) * 1e3:.1f} us (includes the reducer when deferred), "
          f"HBM-ideal {ideal:.1f} us")
    for i, name in MARKS.items():
        v = full[:, i]
        v = (v[v > 0] - t0) / 1e3
        if len(v):
            print(f"    {name:38s} min {v.min():7.2f}  mean {v.mean():7.2f}  max {v.max():7.2f} us   ({len(v)} stamps)")
