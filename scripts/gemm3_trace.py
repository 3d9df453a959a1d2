"""Phase timeline of one gemm3 launch per shape (%globaltimer stamps per CTA, b200_op_gemm_trace): where a decode-shape
GEMM spends its time — prologue, pipeline fill, mainloop (streaming rate), partial exchange, epilogue."""
import ctypes as C
import math
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kubeai_b200 import lib, ops  # noqa: E402

flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def trace(name, T, N, K, force=0, pro=0, epi=0):
    x = torch.randn(T, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    kw = {}
    if pro:
        kw = dict(pro=ops.PRO_NORM, ssq_in=x.float().reshape(T, K // 128, 128).pow(2).sum(-1).contiguous(),
                  norm_w=torch.ones(K, device="cuda").bfloat16())
    if epi == ops.EPI_RESADD:
        kw["out"] = torch.randn(T, N, device="cuda").bfloat16()
    buf = torch.zeros(296 * 8, dtype=torch.int64, device="cuda")
    for rep in range(3):
        flush.zero_()
        torch.cuda.synchronize()
        lib().b200_op_gemm_trace(C.c_void_p(buf.data_ptr()))
        _, sch = ops.gemm3(x, w, force=force, epi=epi, **kw)
        torch.cuda.synchronize()
        lib().b200_op_gemm_trace(C.c_void_p(0))
    g = sch[2]
    t = buf.cpu().numpy().reshape(-1, 8)[:g].astype(np.float64)
    t0 = t[:, 0].min()
    t = (t - t0) / 1e3      # us
    lead = t[::2]           # leader CTAs carry the MMA stamps
    med = lambda a: float(np.median(a))
    span = t[:, 7].max()
    main = lead[:, 3] - lead[:, 2]
    per_cta_bytes = N * K * 2 / g
    print(f"{name:22s} T={T:3d} sch={sch} span {span:6.1f} us | start spread {t[:,0].max():4.1f} | prologue {med(t[:,1]-t[:,0]):4.1f} | "
          f"first MMA at {med(lead[:,2]):5.1f} | mainloop {med(main):5.1f} ({per_cta_bytes/ med(main)/1e3:5.1f} GB/s per SM, "
          f"{N*K*2/1e3/med(main)/1e3:5.2f} TB/s) | last MMA at {med(lead[:,3]):5.1f} | peers wait {med(t[:,6]-t[:,5]):4.1f} | "
          f"epilogue end {med(t[:,7]):5.1f} (max {span:5.1f})")


if __name__ == "__main__":
    import os
    if len(sys.argv) > 1 and sys.argv[1] == "norm":
        for dbg in (0, 1, 2, 4, 6, 7):
            os.environ["B200_GEMM3_DBG"] = str(dbg)
            print("dbg", dbg)
            trace("qkv S=2 norm", 128, 6144, 4096, 2, pro=1)
            trace("gate_up streamK norm", 128, 28672, 4096, 0, pro=1)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "stages":
        print("B200_GEMM3_STAGES =", os.environ.get("B200_GEMM3_STAGES", "6 (default)"))
        trace("o S=1", 128, 4096, 4096, 1)
        trace("down S=1", 128, 4096, 14336, 1)
        trace("qkv S=1", 128, 6144, 4096, 1)
        trace("gate_up S=1 (2 waves)", 128, 28672, 4096, 1) if False else None
        if os.environ.get("B200_GEMM3_STAGES") != "8":
            trace("o S=3", 128, 4096, 4096, 3)
            trace("gate_up streamK", 128, 28672, 4096, 0)
        sys.exit(0)
    for T in (128,):
        trace("o S=1", T, 4096, 4096, 1)
        trace("o S=2", T, 4096, 4096, 2)
        trace("o S=3", T, 4096, 4096, 3)
        trace("o S=3 resadd", T, 4096, 4096, 3, epi=ops.EPI_RESADD)
        trace("down S=1", T, 4096, 14336, 1)
        trace("down S=3", T, 4096, 14336, 3)
        trace("qkv S=2", T, 6144, 4096, 2)
        trace("qkv S=2 norm", T, 6144, 4096, 2, pro=1)
        trace("gate_up streamK", T, 28672, 4096, 0)
        trace("gate_up streamK norm", T, 28672, 4096, 0, pro=1)
        trace("gate_up norm+silu", T, 28672, 4096, 0, pro=1, epi=ops.EPI_SILU)
        trace("lm_head streamK", T, 128256, 4096, 0)
    trace("o S=3", 16, 4096, 4096, 3)
