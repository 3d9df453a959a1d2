"""Does a weight-streaming GEMM run slower when every launch reads a different 235 MB weight tensor out of a 15 GB set (the
decode step's access pattern) than when it re-reads the same tensor (the microbenchmark's pattern)?  L2 cannot hold either;
the difference would be address translation / DRAM page state, not cache hits."""
import math
import sys
import time

import torch

sys.path.insert(0, ".")
from kubeai_b200 import ops  # noqa: E402

T, N, K = 128, 28672, 4096
x = torch.randn(T, K, device="cuda").bfloat16()
nbuf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ws = [(torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16() for _ in range(nbuf)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def run(seq, label, fn):
    for w in seq[:4]:
        fn(x, w)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for w in seq:
        fn(x, w)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / len(seq)
    print(f"{label:58s} {us:7.1f} us/launch  {N * K * 2 / us / 1e6:6.2f} TB/s")


g3 = lambda x, w: ops.gemm3(x, w)
cb = lambda x, w: torch.nn.functional.linear(x, w)
print(f"{nbuf} weight tensors of {N * K * 2 / 1e6:.0f} MB = {nbuf * N * K * 2 / 1e9:.1f} GB; back-to-back launches, no flush (every tensor > L2)")
for name, fn in (("gemm3", g3), ("cuBLAS", cb)):
    run([ws[0], ws[1]] * 32, f"{name}: two tensors alternating (470 MB working set)", fn)
    run(ws[:8] * 8, f"{name}: 8 tensors round robin (1.9 GB)", fn)
    run(ws[:32] * 2, f"{name}: 32 tensors round robin (7.5 GB)", fn)
    run(ws, f"{name}: {nbuf} tensors round robin ({nbuf * 0.235:.1f} GB)", fn)
