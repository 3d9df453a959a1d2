"""TEST/BENCH INFRASTRUCTURE — the CPU path timed beside the GPU engine (bench.py cpu_baseline leg
and `bench.py --impl reference`).

The literal reference path (KubeAI Go proxy -> Ollama/vLLM CPU pod) cannot run here (no Go, no
Ollama, no weights), so the CPU number is the oracle's restatement of the backend
math (oracle/llama_oracle.py) on the host cores: one decode step of the Llama-3-8B shape, sampled as
`layers` decoder layers + the lm_head at batch `batch`, mean context `ctx`, extrapolated to 32
layers.  kind = "port".
"""
from __future__ import annotations

import math
import os
import time

import torch

from . import llama_oracle as O

LLAMA3_8B = dict(num_layers=32, hidden=4096, q_heads=32, kv_heads=8, intermediate=14336, vocab=128256)


class CpuDecodeSample:
    def __init__(self, shape=None, batch=128, ctx=430, seed=0, with_head=True):
        self.s = dict(LLAMA3_8B if shape is None else shape)
        self.batch, self.ctx = batch, ctx
        H, I, D = self.s["hidden"], self.s["intermediate"], 128
        Hq, Hkv = self.s["q_heads"], self.s["kv_heads"]
        g = torch.Generator().manual_seed(seed)
        rnd = lambda *sh, sc=1.0: O.r(torch.randn(*sh, generator=g) * sc)
        self.w = dict(
            wqkv=rnd((Hq + 2 * Hkv) * D, H, sc=1 / math.sqrt(H)), wo=rnd(H, Hq * D, sc=1 / math.sqrt(Hq * D)),
            wgu=rnd(2 * I, H, sc=1 / math.sqrt(H)), wdown=rnd(H, I, sc=1 / math.sqrt(I)),
            norm1=O.r(torch.ones(H)), norm2=O.r(torch.ones(H)))
        self.lm_head = rnd(self.s["vocab"], H, sc=1 / math.sqrt(H)) if with_head else None
        self.x = rnd(batch, H)
        self.res = rnd(batch, H)
        self.k = rnd(ctx, Hkv, D)
        self.v = rnd(ctx, Hkv, D)
        self.cs = O.r(torch.cat([torch.cos(torch.arange(ctx + 1)[:, None] * torch.ones(64)[None]),
                                 torch.sin(torch.arange(ctx + 1)[:, None] * torch.ones(64)[None])], dim=-1))

    @torch.no_grad()
    def layer(self):
        s, w, D = self.s, self.w, 128
        Hq, Hkv, B = s["q_heads"], s["kv_heads"], self.batch
        h, res = O.fused_add_rms_norm(self.x, self.res, w["norm1"], 1e-5)
        qkv = O.gemm(h, w["wqkv"])
        pos = torch.full((B,), self.ctx, dtype=torch.long)
        q = O.rope_neox(qkv[:, :Hq * D].reshape(B, Hq, D), pos, self.cs)
        # every sequence attends over its own `ctx`-token cache (same K/V tensor reused: the arithmetic is identical)
        g = Hq // Hkv
        kk = self.k.repeat_interleave(g, dim=1)
        vv = self.v.repeat_interleave(g, dim=1)
        sc = torch.einsum("bhd,khd->bhk", q, kk) * (D ** -0.5)
        a = O.r(torch.einsum("bhk,khd->bhd", torch.softmax(sc, dim=-1), vv)).reshape(B, Hq * D)
        o = O.gemm(a, w["wo"])
        h, res = O.fused_add_rms_norm(o, res, w["norm2"], 1e-5)
        act = O.silu_and_mul(O.gemm(h, w["wgu"]))
        return O.gemm(act, w["wdown"])

    @torch.no_grad()
    def head(self):
        return torch.argmax(O.gemm(self.x, self.lm_head), dim=-1)


def measure(steps=3, warmup=1, batch=128, ctx=430, shape=None, budget_s=30.0):
    """Returns dict(tok_s, ms_per_step, cores, sample).  One 'step' = one sampled decoder layer."""
    cores = os.cpu_count() or 1
    smp = CpuDecodeSample(shape, batch, ctx)
    L = smp.s["num_layers"]
    # use the thread count that is fastest on this host (all cores is often slower for these BLAS shapes on big
    # multi-socket hosts): one calibration layer per candidate, best one is kept and reported as `cores`
    cands = sorted({c for c in (cores, cores // 2, cores // 4, 32, 16, 8) if 1 <= c <= cores}, reverse=True)
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        smp.layer()                      # warm
        t0 = time.perf_counter()
        smp.layer()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
        if time.perf_counter() - t0 > budget_s / 3:
            break
    torch.set_num_threads(best)
    for _ in range(warmup):
        smp.layer()
    ts = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        smp.layer()
        ts.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    t0 = time.perf_counter()
    smp.head()
    t_head = time.perf_counter() - t0
    t_layer = sum(ts) / len(ts)
    step = L * t_layer + t_head
    return dict(tok_s=batch / step, ms_per_step=step * 1e3, cores=torch.get_num_threads(), layer_s=t_layer, head_s=t_head,
                sample=f"{len(ts)} x 1 of {L} decoder layers + lm_head at batch {batch}, ctx {ctx}, extrapolated to {L} layers "
                       f"(torch-CPU fp32 math over bf16-rounded tensors, {torch.get_num_threads()} threads)")
