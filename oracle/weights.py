"""TEST INFRASTRUCTURE — numpy replica of the engine's seeded weight init.

The engine fills its weights on the GPU with a counter-hash generator
(kubeai_b200/csrc/elementwise.cu: init_uniform_kernel; tensor list and scales in
kubeai_b200/csrc/engine.cu: Engine::alloc_all).  This file restates it in numpy so that the CPU
oracle, HF transformers and the engine all see bit-identical bf16 weights without shipping
checkpoints (there is no network; SURVEY.md §8d "Synthetic weights").
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass

import numpy as np

P1, P2, P3, P4, P5 = (0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9,
                      0x85EBCA77C2B2AE63, 0x27D4EB2F165667C5)
M64 = (1 << 64) - 1


def _rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & M64


def _round(acc, inp):
    return (_rotl((acc + inp * P2) & M64, 31) * P1) & M64


def _merge(acc, v):
    return (((acc ^ _round(0, v)) * P1) + P4) & M64


def xxh64(data: bytes, seed: int = 0) -> int:
    """XXH64 from the published specification (== cespare/xxhash Sum64 for seed 0,
    /root/reference/internal/loadbalancer/balance_chwbl.go:140-142)."""
    n = len(data)
    p = 0
    if n >= 32:
        v1, v2, v3, v4 = (seed + P1 + P2) & M64, (seed + P2) & M64, seed & M64, (seed - P1) & M64
        while p <= n - 32:
            a, b, c, d = struct.unpack_from("<4Q", data, p)
            v1, v2, v3, v4 = _round(v1, a), _round(v2, b), _round(v3, c), _round(v4, d)
            p += 32
        h = (_rotl(v1, 1) + _rotl(v2, 7) + _rotl(v3, 12) + _rotl(v4, 18)) & M64
        for v in (v1, v2, v3, v4):
            h = _merge(h, v)
    else:
        h = (seed + P5) & M64
    h = (h + n) & M64
    while p + 8 <= n:
        (k,) = struct.unpack_from("<Q", data, p)
        h ^= _round(0, k)
        h = (_rotl(h, 27) * P1 + P4) & M64
        p += 8
    if p + 4 <= n:
        (k,) = struct.unpack_from("<I", data, p)
        h ^= (k * P1) & M64
        h = (_rotl(h, 23) * P2 + P3) & M64
        p += 4
    while p < n:
        h ^= (data[p] * P5) & M64
        h = (_rotl(h, 11) * P1) & M64
        p += 1
    h ^= h >> 33
    h = (h * P2) & M64
    h ^= h >> 29
    h = (h * P3) & M64
    h ^= h >> 32
    return h


def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x * np.uint32(0x7FEB352D)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x846CA68B)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + r) >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


def init_uniform(n: int, seed: int, scale: float, offset: float) -> np.ndarray:
    """bf16 bit patterns of element i = offset + scale * u_i, u uniform(-sqrt3, sqrt3); n < 2^32."""
    assert n < (1 << 32)
    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.uint32)
        s0 = _mix32(np.array([seed & 0xFFFFFFFF], dtype=np.uint32))  # i >> 32 == 0
        h = _mix32(i ^ s0)
    u = (h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    u = (u + np.float32(-0.5)).astype(np.float32)
    u = (u * np.float32(3.4641016151)).astype(np.float32)
    v = (np.float32(offset) + (np.float32(scale) * u).astype(np.float32)).astype(np.float32)
    return f32_to_bf16_bits(v)


@dataclass
class ModelCfg:
    num_layers: int = 2
    hidden: int = 512
    q_heads: int = 4
    kv_heads: int = 1
    intermediate: int = 1024
    vocab: int = 512
    rms_eps: float = 1e-5
    rope_theta: float = 500000.0
    max_model_len: int = 256
    seed: int = 0
    init_scale: float = 4.0
    head_dim: int = 128
    # RoPE frequency scaling (b200_config.rope_scaling_type: 0 none, 1 linear, 2 llama3)
    rope_scaling_type: int = 0
    rope_factor: float = 1.0
    rope_low_freq_factor: float = 1.0
    rope_high_freq_factor: float = 4.0
    rope_original_max_pos: int = 8192

    @property
    def qkv_rows(self):
        return (self.q_heads + 2 * self.kv_heads) * self.head_dim


def tensor_specs(cfg: ModelCfg):
    """(name, shape, scale, offset) in the engine's allocation order (engine.cu Engine::alloc_all)."""
    H, I, D = cfg.hidden, cfg.intermediate, cfg.head_dim
    s_h = np.float32(1.0) / np.sqrt(np.float32(H))
    s_i = np.float32(1.0) / np.sqrt(np.float32(I))
    s_a = np.float32(1.0) / np.sqrt(np.float32(cfg.q_heads * D))
    out = [("embed", (cfg.vocab, H), 1.0, 0.0)]
    for l in range(cfg.num_layers):
        p = f"layers.{l}."
        out += [
            (p + "wqkv", (cfg.qkv_rows, H), float(s_h), 0.0),
            (p + "wo", (H, cfg.q_heads * D), float(s_a), 0.0),
            (p + "wgu", (2 * I, H), float(s_h), 0.0),
            (p + "wdown", (H, I), float(s_i), 0.0),
            (p + "norm1", (H,), 0.1, 1.0),
            (p + "norm2", (H,), 0.1, 1.0),
        ]
    out.append(("final_norm", (H,), 0.1, 1.0))
    out.append(("lm_head", (cfg.vocab, H), float(np.float32(s_h) * np.float32(cfg.init_scale)), 0.0))
    return out


def make_weights(cfg: ModelCfg) -> dict:
    """name -> uint16 bf16 bit patterns with the tensor's shape."""
    w = {}
    for name, shape, scale, offset in tensor_specs(cfg):
        seed = xxh64(name.encode(), cfg.seed) & 0xFFFFFFFF
        n = int(np.prod(shape))
        w[name] = init_uniform(n, seed, scale, offset).reshape(shape)
    return w


def cos_sin_cache(cfg: ModelCfg) -> np.ndarray:
    """[max_model_len, 128] fp32 values (cos | sin) before the bf16 cast
    (vllm/model_executor/layers/rotary_embedding/base.py:70-92)."""
    D = cfg.head_dim
    inv = (np.float32(1.0) / np.power(np.float32(cfg.rope_theta),
                                       np.arange(0, D, 2, dtype=np.float32) / np.float32(D))).astype(np.float32)
    t = np.arange(cfg.max_model_len, dtype=np.float32)
    if cfg.rope_scaling_type == 1:      # vllm rotary_embedding/linear_scaling_rope.py: t / factor
        t = (t / np.float32(cfg.rope_factor)).astype(np.float32)
    elif cfg.rope_scaling_type == 2:    # vllm rotary_embedding/llama3_rope.py:37-54
        orig, lo, hi, fac = (np.float32(x) for x in (cfg.rope_original_max_pos, cfg.rope_low_freq_factor,
                                                     cfg.rope_high_freq_factor, cfg.rope_factor))
        wl = (np.float32(2.0 * math.pi) / inv).astype(np.float32)
        smooth = ((orig / wl - lo) / (hi - lo)).astype(np.float32) if lo != hi else np.zeros_like(inv)
        mid = ((np.float32(1.0) - smooth) * inv / fac + smooth * inv).astype(np.float32)
        inv = np.where(wl < orig / hi, inv, np.where(wl > orig / lo, inv / fac, mid)).astype(np.float32)
    f = np.outer(t, inv).astype(np.float32)
    return np.concatenate([np.cos(f), np.sin(f)], axis=-1).astype(np.float32)
