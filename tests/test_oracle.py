"""The CPU oracle pinned against the committed HF-transformers golden fixture
(tests/golden/llama_mini.npz, made by oracle/gen_golden.py).  CPU only."""
from pathlib import Path

import numpy as np
import torch

from oracle import llama_oracle as O
from oracle.weights import ModelCfg, bf16_bits_to_f32, f32_to_bf16_bits, init_uniform, make_weights

GOLD = np.load(Path(__file__).parent / "golden" / "llama_mini.npz")


def test_golden_was_made_for_this_config():
    c = ModelCfg()
    assert GOLD["cfg"].tolist() == [c.num_layers, c.hidden, c.q_heads, c.kv_heads, c.intermediate, c.vocab,
                                    c.max_model_len, c.seed]


def test_oracle_logits_match_hf():
    cfg = ModelCfg()
    o = O.LlamaOracle(cfg, make_weights(cfg))
    logits, _ = o.forward(GOLD["ids"])
    l = logits.numpy()
    for key in ("logits_fp32", "logits_bf16"):
        ref = GOLD[key]
        err = np.abs(l - ref)
        # bf16 pipeline vs HF: vLLM's bf16 tolerance (rtol 1.6e-2) on |logit| up to ~16 plus one ulp
        assert (err <= 0.15 + 1.6e-2 * np.abs(ref)).all(), (key, err.max())  # 0.15 ~ two bf16 ulps at |logit| 8..16
        assert err.mean() < 0.04
    top2 = np.sort(GOLD["logits_fp32"], axis=-1)[:, -2:]
    solid = (top2[:, 1] - top2[:, 0]) > 0.25
    assert (l.argmax(-1)[solid] == GOLD["logits_fp32"].argmax(-1)[solid]).all()


def test_oracle_with_llama3_rope_scaling_matches_hf():
    """Llama-3.1 style frequency scaling (b200_config.rope_scaling_type 2): pinned on HF's own implementation."""
    from oracle.gen_golden import ROPE3
    gold = np.load(Path(__file__).parent / "golden" / "llama_mini_rope3.npz")
    cfg = ModelCfg(**ROPE3)
    o = O.LlamaOracle(cfg, make_weights(cfg))
    l = o.forward(gold["ids"])[0].numpy()
    err = np.abs(l - gold["logits_fp32"])
    assert (err <= 0.15 + 1.6e-2 * np.abs(gold["logits_fp32"])).all() and err.mean() < 0.04, err.max()
    toks, rows = o.generate(gold["ids"].tolist(), 12)
    assert toks == gold["greedy"].tolist()
    # and the scaling is not a no-op: the unscaled model gives different logits
    l0 = O.LlamaOracle(ModelCfg(), make_weights(ModelCfg())).forward(gold["ids"])[0].numpy()
    assert np.abs(l0 - l).max() > 1.0


def test_oracle_greedy_matches_hf_and_kv_cache_path_is_consistent():
    cfg = ModelCfg()
    o = O.LlamaOracle(cfg, make_weights(cfg))
    ids = GOLD["ids"].tolist()
    toks, rows = o.generate(ids, 12)
    assert toks == GOLD["greedy"].tolist()
    # incremental (KV-cached) forward == full forward up to fp32 summation order inside BLAS
    full, _ = o.forward(ids + toks[:5])
    l1, kv = o.forward(ids)
    l2, _ = o.forward(toks[:5], kv_prefix=kv, pos0=len(ids))
    assert torch.allclose(full[len(ids):], l2, atol=0.15, rtol=1.6e-2)
    assert torch.equal(full[len(ids):].argmax(-1), l2.argmax(-1))


def test_bf16_helpers_and_init_replica():
    x = np.array([1.0, 1.00390625, 1.01171875, -3.0e-40, 65504.0, 1e-3], dtype=np.float32)
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(bf16_bits_to_f32(f32_to_bf16_bits(x)), want)
    a = bf16_bits_to_f32(init_uniform(1 << 16, 123, 1.0, 0.0))
    assert abs(a.mean()) < 0.02 and abs(a.std() - 1.0) < 0.02 and a.min() >= -1.74 and a.max() <= 1.74
    assert np.array_equal(init_uniform(100, 123, 1.0, 0.0), init_uniform(200, 123, 1.0, 0.0)[:100])
    assert not np.array_equal(init_uniform(100, 123, 1.0, 0.0), init_uniform(100, 124, 1.0, 0.0))


def test_op_restatements_follow_vllm_rounding_points():
    g = torch.Generator().manual_seed(0)
    x, res, w = (O.r(torch.randn(3, 64, generator=g)) for _ in range(3))
    y, nr = O.fused_add_rms_norm(x, res, w[0], 1e-5)
    # same thing the way HF writes it in real bf16 tensors (modeling_llama.py:325 residual + hidden in bf16, then
    # LlamaRMSNorm :62-67) == vLLM's _C fused_add_rms_norm (sum in scalar_t, norm of the rounded sum)
    xb, rb, wb = x.bfloat16(), res.bfloat16(), w[0].bfloat16()
    r2 = rb + xb
    xf = r2.float()
    var = xf.pow(2).mean(dim=-1, keepdim=True)
    y2 = wb * (xf * torch.rsqrt(var + 1e-5)).to(torch.bfloat16)
    assert torch.equal(y, y2.float()) and torch.equal(nr, r2.float())
    gu = O.r(torch.randn(4, 32, generator=g) * 3)
    d = 16
    ref = (torch.nn.functional.silu(gu.bfloat16()[..., :d]) * gu.bfloat16()[..., d:]).float()
    assert torch.allclose(O.silu_and_mul(gu), ref, atol=0, rtol=8e-3)   # activation.py:138-141
