"""TEST INFRASTRUCTURE — generate tests/golden/llama_mini.npz with HF transformers (run in the build
container; the GPU box only reads the committed fixture).

    python -m oracle.gen_golden

Builds a 2-layer Llama (head_dim 128, GQA 4:1) whose weights come from oracle/weights.py (the numpy
replica of the engine's seeded init), runs transformers' LlamaForCausalLM on CPU in float32 over
bf16-valued weights and in bfloat16, and stores the prompt ids plus both logit sets.  The oracle
(oracle/llama_oracle.py) must agree with these; tests/test_oracle.py checks it.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

from .weights import ModelCfg, bf16_bits_to_f32, make_weights

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "llama_mini.npz"
OUT_ROPE3 = OUT.with_name("llama_mini_rope3.npz")      # same model with Llama-3.1 style rope scaling
ROPE3 = dict(rope_scaling_type=2, rope_factor=8.0, rope_low_freq_factor=1.0, rope_high_freq_factor=4.0, rope_original_max_pos=64)


def hf_model(cfg: ModelCfg, weights: dict, dtype):
    from transformers import LlamaConfig, LlamaForCausalLM

    hc = LlamaConfig(
        vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.intermediate,
        num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.q_heads, num_key_value_heads=cfg.kv_heads,
        head_dim=cfg.head_dim, max_position_embeddings=cfg.max_model_len, rms_norm_eps=cfg.rms_eps,
        tie_word_embeddings=False, attention_bias=False, mlp_bias=False,
        hidden_act="silu",
        rope_parameters=(dict(rope_type="llama3", rope_theta=cfg.rope_theta, factor=cfg.rope_factor,
                              low_freq_factor=cfg.rope_low_freq_factor, high_freq_factor=cfg.rope_high_freq_factor,
                              original_max_position_embeddings=cfg.rope_original_max_pos)
                         if cfg.rope_scaling_type == 2 else dict(rope_type="default", rope_theta=cfg.rope_theta)),
    )
    hc._attn_implementation = "eager"
    m = LlamaForCausalLM(hc).to(dtype)
    t = lambda name: torch.from_numpy(bf16_bits_to_f32(weights[name]).copy()).to(dtype)
    D, Hq, Hkv, I = cfg.head_dim, cfg.q_heads, cfg.kv_heads, cfg.intermediate
    sd = {"model.embed_tokens.weight": t("embed"), "model.norm.weight": t("final_norm"),
          "lm_head.weight": t("lm_head")}
    for l in range(cfg.num_layers):
        p, q = f"layers.{l}.", f"model.layers.{l}."
        wqkv, wgu = t(p + "wqkv"), t(p + "wgu")
        sd[q + "self_attn.q_proj.weight"] = wqkv[: Hq * D]
        sd[q + "self_attn.k_proj.weight"] = wqkv[Hq * D: (Hq + Hkv) * D]
        sd[q + "self_attn.v_proj.weight"] = wqkv[(Hq + Hkv) * D:]
        sd[q + "self_attn.o_proj.weight"] = t(p + "wo")
        sd[q + "mlp.gate_proj.weight"] = wgu[:I]
        sd[q + "mlp.up_proj.weight"] = wgu[I:]
        sd[q + "mlp.down_proj.weight"] = t(p + "wdown")
        sd[q + "input_layernorm.weight"] = t(p + "norm1")
        sd[q + "post_attention_layernorm.weight"] = t(p + "norm2")
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "rotary" not in k], missing
    assert not unexpected, unexpected
    return m.eval()


def main():
    write(ModelCfg(), OUT)
    write(ModelCfg(**ROPE3), OUT_ROPE3)


def write(cfg, OUT):
    w = make_weights(cfg)
    rng = np.random.default_rng(7)
    ids = rng.integers(0, cfg.vocab, size=40).astype(np.int64)
    with torch.no_grad():
        l32 = hf_model(cfg, w, torch.float32)(torch.from_numpy(ids)[None]).logits[0].float().numpy()
        l16 = hf_model(cfg, w, torch.bfloat16)(torch.from_numpy(ids)[None]).logits[0].float().numpy()
        # greedy continuation from the fp32 model (token-level golden)
        m = hf_model(cfg, w, torch.float32)
        gen = m.generate(torch.from_numpy(ids)[None], max_new_tokens=12, do_sample=False)[0, len(ids):].numpy()
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, ids=ids, logits_fp32=l32.astype(np.float32), logits_bf16=l16.astype(np.float32),
                        greedy=gen.astype(np.int64), cfg=np.array([cfg.num_layers, cfg.hidden, cfg.q_heads,
                                                                   cfg.kv_heads, cfg.intermediate, cfg.vocab,
                                                                   cfg.max_model_len, cfg.seed]))
    print("wrote", OUT, l32.shape, "greedy", gen.tolist())
    top2 = np.sort(l32, axis=-1)[:, -2:]
    print("logit std %.3f  mean top1-top2 margin %.3f  min margin %.4f" %
          (l32.std(), (top2[:, 1] - top2[:, 0]).mean(), (top2[:, 1] - top2[:, 0]).min()))


if __name__ == "__main__":
    sys.exit(main())
