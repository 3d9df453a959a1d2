"""Opt-in (B200_VLLM_PARITY=1): greedy parity of the engine against the vLLM wheel of this image — the closest runnable
stand-in for the reference's backend pod (the repo pins vllm/vllm-openai:v0.10.2, charts/kubeai/values.yaml:45; the image
has 0.22) — on the 2-layer test model, both loading the SAME HF checkpoint.  Takes a few minutes of vLLM start-up, so it
is not part of the default GPU suite; the recorded outcome lives in profiles/."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B200_VLLM_PARITY") != "1", reason="set B200_VLLM_PARITY=1 (slow: starts vLLM)")]


def test_greedy_tokens_match_vllm_on_the_same_checkpoint(tmp_path):
    from safetensors.torch import save_file
    from kubeai_b200.engine import Engine, config_from_hf
    from oracle.gen_golden import hf_model
    from oracle.weights import ModelCfg, make_weights
    cfg = ModelCfg()
    m = hf_model(cfg, make_weights(cfg), torch.bfloat16)
    m.config.save_pretrained(tmp_path)
    save_file({k: v.contiguous() for k, v in m.state_dict().items()}, str(tmp_path / "model.safetensors"))

    g = torch.Generator().manual_seed(17)
    prompts = [torch.randint(0, cfg.vocab, (int(n),), generator=g).tolist() for n in (5, 16, 33, 70, 121, 200)]
    N = 24

    ecfg = config_from_hf(tmp_path, max_model_len=256, max_num_seqs=16, max_batched_tokens=256, num_kv_blocks=128,
                          manual_step=1, seed=777)
    with Engine(ecfg) as e:
        e.load_safetensors(tmp_path)
        mine = e.generate(prompts, max_tokens=N)

    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    os.environ.setdefault("VLLM_NO_USAGE_STATS", "1")
    from vllm import LLM, SamplingParams
    llm = LLM(model=str(tmp_path), skip_tokenizer_init=True, dtype="bfloat16", max_model_len=256, enforce_eager=True,
              gpu_memory_utilization=0.3, enable_prefix_caching=False, seed=0)
    sp = SamplingParams(temperature=0.0, max_tokens=N, ignore_eos=True, detokenize=False, logprobs=2)
    outs = llm.generate([{"prompt_token_ids": p} for p in prompts], sp)

    report = []
    for i, (o, got) in enumerate(zip(outs, mine)):
        ref = list(o.outputs[0].token_ids)
        k = next((j for j in range(N) if ref[j] != got[j]), N)
        margin = None
        if k < N:
            lp = sorted((v.logprob for v in o.outputs[0].logprobs[k].values()), reverse=True)
            margin = lp[0] - lp[1] if len(lp) > 1 else float("inf")
        report.append((len(prompts[i]), k, margin))
    print("prompt_len, identical_prefix_of_%d, vllm top-2 logprob margin at the first difference:" % N, report)
    # two bf16 implementations may part ways only where vLLM's own top-2 candidates are a rounding error apart
    for plen, k, margin in report:
        assert k == N or margin < 0.25, (plen, k, margin)
    assert sum(k == N for _, k, _ in report) >= len(report) // 2


@pytest.mark.skipif(os.environ.get("B200_VLLM_PARITY_8B") != "1", reason="set B200_VLLM_PARITY_8B=1 (writes a 16 GB checkpoint)")
def test_full_size_llama3_8b_greedy_tokens_match_vllm(tmp_path):
    """BASELINE config 2 model shape: the engine's seeded Llama-3-8B weights are exported as an HF checkpoint, vLLM loads
    it, and both decode the same prompts greedily."""
    import json
    from safetensors.torch import save_file
    from transformers import LlamaConfig
    from kubeai_b200.engine import Engine, default_config
    L, H, I, V, Hq, Hkv, D = 32, 4096, 14336, 128256, 32, 8, 128
    g = torch.Generator().manual_seed(23)
    prompts = [torch.randint(0, 128000, (int(n),), generator=g).tolist() for n in (5, 17, 64, 150, 333, 700)]
    N = 16

    def bf16(bits: np.ndarray, rows: int, cols: int) -> torch.Tensor:
        return torch.from_numpy(bits.view(np.int16).reshape(rows, cols)).view(torch.bfloat16)

    with Engine(default_config(manual_step=1, max_num_seqs=16, max_batched_tokens=2048, max_model_len=2048, kv_fraction=0.05)) as e:
        mine = e.generate(prompts, max_tokens=N)
        weight_map = {}

        def dump(fn, tensors):
            save_file({k: v.contiguous() for k, v in tensors.items()}, str(tmp_path / fn))
            weight_map.update({k: fn for k in tensors})

        dump("model-head.safetensors", {
            "model.embed_tokens.weight": bf16(e.tensor("embed"), V, H), "lm_head.weight": bf16(e.tensor("lm_head"), V, H),
            "model.norm.weight": bf16(e.tensor("final_norm"), 1, H).reshape(H)})
        for l in range(L):
            p, q = f"layers.{l}.", f"model.layers.{l}."
            wqkv = bf16(e.tensor(p + "wqkv"), (Hq + 2 * Hkv) * D, H)
            wgu = bf16(e.tensor(p + "wgu"), 2 * I, H)
            dump(f"model-layer-{l:02d}.safetensors", {
                q + "self_attn.q_proj.weight": wqkv[:Hq * D], q + "self_attn.k_proj.weight": wqkv[Hq * D:(Hq + Hkv) * D],
                q + "self_attn.v_proj.weight": wqkv[(Hq + Hkv) * D:], q + "self_attn.o_proj.weight": bf16(e.tensor(p + "wo"), H, Hq * D),
                q + "mlp.gate_proj.weight": wgu[:I], q + "mlp.up_proj.weight": wgu[I:],
                q + "mlp.down_proj.weight": bf16(e.tensor(p + "wdown"), H, I),
                q + "input_layernorm.weight": bf16(e.tensor(p + "norm1"), 1, H).reshape(H),
                q + "post_attention_layernorm.weight": bf16(e.tensor(p + "norm2"), 1, H).reshape(H)})
    (tmp_path / "model.safetensors.index.json").write_text(json.dumps({"metadata": {}, "weight_map": weight_map}))
    LlamaConfig(vocab_size=V, hidden_size=H, intermediate_size=I, num_hidden_layers=L, num_attention_heads=Hq,
                num_key_value_heads=Hkv, max_position_embeddings=2048, rms_norm_eps=1e-5, rope_theta=500000.0,
                tie_word_embeddings=False, torch_dtype="bfloat16", bos_token_id=1, eos_token_id=2).save_pretrained(tmp_path)

    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    os.environ.setdefault("VLLM_NO_USAGE_STATS", "1")
    from vllm import LLM, SamplingParams
    llm = LLM(model=str(tmp_path), skip_tokenizer_init=True, dtype="bfloat16", max_model_len=2048, enforce_eager=True,
              gpu_memory_utilization=0.45, enable_prefix_caching=False, seed=0)
    sp = SamplingParams(temperature=0.0, max_tokens=N, ignore_eos=True, detokenize=False, logprobs=2)
    outs = llm.generate([{"prompt_token_ids": p} for p in prompts], sp)
    report = []
    for i, (o, got) in enumerate(zip(outs, mine)):
        ref = list(o.outputs[0].token_ids)
        k = next((j for j in range(N) if ref[j] != got[j]), N)
        margin = None
        if k < N:
            lp = sorted((v.logprob for v in o.outputs[0].logprobs[k].values()), reverse=True)
            margin = lp[0] - lp[1] if len(lp) > 1 else float("inf")
        report.append((len(prompts[i]), k, margin))
    print("8B: prompt_len, identical_prefix_of_%d, vllm top-2 logprob margin at the first difference:" % N, report)
    # random weights over a 128k vocabulary leave many top-2 candidates within one bf16 step of the logit (0.125 at
    # |logit| 16-32): streams may part there and only there
    for plen, k, margin in report:
        assert k == N or margin < 0.25, (plen, k, margin)
    assert any(k == N for _, k, _ in report)
