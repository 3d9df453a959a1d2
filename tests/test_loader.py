"""HF checkpoint loading.  CPU part: safetensors / config.json parsing through the C ABI against
files written by the `safetensors` package.  GPU part: an engine loaded from a checkpoint written by HF transformers
reproduces the golden logits."""
import json
import struct
from pathlib import Path

import numpy as np
import pytest
import torch

from kubeai_b200 import B200Error, lib
from kubeai_b200.engine import config_from_hf, safetensors_list

GOLD = np.load(Path(__file__).parent / "golden" / "llama_mini.npz")


def _write_checkpoint(d: Path, sharded: bool, dtype=torch.bfloat16):
    from safetensors.torch import save_file
    from oracle.gen_golden import hf_model
    from oracle.weights import ModelCfg, make_weights
    cfg = ModelCfg()
    m = hf_model(cfg, make_weights(cfg), dtype)
    sd = {k: v.contiguous() for k, v in m.state_dict().items()}
    m.config.save_pretrained(d)
    if not sharded:
        save_file(sd, str(d / "model.safetensors"))
    else:
        names = sorted(sd)
        half = len(names) // 2
        parts = {"model-00001-of-00002.safetensors": names[:half], "model-00002-of-00002.safetensors": names[half:]}
        for fn, ns in parts.items():
            save_file({n: sd[n] for n in ns}, str(d / fn))
        (d / "model.safetensors.index.json").write_text(json.dumps(
            {"metadata": {}, "weight_map": {n: fn for fn, ns in parts.items() for n in ns}}))
    return cfg, sd


def test_safetensors_listing_and_config_parsing(tmp_path):
    cfg, sd = _write_checkpoint(tmp_path, sharded=True)
    listing = safetensors_list(tmp_path)
    assert set(listing) == set(sd)
    e = listing["model.layers.1.mlp.gate_proj.weight"]
    assert e == {"dtype": "BF16", "shape": [cfg.intermediate, cfg.hidden], "bytes": cfg.intermediate * cfg.hidden * 2}
    c = config_from_hf(tmp_path)
    assert (c.num_layers, c.hidden, c.q_heads, c.kv_heads, c.intermediate, c.vocab) == \
           (cfg.num_layers, cfg.hidden, cfg.q_heads, cfg.kv_heads, cfg.intermediate, cfg.vocab)
    assert abs(c.rms_eps - cfg.rms_eps) < 1e-9 and c.rope_theta == cfg.rope_theta
    # unsupported architectures are refused loudly, not approximated
    bad = json.loads((tmp_path / "config.json").read_text())
    bad["head_dim"] = 64
    (tmp_path / "config.json").write_text(json.dumps(bad))
    with pytest.raises(B200Error, match="head_dim 64"):
        config_from_hf(tmp_path)


def test_corrupt_safetensors_is_rejected(tmp_path):
    p = tmp_path / "model.safetensors"
    p.write_bytes(struct.pack("<Q", 1 << 40) + b"{}")
    with pytest.raises(B200Error, match="header length"):
        safetensors_list(p)
    hdr = json.dumps({"w": {"dtype": "BF16", "shape": [4], "data_offsets": [0, 64]}}).encode()
    p.write_bytes(struct.pack("<Q", len(hdr)) + hdr + b"\0" * 8)
    with pytest.raises(B200Error, match="data_offsets out of range"):
        safetensors_list(p)
    # untrusted header numbers: negative, fractional, wrapping (2^64 - 2048) and overflowing shapes are all refused
    for entry in ({"dtype": "BF16", "shape": [4], "data_offsets": [-8, 0]},
                  {"dtype": "BF16", "shape": [4], "data_offsets": [0, 7.5]},
                  {"dtype": "BF16", "shape": [4], "data_offsets": [18446744073709549568, 18446744073709549576]},
                  {"dtype": "BF16", "shape": [-4], "data_offsets": [0, 8]},
                  {"dtype": "BF16", "shape": [4294967296, 4294967296, 16], "data_offsets": [0, 8]},
                  {"dtype": "BF16", "shape": ["4"], "data_offsets": [0, 8]}):
        hdr = json.dumps({"w": entry}).encode()
        p.write_bytes(struct.pack("<Q", len(hdr)) + hdr + b"\0" * 8)
        with pytest.raises(B200Error, match="out of range"):
            safetensors_list(p)


def test_config_json_rope_scaling_and_unsupported_switches(tmp_path):
    base = dict(model_type="llama", hidden_size=512, num_attention_heads=4, num_key_value_heads=1, intermediate_size=1024,
                num_hidden_layers=2, vocab_size=512, head_dim=128, rms_norm_eps=1e-5, rope_theta=500000.0, hidden_act="silu")

    def cfg_of(**kw):
        d = dict(base)
        d.update(kw)
        (tmp_path / "config.json").write_text(json.dumps(d))
        return config_from_hf(tmp_path)
    c = cfg_of()
    assert c.rope_scaling_type == 0
    # Llama-3.1 style (transformers < 5 key "rope_scaling", >= 5 key "rope_parameters")
    l3 = dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=8192)
    for key in ("rope_scaling", "rope_parameters"):
        c = cfg_of(**{key: dict(l3, rope_theta=500000.0)})
        assert (c.rope_scaling_type, c.rope_factor, c.rope_low_freq_factor, c.rope_high_freq_factor, c.rope_original_max_pos) == \
               (2, 8.0, 1.0, 4.0, 8192)
    c = cfg_of(rope_scaling=dict(type="linear", factor=2.0))
    assert (c.rope_scaling_type, c.rope_factor) == (1, 2.0)
    for kw, msg in ((dict(rope_scaling=dict(rope_type="yarn", factor=4.0)), "yarn"), (dict(model_type="qwen2"), "model_type"),
                    (dict(hidden_act="gelu"), "hidden_act"), (dict(attention_bias=True), "attention_bias"),
                    (dict(mlp_bias=True), "mlp_bias"), (dict(hidden_size=576, num_attention_heads=4, head_dim=128), "hidden"),
                    (dict(vocab_size=509), "vocab")):
        with pytest.raises(B200Error, match=msg):
            cfg_of(**kw)


@pytest.mark.gpu
@pytest.mark.parametrize("sharded,dtype", [(False, torch.bfloat16), (True, torch.float32)])
def test_engine_loaded_from_hf_checkpoint_matches_golden(tmp_path, sharded, dtype):
    from kubeai_b200.engine import Engine
    _write_checkpoint(tmp_path, sharded, dtype)
    cfg = config_from_hf(tmp_path, max_model_len=256, max_num_seqs=16, max_batched_tokens=256, num_kv_blocks=128,
                         manual_step=1, seed=12345)   # different seed: every weight must come from the checkpoint
    with Engine(cfg) as e:
        e.load_safetensors(tmp_path)
        got = e.forward_logits(GOLD["ids"])
        ref = GOLD["logits_fp32"]
        err = np.abs(got - ref)
        assert (err <= 0.15 + 1.6e-2 * np.abs(ref)).all(), err.max()
        # every tensor is bit-identical to the weights the checkpoint was written from ...
        from oracle.weights import ModelCfg, make_weights, tensor_specs
        w = make_weights(ModelCfg())
        for name, _, _, _ in tensor_specs(ModelCfg()):
            assert np.array_equal(e.tensor(name), w[name].reshape(-1)), name
        loaded = e.generate([GOLD["ids"].tolist()], max_tokens=12)[0]
    # ... so generation is identical to an engine that initialised the same weights from the seed
    from kubeai_b200.engine import mini_config
    with Engine(mini_config()) as e0:
        assert loaded == e0.generate([GOLD["ids"].tolist()], max_tokens=12)[0]
    assert loaded[:4] == GOLD["greedy"].tolist()[:4]


@pytest.mark.gpu
def test_b200serve_and_b200bench_processes_end_to_end(tmp_path):
    """The two process entry points (cmd/): an HTTP server over a checkpoint-loaded engine, driven by the load generator
    binary — the standalone equivalent of `kubeai` + `benchmarks/multi-turn-chat-go`."""
    import re
    import subprocess
    import time
    import urllib.request
    from kubeai_b200._lib import LIB_PATH
    bindir = LIB_PATH.parent.parent / "bin"
    _write_checkpoint(tmp_path, sharded=False)
    srv = subprocess.Popen([str(bindir / "b200serve"), "--gpus", "1", "--model", "mini", "--model-dir", str(tmp_path), "--port", "0",
                            "--host", "127.0.0.1", "--max-model-len", "1024", "--max-num-batched-tokens", "512", "--max-num-seqs", "16",
                            "--gpu-memory-utilization", "0.05", "--strategy", "PrefixHash"], stderr=subprocess.PIPE, text=True)
    try:
        port = None
        t0 = time.time()
        while time.time() - t0 < 120:
            line = srv.stderr.readline()
            m = re.search(r"on http://127.0.0.1:(\d+)/", line or "")
            if m:
                port = int(m.group(1))
                break
            assert srv.poll() is None, line
        assert port, "b200serve did not come up"
        models = json.loads(urllib.request.urlopen(f"http://127.0.0.1:{port}/openai/v1/models", timeout=10).read())
        assert models["data"][0]["id"] == "mini"
        out = subprocess.run([str(bindir / "b200bench"), "--base-url", f"http://127.0.0.1:{port}/openai", "--request-model", "mini",
                              "--synthetic-threads", "5", "--synthetic-words", "6", "--vocab", "512", "--max-concurrent-threads", "3",
                              "--max-completion-tokens", "5", "--seed", "2"], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "Failed thread count: 0" in out.stdout and re.search(r"Chunks per request \(mean\): 5\.00", out.stdout)
        metrics = urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=10).read().decode()
        assert "kubeai_inference_requests_hash_lookup_final" in metrics and "b200_engine_generated_tokens_total" in metrics
    finally:
        srv.terminate()
        srv.wait(20)
