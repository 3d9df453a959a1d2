"""BASELINE.json configs[0] (the reference's CPU-runnable plumbing case) on the host cores: the product's load generator
over HTTP -> oracle/cpu_server.py (restated proxy + router + CPU backend).  Tiny model here; bench.py's cpu_baseline leg
runs the same path at qwen2:0.5b's dimensions."""
import json
import urllib.error
import urllib.request

import pytest

from kubeai_b200.server import harness_config, harness_run
from oracle.cpu_server import CpuBackend, Server, chat_prompt, encode

TINY = dict(num_layers=2, hidden=64, q_heads=2, kv_heads=1, intermediate=128, vocab=151936, head_dim=32, rms_eps=1e-6,
            rope_theta=1000000.0, max_model_len=4096)


@pytest.fixture(scope="module")
def srv():
    s = Server(CpuBackend(TINY, threads=2))
    yield s
    s.close()


def test_python_tokenizer_restates_the_native_one():
    from kubeai_b200.server import tokenize
    for text in ("hello wxyz abcd", " aaaa zzzz\n", "héllo  wörld", " abcde", "x abc"):
        assert encode(text, 151936) == tokenize(text, 151936), text
    ids = chat_prompt([{"role": "user", "content": "hi"}], 1000)
    assert ids[0] == 998 and ids[-1] == 10 and 999 in ids


def test_harness_drives_the_cpu_path_with_the_ollama_config(srv):
    hcfg = harness_config(request_model="qwen2:0.5b", max_concurrent_threads=2, max_completion_tokens=10, temperature=0.0,
                          synth_threads=4, seed=2, synth_mean_words=6, request_timeout_s=60.0)
    r = harness_run(hcfg, host="127.0.0.1", port=srv.port)
    assert r["failed_threads"] == 0 and r["first_error"] == ""
    assert r["request_count"] >= 4 * 5                                  # >= 5 user messages per synthetic thread
    assert r["completion_tokens"] == 10 * r["request_count"]
    assert r["chunks_per_request_mean"] == 10.0                         # role chunk + 9 content chunks; the 10th rides on finish
    assert r["cached_prompt_tokens"] > 0.5 * r["prompt_tokens"]         # later turns reuse the conversation's KV
    assert r["run_output_throughput"] > 0 and r["ttft_p50_s"] > 0
    assert srv.proxy.active == 0 and srv.proxy.group.total_in_flight == 0


def test_error_bodies(srv):
    def post(body):
        req = urllib.request.Request(f"http://127.0.0.1:{srv.port}/openai/v1/chat/completions", data=body,
                                     headers={"Content-Type": "application/json"})
        try:
            urllib.request.urlopen(req, timeout=10)
        except urllib.error.HTTPError as e:
            return e.code, json.loads(e.read())["error"]["message"]
        return 200, ""
    assert post(b"{")[0] == 400
    assert post(json.dumps({"messages": [{"role": "user", "content": "x"}]}).encode())[0] == 400
    assert post(json.dumps({"model": "other", "messages": [{"role": "user", "content": "x"}]}).encode()) == (404, "model not found: other")
