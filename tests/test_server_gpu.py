"""Serving shell + load generator on the GPU (mini model): wire framing against the reference's vLLM
capture (api/openai/v1/reference/example-requests.vllm.output:139-160), proxy retry semantics
(internal/modelproxy/handler_test.go:137-168), prefix routing, multi-turn prefix-cache hits."""
import json
import re
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
V = 512


@pytest.fixture()
def stack():
    from kubeai_b200.engine import Engine, mini_config
    from kubeai_b200.server import PREFIX_HASH, Server
    engines = [Engine(mini_config(manual_step=0, max_model_len=1024, max_batched_tokens=512, num_kv_blocks=512))
               for _ in range(2)]
    srv = Server(engines, model="mini", adapters=["lora1"], strategy=PREFIX_HASH, prefix_char_length=10, vocab=V,
                 max_model_len=1024)
    yield srv, engines
    srv.close()
    for e in engines:
        e.close()


def chat(srv, messages, **kw):
    body = dict(model="mini", messages=messages, max_tokens=6, temperature=0, **kw)
    return srv.handle("POST", "/openai/v1/chat/completions", json.dumps(body))


def test_streaming_chat_matches_vllm_framing(stack):
    srv, _ = stack
    r = chat(srv, [{"role": "user", "content": "Hello!"}], stream=True, stream_options={"include_usage": True})
    assert r.status == 200 and r.content_type.startswith("text/event-stream")
    ev = r.sse_events()
    assert ev[-1] == "[DONE]"
    chunks = [json.loads(e) for e in ev[:-1]]
    cid = chunks[0]["id"]
    assert re.fullmatch(r"chatcmpl-[0-9a-f]{32}", cid) and all(c["id"] == cid for c in chunks)
    assert all(c["object"] == "chat.completion.chunk" and c["model"] == "mini" for c in chunks)
    # first chunk: role + empty content; then one chunk per token; the last token chunk carries finish_reason
    assert chunks[0]["choices"] == [{"index": 0, "delta": {"role": "assistant", "content": ""}, "logprobs": None,
                                     "finish_reason": None}]
    toks = chunks[1:-1]
    assert len(toks) == 6 and all(set(c["choices"][0]) >= {"index", "delta", "logprobs", "finish_reason"} for c in toks)
    assert [c["choices"][0]["finish_reason"] for c in toks] == [None] * 5 + ["length"]
    usage = chunks[-1]
    assert usage["choices"] == [] and usage["usage"]["completion_tokens"] == 6
    assert usage["usage"]["total_tokens"] == usage["usage"]["prompt_tokens"] + 6
    # the harness counting rule (runner.go:319-321): role chunk + 5 unfinished token chunks
    counted = [c for c in chunks if c["choices"] and c["choices"][0]["finish_reason"] in (None, "")]
    assert len(counted) == 6


def test_non_stream_and_completions_with_token_ids(stack):
    from kubeai_b200.server import detokenize, tokenize
    srv, engines = stack
    r = chat(srv, [{"role": "system", "content": "Be brief."}, {"role": "user", "content": "Hi"}])
    j = r.json()
    assert r.status == 200 and j["object"] == "chat.completion" and j["choices"][0]["finish_reason"] == "length"
    text = j["choices"][0]["message"]["content"]
    ids = tokenize(text, V)
    assert len(ids) == 6 and detokenize(ids, V) == text
    # same prompt through /v1/completions with explicit ids == what the chat template produced
    prompt = [V - 2] + tokenize("system", V) + [10] + tokenize("Be brief.", V) + [V - 1, 10] + \
             [V - 2] + tokenize("user", V) + [10] + tokenize("Hi", V) + [V - 1, 10] + [V - 2] + tokenize("assistant", V) + [10]
    r2 = srv.handle("POST", "/openai/v1/completions", json.dumps(dict(model="mini", prompt=prompt, max_tokens=6, temperature=0)))
    assert r2.status == 200 and r2.json()["choices"][0]["text"] == text
    assert r2.json()["usage"]["prompt_tokens"] == len(prompt) == j["usage"]["prompt_tokens"]


def test_retry_on_replica_failure_then_bad_gateway(stack):
    srv, _ = stack
    srv.inject_fault(0, 2)
    srv.inject_fault(1, 1)
    r = chat(srv, [{"role": "user", "content": "retry me"}])
    assert r.status == 200                       # at most 3 failures < 1 + maxRetries attempts
    retries = int(re.search(r"b200_request_retries_total (\d+)", srv.metrics()).group(1))
    assert 1 <= retries <= 3                     # each retry re-picks on the ring and may land on the same replica
    srv.inject_fault(0, 100)
    srv.inject_fault(1, 100)
    r = chat(srv, [{"role": "user", "content": "retry me"}])
    assert (r.status, r.body) == (502, b'{"error":"Bad Gateway"}\n')   # handler_test.go "dropped connection"
    assert int(re.search(r"b200_request_retries_total (\d+)", srv.metrics()).group(1)) == retries + 3   # 1 + maxRetries attempts
    srv.inject_fault(0, 0)
    srv.inject_fault(1, 0)
    assert 'kubeai_inference_requests_active{request_model="mini",request_type="http"} 0' in srv.metrics()


def test_prefix_hash_keeps_a_conversation_on_one_replica_and_hits_the_cache(stack):
    srv, engines = stack
    msgs = [{"role": "user", "content": "conversation number one " * 3}]
    for turn in range(3):
        j = chat(srv, msgs).json()
        msgs.append({"role": "assistant", "content": j["choices"][0]["message"]["content"]})
        msgs.append({"role": "user", "content": f"and then {turn}"})
        if turn:
            assert j["usage"]["prompt_tokens_details"]["cached_tokens"] >= 16 * ((j["usage"]["prompt_tokens"] - 40) // 16)
    s0, s1 = engines[0].stats(), engines[1].stats()
    assert (s0.generated_tokens == 0) != (s1.generated_tokens == 0), "all turns must land on one replica"
    # adapter requests route only to replicas that have it (all do here) and use adapter+prefix as key
    r = srv.handle("POST", "/openai/v1/chat/completions", json.dumps(dict(model="mini_lora1", messages=msgs[:1], max_tokens=2, temperature=0)))
    assert r.status == 200 and r.json()["model"] == "mini_lora1"


def test_harness_in_process_and_over_http_agree(stack):
    from kubeai_b200.server import harness_config, harness_run
    srv, engines = stack
    cfg = harness_config(request_model="mini", max_concurrent_threads=4, max_completion_tokens=5, synth_threads=6,
                         synth_mean_words=6, vocab=V, seed=3)
    a = harness_run(cfg, server=srv)
    port = srv.listen()
    b = harness_run(cfg, host="127.0.0.1", port=port)
    for r in (a, b):
        assert r["failed_threads"] == 0, r["first_error"]
        assert r["request_count"] >= 30 and r["completion_tokens"] == 5 * r["request_count"]
        assert r["chunks_per_request_mean"] == 5.0                  # role chunk + 4 unfinished token chunks
        assert r["cached_prompt_tokens"] > 0 and r["ttft_p99_s"] >= r["ttft_p50_s"] > 0
    assert a["prompt_tokens"] == b["prompt_tokens"] and a["request_count"] == b["request_count"]


def test_client_disconnect_aborts_the_sequence(stack):
    srv, engines = stack
    seen = []

    def on_chunk(c):
        seen.append(c)
        return len(seen) >= 3          # "client gone" after three writes

    body = dict(model="mini", messages=[{"role": "user", "content": "x"}], max_tokens=200, temperature=0, stream=True)
    srv.handle("POST", "/openai/v1/chat/completions", json.dumps(body), on_chunk=on_chunk)
    import time
    for _ in range(100):
        if all(e.stats().running == 0 for e in engines):
            break
        time.sleep(0.02)
    st = [e.stats() for e in engines]
    assert all(s.running == 0 and s.kv_blocks_free == s.kv_blocks_total for s in st)
    assert sum(s.generated_tokens for s in st) < 200


def test_failed_replica_leaves_the_endpoint_set(stack):
    """A replica whose engine is in the failed state is dropped from the router like an endpoint that vanished
    (internal/loadbalancer/group.go:119-131): one request pays the retry, later ones never see the dead replica."""
    from kubeai_b200.server import LEAST_LOAD, Server
    _, engines = stack
    with Server(engines, model="mini", strategy=LEAST_LOAD, vocab=V, max_model_len=1024) as srv:
        assert not any(srv._l.b200_engine_is_failed(e._h) for e in engines)
        srv.inject_fault(0, -1)                                    # replica 0 fails every submit from now on
        retried = lambda: int(re.search(r"b200_request_retries_total (\d+)", srv.metrics()).group(1))
        # LeastLoad picks an idle replica; ties go to the first endpoint, so the dead one is tried first
        for i in range(6):
            assert chat(srv, [{"role": "user", "content": f"after the failure {i}"}]).status == 200
        assert retried() == 1, "only the request that found the replica dead retried"
        m = srv.metrics()
        assert 'endpoint="gpu-0"' not in m.split("b200_requests_total")[0] or "gpu-1" in m
        st0, st1 = engines[0].stats(), engines[1].stats()
        assert st0.generated_tokens == 0 and st1.generated_tokens == 6 * 6


def test_destroy_with_idle_keepalive_client_returns_promptly(stack):
    """The destructor shuts client sockets down and waits for every connection thread: no 5 s stall, nothing touches
    the freed server afterwards."""
    import socket
    import time
    from kubeai_b200.server import LEAST_LOAD, Server
    _, engines = stack
    srv = Server(engines, model="mini", strategy=LEAST_LOAD, vocab=V, max_model_len=1024)
    port = srv.listen()
    socks = []
    for _ in range(3):
        c = socket.create_connection(("127.0.0.1", port))
        c.sendall(b"GET /healthz HTTP/1.1\r\nHost: x\r\n\r\n")
        got = b""
        while not got.endswith(b"0\r\n\r\n"):                     # the whole chunked response
            got += c.recv(4096)
        assert b"200 OK" in got
        socks.append(c)                                           # kept open: the server thread sits in recv()
    t0 = time.perf_counter()
    srv.close()
    assert time.perf_counter() - t0 < 3.0                          # the old behaviour was a 5 s stall
    for c in socks:
        assert c.recv(16) == b""                                  # peer closed
        c.close()


def test_real_tokenizer_end_to_end(stack, tmp_path):
    """A tokenizer.json of the Llama-3 pipeline attached to the server: the prompt the engine sees is the chat template's,
    the streamed deltas are valid text whose concatenation is the detokenised generation, non-stream text is the same."""
    tokenizers = pytest.importorskip("tokenizers")
    from tokenizers import Regex, Tokenizer as HFTokenizer, decoders, models, pre_tokenizers, trainers
    from kubeai_b200.tokenizer import Tokenizer
    pat = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
    hf = HFTokenizer(models.BPE(ignore_merges=True))
    hf.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(pat), behavior="isolated"), pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    hf.decoder = decoders.ByteLevel()
    specials = ["<|begin_of_text|>", "<|end_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>"]
    hf.train_from_iterator(["hello world, how are you today? I'm fine thanks; it's 12345 o'clock"] * 50,
                           trainers.BpeTrainer(vocab_size=V - 8, special_tokens=specials, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False))
    hf.save(str(tmp_path / "tokenizer.json"))
    srv, engines = stack
    with Tokenizer(tmp_path / "tokenizer.json") as tok:
        assert tok.vocab_size <= V
        srv.set_tokenizer(tok)
        msgs = [{"role": "user", "content": "hello world, how are you?"}]
        body = dict(model="mini", messages=msgs, max_tokens=24, temperature=0)
        prompt = srv.render_prompt("/v1/chat/completions", json.dumps(body))
        assert prompt == tok.chat(msgs) and prompt[0] == tok.token_id("<|begin_of_text|>")
        # what the engine generates for that prompt, with the stop ids the server adds
        stops = [tok.token_id("<|eot_id|>"), tok.token_id("<|end_of_text|>")]
        rid = engines[0].submit(prompt, max_tokens=24, ignore_eos=False, stop_ids=stops)
        gen = []
        while True:
            assert engines[0].wait(rid, 30.0)
            pr = engines[0].poll(rid)
            gen += pr.tokens
            if pr.finished:
                break
        engines[0].release(rid)
        want = tok.decode([t for t in gen if t not in stops], skip_special=True)
        r = srv.handle("POST", "/openai/v1/chat/completions", json.dumps(dict(body, stream=True)))
        chunks = [json.loads(e) for e in r.sse_events()[:-1]]
        deltas = [c["choices"][0]["delta"].get("content", "") for c in chunks if c["choices"]]
        assert len(deltas) == 1 + len(gen)                  # role chunk + one chunk per token, whatever text each completes
        for d in deltas:
            d.encode("utf-8")                               # every delta is well-formed text on its own
        assert "".join(deltas) == want
        r = srv.handle("POST", "/openai/v1/chat/completions", json.dumps(body))
        assert r.json()["choices"][0]["message"]["content"] == want
        assert r.json()["usage"]["prompt_tokens"] == len(prompt)
        srv.set_tokenizer(None)
