"""Host-side serving shell, CPU-only parts: request parsing / error matrix of the reference
(internal/modelproxy/handler_test.go:60-75, internal/apiutils/request_test.go:13-84,
model_test.go:10-78, api/openai/v1/*_test.go Prefix tables), tokenizer round trip, harness
arithmetic against the reference's mock SSE server (benchmark/runner_test.go:20-116)."""
import json
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

import pytest

from kubeai_b200.server import PREFIX_HASH, Server, detokenize, harness_config, harness_run, tokenize


@pytest.fixture()
def srv():
    # no engine behind it: only paths that end before the engine call are exercised here
    with Server([None], model="model1", adapters=["adapter3"]) as s:
        yield s


def test_error_matrix_matches_reference_strings(srv):
    r = srv.handle("POST", "/openai/v1/chat/completions", "{}")
    assert (r.status, r.body) == (400, b'{"error":"bad request: reading model from body: missing \'model\' field"}\n')
    r = srv.handle("POST", "/openai/v1/chat/completions", '{"model":"does-not-exist"}')
    assert (r.status, r.body) == (404, b'{"error":"model not found: \\"does-not-exist\\""}\n')
    r = srv.handle("POST", "/openai/v1/chat/completions", '{"model":"model1_no-such-adapter","messages":[]}')
    assert (r.status, r.body) == (404, b'{"error":"model not found: \\"model1_no-such-adapter\\""}\n')
    r = srv.handle("POST", "/openai/v1/chat/completions", '{"model": ')
    assert r.status == 400 and r.json()["error"].startswith("bad request: reading model from body: decoding: ")
    r = srv.handle("POST", "/openai/v1/chat/completions", '{"model":"model1","temperature":0.7,"messages":[]}')
    assert r.status == 400 and "greedy" in r.json()["error"]
    assert srv.handle("GET", "/nope").status == 404
    assert srv.handle("GET", "/healthz").body == b"ok\n"
    models = srv.handle("GET", "/openai/v1/models").json()
    assert [m["id"] for m in models["data"]] == ["model1", "model1_adapter3"]
    assert "kubeai_inference_requests_active" in srv.metrics()


def test_tokenizer_round_trips_every_id_and_bytes():
    ids = [0, 65, 255, 256, 4095, 100000, 128255, 128254]
    text = detokenize(ids)
    assert tokenize(text) == ids and len(text) == 5 * len(ids)
    assert tokenize("Hi!") == [72, 105, 33]
    assert tokenize("héllo") == list("héllo".encode())
    # 4-letter lowercase words after a space are single tokens; anything else falls back to bytes
    assert len(tokenize(" abcd efgh")) == 2 and len(tokenize(" abcde")) == 6 and len(tokenize(" zzzz")) == 5
    assert tokenize(" abcd", vocab=512) == list(b" abcd")   # value >= vocab -> bytes


class _MockSSE(BaseHTTPRequestHandler):
    """benchmark/runner_test.go:107-204: TTFT 1 s (scaled), 3 content chunks x 10 tokens, usage chunk."""
    T0, GAP = 0.20, 0.04

    def log_message(self, *a):
        pass

    def do_POST(self):
        req = json.loads(self.rfile.read(int(self.headers["Content-Length"])))
        assert req["stream"] is True and req["stream_options"] == {"include_usage": True}
        assert req["temperature"] == 0.5 and req["max_tokens"] == 7
        self.server.bodies.append(req)
        self.send_response(200)
        self.send_header("Content-Type", "text/event-stream")
        self.end_headers()
        time.sleep(self.T0)
        for i in range(4):
            if i not in (0, 3):
                time.sleep(self.GAP)
            if i == 3:
                ev = {"choices": [], "usage": {"completion_tokens": 30, "prompt_tokens": 5, "total_tokens": 35,
                                               "prompt_tokens_details": {"cached_tokens": 2}}}
            else:
                ev = {"choices": [{"index": 0, "delta": {"role": "assistant", "content": "test chunk text"},
                                   "finish_reason": None}]}
            self.wfile.write(b"data: " + json.dumps(ev).encode() + b"\n\n")
            self.wfile.flush()
        self.wfile.write(b"data: [DONE]\n\n")


def test_harness_arithmetic_against_reference_mock_server():
    httpd = ThreadingHTTPServer(("127.0.0.1", 0), _MockSSE)
    httpd.bodies = []
    th = threading.Thread(target=httpd.serve_forever, daemon=True)
    th.start()
    threads = [
        {"id": "a", "messages": [{"role": "system", "content": "You are helpful."},
                                  {"role": "user", "content": "Hello"}, {"role": "user", "content": "Are you sure?"}]},
        {"id": "b", "messages": [{"role": "user", "content": "Hi"}, {"role": "user", "content": "Are you sure?"}]},
    ]
    cfg = harness_config(request_model="m", max_concurrent_threads=2, max_completion_tokens=7, temperature=0.5)
    r = harness_run(cfg, host="127.0.0.1", port=httpd.server_address[1], threads=threads)
    httpd.shutdown()
    # runner.go iterates over EVERY input message (thread a: 3 requests incl. the system one), 2 + 3 = 5
    assert r["failed_threads"] == 0 and r["request_count"] == 5 and r["input_thread_count"] == 2
    assert r["input_messages_per_thread_mean"] == 2.5 and r["chunks_per_request_mean"] == 3.0
    assert (r["prompt_tokens"], r["cached_prompt_tokens"], r["completion_tokens"], r["total_tokens"]) == (25, 10, 150, 175)
    assert abs(r["ttft_mean_s"] - _MockSSE.T0) < 0.03 and abs(r["ttft_p50_s"] - _MockSSE.T0) < 0.03
    # ITL = sum(later chunk gaps) / sum(completion_tokens * itl_chunks/(itl_chunks+1)) = 2*GAP / (30 * 2/3)
    assert abs(r["itl_mean_s"] - 2 * _MockSSE.GAP / 20.0) < 0.002
    assert r["run_output_throughput"] == pytest.approx(150 / r["duration_s"])
    # history grows by the assistant reply (3 chunks of text) between requests of a thread
    b = [x for x in httpd.bodies if x["messages"][0]["content"] == "Hi"]
    longest = max(b, key=lambda x: len(x["messages"]))
    assert [m["role"] for m in longest["messages"]] == ["user", "assistant", "user"]
    assert longest["messages"][1]["content"] == "test chunk text" * 3


@pytest.fixture()
def psrv():
    with Server([None], model="test-model", adapters=["test-adapter", "my-adapter", "my-adapter_extra"], strategy=PREFIX_HASH) as s:
        yield s


def test_chat_prefix_table(psrv):
    """api/openai/v1/chat_completions_test.go:13-44 (bodies get a "model" because parsing starts at ParseRequest)."""
    M = '"model": "test-model", '
    cases = [
        ('{%s"messages": []}', 9, ""),
        ('{%s"messages": [{"role": "user", "content": "abc"}]}', 0, ""),
        ('{%s"messages": [{"role": "user", "content": "abc"}]}', 9, "abc"),
        ('{%s"messages": [{"role": "user", "content": "abcefghijk"}]}', 9, "abcefghij"),
        ('{%s"messages": [{"role": "user", "content": "世界"}]}', 0, ""),
        ('{%s"messages": [{"role": "user", "content": "世界"}]}', 1, "世"),
        ('{%s"messages": [{"role": "user", "content": "世界"}]}', 2, "世界"),
        ('{%s"messages": [{"role": "user", "content": "世界"}]}', 3, "世界"),
        ('{%s"messages": [{"role": "user", "content": "abc"}, {"role": "user", "content": "xyz"}]}', 9, "abc"),
        ('{%s"messages": [{"role": "system", "content": "abc"}, {"role": "user", "content": "xyz"}]}', 0, ""),
        ('{%s"messages": [{"role": "system", "content": "abc"}, {"role": "user", "content": "xyz"}]}', 9, "xyz"),
        ('{%s"messages": [{"role": "system", "content": "abc"}]}', 9, ""),
        ('{%s"model2": 1}', 9, ""),                                              # no messages at all
        # content as an array of parts is concatenated (chat_completions.go:531-536); null content is "" not a crash
        ('{%s"messages": [{"role": "user", "content": [{"type": "text", "text": "ab"}, {"type": "text", "text": "cd"}]}]}', 3, "abc"),
        ('{%s"messages": [{"role": "user", "content": null}]}', 3, ""),
    ]
    for body, n, exp in cases:
        st, r = psrv.parse_request("/v1/chat/completions", body % M, prefix_chars=n)
        assert st == 0 and r["prefix"] == exp, (body, n, r)


def test_completion_prefix_table_and_first_n_chars(psrv):
    """completions_test.go:14-38, utils_test.go:10-32, completions.go:139-164 (string | [string] | token ids)."""
    M = '"model": "test-model"'
    cases = [('{%s}', 9, ""), ('{%s, "prompt": "abc"}', 0, ""), ('{%s, "prompt": "abc"}', 9, "abc"),
             ('{%s, "prompt": "abcefghijk"}', 9, "abcefghij"), ('{%s, "prompt": "世界"}', 1, "世"),
             ('{%s, "prompt": "世界"}', 2, "世界"), ('{%s, "prompt": "世界"}', 3, "世界"),
             ('{%s, "prompt": ["xyz", "abc"]}', 2, "xy"), ('{%s, "prompt": []}', 2, ""), ('{%s, "prompt": [1, 2, 3]}', 2, "")]
    for body, n, exp in cases:
        st, r = psrv.parse_request("/v1/completions", body % M, prefix_chars=n)
        assert st == 0 and r["prefix"] == exp, (body, n, r)
    for text, n, exp in [("", 0, ""), ("", 1, ""), ("abc", 0, ""), ("abc", 1, "a"), ("abc", 2, "ab"), ("abc", 3, "abc"),
                         ("abc", 4, "abc"), ("世界", 1, "世"), ("世界", 2, "世界"), ("世界", 3, "世界")]:
        st, r = psrv.parse_request("/v1/completions", json.dumps({"model": "test-model", "prompt": text}), prefix_chars=n)
        assert r["prefix"] == exp


def test_parse_request_and_split_model_adapter_tables(psrv):
    """internal/apiutils/request_test.go:13-84 and model_test.go:10-78."""
    cases = [
        ('{"model": "test-model"}', "/v1/chat/completions", "test-model", "", ""),
        ('{"model": "test-model_test-adapter"}', "/v1/chat/completions", "test-model", "test-adapter", ""),
        ('{"model": "test-model", "messages": [{"role": "system", "content": "test"}]}', "/v1/chat/completions", "test-model", "", ""),
        ('{"model": "test-model", "messages": [{"role": "user", "content": "test-prefix"}]}', "/v1/chat/completions", "test-model", "", "test-prefi"),
        ('{"model": "test-model", "prompt": "test-prefix"}', "/v1/completions", "test-model", "", "test-prefi"),
    ]
    for body, path, model, adapter, prefix in cases:
        st, r = psrv.parse_request(path, body, prefix_chars=10)
        assert st == 0 and (r["model"], r["adapter"], r["prefix"]) == (model, adapter, prefix), (body, r)
    # SplitModelAdapter: first "_" splits; the rest belongs to the adapter
    st, r = psrv.parse_request("/v1/chat/completions", '{"model": "test-model_my-adapter_extra"}')
    assert st == 0 and (r["model"], r["adapter"]) == ("test-model", "my-adapter_extra")
    st, r = psrv.parse_request("/v1/chat/completions", '{"model": "test-model_"}')     # trailing separator: no adapter
    assert st == 0 and (r["model"], r["adapter"]) == ("test-model", "")
    st, r = psrv.parse_request("/v1/chat/completions", '{"model": ""}')
    assert st == 400 and r["error"] == "bad request: reading model from body: missing 'model' field"
    st, r = psrv.parse_request("/v1/rerank", '{"model": "test-model", "query": "q", "documents": ["d"]}')
    assert st == 400 and "unknown path" in r["error"]     # rerank/embeddings engines are out of scope here
    # unknown fields (vLLM extensions) never break parsing; \u escapes and surrogate pairs decode to UTF-8
    st, r = psrv.parse_request("/v1/chat/completions", '{"model":"test-model","top_k":5,"nested":{"a":[1,2,{"b":null}]},'
                               '"messages":[{"role":"user","content":"\\u4e16\\ud83d\\ude00x"}]}', prefix_chars=2)
    assert st == 0 and r["prefix"] == "世😀"


def test_unsupported_request_fields_are_refused_not_dropped(psrv):
    """Fields of the reference's schema (api/openai/v1/chat_completions.go:361-470) whose semantics the engine does not
    implement get a 400 at parse time; harmless ones (seed with greedy decoding, user, stream_options) pass."""
    base = {"model": "test-model", "messages": [{"role": "user", "content": "hi"}]}
    ok = [{}, {"seed": 7}, {"user": "u"}, {"n": 1}, {"logprobs": False}, {"stop": []}, {"stop": None}, {"top_logprobs": 0},
          {"presence_penalty": 0}, {"response_format": {"type": "text"}}, {"stop_token_ids": [5]}, {"tools": []}]
    for extra in ok:
        st, out = psrv.parse_request("/openai/v1/chat/completions", json.dumps({**base, **extra}))
        assert st == 0, (extra, out)
    bad = [({"n": 2}, "n > 1"), ({"best_of": 4}, "best_of"), ({"logprobs": True}, "logprobs"), ({"top_logprobs": 5}, "logprobs"),
           ({"stop": "\n\n"}, "stop strings"), ({"stop": ["a", "b"]}, "stop strings"), ({"echo": True}, "echo"),
           ({"frequency_penalty": 0.5}, "penalties"), ({"repetition_penalty": 1.1}, "penalties"),
           ({"logit_bias": {"5": 1}}, "logit_bias"), ({"tools": [{"type": "function"}]}, "tool calling"),
           ({"response_format": {"type": "json_object"}}, "response_format"), ({"min_tokens": 4}, "min_tokens")]
    for extra, msg in bad:
        st, out = psrv.parse_request("/openai/v1/chat/completions", json.dumps({**base, **extra}))
        assert st == 400 and msg in out["error"] and "not supported" in out["error"], (extra, st, out)
    st, out = psrv.parse_request("/openai/v1/completions", json.dumps({"model": "test-model", "prompt": "x", "suffix": "y"}))
    assert st == 400 and "suffix" in out["error"]
