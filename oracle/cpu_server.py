"""TEST/BENCH INFRASTRUCTURE — BASELINE.json configs[0] ("plumbing"): the reference's own CPU-runnable case,
benchmarks/multi-turn-chat-go -> openaiserver/modelproxy -> loadbalancer -> local CPU backend (qwen2:0.5b),
restated so that it can be timed on the GPU box's host cores next to the engine (bench.py cpu_baseline leg).

What runs here, and which reference code each part follows:
  * the load generator is NOT here: bench.py drives this server with the product's own C++ restatement of
    benchmarks/multi-turn-chat-go (b200_harness_run, HTTP transport) and the reference's hack/ollama-config.json
    parameters (qwen2:0.5b, 2 concurrent threads, 4 threads, 10 completion tokens);
  * request path: internal/openaiserver/handler.go:20-49 (route), internal/apiutils/request.go:64-225 (body ->
    model, prefix = first user message's first 100 runes), internal/modelproxy/handler.go:57-159 (in-flight
    accounting around the proxied call), internal/loadbalancer/group.go:60-106 (AwaitBestAddress) — through
    oracle/router_oracle.py, the same restatement the router tests pin on the reference's tables;
  * backend: the reference delegates to an Ollama pod (hack/dev-models/kind-cpu.yaml:1-24, qwen2:0.5b).  Ollama and
    its weights are not in this image, so the backend is the oracle's CPU restatement of the model math
    (oracle/llama_oracle.py) at qwen2:0.5b's dimensions (24 layers, hidden 896, 14 query / 2 KV heads of 64,
    intermediate 4864, vocab 151936; random-init, no QKV bias), greedy, one KV cache per conversation prefix;
  * wire format: SSE chunks shaped like api/openai/v1/reference/example-requests.vllm.output:139-160.

Never imported by kubeai_b200/ (tests/test_abi.py checks); the product path has no CPU backend.
"""
from __future__ import annotations

import json
import math
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

import torch

from . import llama_oracle as O
from .router_oracle import PREFIX_HASH, Group
from .weights import ModelCfg, cos_sin_cache

QWEN2_05B = dict(num_layers=24, hidden=896, q_heads=14, kv_heads=2, intermediate=4864, vocab=151936, head_dim=64,
                 rms_eps=1e-6, rope_theta=1000000.0, max_model_len=4096)


# ---------------------------------------------------------------------------------------------- tokenizer
# the synthetic byte/word tokenizer of kubeai_b200/csrc/hostutil.h, restated (ids 0..255 bytes, " wxyz" = id)
def _piece(t: int) -> str:
    s = []
    for _ in range(4):
        s.append(chr(ord("a") + t % 26))
        t //= 26
    return " " + "".join(reversed(s))


def encode(text: str, vocab: int) -> list:
    b = text.encode("utf-8")
    out, i, n = [], 0, len(b)
    while i < n:
        if b[i] == 0x20 and i + 5 <= n and all(97 <= c <= 122 for c in b[i + 1:i + 5]) and \
                (i + 5 == n or not 97 <= b[i + 5] <= 122):
            v = 0
            for c in b[i + 1:i + 5]:
                v = v * 26 + (c - 97)
            if v < vocab:
                out.append(v)
                i += 5
                continue
        out.append(b[i])
        i += 1
    return out


def chat_prompt(messages: list, vocab: int) -> list:
    im_start, im_end = vocab - 2, vocab - 1
    out = []
    for m in messages:
        out += [im_start] + encode(m.get("role", ""), vocab) + [10] + encode(m.get("content") or "", vocab) + [im_end, 10]
    if not messages or messages[-1].get("role") != "assistant":
        out += [im_start] + encode("assistant", vocab) + [10]
    return out


# ---------------------------------------------------------------------------------------------- backend
class CpuBackend:
    """Greedy decoding with the oracle's forward; finished conversations keep their KV so the next turn of the same
    thread only computes its new tokens (what Ollama's prompt cache / vLLM's prefix cache do)."""

    def __init__(self, shape=None, seed=0, threads=None, keep=8):
        s = dict(QWEN2_05B if shape is None else shape)
        self.cfg = ModelCfg(**s)
        c = self.cfg
        if threads:
            torch.set_num_threads(threads)
        g = torch.Generator().manual_seed(seed)
        rnd = lambda *sh, sc=1.0: O.r(torch.randn(*sh, generator=g) * sc)
        H, I, D = c.hidden, c.intermediate, c.head_dim
        w = {"embed": rnd(c.vocab, H), "final_norm": O.r(torch.ones(H))}
        w["lm_head"] = w["embed"]                     # qwen2:0.5b ties the output projection to the embedding
        for l in range(c.num_layers):
            p = f"layers.{l}."
            w[p + "wqkv"] = rnd(c.qkv_rows, H, sc=1 / math.sqrt(H))
            w[p + "wo"] = rnd(H, c.q_heads * D, sc=1 / math.sqrt(c.q_heads * D))
            w[p + "wgu"] = rnd(2 * I, H, sc=1 / math.sqrt(H))
            w[p + "wdown"] = rnd(H, I, sc=1 / math.sqrt(I))
            w[p + "norm1"] = O.r(torch.ones(H))
            w[p + "norm2"] = O.r(torch.ones(H))
        self.model = O.LlamaOracle.__new__(O.LlamaOracle)
        self.model.cfg = c
        self.model.w = w
        self.model.cs = O.r(torch.from_numpy(cos_sin_cache(c)))
        self.lock = threading.Lock()        # one forward at a time: concurrent requests interleave token by token
        self.cache = []                     # [(tokens, kv)] most recent last
        self.keep = keep
        self.prompt_tokens = self.cached_tokens = self.generated = 0

    def _best_prefix(self, ids):
        best, best_n = None, 0
        for toks, kv in self.cache:
            n = 0
            m = min(len(toks), len(ids) - 1)          # at least one token must be computed
            while n < m and toks[n] == ids[n]:
                n += 1
            if n > best_n:
                best, best_n = kv, n
        return best, best_n

    def generate(self, ids, max_tokens, on_token):
        """Calls on_token(id) for every generated id; returns (prompt_tokens, cached_tokens, completion_tokens)."""
        with self.lock:
            kv, n = self._best_prefix(ids)
            if kv is not None:
                kv = [(k[:n], v[:n]) for k, v in kv]
            logits, kv = self.model.forward(ids[n:], kv_prefix=kv, pos0=n)
        toks = list(ids)
        out = 0
        for _ in range(max_tokens):
            t = int(torch.argmax(logits[-1]))
            toks.append(t)
            out += 1
            on_token(t)
            if out == max_tokens or len(toks) >= self.cfg.max_model_len:
                break
            with self.lock:
                logits, kv = self.model.forward([t], kv_prefix=kv, pos0=len(toks) - 1)
        with self.lock:
            self.cache.append((toks[:-1], kv))       # KV covers every token but the last sampled one
            del self.cache[:-self.keep]
            self.prompt_tokens += len(ids)
            self.cached_tokens += n
            self.generated += out
        return len(ids), n, out


# ---------------------------------------------------------------------------------------------- proxy + HTTP front
class Proxy:
    def __init__(self, backend: CpuBackend, model_name="qwen2:0.5b", prefix_chars=100):
        self.backend, self.model_name, self.prefix_chars = backend, model_name, prefix_chars
        self.group = Group(256)
        self.group.reconcile({"cpu-0": dict(address="127.0.0.1:0")})
        self.mu = threading.Lock()
        self.active = 0          # kubeai_inference_requests_active (internal/metrics/metrics.go:16-27)

    def parse(self, body: bytes):
        try:
            req = json.loads(body)
        except ValueError:
            return None, (400, "unable to parse request: invalid JSON")
        if not isinstance(req, dict) or not isinstance(req.get("model"), str) or not req["model"]:
            return None, (400, "unable to parse model: no model specified")
        if req["model"] != self.model_name:
            return None, (404, f"model not found: {req['model']}")
        msgs = req.get("messages")
        if not isinstance(msgs, list) or not msgs:
            return None, (400, "unable to parse request: no messages")
        prefix = ""
        for m in msgs:                                       # chat_completions.go:525-543
            if m.get("role") == "user":
                prefix = (m.get("content") or "")[:self.prefix_chars]
                break
        return dict(messages=msgs, prefix=prefix, max_tokens=int(req.get("max_tokens") or req.get("max_completion_tokens") or 16),
                    stream=bool(req.get("stream")), usage=bool((req.get("stream_options") or {}).get("include_usage"))), None

    def serve(self, parsed, write):
        with self.mu:
            self.active += 1
            picked = self.group.pick(PREFIX_HASH, "", parsed["prefix"], 125)     # (address, endpoint name)
        try:
            ids = chat_prompt(parsed["messages"], self.backend.cfg.vocab)
            rid, created = f"chatcmpl-{time.time_ns():x}", int(time.time())

            def chunk(delta, finish=None, usage=None):
                d = {"id": rid, "object": "chat.completion.chunk", "created": created, "model": self.model_name,
                     "choices": [] if usage is not None else [{"index": 0, "delta": delta, "logprobs": None, "finish_reason": finish}]}
                if usage is not None:
                    d["usage"] = usage
                write(("data: " + json.dumps(d) + "\n\n").encode())

            held, started = [], []

            def on_token(t):
                if not started:                            # the role chunk leaves when the prefill is done (first token
                    started.append(1)                      # exists), as with vLLM: TTFT covers the prompt computation
                    chunk({"role": "assistant", "content": ""})
                if held:                                   # the last token rides on the finish_reason chunk (vLLM framing)
                    chunk({"content": _piece(held.pop())})
                held.append(t)

            p, c, n = self.backend.generate(ids, parsed["max_tokens"], on_token)
            chunk({"content": _piece(held.pop()) if held else ""}, finish="length")
            if parsed["usage"]:
                chunk(None, usage={"prompt_tokens": p, "completion_tokens": n, "total_tokens": p + n,
                                   "prompt_tokens_details": {"cached_tokens": c}})
            write(b"data: [DONE]\n\n")
        finally:
            with self.mu:
                self.active -= 1
                if picked:
                    self.group.done(picked[1])


class Server:
    def __init__(self, backend: CpuBackend, host="127.0.0.1", port=0, model_name="qwen2:0.5b"):
        proxy = Proxy(backend, model_name)
        class H(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, *a):
                pass

            def do_POST(self):
                body = self.rfile.read(int(self.headers.get("Content-Length") or 0))
                if self.path not in ("/openai/v1/chat/completions", "/v1/chat/completions"):
                    return self._err(404, "not found")
                parsed, err = proxy.parse(body)
                if err:
                    return self._err(*err)
                self.send_response(200)
                self.send_header("Content-Type", "text/event-stream")
                self.send_header("Connection", "close")
                self.end_headers()
                self.close_connection = True

                def write(b):
                    self.wfile.write(b)
                    self.wfile.flush()
                try:
                    proxy.serve(parsed, write)
                except (BrokenPipeError, ConnectionResetError):
                    pass

            def _err(self, code, msg):
                b = json.dumps({"error": {"message": msg}}).encode()
                self.send_response(code)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(b)))
                self.send_header("Connection", "close")
                self.end_headers()
                self.wfile.write(b)
                self.close_connection = True

        self.proxy = proxy
        self.httpd = ThreadingHTTPServer((host, port), H)
        self.httpd.daemon_threads = True
        self.port = self.httpd.server_address[1]
        self.thread = threading.Thread(target=self.httpd.serve_forever, daemon=True)
        self.thread.start()

    def close(self):
        self.httpd.shutdown()
        self.httpd.server_close()
