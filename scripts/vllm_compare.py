#!/usr/bin/env python
"""Side measurement (not bench.py's reference arm): the same multi-turn workload pushed through the vLLM wheel that
ships in this image (0.22, not the 0.10.2 the reference pins — charts/kubeai/values.yaml:45), Llama-3-8B shape with
dummy (random) weights, token-id prompts, greedy, 40 tokens/turn, prefix caching on, 128 sessions in flight, closed loop
like benchmarks/multi-turn-chat-go/benchmark/runner.go:263-352.  Prints one JSON line; run under `timeout`.

  python scripts/vllm_compare.py [--steps 400] [--warmup 30] [--sessions 128]
"""
import argparse
import json
import os
import statistics
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
os.environ.setdefault("VLLM_NO_USAGE_STATS", "1")

VOCAB = 128256
IM_START, IM_END = VOCAB - 2, VOCAB - 1


def turn(role: str, content_ids):
    return [IM_START] + list(role.encode()) + [10] + list(content_ids) + [IM_END, 10]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--sessions", type=int, default=128)
    ap.add_argument("--threads-per-session", type=float, default=2.5)
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--max-batched-tokens", type=int, default=1536)
    args = ap.parse_args()

    from kubeai_b200.server import harness_config, synth_threads, tokenize
    hcfg = harness_config(request_model="llama-3-8b", max_concurrent_threads=args.sessions, max_completion_tokens=40,
                          temperature=0.0, synth_threads=int(args.sessions * args.threads_per_session), seed=2)
    threads = [[tokenize(m["content"]) for m in t["messages"]] for t in synth_threads(hcfg)]

    from transformers import LlamaConfig
    d = tempfile.mkdtemp(prefix="llama8b_cfg_")
    LlamaConfig(vocab_size=VOCAB, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                num_attention_heads=32, num_key_value_heads=8, max_position_embeddings=8192, rms_norm_eps=1e-5,
                rope_theta=500000.0, tie_word_embeddings=False, torch_dtype="bfloat16",
                bos_token_id=IM_START, eos_token_id=IM_END).save_pretrained(d)

    from vllm import LLM, SamplingParams
    t_init = time.perf_counter()
    llm = LLM(model=d, load_format="dummy", skip_tokenizer_init=True, dtype="bfloat16", max_model_len=2048,
              max_num_seqs=args.sessions, max_num_batched_tokens=args.max_batched_tokens, enable_prefix_caching=True,
              gpu_memory_utilization=0.80, enforce_eager=args.eager, seed=0)
    init_s = time.perf_counter() - t_init
    eng = llm.llm_engine
    sp = SamplingParams(temperature=0.0, max_tokens=40, ignore_eos=True, detokenize=False)

    state = {}            # request id -> (thread index, turn index, history ids, submit time)
    next_thread = 0
    ttft, first_seen = [], set()

    def submit(ti, k, hist):
        ids = hist + turn("user", threads[ti][k])
        prompt = ids + [IM_START] + list(b"assistant") + [10]
        rid = f"t{ti}-{k}"
        r = eng.add_request(rid, {"prompt_token_ids": prompt[-2000:]}, sp)
        state[r if isinstance(r, str) and r else rid] = state[rid] = (ti, k, ids, time.perf_counter())

    while next_thread < min(args.sessions, len(threads)):
        submit(next_thread, 0, [])
        next_thread += 1

    steps, marks, out_tokens, prompt_tokens = 0, {}, 0, 0
    seen_len = {}
    step_log = []          # (wall seconds of this step() call, new output tokens, requests that got their first token)
    t_prev = time.perf_counter()
    while eng.has_unfinished_requests() and steps < args.warmup + args.steps:
        outs = eng.step()
        steps += 1
        now = time.perf_counter()
        tok_before, first_before = out_tokens, len(first_seen)
        for o in outs:
            n = len(o.outputs[0].token_ids)
            out_tokens += n - seen_len.get(o.request_id, 0)
            seen_len[o.request_id] = n
            if n > 0 and o.request_id not in first_seen:
                first_seen.add(o.request_id)
                if steps > args.warmup:
                    ttft.append(now - state[o.request_id][3])
            if o.finished:
                ti, k, ids, _ = state.pop(o.request_id)
                state.pop(f"t{ti}-{k}", None)
                seen_len.pop(o.request_id, None)
                prompt_tokens += len(o.prompt_token_ids)
                hist = ids + turn("assistant", o.outputs[0].token_ids)
                if k + 1 < len(threads[ti]):
                    submit(ti, k + 1, hist)
                elif next_thread < len(threads):
                    submit(next_thread, 0, [])
                    next_thread += 1
        if steps > args.warmup:
            step_log.append((now - t_prev, out_tokens - tok_before, len(first_seen) - first_before))
        t_prev = now
        if steps == args.warmup:
            marks["t0"], marks["tok0"] = time.perf_counter(), out_tokens
            t_prev = marks["t0"]
    t1 = time.perf_counter()
    timed_steps = steps - args.warmup
    toks = out_tokens - marks.get("tok0", 0)
    dt = t1 - marks.get("t0", t1)
    ttft.sort()
    dec = sorted(dt for dt, n, f in step_log if f == 0 and n > 0)          # steps that only decoded
    pre = sorted(dt for dt, n, f in step_log if f > 0)                     # steps that finished at least one prefill
    med = lambda v: round(v[len(v) // 2] * 1e3, 3) if v else None
    line = {
        "impl": "vllm-%s (image wheel; reference pins v0.10.2)" % __import__("vllm").__version__,
        "metric": "agg output tok/s, Llama-3-8B multi-turn", "value": round(toks / dt, 1) if dt > 0 else None,
        "unit": "tok/s", "steps": timed_steps, "warmup": args.warmup, "ms_per_step": round(dt / max(1, timed_steps) * 1e3, 3),
        "ttft_ms_p50": round(statistics.median(ttft) * 1e3, 2) if ttft else None,
        "ttft_ms_p99": round(ttft[int(0.99 * (len(ttft) - 1))] * 1e3, 2) if ttft else None,
        "requests_first_token": len(ttft),
        "decode_only_steps": {"n": len(dec), "median_ms": med(dec), "mean_ms": round(sum(dec) / len(dec) * 1e3, 3) if dec else None},
        "steps_completing_prefills": {"n": len(pre), "median_ms": med(pre), "mean_ms": round(sum(pre) / len(pre) * 1e3, 3) if pre else None,
                                      "requests_per_step_mean": round(sum(f for _, _, f in step_log if f > 0) / max(1, len(pre)), 1)}, "init_s": round(init_s, 1), "sessions": args.sessions, "max_num_batched_tokens": args.max_batched_tokens,
        "enforce_eager": args.eager, "timing": "host wall clock around LLMEngine.step() (engine-core process included)",
        "data": "synthetic threads (seed 2), dummy weights",
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
