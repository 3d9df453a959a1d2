"""Live multi-replica routing in ONE process: N engines (one per GPU, each with its own step thread) behind one b200_server,
the load generator driving N x 128 concurrent multi-turn sessions through b200_server_handle; PrefixHash{125, 100, 256}
against LeastLoad.  The experiment of the reference's runs/llama-3.1-8x-l4/run.ipynb:253-435 and
docs/benchmarks/prefix-aware-load-balancing.md:64-80 (same harness parameters: 40 completion tokens, temperature 0,
seeded synthetic conversation threads of the ShareGPT shape).

  python scripts/routing_run.py --gpus 8 [--sessions-per-gpu 128] [--threads-per-session 5] [--out gpurun_out/routing.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def thread_cpu_seconds():
    """CPU seconds per thread name of this process (/proc/self/task/*/stat utime+stime)."""
    out, hz = {}, os.sysconf("SC_CLK_TCK")
    for tid in os.listdir("/proc/self/task"):
        try:
            s = open(f"/proc/self/task/{tid}/stat").read()
        except OSError:
            continue
        name = s[s.index("(") + 1:s.rindex(")")]
        f = s[s.rindex(")") + 2:].split()
        out[name] = out.get(name, 0.0) + (int(f[11]) + int(f[12])) / hz
    return out


def run(strategy_name, args):
    from kubeai_b200.engine import Engine, default_config
    from kubeai_b200.server import LEAST_LOAD, PREFIX_HASH, Server, harness_config, harness_run
    n = args.gpus
    engines = [Engine(default_config(device=i, manual_step=0, max_num_seqs=args.max_num_seqs, max_batched_tokens=1536,
                                     max_model_len=2048, kv_fraction=0.80)) for i in range(n)]
    srv = Server(engines, model="llama-3-8b", strategy=PREFIX_HASH if strategy_name == "PrefixHash" else LEAST_LOAD,
                 mean_load_pct=125, replication=256, prefix_char_length=100)
    sessions = args.sessions_per_gpu * n
    hcfg = harness_config(request_model="llama-3-8b", max_concurrent_threads=sessions, max_completion_tokens=40, temperature=0.0,
                          synth_threads=int(sessions * args.threads_per_session), seed=2)
    cpu0, t0, p0 = thread_cpu_seconds(), time.perf_counter(), time.process_time()
    r = harness_run(hcfg, server=srv)
    wall, cpu = time.perf_counter() - t0, time.process_time() - p0
    cpu1 = thread_cpu_seconds()
    per = []
    for i, e in enumerate(engines):
        st = e.stats()
        per.append({"replica": i, "steps": st.steps, "prompt_tokens": st.prompt_tokens, "cached_prompt_tokens": st.cached_prompt_tokens,
                    "generated_tokens": st.generated_tokens, "preemptions": st.preemptions})
    by_thread = {k: round(cpu1.get(k, 0.0) - cpu0.get(k, 0.0), 2) for k in cpu1 if cpu1.get(k, 0.0) - cpu0.get(k, 0.0) >= 0.5}
    out = {"strategy": strategy_name, "replicas": n, "concurrent_sessions": sessions, "threads": int(sessions * args.threads_per_session),
           "output_tok_s": round(r["run_output_throughput"], 1), "total_tok_s": round(r["run_total_throughput"], 1),
           "ttft_p50_ms": round(r["ttft_p50_s"] * 1e3, 1), "ttft_p99_ms": round(r["ttft_p99_s"] * 1e3, 1), "ttft_mean_ms": round(r["ttft_mean_s"] * 1e3, 1),
           "itl_mean_ms": round(r["itl_mean_s"] * 1e3, 2), "duration_s": round(r["duration_s"], 2), "requests": r["request_count"],
           "failed_threads": r["failed_threads"], "first_error": r.get("first_error", ""), "prompt_tokens": r["prompt_tokens"], "cached_prompt_tokens": r["cached_prompt_tokens"],
           "cached_ratio": round(r["cached_prompt_tokens"] / max(1, r["prompt_tokens"]), 4),
           "completion_tokens": r["completion_tokens"],
           "host_cpu_cores_busy": round(cpu / wall, 2), "host_cpu_seconds_by_thread_name": by_thread, "per_replica": per}
    srv.close()
    for e in engines:
        e.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--sessions-per-gpu", type=int, default=128)
    ap.add_argument("--max-num-seqs", type=int, default=160, help="per-engine batch cap: 125%% of sessions-per-gpu, what CHWBL's "
                    "bounded load lets one replica reach, so an unbalanced pick is served rather than queued")
    ap.add_argument("--threads-per-session", type=float, default=5.0)
    ap.add_argument("--strategies", default="LeastLoad,PrefixHash")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rows = []
    for s in args.strategies.split(","):
        rows.append(run(s, args))
        print(json.dumps(rows[-1]), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
