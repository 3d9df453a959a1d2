// b200bench — the reference's benchmarks/multi-turn-chat-go CLI restated (main.go:26-136,158-174): flags or a JSON
// config, threads file (or synthetic threads), seeded shuffle, run, print the Result (+ p50/p99 TTFT).
//   b200bench --base-url http://127.0.0.1:8000/openai --request-model llama-3-8b --threads threads.json \
//             --thread-count 2000 --max-concurrent-threads 300 --max-completion-tokens 40 --seed 2
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../include/b200engine.h"

static bool read_file(const char* path, std::string* out) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out->append(buf, n);
  fclose(f);
  return true;
}

int main(int argc, char** argv) {
  b200_harness_config cfg;
  b200_harness_config_default(&cfg);
  std::string base = "http://127.0.0.1:8000/openai", model = "llama-3-8b", threads_path;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--base-url") base = next();
    else if (a == "--request-model") model = next();
    else if (a == "--threads") threads_path = next();
    else if (a == "--thread-count") cfg.thread_count = atoi(next());
    else if (a == "--max-concurrent-threads") cfg.max_concurrent_threads = atoi(next());
    else if (a == "--max-completion-tokens") cfg.max_completion_tokens = atoi(next());
    else if (a == "--temperature") cfg.temperature = static_cast<float>(atof(next()));
    else if (a == "--seed") cfg.seed = atoll(next());
    else if (a == "--synthetic-threads") cfg.synth_threads = atoi(next());
    else if (a == "--request-timeout") cfg.request_timeout_s = atof(next());
    else if (a == "--synthetic-words") cfg.synth_mean_words = atoi(next());
    else if (a == "--vocab") cfg.vocab = atoi(next());
    else { fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
  }
  cfg.request_model = model.c_str();
  // http://host:port[/prefix]
  std::string hp = base.substr(base.find("//") == std::string::npos ? 0 : base.find("//") + 2);
  hp = hp.substr(0, hp.find('/'));
  std::string host = hp.substr(0, hp.find(':'));
  int port = hp.find(':') == std::string::npos ? 80 : atoi(hp.c_str() + hp.find(':') + 1);
  std::string threads;
  if (!threads_path.empty() && !read_file(threads_path.c_str(), &threads)) { fprintf(stderr, "cannot read %s\n", threads_path.c_str()); return 1; }
  b200_harness_result r;
  if (b200_harness_run(nullptr, host.c_str(), port, &cfg, threads.empty() ? nullptr : threads.data(), threads.size(), &r)) {
    fprintf(stderr, "run: %s\n", b200_last_error());
    return 1;
  }
  printf("======================= Input =======================\n"
         "         Input thread count: %d\n   Input msgs/thread (mean): %.2f\n"
         "====================== Results ======================\n"
         "                   Duration: %.1fs\n        Failed thread count: %d\n              Request count: %d\n"
         "    Request duration (mean): %.2fms\n  Chunks per request (mean): %.2f\n"
         "              Prompt tokens: %lld (%lld cached)\n          Completion tokens: %lld\n               Total tokens: %lld\n"
         "Output throughput (e2e run): %.2f tok/sec\n Total throughput (e2e run): %.2f tok/sec\n"
         "                TTFT (mean): %.2fms   p50 %.2fms   p99 %.2fms\n                 ITL (mean): %.2fms\n"
         "=====================================================\n",
         r.input_thread_count, r.input_messages_per_thread_mean, r.duration_s, r.failed_threads, r.request_count,
         r.request_duration_mean_s * 1e3, r.chunks_per_request_mean, (long long)r.prompt_tokens, (long long)r.cached_prompt_tokens,
         (long long)r.completion_tokens, (long long)r.total_tokens, r.run_output_throughput, r.run_total_throughput,
         r.ttft_mean_s * 1e3, r.ttft_p50_s * 1e3, r.ttft_p99_s * 1e3, r.itl_mean_s * 1e3);
  if (r.failed_threads) fprintf(stderr, "first error: %s\n", r.first_error);
  return r.failed_threads ? 1 : 0;
}
