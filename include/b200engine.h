/* b200engine.h — C ABI of the in-process B200 inference engine.
 *
 * This is the drop-in boundary described in SURVEY.md §8(b).  The reference (kubeai-project/kubeai)
 * has no FFI: a backend is an HTTP server reached through
 *     internal/modelproxy/handler.go:158   proxy.ServeHTTP(w, pr.httpRequest())
 * after internal/loadbalancer/load_balancer.go:191 AwaitBestAddress picked an "ip:port".
 * The replacement keeps those Go interfaces (modelproxy/handler.go:18-25) and swaps the HTTP hop for
 * the calls below (cgo binding shown in INTEGRATION.md).  Every function returns 0 on success and a
 * negative b200_status on failure; b200_last_error() gives the message for the calling thread.
 * Plain pointers and sizes only; the caller owns every buffer it passes; the engine copies before
 * returning.  No exceptions cross this boundary.
 */
#ifndef B200ENGINE_H
#define B200ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B200_OK = 0,
  B200_ERR_INVALID = -1,   /* bad argument / shape */
  B200_ERR_CUDA = -2,      /* CUDA runtime / launch failure */
  B200_ERR_OOM = -3,       /* device or KV pool exhausted */
  B200_ERR_NOT_FOUND = -4, /* unknown request id / tensor name */
  B200_ERR_NO_DEVICE = -5, /* no CUDA device: the product path never falls back to CPU */
  B200_ERR_CANCELLED = -6, /* context cancelled while waiting (router) */
  B200_ERR_TIMEOUT = -7
} b200_status;

const char* b200_last_error(void);
/* "kubeai-b200 <version> sm_100a" */
const char* b200_version(void);

/* ------------------------------------------------------------------ engine (one per GPU replica)
 * Replaces the backend pod that internal/modelcontroller/engine_vllm.go:82-100 launches. */

typedef struct b200_engine b200_engine;

typedef struct {
  int32_t device;             /* CUDA ordinal */
  /* model architecture (Llama family; head_dim is fixed at 128, q_heads == 4 * kv_heads) */
  int32_t num_layers, hidden, q_heads, kv_heads, intermediate, vocab;
  float rms_eps;              /* 1e-5 */
  float rope_theta;           /* 500000 */
  int32_t max_model_len;      /* vLLM --max-model-len */
  int32_t max_num_seqs;       /* vLLM --max-num-seqs */
  int32_t max_batched_tokens; /* vLLM --max-num-batched-tokens (per-step token budget) */
  int64_t num_kv_blocks;      /* 16-token pages in the pool; 0 = size from kv_fraction of free HBM */
  float kv_fraction;          /* vLLM --gpu-memory-utilization analogue (of free memory), default 0.85 */
  int32_t enable_prefix_caching;
  int32_t eos_token_id;       /* -1 = none */
  uint64_t seed;              /* seeded random-init weights (no checkpoints offline) */
  float init_scale;           /* multiplies the lm_head init so greedy logits are separated; default 4 */
  int32_t manual_step;        /* 1 = no background thread; caller drives b200_engine_step() */
  int32_t record_steps;       /* >0 = keep the device-side inputs of the last N steps for b200_engine_replay */
} b200_config;

typedef struct {
  int32_t max_tokens;
  float temperature;          /* only greedy (temperature < 1e-5, vllm/v1/sample/sampler.py:17) is implemented */
  int32_t ignore_eos;
  int32_t num_stop_ids;
  const int32_t* stop_ids;
} b200_sampling;

typedef struct {
  int32_t prompt_tokens;
  int32_t cached_tokens;      /* usage.prompt_tokens_details.cached_tokens (api/openai/v1/usage.go:5-25) */
  int32_t completion_tokens;
} b200_usage;

enum { B200_RUNNING = 0, B200_FINISH_STOP = 1, B200_FINISH_LENGTH = 2, B200_FINISH_ABORTED = 3, B200_FINISH_ERROR = 4 };

typedef struct {
  int64_t steps;              /* forward steps executed */
  int32_t running, waiting;
  int64_t kv_blocks_total, kv_blocks_free;
  int64_t prompt_tokens, cached_prompt_tokens, generated_tokens;
  int64_t preemptions;
  double last_step_device_us; /* CUDA-event time of the last forward */
  double total_device_us;     /* sum over steps */
  int64_t last_step_tokens;   /* tokens in the last forward (T) */
  int64_t kernel_launches;    /* kernels launched by this engine so far */
} b200_stats;

typedef struct {
  int32_t tokens;             /* T of the step */
  int32_t decode_seqs, prefill_seqs, sampled;
  int64_t kv_tokens_read;     /* sum of context lengths attended by this step */
  double device_us;
} b200_step_info;

void b200_config_default(b200_config* cfg);          /* Llama-3-8B shape, vLLM-like limits */
int b200_engine_create(const b200_config* cfg, b200_engine** out);
void b200_engine_destroy(b200_engine* e);

int b200_submit(b200_engine* e, const int32_t* prompt_ids, int32_t n, const b200_sampling* sp, uint64_t* req_id);
/* Drain newly generated token ids (at most cap).  *finished gets a B200_FINISH_* code or 0. */
int b200_poll(b200_engine* e, uint64_t req_id, int32_t* out_ids, int32_t cap, int32_t* n_out, int32_t* finished,
              b200_usage* usage);
/* Block until the request has undrained tokens or is finished (B200_ERR_TIMEOUT otherwise). */
int b200_wait(b200_engine* e, uint64_t req_id, int64_t timeout_us);
int b200_abort(b200_engine* e, uint64_t req_id);
/* Forget a finished request (frees its host-side record). */
int b200_release(b200_engine* e, uint64_t req_id);
int b200_stats_get(b200_engine* e, b200_stats* out);

/* manual_step mode: run one scheduler iteration + forward; *info may be NULL. Returns 1 if a step ran, 0 if idle. */
int b200_engine_step(b200_engine* e, b200_step_info* info);
/* Re-run the forward passes of the last `n` recorded steps from their HBM-resident inputs,
 * `repeat` times back to back; returns CUDA-event milliseconds and what those steps carried. */
int b200_engine_replay(b200_engine* e, int32_t n, int32_t repeat, double* ms_total, int64_t* tokens,
                       int64_t* sampled, int64_t* kv_tokens_read, int64_t* launches);
/* Drop every cached prefix block (after a replay, which clobbers KV contents). */
int b200_engine_reset_prefix_cache(b200_engine* e);

/* Weights / buffers by name ("embed", "lm_head", "final_norm", "layers.<i>.{wqkv,wo,wgu,wdown,norm1,norm2}",
 * "cos_sin").  Tests copy weights out for the CPU oracle or write their own in. */
int b200_engine_tensor_info(b200_engine* e, const char* name, uint64_t* num_bytes, void** device_ptr);
int b200_engine_tensor_read(b200_engine* e, const char* name, void* host_dst, uint64_t cap);
int b200_engine_tensor_write(b200_engine* e, const char* name, const void* host_src, uint64_t n);
/* Debug/parity: run a plain forward over `n` tokens of ONE sequence (positions 0..n-1, fresh KV pages taken from
 * the pool and returned afterwards) and copy the bf16 logits of every position to host_logits [n, vocab]. */
int b200_engine_forward_logits(b200_engine* e, const int32_t* ids, int32_t n, void* host_logits_bf16);

/* ------------------------------------------------------------------ router
 * Port of internal/loadbalancer group / CHWBL / LeastLoad (group.go:25-150, balance_chwbl.go:14-162,
 * balance_least_load.go:3-23).  "Endpoints" are GPU replicas; address strings are kept so the
 * reference's tests can be replayed literally. */
typedef struct b200_router b200_router;
enum { B200_LB_LEAST_LOAD = 0, B200_LB_PREFIX_HASH = 1 };

int b200_router_create(int32_t replication, b200_router** out);
void b200_router_destroy(b200_router* r);
/* reconcileEndpoints (group.go:108-137): names/addresses parallel arrays; adapters[i] is a comma-separated list or NULL. */
int b200_router_set_endpoints(b200_router* r, const char* const* names, const char* const* addresses,
                              const char* const* adapters, int32_t n);
/* getBestAddr (group.go:53-88).  Blocks until an endpoint (serving `adapter`) exists or timeout_us passes
 * (<0 = forever, 0 = don't block).  addr_out receives the chosen address; *endpoint_token identifies the
 * endpoint for b200_router_done (the `done func()` of the Go API). */
int b200_router_pick(b200_router* r, int32_t strategy, const char* adapter, const char* prefix, int32_t prefix_len,
                     int32_t mean_load_pct, int64_t timeout_us, char* addr_out, int32_t addr_cap,
                     uint64_t* endpoint_token);
int b200_router_done(b200_router* r, uint64_t endpoint_token);
/* group.addInFlight (group.go:147-150) on a named endpoint; used by the replayed reference tests */
int b200_router_add_inflight(b200_router* r, const char* name, int64_t delta);
int b200_router_inflight(b200_router* r, const char* name, int64_t* endpoint_inflight, int64_t* total_inflight);
/* cespare/xxhash v1.1.0 Sum64 (seed 0) as used at balance_chwbl.go:140-142 */
uint64_t b200_xxh64(const void* data, size_t len);

/* ------------------------------------------------------------------ op-level entry points
 * Raw device pointers (bf16 unless noted) + a cudaStream_t passed as void* (NULL = default stream).
 * These are what the per-kernel parity tests and the ncu captures call. */
int b200_op_gemm(const void* w, const void* x, void* out, int32_t N, int32_t T, int32_t K, void* stream);
int b200_op_embed(const void* table, const int32_t* ids, void* out, int32_t T, int32_t H, int32_t vocab, void* stream);
int b200_op_rmsnorm(const void* x, void* residual, const void* w, void* out, const int32_t* row_index, int32_t rows,
                    int32_t H, float eps, void* stream);
int b200_op_rope_kvwrite(void* qkv, const int32_t* positions, const int32_t* slots, const void* cos_sin,
                         void* kv_layer, int32_t T, int32_t q_heads, int32_t kv_heads, int32_t max_pos, void* stream);
int b200_op_silu_mul(const void* gate_up, void* out, int32_t T, int32_t I, void* stream);
int b200_op_argmax(const void* logits, int32_t* out, int32_t S, int32_t V, int32_t ld, void* stream);
/* work: int32[num_work][4] = {q_tok0, q_count, q_pos0, seq}; kv_layer: [block][2][kv_heads][16][128] */
int b200_op_paged_attn(const void* q, int32_t ldq, void* out, int32_t ldo, const void* kv_layer,
                       const int32_t* block_tables, int32_t max_blocks, const int32_t* work, int32_t num_work,
                       int32_t q_heads, int32_t kv_heads, float scale, int32_t decode, void* stream);
int b200_op_init_uniform(void* p, uint64_t n, uint32_t seed, float scale, float offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200ENGINE_H */
