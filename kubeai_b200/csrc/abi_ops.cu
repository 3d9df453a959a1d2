// Op-level C-ABI entry points (include/b200engine.h "op-level entry points") + error plumbing.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "../../include/b200engine.h"
#include "errors.h"
#include "gemm.h"
#include "gemm3.h"
#include "epi_pass.cuh"
#include "chain.h"
#include "kernels.h"

namespace b200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("no CUDA device available (%s): the B200 path has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return B200_ERR_NO_DEVICE;
  }
  return 0;
}

int cuda_fail(const char* what, int rc) {
  cudaError_t e = cudaGetLastError();
  set_error("%s failed (rc=%d, cuda=%s)", what, rc, cudaGetErrorString(e));
  return rc == -1 ? B200_ERR_INVALID : B200_ERR_CUDA;
}

namespace {
// scratch for b200_op_gemm (the engine owns its own)
std::mutex g_mu;
float* g_ws = nullptr;
size_t g_ws_bytes = 0;
int* g_counters = nullptr;
constexpr int kCounterInts = 1 << 20;

int ensure_scratch() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_ws) return 0;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  g_ws_bytes = std::max<size_t>(gemm_workspace_bytes(sms), 256ull << 20);
  if (cudaMalloc(&g_ws, g_ws_bytes) != cudaSuccess) return -1;
  if (cudaMalloc(&g_counters, kCounterInts * sizeof(int)) != cudaSuccess) return -1;
  cudaMemset(g_counters, 0, kCounterInts * sizeof(int));
  return 0;
}
}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

const char* b200_last_error(void) { return g_err; }
const char* b200_version(void) { return "kubeai-b200 0.1.0 sm_100a"; }

int b200_op_gemm(const void* w, const void* x, void* out, int32_t N, int32_t T, int32_t K, void* stream) {
  if (int rc = require_device()) return rc;
  if (!w || !x || !out || N <= 0 || T < 0 || K <= 0 || K % 8) {
    set_error("b200_op_gemm: bad arguments N=%d T=%d K=%d", N, T, K);
    return B200_ERR_INVALID;
  }
  if (T == 0) return 0;
  if (ensure_scratch()) {
    set_error("b200_op_gemm: workspace allocation failed");
    return B200_ERR_OOM;
  }
  GemmPlan plan;
  int rc = gemm_plan_init(&plan, w, N, K, K, g_ws, g_counters, 0);
  if (rc) return cuda_fail("gemm_plan_init", rc);
  const int bn = gemm_block_n_for(T);
  const int slabs = (N + 127) / 128, ntt = (T + bn - 1) / bn;
  if (2ll * slabs * ntt > kCounterInts) {
    set_error("b200_op_gemm: problem too large for the op-level scratch");
    return B200_ERR_INVALID;
  }
  CUtensorMap tmx;
  rc = gemm_make_x_map(&tmx, x, T, K, K, bn);
  if (rc) return cuda_fail("gemm_make_x_map", rc);
  rc = gemm_run(plan, tmx, bn, out, N, T, static_cast<cudaStream_t>(stream));
  if (rc) return cuda_fail("gemm_run", rc);
  return 0;
}

int b200_op_gemm_deferred(const void* w, const void* x, void* out, int32_t N, int32_t T, int32_t K, void* stream) {
  if (int rc = require_device()) return rc;
  if (!w || !x || !out || N <= 0 || T <= 0 || K <= 0 || K % 8 || gemm_variant() != 2) {
    set_error("b200_op_gemm_deferred: bad arguments N=%d T=%d K=%d (needs the pair kernel)", N, T, K);
    return B200_ERR_INVALID;
  }
  if (ensure_scratch()) { set_error("workspace allocation failed"); return B200_ERR_OOM; }
  GemmPlan plan;
  int rc = gemm_plan_init(&plan, w, N, K, K, g_ws, g_counters, 0);
  if (rc) return cuda_fail("gemm_plan_init", rc);
  plan.ws_bytes = g_ws_bytes;
  if ((rc = gemm_plan_build_table(&plan, T))) return cuda_fail("gemm_plan_build_table", rc);
  const int bn = gemm_block_n_for(T);
  CUtensorMap tmx;
  rc = gemm_make_x_map(&tmx, x, T, K, K, bn);
  PartialView pv = no_partials();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // complete tiles land in `out` directly; the reducer then fills in the split tiles (in place)
  if (!rc) rc = gemm_run_deferred(plan, tmx, bn, out, N, T, st, &pv);
  if (!rc) rc = reduce_partials(pv, out, N, T, N, st);
  cudaStreamSynchronize(st);  // the segment table is freed below
  gemm_plan_destroy(&plan);
  return rc ? cuda_fail("gemm_deferred", rc) : 0;
}

static long long* g_gemm3_trace = nullptr;

int b200_op_gemm3(const b200_gemm3_args* a, void* stream, int32_t* schedule_out) {
  if (int rc = require_device()) return rc;
  if (!a || !a->w || !a->x || a->N <= 0 || a->T <= 0 || a->K <= 0 || a->x_rows < a->T) {
    set_error("b200_op_gemm3: bad arguments");
    return B200_ERR_INVALID;
  }
  if (ensure_scratch()) { set_error("workspace allocation failed"); return B200_ERR_OOM; }
  static float2* cand = nullptr;          // [128 tokens][slabs]
  static int* flags = nullptr;            // stream-K neighbour flags [0, 4096) + fused-norm row flags [4096, 8192)
  static int epoch = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!flags) {
    if (cudaMalloc(&flags, 8192 * sizeof(int)) != cudaSuccess || cudaMemset(flags, 0, 8192 * sizeof(int)) != cudaSuccess) return B200_ERR_OOM;
    if (cudaMalloc(&cand, sizeof(float2) * 128 * 4096) != cudaSuccess) return B200_ERR_OOM;
  }
  if (a->epi == GEMM3_EPI_ARGMAX && (a->N / 128 > 4096 || !a->argmax_out)) { set_error("b200_op_gemm3: argmax needs argmax_out and N <= 524288"); return B200_ERR_INVALID; }
  if (a->T > 128) {
    // steps of more than 128 tokens: the pair kernel (gemm2) with the same fused epilogues, split tiles finished in-kernel
    if (a->pro != GEMM3_PRO_NONE || a->epi == GEMM3_EPI_ARGMAX || a->normed_out || gemm_variant() != 2) {
      set_error("b200_op_gemm3: T > 128 serves PRO_NONE with the PLAIN / RESADD / SILU / ROPE_KV epilogues");
      return B200_ERR_INVALID;
    }
    gemm2_read_env();
    GemmPlan plan;
    int rc = gemm_plan_init(&plan, a->w, a->N, a->K, a->K, g_ws, g_counters, 0);
    if (rc) return cuda_fail("gemm_plan_init", rc);
    plan.ws_bytes = g_ws_bytes;
    const int bn = a->force > 0 ? a->force : gemm_block_n_for(a->T);   // force = token-tile size (256 / 512) in this regime
    CUtensorMap tmx;
    rc = gemm_make_x_map(&tmx, a->x, a->x_rows, a->K, a->K, bn);
    if (rc) return cuda_fail("gemm_make_x_map", rc);
    Gemm2Epi e;
    memset(&e, 0, sizeof(e));
    e.epi = a->epi;
    e.out = static_cast<__nv_bfloat16*>(a->out); e.ldo = a->ldo;
    e.positions = a->positions; e.slots = a->slots; e.cos_sin = static_cast<const __nv_bfloat16*>(a->cos_sin);
    e.kv_layer = static_cast<__nv_bfloat16*>(a->kv_layer); e.Hq = a->q_heads; e.Hkv = a->kv_heads; e.max_pos = a->max_pos;
    e.flags = flags; e.epoch = ++epoch;
    if (schedule_out) { schedule_out[0] = bn; schedule_out[1] = 2; schedule_out[2] = 2 * gemm2_units_for(plan, (a->T + bn - 1) / bn); }
    rc = gemm2_run_fused(plan, tmx, bn, a->T, e, st);
    return rc ? cuda_fail("gemm2_run_fused", rc) : 0;
  }
  Gemm3Schedule sch;
  int rc = gemm3_schedule(a->N, a->K, a->T, a->pro, a->force, &sch);
  if (rc) { set_error("b200_op_gemm3: shape / schedule not served (rc=%d, N=%d K=%d T=%d force=%d)", rc, a->N, a->K, a->T, a->force); return B200_ERR_INVALID; }
  if (gemm3_ws_bytes(sch.grid) > g_ws_bytes) { set_error("b200_op_gemm3: workspace too small"); return B200_ERR_INVALID; }
  GemmPlan plan;
  rc = gemm_plan_init(&plan, a->w, a->N, a->K, a->K, g_ws, g_counters, 0);
  if (rc) return cuda_fail("gemm_plan_init", rc);
  Gemm3Params p;
  memset(&p, 0, sizeof(p));
  p.tm_w = plan.tm_w;
  rc = gemm_make_x_map(&p.tm_x, a->x, a->x_rows, a->K, a->K, 128);
  if (rc) return cuda_fail("gemm_make_x_map", rc);
  p.N = a->N; p.T = a->T; p.K = a->K; p.pro = a->pro; p.epi = a->epi;
  p.ssq_in = a->ssq_in; p.ssq_slabs = a->ssq_slabs; p.norm_w = static_cast<const __nv_bfloat16*>(a->norm_w); p.eps = a->eps;
  p.out = static_cast<__nv_bfloat16*>(a->out); p.ldo = a->ldo; p.ssq_out = a->ssq_out;
  p.positions = a->positions; p.slots = a->slots; p.cos_sin = static_cast<const __nv_bfloat16*>(a->cos_sin);
  p.kv_layer = static_cast<__nv_bfloat16*>(a->kv_layer); p.Hq = a->q_heads; p.Hkv = a->kv_heads; p.max_pos = a->max_pos;
  p.cand = cand; p.n_valid = a->n_valid > 0 ? a->n_valid : a->N;
  if (a->epi == GEMM3_EPI_RESADD && a->normed_out) {
    if (sch.streamk || !a->norm_w_out || a->N / 128 > 1024) { set_error("b200_op_gemm3: the fused norm needs the cluster schedule and norm_w_out"); return B200_ERR_INVALID; }
    p.normed_out = static_cast<__nv_bfloat16*>(a->normed_out);
    p.norm_w_out = static_cast<const __nv_bfloat16*>(a->norm_w_out);
    p.row_flags = flags + 4096;
  }
  p.ws = g_ws; p.flags = flags; p.epoch = ++epoch;
  p.trace = g_gemm3_trace;
  { const char* e = getenv("B200_GEMM3_DBG"); p.dbg = e ? atoi(e) : 0; }
  if (schedule_out) { schedule_out[0] = sch.S; schedule_out[1] = sch.streamk; schedule_out[2] = sch.grid; }
  rc = gemm3_launch(p, sch, st);
  if (rc) return cuda_fail("gemm3_launch", rc);
  if (a->epi == GEMM3_EPI_ARGMAX) {
    rc = argmax_candidates(cand, a->argmax_out, a->T, a->N / 128, st);
    if (rc) return cuda_fail("argmax_candidates", rc);
    cudaStreamSynchronize(st);   // the shared candidate buffer is reused by the next call
  }
  return 0;
}

int b200_schedule_query(int32_t N, int32_t K, int32_t T, int32_t sms, int32_t* out8) {
  if (N <= 0 || K <= 0 || T <= 0 || sms < 2 || !out8) { set_error("b200_schedule_query: bad arguments"); return B200_ERR_INVALID; }
  gemm2_schedule_query(N, K, T, sms, out8);
  return 0;
}

int b200_attn_split_query(int32_t num_work, int32_t kv_heads, int32_t max_ctx, int32_t sms) {
  if (num_work <= 0 || kv_heads <= 0 || max_ctx <= 0 || sms <= 0) return 1;
  return attn_decode_split(num_work, kv_heads, max_ctx, sms);
}

int b200_set_gemm_variant(int32_t v) {
  gemm_set_variant(v);
  return gemm_variant();
}

int b200_op_gemm_trace(void* trace_dev) {
  gemm2_set_trace(static_cast<long long*>(trace_dev));
  g_gemm3_trace = static_cast<long long*>(trace_dev);   // b200_op_gemm3 launches stamp 8 values per CTA while this is set
  chain_set_trace(static_cast<long long*>(trace_dev));  // chain launches: 40 slots x 148 CTAs x 32 stamps
  return 0;
}

int b200_op_embed(const void* table, const int32_t* ids, void* out, int32_t T, int32_t H, int32_t vocab,
                  void* stream) {
  if (int rc = require_device()) return rc;
  int rc = embed_gather(table, ids, out, T, H, vocab, static_cast<cudaStream_t>(stream));
  return rc ? cuda_fail("embed_gather", rc) : 0;
}

int b200_op_rmsnorm(const void* x, void* residual, const void* w, void* out, const int32_t* row_index,
                    int32_t rows, int32_t H, float eps, void* stream) {
  if (int rc = require_device()) return rc;
  int rc = rmsnorm(x, residual, w, out, row_index, rows, H, eps, static_cast<cudaStream_t>(stream));
  return rc ? cuda_fail("rmsnorm", rc) : 0;
}

int b200_op_rope_kvwrite(void* qkv, const int32_t* positions, const int32_t* slots, const void* cos_sin,
                         void* kv_layer, int32_t T, int32_t q_heads, int32_t kv_heads, int32_t max_pos,
                         void* stream) {
  if (int rc = require_device()) return rc;
  int rc = rope_kv_write(qkv, positions, slots, cos_sin, kv_layer, T, q_heads, kv_heads, max_pos,
                         static_cast<cudaStream_t>(stream));
  return rc ? cuda_fail("rope_kv_write", rc) : 0;
}

int b200_op_silu_mul(const void* gate_up, void* out, int32_t T, int32_t I, void* stream) {
  if (int rc = require_device()) return rc;
  int rc = silu_mul(gate_up, out, T, I, static_cast<cudaStream_t>(stream));
  return rc ? cuda_fail("silu_mul", rc) : 0;
}

int b200_op_argmax(const void* logits, int32_t* out, int32_t S, int32_t V, int32_t ld, void* stream) {
  if (int rc = require_device()) return rc;
  int rc = argmax_rows(logits, out, S, V, ld, static_cast<cudaStream_t>(stream));
  return rc ? cuda_fail("argmax_rows", rc) : 0;
}

int b200_op_paged_attn(const void* q, int32_t ldq, void* out, int32_t ldo, const void* kv_layer,
                       const int32_t* block_tables, int32_t max_blocks, const int32_t* work, int32_t num_work,
                       int32_t q_heads, int32_t kv_heads, float scale, int32_t decode, void* stream) {
  if (int rc = require_device()) return rc;
  static_assert(sizeof(AttnWork) == 16, "AttnWork is int32[4] on the wire");
  int rc = paged_attention(q, ldq, out, ldo, kv_layer, block_tables, max_blocks,
                           reinterpret_cast<const AttnWork*>(work), num_work, q_heads, kv_heads, scale, decode,
                           static_cast<cudaStream_t>(stream));
  return rc ? cuda_fail("paged_attention", rc) : 0;
}

int b200_op_paged_attn_decode_split(const void* q, int32_t ldq, void* out, int32_t ldo, const void* kv_layer,
                                    const int32_t* block_tables, int32_t max_blocks, const int32_t* work, int32_t num_work,
                                    int32_t q_heads, int32_t kv_heads, float scale, int32_t split, void* stream) {
  if (int rc = require_device()) return rc;
  if (split < 1 || split > 64 || num_work <= 0) { set_error("b200_op_paged_attn_decode_split: bad arguments"); return B200_ERR_INVALID; }
  static float* ws = nullptr;
  static size_t ws_bytes = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  const size_t need = attn_split_ws_bytes(num_work, kv_heads, split);
  if (need > ws_bytes) {
    cudaDeviceSynchronize();
    if (ws) cudaFree(ws);
    ws = nullptr;
    ws_bytes = 0;
    if (cudaMalloc(&ws, need) != cudaSuccess) { set_error("workspace allocation failed"); return B200_ERR_OOM; }
    ws_bytes = need;
  }
  int rc = paged_attention(q, ldq, out, ldo, kv_layer, block_tables, max_blocks, reinterpret_cast<const AttnWork*>(work), num_work,
                           q_heads, kv_heads, scale, 1, static_cast<cudaStream_t>(stream), ws, split);
  return rc ? cuda_fail("paged_attention(split)", rc) : 0;
}

int b200_op_paged_attn_prefill_tc(const void* qkv, int32_t q_rows, int32_t ldq, void* out, int32_t ldo, const void* kv_layer,
                                  const int32_t* block_tables, int32_t max_blocks, const int32_t* work, int32_t num_work,
                                  int32_t q_heads, int32_t kv_heads, float scale, void* stream) {
  if (int rc = require_device()) return rc;
  int rc = paged_attention_prefill_tc(qkv, q_rows, ldq, out, ldo, kv_layer, block_tables, max_blocks,
                                      reinterpret_cast<const AttnWork*>(work), num_work, q_heads, kv_heads, scale,
                                      static_cast<cudaStream_t>(stream));
  return rc ? cuda_fail("paged_attention_prefill_tc", rc) : 0;
}

int b200_op_init_uniform(void* p, uint64_t n, uint32_t seed, float scale, float offset, void* stream) {
  if (int rc = require_device()) return rc;
  int rc = init_uniform(p, n, seed, scale, offset, static_cast<cudaStream_t>(stream));
  return rc ? cuda_fail("init_uniform", rc) : 0;
}

}  // extern "C"
