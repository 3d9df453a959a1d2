// Host interface of the decode-step projection chain (chain_tcgen05.cu): up to four stream-K GEMM phases in one persistent
// launch, software grid barriers between them, elementwise phases that sum the split-K segments on load.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int kChainMaxGemm = 4;
enum { CHAIN_DEFERRED = 0, CHAIN_SILU = 1 };
enum { CHAIN_REDUCE_NONE = 0, CHAIN_REDUCE_RESADD_NORM = 1, CHAIN_REDUCE_ROPE_KV = 2 };

struct ChainGemm {
  CUtensorMap tm_w;        // weight [N, K], box 128 rows x 64 cols (GemmPlan::tm_w)
  CUtensorMap tm_x;        // token tile source, box 64 rows x 64 cols
  int N, K;
  int units;               // CTA pairs that take part (stream-K ranges over them); <= launched pairs
  int mode;                // CHAIN_DEFERRED: split tiles leave fp32 segments (seg_table), complete tiles bf16 in `out`;
                           // CHAIN_SILU: gate_up with rows interleaved in 64-row blocks, act = silu(gate) * up in `out`
  const int2* seg_table;   // CHAIN_DEFERRED: per tile {first segment, #segments} for `units` units (gemm_plan_build_table, ntt = 1)
  __nv_bfloat16* out;
  int ldo;
  int wait_barrier;        // grid barrier that publishes this phase's token tile (0: only the launch dependency)
  int done_barrier;        // grid barrier every CTA arrives at when its part of the phase is written
  int reduce;              // elementwise phase after done_barrier (one token row per CTA)
  int reduce_barrier;      // grid barrier arrived at after the elementwise phase (0: none follows)
  const __nv_bfloat16* norm_w;   // CHAIN_REDUCE_RESADD_NORM: RMSNorm weight of the next projection's input, or nullptr (add only)
};

struct ChainParams {
  ChainGemm g[kChainMaxGemm];
  int n_gemm;
  int T;
  float* ws_def;                 // fp32 segments of the deferred phases [segment][rank][128 tokens][128 rows]
  float* ws_sk;                  // gate_up neighbour exchange [unit][rank][128 tokens][128 rows]
  int* flags;                    // one per (unit, rank), stamped with `epoch`
  int epoch;
  unsigned long long* bar;       // grid barrier counter (monotonic)
  unsigned long long bar_base;   // its value when this launch starts: barrier k passes at bar_base + k * CTAs
  // elementwise phases
  __nv_bfloat16* res;            // residual [T, H], in/out
  __nv_bfloat16* normed;         // RMSNorm output [T, H]
  float eps;
  const int* positions;
  const int* slots;
  const __nv_bfloat16* cos_sin;
  __nv_bfloat16* kv_layer;       // KV pages of the layer whose qkv projection is phase 3
  int Hq, Hkv, max_pos;
  int dbg;                       // timing experiments only: 1 = epilogues release the accumulators undrained
  int prefetch;                  // weight boxes (16 KB) a CTA pulls into L2 ahead of its ring while its token tiles are blocked
  long long* trace;              // debug: 32 %globaltimer stamps per CTA, or nullptr
};

int chain_smem_bytes();
int chain_max_ctas(int* out);    // co-resident CTAs (148 on a B200): the grid of every chain launch
int chain_launch(const ChainParams& p, int ctas, cudaStream_t st);
void chain_set_trace(long long* dev);   // debug (b200_op_gemm_trace)

}  // namespace b200
