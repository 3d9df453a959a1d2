// Chunked-prefill paged attention on the 5th-generation tensor cores: softmax(q K^T / sqrt(128)) V for
// blocks of new tokens on top of a cached prefix, GQA 4:1, the operation the reference backend delegates to FlashInfer's
// prefill kernel (vllm/v1/attention/backends/flashinfer.py; trtllm-gen FMHA on sm_100).
//
// Round 1 ran prefill on the decode kernel's mma.sync path with 16-query tiles: K/V were re-read from L2 once per 16
// queries and the math was issue-bound (VERDICT r1 "What's missing" 4).  Here a CTA owns 64 query tokens x the 4 query heads
// of one KV head = 256 query rows, as TWO 128-row UMMA tiles (32 tokens x 4 heads each) that share every K/V tile:
//   * one driver thread gathers a 64-token K/V tile by TMA (16 boxes: 4 pages x {K,V} x two 64-column halves, 128B swizzle —
//     the same page gather as the decode kernel) and issues every tcgen05.mma:
//         S_i = Q_i K^T          M=128 (rows = token x head), N=64 kv tokens, K=128 dims; Q_i lands by a 3-D TMA box
//                                (64 dims x 4 heads x 32 tokens) straight from the fused qkv buffer, K-major, swizzled
//         PV_i = P_i V           M=128, N=128 dims, K=64 kv tokens; P_i is written to shared memory by the softmax warps as
//                                a K-major swizzled A operand, V is read in place as an MN-major B operand (its rows are kv
//                                tokens = the contraction index), so no transpose of V is ever made
//     accumulators live in TMEM: S_0, S_1 (64 columns each) and PV_0, PV_1 (128 columns each);
//   * two softmax warpgroups (4 warps each, one per 128-row tile, thread = row as tcgen05.ld 32x32b delivers it) run the
//     online softmax in fp32 with base-2 exponentials, rescale the running output in TMEM (tcgen05.ld / .st) only when a row
//     of the warp moved its maximum, and hand P back; while one group is in its softmax the tensor core works for the other.
// Causal masking by absolute position; K/V rows of a page beyond the sequence end are finite (the pool is zero-filled) and
// masked.  Output rounding: one bf16 rounding of the normalised fp32 result, as the decode kernel.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <mutex>
#include <unordered_map>

#include "gemm.h"
#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kD = 128;
constexpr int kKvTile = 64;                     // kv tokens per tile (4 pages)
constexpr int kQTok = 32;                       // query tokens per 128-row UMMA tile (x 4 heads)
constexpr int kHalf = kKvTile * 128;            // [64 tokens][64 dims] bf16: 8 KB
constexpr int kKTile = 2 * kHalf;               // K tile: two 64-column halves, 16 KB
constexpr int kKvStage = 2 * kKTile;            // K + V: 32 KB
constexpr int kStages = 2;
constexpr int kQHalf = 128 * 128;               // [128 rows][64 dims] bf16: 16 KB
constexpr int kQTile = 2 * kQHalf;              // 32 KB per UMMA tile
constexpr int kPTile = 128 * 128;               // [128 rows][64 kv] bf16: 16 KB
constexpr int kSmem = 1024 + kStages * kKvStage + 2 * kQTile + 2 * kPTile + 256;   // 161 KB
constexpr int kThreads = 288;                   // warp 0: driver (TMA + MMA); warps 1-4: rows of tile 0; warps 5-8: tile 1

// Instruction descriptor: D=f32, A=B=bf16, A K-major, B K-major (b_mn = 0) or MN-major (b_mn = 1), M x N.
__device__ __forceinline__ uint32_t idesc_bf16(uint32_t m, uint32_t n, uint32_t b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
// Shared-memory descriptor, 128B swizzle, explicit leading / stride byte offsets (units of 16 B in the descriptor).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const void* tmap, uint32_t bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
prefill_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_q,
                       __nv_bfloat16* __restrict__ out, int ldo, const int* __restrict__ block_tables, int max_blocks,
                       const AttnWork* __restrict__ work, int Hkv, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t kv_s = sbase;                                  // [stage][K half0 | K half1 | V half0 | V half1]
  const uint32_t q_s = kv_s + kStages * kKvStage;               // [tile][half0 | half1]
  const uint32_t p_s = q_s + 2 * kQTile;                        // [tile][128 rows x 128 B]
  const uint32_t bars = p_s + 2 * kPTile;
  auto kv_full = [&](int s) { return bars + 8u * s; };          // TMA -> driver
  auto kv_empty = [&](int s) { return bars + 8u * (2 + s); };   // MMA commit -> driver (stage can be refilled)
  auto s_full = [&](int i) { return bars + 8u * (4 + i); };     // MMA commit -> group i
  auto p_ready = [&](int i) { return bars + 8u * (6 + i); };    // group i (4 warp arrivals) -> driver: P_i written, S_i read, O_i rescaled
  auto pv_full = [&](int i) { return bars + 8u * (8 + i); };    // MMA commit -> group i
  const uint32_t q_bar = bars + 8u * 12;
  const uint32_t tmem_slot = bars + 8u * 13;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - sbase));

  const AttnWork wk = work[blockIdx.y];     // q_count <= 64 query tokens of one sequence
  const int kvh = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv_end = wk.q_pos0 + wk.q_count;                    // kv tokens visible to the last query of the block
  const int ntiles = (kv_end + kKvTile - 1) / kKvTile;
  const int last_page = (kv_end - 1) >> 4;
  const int* btab = block_tables + static_cast<size_t>(wk.seq) * max_blocks;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_kv);
      tma_prefetch_desc(&tm_q);
      for (int s = 0; s < kStages; ++s) {
        mbar_init(kv_full(s), 1);
        mbar_init(kv_empty(s), 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(s_full(i), 1);
        mbar_init(p_ready(i), 4);
        mbar_init(pv_full(i), 1);
      }
      mbar_init(q_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tm_S[2] = {tmem_base, tmem_base + 64};
  const uint32_t tm_PV[2] = {tmem_base + 128, tmem_base + 256};
  // prefill tiles may hold K/V rows and q rows written by this step's upstream kernels: wait before any load
  griddep_enter();

  if (warp == 0) {
    if (lane == 0) {
      auto issue_tile = [&](int tile, int stage) {
        const uint32_t dst = kv_s + stage * kKvStage;
        mbar_arrive_expect_tx(kv_full(stage), kKvStage);
#pragma unroll
        for (int pg = 0; pg < 4; ++pg) {
          int page = tile * 4 + pg;
          page = page > last_page ? last_page : page;
          const int blk = __ldg(btab + page);
          const int row_k = ((blk * 2) * Hkv + kvh) * 16;
          const int row_v = row_k + Hkv * 16;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tma_load_2d(dst + half * kHalf + pg * 2048, &tm_kv, kv_full(stage), half * 64, row_k, kEvictNormal);
            tma_load_2d(dst + kKTile + half * kHalf + pg * 2048, &tm_kv, kv_full(stage), half * 64, row_v, kEvictNormal);
          }
        }
      };
      // Q: two tiles x two 64-dim halves, rows ordered (token, head) by the 3-D box
      mbar_arrive_expect_tx(q_bar, 2 * kQTile);
      for (int i = 0; i < 2; ++i)
        for (int half = 0; half < 2; ++half)
          tma_load_3d(q_s + i * kQTile + half * kQHalf, &tm_q, q_bar, half * 64, kvh * 4, wk.q_tok0 + i * kQTok);
      for (int s = 0; s < kStages && s < ntiles; ++s) issue_tile(s, s);
      mbar_wait(q_bar, 0);

      const uint32_t idesc_s = idesc_bf16(128, kKvTile, 0);
      const uint32_t idesc_pv = idesc_bf16(128, kD, 1);
      uint32_t ph_p[2] = {0, 0};
      for (int t = 0; t < ntiles; ++t) {
        const int stage = t % kStages;
        mbar_wait(kv_full(stage), static_cast<uint32_t>(t / kStages) & 1u);
        tc_fence_after();
        const uint32_t ks = kv_s + stage * kKvStage, vs = ks + kKTile;
        // S_i = Q_i K^T.  S_i of the previous tile has been read: group i's p_ready arrival of tile t-1 (waited below
        // before PV_i(t-1) was issued) came after its tcgen05.ld of S_i.
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) {
            const uint32_t koff = (k >> 2) * kQHalf + (k & 3) * 32;     // 64 dims per half, 32 B per k-step inside it
            const uint32_t kboff = (k >> 2) * kHalf + (k & 3) * 32;
            umma_bf16(tm_S[i], smem_desc_sw128(q_s + i * kQTile + koff, 16, 1024), smem_desc_sw128(ks + kboff, 16, 1024), idesc_s,
                      k > 0 ? 1u : 0u);
          }
          umma_commit(s_full(i));
        }
        // PV_i += P_i V once group i has written P_i (and rescaled the running output to this tile's maximum)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          mbar_wait(p_ready(i), ph_p[i]);
          ph_p[i] ^= 1u;
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kKvTile / 16; ++k) {
            // A = P_i: K-major rows of 128 B (64 kv), k-step = 32 B; B = V: MN-major, 16 kv rows per k-step = 2048 B,
            // leading offset = the second 64-dim half, stride = 8 rows
            umma_bf16(tm_PV[i], smem_desc_sw128(p_s + i * kPTile + k * 32, 16, 1024), smem_desc_sw128(vs + k * 2048, kHalf, 1024), idesc_pv,
                      (t > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(pv_full(i));
        }
        umma_commit(kv_empty(stage));
        if (t + kStages < ntiles) {
          mbar_wait(kv_empty(stage), static_cast<uint32_t>(t / kStages) & 1u);
          issue_tile(t + kStages, stage);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax groups: thread = one (token, head) row
    const int gi = (warp - 1) >> 2;                 // 0: rows of tile 0, 1: rows of tile 1
    const int q = warp & 3;                         // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;                  // row of the 128-row tile = token_local * 4 + head
    const int tl = gi * kQTok + (row >> 2);         // token inside the 64-token block
    const int head = kvh * 4 + (row & 3);
    const bool valid = tl < wk.q_count;
    const int q_pos = wk.q_pos0 + tl;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* prow = smem + (p_s - sbase) + gi * kPTile + row * 128;

    float m = -INFINITY, l = 0.f;
    uint32_t ph_s = 0, ph_pv = 0;

    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(s_full(gi), ph_s);
      ph_s ^= 1u;
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(tm_S[gi] + lane_off, s0);
      tmem_ld_32x32(tm_S[gi] + lane_off + 32, s1);
      tmem_ld_wait();
      const int kv0 = t * kKvTile;
      const bool edge = kv0 + kKvTile - 1 > q_pos || !valid;     // tile reaches past this row's position: mask
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float a = __uint_as_float(s0[c]), b = __uint_as_float(s1[c]);
        if (edge && (kv0 + c > q_pos || !valid)) a = -INFINITY;
        if (edge && (kv0 + 32 + c > q_pos || !valid)) b = -INFINITY;
        s0[c] = __float_as_uint(a);
        s1[c] = __float_as_uint(b);
        mx = fmaxf(mx, fmaxf(a, b));
      }
      const float mn = fmaxf(m, mx * scale_log2);
      const float mu = mn == -INFINITY ? 0.f : mn;
      const float alpha = exp2f(m - mu);            // m = -inf -> 0
      m = mn;
      float ps = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float p0 = exp2f(__uint_as_float(s0[c]) * scale_log2 - mu), p1 = exp2f(__uint_as_float(s0[c + 1]) * scale_log2 - mu);
        const float p2 = exp2f(__uint_as_float(s1[c]) * scale_log2 - mu), p3 = exp2f(__uint_as_float(s1[c + 1]) * scale_log2 - mu);
        ps += (p0 + p1) + (p2 + p3);
        pk[c >> 1] = pack_bf16x2(p0, p1);
        pk[16 + (c >> 1)] = pack_bf16x2(p2, p3);
      }
      l = l * alpha + ps;
      // the running output lives in TMEM (PV_i accumulates over the tiles): once the previous tile's P V has landed,
      // rescale it to this tile's reference maximum — only when some row of the warp actually moved its maximum
      if (t > 0) {
        mbar_wait(pv_full(gi), ph_pv);
        ph_pv ^= 1u;
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
          for (int c0 = 0; c0 < kD; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tm_PV[gi] + lane_off + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) * alpha);
            tmem_st_32x32(tm_PV[gi] + lane_off + c0, v);
          }
          tmem_st_wait();
        }
      }
      // P row -> smem, K-major 128B-swizzled: logical 16-byte chunk ch lands at ch ^ (row & 7)
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 val = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4)) = val;
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(gi));
    }
    // the finished output
    mbar_wait(pv_full(gi), ph_pv);
    tc_fence_after();
    const float inv = l > 0.f ? 1.f / l : 0.f;
    __nv_bfloat16* dst = out + static_cast<size_t>(wk.q_tok0 + tl) * ldo + head * kD;
#pragma unroll
    for (int c0 = 0; c0 < kD; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tm_PV[gi] + lane_off + c0, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[c]) * inv, __uint_as_float(v[c + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(v[c + 2]) * inv, __uint_as_float(v[c + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(v[c + 4]) * inv, __uint_as_float(v[c + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(v[c + 6]) * inv, __uint_as_float(v[c + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c0 + c) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// 3-D view of the fused qkv buffer [rows][heads][128]: box = 64 dims x 4 heads x 32 tokens, 128B swizzle
int q_map_for(const void* qkv, int rows, int ldq, CUtensorMap* out) {
  struct Key {
    const void* p;
    int rows, ld;
    bool operator==(const Key& o) const { return p == o.p && rows == o.rows && ld == o.ld; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const { return std::hash<const void*>()(k.p) ^ (static_cast<size_t>(k.rows) << 20) ^ static_cast<size_t>(k.ld); }
  };
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  const Key k{qkv, rows, ldq};
  auto it = cache.find(k);
  if (it == cache.end()) {
    CUtensorMap tm;
    if (ldq % kD) return -5;
    if (int rc = tmap_encode_bf16_3d(&tm, qkv, kD, static_cast<uint64_t>(ldq / kD), static_cast<uint64_t>(rows), kD, static_cast<uint64_t>(ldq), 64, 4, kQTok)) return rc;
    if (cache.size() > 1024) cache.clear();
    it = cache.emplace(k, tm).first;
  }
  *out = it->second;
  return 0;
}

}  // namespace

int kv_map_for(const void* kv_layer, CUtensorMap* out);   // attention.cu

int paged_attention_prefill_tc(const void* qkv, int qkv_rows, int ldq, void* out, int ldo, const void* kv_layer, const int* block_tables,
                               int max_blocks, const AttnWork* work, int num_work, int Hq, int Hkv, float scale, cudaStream_t st) {
  if (num_work <= 0) return 0;
  if (Hq != 4 * Hkv) return -1;
  CUtensorMap tm_kv, tm_q;
  if (int rc = kv_map_for(kv_layer, &tm_kv)) return rc;
  if (int rc = q_map_for(qkv, qkv_rows, ldq, &tm_q)) return rc;
  static std::atomic<unsigned long long> attr_done{0};
  if (!ensure_dynamic_smem(prefill_attn_tc_kernel, kSmem, &attr_done)) return -3;
  launch_pdl(prefill_attn_tc_kernel, dim3(Hkv, num_work), dim3(kThreads), kSmem, st, tm_kv, tm_q, static_cast<__nv_bfloat16*>(out), ldo,
             block_tables, max_blocks, work, Hkv, scale * 1.4426950408889634f);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace b200
