// Paged causal attention for the per-step hot path (SURVEY.md §8a K6 decode, K7 chunked prefill).
// Computes softmax(q . K^T / sqrt(128)) . V over the tokens reachable through a sequence's block
// table, GQA 4:1 — the operation the reference backend delegates to FlashInfer/FlashAttention
// (vllm/model_executor/models/llama.py:223-233, vllm/v1/attention/backends/flashinfer.py).
//
// B200 design: this op is an HBM stream (decode reads 4 KiB of K+V per context token per layer),
// so the kernel is organised around the page gather, not around the math:
//   * KV layout is [block][K|V][kv_head][16 tokens][128] so one (page, head) is a contiguous 4 KiB
//     run; 128 threads pull a 64-token tile (4 pages) with 16-byte cp.async into an XOR-swizzled,
//     double-buffered smem tile (32 KiB per stage, 3 CTAs/SM => ~96 KiB in flight per SM).
//   * the 4 query heads of a KV head ride in one m16n8k16 tensor-core tile so K/V bytes are read
//     once per group (fp32 softmax, warp-shuffle row reductions, P rounded to bf16 for P.V).
//   * decode: the CTA's 4 warps each own a 16-token strip of every tile and merge their
//     (max, sum, acc) through smem at the end; prefill: a CTA owns 16 query tokens x 4 heads, one
//     head per warp, and walks the causal prefix.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels.h"
#include "launch.h"
#include "partials.cuh"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kD = 128;
constexpr int kTile = 64;                       // tokens per smem tile
constexpr int kTileBytes = kTile * kD * 2;      // 16 KiB for K, same for V
constexpr int kStageBytes = 2 * kTileBytes;     // K + V
constexpr int kAttnSmem = 2 * kStageBytes;      // double buffered: 64 KiB
constexpr int kFusedStage = (4 + 2) * kD * 2;   // fused decode: rotated q (4 heads), new k, new v rows (bf16)

struct Vec8 {
  union {
    uint4 u;
    __nv_bfloat16 h[8];
  };
};

// FUSED (decode only): the kernel also does this token's RoPE + KV write (K5), reading q|k|v straight from the
// QKV GEMM's output (bf16 or deferred fp32 partials): q is rotated into smem, the new k/v row goes to the paged
// cache for later steps and is patched into the last smem tile for this one — one launch fewer per layer, and the
// gather of the cached context starts before the QKV GEMM has drained (those pages were written by earlier steps).
template <bool DECODE, bool FUSED>
__global__ void __launch_bounds__(128)  // 178 regs -> 2 CTAs/SM; forcing 3 (168 regs, spills) measured no faster
paged_attn_kernel(const __nv_bfloat16* __restrict__ q, int ldq, __nv_bfloat16* __restrict__ out, int ldo,
                  __nv_bfloat16* __restrict__ kv, const int* __restrict__ block_tables, int max_blocks,
                  const AttnWork* __restrict__ work, int Hkv, float scale_log2,
                  const __nv_bfloat16* __restrict__ cos_sin, int max_pos, PartialView pv) {
  static_assert(DECODE || !FUSED, "the fused RoPE/KV-write prologue exists for decode only");
  constexpr int WT = DECODE ? 16 : 64;  // tokens of each tile handled by one warp
  constexpr int NT = WT / 8;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);

  // work[] and block_tables were uploaded before the first kernel of the step: readable ahead of the dependency wait.
  const AttnWork wk = work[blockIdx.y];  // work items are sorted longest-first; heads are the fast grid dimension
  const int kvh = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;

  const int kv_end = wk.q_pos0 + wk.q_count;  // tokens [0, kv_end) are visible to the last query
  const int ntiles = (kv_end + kTile - 1) / kTile;
  const int* btab = block_tables + static_cast<size_t>(wk.seq) * max_blocks;
  const size_t head_page = static_cast<size_t>(16) * kD;            // elements per (page, head)
  const size_t kv_page = static_cast<size_t>(Hkv) * head_page;      // elements per K (or V) page

  // tokens [0, kv_stored) are gathered from the paged cache; FUSED keeps the newest one in smem instead
  const int kv_stored = FUSED ? kv_end - 1 : kv_end;
  auto load_tile = [&](int tile, int stage) {
    const uint32_t kdst = sbase + stage * kStageBytes;
    const uint32_t vdst = kdst + kTileBytes;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = j * 8 + (tid >> 4);  // token row inside the tile
      const int cc = tid & 15;           // 16-byte chunk inside the 256-byte row
      const int tok = tile * kTile + r;
      const bool valid = tok < kv_stored;
      const int blk = valid ? __ldg(btab + (tok >> 4)) : 0;
      const __nv_bfloat16* ksrc =
          kv + (static_cast<size_t>(blk) * 2) * kv_page + kvh * head_page + (tok & 15) * kD + cc * 8;
      const uint32_t off = r * 256 + ((cc ^ (r & 7)) << 4);
      cp_async_16(kdst + off, ksrc, valid);
      cp_async_16(vdst + off, ksrc + kv_page, valid);
    }
  };
  // Both stages are put in flight as early as the data allows.  Decode: every cached token except the newest was
  // written by an earlier step, and work[] / block_tables are step inputs, so tiles that end before the newest token are
  // gathered BEFORE the dependency wait (ptx.cuh griddep_enter) and overlap the RoPE/KV-write kernel upstream.
  // FUSED never reads the newest row from the cache; prefill tiles may hold rows written by this step: wait first.
  const int safe_tiles = FUSED ? ntiles : (DECODE ? wk.q_pos0 / kTile : 0);
  bool waited = FUSED;  // FUSED waits inside its prologue below
  if (!waited && safe_tiles < 1) {
    griddep_enter();
    waited = true;
  }
  load_tile(0, 0);
  cp_async_commit();
  if (!waited && safe_tiles < 2) {
    griddep_enter();
    waited = true;
  }
  if (ntiles > 1) load_tile(1, 1);
  cp_async_commit();
  if (!waited) griddep_enter();

  // ---- Q fragments (A operand, 16 rows x 128 d as 8 k-steps)
  uint32_t qf[8][4];
  if (FUSED) {
    __nv_bfloat16* stg = reinterpret_cast<__nv_bfloat16*>(smem + kAttnSmem);  // [4 q heads | k | v][128]
    griddep_enter();
    const int Hq = 4 * Hkv, HALF = kD / 2;
    const int t = wk.q_tok0;
    const int posr = wk.q_pos0;
    const int pos = posr < 0 ? 0 : (posr >= max_pos ? max_pos - 1 : posr);
    const int blk = __ldg(btab + (posr >> 4));
    __nv_bfloat16* kdst = kv + (static_cast<size_t>(blk) * 2) * kv_page + kvh * head_page + (posr & 15) * kD;
    if (tid < 40) {  // 5 heads (4 q + k) x 8 rotation tasks; arithmetic identical to rope_kv_kernel (elementwise.cu)
      const int h5 = tid >> 3, c = tid & 7;
      const int col = (h5 < 4 ? kvh * 4 + h5 : Hq + kvh) * kD;
      float xa[8], xb[8];
      if (pv.ws) {
        load8_partials(pv, t, col + c * 8, xa);
        load8_partials(pv, t, col + HALF + c * 8, xb);
      } else {
        Vec8 x1, x2;
        const __nv_bfloat16* hp = q + static_cast<size_t>(t) * ldq + col;
        x1.u = *reinterpret_cast<const uint4*>(hp + c * 8);
        x2.u = *reinterpret_cast<const uint4*>(hp + HALF + c * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xa[j] = __bfloat162float(x1.h[j]);
          xb[j] = __bfloat162float(x2.h[j]);
        }
      }
      Vec8 co, si, o1, o2;
      const __nv_bfloat16* cs = cos_sin + static_cast<size_t>(pos) * kD;
      co.u = __ldg(reinterpret_cast<const uint4*>(cs + c * 8));
      si.u = __ldg(reinterpret_cast<const uint4*>(cs + HALF + c * 8));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = xa[j], b = xb[j];
        const float cc = __bfloat162float(co.h[j]), sn = __bfloat162float(si.h[j]);
        const float ac = __bfloat162float(__float2bfloat16_rn(__fmul_rn(a, cc)));
        const float bs = __bfloat162float(__float2bfloat16_rn(__fmul_rn(b, sn)));
        const float bc = __bfloat162float(__float2bfloat16_rn(__fmul_rn(b, cc)));
        const float as = __bfloat162float(__float2bfloat16_rn(__fmul_rn(a, sn)));
        o1.h[j] = __float2bfloat16_rn(ac - bs);
        o2.h[j] = __float2bfloat16_rn(bc + as);
      }
      *reinterpret_cast<uint4*>(stg + h5 * kD + c * 8) = o1.u;
      *reinterpret_cast<uint4*>(stg + h5 * kD + HALF + c * 8) = o2.u;
      if (h5 == 4) {
        *reinterpret_cast<uint4*>(kdst + c * 8) = o1.u;
        *reinterpret_cast<uint4*>(kdst + HALF + c * 8) = o2.u;
      }
    } else if (tid < 56) {  // the v row: 16 chunks
      const int c = tid - 40;
      const int col = (Hq + Hkv + kvh) * kD + c * 8;
      uint4 val;
      if (pv.ws) {
        float f[8];
        load8_partials(pv, t, col, f);
        val.x = pack_bf16x2(f[0], f[1]);
        val.y = pack_bf16x2(f[2], f[3]);
        val.z = pack_bf16x2(f[4], f[5]);
        val.w = pack_bf16x2(f[6], f[7]);
      } else {
        val = *reinterpret_cast<const uint4*>(q + static_cast<size_t>(t) * ldq + col);
      }
      *reinterpret_cast<uint4*>(stg + 5 * kD + c * 8) = val;
      *reinterpret_cast<uint4*>(kdst + kv_page + c * 8) = val;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const int d = ks * 16 + tq * 2;
      qf[ks][0] = g < 4 ? *reinterpret_cast<const uint32_t*>(stg + g * kD + d) : 0u;
      qf[ks][1] = 0u;
      qf[ks][2] = g < 4 ? *reinterpret_cast<const uint32_t*>(stg + g * kD + d + 8) : 0u;
      qf[ks][3] = 0u;
    }
  } else {
    const __nv_bfloat16* r0p = nullptr;
    const __nv_bfloat16* r1p = nullptr;
    if (DECODE) {
      if (g < 4) r0p = q + static_cast<size_t>(wk.q_tok0) * ldq + (kvh * 4 + g) * kD;
    } else {
      const int head = kvh * 4 + warp;
      if (g < wk.q_count) r0p = q + static_cast<size_t>(wk.q_tok0 + g) * ldq + head * kD;
      if (g + 8 < wk.q_count) r1p = q + static_cast<size_t>(wk.q_tok0 + g + 8) * ldq + head * kD;
    }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const int d = ks * 16 + tq * 2;
      qf[ks][0] = r0p ? *reinterpret_cast<const uint32_t*>(r0p + d) : 0u;
      qf[ks][1] = r1p ? *reinterpret_cast<const uint32_t*>(r1p + d) : 0u;
      qf[ks][2] = r0p ? *reinterpret_cast<const uint32_t*>(r0p + d + 8) : 0u;
      qf[ks][3] = r1p ? *reinterpret_cast<const uint32_t*>(r1p + d + 8) : 0u;
    }
  }

  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int wo = DECODE ? warp * 16 : 0;  // this warp's token offset inside a tile

  for (int t = 0; t < ntiles; ++t) {
    cp_async_wait<1>();  // tile t has landed; tile t+1 may still be in flight
    if (FUSED && t == ntiles - 1) {
      // the newest token's k/v row lives in the staging area: the thread that zero-filled its chunk overwrites it
      const int r = kv_end - 1 - t * kTile;
      if ((r & 7) == (tid >> 4)) {
        const int cc = tid & 15;
        const uint8_t* stg = smem + kAttnSmem + 4 * kD * 2;
        uint8_t* kt = smem + (t & 1) * kStageBytes;
        const uint32_t off = r * 256 + ((cc ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(kt + off) = *reinterpret_cast<const uint4*>(stg + cc * 16);
        *reinterpret_cast<uint4*>(kt + kTileBytes + off) = *reinterpret_cast<const uint4*>(stg + kD * 2 + cc * 16);
      }
    }
    __syncthreads();

    const int tok_base = t * kTile + wo;
    if (tok_base < kv_end) {
      const uint32_t kb = sbase + (t & 1) * kStageBytes;
      const uint32_t vb = kb + kTileBytes;
      float s[NT][4];
#pragma unroll
      for (int i = 0; i < NT; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      const int mi = lane >> 3, ri = lane & 7;
      // ---- S = Q K^T
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
        for (int tp = 0; tp < WT / 16; ++tp) {
          const int tok = wo + tp * 16 + (mi >> 1) * 8 + ri;
          const int chunk = ks * 2 + (mi & 1);
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(kb + tok * 256 + ((chunk ^ (tok & 7)) << 4), b0, b1, b2, b3);
          mma_bf16_16816(s[tp * 2], qf[ks], b0, b1);
          mma_bf16_16816(s[tp * 2 + 1], qf[ks], b2, b3);
        }
      }
      // ---- causal mask, online softmax (fp32, base-2)
      const int qp0 = DECODE ? wk.q_pos0 : wk.q_pos0 + g;
      const int qp1 = DECODE ? wk.q_pos0 : wk.q_pos0 + g + 8;
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int kp = tok_base + nt * 8 + tq * 2;
        if (kp > qp0) s[nt][0] = -INFINITY;
        if (kp + 1 > qp0) s[nt][1] = -INFINITY;
        if (kp > qp1) s[nt][2] = -INFINITY;
        if (kp + 1 > qp1) s[nt][3] = -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0 * scale_log2), mn1 = fmaxf(m1, mx1 * scale_log2);
      const float mu0 = mn0 == -INFINITY ? 0.f : mn0, mu1 = mn1 == -INFINITY ? 0.f : mn1;
      const float a0 = exp2f(m0 - mu0), a1 = exp2f(m1 - mu1);
      m0 = mn0;
      m1 = mn1;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = exp2f(s[nt][0] * scale_log2 - mu0);
        s[nt][1] = exp2f(s[nt][1] * scale_log2 - mu0);
        s[nt][2] = exp2f(s[nt][2] * scale_log2 - mu1);
        s[nt][3] = exp2f(s[nt][3] * scale_log2 - mu1);
        ps0 += s[nt][0] + s[nt][1];
        ps1 += s[nt][2] + s[nt][3];
      }
      l0 = l0 * a0 + ps0;
      l1 = l1 * a1 + ps1;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        o[i][0] *= a0;
        o[i][1] *= a0;
        o[i][2] *= a1;
        o[i][3] *= a1;
      }
      // ---- O += P V
#pragma unroll
      for (int kk = 0; kk < WT / 16; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < 8; ++dp) {
          const int tok = wo + kk * 16 + (mi & 1) * 8 + ri;
          const int chunk = dp * 2 + (mi >> 1);
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(vb + tok * 256 + ((chunk ^ (tok & 7)) << 4), b0, b1, b2, b3);
          mma_bf16_16816(o[dp * 2], pa, b0, b1);
          mma_bf16_16816(o[dp * 2 + 1], pa, b2, b3);
        }
      }
    }
    __syncthreads();
    if (t + 2 < ntiles) load_tile(t + 2, t & 1);  // refill the stage just consumed
    cp_async_commit();
  }
  cp_async_wait<0>();

  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

  if (!DECODE) {
    const int head = kvh * 4 + warp;
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    if (g < wk.q_count) {
      __nv_bfloat16* dst = out + static_cast<size_t>(wk.q_tok0 + g) * ldo + head * kD + tq * 2;
#pragma unroll
      for (int dn = 0; dn < 16; ++dn)
        *reinterpret_cast<uint32_t*>(dst + dn * 8) = pack_bf16x2(o[dn][0] * i0, o[dn][1] * i0);
    }
    if (g + 8 < wk.q_count) {
      __nv_bfloat16* dst = out + static_cast<size_t>(wk.q_tok0 + g + 8) * ldo + head * kD + tq * 2;
#pragma unroll
      for (int dn = 0; dn < 16; ++dn)
        *reinterpret_cast<uint32_t*>(dst + dn * 8) = pack_bf16x2(o[dn][2] * i1, o[dn][3] * i1);
    }
  } else {
    // merge the 4 warps' partial softmax states: rows 0..3 of each warp tile are the 4 heads
    float* so = reinterpret_cast<float*>(smem);            // [4 warps][4 heads][128]
    float* sm = so + 4 * 4 * kD;                           // [4][4]
    float* sl = sm + 16;                                   // [4][4]
    if (g < 4) {
#pragma unroll
      for (int dn = 0; dn < 16; ++dn) {
        float2 v = make_float2(o[dn][0], o[dn][1]);
        *reinterpret_cast<float2*>(so + (warp * 4 + g) * kD + dn * 8 + tq * 2) = v;
      }
      if (tq == 0) {
        sm[warp * 4 + g] = m0;
        sl[warp * 4 + g] = l0;
      }
    }
    __syncthreads();
    const int d = tid;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float mm = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) mm = fmaxf(mm, sm[w * 4 + h]);
      float num = 0.f, den = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float mw = sm[w * 4 + h];
        const float wgt = (mw == -INFINITY) ? 0.f : exp2f(mw - mm);
        num += so[(w * 4 + h) * kD + d] * wgt;
        den += sl[w * 4 + h] * wgt;
      }
      out[static_cast<size_t>(wk.q_tok0) * ldo + (kvh * 4 + h) * kD + d] =
          __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
    }
  }
}

}  // namespace

static int attn_attrs() {
  static int state = 0;  // 0 = not set, 1 = ok, -1 = failed
  if (state == 0) {
    const bool ok =
        cudaFuncSetAttribute(paged_attn_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem) ==
            cudaSuccess &&
        cudaFuncSetAttribute(paged_attn_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem) ==
            cudaSuccess &&
        cudaFuncSetAttribute(paged_attn_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kAttnSmem + kFusedStage) == cudaSuccess;
    state = ok ? 1 : -1;
  }
  return state == 1 ? 0 : -3;
}

int paged_attention(const void* q, int ldq, void* out, int ldo, const void* kv_layer, const int* block_tables,
                    int max_blocks, const AttnWork* work, int num_work, int Hq, int Hkv, float scale,
                    int decode, cudaStream_t st) {
  if (num_work <= 0) return 0;
  if (Hq != 4 * Hkv) return -1;
  if (int rc = attn_attrs()) return rc;
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(Hkv, num_work);
  const __nv_bfloat16* qq = static_cast<const __nv_bfloat16*>(q);
  __nv_bfloat16* oo = static_cast<__nv_bfloat16*>(out);
  __nv_bfloat16* kk = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(kv_layer));
  const __nv_bfloat16* none = nullptr;
  if (decode)
    launch_pdl(paged_attn_kernel<true, false>, grid, dim3(128), kAttnSmem, st, qq, ldq, oo, ldo, kk, block_tables,
               max_blocks, work, Hkv, scale_log2, none, 0, no_partials());
  else
    launch_pdl(paged_attn_kernel<false, false>, grid, dim3(128), kAttnSmem, st, qq, ldq, oo, ldo, kk, block_tables,
               max_blocks, work, Hkv, scale_log2, none, 0, no_partials());
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int paged_attention_rope_decode(const void* qkv, int ldq, void* out, int ldo, void* kv_layer, const int* block_tables,
                                int max_blocks, const AttnWork* work, int num_work, int Hq, int Hkv, float scale,
                                const void* cos_sin, int max_pos, cudaStream_t st, PartialView pv) {
  if (num_work <= 0) return 0;
  if (Hq != 4 * Hkv) return -1;
  if (int rc = attn_attrs()) return rc;
  const float scale_log2 = scale * 1.4426950408889634f;
  launch_pdl(paged_attn_kernel<true, true>, dim3(Hkv, num_work), dim3(128), kAttnSmem + kFusedStage, st,
             static_cast<const __nv_bfloat16*>(qkv), ldq, static_cast<__nv_bfloat16*>(out), ldo,
             static_cast<__nv_bfloat16*>(kv_layer), block_tables, max_blocks, work, Hkv, scale_log2,
             static_cast<const __nv_bfloat16*>(cos_sin), max_pos, pv);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace b200
