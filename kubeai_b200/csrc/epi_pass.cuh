// Token-major pass of the fused GEMM epilogues over one staged chunk: bf16 [32 tokens][128 weight rows] in shared memory
// (the accumulator already rounded once to bf16, the rounding point of a bf16 GEMM output), 128 epilogue threads.
// Same arithmetic, in the same order, as the standalone kernels of elementwise.cu and the decode-shape epilogues of
// gemm3_tcgen05.cu (vllm/model_executor/models/llama.py:81-121,316-340, activation.py:138-148,
// rotary_embedding/base.py:140-198 restated).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "gemm3.h"
#include "ptx.cuh"

namespace b200 {

// destination of the pair kernel's fused epilogue on steps of more than 128 tokens (gemm2_tcgen05.cu, mode 2)
struct Gemm2Epi {
  int epi;                 // GEMM3_EPI_PLAIN / RESADD / SILU / ROPE_KV
  __nv_bfloat16* out;      // PLAIN out [T, N]; RESADD residual in/out [T, N]; SILU act [T, N/2]; ROPE_KV the fused qkv buffer
  int ldo;
  const int* positions;    // ROPE_KV
  const int* slots;
  const __nv_bfloat16* cos_sin;
  __nv_bfloat16* kv_layer;
  int Hq, Hkv, max_pos;
  int* flags;              // one int per (unit, CTA rank): stamped with `epoch` when that unit's partial tile is in the workspace
  int epoch;
  int hybrid;              // whole-tile waves first, stream-K over the remaining tiles (A/B knob B200_GEMM_HYBRID=0)
  int wide;                // 64-token epilogue rounds (two tcgen05.ld in flight); 0 = 32-token rounds (A/B knob B200_GEMM_WIDE_EPI=0)
};

namespace epi {

constexpr int kRows = 128;   // weight rows per CTA slab
union V8 {
  uint4 u;
  __nv_bfloat16 h[8];
};
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// residual vectors of this thread's four (token, 8-row) items of the chunk; issued before the accumulator is drained
__device__ __forceinline__ void resadd_prefetch(const Gemm2Epi& E, int et, int t_base, int T, int n0, uint4 (&pref)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int item = et + i * 128, j = item >> 4, vv = item & 15, t = t_base + j;
    pref[i] = (t < T) ? __ldcg(reinterpret_cast<const uint4*>(E.out + static_cast<size_t>(t) * E.ldo + n0 + vv * 8)) : make_uint4(0, 0, 0, 0);
  }
}

// ob: the staged chunk; slab: global 128-row slab index of this CTA's rows; tokens [t_base, t_base + 32) clipped to T
__device__ __forceinline__ void pass(const Gemm2Epi& E, const __nv_bfloat16* ob, int et, int t_base, int T, int slab, const uint4 (&res_pref)[4]) {
  const int n0 = slab * kRows;
  if (E.epi == GEMM3_EPI_PLAIN) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int item = et + i * 128, j = item >> 4, vv = item & 15, t = t_base + j;
      if (t < T) *reinterpret_cast<uint4*>(E.out + static_cast<size_t>(t) * E.ldo + n0 + vv * 8) = *reinterpret_cast<const uint4*>(ob + j * kRows + vv * 8);
    }
  } else if (E.epi == GEMM3_EPI_RESADD) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int item = et + i * 128, j = item >> 4, vv = item & 15, t = t_base + j;
      V8 x, r, z;
      x.u = *reinterpret_cast<const uint4*>(ob + j * kRows + vv * 8);
      r.u = res_pref[i];
#pragma unroll
      for (int e = 0; e < 8; ++e) z.h[e] = __float2bfloat16_rn(__bfloat162float(x.h[e]) + __bfloat162float(r.h[e]));
      if (t < T) *reinterpret_cast<uint4*>(E.out + static_cast<size_t>(t) * E.ldo + n0 + vv * 8) = z.u;
    }
  } else if (E.epi == GEMM3_EPI_SILU) {
    // rows 0..63 of the slab are gate rows, 64..127 the matching up rows; act columns [slab*64, slab*64+64)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int item = et + i * 128, j = item >> 3, vv = item & 7, t = t_base + j;
      V8 g, u, o;
      g.u = *reinterpret_cast<const uint4*>(ob + j * kRows + vv * 8);
      u.u = *reinterpret_cast<const uint4*>(ob + j * kRows + 64 + vv * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const __nv_bfloat16 s = silu_bf16(__bfloat162float(g.h[e]));
        o.h[e] = __float2bfloat16_rn(__bfloat162float(s) * __bfloat162float(u.h[e]));
      }
      if (t < T) *reinterpret_cast<uint4*>(E.out + static_cast<size_t>(t) * E.ldo + slab * 64 + vv * 8) = o.u;
    }
  } else {  // GEMM3_EPI_ROPE_KV: one 128-row slab = one head of the fused qkv projection
    const int head = slab;
    if (head < E.Hq + E.Hkv) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int item = et + i * 128, j = item >> 3, vv = item & 7, t = t_base + j;
        if (t >= T) continue;
        int pos = __ldg(E.positions + t);
        pos = pos < 0 ? 0 : (pos >= E.max_pos ? E.max_pos - 1 : pos);
        const __nv_bfloat16* cs = E.cos_sin + static_cast<size_t>(pos) * 128;
        V8 x1, x2, co, si, o1, o2;
        x1.u = *reinterpret_cast<const uint4*>(ob + j * kRows + vv * 8);
        x2.u = *reinterpret_cast<const uint4*>(ob + j * kRows + 64 + vv * 8);
        co.u = __ldg(reinterpret_cast<const uint4*>(cs + vv * 8));
        si.u = __ldg(reinterpret_cast<const uint4*>(cs + 64 + vv * 8));
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float a = __bfloat162float(x1.h[e]), b = __bfloat162float(x2.h[e]);
          const float cc = __bfloat162float(co.h[e]), sn = __bfloat162float(si.h[e]);
          // every bf16 op rounds, as in elementwise.cu rope_kv_kernel
          const float ac = bf16r(__fmul_rn(a, cc)), bs = bf16r(__fmul_rn(b, sn));
          const float bc = bf16r(__fmul_rn(b, cc)), as = bf16r(__fmul_rn(a, sn));
          o1.h[e] = __float2bfloat16_rn(ac - bs);
          o2.h[e] = __float2bfloat16_rn(bc + as);
        }
        if (head < E.Hq) {
          __nv_bfloat16* dst = E.out + static_cast<size_t>(t) * E.ldo + head * 128;
          *reinterpret_cast<uint4*>(dst + vv * 8) = o1.u;
          *reinterpret_cast<uint4*>(dst + 64 + vv * 8) = o2.u;
        } else {
          const int slot = __ldg(E.slots + t);
          if (slot >= 0) {
            const size_t page_stride = static_cast<size_t>(E.Hkv) * 16 * 128;
            __nv_bfloat16* dst = E.kv_layer + static_cast<size_t>(slot >> 4) * 2 * page_stride +
                                 (static_cast<size_t>(head - E.Hq) * 16 + (slot & 15)) * 128;
            *reinterpret_cast<uint4*>(dst + vv * 8) = o1.u;
            *reinterpret_cast<uint4*>(dst + 64 + vv * 8) = o2.u;
          }
        }
      }
    } else {
      const int vh = head - E.Hq - E.Hkv;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int item = et + i * 128, j = item >> 4, vv = item & 15, t = t_base + j;
        if (t >= T) continue;
        const int slot = __ldg(E.slots + t);
        if (slot < 0) continue;
        const size_t page_stride = static_cast<size_t>(E.Hkv) * 16 * 128;
        __nv_bfloat16* dst = E.kv_layer + static_cast<size_t>(slot >> 4) * 2 * page_stride + page_stride +
                             (static_cast<size_t>(vh) * 16 + (slot & 15)) * 128;
        *reinterpret_cast<uint4*>(dst + vv * 8) = *reinterpret_cast<const uint4*>(ob + j * kRows + vv * 8);
      }
    }
  }
}

}  // namespace epi
}  // namespace b200
