// Decode-shape projection GEMM with in-kernel split-K reduction and fused elementwise neighbours (T <= 128 tokens).
//     acc[t, n] = sum_k X'[t, k] * W[n, k]          (bf16 x bf16 -> fp32 in TMEM)
// It replaces, for the steps that are weight streams (every sequence contributes one token: ~80% of the bench's steps),
// the chain  GEMM -> fp32 stream-K segments in L2 -> elementwise kernel -> GEMM  of gemm2_tcgen05.cu + elementwise.cu:
// round 1 measured ~11 us per such boundary, 54 MB of fp32 segments per layer through L2, and a decode step at 0.57 of the
// HBM roofline (profiles/r01_elementwise_ncu.md, VERDICT r1 "What's weak" 4).  Here the reduction finishes inside the GEMM
// and the elementwise neighbours ride in its prologue / epilogue, so a decoder layer is 4 GEMMs + attention, no
// elementwise launch and no fp32 round trip:
//
//   prologue (B operand, the token tile)
//     PRO_NONE   X' = X (bf16 activations, TMA)
//     PRO_NORM   X' = RMSNorm(residual): the residual tile lands by TMA, four "transform" warps normalise it in place
//                in shared memory (x * inv_rms[t] -> bf16 -> * w[k] -> bf16; inv_rms from the per-slab sums of squares the
//                producing GEMM's epilogue left) and hand the stage to the MMA thread.  Rounding points = vLLM's
//                fused_add_rms_norm CUDA op / HF transformers (modeling_llama.py:62-67,325).
//   mainloop    as gemm2: CTA pair, tcgen05.mma.cta_group::2 M=256 (weights are the M side, tokens the N side), TMA
//               128B-swizzle ring, TMEM accumulators, PDL with the weight ring issued before the dependency wait.
//   reduction   cluster mode (N small: qkv / o / down): the S CTA pairs of a cluster split K of ONE 256-row tile; partial
//               accumulators are reduce-scattered through distributed shared memory (each CTA ships the 32-token chunks
//               it does not own to their owners with cp.async.bulk smem->smem, owners add in fp32 in a fixed order).
//               stream-K mode (N large: gate_up / lm_head): contiguous (tile, k-block) ranges over the pairs as in gemm2;
//               a unit's range starts with the TAIL of a tile and ends with the HEAD of another.  The tail is computed
//               first and parked as fp32 in L2 with a release flag — tens of microseconds before the neighbouring unit,
//               which owns the head of that tile and computes it LAST, needs it — so the head owner prefetches it into
//               shared memory under its own mainloop and finishes the tile without a handshake on the critical path.
//   epilogue    on the finished fp32 tile, rounded once to bf16 (the rounding point of a bf16 GEMM output):
//     EPI_PLAIN    out[t, n]
//     EPI_RESADD   residual[t, n] = bf16(acc + residual); per (token, 128-row slab) sum of squares for the next PRO_NORM
//     EPI_SILU     act[t, i] = bf16(bf16(silu(gate)) * up) with gate/up rows interleaved in 64-row blocks in the weight
//     EPI_ROPE_KV  neox RoPE on q/k heads (a 128-row slab is one head), q -> qkv buffer, k/v -> paged KV cache
//     EPI_ARGMAX   per (token, slab) best logit + index; logits are never materialised
//
// Reference behaviour restated: vllm/model_executor/models/llama.py:81-121,223-233,316-340 (layer structure),
// activation.py:138-148, rotary_embedding/base.py:140-198, _custom_ops.py:323-327, v1/sample/sampler.py:91.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gemm.h"
#include "gemm3.h"
#include "launch.h"
#include "ptx.cuh"
#include "umma2.cuh"

namespace b200 {

namespace {

constexpr int kSlab = 128;     // weight rows per CTA (UMMA M = 256 per pair)
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kBN = 128;       // token tile (one tile per launch: T <= 128)
constexpr int kChunkTok = 32;  // tokens per epilogue chunk (one tcgen05.ld 32x32b.x32)
constexpr int kNumChunks = kBN / kChunkTok;
constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue, warps 6-9 transform
constexpr int kABytes = kSlab * kBlockK * 2;            // 16 KB
constexpr int kBBytes = (kBN / 2) * kBlockK * 2;        // 8 KB: this CTA's half of the token tile
constexpr int kStageBytes = kABytes + kBBytes;          // 24 KB
constexpr int kStagesDefault = 6;                       // 144 KB ring
constexpr int kMaxStages = 8;
constexpr int kChunkF32 = kChunkTok * kSlab * 4;        // 16 KB: one chunk of fp32 partials [32 tokens][128 rows]
constexpr int kXbufDefault = 4 * kChunkF32;             // 64 KB: DSMEM receive slots / stream-K partner partial
constexpr int kOpStage = kChunkTok * kSlab * 2;         // 8 KB: bf16 [32 tokens][128 rows] for the fused epilogue pass
constexpr int kWnormMax = 8192 * 2;                     // norm weight (K <= 8192 for PRO_NORM)
constexpr int kMisc = 1024;                             // barriers, tmem slot, inv table (64 floats)
constexpr int smem_bytes_for(int stages, int xbuf) { return 1024 + stages * kStageBytes + xbuf + kOpStage + kMisc; }   // + K*2 for PRO_NORM
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void xf_bar_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

union V8 {
  uint4 u;
  __nv_bfloat16 h[8];
};
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct Seg {
  int tile, kb0, kb1;
};

// kStages x 24 KB ring; kXbuf = 64 KB in production, 0 in the ring-depth experiment (plain S = 1 launches only)
template <int kStages, int kXbuf>
__global__ void __launch_bounds__(kThreads, 1) gemm3_kernel(const __grid_constant__ Gemm3Params P) {
  constexpr int kRing = kStages * kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t xbuf = smem_base + kRing;
  const uint32_t opst = xbuf + kXbuf;
  const uint32_t misc = opst + kOpStage;
  const uint32_t wnorm = misc + kMisc;
  auto full_bar = [&](int s) { return misc + 8u * s; };                    // leader's copy is live: A (and B if PRO_NONE)
  auto empty_bar = [&](int s) { return misc + 8u * (kStages + s); };       // both CTAs (multicast commit)
  auto bfull_bar = [&](int s) { return misc + 8u * (2 * kStages + s); };   // own B half landed (PRO_NORM)
  auto xf_bar = [&](int s) { return misc + 8u * (3 * kStages + s); };      // leader: B halves transformed (8 warp arrivals)
  auto tfull_bar = [&](int a) { return misc + 8u * (4 * kStages + a); };   // both CTAs (multicast commit)
  auto tempty_bar = [&](int a) { return misc + 8u * (4 * kStages + 2 + a); };  // leader: 8 epilogue-warp arrivals
  const uint32_t recv_bar = misc + 8u * (4 * kStages + 4);                 // DSMEM partials / partner partial landed
  const uint32_t tmem_slot = misc + 8u * (4 * kStages + 5);
  float* inv_s = reinterpret_cast<float*>(smem + (misc - smem_base) + 512);   // [64] 1/rms of this CTA's token half
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  // optional phase trace (debug): 8 %globaltimer stamps (ns, one clock for the whole GPU) per CTA
  auto mark = [&](int i) {
    if (P.trace) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      P.trace[static_cast<size_t>(blockIdx.x) * 8 + i] = t;
    }
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 64) mark(0);
  const uint32_t crank = cluster_ctarank();       // rank in the cluster (2S CTAs)
  const uint32_t rank = crank & 1u;               // rank in the pair
  const uint32_t pair = crank >> 1;               // pair index inside the cluster
  const bool leader = rank == 0;
  const uint32_t leader_cta = crank & ~1u;
  const uint16_t pair_mask = static_cast<uint16_t>(3u << leader_cta);

  const int T = P.T, N = P.N, K = P.K;
  const int KB = (K + kBlockK - 1) / kBlockK;
  const int nc = (T + 15) & ~15;                  // MMA N: tokens padded to 16
  const int n_eff = nc;
  const int S = P.S;                              // pairs per tile (cluster mode); stream-K has S == 1
  const int tiles = N / (2 * kSlab);

  // ---- this pair's work: segments (tile, kb0, kb1) in processing order
  long long it_begin, it_end;                     // stream-K iteration range
  int unit = 0, units = 1;
  if (P.streamk) {
    units = gridDim.x >> 1;
    unit = blockIdx.x >> 1;
    const long long total = static_cast<long long>(tiles) * KB;
    it_begin = range_begin(unit, total, units);
    it_end = range_begin(unit + 1, total, units);
  } else {
    const int tile = blockIdx.x / (2 * S);
    const int k0 = static_cast<int>((static_cast<long long>(pair) * KB) / S), k1 = static_cast<int>((static_cast<long long>(pair + 1) * KB) / S);
    it_begin = static_cast<long long>(tile) * KB + k0;
    it_end = static_cast<long long>(tile) * KB + k1;
  }
  auto seg_at = [&](long long it) {
    Seg s;
    s.tile = static_cast<int>(it / KB);
    s.kb0 = static_cast<int>(it - static_cast<long long>(s.tile) * KB);
    const long long rem = it_end - it;
    s.kb1 = (KB - s.kb0 <= rem) ? KB : s.kb0 + static_cast<int>(rem);
    return s;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tm_w);
    tma_prefetch_desc(&P.tm_x);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
        mbar_init(bfull_bar(s), 1);
        mbar_init(xf_bar(s), 8);       // 4 transform warps of each CTA of the pair
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 8);
      }
      mbar_init(recv_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc2(tmem_slot, 2 * kBN);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation visible to every CTA of the cluster
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_launch();
  if (threadIdx.x == 64) mark(1);

  const bool pro_norm = P.pro == GEMM3_PRO_NORM;
  const int row_half0 = static_cast<int>(rank) * (nc >> 1);   // first token of this CTA's half of the tile

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      const uint64_t w_hint = kEvictFirst;
      int pre = 0;
      for (long long it = it_begin; it < it_end && pre < kStages; ++it, ++pre) {
        const int tile = static_cast<int>(it / KB), kb = static_cast<int>(it - static_cast<long long>(tile) * KB);
        if (leader) mbar_arrive_expect_tx(full_bar(pre), pro_norm ? 2u * kABytes : 2u * kStageBytes);
        tma_load_2d_pair(smem_base + pre * kStageBytes, &P.tm_w, full_bar(pre), kb * kBlockK, tile * 2 * kSlab + static_cast<int>(rank) * kSlab,
                         w_hint);
      }
      griddep_wait();
      int stage = 0, idx = 0;
      uint32_t phase = 0;
      for (long long it = it_begin; it < it_end; ++it, ++idx) {
        const int tile = static_cast<int>(it / KB), kb = static_cast<int>(it - static_cast<long long>(tile) * KB);
        const uint32_t sa = smem_base + stage * kStageBytes;
        if (idx >= pre) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (leader) mbar_arrive_expect_tx(full_bar(stage), pro_norm ? 2u * kABytes : 2u * kStageBytes);
          tma_load_2d_pair(sa, &P.tm_w, full_bar(stage), kb * kBlockK, tile * 2 * kSlab + static_cast<int>(rank) * kSlab, w_hint);
        }
        if (pro_norm) {
          mbar_arrive_expect_tx(bfull_bar(stage), kBBytes);
          tma_load_2d(sa + kABytes, &P.tm_x, bfull_bar(stage), kb * kBlockK, row_half0, kEvictLast);
        } else {
          tma_load_2d_pair(sa + kABytes, &P.tm_x, full_bar(stage), kb * kBlockK, row_half0, kEvictLast);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA of the pair)
    if (lane == 0 && leader) {
      griddep_wait();
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t idesc = umma_idesc_bf16(2 * kSlab, nc);
      for (long long it = it_begin; it < it_end;) {
        const Seg sg = seg_at(it);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if (pro_norm) mbar_wait_cluster(xf_bar(stage), phase);
          tc_fence_after();
          if (it == it_begin && kb == sg.kb0) mark(2);
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint64_t a_desc = umma_desc_kmajor_sw128(sa);
          const uint64_t b_desc = umma_desc_kmajor_sw128(sa + kABytes);
          const uint32_t d_tmem = tmem_base + acc * kBN;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma2_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb > sg.kb0 || k > 0) ? 1u : 0u);
          umma2_commit_pair(empty_bar(stage), pair_mask);
          if (kb == sg.kb1 - 1) umma2_commit_pair(tfull_bar(acc), pair_mask);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
        it += sg.kb1 - sg.kb0;
      }
      mark(3);
    }
  } else if (warp >= 6) {
    // ------------------------------------------------------------ transform warps: RMSNorm of the token tile in smem
    if (pro_norm) {
      const int xt = threadIdx.x - 192;   // 0..127
      // static: norm weight -> smem (before the dependency wait)
      for (int i = xt; i < K / 8; i += 128)
        reinterpret_cast<uint4*>(smem + (wnorm - smem_base))[i] = __ldg(reinterpret_cast<const uint4*>(P.norm_w) + i);
      griddep_wait();
      {
        // 1/rms of this CTA's 64 token rows from the per-slab sums of squares: two threads per token, loads issued eight
        // at a time (a dependent chain of L2 loads per slab measured ~10 us per launch)
        const int tr = xt >> 1, part = xt & 1;
        const int t = row_half0 + tr;
        float ss = 0.f;
        if (tr < (nc >> 1) && t < T) {
          const float* p = P.ssq_in + static_cast<size_t>(t) * P.ssq_slabs;
          const int half = (P.ssq_slabs + 1) >> 1, s0 = part * half, s1 = min(P.ssq_slabs, s0 + half);
          for (int s = s0; s < s1; s += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (s + u < s1) ? __ldcg(p + s + u) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) ss += v[u];
          }
        }
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        if (part == 0) inv_s[tr] = (tr < (nc >> 1) && t < T) ? rsqrtf(ss / static_cast<float>(K) + P.eps) : 0.f;
      }
      xf_bar_sync();
      int stage = 0;
      uint32_t phase = 0;
      const int nrows = nc >> 1;  // rows of the box that the MMA reads
      for (long long it = it_begin; it < it_end; ++it) {
        const int kb = static_cast<int>(it % KB);
        mbar_wait(bfull_bar(stage), phase);
        uint8_t* b = smem + (stage * kStageBytes + kABytes);
        const uint8_t* wsm = smem + (wnorm - smem_base) + kb * kBlockK * 2;
        // 64 rows x 8 sixteen-byte chunks, four per thread; 128B swizzle: physical chunk p of row r holds logical chunk
        // p ^ (r & 7).  All loads first, then the math: the four chunks' shared-memory latencies overlap.
        uint4 xv[4], wv[4];
        float inv[4];
        const int nchunks = (P.dbg & 1) ? 0 : nrows * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = xt + i * 128;
          if (c < nchunks) {
            const int r = c >> 3, pch = c & 7;
            xv[i] = *reinterpret_cast<const uint4*>(b + r * 128 + pch * 16);
            wv[i] = *reinterpret_cast<const uint4*>(wsm + ((pch ^ (r & 7)) << 4));
            inv[i] = inv_s[r];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = xt + i * 128;
          if (c < nchunks) {
            // packed math: x * inv in fp32 -> one packed rounding; the product with the norm weight is a bf16 x bf16
            // multiply, exact in fp32, so __hmul2's single rounding equals bf16(float(nb) * float(w))
            const uint32_t xw[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w}, ww[4] = {wv[i].x, wv[i].y, wv[i].z, wv[i].w};
            uint32_t ow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float lo = __uint_as_float(xw[j] << 16) * inv[i], hi = __uint_as_float(xw[j] & 0xffff0000u) * inv[i];
              const __nv_bfloat162 nb = __floats2bfloat162_rn(lo, hi);
              const __nv_bfloat162 prod = __hmul2(nb, *reinterpret_cast<const __nv_bfloat162*>(&ww[j]));
              ow[j] = *reinterpret_cast<const uint32_t*>(&prod);
            }
            *reinterpret_cast<uint4*>(b + (c >> 3) * 128 + (c & 7) * 16) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        // the peer's warps signal the leader's barrier directly with a plain remote arrive (the form CUTLASS's ClusterBarrier
        // uses between CTAs): the writes above were made visible to the async proxy by fence.proxy.async in this thread
        // before the arrive is sent.  A cluster-scope release here cost ~1 us per stage — per warp, or in a relay thread —
        // and tripled the GEMM's time (profiles/r02_gemm3_trace.md).
        if (lane == 0) {
          if (leader) mbar_arrive(xf_bar(stage));
          else mbar_arrive_cluster_relaxed(xf_bar(stage), leader_cta);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (both CTAs, own 128 weight rows)
    const int q = warp & 3;
    const int row = q * 32 + lane;                    // TMEM lane = weight row inside the slab
    const int et = threadIdx.x - 64;                  // 0..127
    griddep_wait();
    if (et == 0) mark(4);
    int acc = 0;
    uint32_t acc_phase = 0, recv_phase = 0;
    float* xb = reinterpret_cast<float*>(smem + (xbuf - smem_base));
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(smem + (opst - smem_base));
    // fused-norm epilogue: bf16 [4 chunks][32 tokens][128 rows] of summed rows, in the (idle) ring behind the send staging
    __nv_bfloat16* zkeep = reinterpret_cast<__nv_bfloat16*>(smem + 3 * kChunkF32);
    float* inv_tok = reinterpret_cast<float*>(smem + 3 * kChunkF32 + kNumChunks * kOpStage);   // [4][32]
    const int valid_chunks = (n_eff + kChunkTok - 1) / kChunkTok;

    for (long long it = it_begin; it < it_end;) {
      const Seg sg = seg_at(it);
      const int slab = sg.tile * 2 + static_cast<int>(rank);   // global 128-row slab index of this CTA's rows
      const int n0 = slab * kSlab;
      const bool head_only = P.streamk && sg.kb0 == 0 && sg.kb1 < KB;   // the neighbour unit holds the tail
      const bool tail_only = P.streamk && sg.kb0 > 0;                   // publish for the neighbour

      // head of a shared tile: fetch the partner's partial (published long ago) under our own mainloop
      if (head_only) {
        if (et == 0) {
          const int* flag = P.flags + (unit + 1) * 2 + static_cast<int>(rank);
          while (ld_acquire_gpu(flag) != P.epoch) __nanosleep(32);
          asm volatile("fence.proxy.async;" ::: "memory");
          const uint32_t bytes = static_cast<uint32_t>(valid_chunks) * kChunkF32;
          mbar_arrive_expect_tx(recv_bar, bytes);
          const float* src = P.ws + (static_cast<size_t>(unit + 1) * 2 + rank) * (kBN * kSlab);
          for (int c = 0; c < valid_chunks; ++c) bulk_load_1d(xbuf + c * kChunkF32, src + static_cast<size_t>(c) * (kChunkTok * kSlab), kChunkF32, recv_bar);
        }
      } else if (!P.streamk && S > 1 && et == 0) {
        // cluster mode: arm the receive barrier for the chunks this CTA owns
        int owned = 0;
        for (int c = 0; c < valid_chunks; ++c) owned += (c % S) == static_cast<int>(pair);
        if (owned) mbar_arrive_expect_tx(recv_bar, static_cast<uint32_t>(owned) * (S - 1) * kChunkF32);
      }

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBN;

      if (tail_only) {
        // park this partial in L2 for the unit that owns the head of the tile (it reads it much later)
        float* dst = P.ws + (static_cast<size_t>(unit) * 2 + rank) * (kBN * kSlab);
        for (int c = 0; c < valid_chunks; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * kChunkTok, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[(c * kChunkTok + j) * kSlab + row] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(tempty_bar(acc));
          else mbar_arrive_cluster(tempty_bar(acc), leader_cta);
        }
        __threadfence();
        epi_bar();
        if (et == 0) st_release_gpu(P.flags + unit * 2 + static_cast<int>(rank), P.epoch);
      } else {
        // ---- cluster mode, phase 1: ship the chunks other pairs own (fp32, staged in the idle ring)
        if (!P.streamk && S > 1) {
          int sent = 0;
          for (int c = 0; c < valid_chunks; ++c) {
            const int owner = c % S;
            if (owner == static_cast<int>(pair)) continue;
            uint32_t v[32];
            tmem_ld_32x32(taddr + c * kChunkTok, v);
            tmem_ld_wait();
            float* sf = reinterpret_cast<float*>(smem + sent * kChunkF32);
#pragma unroll
            for (int j = 0; j < 32; ++j) sf[j * kSlab + row] = __uint_as_float(v[j]);
            fence_proxy_async();
            epi_bar();
            if (et == 0) {
              const uint32_t dst_cta = static_cast<uint32_t>(owner) * 2 + rank;
              const int oi = c / S;                                       // index among the owner's chunks
              const int si = static_cast<int>(pair) < owner ? static_cast<int>(pair) : static_cast<int>(pair) - 1;   // sender slot
              dsmem_copy(mapa_u32(xbuf + (oi * (S - 1) + si) * kChunkF32, dst_cta), smem_base + sent * kChunkF32, kChunkF32,
                         mapa_u32(recv_bar, dst_cta));
            }
            ++sent;
          }
        }
        const bool need_recv = head_only || (!P.streamk && S > 1);
        bool recv_waited = false;
        // ---- finish the chunks this CTA owns
        for (int c = 0; c < valid_chunks; ++c) {
          if (!P.streamk && S > 1 && (c % S) != static_cast<int>(pair)) continue;
          // operands of the fused op that do not depend on the accumulator: issue their loads first
          const int t_base = c * kChunkTok;
          uint4 res_pref[4];
          if (P.epi == GEMM3_EPI_RESADD) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int item = et + i * 128, j = item >> 4, vv = item & 15, t = t_base + j;
              res_pref[i] = (t < T) ? __ldcg(reinterpret_cast<const uint4*>(P.out + static_cast<size_t>(t) * P.ldo + n0 + vv * 8))
                                    : make_uint4(0, 0, 0, 0);
            }
          }
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * kChunkTok, v);
          tmem_ld_wait();
          if (need_recv && !recv_waited) {
            if (et == 0 && it + (sg.kb1 - sg.kb0) >= it_end) mark(5);
            mbar_wait(recv_bar, recv_phase);
            recv_phase ^= 1u;
            recv_waited = true;
            if (et == 0 && it + (sg.kb1 - sg.kb0) >= it_end) mark(6);
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (head_only) {
            const float* pp = xb + c * (kChunkTok * kSlab);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += pp[j * kSlab + row];
          } else if (!P.streamk && S > 1) {
            const int oi = c / S;
            for (int s = 0; s < S - 1; ++s) {
              const float* pp = xb + (oi * (S - 1) + s) * (kChunkTok * kSlab);
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] += pp[j * kSlab + row];
            }
          }
          // the rounding point of a bf16 GEMM output; staged [token][row] for the token-major pass
          epi_bar();   // previous chunk's pass is done with the staging tile
#pragma unroll
          for (int j = 0; j < 32; ++j) ob[j * kSlab + row] = __float2bfloat16_rn(f[j]);
          epi_bar();

          // ---------------- token-major pass: item = (token j, 16-byte vector vv of 8 rows)
          if (P.epi == GEMM3_EPI_PLAIN) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int item = et + i * 128, j = item >> 4, vv = item & 15, t = t_base + j;
              if (t < T) *reinterpret_cast<uint4*>(P.out + static_cast<size_t>(t) * P.ldo + n0 + vv * 8) = *reinterpret_cast<const uint4*>(ob + j * kSlab + vv * 8);
            }
          } else if (P.epi == GEMM3_EPI_RESADD) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int item = et + i * 128, j = item >> 4, vv = item & 15, t = t_base + j;
              V8 x, r, z;
              x.u = *reinterpret_cast<const uint4*>(ob + j * kSlab + vv * 8);
              r.u = res_pref[i];
              float ss = 0.f;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                z.h[e] = __float2bfloat16_rn(__bfloat162float(x.h[e]) + __bfloat162float(r.h[e]));
                const float zf = __bfloat162float(z.h[e]);
                ss += zf * zf;
              }
              // 16 consecutive lanes hold one token's 128 rows
              ss += __shfl_xor_sync(0xffffffffu, ss, 1);
              ss += __shfl_xor_sync(0xffffffffu, ss, 2);
              ss += __shfl_xor_sync(0xffffffffu, ss, 4);
              ss += __shfl_xor_sync(0xffffffffu, ss, 8);
              if (t < T) {
                *reinterpret_cast<uint4*>(P.out + static_cast<size_t>(t) * P.ldo + n0 + vv * 8) = z.u;
                if (vv == 0) P.ssq_out[static_cast<size_t>(t) * (N / kSlab) + slab] = ss;
              }
              // fused norm: the summed row stays in shared memory until every slab's sum of squares has been published
              if (P.normed_out) *reinterpret_cast<uint4*>(zkeep + c * (kChunkTok * kSlab) + j * kSlab + vv * 8) = z.u;
            }
          } else if (P.epi == GEMM3_EPI_SILU) {
            // rows 0..63 of the slab are gate rows, 64..127 the matching up rows; act columns [slab*64, slab*64+64)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int item = et + i * 128, j = item >> 3, vv = item & 7, t = t_base + j;
              V8 g, u, o;
              g.u = *reinterpret_cast<const uint4*>(ob + j * kSlab + vv * 8);
              u.u = *reinterpret_cast<const uint4*>(ob + j * kSlab + 64 + vv * 8);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float gv = __bfloat162float(g.h[e]);
                const __nv_bfloat16 s = silu_bf16(gv);
                o.h[e] = __float2bfloat16_rn(__bfloat162float(s) * __bfloat162float(u.h[e]));
              }
              if (t < T) *reinterpret_cast<uint4*>(P.out + static_cast<size_t>(t) * P.ldo + slab * 64 + vv * 8) = o.u;
            }
          } else if (P.epi == GEMM3_EPI_ROPE_KV) {
            const int head = slab;   // one 128-row slab = one head of the fused qkv projection
            if (head < P.Hq + P.Hkv) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int item = et + i * 128, j = item >> 3, vv = item & 7, t = t_base + j;
                if (t >= T) continue;
                int pos = __ldg(P.positions + t);
                pos = pos < 0 ? 0 : (pos >= P.max_pos ? P.max_pos - 1 : pos);
                const __nv_bfloat16* cs = P.cos_sin + static_cast<size_t>(pos) * 128;
                V8 x1, x2, co, si, o1, o2;
                x1.u = *reinterpret_cast<const uint4*>(ob + j * kSlab + vv * 8);
                x2.u = *reinterpret_cast<const uint4*>(ob + j * kSlab + 64 + vv * 8);
                co.u = __ldg(reinterpret_cast<const uint4*>(cs + vv * 8));
                si.u = __ldg(reinterpret_cast<const uint4*>(cs + 64 + vv * 8));
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float a = __bfloat162float(x1.h[e]), b = __bfloat162float(x2.h[e]);
                  const float cc = __bfloat162float(co.h[e]), sn = __bfloat162float(si.h[e]);
                  // every bf16 op rounds, as in the reference kernel (elementwise.cu rope_kv_kernel)
                  const float ac = bf16r(__fmul_rn(a, cc)), bs = bf16r(__fmul_rn(b, sn));
                  const float bc = bf16r(__fmul_rn(b, cc)), as = bf16r(__fmul_rn(a, sn));
                  o1.h[e] = __float2bfloat16_rn(ac - bs);
                  o2.h[e] = __float2bfloat16_rn(bc + as);
                }
                if (head < P.Hq) {
                  __nv_bfloat16* dst = P.out + static_cast<size_t>(t) * P.ldo + head * 128;
                  *reinterpret_cast<uint4*>(dst + vv * 8) = o1.u;
                  *reinterpret_cast<uint4*>(dst + 64 + vv * 8) = o2.u;
                } else {
                  const int slot = __ldg(P.slots + t);
                  if (slot >= 0) {
                    const size_t page_stride = static_cast<size_t>(P.Hkv) * 16 * 128;
                    __nv_bfloat16* dst = P.kv_layer + static_cast<size_t>(slot >> 4) * 2 * page_stride +
                                         (static_cast<size_t>(head - P.Hq) * 16 + (slot & 15)) * 128;
                    *reinterpret_cast<uint4*>(dst + vv * 8) = o1.u;
                    *reinterpret_cast<uint4*>(dst + 64 + vv * 8) = o2.u;
                  }
                }
              }
            } else {
              const int vh = head - P.Hq - P.Hkv;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int item = et + i * 128, j = item >> 4, vv = item & 15, t = t_base + j;
                if (t >= T) continue;
                const int slot = __ldg(P.slots + t);
                if (slot < 0) continue;
                const size_t page_stride = static_cast<size_t>(P.Hkv) * 16 * 128;
                __nv_bfloat16* dst = P.kv_layer + static_cast<size_t>(slot >> 4) * 2 * page_stride + page_stride +
                                     (static_cast<size_t>(vh) * 16 + (slot & 15)) * 128;
                *reinterpret_cast<uint4*>(dst + vv * 8) = *reinterpret_cast<const uint4*>(ob + j * kSlab + vv * 8);
              }
            }
          } else {  // GEMM3_EPI_ARGMAX
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int item = et + i * 128, j = item >> 4, vv = item & 15, t = t_base + j;
              V8 x;
              x.u = *reinterpret_cast<const uint4*>(ob + j * kSlab + vv * 8);
              float best = -INFINITY;
              int bi = 0x7fffffff;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float fv = __bfloat162float(x.h[e]);
                const int idx = n0 + vv * 8 + e;
                if (idx < P.n_valid && (fv > best || (fv == best && idx < bi))) {
                  best = fv;
                  bi = idx;
                }
              }
#pragma unroll
              for (int o = 1; o < 16; o <<= 1) {
                const float ob2 = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi2 = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob2 > best || (ob2 == best && oi2 < bi)) {
                  best = ob2;
                  bi = oi2;
                }
              }
              if (vv == 0 && t < T) P.cand[static_cast<size_t>(t) * (N / kSlab) + slab] = make_float2(best, __int_as_float(bi));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(tempty_bar(acc));
          else mbar_arrive_cluster(tempty_bar(acc), leader_cta);
        }
        if (P.epi == GEMM3_EPI_RESADD && P.normed_out) {
          // ---- fused RMSNorm of the summed residual (the input of the NEXT projection): every CTA has published the sums
          // of squares of its (chunk, slab); a row's norm needs all N/128 slabs, i.e. the other clusters — they finish
          // within a microsecond or two of each other (all clusters are co-resident in cluster mode), so each CTA marks its
          // slabs with the launch epoch and waits for the flags of the chunks it owns.  Deterministic: the per-slab sums are
          // added in slab order.  Rounding as the standalone kernel: bf16(bf16(z * inv) * w).
          const int slabs = N / kSlab;
          __threadfence();
          epi_bar();
          if (et == 0)
            for (int c = 0; c < valid_chunks; ++c)
              if (S == 1 || (c % S) == static_cast<int>(pair)) st_release_gpu(P.row_flags + c * slabs + slab, P.epoch);
          if (warp == 2) {
            for (int c = 0; c < valid_chunks; ++c) {
              if (S > 1 && (c % S) != static_cast<int>(pair)) continue;
              for (int s0 = 0; s0 < slabs; s0 += 32) {
                const int sidx = s0 + lane;
                while (!__all_sync(0xffffffffu, sidx >= slabs || ld_acquire_gpu(P.row_flags + c * slabs + sidx) == P.epoch)) __nanosleep(40);
              }
            }
          }
          epi_bar();
          // 1/rms per owned token: 4 threads per token, each adds a quarter of the slabs in order, then a fixed-order combine
          for (int c = 0; c < valid_chunks; ++c) {
            if (S > 1 && (c % S) != static_cast<int>(pair)) continue;
            const int j = et >> 2, part = et & 3, t = c * kChunkTok + j;
            float ss = 0.f;
            if (t < T) {
              const float* sp = P.ssq_out + static_cast<size_t>(t) * slabs;
              const int q4 = (slabs + 3) >> 2, a0 = part * q4, a1 = min(slabs, a0 + q4);
              for (int a = a0; a < a1; a += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (a + u < a1) ? __ldcg(sp + a + u) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) ss += v[u];
              }
            }
            const float s1 = __shfl_xor_sync(0xffffffffu, ss, 1);
            const float lo = (part & 1) ? s1 + ss : ss + s1;            // (part0 + part1) or (part2 + part3), same order in both lanes
            const float s2 = __shfl_xor_sync(0xffffffffu, lo, 2);
            const float tot = (part & 2) ? s2 + lo : lo + s2;
            if (part == 0) inv_tok[c * kChunkTok + j] = t < T ? rsqrtf(tot / static_cast<float>(N) + P.eps) : 0.f;
          }
          epi_bar();
          for (int c = 0; c < valid_chunks; ++c) {
            if (S > 1 && (c % S) != static_cast<int>(pair)) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int item = et + i * 128, j = item >> 4, vv = item & 15, t = c * kChunkTok + j;
              if (t >= T) continue;
              const uint4 zv = *reinterpret_cast<const uint4*>(zkeep + c * (kChunkTok * kSlab) + j * kSlab + vv * 8);
              const uint4 wv = __ldg(reinterpret_cast<const uint4*>(P.norm_w_out + n0 + vv * 8));
              const float inv = inv_tok[c * kChunkTok + j];
              const uint32_t zw[4] = {zv.x, zv.y, zv.z, zv.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
              uint32_t ow[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 nb = __floats2bfloat162_rn(__uint_as_float(zw[e] << 16) * inv, __uint_as_float(zw[e] & 0xffff0000u) * inv);
                const __nv_bfloat162 prod = __hmul2(nb, *reinterpret_cast<const __nv_bfloat162*>(&ww[e]));
                ow[e] = *reinterpret_cast<const uint32_t*>(&prod);
              }
              *reinterpret_cast<uint4*>(P.normed_out + static_cast<size_t>(t) * P.ldo + n0 + vv * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
          }
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
      it += sg.kb1 - sg.kb0;
    }
  }

  if (threadIdx.x == 64) mark(7);
  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while a peer can still signal its barriers or read / write its smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 2 * kBN);
  }
}

// out[s] = index of the best logit of row s over its per-slab candidates (lowest index wins ties, sampler.py:91)
__global__ void argmax_cand_kernel(const float2* __restrict__ cand, int* __restrict__ out, int slabs) {
  griddep_enter();
  const int s = blockIdx.x;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < slabs; i += blockDim.x) {
    const float2 c = __ldcg(cand + static_cast<size_t>(s) * slabs + i);
    const int idx = __float_as_int(c.y);
    if (c.x > best || (c.x == best && idx < bi)) {
      best = c.x;
      bi = idx;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) {
      best = ob;
      bi = oi;
    }
  }
  __shared__ float sb[32];
  __shared__ int si[32];
  if ((threadIdx.x & 31) == 0) {
    sb[threadIdx.x >> 5] = best;
    si[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    best = threadIdx.x < nw ? sb[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < nw ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    if (threadIdx.x == 0) out[s] = bi == 0x7fffffff ? 0 : bi;
  }
}

// co-resident clusters of each size with this kernel's footprint, per device (queried once)
struct Occupancy {
  int max_clusters[5];  // index S = pairs per cluster (1..4)
};
bool query_occupancy(int smem_bytes, Occupancy* o) {
  for (int S = 1; S <= 4; ++S) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * S * 64);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2 * S;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm3_kernel<kStagesDefault, kXbufDefault>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    o->max_clusters[S] = n;
  }
  return o->max_clusters[1] > 0;
}

}  // namespace

int gemm3_stages() {   // B200_GEMM3_STAGES=4|8: ring-depth experiment (8: no exchange buffer, plain S=1 cluster launches only)
  static const int v = [] {
    const char* e = getenv("B200_GEMM3_STAGES");
    const int n = e ? atoi(e) : kStagesDefault;
    return n == 4 || n == 8 ? n : kStagesDefault;
  }();
  return v;
}
int gemm3_smem_bytes(int K, int pro) {
  const int st = gemm3_stages();
  return smem_bytes_for(st, st == 8 ? 0 : kXbufDefault) + (pro == GEMM3_PRO_NORM ? K * 2 : 0);
}

int gemm3_schedule(int N, int K, int T, int pro, int force, Gemm3Schedule* out) {
  if (T < 1 || T > kBN || N % (2 * kSlab) != 0 || K % kBlockK != 0) return -1;
  static Occupancy occ[64];
  static bool have[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -2;
  dev &= 63;
  const int smem = gemm3_smem_bytes(4096, GEMM3_PRO_NORM);   // the largest footprint the engine launches (K = 4096)
  static std::atomic<unsigned long long> attr_done{0}, attr_done4{0}, attr_done8{0};
  constexpr int kMaxDynSmem = 232448;   // 227 KB: the per-CTA limit of sm_100
  if (!ensure_dynamic_smem(gemm3_kernel<kStagesDefault, kXbufDefault>, kMaxDynSmem, &attr_done) ||
      !ensure_dynamic_smem(gemm3_kernel<4, kXbufDefault>, kMaxDynSmem, &attr_done4) ||
      !ensure_dynamic_smem(gemm3_kernel<8, 0>, kMaxDynSmem, &attr_done8))
    return -3;
  if (!have[dev]) {
    if (!query_occupancy(smem, &occ[dev])) return -3;
    have[dev] = true;
  }
  if (gemm3_smem_bytes(K, pro) > kMaxDynSmem) return -1;
  const int tiles = N / (2 * kSlab), KB = K / kBlockK;
  const int pairs1 = occ[dev].max_clusters[1];   // co-resident pairs (74 on a 148-SM part)
  // stream-K needs every unit's range to be at least one tile long (a tile is then shared by at most two units)
  if ((force == 0 && tiles >= pairs1) || (force < 0 && tiles >= 2)) {
    out->streamk = 1;
    out->S = 1;
    out->units = tiles < pairs1 ? tiles : pairs1;
    if (force < -1 && -force <= out->units) out->units = -force;   // tests: a chosen number of units
    out->grid = 2 * out->units;
    return 0;
  }
  int S = 1;
  for (int s = 4; s >= 1; --s)
    if (occ[dev].max_clusters[s] >= tiles && s <= KB) {
      S = s;
      break;
    }
  if (force >= 1 && force <= 4) {
    if (force > KB || occ[dev].max_clusters[force] < tiles) return -4;
    S = force;
  }
  if (occ[dev].max_clusters[S] < tiles) return -4;   // cluster mode needs every cluster co-resident (DSMEM exchange)
  out->streamk = 0;
  out->S = S;
  out->units = tiles * S;
  out->grid = tiles * 2 * S;
  return 0;
}

int gemm3_launch(const Gemm3Params& p, const Gemm3Schedule& sch, cudaStream_t st) {
  const int smem = gemm3_smem_bytes(p.K, p.pro);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sch.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * sch.S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  Gemm3Params q = p;
  q.S = sch.S;
  q.streamk = sch.streamk;
  const int st_n = gemm3_stages();
  if (st_n == 8 && (q.streamk || q.S > 1)) return -6;   // the 8-stage variant has no exchange buffer
  cudaError_t e = st_n == 4 ? cudaLaunchKernelEx(&cfg, gemm3_kernel<4, kXbufDefault>, q)
                : st_n == 8 ? cudaLaunchKernelEx(&cfg, gemm3_kernel<8, 0>, q)
                            : cudaLaunchKernelEx(&cfg, gemm3_kernel<kStagesDefault, kXbufDefault>, q);
  return e == cudaSuccess ? 0 : -4;
}

int argmax_candidates(const void* cand, int* out, int S, int slabs, cudaStream_t st) {
  if (S <= 0) return 0;
  launch_pdl(argmax_cand_kernel, dim3(S), dim3(256), 0, st, static_cast<const float2*>(cand), out, slabs);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace b200
