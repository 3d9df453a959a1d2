// Decode-step "projection chain": o_proj -> (residual add + RMSNorm) -> gate_up (+SiLU) -> down_proj -> (residual add +
// RMSNorm) -> next layer's qkv_proj -> (RoPE + KV write) as ONE persistent kernel, T <= 128 tokens.
//
// Why.  Round 1 ran each projection and each elementwise op as its own launch: ~11 us per GEMM -> elementwise -> GEMM
// boundary, a decode step at 0.57 of the HBM roofline.  Fusing the elementwise ops into per-GEMM launches (gemm3_tcgen05.cu)
// removed the elementwise kernels but not the per-launch cost: every GEMM launch pays ~4 us before its first MMA (CTA start,
// barrier / TMEM set-up, the first HBM round trip) and ~3-5 us after its last one, the weight stream stops in between, and
// the projections with few 256-row tiles (N = 4096, 6144) could only use the CTAs of co-resident clusters — 96 of 148 SMs at
// the measured ~49 GB/s per SM cap, 4.6 TB/s (profiles/r02_gemm3_trace.md).  Measured: no faster than the unfused path.
//
// Here the four projections between two attention calls share one launch of 148 CTAs (74 CTA pairs, one per SM pair):
//   * every projection is a stream-K phase over all 74 pairs (tcgen05.mma.cta_group::2, M = 256, TMA 128B-swizzle ring, TMEM
//     accumulators, as gemm2 / gemm3), so all SMs stream weights in every phase;
//   * the TMA producer runs AHEAD across phase boundaries: weights are static, so while the pairs drain a phase, exchange
//     partials and wait at a grid barrier, the ring already fills with the next projection's weights — HBM keeps streaming;
//     only the token-tile loads of the next phase wait for the barrier;
//   * split-K: o / down / qkv tiles are shared by several pairs (fewer tiles than pairs): fp32 segments go to an L2
//     workspace in the gemm2 layout, a software grid barrier follows, and the elementwise phase (one token row per CTA) sums
//     the segments while it does its own work — exactly the data flow of the unfused path, minus six kernel boundaries;
//     gate_up has more tiles than pairs: tails are parked early in L2 and the head owner finishes the tile (gemm3's scheme),
//     SiLU*up runs in the epilogue, no extra barrier;
//   * grid barrier = one 64-bit counter in HBM, one atomic arrival per CTA, acquire polling; targets are base + k * CTAs.
// Rounding points are those of the unfused kernels (elementwise.cu) and of the reference backend:
//   vllm/model_executor/models/llama.py:81-121,223-233,316-340, _custom_ops.py:323-327, activation.py:138-148,
//   rotary_embedding/base.py:140-198.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "chain.h"
#include "gemm.h"
#include "launch.h"
#include "partials.cuh"
#include "ptx.cuh"
#include "umma2.cuh"

namespace b200 {

namespace {

constexpr int kSlab = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kBN = 128;
constexpr int kChunkTok = 32;
constexpr int kThreads = 192;                          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue / elementwise
constexpr int kABytes = kSlab * kBlockK * 2;           // 16 KB
constexpr int kBBytes = (kBN / 2) * kBlockK * 2;       // 8 KB
constexpr int kStageBytes = kABytes + kBBytes;         // 24 KB
constexpr int kStages = 6;
constexpr int kRing = kStages * kStageBytes;           // 144 KB
constexpr int kChunkF32 = kChunkTok * kSlab * 4;       // 16 KB
constexpr int kXbuf = 4 * kChunkF32;                   // 64 KB: gate_up partner partial
constexpr int kOpStage = kChunkTok * kSlab * 2;        // 8 KB
constexpr int kMisc = 1024;
constexpr int kSmemBytes = 1024 + kRing + kXbuf + kOpStage + kMisc;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int unit_of_iter(long long x, long long total, int units) {
  return static_cast<int>(((x + 1) * units + total - 1) / total - 1);
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

union V8 {
  uint4 u;
  __nv_bfloat16 h[8];
};
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct Seg {
  int tile, kb0, kb1;
};

// iteration range of `unit` in GEMM phase g (empty when the unit does not take part)
__device__ __forceinline__ void phase_range(const ChainGemm& g, int unit, long long* b, long long* e) {
  const long long total = static_cast<long long>(g.N / (2 * kSlab)) * (g.K / kBlockK);
  if (unit >= g.units) {
    *b = *e = 0;
    return;
  }
  *b = range_begin(unit, total, g.units);
  *e = range_begin(unit + 1, total, g.units);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) chain_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t xbuf = smem_base + kRing;
  const uint32_t opst = xbuf + kXbuf;
  const uint32_t misc = opst + kOpStage;
  auto full_bar = [&](int s) { return misc + 8u * s; };
  auto empty_bar = [&](int s) { return misc + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return misc + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return misc + 8u * (2 * kStages + 2 + a); };
  const uint32_t recv_bar = misc + 8u * (2 * kStages + 4);
  const uint32_t tmem_slot = misc + 8u * (2 * kStages + 5);
  float* red = reinterpret_cast<float*>(smem + (misc - smem_base) + 512);   // [8] block-reduction scratch
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  // optional phase trace (debug): 32 %globaltimer stamps per CTA
  //   [4p+0] epilogue warps done with phase p's segments  [4p+1] done_barrier passed  [4p+2] elementwise phase done
  //   [16+2p] first MMA of phase p  [17+2p] last MMA of phase p issued  [24+p] first token-tile load of phase p issued
  //   [30] CTA start  [31] CTA end
  auto mark = [&](int i) {
    if (P.trace) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      P.trace[static_cast<size_t>(blockIdx.x) * 32 + i] = t;
    }
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 64) mark(30);
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int unit = blockIdx.x >> 1;
  const int T = P.T;
  const int nc = (T + 15) & ~15;
  const int valid_chunks = (nc + kChunkTok - 1) / kChunkTok;
  const int row_half0 = static_cast<int>(rank) * (nc >> 1);
  const unsigned long long nctas = gridDim.x;
  auto bar_target = [&](int k) { return P.bar_base + static_cast<unsigned long long>(k) * nctas; };

  if (warp == 0 && lane == 0)
    for (int p = 0; p < P.n_gemm; ++p) {
      tma_prefetch_desc(&P.g[p].tm_w);
      tma_prefetch_desc(&P.g[p].tm_x);
    }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
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
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: weights run ahead across phases, token
    // tiles of a phase wait for the grid barrier that publishes them
    if (lane == 0) {
      // Two cursors walk this pair's iterations over all phases: `ca` issues weight loads (never waits for anything but a
      // free ring slot), `cb` issues token-tile loads (waits for the phase's grid barrier).  Pure spinning: a __nanosleep in
      // this loop let the ring run dry (measured: every mainloop 2x slower, profiles/r02_chain_trace.md).
      struct Cur {
        int ph, tile, kb, KB;
        long long left;      // iterations left in the current phase
        long long idx;       // flattened index
      };
      auto enter = [&](Cur& c) {   // position the cursor at the first iteration of the first non-empty phase >= c.ph
        while (c.ph < P.n_gemm) {
          long long b0, e0;
          phase_range(P.g[c.ph], unit, &b0, &e0);
          if (e0 > b0) {
            c.KB = P.g[c.ph].K / kBlockK;
            c.tile = static_cast<int>(b0 / c.KB);
            c.kb = static_cast<int>(b0 - static_cast<long long>(c.tile) * c.KB);
            c.left = e0 - b0;
            return;
          }
          ++c.ph;
        }
      };
      auto advance = [&](Cur& c) {
        ++c.idx;
        if (--c.left == 0) {
          ++c.ph;
          enter(c);
        } else if (++c.kb == c.KB) {
          c.kb = 0;
          ++c.tile;
        }
      };
      Cur ca{0, 0, 0, 1, 0, 0}, cb{0, 0, 0, 1, 0, 0}, cp{0, 0, 0, 1, 0, 0};
      enter(ca);
      enter(cb);
      enter(cp);
      // While the token tiles are blocked (upstream kernel still running, or a grid barrier not yet passed) and the ring is
      // full, the weight stream would stop: instead the next kPrefetch weight boxes beyond the ring are pulled into L2, so the
      // ring refills from L2 once the phase runs.  Only in those gaps: in the steady mainloop HBM is busy anyway.
      const int kPrefetch = P.prefetch;
      bool waited_dep = false;
      int passed = 0;                 // highest grid barrier known to have been passed
      while (cb.ph < P.n_gemm) {
        if (ca.ph < P.n_gemm && ca.idx < cb.idx + kStages) {
          const int stage = static_cast<int>(ca.idx % kStages);
          const uint32_t use = static_cast<uint32_t>(ca.idx / kStages);
          if (use == 0 || mbar_try_wait(empty_bar(stage), (use & 1u) ^ 1u)) {
            if (leader) mbar_arrive_expect_tx(full_bar(stage), 2u * kStageBytes);
            tma_load_2d_pair(smem_base + stage * kStageBytes, &P.g[ca.ph].tm_w, full_bar(stage), ca.kb * kBlockK,
                             ca.tile * 2 * kSlab + static_cast<int>(rank) * kSlab, kEvictFirst);
            advance(ca);
          }
        }
        bool blocked = false;
        if (cb.idx < ca.idx) {
          if (!waited_dep) {          // first token tile of the launch: the upstream kernel (attention) must be complete
            // (griddepcontrol.wait blocks: pull this pair's first weights into L2 before sitting in it)
            while (cp.ph < P.n_gemm && cp.idx < ca.idx) advance(cp);
            for (int i = 0; i < kPrefetch && cp.ph < P.n_gemm; ++i) {
              tma_prefetch_l2_2d(&P.g[cp.ph].tm_w, cp.kb * kBlockK, cp.tile * 2 * kSlab + static_cast<int>(rank) * kSlab);
              advance(cp);
            }
            griddep_wait();
            waited_dep = true;
          }
          const int need = P.g[cb.ph].wait_barrier;
          if (need > passed && ld_acquire_u64(P.bar) >= bar_target(need)) passed = need;
          blocked = need > passed;
          if (need <= passed) {
            const int stage = static_cast<int>(cb.idx % kStages);
            tma_load_2d_pair(smem_base + stage * kStageBytes + kABytes, &P.g[cb.ph].tm_x, full_bar(stage), cb.kb * kBlockK, row_half0, kEvictLast);
            advance(cb);
          }
        }
        if (blocked && ca.idx >= cb.idx + kStages) {   // ring full of the next phase's weights, token tiles not published yet
          while (cp.ph < P.n_gemm && cp.idx < ca.idx) advance(cp);
          if (cp.ph < P.n_gemm && cp.idx < ca.idx + kPrefetch) {
            tma_prefetch_l2_2d(&P.g[cp.ph].tm_w, cp.kb * kBlockK, cp.tile * 2 * kSlab + static_cast<int>(rank) * kSlab);
            advance(cp);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && leader) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t idesc = umma_idesc_bf16(2 * kSlab, nc);
      for (int p = 0; p < P.n_gemm; ++p) {
        long long it_begin, it_end;
        phase_range(P.g[p], unit, &it_begin, &it_end);
        const int KB = P.g[p].K / kBlockK;
        for (long long it = it_begin; it < it_end;) {
          Seg sg;
          sg.tile = static_cast<int>(it / KB);
          sg.kb0 = static_cast<int>(it - static_cast<long long>(sg.tile) * KB);
          const long long rem = it_end - it;
          sg.kb1 = (KB - sg.kb0 <= rem) ? KB : sg.kb0 + static_cast<int>(rem);
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            if (it == it_begin && kb == sg.kb0) mark(16 + 2 * p);
            const uint32_t sa = smem_base + stage * kStageBytes;
            const uint64_t a_desc = umma_desc_kmajor_sw128(sa);
            const uint64_t b_desc = umma_desc_kmajor_sw128(sa + kABytes);
            const uint32_t d_tmem = tmem_base + acc * kBN;
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma2_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb > sg.kb0 || k > 0) ? 1u : 0u);
            umma2_commit_pair(empty_bar(stage), 3);
            if (kb == sg.kb1 - 1) umma2_commit_pair(tfull_bar(acc), 3);
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
        mark(17 + 2 * p);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue / elementwise warps
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    griddep_wait();
    int acc = 0;
    uint32_t acc_phase = 0, recv_phase = 0;
    float* xb = reinterpret_cast<float*>(smem + (xbuf - smem_base));
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(smem + (opst - smem_base));

    // One arrival per CTA per barrier, after all of this CTA's global writes of the phase.  The counter is cumulative, so a
    // CTA may only arrive at barrier k once barrier k-1 has passed — otherwise a CTA with no work in a phase could run ahead
    // and its extra arrivals would stand in for a CTA that has not arrived yet.
    auto grid_arrive = [&](int k) {
      __threadfence();
      epi_bar();
      if (et == 0) {
        if (k > 1) {
          const unsigned long long prev = bar_target(k - 1);
          while (ld_acquire_u64(P.bar) < prev) {
          }
        }
        atomicAdd(P.bar, 1ULL);
      }
    };
    auto grid_wait = [&](int k) {
      if (et == 0) {
        const unsigned long long tgt = bar_target(k);
        while (ld_acquire_u64(P.bar) < tgt) {
        }
      }
      epi_bar();
    };
    auto block_sum = [&](float v) {   // sum over the 128 elementwise threads, same value in every thread
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      epi_bar();                      // scratch free
      if (lane == 0) red[q] = v;
      epi_bar();
      return (red[0] + red[1]) + (red[2] + red[3]);
    };

    for (int p = 0; p < P.n_gemm; ++p) {
      const ChainGemm& g = P.g[p];
      const int KB = g.K / kBlockK;
      long long it_begin, it_end;
      phase_range(g, unit, &it_begin, &it_end);
      const long long total = static_cast<long long>(g.N / (2 * kSlab)) * KB;

      for (long long it = it_begin; it < it_end;) {
        Seg sg;
        sg.tile = static_cast<int>(it / KB);
        sg.kb0 = static_cast<int>(it - static_cast<long long>(sg.tile) * KB);
        const long long rem = it_end - it;
        sg.kb1 = (KB - sg.kb0 <= rem) ? KB : sg.kb0 + static_cast<int>(rem);
        const int slab = sg.tile * 2 + static_cast<int>(rank);
        const int n0 = slab * kSlab;
        const bool complete = sg.kb0 == 0 && sg.kb1 == KB;
        const bool silu = g.mode == CHAIN_SILU;
        const bool head_only = silu && sg.kb0 == 0 && sg.kb1 < KB;
        const bool tail_only = silu && sg.kb0 > 0;

        if (head_only && et == 0 && !(P.dbg & 1)) {   // partner's tail of this tile was parked long ago: fetch it under our mainloop
          const int* flag = P.flags + (unit + 1) * 2 + static_cast<int>(rank);
          while (ld_acquire_gpu(flag) != P.epoch) __nanosleep(32);
          asm volatile("fence.proxy.async;" ::: "memory");
          mbar_arrive_expect_tx(recv_bar, static_cast<uint32_t>(valid_chunks) * kChunkF32);
          const float* src = P.ws_sk + (static_cast<size_t>(unit + 1) * 2 + rank) * (kBN * kSlab);
          for (int c = 0; c < valid_chunks; ++c)
            bulk_load_1d(xbuf + c * kChunkF32, src + static_cast<size_t>(c) * (kChunkTok * kSlab), kChunkF32, recv_bar);
        }

        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBN;

        if (P.dbg & 1) {   // timing experiment: accumulators are released undrained (results are garbage)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(tempty_bar(acc));
            else mbar_arrive_cluster(tempty_bar(acc), 0);
          }
          if (tail_only) {
            epi_bar();
            if (et == 0) st_release_gpu(P.flags + unit * 2 + static_cast<int>(rank), P.epoch);
          }
        } else if (tail_only || (!silu && !complete)) {
          // fp32 partial [token][128 rows] to L2: the stream-K neighbour (gate_up) or the elementwise phase sums it
          float* dst;
          if (silu) {
            dst = P.ws_sk + (static_cast<size_t>(unit) * 2 + rank) * (kBN * kSlab);
          } else {
            const int u0 = unit_of_iter(static_cast<long long>(sg.tile) * KB, total, g.units);
            const int seg = __ldg(&g.seg_table[sg.tile]).x + (unit - u0);
            dst = P.ws_def + (static_cast<size_t>(seg) * 2 + rank) * (kBN * kSlab);
          }
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
            else mbar_arrive_cluster(tempty_bar(acc), 0);
          }
          if (silu) {
            __threadfence();
            epi_bar();
            if (et == 0) st_release_gpu(P.flags + unit * 2 + static_cast<int>(rank), P.epoch);
          }
        } else {
          bool recv_waited = false;
          for (int c = 0; c < valid_chunks; ++c) {
            const int t_base = c * kChunkTok;
            uint32_t v[32];
            tmem_ld_32x32(taddr + c * kChunkTok, v);
            tmem_ld_wait();
            if (head_only && !recv_waited) {
              mbar_wait(recv_bar, recv_phase);
              recv_phase ^= 1u;
              recv_waited = true;
            }
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (head_only) {
              const float* pp = xb + c * (kChunkTok * kSlab);
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] += pp[j * kSlab + row];
            }
            epi_bar();
#pragma unroll
            for (int j = 0; j < 32; ++j) ob[j * kSlab + row] = __float2bfloat16_rn(f[j]);
            epi_bar();
            if (!silu) {
              // complete tile of a deferred projection: bf16 to the dense tensor (the elementwise phase reads it there)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int item = et + i * 128, j = item >> 4, vv = item & 15, t = t_base + j;
                if (t < T) *reinterpret_cast<uint4*>(g.out + static_cast<size_t>(t) * g.ldo + n0 + vv * 8) = *reinterpret_cast<const uint4*>(ob + j * kSlab + vv * 8);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int item = et + i * 128, j = item >> 3, vv = item & 7, t = t_base + j;
                V8 gg, uu, o;
                gg.u = *reinterpret_cast<const uint4*>(ob + j * kSlab + vv * 8);
                uu.u = *reinterpret_cast<const uint4*>(ob + j * kSlab + 64 + vv * 8);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float gv = __bfloat162float(gg.h[e]);
                  const __nv_bfloat16 s = silu_bf16(gv);
                  o.h[e] = __float2bfloat16_rn(__bfloat162float(s) * __bfloat162float(uu.h[e]));
                }
                if (t < T) *reinterpret_cast<uint4*>(g.out + static_cast<size_t>(t) * g.ldo + slab * 64 + vv * 8) = o.u;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(tempty_bar(acc));
            else mbar_arrive_cluster(tempty_bar(acc), 0);
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
        it += sg.kb1 - sg.kb0;
      }

      // ---- this CTA's part of projection p is out
      if (et == 0) mark(4 * p);
      if (g.done_barrier > 0) grid_arrive(g.done_barrier);
      if (g.reduce == CHAIN_REDUCE_NONE) continue;

      // ---- elementwise phase: token row blockIdx.x (T <= 128 <= CTAs), the projection's split tiles summed on load.
      // Everything that does not come from the projection — segment-table entries (static), the residual row (final since
      // the previous elementwise phase), the norm weight — is loaded BEFORE the barrier wait.
      PartialView pv;
      pv.ws = P.ws_def;
      pv.table = g.seg_table;
      pv.dense = g.out;
      pv.ld_dense = g.ldo;
      pv.slot = kBN * kSlab;
      pv.ntt = 1;
      pv.block_n = kBN;
      pv.bn_shift = 7;
      const int t = blockIdx.x;
      if (g.reduce == CHAIN_REDUCE_RESADD_NORM) {
        const int H = g.N;
        const __nv_bfloat16* nw = g.norm_w;
        float ss = 0.f;
        constexpr int kMaxV = 8;                 // H <= 8192
        uint4 zv[kMaxV];
        const bool fast = t < T && H == 4096;
        int nn[4] = {0, 0, 0, 0};
        int2 ent[4];
        uint4 rr[4], wpre[4];
        if (fast) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            nn[i] = (et + i * 128) * 8;
            rr[i] = *reinterpret_cast<const uint4*>(P.res + static_cast<size_t>(t) * H + nn[i]);
            if (nw) wpre[i] = __ldg(reinterpret_cast<const uint4*>(nw) + et + i * 128);
          }
          partial_entries<4>(pv, t, nn, ent);
        }
        grid_wait(g.done_barrier);
        if (et == 0) mark(4 * p + 1);
        if (fast) {
          // the common shape: all four vectors' segment loads issued before the first add (one L2 round trip)
          float xa[4][8];
          load8xM_entries<4>(pv, ent, t, nn, xa);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            V8 r, z;
            r.u = rr[i];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              z.h[e] = __float2bfloat16_rn(xa[i][e] + __bfloat162float(r.h[e]));
              const float zf = __bfloat162float(z.h[e]);
              ss += zf * zf;
            }
            *reinterpret_cast<uint4*>(P.res + static_cast<size_t>(t) * H + nn[i]) = z.u;
            zv[i] = z.u;
          }
        } else if (t < T) {
#pragma unroll
          for (int i = 0; i < kMaxV; ++i) {
            const int idx = et + i * 128;
            if (idx * 8 < H) {
              float xa[8];
              load8_partials(pv, t, idx * 8, xa);
              V8 r, z;
              r.u = *reinterpret_cast<const uint4*>(P.res + static_cast<size_t>(t) * H + idx * 8);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                z.h[e] = __float2bfloat16_rn(xa[e] + __bfloat162float(r.h[e]));
                const float zf = __bfloat162float(z.h[e]);
                ss += zf * zf;
              }
              *reinterpret_cast<uint4*>(P.res + static_cast<size_t>(t) * H + idx * 8) = z.u;
              zv[i] = z.u;
            }
          }
        }
        if (nw) {
          const float tot = block_sum(ss);
          const float inv = rsqrtf(tot / static_cast<float>(H) + P.eps);
          if (t < T) {
#pragma unroll
            for (int i = 0; i < kMaxV; ++i) {
              const int idx = et + i * 128;
              if (idx * 8 < H) {
                const uint4 wv = fast ? wpre[i & 3] : __ldg(reinterpret_cast<const uint4*>(nw) + idx);
                const uint32_t zw[4] = {zv[i].x, zv[i].y, zv[i].z, zv[i].w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
                uint32_t ow[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const __nv_bfloat162 nb = __floats2bfloat162_rn(__uint_as_float(zw[e] << 16) * inv, __uint_as_float(zw[e] & 0xffff0000u) * inv);
                  const __nv_bfloat162 prod = __hmul2(nb, *reinterpret_cast<const __nv_bfloat162*>(&ww[e]));
                  ow[e] = *reinterpret_cast<const uint32_t*>(&prod);
                }
                *reinterpret_cast<uint4*>(P.normed + static_cast<size_t>(t) * H + idx * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
              }
            }
          }
        }
      } else if (g.reduce == CHAIN_REDUCE_ROPE_KV) {
        grid_wait(g.done_barrier);
        if (et == 0) mark(4 * p + 1);
        if (t < T) {
        // neox RoPE on the q / k heads of token t, q back into the qkv buffer, k / v into the paged cache
        // (elementwise.cu rope_kv_kernel, 128 threads)
        constexpr int D = 128, HALF = 64;
        const int Hq = P.Hq, Hkv = P.Hkv;
        int pos = __ldg(P.positions + t);
        pos = pos < 0 ? 0 : (pos >= P.max_pos ? P.max_pos - 1 : pos);
        const int slot = __ldg(P.slots + t);
        __nv_bfloat16* rowp = g.out + static_cast<size_t>(t) * g.ldo;
        const __nv_bfloat16* cs = P.cos_sin + static_cast<size_t>(pos) * D;
        const size_t page_stride = static_cast<size_t>(Hkv) * 16 * D;
        __nv_bfloat16* kbase = P.kv_layer + static_cast<size_t>(slot >> 4) * 2 * page_stride;
        __nv_bfloat16* vbase = kbase + page_stride;
        const int off = slot & 15;
        const int rot_tasks = (Hq + Hkv) * (HALF / 8), v_tasks = Hkv * (D / 8);
        for (int task = et; task < rot_tasks; task += 128) {
          const int head = task >> 3, c = task & 7;
          const int nn[2] = {head * D + c * 8, head * D + HALF + c * 8};
          float xx[2][8];
          load8xM_partials<2>(pv, t, nn, xx);
          V8 co, si, o1, o2;
          co.u = __ldg(reinterpret_cast<const uint4*>(cs + c * 8));
          si.u = __ldg(reinterpret_cast<const uint4*>(cs + HALF + c * 8));
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float a = xx[0][e], b = xx[1][e];
            const float cc = __bfloat162float(co.h[e]), sn = __bfloat162float(si.h[e]);
            const float ac = bf16r(__fmul_rn(a, cc)), bs = bf16r(__fmul_rn(b, sn));
            const float bc = bf16r(__fmul_rn(b, cc)), as = bf16r(__fmul_rn(a, sn));
            o1.h[e] = __float2bfloat16_rn(ac - bs);
            o2.h[e] = __float2bfloat16_rn(bc + as);
          }
          if (head < Hq) {
            *reinterpret_cast<uint4*>(rowp + head * D + c * 8) = o1.u;
            *reinterpret_cast<uint4*>(rowp + head * D + HALF + c * 8) = o2.u;
          } else if (slot >= 0) {
            __nv_bfloat16* dst = kbase + (static_cast<size_t>(head - Hq) * 16 + off) * D;
            *reinterpret_cast<uint4*>(dst + c * 8) = o1.u;
            *reinterpret_cast<uint4*>(dst + HALF + c * 8) = o2.u;
          }
        }
        if (slot >= 0) {
          for (int task = et; task < v_tasks; task += 128) {
            const int head = task >> 4, c = task & 15;
            float fv[8];
            load8_partials(pv, t, (Hq + Hkv + head) * D + c * 8, fv);
            uint4 val;
            val.x = pack_bf16x2(fv[0], fv[1]);
            val.y = pack_bf16x2(fv[2], fv[3]);
            val.z = pack_bf16x2(fv[4], fv[5]);
            val.w = pack_bf16x2(fv[6], fv[7]);
            *reinterpret_cast<uint4*>(vbase + (static_cast<size_t>(head) * 16 + off) * D + c * 8) = val;
          }
        }
        }
      }
      if (et == 0) mark(4 * p + 2);
      if (g.reduce_barrier > 0) grid_arrive(g.reduce_barrier);
    }
  }
  if (threadIdx.x == 64) mark(31);

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 2 * kBN);
  }
}

}  // namespace

static long long* g_chain_trace = nullptr;
static int g_chain_trace_launch = 0;
void chain_set_trace(long long* dev) {   // debug: 40 slots of CTAs x 32 stamps, launch i writes slot i % 40
  g_chain_trace = dev;
  g_chain_trace_launch = 0;
}

int chain_smem_bytes() { return kSmemBytes; }

int chain_max_ctas(int* out) {
  static std::atomic<unsigned long long> attr_done{0};
  if (!ensure_dynamic_smem(chain_kernel, kSmemBytes, &attr_done)) return -3;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 64);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, chain_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return -3;
  }
  *out = 2 * n;
  return 0;
}

// grid: every launched CTA must be resident at the same time (software grid barrier): callers pass chain_max_ctas()
int chain_launch(const ChainParams& p, int ctas, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_done{0};
  if (!ensure_dynamic_smem(chain_kernel, kSmemBytes, &attr_done)) return -3;
  if (p.T < 1 || p.T > kBN || p.T > ctas) return -1;
  ChainParams q = p;
  q.trace = g_chain_trace ? g_chain_trace + static_cast<size_t>(g_chain_trace_launch++ % 40) * 148 * 32 : nullptr;
  return launch_pdl(chain_kernel, dim3(ctas), dim3(kThreads), kSmemBytes, st, q) == cudaSuccess ? 0 : -4;
}

}  // namespace b200
