// Projection GEMM, CTA-pair version (tcgen05 cta_group::2) — the default for every step.
//     out[t, n] = sum_k X[t, k] * W[n, k]        (bf16 x bf16 -> fp32 -> bf16)
// Same contract and stream-K schedule as gemm_tcgen05.cu (see there for the reference citations);
// what changes is the unit of work, because the single-CTA kernel measured L2-bound, not HBM-bound:
// every 128-row weight slab re-read the whole activation tile from L2 (1 byte of X per byte of W at
// T=128, 2:1 and W twice at T=384) and the L2 output port (~11 TB/s measured) saturated first.
//   * Two SMs (a cluster of 2) issue ONE tcgen05.mma.cta_group::2 with M=256: each CTA stages its own
//     128 weight rows and only HALF of the token tile; the tensor core reads both halves across the
//     pair.  Activation traffic per weight byte is halved and per-CTA smem per stage shrinks.
//   * The token tile is up to 512 wide (two N<=256 instructions per k-step into two TMEM regions), so
//     a mixed decode+prefill step of <=512 tokens still streams every weight byte exactly once.
//   * stream-K over CTA pairs (74 ranges); the split-tile fix-up, the fp32 L2 workspace, the
//     cp.async.bulk pull, PDL with early weight prefetch are as in the single-CTA kernel, per CTA.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "epi_pass.cuh"
#include "gemm.h"
#include "launch.h"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kSlab = 128;   // weight rows per CTA (UMMA M = 256 per pair)
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;
constexpr int kABytes = kSlab * kBlockK * 2;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> rank 0's copy

template <int BLOCK_N>
struct Cfg2 {
  static constexpr int kChunk = BLOCK_N > 256 ? 256 : BLOCK_N;  // tokens per MMA instruction
  static constexpr int kNch = (BLOCK_N + kChunk - 1) / kChunk;  // instructions per k-step (1 or 2; 384 = 256 + 128)
  static constexpr int kHalf = kChunk / 2;                      // token rows of one chunk staged by one CTA
  static constexpr int kBBytes = kNch * kHalf * kBlockK * 2;    // B bytes per CTA per stage
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kAccStages = (2 * BLOCK_N <= 512) ? 2 : 1;
  static constexpr int kTmemNeed = kAccStages * BLOCK_N < 32 ? 32 : kAccStages * BLOCK_N;
  static constexpr int kTmemCols = kTmemNeed <= 32 ? 32 : kTmemNeed <= 64 ? 64 : kTmemNeed <= 128 ? 128 : kTmemNeed <= 256 ? 256 : 512;  // power of two
  static constexpr int kStoreBuf = 32 * kSlab * 4;        // one staging buffer: [32 tokens][128 rows], fp32 or bf16
  static constexpr int kStoreStage = 2 * kStoreBuf;       // double buffered
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kStoreStage;
};

struct Seg {
  int tile, kb0, kb1;
};

// whole-tile waves ahead of the stream-K tail in fused mode (see the schedule comment in the kernel); host and device
__host__ __device__ __forceinline__ int hybrid_dp_waves(int tiles_total, int units, int KB) {
  int w = tiles_total >= units ? tiles_total / units : 0;
  // a short tail would cut each of its few tiles into many pieces for one finisher to add (4 tiles over 74 pairs: 18 pieces
  // each, +120 us at T = 2048): when less than half a wave is left, one whole wave joins the stream-K part
  if (w > 0 && (tiles_total - w * units) * 2 < units && tiles_total != w * units) --w;
  // every pair must own at least one iteration of the tail: a pair with an empty range publishes nothing, and the finisher of
  // a tile counts on a piece from every pair between the owners of the tile's first and last k-block
  const long long tail = static_cast<long long>(tiles_total - w * units) * KB;
  if (tail != 0 && tail < units) w = 0;
  return w;
}
__device__ __forceinline__ long long range_begin(int unit, long long total, int units) {
  return (static_cast<long long>(unit) * total) / units;
}
__device__ __forceinline__ int unit_of_iter(long long x, long long total, int units) {
  return static_cast<int>(((x + 1) * units + total - 1) / total - 1);
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-D TMA load into OWN smem, completion bytes credited to the LEADER CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst_smem, const void* tmap, uint32_t bar, int32_t c0,
                                                 int32_t c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & kPeerMask), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, 256 x N over the pair] (+)= A[256 x 16] * B[N x 16]; issued by the leader CTA only.
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// Completion of all prior MMAs of this thread arrives on the same-offset mbarrier of BOTH CTAs.
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(bar),
      "r"(cta)
      : "memory");
}

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_streamk_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                     const __grid_constant__ CUtensorMap tm_out, int tma_store, __nv_bfloat16* __restrict__ out, int ldo, float* __restrict__ ws, int* __restrict__ counters,
                     int N, int T, int K, const int2* __restrict__ seg_table, int mode, const __grid_constant__ Gemm2Epi E,
                     long long* __restrict__ trace) {
  // mode 0: in-kernel fix-up through counters; 1: deferred (split tiles stay fp32 segments for the consumer kernel);
  // 2: fused — split tiles are finished in-kernel by the unit that owns their head, every finished tile leaves through
  //    the fused epilogue E (epi_pass.cuh).  See the epilogue branch below.
  const int deferred = mode == 1;
  const bool fused = mode == 2;
  using C = Cfg2<BLOCK_N>;
  // optional phase trace (debug): 16 %globaltimer stamps (ns, one clock for the whole GPU) per CTA
  auto mark = [&](int i) {
    if (trace) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      trace[static_cast<size_t>(blockIdx.x) * 16 + i] = t;
    }
  };
  if (threadIdx.x == 64) mark(0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // leader's copy is the live one
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };       // both CTAs (multicast commit)
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };   // both CTAs (multicast commit)
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + 2 + a); };  // leader's copy, 8 arrivals
  const uint32_t fix_bar = bar_base + 8u * (2 * C::kStages + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const int pairs_n = (N + 2 * kSlab - 1) / (2 * kSlab);  // 256-row tiles
  const int ntt = (T + BLOCK_N - 1) / BLOCK_N;
  const int KB = (K + kBlockK - 1) / kBlockK;
  const int units = gridDim.x >> 1;
  const int unit = blockIdx.x >> 1;
  const int cta = blockIdx.x;
  // Schedule.  Stream-K: one contiguous range of (tile, k-block) iterations per pair.  Fused mode with more tiles than pairs
  // (gate_up of a large step) runs a HYBRID: first `dp_waves` waves of whole tiles, wave w giving pair u tile w*units + u,
  // then stream-K over the tiles that are left.  Tiles are ordered token-tile-minor, so in every wave each 256-row weight
  // slab is in use by ntt neighbouring pairs at the same time and is fetched from HBM once; with contiguous ranges the
  // pairs sit ~4.5 tiles apart, every pair streams its own slab, 74 slabs x 2 MB do not fit L2 together with the
  // activations and each slab came from DRAM once per token tile (ncu, T = 2048: 2.95x the algorithmic bytes).
  const int tiles_total = pairs_n * ntt;
  const int dp_waves = (mode == 2 && E.hybrid) ? hybrid_dp_waves(tiles_total, units, KB) : 0;
  const int dp_tiles = dp_waves * units;
  const long long total = static_cast<long long>(tiles_total - dp_tiles) * KB;   // the stream-K part
  const long long it_begin = range_begin(unit, total, units);
  const long long it_end = range_begin(unit + 1, total, units);
  // this pair's iterations as one virtual index v: [0, v_dp) the whole tiles, [v_dp, v_end) its stream-K range
  const long long v_dp = static_cast<long long>(dp_waves) * KB;
  const long long v_end = v_dp + (it_end - it_begin);
  auto tile_kb_at = [&](long long v, int* kb) {
    if (v < v_dp) {
      const int w = static_cast<int>(v / KB);
      *kb = static_cast<int>(v - static_cast<long long>(w) * KB);
      return w * units + unit;
    }
    const long long it = it_begin + (v - v_dp);
    const int t = static_cast<int>(it / KB);
    *kb = static_cast<int>(it - static_cast<long long>(t) * KB);
    return dp_tiles + t;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_x);
    if (tma_store) tma_prefetch_desc(&tm_out);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < C::kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 8);  // 4 epilogue warps of each CTA
      }
      mbar_init(fix_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc2(tmem_slot, C::kTmemCols);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync();  // barrier inits + TMEM allocation visible to the peer CTA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_launch();
  if (threadIdx.x == 64) mark(1);

  auto seg_at = [&](long long v) {
    Seg s;
    s.tile = tile_kb_at(v, &s.kb0);
    const long long rem = v_end - v;
    s.kb1 = (v < v_dp || KB - s.kb0 <= rem) ? KB : s.kb0 + static_cast<int>(rem);
    return s;
  };
  // token layout of a tile: chunk c covers tokens [t0 + c*256, ...), nc_c = its (16-padded) width
  auto chunk_n = [&](int tt, int c) {
    const int tile_end = (tt + 1) * BLOCK_N < T ? (tt + 1) * BLOCK_N : T;   // a 384-token tile's second chunk is 128 wide
    const int rem = tile_end - tt * BLOCK_N - c * C::kChunk;
    return rem <= 0 ? 0 : (rem >= C::kChunk ? C::kChunk : ((rem + 15) & ~15));
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      const uint64_t w_hint = ntt > 1 ? kEvictNormal : kEvictFirst;
      int pre = 0;
      for (long long v = 0; v < v_end && pre < C::kStages; ++v, ++pre) {
        int kb;
        const int tile = tile_kb_at(v, &kb);
        if (leader) mbar_arrive_expect_tx(full_bar(pre), 2u * C::kStageBytes);
        tma_load_2d_pair(smem_base + pre * C::kStageBytes, &tm_w, full_bar(pre), kb * kBlockK,
                         (tile / ntt) * 2 * kSlab + static_cast<int>(rank) * kSlab, w_hint);
      }
      griddep_wait();
      mark(12);
      int stage = 0, idx = 0;
      uint32_t phase = 0;
      for (long long v = 0; v < v_end; ++v, ++idx) {
        int kb;
        const int tile = tile_kb_at(v, &kb);
        const int slab2 = tile / ntt, tt = tile - slab2 * ntt;
        const uint32_t sa = smem_base + stage * C::kStageBytes;
        if (idx >= pre) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2u * C::kStageBytes);
          tma_load_2d_pair(sa, &tm_w, full_bar(stage), kb * kBlockK, slab2 * 2 * kSlab + static_cast<int>(rank) * kSlab, w_hint);
        }
#pragma unroll
        for (int c = 0; c < C::kNch; ++c) {
          // this CTA stages the rank-th half of chunk c (the box is always kHalf rows; rows past nc/2 are ignored)
          const int nc = chunk_n(tt, c);
          const int row0 = tt * BLOCK_N + c * C::kChunk + static_cast<int>(rank) * (nc >> 1);
          tma_load_2d_pair(sa + kABytes + c * (C::kHalf * kBlockK * 2), &tm_x, full_bar(stage), kb * kBlockK,
                           nc > 0 ? row0 : T /* fully out of range -> zero fill */, kEvictLast);
        }
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      mark(13);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && leader) {
      griddep_wait();
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long it = 0; it < v_end;) {
        Seg sg = seg_at(it);
        const int slab2 = sg.tile / ntt, tt = sg.tile - slab2 * ntt;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (it == 0 && kb == sg.kb0) mark(10);
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint64_t a_desc = umma_desc_kmajor_sw128(sa);
#pragma unroll
          for (int c = 0; c < C::kNch; ++c) {
            const int nc = chunk_n(tt, c);
            if (nc == 0) continue;
            const uint32_t idesc = umma_idesc_bf16(2 * kSlab, nc);
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N + c * C::kChunk;
            const uint64_t b_desc = umma_desc_kmajor_sw128(sa + kABytes + c * (C::kHalf * kBlockK * 2));
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma2_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb > sg.kb0 || k > 0) ? 1u : 0u);
          }
          umma2_commit_both(empty_bar(stage));
          if (kb == sg.kb1 - 1) umma2_commit_both(tfull_bar(acc));
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (++acc == C::kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
        it += sg.kb1 - sg.kb0;
      }
      mark(11);
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (both CTAs, own 128 rows)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    griddep_wait();
    const int epi_tid = threadIdx.x - 64;
    constexpr int kSlot = BLOCK_N * kSlab;
    int acc = 0;
    uint32_t acc_phase = 0;
    int fix_tile[2] = {-1, -1};
    // complete tiles leave through smem: bf16 rows staged as [32 tokens][128 weight rows] and written by TMA, so the
    // accumulator drain is tcgen05.ld + st.shared only (scalar 2-byte global stores measured 15-18 us per 512-token
    // tile and left the tensor pipe idle a third of the time at T=2048)
    // (deferred) split tiles leave the same way: fp32 [32 tokens][128 rows] chunks are contiguous in the workspace
    // slot, so each is one cp.async.bulk store.
    const uint32_t store_stage = bar_base + 256;
    uint8_t* store_ptr = smem_raw + (store_stage - smem_u32(smem_raw));
    int sbuf = 0;
    for (long long it = 0; it < v_end;) {   // `it` is the virtual index here (== offset into the stream-K range when dp_waves == 0)
      Seg sg = seg_at(it);
      const int slab2 = sg.tile / ntt, tt = sg.tile - slab2 * ntt;
      const int t0 = tt * BLOCK_N;
      int n_eff = 0;
#pragma unroll
      for (int c = 0; c < C::kNch; ++c) n_eff += chunk_n(tt, c);
      const int n = (slab2 * 2 + static_cast<int>(rank)) * kSlab + row;
      const bool complete = (sg.kb0 == 0 && sg.kb1 == KB);
      const int slot = (it == 0) ? 0 : 1;
      float* wslot;
      if (deferred) {
        // partial segment index = table[tile].first + (this unit's position among the units sharing the tile)
        const int u0 = unit_of_iter(static_cast<long long>(sg.tile - dp_tiles) * KB, total, units);
        const int seg = complete ? 0 : __ldg(&seg_table[sg.tile]).x + (unit - u0);
        wslot = ws + (static_cast<size_t>(seg) * 2 + rank) * kSlot;
      } else {
        wslot = ws + (static_cast<size_t>(cta) * 2 + slot) * kSlot;
      }

      if (fused) {
        // A unit's range starts with the TAIL (or a middle piece) of a tile and ends with the HEAD of another.  Pieces that
        // do not contain k-block 0 are parked in the workspace as fp32 [token][128 rows] with a release flag per unit; the
        // unit that owns the head computes it LAST in its range, so the other pieces were published long before (a piece
        // that is a unit's whole range finishes at the same moment: a wait of one flag round trip), adds them in unit order
        // and runs the fused epilogue.  No unit ever waits for the finisher: no cycle, and every CTA is co-resident.
        const bool publish = sg.kb0 > 0;
        const bool head = sg.kb0 == 0 && sg.kb1 < KB;
        const int slab = slab2 * 2 + static_cast<int>(rank);
        int parts = 0;
        if (head) {
          parts = unit_of_iter(static_cast<long long>(sg.tile - dp_tiles + 1) * KB - 1, total, units) - unit;
          if (epi_tid == 0)
            for (int p = 1; p <= parts; ++p)
              while (ld_acquire(E.flags + (unit + p) * 2 + static_cast<int>(rank)) != E.epoch) __nanosleep(32);
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        if (epi_tid == 0 && it == 0) mark(2);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
        if (publish) {
          float* dst = ws + (static_cast<size_t>(unit) * 2 + rank) * kSlot;
          for (int c0 = 0; c0 < n_eff; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < n_eff) dst[static_cast<size_t>(c0 + j) * kSlab + row] = __uint_as_float(v[j]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(tempty_bar(acc));
            else mbar_arrive_remote(tempty_bar(acc), 0);
          }
          __threadfence();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (epi_tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(E.flags + unit * 2 + static_cast<int>(rank)), "r"(E.epoch) : "memory");
        } else {
          // 64 tokens per round: two tcgen05.ld in flight, one wait, one barrier — the accumulator drain is exposed at
          // 512-token tiles (TMEM has a single accumulator stage), so its latency chain is what the tensor pipe waits for
          for (int c0 = 0; c0 < n_eff; c0 += E.wide ? 64 : 32) {
            const bool two = E.wide && c0 + 32 < n_eff;
            uint4 res_pref[2][4];
            if (E.epi == GEMM3_EPI_RESADD) {
              epi::resadd_prefetch(E, epi_tid, t0 + c0, T, slab * kSlab, res_pref[0]);
              if (two) epi::resadd_prefetch(E, epi_tid, t0 + c0 + 32, T, slab * kSlab, res_pref[1]);
            }
            uint32_t v0[32], v1[32];
            tmem_ld_32x32(taddr + c0, v0);
            if (two) tmem_ld_32x32(taddr + c0 + 32, v1);
            tmem_ld_wait();
            float f0[32], f1[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              f0[j] = __uint_as_float(v0[j]);
              f1[j] = two ? __uint_as_float(v1[j]) : 0.f;
            }
            for (int p = 1; p <= parts; ++p) {
              const float* pp = ws + (static_cast<size_t>(unit + p) * 2 + rank) * kSlot + static_cast<size_t>(c0) * kSlab + row;
#pragma unroll
              for (int j = 0; j < 32; ++j) f0[j] += (c0 + j < n_eff) ? __ldcg(pp + j * kSlab) : 0.f;
              if (two) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f1[j] += (c0 + 32 + j < n_eff) ? __ldcg(pp + (32 + j) * kSlab) : 0.f;
              }
            }
            __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(store_ptr + sbuf * C::kStoreBuf);   // 64 tokens x 128 rows bf16
#pragma unroll
            for (int j = 0; j < 32; ++j) sb[j * kSlab + row] = __float2bfloat16_rn(f0[j]);
            if (two) {
#pragma unroll
              for (int j = 0; j < 32; ++j) sb[(32 + j) * kSlab + row] = __float2bfloat16_rn(f1[j]);
            }
            // one barrier per round: a thread past it has finished its pass over the buffer staged two rounds ago
            asm volatile("bar.sync 1, 128;" ::: "memory");
            epi::pass(E, sb, epi_tid, t0 + c0, T, slab, res_pref[0]);
            if (two) epi::pass(E, sb + 32 * kSlab, epi_tid, t0 + c0 + 32, T, slab, res_pref[1]);
            sbuf ^= 1;
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(tempty_bar(acc));
            else mbar_arrive_remote(tempty_bar(acc), 0);
          }
        }
        if (++acc == C::kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
        it += sg.kb1 - sg.kb0;
        continue;
      }

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (epi_tid == 0 && it == 0) mark(2);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
      const bool wide = complete && tma_store && BLOCK_N >= 256 && E.wide;
      // complete tiles of the large steps: 64 tokens per round (two tcgen05.ld in flight, one pair of barriers, two TMA
      // stores in one bulk group) — the drain of a 512-token tile is exposed, its latency chain is tensor-pipe idle time
      for (int c0 = 0; wide && c0 < n_eff; c0 += 64) {
        const bool two = c0 + 32 < n_eff;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(taddr + c0, v0);
        if (two) tmem_ld_32x32(taddr + c0 + 32, v1);
        tmem_ld_wait();
        if (epi_tid == 0) bulk_wait_group_read<1>();  // the stores issued two rounds ago have left their buffer
        asm volatile("bar.sync 1, 128;" ::: "memory");
        __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(store_ptr + sbuf * C::kStoreBuf);
#pragma unroll
        for (int j = 0; j < 32; ++j) sb[j * kSlab + row] = __float2bfloat16_rn(__uint_as_float(v0[j]));
        if (two) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sb[(32 + j) * kSlab + row] = __float2bfloat16_rn(__uint_as_float(v1[j]));
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (epi_tid == 0) {
          const int n_first = (slab2 * 2 + static_cast<int>(rank)) * kSlab;
          tma_store_2d(&tm_out, store_stage + sbuf * C::kStoreBuf, n_first, t0 + c0);
          if (two) tma_store_2d(&tm_out, store_stage + sbuf * C::kStoreBuf + 32 * kSlab * 2, n_first, t0 + c0 + 32);
          bulk_commit_group();
        }
        sbuf ^= 1;
      }
      for (int c0 = 0; !wide && c0 < n_eff; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        if (complete && tma_store) {
          if (epi_tid == 0) bulk_wait_group_read<1>();  // the store issued two chunks ago has left its buffer
          asm volatile("bar.sync 1, 128;" ::: "memory");
          __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(store_ptr + sbuf * C::kStoreBuf);
#pragma unroll
          for (int j = 0; j < 32; ++j) sb[j * kSlab + row] = __float2bfloat16_rn(__uint_as_float(v[j]));
          fence_proxy_async();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (epi_tid == 0) {
            // rows t >= T and columns n >= N are clipped by the tensor map
            tma_store_2d(&tm_out, store_stage + sbuf * C::kStoreBuf, (slab2 * 2 + static_cast<int>(rank)) * kSlab,
                         t0 + c0);
            bulk_commit_group();
          }
          sbuf ^= 1;
        } else if (!complete && deferred && tma_store) {
          if (epi_tid == 0) bulk_wait_group_read<1>();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          float* sf = reinterpret_cast<float*>(store_ptr + sbuf * C::kStoreBuf);
#pragma unroll
          for (int j = 0; j < 32; ++j) sf[j * kSlab + row] = __uint_as_float(v[j]);
          fence_proxy_async();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (epi_tid == 0) {
            const int cols = (n_eff - c0) < 32 ? (n_eff - c0) : 32;
            bulk_store_1d(wslot + static_cast<size_t>(c0) * kSlab, store_stage + sbuf * C::kStoreBuf,
                          static_cast<uint32_t>(cols) * kSlab * 4);
            bulk_commit_group();
          }
          sbuf ^= 1;
        } else if (complete) {
          if (n < N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int t = t0 + c0 + j;
              if (t < T) out[static_cast<size_t>(t) * ldo + n] = __float2bfloat16_rn(__uint_as_float(v[j]));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < n_eff) wslot[(c0 + j) * kSlab + row] = __uint_as_float(v[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_remote(tempty_bar(acc), 0);
      }
      if (!complete && !deferred) {
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (epi_tid == 0) atomicAdd(&counters[2 * (sg.tile * 2 + static_cast<int>(rank))], 1);
        fix_tile[slot] = sg.tile;
      }
      if (++acc == C::kAccStages) {
        acc = 0;
        acc_phase ^= 1u;
      }
      it += sg.kb1 - sg.kb0;
    }

    if (epi_tid == 0) mark(3);
    // -------- fix-up (per CTA, same-rank CTAs of the other pairs are the peers)
    struct Fix {
      int j, u0, nseg, cb, ncol, t0, n, cid;
    };
    Fix fx[2];
    int nfix = 0;
    uint32_t total_bytes = 0;
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int j = fix_tile[f];
      if (j < 0) continue;
      Fix& x = fx[nfix];
      x.j = j;
      x.cid = j * 2 + static_cast<int>(rank);
      x.u0 = unit_of_iter(static_cast<long long>(j) * KB, total, units);
      const int u1 = unit_of_iter(static_cast<long long>(j + 1) * KB - 1, total, units);
      x.nseg = u1 - x.u0 + 1;
      const int si = unit - x.u0;
      const int slab2 = j / ntt, tt = j - slab2 * ntt;
      x.t0 = tt * BLOCK_N;
      const int cols = (T - x.t0) >= BLOCK_N ? BLOCK_N : (T - x.t0);
      x.cb = (si * cols) / x.nseg;
      x.ncol = ((si + 1) * cols) / x.nseg - x.cb;
      x.n = (slab2 * 2 + static_cast<int>(rank)) * kSlab + row;
      total_bytes += static_cast<uint32_t>(x.nseg) * x.ncol * kSlab * 4;
      ++nfix;
    }
    constexpr uint32_t kRing = C::kStages * C::kStageBytes;
    const float* fix_smem = reinterpret_cast<const float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
    uint32_t fix_phase = 0;
    auto src_of = [&](const Fix& x, int p, int col) {
      const long long pbeg = range_begin(x.u0 + p, total, units);
      const int pslot = (static_cast<int>(pbeg / KB) == x.j) ? 0 : 1;
      const int pcta = (x.u0 + p) * 2 + static_cast<int>(rank);
      return ws + (static_cast<size_t>(pcta) * 2 + pslot) * kSlot + static_cast<size_t>(col) * kSlab;
    };
    auto finish = [&](const Fix& x) {
      if (atomicAdd(&counters[2 * x.cid + 1], 1) == x.nseg - 1) {
        counters[2 * x.cid] = 0;
        counters[2 * x.cid + 1] = 0;
      }
    };
    if (nfix > 0 && total_bytes <= kRing) {
      // fast path (decode shapes): both split tiles in ONE L2 round trip
      uint32_t off[2] = {0, 0};
      if (nfix == 2) off[1] = static_cast<uint32_t>(fx[0].nseg) * fx[0].ncol * kSlab * 4;
      if (epi_tid == 0) {
        for (int f = 0; f < nfix; ++f)
          while (ld_acquire(&counters[2 * fx[f].cid]) < fx[f].nseg) __nanosleep(20);
        mark(4);
        if (total_bytes) {
          asm volatile("fence.proxy.async;" ::: "memory");
          mbar_arrive_expect_tx(fix_bar, total_bytes);
          for (int f = 0; f < nfix; ++f) {
            const uint32_t pb = static_cast<uint32_t>(fx[f].ncol) * kSlab * 4;
            if (pb == 0) continue;
            for (int p = 0; p < fx[f].nseg; ++p)
              bulk_load_1d(smem_base + off[f] + static_cast<uint32_t>(p) * pb, src_of(fx[f], p, fx[f].cb), pb, fix_bar);
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (total_bytes) mbar_wait(fix_bar, fix_phase);
      if (epi_tid == 0) mark(5);
      for (int f = 0; f < nfix; ++f) {
        const Fix& x = fx[f];
        const float* base = fix_smem + off[f] / 4;
        for (int col = 0; col < x.ncol; ++col) {
          float sum = 0.f;
          for (int p = 0; p < x.nseg; ++p) sum += base[(p * x.ncol + col) * kSlab + row];
          if (x.n < N) out[static_cast<size_t>(x.t0 + x.cb + col) * ldo + x.n] = __float2bfloat16_rn(sum);
        }
      }
      if (epi_tid == 0) mark(8);
      if (epi_tid == 0)
        for (int f = 0; f < nfix; ++f) finish(fx[f]);
      if (epi_tid == 0) mark(9);
    } else {
      // general path: one tile at a time, the column slice in pieces that fit the ring
      for (int f = 0; f < nfix; ++f) {
        const Fix& x = fx[f];
        if (epi_tid == 0)
          while (ld_acquire(&counters[2 * x.cid]) < x.nseg) __nanosleep(20);
        const int cstep = static_cast<int>(kRing / (static_cast<uint32_t>(x.nseg) * kSlab * 4));
        for (int c0 = 0; c0 < x.ncol; c0 += cstep) {
          const int nc = (x.ncol - c0) < cstep ? (x.ncol - c0) : cstep;
          const uint32_t pb = static_cast<uint32_t>(nc) * kSlab * 4;
          if (epi_tid == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            mbar_arrive_expect_tx(fix_bar, pb * x.nseg);
            for (int p = 0; p < x.nseg; ++p)
              bulk_load_1d(smem_base + static_cast<uint32_t>(p) * pb, src_of(x, p, x.cb + c0), pb, fix_bar);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          mbar_wait(fix_bar, fix_phase);
          fix_phase ^= 1u;
          for (int col = 0; col < nc; ++col) {
            float sum = 0.f;
            for (int p = 0; p < x.nseg; ++p) sum += fix_smem[(p * nc + col) * kSlab + row];
            if (x.n < N) out[static_cast<size_t>(x.t0 + x.cb + c0 + col) * ldo + x.n] = __float2bfloat16_rn(sum);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (x.ncol == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        if (epi_tid == 0) finish(x);
      }
    }
  }

  if (threadIdx.x == 64) {
    // thread 64 (epi_tid 0) issued every bulk store of this CTA: its smem source must outlive the reads; the writes
    // themselves are ordered before the dependent kernel by grid completion (as in any TMA-store epilogue)
    bulk_wait_group_read<0>();
    mark(6);
  }
  tc_fence_before();
  cluster_sync();  // neither CTA may exit (or free TMEM) while the peer can still signal its barriers / read its smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, C::kTmemCols);
  }
  if (threadIdx.x == 64) mark(7);
}

// debug (b200_op_gemm_trace): device buffer of B200_GEMM_TRACE_LAUNCHES (default 1) blocks of 148 x 16 stamps;
// launch i after the buffer was set writes block i mod that count
long long* g_trace = nullptr;
int g_trace_launch = 0;
long long* trace_block() {
  static const int blocks = [] { const char* e = getenv("B200_GEMM_TRACE_LAUNCHES"); return e && atoi(e) > 0 ? atoi(e) : 1; }();
  if (!g_trace) return nullptr;
  return g_trace + static_cast<size_t>(g_trace_launch++ % blocks) * 148 * 16;
}

int units_for(const GemmPlan& p, int ntt) {
  const int pairs_n = (p.N + 2 * kSlab - 1) / (2 * kSlab);
  const int KB = (p.K + kBlockK - 1) / kBlockK;
  const long long total = static_cast<long long>(pairs_n) * ntt * KB;
  long long u = total / 4;
  if (u < 1) u = 1;
  const int max_units = p.max_ctas / 2;
  // Several token tiles, a shallow K and almost as many tiles as pairs (qkv / o_proj of a prefill burst: 72 and 48 tiles
  // for 74 pairs): one whole tile per pair.  Stream-K over all pairs would cut EVERY tile into 2-3 fp32 pieces — two exposed
  // 512-token epilogues per pair and ~75 MB of segments for the consumer kernel to sum — for the sake of a mainloop that is
  // only ~20 us long; idling a third of the pairs costs less (ncu, T = 1481: RMSNorm after o_proj 30 us -> bf16 input only).
  const int tiles = pairs_n * ntt;
  if (ntt > 1 && KB <= 64 && tiles <= max_units && tiles * 5 >= max_units * 3) return tiles;
  return static_cast<int>(u < max_units ? u : max_units);
}

// Output tensor maps, one per (buffer, rows, N, ld): a handful in the engine (x, qkv, gate_up, logits x the T of the
// step); encoding is a host-only driver call.  B200_GEMM_TMA_STORE=0 falls back to per-thread stores.
CUtensorMap out_map_for(const void* out, int T, int N, int ldo, int* ok) {
  struct Key {
    const void* p;
    int T, N, ld;
    bool operator==(const Key& o) const { return p == o.p && T == o.T && N == o.N && ld == o.ld; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      return std::hash<const void*>()(k.p) ^ (static_cast<size_t>(k.T) * 0x9E3779B97F4A7C15ull) ^
             (static_cast<size_t>(k.N) << 20) ^ static_cast<size_t>(k.ld);
    }
  };
  static const bool enabled = [] { const char* e = getenv("B200_GEMM_TMA_STORE"); return !e || atoi(e) != 0; }();
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  static std::mutex mu;
  const CUtensorMap dummy = {};
  *ok = 0;
  if (!enabled) return dummy;
  std::lock_guard<std::mutex> lk(mu);
  const Key k{out, T, N, ldo};
  auto it = cache.find(k);
  if (it == cache.end()) {
    CUtensorMap tm;
    if (gemm_make_out_map(&tm, out, T, N, ldo) != 0) return dummy;
    if (cache.size() > 16384) cache.clear();
    it = cache.emplace(k, tm).first;
  }
  *ok = 1;
  return it->second;
}

// A/B knobs, re-read from the environment whenever an engine (or the op-level ABI) initialises: B200_GEMM_WIDE_EPI=0 selects
// 32-token epilogue rounds, B200_GEMM_HYBRID=0 plain stream-K ranges in fused mode
int g_wide_epi = 1, g_hybrid = 1;

template <int BLOCK_N>
int launch2(const GemmPlan& p, const CUtensorMap& tm_x, __nv_bfloat16* out, int ldo, int T, cudaStream_t st,
            int mode = 0, const Gemm2Epi* fe = nullptr) {
  using C = Cfg2<BLOCK_N>;
  static std::atomic<unsigned long long> attr_done{0};
  if (!ensure_dynamic_smem(gemm2_streamk_kernel<BLOCK_N>, C::kSmemBytes, &attr_done)) return -3;
  const int ntt = (T + BLOCK_N - 1) / BLOCK_N;
  const int units = units_for(p, ntt);
  int tma_store = 0;
  Gemm2Epi E;
  memset(&E, 0, sizeof(E));
  CUtensorMap tm_out = {};
  if (mode == 2) E = *fe;
  else tm_out = out_map_for(out, T, p.N, ldo, &tma_store);
  E.wide = g_wide_epi;
  E.hybrid = g_hybrid;
  cudaError_t e = launch_pdl(gemm2_streamk_kernel<BLOCK_N>, dim3(2 * units), dim3(kThreads), C::kSmemBytes, st, p.tm_w, tm_x,
                             tm_out, tma_store, out, ldo, p.ws, p.counters, p.N, T, p.K, static_cast<const int2*>(p.seg_table), mode, E,
                             trace_block());
  return e == cudaSuccess ? 0 : -4;
}

// out[t, n] = bf16(sum of the tile's segments): the generic consumer of deferred partials.
__global__ void reduce_partials_kernel(PartialView v, __nv_bfloat16* __restrict__ out, int ldo, int T, int N) {
  griddep_wait();
  griddep_launch();
  const int t = blockIdx.y;
  const int n0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (n0 >= N) return;
  float f[8];
  load8_partials(v, t, n0, f);
  if (n0 + 8 <= N) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]);
    u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]);
    u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * ldo + n0) = u;
  } else {
    for (int j = 0; j < 8 && n0 + j < N; ++j) out[static_cast<size_t>(t) * ldo + n0 + j] = __float2bfloat16_rn(f[j]);
  }
}

}  // namespace

int gemm2_units_for(const GemmPlan& p, int ntt) { return units_for(p, ntt); }

// Host-only view of the large-step schedule for (N, K, T) on a device with `sms` SMs: what gemm2_block_n_for_plan, units_for and
// the engine's fusion rule decide.  out[0] token-tile size, [1] token tiles, [2] tiles, [3] CTA pairs launched, [4] 1 = one whole
// tile per pair (no split tile), [5] 1 = the engine fuses the elementwise neighbour into this launch (a tile per pair or
// more), [6] whole-tile waves ahead of the stream-K tail in that fused form, [7] 0.
void gemm2_schedule_query(int N, int K, int T, int sms, int* out) {
  GemmPlan p;
  memset(&p, 0, sizeof(p));
  p.N = N;
  p.K = K;
  p.max_ctas = sms;
  const int bn = gemm2_block_n_for_plan(p, T);
  const int ntt = (T + bn - 1) / bn;
  const int pairs_n = (N + 2 * kSlab - 1) / (2 * kSlab);
  const int tiles = pairs_n * ntt;
  const int units = units_for(p, ntt);
  out[0] = bn;
  out[1] = ntt;
  out[2] = tiles;
  out[3] = units;
  out[4] = units == tiles && ntt > 1;
  out[5] = T > 128 && tiles >= sms / 2;
  out[6] = out[5] && g_hybrid ? hybrid_dp_waves(tiles, units, (K + kBlockK - 1) / kBlockK) : 0;
  out[7] = 0;
}

void gemm2_read_env() {
  const char* w = getenv("B200_GEMM_WIDE_EPI");
  const char* h = getenv("B200_GEMM_HYBRID");
  g_wide_epi = w ? atoi(w) : 1;
  g_hybrid = h ? atoi(h) : 1;
}

int gemm_plan_build_table(GemmPlan* p, int max_tokens) {
  if (p->seg_table) return 0;
  const int pairs_n = (p->N + 2 * kSlab - 1) / (2 * kSlab);
  const int KB = (p->K + kBlockK - 1) / kBlockK;
  int max_ntt = (max_tokens + 383) / 384;   // 384-token tiles are the narrowest multi-tile form
  if (max_ntt < 1) max_ntt = 1;
  if (max_ntt > 64) max_ntt = 64;
  p->max_ntt = max_ntt;
  std::vector<int2> tab;
  for (int ntt = 1; ntt <= max_ntt; ++ntt) {
    const long long total = static_cast<long long>(pairs_n) * ntt * KB;
    const int units = units_for(*p, ntt);
    auto unit_of = [&](long long x) { return static_cast<int>(((x + 1) * units + total - 1) / total - 1); };
    p->table_off[ntt] = static_cast<int>(tab.size());
    int seg = 0;
    for (int j = 0; j < pairs_n * ntt; ++j) {
      const int u0 = unit_of(static_cast<long long>(j) * KB), u1 = unit_of(static_cast<long long>(j + 1) * KB - 1);
      const int nseg = u1 - u0 + 1;
      tab.push_back(make_int2(nseg > 1 ? seg : 0, nseg));
      if (nseg > 1) seg += nseg;
    }
    p->table_segs[ntt] = seg;
  }
  if (cudaMalloc(&p->seg_table, tab.size() * sizeof(int2)) != cudaSuccess) return -7;
  if (cudaMemcpy(p->seg_table, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess) return -7;
  return 0;
}

void gemm_plan_destroy(GemmPlan* p) {
  if (p->seg_table) cudaFree(p->seg_table);
  p->seg_table = nullptr;
}

// every unit has at most two partial segments (the head and the tail of its range)
size_t gemm_deferred_ws_bytes(int max_ctas) {
  return static_cast<size_t>(max_ctas / 2) * 2 * 2 * 512 * kSlab * sizeof(float);
}

int gemm_run_deferred(const GemmPlan& p, const CUtensorMap& tm_x, int block_n, void* out, int ldo, int T, cudaStream_t st,
                      PartialView* view) {
  if (T <= 0 || !p.seg_table) return -8;
  if (ldo % 8 != 0) return -5;  // consumers read the dense tiles with 16-byte loads
  const int ntt = (T + block_n - 1) / block_n;
  const bool whole = units_for(p, ntt) == ((p.N + 2 * kSlab - 1) / (2 * kSlab)) * ntt;   // one whole tile per pair: no segments
  if (ntt > 1 && block_n != 512 && !(block_n == 384 && whole)) return -8;   // consumers index segment tables by 512-token tiles
  if (ntt > p.max_ntt) return -8;
  const size_t need = static_cast<size_t>(p.table_segs[ntt]) * 2 * block_n * kSlab * sizeof(float);
  if (need > p.ws_bytes) return -9;
  GemmPlan q = p;
  q.seg_table = p.seg_table + p.table_off[ntt];
  view->ws = p.ws;
  view->table = q.seg_table;
  view->dense = static_cast<const __nv_bfloat16*>(out);
  view->ld_dense = ldo;
  view->slot = block_n * kSlab;
  view->ntt = ntt;
  view->block_n = block_n;
  view->bn_shift = block_n == 32 ? 5 : block_n == 64 ? 6 : block_n == 128 ? 7 : block_n == 256 ? 8 : 9;
  if (whole && ntt > 1) *view = no_partials();   // every tile is complete: the consumer reads the bf16 tensor, no table look-ups
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  switch (block_n) {
    case 32: return launch2<32>(q, tm_x, o, ldo, T, st, 1);
    case 64: return launch2<64>(q, tm_x, o, ldo, T, st, 1);
    case 128: return launch2<128>(q, tm_x, o, ldo, T, st, 1);
    case 256: return launch2<256>(q, tm_x, o, ldo, T, st, 1);
    case 384: return launch2<384>(q, tm_x, o, ldo, T, st, 1);
    case 512: return launch2<512>(q, tm_x, o, ldo, T, st, 1);
    default: return -6;
  }
}

int gemm2_run_fused(const GemmPlan& p, const CUtensorMap& tm_x, int block_n, int T, const Gemm2Epi& e, cudaStream_t st) {
  if (T <= 0) return 0;
  if (p.N % (2 * kSlab) != 0 || e.ldo % 8 != 0 || !e.flags || !e.out) return -5;
  if (e.epi != GEMM3_EPI_PLAIN && e.epi != GEMM3_EPI_RESADD && e.epi != GEMM3_EPI_SILU && e.epi != GEMM3_EPI_ROPE_KV) return -6;
  const int ntt = (T + block_n - 1) / block_n;
  // one parked piece per unit: [unit][rank][block_n tokens][128 rows] fp32
  if (static_cast<size_t>(units_for(p, ntt)) * 2 * block_n * kSlab * sizeof(float) > p.ws_bytes) return -9;
  switch (block_n) {
    case 32: return launch2<32>(p, tm_x, nullptr, 0, T, st, 2, &e);
    case 64: return launch2<64>(p, tm_x, nullptr, 0, T, st, 2, &e);
    case 128: return launch2<128>(p, tm_x, nullptr, 0, T, st, 2, &e);
    case 256: return launch2<256>(p, tm_x, nullptr, 0, T, st, 2, &e);
    case 384: return launch2<384>(p, tm_x, nullptr, 0, T, st, 2, &e);
    case 512: return launch2<512>(p, tm_x, nullptr, 0, T, st, 2, &e);
    default: return -6;
  }
}

int reduce_partials(const PartialView& v, void* out, int ldo, int T, int N, cudaStream_t st) {
  if (T <= 0 || !v.ws) return 0;   // no segments (every tile complete): `out` already holds the bf16 result
  dim3 grid((N / 8 + 127) / 128 + 1, T);
  cudaError_t e = launch_pdl(reduce_partials_kernel, grid, dim3(128), 0, st, v, static_cast<__nv_bfloat16*>(out), ldo, T, N);
  return e == cudaSuccess ? 0 : -4;
}

void gemm2_set_trace(long long* dev_ptr) {
  g_trace = dev_ptr;
  g_trace_launch = 0;
}

int gemm2_block_n_for(int T) { return T <= 32 ? 32 : T <= 64 ? 64 : T <= 128 ? 128 : T <= 256 ? 256 : 512; }

// Token-tile size for one projection of a step: 512 (256 up to 256 tokens, ...) unless 384-token tiles turn a launch with
// fewer tiles than pairs into one whole tile per pair with more pairs busy (o_proj of a prefill burst: 48 -> 64 tiles for 74).
int gemm2_block_n_for_plan(const GemmPlan& p, int T) {
  const int bn = gemm2_block_n_for(T);
  static const int allow = [] { const char* e = getenv("B200_GEMM_BN384"); return e ? atoi(e) : 1; }();
  if (bn != 512 || T <= 512 || !allow) return bn;
  const int pairs_n = (p.N + 2 * kSlab - 1) / (2 * kSlab), pairs = p.max_ctas / 2;
  const int KB = (p.K + kBlockK - 1) / kBlockK;
  const int t512 = pairs_n * ((T + 511) / 512), t384 = pairs_n * ((T + 383) / 384);
  if (KB <= 64 && t512 < pairs && t384 <= pairs && t384 > t512 && units_for(p, (T + 383) / 384) == t384) return 384;
  return bn;
}
int gemm2_x_box_rows(int block_n) { return (block_n > 256 ? 256 : block_n) / 2; }

int gemm2_run(const GemmPlan& p, const CUtensorMap& tm_x, int block_n, void* out, int ldo, int T, cudaStream_t st) {
  if (T <= 0) return 0;
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  switch (block_n) {
    case 32: return launch2<32>(p, tm_x, o, ldo, T, st);
    case 64: return launch2<64>(p, tm_x, o, ldo, T, st);
    case 128: return launch2<128>(p, tm_x, o, ldo, T, st);
    case 256: return launch2<256>(p, tm_x, o, ldo, T, st);
    case 384: return launch2<384>(p, tm_x, o, ldo, T, st);
    case 512: return launch2<512>(p, tm_x, o, ldo, T, st);
    default: return -6;
  }
}

}  // namespace b200
