// Dense projection GEMM for the per-step hot path:
//     out[t, n] = sum_k X[t, k] * W[n, k]        (bf16 x bf16 -> fp32 -> bf16)
// which is what the reference backend computes for qkv_proj / o_proj /
// gate_up_proj / down_proj / lm_head (vLLM model_executor/models/llama.py:81-121,
// :157-233; weights stored [N, K] K-major, outputs rounded to bf16 between ops).
//
// B200 design (not a port of any library kernel):
//   * "swap-AB": the 128-row weight slab is the UMMA M operand, the token tile
//     (32..256 tokens, runtime N in multiples of 16) is the UMMA N operand, so a
//     decode step with T=128 tokens streams every weight byte exactly once
//     through one tcgen05.mma M=128 tile per slab, and T is padded to 16 not 128.
//   * TMA (cp.async.bulk.tensor, 128B swizzle) feeds a multi-stage smem ring;
//     one elected thread issues tcgen05.mma into a double-buffered TMEM
//     accumulator; 4 epilogue warps drain TMEM with tcgen05.ld.
//   * stream-K persistent schedule: the (tile x k-block) iteration space is cut
//     into one contiguous range per SM, so every SM streams the same number of
//     weight bytes whatever N/128 is (32 slabs for o_proj would otherwise light
//     32 of 148 SMs).  Tiles that straddle CTAs are summed through an fp32 L2
//     workspace with a reduce-scatter fix-up at the end of the kernel.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "gemm.h"
#include "launch.h"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kSlab = 128;    // weight rows per tile == UMMA M
constexpr int kBlockK = 64;   // bf16 per k-block == one 128-byte swizzle span
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;  // warp0 TMA, warp1 MMA(+TMEM alloc), warps2-5 epilogue
constexpr int kEpiThreads = 128;
constexpr int kABytes = kSlab * kBlockK * 2;  // 16 KiB

template <int BLOCK_N>
struct Cfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;  // power of two for 32..256
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct Seg {
  int tile, kb0, kb1;
};

__device__ __forceinline__ long long it_begin_of(int cta, long long total, int grid) {
  return (static_cast<long long>(cta) * total) / grid;
}
// CTA whose range contains iteration x: largest c with floor(c*total/grid) <= x.
__device__ __forceinline__ int cta_of_iter(long long x, long long total, int grid) {
  long long c = ((x + 1) * grid + total - 1) / total - 1;
  return static_cast<int>(c);
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
gemm_streamk_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                    __nv_bfloat16* __restrict__ out, int ldo, float* __restrict__ ws,
                    int* __restrict__ counters, int N, int T, int K) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::kStages * C::kStageBytes;
  // barrier layout: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + 2 + a); };
  const uint32_t fix_bar = bar_base + 8u * (2 * C::kStages + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 5);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int slabs = (N + kSlab - 1) / kSlab;
  const int ntt = (T + BLOCK_N - 1) / BLOCK_N;
  const int KB = (K + kBlockK - 1) / kBlockK;
  const long long total = static_cast<long long>(slabs) * ntt * KB;
  const int grid = gridDim.x;
  const int cta = blockIdx.x;
  const long long it_begin = it_begin_of(cta, total, grid);
  const long long it_end = it_begin_of(cta + 1, total, grid);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < C::kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
      }
      mbar_init(fix_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_launch();  // the next kernel may start its own prologue; it waits for our completion before reading `out`

  auto seg_at = [&](long long it) {
    Seg s;
    s.tile = static_cast<int>(it / KB);
    s.kb0 = static_cast<int>(it - static_cast<long long>(s.tile) * KB);
    long long rem = it_end - it;
    s.kb1 = (KB - s.kb0 <= rem) ? KB : s.kb0 + static_cast<int>(rem);
    return s;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint64_t w_hint = ntt > 1 ? kEvictNormal : kEvictFirst;  // weights are read once per step
      // Weights do not depend on the previous kernel: fill the ring's weight halves BEFORE the
      // programmatic-dependency wait, so launch latency, prologue and the DRAM ramp of this GEMM
      // overlap the tail of the previous kernel.  The activation halves follow after the wait.
      int pre = 0;
      for (long long it = it_begin; it < it_end && pre < C::kStages; ++it, ++pre) {
        const int tile = static_cast<int>(it / KB), kb = static_cast<int>(it - static_cast<long long>(tile) * KB);
        mbar_arrive_expect_tx(full_bar(pre), C::kStageBytes);
        tma_load_2d(smem_base + pre * C::kStageBytes, &tm_w, full_bar(pre), kb * kBlockK, (tile / ntt) * kSlab, w_hint);
      }
      griddep_wait();
      int stage = 0, idx = 0;
      uint32_t phase = 0;
      for (long long it = it_begin; it < it_end; ++it, ++idx) {
        const int tile = static_cast<int>(it / KB), kb = static_cast<int>(it - static_cast<long long>(tile) * KB);
        const int slab = tile / ntt, tt = tile - slab * ntt;
        const uint32_t sa = smem_base + stage * C::kStageBytes;
        if (idx >= pre) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), C::kStageBytes);
          tma_load_2d(sa, &tm_w, full_bar(stage), kb * kBlockK, slab * kSlab, w_hint);
        }
        // activations are re-read by every slab: keep them in L2
        tma_load_2d(sa + kABytes, &tm_x, full_bar(stage), kb * kBlockK, tt * BLOCK_N, kEvictLast);
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      griddep_wait();
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long it = it_begin; it < it_end;) {
        Seg sg = seg_at(it);
        const int slab = sg.tile / ntt, tt = sg.tile - slab * ntt;
        const int rem_t = T - tt * BLOCK_N;
        const int n_eff = rem_t >= BLOCK_N ? BLOCK_N : ((rem_t + 15) & ~15);
        const uint32_t idesc = umma_idesc_bf16(kSlab, n_eff);
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint64_t a_desc = umma_desc_kmajor_sw128(sa);
          const uint64_t b_desc = umma_desc_kmajor_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // +32 B per UMMA_K step inside the 128 B swizzle span (start address is in 16 B units)
            umma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb > sg.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (kb == sg.kb1 - 1) umma_commit(tfull_bar(acc));
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
        it += sg.kb1 - sg.kb0;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    griddep_wait();                 // before the first write to out / workspace / counters
    const int q = warp & 3;         // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;  // weight row inside the slab == TMEM lane
    const int epi_tid = threadIdx.x - 64;
    constexpr int kSlot = BLOCK_N * kSlab;  // fp32 elements per partial slot
    int acc = 0;
    uint32_t acc_phase = 0;
    int fix_tile[2] = {-1, -1};
    for (long long it = it_begin; it < it_end;) {
      Seg sg = seg_at(it);
      const int slab = sg.tile / ntt, tt = sg.tile - slab * ntt;
      const int t0 = tt * BLOCK_N;
      const int rem_t = T - t0;
      const int n_eff = rem_t >= BLOCK_N ? BLOCK_N : ((rem_t + 15) & ~15);
      const int n = slab * kSlab + row;
      const bool complete = (sg.kb0 == 0 && sg.kb1 == KB);
      const int slot = (it == it_begin) ? 0 : 1;
      float* wslot = ws + (static_cast<size_t>(cta) * 2 + slot) * kSlot;

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
      for (int c0 = 0; c0 < n_eff; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        if (complete) {
          if (n < N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int t = t0 + c0 + j;
              if (t < T) out[static_cast<size_t>(t) * ldo + n] = __float2bfloat16_rn(__uint_as_float(v[j]));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (c0 + j < n_eff) wslot[(c0 + j) * kSlab + row] = __uint_as_float(v[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (!complete) {
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (epi_tid == 0) atomicAdd(&counters[2 * sg.tile], 1);
        fix_tile[slot] = sg.tile;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
      it += sg.kb1 - sg.kb0;
    }
    // -------- fix-up: every CTA that holds a partial of tile j reduces a 1/nseg token slice of it.
    // All of this CTA's MMAs have completed (the last segment's tmem_full was waited on), so the
    // smem ring is idle: the peers' fp32 slices are pulled into it with one cp.async.bulk each
    // (one L2 round trip for the whole slice instead of one per element) and summed from smem.
    // Both of a CTA's split tiles (head of its range, tail of its range) share ONE round trip
    // when their slices fit the ring together.
    struct Fix {
      int j, c0, nseg, cb, ncol, t0, n;
      uint32_t bytes, off;
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
      x.c0 = cta_of_iter(static_cast<long long>(j) * KB, total, grid);
      const int c1 = cta_of_iter(static_cast<long long>(j + 1) * KB - 1, total, grid);
      x.nseg = c1 - x.c0 + 1;
      const int si = cta - x.c0;
      const int slab = j / ntt, tt = j - slab * ntt;
      x.t0 = tt * BLOCK_N;
      const int cols = (T - x.t0) >= BLOCK_N ? BLOCK_N : (T - x.t0);
      x.cb = (si * cols) / x.nseg;
      x.ncol = ((si + 1) * cols) / x.nseg - x.cb;
      x.n = slab * kSlab + row;
      x.bytes = static_cast<uint32_t>(x.nseg) * x.ncol * kSlab * 4;
      x.off = total_bytes;
      total_bytes += x.bytes;
      ++nfix;
    }
    const bool merged = total_bytes <= static_cast<uint32_t>(C::kStages * C::kStageBytes);
    uint32_t fix_phase = 0;
    const float* fix_smem = reinterpret_cast<const float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
    auto pull = [&](int f0, int f1) {  // one thread: wait for the peers' partials, start the bulk copies
      uint32_t tot = 0;
      for (int f = f0; f < f1; ++f) {
        while (ld_acquire(&counters[2 * fx[f].j]) < fx[f].nseg) __nanosleep(20);
        tot += fx[f].bytes;
      }
      if (tot == 0) return;
      asm volatile("fence.proxy.async;" ::: "memory");  // peers wrote with generic stores; bulk copy reads via the async proxy
      mbar_arrive_expect_tx(fix_bar, tot);
      for (int f = f0; f < f1; ++f) {
        const Fix& x = fx[f];
        if (x.ncol == 0) continue;
        const uint32_t pb_bytes = static_cast<uint32_t>(x.ncol) * kSlab * 4;
        for (int p = 0; p < x.nseg; ++p) {
          const long long pbeg = it_begin_of(x.c0 + p, total, grid);
          const int pslot = (static_cast<int>(pbeg / KB) == x.j) ? 0 : 1;
          bulk_load_1d(smem_base + (merged ? x.off : 0u) + static_cast<uint32_t>(p) * pb_bytes,
                       ws + (static_cast<size_t>(x.c0 + p) * 2 + pslot) * kSlot + static_cast<size_t>(x.cb) * kSlab,
                       pb_bytes, fix_bar);
        }
      }
    };
    auto reduce = [&](int f) {
      const Fix& x = fx[f];
      const float* base = fix_smem + (merged ? x.off : 0u) / 4;
      for (int col = 0; col < x.ncol; ++col) {
        float sum = 0.f;
        for (int p = 0; p < x.nseg; ++p) sum += base[(p * x.ncol + col) * kSlab + row];
        if (x.n < N) out[static_cast<size_t>(x.t0 + x.cb + col) * ldo + x.n] = __float2bfloat16_rn(sum);
      }
    };
    auto finish = [&](int f) {  // one thread: last reducer of a tile resets its counters for the next launch
      if (atomicAdd(&counters[2 * fx[f].j + 1], 1) == fx[f].nseg - 1) {
        counters[2 * fx[f].j] = 0;
        counters[2 * fx[f].j + 1] = 0;
      }
    };
    if (nfix > 0) {
      if (merged) {
        if (epi_tid == 0) pull(0, nfix);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (total_bytes > 0) mbar_wait(fix_bar, fix_phase);
        for (int f = 0; f < nfix; ++f) reduce(f);
        if (epi_tid == 0)
          for (int f = 0; f < nfix; ++f) finish(f);
      } else {
        for (int f = 0; f < nfix; ++f) {
          if (epi_tid == 0) pull(f, f + 1);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (fx[f].bytes > 0) {
            mbar_wait(fix_bar, fix_phase);
            fix_phase ^= 1u;
          }
          reduce(f);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (epi_tid == 0) finish(f);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128B swizzle.
int make_tmap(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BLOCK_N>
int launch(const GemmPlan& p, const CUtensorMap& tm_x, __nv_bfloat16* out, int ldo, int T, cudaStream_t st) {
  using C = Cfg<BLOCK_N>;
  static std::atomic<unsigned long long> attr_done{0};
  if (!ensure_dynamic_smem(gemm_streamk_kernel<BLOCK_N>, C::kSmemBytes, &attr_done)) return -3;
  const int slabs = (p.N + kSlab - 1) / kSlab;
  const int ntt = (T + BLOCK_N - 1) / BLOCK_N;
  const int KB = (p.K + kBlockK - 1) / kBlockK;
  const long long total = static_cast<long long>(slabs) * ntt * KB;
  long long g = total / 4;  // at least ~4 k-blocks per CTA
  if (g < 1) g = 1;
  int grid = static_cast<int>(g < p.max_ctas ? g : p.max_ctas);
  cudaError_t e = launch_pdl(gemm_streamk_kernel<BLOCK_N>, dim3(grid), dim3(kThreads), C::kSmemBytes, st, p.tm_w, tm_x,
                             out, ldo, p.ws, p.counters, p.N, T, p.K);
  return e == cudaSuccess ? 0 : -4;
}

}  // namespace

static int g_variant = 2;
void gemm_set_variant(int v) { g_variant = (v == 1) ? 1 : 2; }
int gemm_variant() { return g_variant; }

int gemm_block_n_for(int T) {
  if (g_variant == 2) return gemm2_block_n_for(T);
  return T <= 32 ? 32 : T <= 64 ? 64 : T <= 128 ? 128 : 256;
}
int gemm_block_n_index(int bn) { return bn == 32 ? 0 : bn == 64 ? 1 : bn == 128 ? 2 : bn == 256 ? 3 : 4; }
int gemm_x_box_rows(int bn) { return g_variant == 2 ? gemm2_x_box_rows(bn) : (bn > 256 ? 256 : bn); }

size_t gemm_workspace_bytes(int max_ctas) {  // in-kernel fix-up slots (2 per CTA) == deferred segments (2 per unit x 2 ranks)
  return static_cast<size_t>(max_ctas) * 2 * 512 * kSlab * sizeof(float);
}

int gemm_plan_init(GemmPlan* p, const void* W, int N, int K, int ldw, float* ws, int* counters, int max_ctas) {
  memset(p, 0, sizeof(*p));
  p->N = N;
  p->K = K;
  p->w_ptr = W;
  p->ldw = ldw;
  p->ws = ws;
  p->counters = counters;
  p->max_ctas = max_ctas > 0 ? max_ctas : num_sms();
  p->seg_table = nullptr;
  p->max_ntt = 0;
  p->ws_bytes = gemm_workspace_bytes(p->max_ctas);
  if (K % 8 != 0 || ldw % 8 != 0) return -5;
  return make_tmap(&p->tm_w, W, N, K, ldw, kSlab);
}

// Generic 2-D bf16 map: row-major [rows, cols] with leading dimension ld (elements), box = box_rows x box_cols,
// 128-byte swizzle when box_cols * 2 == 128 and swizzle128 != 0.
int tmap_encode_bf16_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, int box_rows,
                        int box_cols, int swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// 3-D bf16 map [d2][d1][d0] (d0 innermost, contiguous): strides in elements, box b2 x b1 x b0, 128-byte swizzle (b0 * 2 == 128).
int tmap_encode_bf16_3d(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1, uint64_t stride2,
                        int b0, int b1, int b2) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1 * 2, stride2 * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(b0), static_cast<cuuint32_t>(b1), static_cast<cuuint32_t>(b2)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// Output map for the TMA-store epilogue of the pair kernel: out row-major [rows, N] bf16, box = 32 rows x 128 cols,
// no swizzle (the staging tile in smem is plain row-major); rows/cols outside the tensor are clipped by the hardware.
int gemm_make_out_map(CUtensorMap* tm, const void* out, int rows, int N, int ldo) {
  if (ldo % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0) return -5;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldo) * 2};
  cuuint32_t box[2] = {128, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(out), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

int gemm_make_x_map(CUtensorMap* tm, const void* X, int rows, int K, int ldx, int block_n) {
  if (ldx % 8 != 0) return -5;
  return make_tmap(tm, X, rows, K, ldx, gemm_x_box_rows(block_n));
}

int gemm_run(const GemmPlan& p, const CUtensorMap& tm_x, int block_n, void* out, int ldo, int T,
             cudaStream_t st) {
  if (T <= 0) return 0;
  if (g_variant == 2) return gemm2_run(p, tm_x, block_n, out, ldo, T, st);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  switch (block_n) {
    case 32: return launch<32>(p, tm_x, o, ldo, T, st);
    case 64: return launch<64>(p, tm_x, o, ldo, T, st);
    case 128: return launch<128>(p, tm_x, o, ldo, T, st);
    case 256: return launch<256>(p, tm_x, o, ldo, T, st);
    default: return -6;
  }
}

}  // namespace b200
