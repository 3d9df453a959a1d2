// Probe: where do the rows of D land in tensor memory for tcgen05.mma.cta_group::2 with M = 128 (64 rows per CTA)?
// A[r][0] = r, A[r][1] = 1, B[n][0] = 1, B[n][1] = n / 256  =>  D[r][n] = r + n / 256 exactly (fp32 accumulate).
// Each CTA dumps TMEM lanes 0..127 x columns 0..N-1 (tcgen05.ld 32x32b) so the (row, column) -> (cta, lane, column) map
// can be read off.  Operands are built by hand in the canonical K-major 128B-swizzle layout (no TMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/exp/umma_layout_probe scripts/exp/umma_layout_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int M, int N>
__global__ void __cluster_dims__(2, 1, 1) probe(float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_s = base, b_s = base + 16384, bar = base + 32768, slot = bar + 16;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int MR = M / 2, NR = N / 2;   // rows of A / B held by this CTA
  for (int i = tid; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  // element (row r, k) of a K-major SW128 tile: byte r*128 + ((k/8) ^ (r%8))*16 + (k%8)*2
  for (int r = tid; r < MR; r += blockDim.x) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(smem + r * 128 + ((0 ^ (r & 7)) << 4));
    p[0] = __float2bfloat16(static_cast<float>(rank * MR + r));
    p[1] = __float2bfloat16(1.0f);
  }
  for (int n = tid; n < NR; n += blockDim.x) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(smem + 16384 + n * 128 + ((0 ^ (n & 7)) << 4));
    p[0] = __float2bfloat16(1.0f);
    p[1] = __float2bfloat16(static_cast<float>(rank * NR + n) / 256.0f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 32768 + 16);
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}\n" ::"r"(tmem),
        "l"(desc_sw128(a_s)), "l"(desc_sw128(b_s)), "r"(idesc), "r"(0u), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
  }
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // dump lanes 32*(warp%4).. x N columns
  const int q = warp & 3, lane = tid & 31;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(static_cast<size_t>(rank) * 128 + q * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}

template <int M, int N>
void run() {
  float* d;
  cudaMalloc(&d, 2 * 128 * N * 4);
  cudaMemset(d, 0xff, 2 * 128 * N * 4);
  cudaFuncSetAttribute(probe<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  probe<M, N><<<2, 128, 40000>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("M=%d N=%d: %s\n", M, N, cudaGetErrorString(e));
  if (e != cudaSuccess) return;
  std::vector<float> h(2 * 128 * N);
  cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  // for every (cta, lane): which row, and columns pattern of the first 4 + middle entries
  for (int cta = 0; cta < 2; ++cta)
    for (int lane = 0; lane < 128; lane += (lane < 4 || (lane >= 14 && lane < 18) || (lane >= 30 && lane < 34) || (lane >= 62 && lane < 66) || lane >= 126) ? 1 : 1) {
      const float* p = &h[(cta * 128 + lane) * N];
      auto dec = [&](float v) { int r = static_cast<int>(v); int n = static_cast<int>((v - r) * 256.0f + 0.5f); return r * 1000 + n; };
      if (lane % 8 == 0 || lane % 8 == 7)
        printf(" cta %d lane %3d: col0 -> (row,col)=%6d  col1 -> %6d  col%d -> %6d  col%d -> %6d\n", cta, lane, dec(p[0]), dec(p[1]), N / 2, dec(p[N / 2]),
               N - 1, dec(p[N - 1]));
    }
}

int main() {
  run<256, 64>();
  run<128, 64>();
  return 0;
}
