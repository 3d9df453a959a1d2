// Probe for the cluster split-K design: how many clusters of each size are co-resident on this B200 with a GEMM-sized
// CTA (one per SM), and how fast a CTA can push an fp32 partial tile into a peer's shared memory (DSMEM), by plain
// st.shared::cluster.v4 from 128 threads and by cp.async.bulk smem->remote smem.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/exp/cluster_probe scripts/exp/cluster_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

extern __shared__ __align__(1024) uint8_t smem[];

__global__ void dummy_kernel(int* out) {
  if (threadIdx.x == 0) out[blockIdx.x] = 1;
}

// every CTA pushes `bytes` to the next CTA of its cluster (ring), mode 0: st.shared::cluster.v4, mode 1: cp.async.bulk
__global__ void dsmem_kernel(int bytes, int mode, long long* ns_out, unsigned* smid_out, int reps) {
  const uint32_t rank = cluster_rank(), n = cluster_size();
  const uint32_t recv = smem_u32(smem);                 // [0, bytes): receive buffer
  const uint32_t send = recv + 96 * 1024;               // local staging for the bulk copy
  const uint32_t bar = recv + 200 * 1024;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smid_out[blockIdx.x] = s;
  }
  float4 v = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
  for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x)
    reinterpret_cast<float4*>(smem + 96 * 1024)[i] = v;
  __syncthreads();
  cluster_sync();
  const uint32_t peer = (rank + 1) % n;
  const uint32_t dst = mapa(recv, peer), dst_bar = mapa(bar, peer);
  long long t0 = gtime();
  for (int r = 0; r < reps; ++r) {
    if (mode == 0) {
      for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x)
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + i * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      cluster_sync();
    } else {
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      }
      cluster_sync();  // everyone armed their barrier
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "r"(send), "r"(bytes), "r"(dst_bar) : "memory");
      }
      // wait for MY buffer to be filled by my predecessor
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(r & 1) : "memory");
      }
      cluster_sync();
    }
  }
  long long t1 = gtime();
  if (threadIdx.x == 0) ns_out[blockIdx.x] = t1 - t0;
  cluster_sync();
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  printf("%s SMs %d\n", prop.name, prop.multiProcessorCount);
  int* d_out; cudaMalloc(&d_out, 4096 * 4);
  long long* d_ns; cudaMalloc(&d_ns, 4096 * 8);
  unsigned* d_sm; cudaMalloc(&d_sm, 4096 * 4);
  for (int smem_kb : {100, 160, 208, 224}) {
    cudaFuncSetAttribute(dummy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    cudaFuncSetAttribute(dummy_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 6, 8, 12, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem_kb * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy_kernel, &cfg);
      printf("smem %3d KB cluster %2d: max active clusters %3d (%4d CTAs) %s\n", smem_kb, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
      cudaGetLastError();
    }
  }
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {2, 4, 6, 8}) {
    for (int mode = 0; mode < 2; ++mode) {
      for (int kb : {16, 48, 64}) {
        int nclusters = cs == 8 ? 16 : cs == 6 ? 20 : cs == 4 ? 32 : 74;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * nclusters); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 208 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const int reps = 20;
        cudaError_t e = cudaLaunchKernelEx(&cfg, dsmem_kernel, kb * 1024, mode, d_ns, d_sm, reps);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cluster %d mode %d %d KB: %s\n", cs, mode, kb, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        std::vector<long long> ns(cs * nclusters);
        cudaMemcpy(ns.data(), d_ns, ns.size() * 8, cudaMemcpyDeviceToHost);
        std::sort(ns.begin(), ns.end());
        double med = ns[ns.size() / 2] / double(reps), mx = ns.back() / double(reps);
        printf("cluster %d x%d %s %2d KB/CTA: median %.0f ns, max %.0f ns per push+sync  -> %.1f GB/s per SM\n", cs, nclusters,
               mode ? "cp.async.bulk" : "st.cluster.v4", kb, med, mx, kb * 1024 / med);
      }
    }
  }
  return 0;
}
