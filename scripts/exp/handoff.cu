// Kernel-to-kernel handoff latency: programmatic dependent launch (griddepcontrol.wait on the producer grid) versus a
// software flag (producer CTAs bump a counter after their last store, consumer CTAs poll it).  Chain of N tiny kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o handoff handoff.cu && ./handoff
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void k_pdl(float* buf, int i) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  buf[blockIdx.x * blockDim.x + threadIdx.x] += 1.0f;
}
__global__ void k_flag(float* buf, unsigned long long* ctr, unsigned long long target) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
  buf[blockIdx.x * blockDim.x + threadIdx.x] += 1.0f;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(ctr, 1ull);
}
__global__ void k_plain(float* buf) { buf[blockIdx.x * blockDim.x + threadIdx.x] += 1.0f; }

template <typename... A, typename... B>
void launch(void (*k)(A...), int grid, int block, cudaStream_t st, bool pdl, B... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k, static_cast<A>(args)...);
}

int main() {
  const int N = 400, grid = 128, block = 512;
  float* buf;
  unsigned long long* ctr;
  cudaMalloc(&buf, grid * block * 4);
  cudaMemset(buf, 0, grid * block * 4);
  cudaMalloc(&ctr, 8);
  cudaMemset(ctr, 0, 8);
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, st);
    for (int i = 0; i < N; ++i) launch(k_plain, grid, block, st, false, buf);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("plain stream order : %.2f us per kernel\n", ms * 1e3 / N);
    cudaEventRecord(e0, st);
    for (int i = 0; i < N; ++i) launch(k_pdl, grid, block, st, true, buf, i);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("PDL wait/launch    : %.2f us per kernel\n", ms * 1e3 / N);
    cudaMemsetAsync(ctr, 0, 8, st);
    cudaEventRecord(e0, st);
    for (int i = 0; i < N; ++i) launch(k_flag, grid, block, st, true, buf, ctr, (unsigned long long)i * grid);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("flag poll handoff  : %.2f us per kernel  (%s)\n", ms * 1e3 / N, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
