// Micro-benchmark (design input for the RoIAlign backward rewrite): throughput of fp32 shared-memory atomic adds
// vs global red.add on sm_100a, conflict-free and with 2..4 lanes per address.
// build+run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/atoms benchmarks/micro/atoms.cu && /tmp/atoms
#include <cstdio>
#include <cuda_runtime.h>

template <int kShare>  // kShare lanes hit the same address
__global__ void __launch_bounds__(512) smem_atoms(float* out, int iters) {
  extern __shared__ float s[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) s[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned idx = warp * 997 + lane / kShare;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      atomicAdd(&s[(idx + u * 37) & 16383], 1.0f);
    }
    idx += 331;
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = s[0] + s[5];
}

__global__ void __launch_bounds__(512) gmem_reds(float* g, size_t n, int iters) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t idx = (tid * 7) % n;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) atomicAdd(&g[(idx + u * 4099) % n], 1.0f);
    idx = (idx + 1000003) % n;
  }
}

__global__ void __launch_bounds__(512) gmem_reds_coalesced(float* g, size_t n, int iters) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t idx = tid % n;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) atomicAdd(&g[(idx + (size_t)u * 1048576) % n], 1.0f);
    idx = (idx + 75776) % n;
  }
}

int main() {
  float *out, *g;
  const size_t n = 64u << 20;
  cudaMalloc(&out, 4096);
  cudaMalloc(&g, n * 4);
  cudaMemset(g, 0, n * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int iters = 2000;
  auto run = [&](const char* name, auto launch, double ops) {
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("%-28s %8.3f ms  %8.1f G lane-atomics/s\n", name, ms, ops / ms / 1e6);
  };
  const double ops = 148.0 * 512 * iters * 8;
  cudaFuncSetAttribute(smem_atoms<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(smem_atoms<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(smem_atoms<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  run("smem atomicAdd f32 spread", [&] { smem_atoms<1><<<148, 512, 65536>>>(out, iters); }, ops);
  run("smem atomicAdd f32 2/addr", [&] { smem_atoms<2><<<148, 512, 65536>>>(out, iters); }, ops);
  run("smem atomicAdd f32 4/addr", [&] { smem_atoms<4><<<148, 512, 65536>>>(out, iters); }, ops);
  run("global red scattered", [&] { gmem_reds<<<148 * 4, 512>>>(g, n, iters / 4); }, ops);
  run("global red coalesced", [&] { gmem_reds_coalesced<<<148 * 4, 512>>>(g, n, iters / 4); }, ops);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
