// tlb_probe2.cu — latency / throughput of dependent random 160 B block reads when every SM (one CTA per SM) stays inside
// its own contiguous region of `region_mb` MB, for several region sizes and thread counts: how large may a pool
// region of the self-play engine be before the SM's TLB stops covering it, and what is the unloaded latency?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/tlb_probe2.cu -o build/tlb_probe2
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_fill(uint4* p, size_t n_vec) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long x = i * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    p[i] = make_uint4((unsigned)x, (unsigned)(x >> 32), (unsigned)(x * 3), (unsigned)(x * 7 >> 17));
  }
}
__global__ void k_walk(const uint4* __restrict__ p, size_t region_blocks, int steps, unsigned* out) {
  const size_t base = (size_t)blockIdx.x * region_blocks;
  unsigned long long h = (blockIdx.x * 1024ull + threadIdx.x) * 0x2545F4914F6CDD1Dull + 12345;
  unsigned acc = 0;
  for (int s = 0; s < steps; ++s) {
    const uint4* q = p + (base + (size_t)(h % region_blocks)) * 10;
    uint4 v[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) v[j] = q[j];
    unsigned x = 0;
#pragma unroll
    for (int j = 0; j < 10; ++j) x ^= v[j].x + v[j].y * 3u + v[j].z * 5u + v[j].w * 7u;
    acc += x;
    h = h * 6364136223846793005ull + x + 1442695040888963407ull;
    h ^= h >> 31;
  }
  out[blockIdx.x * 1024 + threadIdx.x] = acc;
}
int main() {
  const size_t total_gb = 80;
  const size_t n_vec = (total_gb << 30) / 16;
  uint4* p = nullptr;
  if (cudaMalloc(&p, n_vec * sizeof(uint4)) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  unsigned* out; cudaMalloc(&out, 148 * 1024 * 4);
  k_fill<<<148 * 8, 256>>>(p, n_vec);
  cudaDeviceSynchronize();
  const int steps = 300;
  for (int threads : {32, 128, 448}) {
    for (size_t region_mb : {16, 32, 64, 128, 256, 512}) {
      const size_t region_blocks = (region_mb << 20) / 160;
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      k_walk<<<148, threads>>>(p, region_blocks, 30, out);
      cudaEventRecord(a);
      k_walk<<<148, threads>>>(p, region_blocks, steps, out);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms = 0; cudaEventElapsedTime(&ms, a, b);
      printf("{\"threads_per_sm\": %d, \"region_mb_per_sm\": %zu, \"ns_per_dependent_hop\": %.1f, \"Ghops_per_s\": %.3f}\n",
             threads, region_mb, ms * 1e6 / steps, 148.0 * threads * steps / ms / 1e6);
    }
  }
  return 0;
}
