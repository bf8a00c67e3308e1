// tlb_probe.cu — dependent random 160 B block reads over footprints of different sizes: does the latency of the
// step kernels' tree-block loads depend on how much HBM the pool spans (TLB reach)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/tlb_probe.cu -o build/tlb_probe && build/tlb_probe
// Each thread walks `steps` dependent hops; a hop loads ten 16 B vectors of one 160 B block and derives the next block
// index from the data (so hops cannot overlap). `spread` = how many distinct blocks a thread's hops are confined to
// (0 = anywhere in the footprint), i.e. the locality a per-tree arena would give.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void k_fill(uint4* p, size_t n_vec) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long x = i * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    p[i] = make_uint4((unsigned)x, (unsigned)(x >> 32), (unsigned)(x * 3), (unsigned)(x * 7 >> 17));
  }
}
__global__ void k_walk(const uint4* __restrict__ p, size_t n_blocks, size_t region_blocks, int steps, unsigned* out) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  // a thread's home region: region_blocks consecutive blocks (or the whole footprint)
  const size_t span = region_blocks ? region_blocks : n_blocks;
  const size_t base = region_blocks ? (t * (n_blocks / nthreads)) : 0;
  unsigned long long h = t * 0x2545F4914F6CDD1Dull + 12345;
  unsigned acc = 0;
  for (int s = 0; s < steps; ++s) {
    const size_t b = base + (size_t)(h % span);
    const uint4* q = p + b * 10;
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
  out[t] = acc;
}
int main() {
  const size_t sizes_gb[] = {1, 4, 16, 36, 72};
  const int steps = 400;
  for (int threads_per_sm : {448, 1024}) {
    for (size_t gb : sizes_gb) {
      const size_t bytes = gb << 30, n_blocks = bytes / 160, n_vec = n_blocks * 10;
      uint4* p = nullptr;
      if (cudaMalloc(&p, n_vec * sizeof(uint4)) != cudaSuccess) { printf("alloc %zu GB failed\n", gb); continue; }
      unsigned* out; cudaMalloc(&out, 148 * 1024 * 4);
      k_fill<<<148 * 8, 256>>>(p, n_vec);
      cudaDeviceSynchronize();
      for (size_t region : {(size_t)0, (size_t)3456}) {  // anywhere | a 553 KB home region per thread (two trees' arenas)
        const int block = threads_per_sm == 448 ? 448 : 1024;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k_walk<<<148, block>>>(p, n_blocks, region, 50, out);
        cudaEventRecord(a);
        k_walk<<<148, block>>>(p, n_blocks, region, steps, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        const double hops = 148.0 * block * steps;
        printf("{\"threads_per_sm\": %d, \"footprint_gb\": %zu, \"home_region_blocks\": %zu, \"ns_per_dependent_hop\": %.1f, "
               "\"Ghops_per_s\": %.3f, \"GBps\": %.1f}\n", block, gb, region, ms * 1e6 / steps, hops / ms / 1e6, hops * 160 / ms / 1e6);
      }
      cudaFree(p); cudaFree(out);
    }
  }
  return 0;
}
