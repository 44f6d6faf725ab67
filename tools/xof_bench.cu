// Scratch micro-benchmark (not product code): cycles and nanoseconds per Keccak-p[1600,12] of the device expander.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/xof_bench tools/xof_bench.cu
#include <cstdio>
#include <string>
#include "../chalametpir_b200/csrc/expand.cu"
namespace chpir { void set_last_cuda_error(cudaError_t, const char *) {} }

__global__ void clock_probe(unsigned long long *out, int iters) {
  unsigned long long t0, t1, c0, c1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  c0 = clock64();
  unsigned x = threadIdx.x;
  for (int i = 0; i < iters; i++) x = x * 1664525u + 1013904223u;
  c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  out[0] = c1 - c0; out[1] = t1 - t0; out[2] = x;
}

int main() {
  uint8_t seed[32]; for (int i = 0; i < 32; i++) seed[i] = i;
  uint8_t *out, *scratch; unsigned long long *probe;
  const uint64_t total = 256ull << 20;
  cudaMalloc(&out, total); cudaMalloc(&scratch, 512); cudaMalloc(&probe, 64);
  for (int rep = 0; rep < 3; rep++) {
    clock_probe<<<1, 32>>>(probe, 1 << 22);
    unsigned long long h[3]; cudaMemcpy(h, probe, 24, cudaMemcpyDeviceToHost);
    printf("clock probe: %llu cycles in %llu ns -> %.0f MHz\n", h[0], h[1], h[0] * 1e3 / h[1]);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    chpir::launch_expand(seed, out, total, scratch, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double perms = total / 168.0;
    printf("expand u32: %.1f ms for %.0f permutations -> %.1f ns/perm, %.1f MB/s\n", ms, perms, ms * 1e6 / perms, total / ms / 1e3);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
