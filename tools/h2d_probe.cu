// Scratch: pinned host -> device copy rate vs copy size, back-to-back on one stream (cudaMemcpyAsync), and with 2/4 streams.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <chrono>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  const size_t total = 1ull << 30;
  char *h, *d;
  cudaHostAlloc(&h, total, cudaHostAllocDefault);
  cudaMalloc(&d, total);
  for (size_t i = 0; i < total; i += 4096) h[i] = 1;
  cudaStream_t s[4];
  for (auto &x : s) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
  for (int pass = 0; pass < 2; pass++)
    for (size_t sz : {1ull << 20, 4718600ull, 16ull << 20, 75ull << 20, 256ull << 20, 1ull << 30}) {
      for (int ns : {1, 2, 4}) {
        const size_t n = total / sz;
        cudaDeviceSynchronize();
        double t = now();
        for (size_t i = 0; i < n; i++) cudaMemcpyAsync(d + i * sz, h + i * sz, sz, cudaMemcpyHostToDevice, s[i % ns]);
        cudaDeviceSynchronize();
        double dt = now() - t;
        if (pass) printf("H2D size %8.2f MB, %d stream(s): %6.1f GB/s\n", sz / 1e6, ns, n * sz / dt / 1e9);
      }
    }
  // same region copied repeatedly (the e2e bench reuses its 16 query buffers)
  for (size_t sz : {4718600ull, 75ull << 20}) {
    cudaDeviceSynchronize();
    double t = now();
    for (int i = 0; i < 200; i++) cudaMemcpyAsync(d, h, sz, cudaMemcpyHostToDevice, s[0]);
    cudaDeviceSynchronize();
    printf("H2D same %8.2f MB buffer x200: %6.1f GB/s\n", sz / 1e6, 200 * sz / (now() - t) / 1e9);
  }
  return 0;
}
