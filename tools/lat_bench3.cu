// Scratch: smem exchange cost vs number/width of stores and loads, conflict-free vs permuted addressing.
#include <cstdio>
#include <cstdint>
#define N 4096
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }
__device__ __forceinline__ void lds64(uint32_t addr, uint32_t &a, uint32_t &b) { asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory"); }
__device__ __forceinline__ void lds32(uint32_t addr, uint32_t &a) { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(addr) : "memory"); }

template <int NST32, int NLD, bool PERM, bool W64>
__global__ void k(unsigned long long *out, uint32_t v) {
  __shared__ __align__(16) uint8_t smem[2048];
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  int lane = threadIdx.x;
  int slot = PERM ? (lane * 7 + 3) % 32 : lane;
  uint32_t dst = base + slot * 8;
  uint32_t l0 = base + ((lane + 1) & 31) * 8, l1 = base + ((lane + 2) & 31) * 8, l2 = base + ((lane + 5) & 31) * 8;
  uint32_t lo = v * (lane + 1), hi = v * 7 + lane;
  unsigned long long c0 = clock64();
  for (int i = 0; i < N; i++) {
    if (NST32 == 0) sts64(dst, lo, hi);
    if (NST32 >= 1) sts32(dst, lo);
    if (NST32 >= 2) sts32(dst + 4, hi);
    __syncwarp();
    uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0_ = 0, c1_ = 0;
    if (W64) {
      lds64(l0, a0, a1);
      if (NLD >= 2) lds64(l1, b0, b1);
      if (NLD >= 3) lds64(l2, c0_, c1_);
    } else {
      lds32(l0, a0); lds32(l0 + 4, a1);
      if (NLD >= 2) { lds32(l1, b0); lds32(l1 + 4, b1); }
      if (NLD >= 3) { lds32(l2, c0_); lds32(l2 + 4, c1_); }
    }
    lo = a0 ^ (~b0 & c0_); hi = a1 ^ (~b1 & c1_);
  }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = lo + hi;
}
int main() {
  unsigned long long *d, h[2]; cudaMalloc(&d, 16);
#define RUN(...) for (int r = 0; r < 2; r++) { k<__VA_ARGS__><<<1, 32>>>(d, 12345u); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); if (r) printf("%-24s %.1f cycles/iter\n", #__VA_ARGS__, double(h[0]) / N); }
  printf("<#sts32 (0 = one sts64), #loads, permuted slots, 64-bit loads>\n");
  RUN(0, 1, false, true); RUN(0, 2, false, true); RUN(0, 3, false, true);
  RUN(2, 1, false, true); RUN(2, 3, false, true);
  RUN(0, 3, true, true); RUN(2, 3, true, true);
  RUN(1, 1, false, false); RUN(2, 3, false, false); RUN(0, 3, false, false);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
