// Scratch: where do the cycles of one smem-exchange Keccak round go?
#include <cstdio>
#include <cstdint>
#define N 4096
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) { asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory"); }
__device__ __forceinline__ void lds64(uint32_t addr, uint32_t &a, uint32_t &b) { asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory"); }

// exchange 1 only: sts64 own, 4x lds128 + 2x lds64 of two columns, xor tree -> feeds next store
template <int MODE>
__global__ void k_ex1(unsigned long long *out, uint32_t v) {
  __shared__ __align__(16) uint8_t smem[1024];
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  int lane = threadIdx.x, t = lane < 25 ? lane : lane - 25, x = t % 5, y = t / 5;
  uint32_t own = base + (lane < 25 ? x * 48 + y * 8 : 512 + lane * 8);
  uint32_t cm = base + ((x + 4) % 5) * 48, cp = base + ((x + 1) % 5) * 48;
  uint32_t lo = v * (lane + 1), hi = v * 7 + lane;
  unsigned long long c0 = clock64();
  for (int i = 0; i < N; i++) {
    sts64(own, lo, hi);
    __syncwarp();
    uint32_t m[10], p[10];
    if (MODE == 0) {
      lds128(cm, m[0], m[1], m[2], m[3]); lds128(cp, p[0], p[1], p[2], p[3]);
      lds128(cm + 16, m[4], m[5], m[6], m[7]); lds128(cp + 16, p[4], p[5], p[6], p[7]);
      lds64(cm + 32, m[8], m[9]); lds64(cp + 32, p[8], p[9]);
    } else if (MODE == 1) {  // all 64-bit loads
#pragma unroll
      for (int j = 0; j < 5; j++) { lds64(cm + 8 * j, m[2 * j], m[2 * j + 1]); lds64(cp + 8 * j, p[2 * j], p[2 * j + 1]); }
    } else {  // only one column (5 values)
      lds128(cm, m[0], m[1], m[2], m[3]); lds128(cm + 16, m[4], m[5], m[6], m[7]); lds64(cm + 32, m[8], m[9]);
#pragma unroll
      for (int j = 0; j < 10; j++) p[j] = m[j] * 3;
    }
    uint32_t cml = (m[0] ^ m[2] ^ m[4]) ^ m[6] ^ m[8], cmh = (m[1] ^ m[3] ^ m[5]) ^ m[7] ^ m[9];
    uint32_t cpl = (p[0] ^ p[2] ^ p[4]) ^ p[6] ^ p[8], cph = (p[1] ^ p[3] ^ p[5]) ^ p[7] ^ p[9];
    lo = lo ^ cml ^ __funnelshift_l(cph, cpl, 1);
    hi = hi ^ cmh ^ __funnelshift_l(cpl, cph, 1);
  }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = lo + hi;
}
// exchange 2 only: 2x sts32, 3x lds64, chi
__global__ void k_ex2(unsigned long long *out, uint32_t v) {
  __shared__ __align__(16) uint8_t smem[1024];
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  int lane = threadIdx.x, t = lane < 25 ? lane : lane - 25, x = t % 5, y = t / 5;
  int X = y, Y = (2 * x + 3 * y) % 5;
  uint32_t dst = base + (lane < 25 ? (X + 5 * Y) * 8 : 512 + lane * 8);
  uint32_t b0 = base + (x + 5 * y) * 8, b1 = base + ((x + 1) % 5 + 5 * y) * 8, b2 = base + ((x + 2) % 5 + 5 * y) * 8;
  uint32_t lo = v * (lane + 1), hi = v * 7 + lane, rot = lane & 31;
  unsigned long long c0 = clock64();
  for (int i = 0; i < N; i++) {
    sts32(dst, __funnelshift_l(hi, lo, rot));
    sts32(dst + 4, __funnelshift_l(lo, hi, rot));
    __syncwarp();
    uint32_t b0l, b0h, b1l, b1h, b2l, b2h;
    lds64(b0, b0l, b0h); lds64(b1, b1l, b1h); lds64(b2, b2l, b2h);
    lo = b0l ^ (~b1l & b2l); hi = b0h ^ (~b1h & b2h);
  }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = lo + hi;
}
// shuffle versions of the same two stages
__global__ void k_sh2(unsigned long long *out, uint32_t v) {
  int lane = threadIdx.x, t = lane < 25 ? lane : lane - 25, x = t % 5, y = t / 5;
  int s0 = (x + 3 * y) % 5 + 5 * x, x1 = (x + 1) % 5, x2 = (x + 2) % 5;
  int s1 = (x1 + 3 * y) % 5 + 5 * x1, s2 = (x2 + 3 * y) % 5 + 5 * x2;
  uint32_t lo = v * (lane + 1), hi = v * 7 + lane, rot = lane & 31;
  unsigned long long c0 = clock64();
  for (int i = 0; i < N; i++) {
    uint32_t rl = __funnelshift_l(hi, lo, rot), rh = __funnelshift_l(lo, hi, rot);
    uint32_t b0l = __shfl_sync(~0u, rl, s0), b0h = __shfl_sync(~0u, rh, s0);
    uint32_t b1l = __shfl_sync(~0u, rl, s1), b1h = __shfl_sync(~0u, rh, s1);
    uint32_t b2l = __shfl_sync(~0u, rl, s2), b2h = __shfl_sync(~0u, rh, s2);
    lo = b0l ^ (~b1l & b2l); hi = b0h ^ (~b1h & b2h);
  }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = lo + hi;
}
int main() {
  unsigned long long *d, h[2]; cudaMalloc(&d, 16);
#define RUN(K, per) for (int r = 0; r < 2; r++) { K<<<1, 32>>>(d, 12345u); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); if (r) printf("%-12s %.1f cycles/iter  (%s)\n", #K, double(h[0]) / N, per); }
  RUN(k_ex1<0>, "exchange1: sts64, 4 lds128 + 2 lds64, theta alu");
  RUN(k_ex1<1>, "exchange1 with 10 lds64");
  RUN(k_ex1<2>, "exchange1 one column only (2 lds128 + lds64)");
  RUN(k_ex2, "exchange2: 2 sts32, 3 lds64, chi");
  RUN(k_sh2, "shuffle chi stage: 6 shfl");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
