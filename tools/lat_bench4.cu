// Scratch: candidate exchange layouts for the Keccak round (single warp).
#include <cstdio>
#include <cstdint>
#define N 4096
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) { asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory"); }
__device__ __forceinline__ void lds64(uint32_t addr, uint32_t &a, uint32_t &b) { asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory"); }

// ex2 variant: rows of 8 slots with mirrors, 2 x lds128 + bitwise select
__global__ void k_ex2_wide(unsigned long long *out, uint32_t v) {
  __shared__ __align__(16) uint8_t smem[1024];
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  int lane = threadIdx.x, t = lane < 25 ? lane : lane - 25, x = t % 5, y = t / 5;
  int X = y, Y = (2 * x + 3 * y) % 5;
  uint32_t dst = base + (lane < 25 ? (X + 8 * Y) * 8 : 512 + lane * 8);
  uint32_t dstm = base + ((lane < 25 && X < 2) ? (X + 5 + 8 * Y) * 8 : 768 + lane * 8);
  uint32_t ld = base + ((x & ~1) + 8 * y) * 8;
  uint32_t odd = (x & 1) ? 0xffffffffu : 0u;
  uint32_t lo = v * (lane + 1), hi = v * 7 + lane, rot = lane & 31;
  unsigned long long c0 = clock64();
  for (int i = 0; i < N; i++) {
    uint32_t rl = __funnelshift_l(hi, lo, rot), rh = __funnelshift_l(lo, hi, rot);
    sts32(dst, rl); sts32(dst + 4, rh); sts32(dstm, rl); sts32(dstm + 4, rh);
    __syncwarp();
    uint32_t v0l, v0h, v1l, v1h, v2l, v2h, v3l, v3h;
    lds128(ld, v0l, v0h, v1l, v1h); lds128(ld + 16, v2l, v2h, v3l, v3h);
    uint32_t b0l = (v1l & odd) | (v0l & ~odd), b0h = (v1h & odd) | (v0h & ~odd);
    uint32_t b1l = (v2l & odd) | (v1l & ~odd), b1h = (v2h & odd) | (v1h & ~odd);
    uint32_t b2l = (v3l & odd) | (v2l & ~odd), b2h = (v3h & odd) | (v2h & ~odd);
    lo = b0l ^ (~b1l & b2l); hi = b0h ^ (~b1h & b2h);
  }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = lo + hi;
}
// ex1 variant: two-step theta. step a: column replicated 12 slots, 2 x lds128 give the other four; step b: parity exchange 2 x lds64
__global__ void k_ex1_two(unsigned long long *out, uint32_t v) {
  __shared__ __align__(16) uint8_t smem[2048];
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  int lane = threadIdx.x, t = lane < 25 ? lane : lane - 25, x = t % 5, y = t / 5;
  // column x region: 12 slots * 8 B = 96 B
  uint32_t col = base + x * 96;
  uint32_t st0 = lane < 25 ? col + y * 8 : base + 1024 + lane * 8;
  uint32_t st1 = lane < 25 ? col + (y + 5) * 8 : base + 1280 + lane * 8;
  uint32_t st2 = (lane < 25 && y < 2) ? col + (y + 10) * 8 : base + 1536 + lane * 8;
  const int start[5] = {6, 2, 8, 4, 0};  // window of 4 slots that misses own y
  uint32_t ld = col + start[y] * 8;
  uint32_t pbase = base + 512;  // parities: 5 * 8 B
  uint32_t pst = lane < 25 ? pbase + x * 8 : base + 1792 + lane * 8;
  uint32_t pm = pbase + ((x + 4) % 5) * 8, pp = pbase + ((x + 1) % 5) * 8;
  uint32_t lo = v * (lane + 1), hi = v * 7 + lane;
  unsigned long long c0 = clock64();
  for (int i = 0; i < N; i++) {
    sts64(st0, lo, hi); sts64(st1, lo, hi); sts64(st2, lo, hi);
    __syncwarp();
    uint32_t a0, a1, a2, a3, a4, a5, a6, a7;
    lds128(ld, a0, a1, a2, a3); lds128(ld + 16, a4, a5, a6, a7);
    uint32_t cl = (lo ^ a0 ^ a2) ^ a4 ^ a6, ch = (hi ^ a1 ^ a3) ^ a5 ^ a7;
    sts64(pst, cl, ch);
    __syncwarp();
    uint32_t ml, mh, pl, ph;
    lds64(pm, ml, mh); lds64(pp, pl, ph);
    lo = lo ^ ml ^ __funnelshift_l(ph, pl, 1);
    hi = hi ^ mh ^ __funnelshift_l(pl, ph, 1);
  }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = lo + hi;
}
// ex1 variant: step b by shuffles instead of smem
__global__ void k_ex1_two_shfl(unsigned long long *out, uint32_t v) {
  __shared__ __align__(16) uint8_t smem[2048];
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  int lane = threadIdx.x, t = lane < 25 ? lane : lane - 25, x = t % 5, y = t / 5;
  uint32_t col = base + x * 96;
  uint32_t st0 = lane < 25 ? col + y * 8 : base + 1024 + lane * 8;
  uint32_t st1 = lane < 25 ? col + (y + 5) * 8 : base + 1280 + lane * 8;
  uint32_t st2 = (lane < 25 && y < 2) ? col + (y + 10) * 8 : base + 1536 + lane * 8;
  const int start[5] = {6, 2, 8, 4, 0};
  uint32_t ld = col + start[y] * 8;
  int lm = (x + 4) % 5 + 5 * y, lp = (x + 1) % 5 + 5 * y;
  uint32_t lo = v * (lane + 1), hi = v * 7 + lane;
  unsigned long long c0 = clock64();
  for (int i = 0; i < N; i++) {
    sts64(st0, lo, hi); sts64(st1, lo, hi); sts64(st2, lo, hi);
    __syncwarp();
    uint32_t a0, a1, a2, a3, a4, a5, a6, a7;
    lds128(ld, a0, a1, a2, a3); lds128(ld + 16, a4, a5, a6, a7);
    uint32_t cl = (lo ^ a0 ^ a2) ^ a4 ^ a6, ch = (hi ^ a1 ^ a3) ^ a5 ^ a7;
    uint32_t ml = __shfl_sync(~0u, cl, lm), mh = __shfl_sync(~0u, ch, lm), pl = __shfl_sync(~0u, cl, lp), ph = __shfl_sync(~0u, ch, lp);
    lo = lo ^ ml ^ __funnelshift_l(ph, pl, 1);
    hi = hi ^ mh ^ __funnelshift_l(pl, ph, 1);
  }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = lo + hi;
}
// ex1 variant: records [col x-1 | col x+1] contiguous 80 B -> 5 x lds128
__global__ void k_ex1_rec(unsigned long long *out, uint32_t v) {
  __shared__ __align__(16) uint8_t smem[2048];
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  int lane = threadIdx.x, t = lane < 25 ? lane : lane - 25, x = t % 5, y = t / 5;
  // record of reader column c at c*80: slots 0..4 = column c-1, 5..9 = column c+1
  uint32_t stA = lane < 25 ? base + ((x + 1) % 5) * 80 + y * 8 : base + 1024 + lane * 8;          // I am column c-1 of reader c = x+1
  uint32_t stB = lane < 25 ? base + ((x + 4) % 5) * 80 + (5 + y) * 8 : base + 1280 + lane * 8;    // I am column c+1 of reader c = x-1
  uint32_t ld = base + x * 80;
  uint32_t lo = v * (lane + 1), hi = v * 7 + lane;
  unsigned long long c0 = clock64();
  for (int i = 0; i < N; i++) {
    sts64(stA, lo, hi); sts64(stB, lo, hi);
    __syncwarp();
    uint32_t m[10], p[10];
    lds128(ld, m[0], m[1], m[2], m[3]); lds128(ld + 16, m[4], m[5], m[6], m[7]); lds128(ld + 32, m[8], m[9], p[0], p[1]);
    lds128(ld + 48, p[2], p[3], p[4], p[5]); lds128(ld + 64, p[6], p[7], p[8], p[9]);
    uint32_t cml = (m[0] ^ m[2] ^ m[4]) ^ m[6] ^ m[8], cmh = (m[1] ^ m[3] ^ m[5]) ^ m[7] ^ m[9];
    uint32_t cpl = (p[0] ^ p[2] ^ p[4]) ^ p[6] ^ p[8], cph = (p[1] ^ p[3] ^ p[5]) ^ p[7] ^ p[9];
    lo = lo ^ cml ^ __funnelshift_l(cph, cpl, 1);
    hi = hi ^ cmh ^ __funnelshift_l(cpl, cph, 1);
  }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = lo + hi;
}
int main() {
  unsigned long long *d, h[2]; cudaMalloc(&d, 16);
#define RUN(K, per) for (int r = 0; r < 2; r++) { K<<<1, 32>>>(d, 12345u); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); if (r) printf("%-16s %.1f cycles/iter  (%s)\n", #K, double(h[0]) / N, per); }
  RUN(k_ex2_wide, "chi exchange: 4 sts32, 2 lds128, select, chi   [was 72]");
  RUN(k_ex1_two, "theta two-step smem: 3 sts64, 2 lds128, sts64, 2 lds64   [one-step was 92]");
  RUN(k_ex1_two_shfl, "theta two-step: smem column gather + 4 shfl");
  RUN(k_ex1_rec, "theta one-step, 80 B records, 5 lds128");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
