// Scratch latency micro-benchmarks (single warp): dependent chains of SHFL, STS->LDS, LOP3.
#include <cstdio>
#include <cstdint>
#define N 4096
__global__ void k_shfl(unsigned long long *out, uint32_t v) {
  unsigned long long c0 = clock64();
  int src = (threadIdx.x + 7) & 31;
  for (int i = 0; i < N; i++) v = __shfl_sync(0xffffffffu, v, src) ^ 0x9e37u;
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = v;
}
__global__ void k_shfl2(unsigned long long *out, uint32_t v) {  // two independent shuffles then combine
  unsigned long long c0 = clock64();
  int src = (threadIdx.x + 7) & 31, src2 = (threadIdx.x + 11) & 31;
  uint32_t w = v * 3;
  for (int i = 0; i < N; i++) { uint32_t a = __shfl_sync(0xffffffffu, v, src), b = __shfl_sync(0xffffffffu, w, src2); v = a ^ b; w = a + b; }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = v + w;
}
__global__ void k_smem32(unsigned long long *out, uint32_t v) {
  __shared__ uint32_t s[64];
  unsigned long long c0 = clock64();
  uint32_t st = (uint32_t)__cvta_generic_to_shared(&s[threadIdx.x]), ld = (uint32_t)__cvta_generic_to_shared(&s[(threadIdx.x + 7) & 31]);
  for (int i = 0; i < N; i++) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(st), "r"(v) : "memory");
    __syncwarp();
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ld) : "memory");
    v ^= 0x9e37u;
  }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = v;
}
__global__ void k_smem64(unsigned long long *out, uint32_t v) {
  __shared__ __align__(16) uint32_t s[128];
  unsigned long long c0 = clock64();
  uint32_t st = (uint32_t)__cvta_generic_to_shared(&s[2 * threadIdx.x]), ld = (uint32_t)__cvta_generic_to_shared(&s[2 * ((threadIdx.x + 7) & 31)]);
  uint32_t w = v * 3;
  for (int i = 0; i < N; i++) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(st), "r"(v), "r"(w) : "memory");
    __syncwarp();
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(w) : "r"(ld) : "memory");
    v ^= 0x9e37u;
  }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = v + w;
}
__global__ void k_smem_nosync(unsigned long long *out, uint32_t v) {  // same lane reads back its own slot: no syncwarp
  __shared__ uint32_t s[64];
  unsigned long long c0 = clock64();
  uint32_t st = (uint32_t)__cvta_generic_to_shared(&s[threadIdx.x]);
  for (int i = 0; i < N; i++) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(st), "r"(v) : "memory");
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(st) : "memory");
    v ^= 0x9e37u;
  }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = v;
}
__global__ void k_lds_only(unsigned long long *out, uint32_t v) {  // pointer chase in smem
  __shared__ uint32_t s[64];
  s[threadIdx.x] = (uint32_t)__cvta_generic_to_shared(&s[(threadIdx.x + 7) & 31]);
  __syncwarp();
  uint32_t p = (uint32_t)__cvta_generic_to_shared(&s[threadIdx.x]);
  unsigned long long c0 = clock64();
  for (int i = 0; i < N; i++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(p) : "r"(p) : "memory");
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = p + v;
}
__global__ void k_lop(unsigned long long *out, uint32_t v) {
  unsigned long long c0 = clock64();
  uint32_t a = v, b = v * 5, c = v * 7;
#pragma unroll 16
  for (int i = 0; i < N; i++) { a = a ^ (~b & c); b = __funnelshift_l(a, b, 7); }
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = a + b;
}
__global__ void k_redux(unsigned long long *out, uint32_t v) {
  unsigned long long c0 = clock64();
  for (int i = 0; i < N; i++) v = __reduce_xor_sync(0xffffffffu, v + threadIdx.x) ;
  unsigned long long c1 = clock64(); out[0] = c1 - c0; out[1] = v;
}
int main() {
  unsigned long long *d, h[2]; cudaMalloc(&d, 16);
#define RUN(K, per) for (int r = 0; r < 2; r++) { K<<<1, 32>>>(d, 12345u); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); if (r) printf("%-14s %.1f cycles/iter  (%s)\n", #K, double(h[0]) / N, per); }
  RUN(k_shfl, "shfl + lop3");
  RUN(k_shfl2, "2 parallel shfl + alu");
  RUN(k_smem32, "sts32 + syncwarp + lds32 + lop3");
  RUN(k_smem64, "sts64 + syncwarp + lds64 + lop3");
  RUN(k_smem_nosync, "sts32 + lds32 same lane + lop3");
  RUN(k_lds_only, "lds32 pointer chase");
  RUN(k_lop, "lop3 + shf dependent pair");
  RUN(k_redux, "redux.xor + iadd");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
