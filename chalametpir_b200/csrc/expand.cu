// expand.cu -- Matrix::generate_from_seed (chalametpir_common/src/matrix.rs:541-558) on device:
//   TurboShake128::default(); absorb(seed[32]); finalize::<0x1F>(); squeeze(4*rows*cols bytes)
// written straight into the element memory of A (row-major u32, little endian), bit-exactly.
//
// TurboSHAKE128 = Keccak-p[1600, 12 rounds] sponge, rate 168 bytes (RFC 9861; `turboshake` crate =0.4.1 is the
// reference's dependency).  The squeeze is ONE serial chain -- block i+1 is a permutation of block i -- so there is no
// data parallelism across blocks; the only parallelism is inside one permutation.  The kernel therefore runs on a
// single warp and is latency-bound by construction:
//   * lane t = x + 5y (t < 25) owns the 64-bit Keccak lane A[x,y] as two registers;
//   * theta: 4 shuffles fetch the rest of the column, 2 more fetch the neighbour columns' parities;
//   * rho: per-lane constant rotate (funnel shifts);
//   * pi + chi fused: the three chi operands are fetched with 3 shuffles directly from their pre-pi positions;
//   * iota: folded into lane 0.
// After every permutation lanes 0..20 store the 168-byte rate with one coalesced 8-byte-per-lane store.
#include "common.cuh"

namespace chpir {
namespace {

__constant__ uint32_t kRho[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
// round constants of rounds 12..23 of Keccak-f[1600], split (lo, hi)
__constant__ uint32_t kRcLo[12] = {0x8000808bu, 0x0000008bu, 0x00008089u, 0x00008003u, 0x00008002u, 0x00000080u,
                                   0x0000800au, 0x8000000au, 0x80008081u, 0x00008080u, 0x80000001u, 0x80008008u};
__constant__ uint32_t kRcHi[12] = {0x00000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u,
                                   0x00000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0x00000000u, 0x80000000u};

constexpr unsigned kFull = 0xffffffffu;

struct LaneCfg {
  int col1, col2, col3, col4;  // the other four lanes of my column
  int cm, cp;                  // a lane of column x-1 / x+1
  int s0, s1, s2;              // pre-pi sources of B[x,y], B[x+1,y], B[x+2,y]
  uint32_t rot;                // rho amount & 31
  bool swap;                   // rho amount >= 32
  uint32_t iota_mask;          // all ones on lane 0
};

__device__ __forceinline__ LaneCfg make_cfg(int lane) {
  LaneCfg c;
  const int t = lane < 25 ? lane : lane - 25;  // lanes 25..31 shadow lanes 0..6: they only need legal shuffle sources
  const int x = t % 5, y = t / 5;
  c.col1 = (t + 5) % 25, c.col2 = (t + 10) % 25, c.col3 = (t + 15) % 25, c.col4 = (t + 20) % 25;
  c.cm = (x + 4) % 5 + 5 * y;
  c.cp = (x + 1) % 5 + 5 * y;
  const int x1 = (x + 1) % 5, x2 = (x + 2) % 5;
  c.s0 = (x + 3 * y) % 5 + 5 * x;
  c.s1 = (x1 + 3 * y) % 5 + 5 * x1;
  c.s2 = (x2 + 3 * y) % 5 + 5 * x2;
  const uint32_t r = kRho[t];
  c.rot = r & 31u;
  c.swap = r >= 32u;
  c.iota_mask = lane == 0 ? 0xffffffffu : 0u;
  return c;
}

__device__ __forceinline__ void keccak_p12_warp(uint32_t &lo, uint32_t &hi, const LaneCfg &c) {
#pragma unroll
  for (int round = 0; round < 12; round++) {
    // theta
    uint32_t pl = lo ^ __shfl_sync(kFull, lo, c.col1) ^ __shfl_sync(kFull, lo, c.col2);
    uint32_t ph = hi ^ __shfl_sync(kFull, hi, c.col1) ^ __shfl_sync(kFull, hi, c.col2);
    pl ^= __shfl_sync(kFull, lo, c.col3) ^ __shfl_sync(kFull, lo, c.col4);
    ph ^= __shfl_sync(kFull, hi, c.col3) ^ __shfl_sync(kFull, hi, c.col4);
    const uint32_t ml = __shfl_sync(kFull, pl, c.cm), mh = __shfl_sync(kFull, ph, c.cm);
    const uint32_t nl = __shfl_sync(kFull, pl, c.cp), nh = __shfl_sync(kFull, ph, c.cp);
    lo ^= ml ^ __funnelshift_l(nh, nl, 1);
    hi ^= mh ^ __funnelshift_l(nl, nh, 1);
    // rho
    const uint32_t l = c.swap ? hi : lo, h = c.swap ? lo : hi;
    const uint32_t rl = __funnelshift_l(h, l, c.rot), rh = __funnelshift_l(l, h, c.rot);
    // pi + chi
    const uint32_t b0l = __shfl_sync(kFull, rl, c.s0), b0h = __shfl_sync(kFull, rh, c.s0);
    const uint32_t b1l = __shfl_sync(kFull, rl, c.s1), b1h = __shfl_sync(kFull, rh, c.s1);
    const uint32_t b2l = __shfl_sync(kFull, rl, c.s2), b2h = __shfl_sync(kFull, rh, c.s2);
    // iota
    lo = b0l ^ (~b1l & b2l) ^ (kRcLo[round] & c.iota_mask);
    hi = b0h ^ (~b1h & b2h) ^ (kRcHi[round] & c.iota_mask);
  }
}

// state: 25 x u64 (lane-major) + u64 block counter.  One warp.
__global__ void __launch_bounds__(32, 1) expand_kernel(const uint8_t *__restrict__ seed, uint64_t *__restrict__ state, int first,
                                                        uint8_t *__restrict__ out, uint64_t total_bytes, uint64_t block_begin,
                                                        uint64_t block_count) {
  const int lane = threadIdx.x;
  const LaneCfg c = make_cfg(lane);
  uint32_t lo = 0, hi = 0;
  if (first) {
    // absorb: seed -> lanes 0..3; 0x1F at byte 32 (lane 4); 0x80 at byte 167 (lane 20, top byte)
    if (lane < 4) {
      const uint8_t *p = seed + 8 * lane;
      lo = uint32_t(p[0]) | uint32_t(p[1]) << 8 | uint32_t(p[2]) << 16 | uint32_t(p[3]) << 24;
      hi = uint32_t(p[4]) | uint32_t(p[5]) << 8 | uint32_t(p[6]) << 16 | uint32_t(p[7]) << 24;
    } else if (lane == 4) {
      lo = 0x1fu;
    } else if (lane == 20) {
      hi = 0x80000000u;
    }
  } else if (lane < 25) {
    const uint64_t v = state[lane];
    lo = uint32_t(v), hi = uint32_t(v >> 32);
  }
  for (uint64_t blk = block_begin; blk < block_begin + block_count; blk++) {
    keccak_p12_warp(lo, hi, c);
    if (lane < 21) {
      const uint64_t off = blk * 168ull + 8ull * lane;
      if (off + 8 <= total_bytes) {
        *reinterpret_cast<uint2 *>(out + off) = make_uint2(lo, hi);
      } else if (off + 4 <= total_bytes) {
        *reinterpret_cast<uint32_t *>(out + off) = lo;
      }
    }
  }
  if (lane < 25) state[lane] = uint64_t(lo) | uint64_t(hi) << 32;
}

}  // namespace

int launch_expand(const uint8_t seed[32], uint8_t *out_dev, uint64_t total_bytes, uint8_t *scratch_dev, cudaStream_t s) {
  // scratch: [0,32) seed copy, [64, 64+200) sponge state
  CHPIR_CUDA(cudaMemcpyAsync(scratch_dev, seed, 32, cudaMemcpyHostToDevice, s), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  uint64_t *state = reinterpret_cast<uint64_t *>(scratch_dev + 64);
  const uint64_t blocks = (total_bytes + 167) / 168;
  const uint64_t per_launch = 1ull << 20;  // ~1 s of chain per launch: keeps any watchdog and the stream responsive
  for (uint64_t b0 = 0; b0 < blocks; b0 += per_launch) {
    const uint64_t cnt = blocks - b0 < per_launch ? blocks - b0 : per_launch;
    expand_kernel<<<1, 32, 0, s>>>(scratch_dev, state, b0 == 0 ? 1 : 0, out_dev, total_bytes, b0, cnt);
    if (cudaGetLastError() != cudaSuccess) return CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
  }
  return CHPIR_OK;
}

}  // namespace chpir
