// expand.cu -- Matrix::generate_from_seed (chalametpir_common/src/matrix.rs:541-558) on device:
//   TurboShake128::default(); absorb(seed[32]); finalize::<0x1F>(); squeeze(4*rows*cols bytes)
// bit-exactly, written either as row-major u32 (the reference's element memory) or straight into the K-major byte
// ("limb") planes the tensor-core hint GEMM consumes, so that A never exists as u32 in HBM during setup.
//
// TurboSHAKE128 = Keccak-p[1600, 12 rounds] sponge, rate 168 bytes (RFC 9861; `turboshake` crate =0.4.1 is the
// reference's dependency).  The squeeze is ONE serial chain -- block i+1 is a permutation of block i -- so there is no
// parallelism across blocks; the kernel is a single warp and is bound by the latency of one round.  The design goal is
// therefore the shortest dependent chain per round, not throughput:
//   * lane t = x + 5y (t < 25) owns the 64-bit Keccak lane A[x,y] in two registers;
//   * two shared-memory exchanges per round (a store, a __syncwarp and vector loads -- fewer instructions and a
//     shorter chain than the 18 shuffles of a shuffle-only round):
//       1. theta: every lane stores A and reads the two neighbouring COLUMNS whole (3 vector loads each), computes
//          both column parities itself and applies D = C[x-1] ^ rol(C[x+1], 1) -- the separate "exchange the
//          parities" step of the textbook formulation disappears;
//       2. rho + pi + chi: every lane stores its rotated lane at its pi destination and reads the three lanes of
//          its chi row;
//   * rho rotations >= 32 cost nothing: the two halves are simply stored to swapped addresses;
//   * iota is taken off the chain: lane (0,0) keeps its value WITHOUT the round constant; the constant reaches the
//     next round's theta through a per-lane, per-round compile-time constant XORed into the lane's own value while
//     the loads are in flight (C[0] is linear in A[0,0]), and is applied for real only when a block is written out.
#include "common.cuh"

namespace chpir {
namespace {

__constant__ uint32_t kRho[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

// round constants of rounds 12..23 of Keccak-f[1600]
__host__ __device__ constexpr uint64_t rc_of(int i) {
  constexpr uint64_t rc[12] = {0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
                               0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
                               0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  return rc[i];
}

// shared memory map (bytes)
constexpr uint32_t kS1 = 0;          // theta exchange: column x at x*48, lane y at +8y
constexpr uint32_t kS2 = 256;        // chi exchange: B[X][Y] at (X + 5Y)*8
constexpr uint32_t kDump = 512;      // lanes 25..31 store here (one 8-byte slot each, per exchange)
constexpr uint32_t kSmemBytes = 640;

struct LaneCfg {
  uint32_t s1_own, s1_m, s1_p;   // own slot, base of column x-1, base of column x+1
  uint32_t s2_st_a, s2_st_b;     // where the two words of the rotated lane go (swap for rho >= 32 folded in)
  uint32_t s2_b0, s2_b1, s2_b2;  // chi operands B[x,y], B[x+1,y], B[x+2,y]
  uint32_t rot;                  // rho amount & 31
  uint32_t mask_a, mask_b;       // iota folding: lanes that see RC directly (lane 0, column 1) / rotated (column 4)
};

__device__ __forceinline__ LaneCfg make_cfg(uint32_t smem_base, int lane) {
  LaneCfg c;
  const bool real = lane < 25;
  const int t = real ? lane : lane - 25;
  const int x = t % 5, y = t / 5;
  c.s1_own = smem_base + (real ? kS1 + x * 48 + y * 8 : kDump + (lane - 25) * 8);
  c.s1_m = smem_base + kS1 + ((x + 4) % 5) * 48;
  c.s1_p = smem_base + kS1 + ((x + 1) % 5) * 48;
  const uint32_t r = kRho[t];
  c.rot = r & 31u;
  const int X = y, Y = (2 * x + 3 * y) % 5;  // pi
  const uint32_t dst = smem_base + (real ? kS2 + (X + 5 * Y) * 8 : kDump + 64 + (lane - 25) * 8);
  c.s2_st_a = dst + (r >= 32u ? 4 : 0);
  c.s2_st_b = dst + (r >= 32u ? 0 : 4);
  c.s2_b0 = smem_base + kS2 + (x + 5 * y) * 8;
  c.s2_b1 = smem_base + kS2 + ((x + 1) % 5 + 5 * y) * 8;
  c.s2_b2 = smem_base + kS2 + ((x + 2) % 5 + 5 * y) * 8;
  c.mask_a = (real && (t == 0 || x == 1)) ? 0xffffffffu : 0u;
  c.mask_b = (real && x == 4) ? 0xffffffffu : 0u;
  return c;
}

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t a) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory"); }
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ void lds64(uint32_t addr, uint32_t &a, uint32_t &b) {
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory");
}

// One Keccak-p[1600,12] on the warp.  (lo, hi) is the lane's value with the iota constant of the PREVIOUS round
// (round 11 of the previous permutation for round 0) still pending on lane 0.
__device__ __forceinline__ void keccak_p12_warp(uint32_t &lo, uint32_t &hi, const LaneCfg &c) {
#pragma unroll
  for (int round = 0; round < 12; round++) {
    // ---- theta
    sts64(c.s1_own, lo, hi);
    __syncwarp();
    uint32_t m[10], p[10];
    lds128(c.s1_m, m[0], m[1], m[2], m[3]);
    lds128(c.s1_p, p[0], p[1], p[2], p[3]);
    lds128(c.s1_m + 16, m[4], m[5], m[6], m[7]);
    lds128(c.s1_p + 16, p[4], p[5], p[6], p[7]);
    lds64(c.s1_m + 32, m[8], m[9]);
    lds64(c.s1_p + 32, p[8], p[9]);
    // pending iota, folded while the loads are in flight
    const uint64_t pend = rc_of((round + 11) % 12);
    const uint32_t pl = uint32_t(pend), ph = uint32_t(pend >> 32);
    const uint32_t ql = (pl << 1) | (ph >> 31), qh = (ph << 1) | (pl >> 31);
    const uint32_t fl = lo ^ (pl & c.mask_a) ^ (ql & c.mask_b);
    const uint32_t fh = hi ^ (ph & c.mask_a) ^ (qh & c.mask_b);
    const uint32_t cml = (m[0] ^ m[2] ^ m[4]) ^ m[6] ^ m[8], cmh = (m[1] ^ m[3] ^ m[5]) ^ m[7] ^ m[9];
    const uint32_t cpl = (p[0] ^ p[2] ^ p[4]) ^ p[6] ^ p[8], cph = (p[1] ^ p[3] ^ p[5]) ^ p[7] ^ p[9];
    const uint32_t al = fl ^ cml ^ __funnelshift_l(cph, cpl, 1);
    const uint32_t ah = fh ^ cmh ^ __funnelshift_l(cpl, cph, 1);
    // ---- rho (halves land swapped in shared memory when the amount is >= 32) + pi
    sts32(c.s2_st_a, __funnelshift_l(ah, al, c.rot));
    sts32(c.s2_st_b, __funnelshift_l(al, ah, c.rot));
    __syncwarp();
    // ---- chi (iota stays pending)
    uint32_t b0l, b0h, b1l, b1h, b2l, b2h;
    lds64(c.s2_b0, b0l, b0h);
    lds64(c.s2_b1, b1l, b1h);
    lds64(c.s2_b2, b2l, b2h);
    lo = b0l ^ (~b1l & b2l);
    hi = b0h ^ (~b1h & b2h);
  }
}

struct ExpandOut {
  // u32 mode
  uint8_t *out;          // row-major u32 stream
  uint64_t total_bytes;
  // limb-plane mode: ring of 2 panel buffers, each [4 planes][128 rows][kp bytes]
  uint8_t *ring;
  uint64_t cols;         // K
  uint64_t kp;           // plane row pitch (bytes)
  uint64_t rows;         // total rows of A
};

constexpr uint32_t kPanelRows = 128;

// state: 25 x u64 (lane-major, with the iota of the last round pending).  One warp.
template <bool PLANES>
__global__ void __launch_bounds__(32, 1) expand_kernel(const uint8_t *__restrict__ seed, uint64_t *__restrict__ state, int first, ExpandOut o,
                                                        uint64_t block_begin, uint64_t block_count) {
  __shared__ __align__(16) uint8_t smem[kSmemBytes];
  const int lane = threadIdx.x;
  const LaneCfg c = make_cfg(static_cast<uint32_t>(__cvta_generic_to_shared(smem)), lane);
  const uint64_t last_rc = rc_of(11);
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
    // round 0 always cancels a pending round-11 constant: pre-apply it so that nothing is pending in effect
    if (lane == 0) lo ^= uint32_t(last_rc), hi ^= uint32_t(last_rc >> 32);
  } else if (lane < 25) {
    const uint64_t v = state[lane];
    lo = uint32_t(v), hi = uint32_t(v >> 32);
  }

  // position of this lane's first element (2 u32 elements per lane per block, 42 per block)
  uint64_t row = 0, col = 0;
  if (PLANES) {
    const uint64_t e = block_begin * 42ull + 2ull * lane;
    row = e / o.cols;
    col = e - row * o.cols;
  }
  const uint64_t plane_bytes = uint64_t(kPanelRows) * o.kp;

  for (uint64_t blk = block_begin; blk < block_begin + block_count; blk++) {
    keccak_p12_warp(lo, hi, c);
    const uint32_t vl = lane == 0 ? lo ^ uint32_t(last_rc) : lo;
    const uint32_t vh = lane == 0 ? hi ^ uint32_t(last_rc >> 32) : hi;
    if (!PLANES) {
      if (lane < 21) {
        const uint64_t off = blk * 168ull + 8ull * lane;
        if (off + 8 <= o.total_bytes) {
          *reinterpret_cast<uint2 *>(o.out + off) = make_uint2(vl, vh);
        } else if (off + 4 <= o.total_bytes) {
          *reinterpret_cast<uint32_t *>(o.out + off) = vl;
        }
      }
    } else {
      if (lane < 21) {
        // element (row, col) -> ring buffer (row / 128) & 1, plane l, local row row % 128
        uint8_t *dst = o.ring + ((row / kPanelRows) & 1) * 4 * plane_bytes + (row % kPanelRows) * o.kp + col;
        if (col + 1 < o.cols && !(col & 1)) {
          if (row < o.rows) {
#pragma unroll
            for (int l = 0; l < 4; l++)
              *reinterpret_cast<uint16_t *>(dst + l * plane_bytes) = uint16_t(((vl >> (8 * l)) & 0xffu) | (((vh >> (8 * l)) & 0xffu) << 8));
          }
        } else {
          if (row < o.rows) {
#pragma unroll
            for (int l = 0; l < 4; l++) dst[l * plane_bytes] = uint8_t(vl >> (8 * l));
          }
          uint64_t r2 = row, c2 = col + 1;
          if (c2 >= o.cols) c2 = 0, r2++;
          if (r2 < o.rows) {
            uint8_t *d2 = o.ring + ((r2 / kPanelRows) & 1) * 4 * plane_bytes + (r2 % kPanelRows) * o.kp + c2;
#pragma unroll
            for (int l = 0; l < 4; l++) d2[l * plane_bytes] = uint8_t(vh >> (8 * l));
          }
        }
      }
      col += 42;
      while (col >= o.cols) col -= o.cols, row++;
    }
  }
  if (lane < 25) state[lane] = uint64_t(lo) | uint64_t(hi) << 32;
}

}  // namespace

// scratch: [0,32) seed copy, [64, 64+200) sponge state
int expand_begin(const uint8_t seed[32], uint8_t *scratch_dev, cudaStream_t s) {
  CHPIR_CUDA(cudaMemcpyAsync(scratch_dev, seed, 32, cudaMemcpyHostToDevice, s), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  return CHPIR_OK;
}

int launch_expand(const uint8_t seed[32], uint8_t *out_dev, uint64_t total_bytes, uint8_t *scratch_dev, cudaStream_t s) {
  if (int rc = expand_begin(seed, scratch_dev, s); rc != CHPIR_OK) return rc;
  uint64_t *state = reinterpret_cast<uint64_t *>(scratch_dev + 64);
  const uint64_t blocks = (total_bytes + 167) / 168;
  const uint64_t per_launch = 1ull << 21;  // ~1 s of chain per launch: keeps the stream responsive
  ExpandOut o{};
  o.out = out_dev;
  o.total_bytes = total_bytes;
  o.cols = 1;
  for (uint64_t b0 = 0; b0 < blocks; b0 += per_launch) {
    const uint64_t cnt = blocks - b0 < per_launch ? blocks - b0 : per_launch;
    expand_kernel<false><<<1, 32, 0, s>>>(scratch_dev, state, b0 == 0 ? 1 : 0, o, b0, cnt);
    if (cudaGetLastError() != cudaSuccess) return CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
  }
  return CHPIR_OK;
}

// XOF blocks [block_begin, block_begin + block_count) of the stream of an (rows x cols) u32 matrix, scattered into the
// two-panel ring of limb planes.  expand_begin() must have been enqueued before the first call (block_begin == 0).
int launch_expand_planes(uint8_t *ring_dev, uint64_t rows, uint64_t cols, uint64_t kp, uint8_t *scratch_dev, uint64_t block_begin,
                         uint64_t block_count, cudaStream_t s) {
  if (block_count == 0) return CHPIR_OK;
  ExpandOut o{};
  o.ring = ring_dev;
  o.rows = rows;
  o.cols = cols;
  o.kp = kp;
  uint64_t *state = reinterpret_cast<uint64_t *>(scratch_dev + 64);
  expand_kernel<true><<<1, 32, 0, s>>>(scratch_dev, state, block_begin == 0 ? 1 : 0, o, block_begin, block_count);
  return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}

}  // namespace chpir
