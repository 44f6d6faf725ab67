// gemm_tc.cu -- the setup hint GEMM  M = A . D mod 2^32  on the 5th-generation tensor cores (tcgen05, kind::i8).
//
// Replaces (reference): `impl Mul<&Matrix> for &Matrix` (chalametpir_common/src/matrix.rs:1040-1059) and the Vulkan
// offload shaders/mat_x_mat.glsl:27-47 + shaders/mat_transpose.glsl:20-37.
//
// Arithmetic.  A is full-range u32 = 4 unsigned byte limbs a0..a3; D has b <= 14 bits = NB <= 2 byte limbs d0,d1.
//   A.D mod 2^32 = sum_{i+j<=3} 2^{8(i+j)} (A_i . D_j)            (limb products with i+j >= 4 vanish mod 2^32)
// Every A_i . D_j is a u8 x u8 GEMM with int32 accumulation.  int32 accumulation that WRAPS (instruction descriptor
// saturate bit = 0) is exactly accumulation mod 2^32, so K needs no chunking; products with equal shift s = i+j share
// one accumulator, giving 4 TMEM accumulators and 7 (NB=2) or 4 (NB=1) MMAs per 32-deep k-step.  The epilogue
// recombines  acc0 + (acc1<<8) + (acc2<<16) + (acc3<<24)  and adds it into C with u32 atomics (split-K over CTAs is
// exact because addition mod 2^32 is order independent).
//
// Data movement.  Two streaming pre-passes put the operands in the K-major byte layout the MMA wants:
//   split_a_limbs      A[m][k] u32            -> A8[4][m][kp]   u8   (byte planes; the LE bytes of a u32 ARE the limbs)
//   split_transpose_b  D[k][n] u32 (ld)       -> B8[NB][n][kp]  u8   (limb split + shared-memory tiled transpose)
// The main kernel is warp specialised: warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle, OOB zero fill for
// the ragged M/N/K edges), warp 1 = single-thread tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld).
//
// Clusters.  TMEM holds one 128 x 128 output tile per SM (4 shift accumulators x 128 columns = all 512 columns), so a k-block costs
// every CTA 64 KB of A (four limb tiles) + 32 KB of B from L2, and round 1 measured the kernel bound by exactly that: ~4700 B/clk
// of L2 -> SM delivery chip-wide, 70 % tensor-pipe activity.  The N tiles of one K range all read the SAME A tile, so they form a
// thread-block cluster along N (2, 4 or 8 CTAs) in which every CTA fetches 1/csz of the A tile and TMA-multicasts it to all of
// them: L2 reads per CTA and k-block drop from 96 KB to 32 + 64/csz KB.  A stage is released to the producers by the MMA commits
// of ALL CTAs of the cluster (tcgen05.commit ... multicast::cluster onto every CTA's `empty` barrier), since any of them may be
// written by any peer.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace chpir {
namespace {

constexpr int BM = 128;        // UMMA M (cta_group::1)
constexpr int BN_MAX = 128;    // accumulator columns per shift; 4 * 128 = all 512 TMEM columns
// Bytes of K per stage row = one swizzle row, and pipeline depth.  A stage holds 4 A limbs + up to 2 B limbs of 128 rows each:
// 96 KB at BK = 128 (two stages fit), 48 KB at BK = 64 (four stages).  The two-stage pipeline leaves the tensor pipe 70 % active
// (ncu, round 1).  What bounds it is the rate at which TMA fills shared memory, not latency: a k-block takes ~3100 cycles for 768
// 128-byte rows (32 B/clk per SM, 4700 B/clk chip-wide) where its MMAs need 1728.  Doubling the TMA work per byte of K makes the
// kernel proportionally slower: four half-size stages (-DCHPIR_GEMM_BK=64, 64B swizzle, 12 loads per 128 B of K) 19.9 ms against
// 11.4 ms for the 2^20 hint GEMM; L2 prefetches (cp.async.bulk.prefetch.tensor) of A and B ahead of the loads 19.9 ms, of B alone
// 14.0 ms.  So 128 B x 2 stages stays, and the next step for this kernel is fewer L2 reads per MAC (A multicast across a cluster
// of N-tile CTAs), not a deeper pipeline.
#ifndef CHPIR_GEMM_BK
#define CHPIR_GEMM_BK 128
#endif
constexpr int BK = CHPIR_GEMM_BK;
static_assert(BK == 128 || BK == 64, "BK is one 128B- or 64B-swizzle row");
constexpr int STAGES = BK == 128 ? 2 : 4;
constexpr int A_TILE = BM * BK;      // 16 KB per limb
constexpr int B_TILE = BN_MAX * BK;  // 16 KB per limb
constexpr int kThreads = 192;

// ----------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, uint16_t cta_mask) {
  // the tile lands at the same shared-memory offset of every CTA in cta_mask and completes bytes on the barrier at `bar`'s offset there
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major, swizzled operand tile: rows of BK bytes (128B or 64B swizzle), 8-row groups 8*BK bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);  // start address
  d |= uint64_t(1) << 16;                     // leading byte offset (ignored for swizzled K-major), canonical value 1
  d |= uint64_t((8 * BK) >> 4) << 32;         // stride byte offset: 8 rows * BK bytes
  d |= uint64_t(1) << 46;                     // descriptor version (sm_100)
  d |= uint64_t(BK == 128 ? 2 : 4) << 61;     // SWIZZLE_128B / SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

struct __align__(8) Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t acc_full;
  uint32_t tmem_base;
};

// grid: (tiles_m * tiles_n, splits), cluster (csz, 1, 1) along the N tiles.  One output tile x one K range per CTA.
// map_a: box of 128 rows (csz = 1); map_a64: box of 64 rows, the unit of the multicast A slices (csz > 1).
template <int NB>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a64, const __grid_constant__ CUtensorMap map_b,
                   uint32_t *__restrict__ C, uint32_t m, uint32_t n, uint32_t tiles_n, uint32_t bn, uint32_t kblocks_total, uint32_t kblocks_per_split,
                   uint32_t csz) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Barriers bars;
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int STAGE_BYTES = 4 * A_TILE + NB * B_TILE;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t tile_m = blockIdx.x / tiles_n, tile_n = blockIdx.x % tiles_n;
  const uint32_t kb0 = blockIdx.y * kblocks_per_split;
  uint32_t kb1 = kb0 + kblocks_per_split;
  if (kb1 > kblocks_total) kb1 = kblocks_total;
  const uint32_t nkb = kb1 > kb0 ? kb1 - kb0 : 0;

  const uint32_t crank = csz > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = uint16_t((1u << csz) - 1u);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(csz > 1 ? &map_a64 : &map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], csz);  // one commit from every CTA that reads (and whose peers write) this stage
    }
    mbar_init(&bars.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (csz > 1) cluster_sync_all();  // every CTA's barriers exist before a peer's multicast or commit can reach them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = bars.tmem_base;

  if (nkb > 0) {
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      if (lane == 0) {
        const uint32_t tx_bytes = 4 * A_TILE + NB * bn * BK;
        for (uint32_t i = 0; i < nkb; i++) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          mbar_wait(&bars.empty[s], ph ^ 1);
          mbar_expect_tx(&bars.full[s], tx_bytes);
          uint8_t *st = smem + s * STAGE_BYTES;
          const int kc = int((kb0 + i) * BK);
          if (csz == 1) {
#pragma unroll
            for (int l = 0; l < 4; l++) tma_load_3d(st + l * A_TILE, &map_a, &bars.full[s], kc, int(tile_m * BM), l);
          } else {
            // the A stage is 8 half-limb slices of 64 rows; this CTA fetches 8 / csz of them for everybody
            const uint32_t per = 8u / csz;
            for (uint32_t j = crank * per; j < (crank + 1) * per; j++) {
              const uint32_t l = j >> 1, half = j & 1u;
              tma_load_3d_mc(st + l * A_TILE + half * (64 * BK), &map_a64, &bars.full[s], kc, int(tile_m * BM + half * 64), int(l), cmask);
            }
          }
#pragma unroll
          for (int l = 0; l < NB; l++) tma_load_3d(st + 4 * A_TILE + l * B_TILE, &map_b, &bars.full[s], kc, int(tile_n * bn), l);
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer (one thread)
      if (lane == 0) {
        // kind::i8, u8 x u8 -> s32 (wrapping), K-major A and B, M = 128, N = bn
        const uint32_t idesc = (2u << 4) | ((bn >> 3) << 17) | (uint32_t(BM >> 4) << 24);
        const uint32_t idesc_wide = (2u << 4) | (((BN_MAX + bn) >> 3) << 17) | (uint32_t(BM >> 4) << 24);
        for (uint32_t i = 0; i < nkb; i++) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          mbar_wait(&bars.full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_base = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_base = a_base + 4 * A_TILE;
#pragma unroll
          for (int kk = 0; kk < BK / 32; kk++) {
            const uint32_t first = (i == 0 && kk == 0) ? 0u : 1u;
            uint64_t da[4], db[NB];
#pragma unroll
            for (int l = 0; l < 4; l++) da[l] = umma_desc_sw128(a_base + l * A_TILE + kk * 32);
#pragma unroll
            for (int l = 0; l < NB; l++) db[l] = umma_desc_sw128(b_base + l * B_TILE + kk * 32);
            // accumulator s = i + j lives at TMEM columns [s*128, s*128 + bn)
            if (NB == 2 && first != 0u) {
              // Steady state: the d0 and d1 tiles are adjacent in shared memory (rows 0..127 and 128..128+bn of one
              // K-major tile) and the accumulators of shift i and i+1 are adjacent in TMEM, so A_i x [d0; d1] is ONE
              // MMA of N = 128 + bn that lands A_i.d0 in acc_i and A_i.d1 in acc_{i+1}.  4 MMAs instead of 7 for the same
              // tensor work, and every A_i tile is read from shared memory once instead of twice.
              umma_i8(tmem + 0 * BN_MAX, da[0], db[0], idesc_wide, 1u);
              umma_i8(tmem + 1 * BN_MAX, da[1], db[0], idesc_wide, 1u);
              umma_i8(tmem + 2 * BN_MAX, da[2], db[0], idesc_wide, 1u);
              umma_i8(tmem + 3 * BN_MAX, da[3], db[0], idesc, 1u);
            } else {
              // very first k-step (accumulators are overwritten, not accumulated) and the single-limb case
              umma_i8(tmem + 0 * BN_MAX, da[0], db[0], idesc, first);
              umma_i8(tmem + 1 * BN_MAX, da[1], db[0], idesc, first);
              umma_i8(tmem + 2 * BN_MAX, da[2], db[0], idesc, first);
              umma_i8(tmem + 3 * BN_MAX, da[3], db[0], idesc, first);
              if (NB == 2) {
                umma_i8(tmem + 1 * BN_MAX, da[0], db[NB - 1], idesc, 1u);
                umma_i8(tmem + 2 * BN_MAX, da[1], db[NB - 1], idesc, 1u);
                umma_i8(tmem + 3 * BN_MAX, da[2], db[NB - 1], idesc, 1u);
              }
            }
          }
          // frees the smem stage once these MMAs have read it -- in every CTA of the cluster, whose producers all write into it
          if (csz == 1)
            umma_commit(&bars.empty[s]);
          else
            umma_commit_mc(&bars.empty[s], cmask);
        }
        umma_commit(&bars.acc_full);
      }
    } else {
      // ---------------------------------------------------------------- epilogue: TMEM -> registers -> atomics on C
      mbar_wait(&bars.acc_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_group = warp % 4;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
      const uint32_t row = tile_m * BM + lane_group * 32 + lane;
      const uint32_t taddr = tmem + ((lane_group * 32) << 16);
      for (uint32_t c0 = 0; c0 < bn; c0 += 16) {
        uint32_t v0[16], v1[16], v2[16], v3[16];
        tmem_ld16(taddr + 0 * BN_MAX + c0, v0);
        tmem_ld16(taddr + 1 * BN_MAX + c0, v1);
        tmem_ld16(taddr + 2 * BN_MAX + c0, v2);
        tmem_ld16(taddr + 3 * BN_MAX + c0, v3);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < m) {
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const uint32_t col = tile_n * bn + c0 + j;
            const uint32_t v = v0[j] + (v1[j] << 8) + (v2[j] << 16) + (v3[j] << 24);
            if (col < n && c0 + j < bn) atomicAdd(C + uint64_t(row) * n + col, v);
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (csz > 1) cluster_sync_all();  // no CTA leaves while a peer's commit may still be on its way to this CTA's barriers
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------- CTA pairs (cta_group::2)
// The same product with the roles of the operands swapped and two SMs working on one MMA:  C^T[n][m] = D^T . A^T.
//   M side of the MMA (256 rows per pair, 128 per CTA) = rows of D's limb planes d0, d1 (columns of the hint / response),
//   N side                                              = rows of A's limb planes (rows of A / queries), mq <= 128 of them per pass.
// A cta_group::2 MMA takes half of its N-side operand from each CTA of the pair, so a CTA stages only
//   d0, d1 of its own 128 columns (2 x 16 KB), ONE of the limbs A0 / A1 for all mq rows (CTA 0: A0, CTA 1: A1), and its half of the rows
//   of A2 and A3                                                                          = 64 KB per k-block of 128 at mq = 128
// against 96 KB for the one-SM kernel above (three stages fit instead of two), and every MMA reads less shared memory per MAC
// (34 KB per 32-deep k-step instead of 44 KB).  Accumulators: shift s = i + j at TMEM columns [s * mq, (s + 1) * mq), lanes = the CTA's
// 128 columns of D.  Per 32-deep k-step the leader CTA issues
//   d0 x [A0 | A1] -> (s0, s1)   N = 2 mq        d0 x A2 -> s2   N = mq        d0 x A3 -> s3   N = mq
//   d1 x [A0 | A1] -> (s1, s2)   N = 2 mq        d1 x A2 -> s3   N = mq                         (d1 x A3 vanishes mod 2^32)
// in exactly this order: the first three overwrite all four accumulators on the very first k-step, the last two always accumulate.
// mq is the number of A rows rounded up to 32 (an N = mq MMA of a pair needs N % 32 == 0): a 64-query batch pays for 64 rows, not 128.
// Both CTAs run a TMA producer (signalling the LEADER's `full` barrier), only the leader issues MMAs, and its commits are multicast
// to both CTAs' `empty` / `acc_full` barriers.  The epilogue thread owns one column of D and walks the mq rows: for a fixed row the
// 32 lanes of a warp add into 32 consecutive words of C (one coalesced reduction per instruction).
constexpr int D_TILE = 128 * BK;  // 16 KB: one limb of D, the CTA's 128 columns
template <int NB>
__host__ __device__ constexpr int pair_stage_bytes() {
  return NB * D_TILE + 2 * BM * BK;  // d limbs + (A0|A1: 128 rows) + (A2, A3: 64 rows each)
}
template <int NB>
__host__ __device__ constexpr int pair_stages() {
  return NB == 2 ? 3 : 4;
}
constexpr int kPairMaxStages = 4;

struct __align__(8) PairBarriers {
  uint64_t full[kPairMaxStages];
  uint64_t empty[kPairMaxStages];
  uint64_t acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
// box -> this CTA's shared memory, bytes completed on a barrier that may live in the peer CTA (the leader's `full`)
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}

// grid: (2 * pairs, splits), cluster (2, 1, 1): CTA x owns columns [128 x, 128 x + 128) of D; one K range per blockIdx.y.
// map_ah: A's limb planes [4][128][kp], box = mq / 2 rows; map_d: D's limb planes [NB][n][kp], box = 128 rows.
template <int NB>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_d, uint32_t *__restrict__ C, uint32_t m,
                        uint32_t n, uint32_t mq, uint32_t kblocks_total, uint32_t kblocks_per_split) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ PairBarriers bars;
  constexpr int ST = pair_stages<NB>();
  constexpr int STAGE_BYTES = pair_stage_bytes<NB>();
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t crank = cluster_ctarank();  // 0 = leader (issues the MMAs, owns the `full` barriers)
  const uint32_t n0 = blockIdx.x * 128u;
  const uint32_t kb0 = blockIdx.y * kblocks_per_split;
  uint32_t kb1 = kb0 + kblocks_per_split;
  if (kb1 > kblocks_total) kb1 = kblocks_total;
  const uint32_t nkb = kb1 > kb0 ? kb1 - kb0 : 0;
  const uint32_t half = mq / 2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_ah)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_d)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < ST; s++) {
      mbar_init(&bars.full[s], 1);   // the leader's producer arrives once with the bytes of BOTH CTAs' loads
      mbar_init(&bars.empty[s], 1);  // one multicast commit of the leader's MMA thread
    }
    mbar_init(&bars.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before a load or a commit can reach them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = bars.tmem_base;

  if (nkb > 0) {
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (both CTAs)
      if (lane == 0) {
        const uint32_t tx_bytes = 2u * (NB * D_TILE + mq * 2u * BK);  // both CTAs of the pair
        for (uint32_t i = 0; i < nkb; i++) {
          const int s = i % ST;
          const uint32_t ph = (i / ST) & 1;
          mbar_wait(&bars.empty[s], ph ^ 1);
          if (crank == 0) mbar_expect_tx(&bars.full[s], tx_bytes);
          const uint32_t full = mapa_u32(smem_u32(&bars.full[s]), 0);
          const uint32_t st = smem + s * STAGE_BYTES;
          const int kc = int((kb0 + i) * BK);
#pragma unroll
          for (int l = 0; l < NB; l++) tma_load_3d_pair(st + l * D_TILE, &map_d, full, kc, int(n0), l);
          const uint32_t ax = st + NB * D_TILE;
          tma_load_3d_pair(ax, &map_ah, full, kc, 0, int(crank));                       // A0 (leader) / A1 (peer), rows [0, mq/2)
          tma_load_3d_pair(ax + half * BK, &map_ah, full, kc, int(half), int(crank));   //                        rows [mq/2, mq)
          tma_load_3d_pair(ax + mq * BK, &map_ah, full, kc, int(crank * half), 2);      // this CTA's half of A2
          tma_load_3d_pair(ax + mq * BK + half * BK, &map_ah, full, kc, int(crank * half), 3);  // and of A3
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer (one thread of the leader CTA)
      if (lane == 0 && crank == 0) {
        // kind::i8, u8 x u8 -> s32 (wrapping), K-major operands, M = 256 over the pair
        const uint32_t idesc_wide = (2u << 4) | (((2u * mq) >> 3) << 17) | (uint32_t(256 >> 4) << 24);
        const uint32_t idesc_half = (2u << 4) | ((mq >> 3) << 17) | (uint32_t(256 >> 4) << 24);
        for (uint32_t i = 0; i < nkb; i++) {
          const int s = i % ST;
          const uint32_t ph = (i / ST) & 1;
          mbar_wait(&bars.full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_base = smem + s * STAGE_BYTES;
          const uint32_t ax = d_base + NB * D_TILE, a2 = ax + mq * BK, a3 = a2 + half * BK;
#pragma unroll
          for (int kk = 0; kk < BK / 32; kk++) {
            const uint32_t acc = (i == 0 && kk == 0) ? 0u : 1u;
            const uint64_t dd0 = umma_desc_sw128(d_base + kk * 32), dax = umma_desc_sw128(ax + kk * 32);
            const uint64_t da2 = umma_desc_sw128(a2 + kk * 32), da3 = umma_desc_sw128(a3 + kk * 32);
            umma_i8_pair(tmem + 0 * mq, dd0, dax, idesc_wide, acc);
            umma_i8_pair(tmem + 2 * mq, dd0, da2, idesc_half, acc);
            umma_i8_pair(tmem + 3 * mq, dd0, da3, idesc_half, acc);
            if (NB == 2) {
              const uint64_t dd1 = umma_desc_sw128(d_base + D_TILE + kk * 32);
              umma_i8_pair(tmem + 1 * mq, dd1, dax, idesc_wide, 1u);
              umma_i8_pair(tmem + 3 * mq, dd1, da2, idesc_half, 1u);
            }
          }
          umma_commit_pair(&bars.empty[s], 0b11);  // both CTAs' producers may refill the stage once these MMAs have read it
        }
        umma_commit_pair(&bars.acc_full, 0b11);
      }
    } else {
      // ---------------------------------------------------------------- epilogue (both CTAs): TMEM -> registers -> atomics on C
      mbar_wait(&bars.acc_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_group = warp % 4;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
      const uint32_t col = n0 + lane_group * 32 + lane;
      const uint32_t taddr = tmem + ((lane_group * 32) << 16);
      for (uint32_t r0 = 0; r0 < mq; r0 += 16) {
        uint32_t v0[16], v1[16], v2[16], v3[16];
        tmem_ld16(taddr + 0 * mq + r0, v0);
        tmem_ld16(taddr + 1 * mq + r0, v1);
        tmem_ld16(taddr + 2 * mq + r0, v2);
        tmem_ld16(taddr + 3 * mq + r0, v3);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (col < n) {
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const uint32_t v = v0[j] + (v1[j] << 8) + (v2[j] << 16) + (v3[j] << 24);
            if (r0 + j < m) atomicAdd(C + uint64_t(r0 + j) * n + col, v);
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // the leader's MMAs wrote this CTA's TMEM and its commits target this CTA's barriers: leave together
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// A[m][k] u32 -> planes[l][plane_rows][kp] u8 (rows 0..m-1 filled), l = byte index.  One thread per 4 consecutive k.
__global__ void split_a_limbs(const uint32_t *__restrict__ A, uint32_t m, uint64_t k, uint64_t kp, uint32_t plane_rows,
                              uint8_t *__restrict__ planes) {
  const uint64_t groups = kp / 4;
  const uint64_t total = uint64_t(m) * groups;
  const uint64_t plane = uint64_t(plane_rows) * kp;
  for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < total; idx += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t r = idx / groups, g = idx - r * groups;
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint64_t kk = 4 * g + j;
      v[j] = kk < k ? __ldg(A + r * k + kk) : 0u;
    }
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const uint32_t w = ((v[0] >> (8 * l)) & 0xffu) | (((v[1] >> (8 * l)) & 0xffu) << 8) | (((v[2] >> (8 * l)) & 0xffu) << 16) |
                         (((v[3] >> (8 * l)) & 0xffu) << 24);
      *reinterpret_cast<uint32_t *>(planes + l * plane + r * kp + 4 * g) = w;
    }
  }
}

// D[k][n] u32 (leading dimension ld) -> planes[l][n][kp] u8: limb split + transpose through a shared-memory tile.
// Block = 256 threads, tile = 128 k x 32 n.
template <int NB>
__global__ void __launch_bounds__(256) split_transpose_b(const uint32_t *__restrict__ B, uint64_t k, uint32_t n, uint32_t ld, uint64_t kp,
                                                          uint8_t *__restrict__ planes) {
  __shared__ __align__(4) uint8_t tile[NB][32][128 + 4];
  // the (1D) grid walks the n tiles of one band of 128 k-rows first: a row of B is not sector aligned (3760 bytes at N = 940), and the sectors
  // two neighbouring tiles share are then requested by CTAs that run together (one DRAM read, an L2 hit for the other) instead of
  // thousands of CTAs apart (round 1: 8.4 GB read for a 4.4 GB operand)
  const uint32_t tiles_n = (n + 31) / 32;
  const uint64_t k0 = uint64_t(blockIdx.x / tiles_n) * 128;
  const uint32_t n0 = (blockIdx.x % tiles_n) * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;  // ty in 0..7
  for (int kk = ty; kk < 128; kk += 8) {
    const uint64_t gk = k0 + kk;
    const uint32_t gn = n0 + tx;
    const uint32_t v = (gk < k && gn < n) ? __ldg(B + gk * ld + gn) : 0u;
#pragma unroll
    for (int l = 0; l < NB; l++) tile[l][tx][kk] = uint8_t(v >> (8 * l));
  }
  __syncthreads();
  const uint64_t plane = uint64_t(n) * kp;
  for (int rowi = ty; rowi < NB * 32; rowi += 8) {
    const int l = rowi / 32, nn = rowi % 32;
    const uint32_t gn = n0 + nn;
    const uint64_t gk = k0 + 4 * tx;
    if (gn < n && gk < kp) {
      const uint32_t w = *reinterpret_cast<const uint32_t *>(&tile[l][nn][4 * tx]);
      *reinterpret_cast<uint32_t *>(planes + l * plane + uint64_t(gn) * kp + gk) = w;
    }
  }
}

// ----------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// u8 tensor [planes][rows][k] with row pitch kp; box = 128 k-bytes x box_rows rows x 1 plane; 128B swizzle; OOB -> 0.
int make_map(CUtensorMap *map, void *base, uint64_t k, uint64_t kp, uint64_t rows, uint32_t planes, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
  const cuuint64_t dims[3] = {k, rows, planes};
  const cuuint64_t strides[2] = {kp, rows * kp};
  const cuuint32_t box[3] = {BK, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  BK == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}

}  // namespace

// Everything the hint GEMM needs that depends only on the B operand (D): its K-major limb planes, the TMA map over
// them and the launch geometry.  Built once per setup, then used for every 128-row panel of A.
struct GemmTcB {
  uint8_t *b8 = nullptr;
  uint8_t *a_ring = nullptr;  // two panel buffers [4][128][kp]
  uint64_t k = 0, kp = 0;
  uint32_t n = 0, nb = 0, bn = 0, tiles_n = 0, kblocks = 0, kbps = 0, splits = 0;
  uint32_t csz = 1;  // thread-block cluster along the N tiles (A multicast), 1 = none
  CUtensorMap map_b{}, map_a[2]{}, map_a64[2]{};
  // CTA-pair kernel (cta_group::2, the default): 2 * pairs CTAs of 128 columns, its own K split; A maps with boxes of 16 / 32 / 48 / 64 rows
  bool pair = true;
  uint32_t p_ctas = 0, p_kbps = 0, p_splits = 0;
  CUtensorMap map_d{}, map_ah[2][4]{};
  // Last reader of each ring buffer (the panel GEMM that consumed it), on whatever stream it ran: callers on different streams
  // -- concurrent respond_batch / coalesced respond / respond_device_tc on one server -- order themselves against it before
  // refilling the buffer (gemm_tc_buf_acquire / gemm_tc_buf_release), so the ring can be shared without serialising execution.
  cudaEvent_t buf_done[2] = {nullptr, nullptr};
  ~GemmTcB() {
    for (auto e : buf_done)
      if (e) cudaEventDestroy(e);
    if (b8) cudaFree(b8);
    if (a_ring) cudaFree(a_ring);
  }
};

uint64_t gemm_tc_panel_bytes(const GemmTcB *g) { return 4ull * BM * g->kp; }
uint8_t *gemm_tc_ring(const GemmTcB *g) { return g->a_ring; }
uint64_t gemm_tc_kp(const GemmTcB *g) { return g->kp; }
int gemm_tc_buf_acquire(const GemmTcB *g, int buf, cudaStream_t s) {
  return cudaStreamWaitEvent(s, g->buf_done[buf & 1], 0) == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}
int gemm_tc_buf_release(const GemmTcB *g, int buf, cudaStream_t s) {
  return cudaEventRecord(g->buf_done[buf & 1], s) == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}
void gemm_tc_free(GemmTcB *g) { delete g; }

int gemm_tc_prepare(const uint32_t *B, uint32_t ldb, uint64_t k, uint32_t n, uint32_t b_bits, int sm_count, cudaStream_t s, GemmTcB **out) {
  *out = nullptr;
  if (b_bits == 0 || b_bits > 16) return CHPIR_ERR_INVALID_ARGUMENT;  // two byte limbs cover every legal element width (4..14)
  if (k > 0x7fffffffull - BK) return CHPIR_ERR_INVALID_ARGUMENT;      // TMA coordinates are int32
  GemmTcB *g = new GemmTcB();
  g->k = k;
  g->kp = (k + 15) / 16 * 16;
  g->n = n;
  g->nb = b_bits <= 8 ? 1 : 2;
  int rc = CHPIR_OK;
  do {
    if (cudaMalloc(&g->b8, uint64_t(g->nb) * n * g->kp) != cudaSuccess || cudaMalloc(&g->a_ring, 2 * gemm_tc_panel_bytes(g)) != cudaSuccess ||
        cudaEventCreateWithFlags(&g->buf_done[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&g->buf_done[1], cudaEventDisableTiming) != cudaSuccess) {
      rc = CHPIR_ERR_CUDA_ALLOCATION_FAILED;
      break;
    }
    const uint64_t tr_tiles = uint64_t((n + 31) / 32) * ((g->kp + 127) / 128);
    if (tr_tiles > 0x7fffffffull) {
      rc = CHPIR_ERR_INVALID_ARGUMENT;
      break;
    }
    dim3 grid{unsigned(tr_tiles)};
    if (g->nb == 1)
      split_transpose_b<1><<<grid, 256, 0, s>>>(B, k, n, ldb, g->kp, g->b8);
    else
      split_transpose_b<2><<<grid, 256, 0, s>>>(B, k, n, ldb, g->kp, g->b8);
    if (cudaGetLastError() != cudaSuccess) {
      rc = CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
      break;
    }
    const uint32_t tiles_n0 = (n + BN_MAX - 1) / BN_MAX;
    // Cluster size.  Measured on a B200 (profiles/r2_gemm_sweep.txt, 128-query batch incl. the limb split of the queries):
    //   2^20 shape, 8 N tiles: 1.029 ms without clusters, 0.962 / 1.007 / 1.078 ms with clusters of 2 / 4 / 8;
    //   2^18 shape, 7 N tiles: 0.235 ms without, 0.258+ with (the tile count is rounded up to a multiple of the cluster size -- a tile
    //   past column n loads zeros and publishes nothing -- so an eighth, empty tile is paid for).
    // Multicast removes L2 READS, and the gain is small: what bounds the kernel is the bytes DELIVERED to each SM's shared memory
    // (96 KB per k-block whoever fetched them; ~31 B/clk per SM with all SMs streaming), and TMEM fixes the tile at 128 x 128 x 4
    // accumulators, so the bytes per MAC cannot shrink further without pairing SMs (cta_group::2 would share the B tile: 80 KB).
    // Default: pairs when the tile count is even, none otherwise; CHPIR_GEMM_CLUSTER = 1 / 2 / 4 / 8 overrides.
    uint32_t csz = tiles_n0 % 2 == 0 ? 2 : 1;
    if (const char *v = std::getenv("CHPIR_GEMM_CLUSTER"); v && *v) {
      csz = uint32_t(std::strtoul(v, nullptr, 10));
      if (csz != 1 && csz != 2 && csz != 4 && csz != 8) csz = 1;
    }
    while (csz > 1 && csz > tiles_n0) csz /= 2;
    g->csz = csz;
    const uint32_t tiles = (tiles_n0 + csz - 1) / csz * csz;
    uint32_t bn = ((n + tiles - 1) / tiles + 15) / 16 * 16;
    if (bn < 16) bn = 16;
    g->bn = bn;
    g->tiles_n = csz > 1 ? tiles : (n + bn - 1) / bn;
    g->kblocks = uint32_t((k + BK - 1) / BK);
    constexpr int smem1 = STAGES * (4 * A_TILE + 1 * B_TILE) + 1024, smem2 = STAGES * (4 * A_TILE + 2 * B_TILE) + 1024;
    if (cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1) != cudaSuccess ||
        cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2) != cudaSuccess) {
      rc = CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
      break;
    }
    // CTAs of one wave: every SM, or -- with clusters -- as many whole clusters as the GPCs can host at once
    uint32_t wave = uint32_t(sm_count);
    if (csz > 1) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(g->tiles_n, 1, 1), cfg.blockDim = dim3(kThreads, 1, 1);
      cfg.dynamicSmemBytes = g->nb == 1 ? smem1 : smem2;
      cudaLaunchAttribute at{};
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = csz, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
      cfg.attrs = &at, cfg.numAttrs = 1;
      int clusters = 0;
      const cudaError_t e = g->nb == 1 ? cudaOccupancyMaxActiveClusters(&clusters, gemm_tc_kernel<1>, &cfg)
                                       : cudaOccupancyMaxActiveClusters(&clusters, gemm_tc_kernel<2>, &cfg);
      if (e != cudaSuccess || clusters < 1) {
        (void)cudaGetLastError();
        clusters = sm_count / int(csz);
      }
      wave = uint32_t(clusters) * csz;
    }
    // one launch = one 128-row panel: split K so that tiles * splits fills whole waves
    auto split_k = [&](uint32_t tiles, uint32_t wave_ctas, uint32_t *kbps, uint32_t *splits) {
      uint32_t best_splits = 1;
      double best_eff = 0.0;
      uint32_t max_splits = g->kblocks / (1024 / BK);  // every split keeps >= 1024 k of mainloop per epilogue
      if (max_splits < 1) max_splits = 1;
      if (max_splits > 148) max_splits = 148;
      for (uint32_t sp = 1; sp <= max_splits; sp++) {
        const uint64_t units = uint64_t(tiles) * sp;
        const uint64_t waves = (units + wave_ctas - 1) / wave_ctas;
        const double eff = double(units) / double(waves * wave_ctas);
        if (eff > best_eff + 0.02) best_eff = eff, best_splits = sp;
      }
      *kbps = (g->kblocks + best_splits - 1) / best_splits;
      *splits = (g->kblocks + *kbps - 1) / *kbps;
    };
    split_k(g->tiles_n, wave, &g->kbps, &g->splits);
    if ((rc = make_map(&g->map_b, g->b8, k, g->kp, n, g->nb, bn)) != CHPIR_OK) break;
    // CTA pairs (cta_group::2): the default; CHPIR_GEMM_KERNEL=1sm keeps the one-SM kernel (the measured comparison)
    if (const char *v = std::getenv("CHPIR_GEMM_KERNEL"); v && v[0] == '1') g->pair = false;
    {
      constexpr int psmem1 = pair_stages<1>() * pair_stage_bytes<1>() + 1024, psmem2 = pair_stages<2>() * pair_stage_bytes<2>() + 1024;
      if (cudaFuncSetAttribute(gemm_tc_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem1) != cudaSuccess ||
          cudaFuncSetAttribute(gemm_tc_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem2) != cudaSuccess) {
        rc = CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
        break;
      }
      g->p_ctas = 2 * ((n + 255) / 256);
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(g->p_ctas, 1, 1), cfg.blockDim = dim3(kThreads, 1, 1);
      cfg.dynamicSmemBytes = g->nb == 1 ? psmem1 : psmem2;
      cudaLaunchAttribute at{};
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = 2, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
      cfg.attrs = &at, cfg.numAttrs = 1;
      int clusters = 0;
      const cudaError_t e = g->nb == 1 ? cudaOccupancyMaxActiveClusters(&clusters, gemm_tc_pair_kernel<1>, &cfg)
                                       : cudaOccupancyMaxActiveClusters(&clusters, gemm_tc_pair_kernel<2>, &cfg);
      if (e != cudaSuccess || clusters < 1) {
        (void)cudaGetLastError();
        clusters = sm_count / 2;
      }
      split_k(g->p_ctas, uint32_t(clusters) * 2u, &g->p_kbps, &g->p_splits);
      if ((rc = make_map(&g->map_d, g->b8, k, g->kp, n, g->nb, 128)) != CHPIR_OK) break;
      for (int i = 0; i < 2 && rc == CHPIR_OK; i++)
        for (int h = 0; h < 4 && rc == CHPIR_OK; h++)
          rc = make_map(&g->map_ah[i][h], g->a_ring + i * gemm_tc_panel_bytes(g), k, g->kp, BM, 4, 16u * uint32_t(h + 1));
      if (rc != CHPIR_OK) break;
    }
    for (int i = 0; i < 2; i++) {
      if ((rc = make_map(&g->map_a[i], g->a_ring + i * gemm_tc_panel_bytes(g), k, g->kp, BM, 4, BM)) != CHPIR_OK) break;
      if ((rc = make_map(&g->map_a64[i], g->a_ring + i * gemm_tc_panel_bytes(g), k, g->kp, BM, 4, 64)) != CHPIR_OK) break;
    }
    if (rc != CHPIR_OK) break;
  } while (false);
  if (rc != CHPIR_OK) {
    if (rc != CHPIR_ERR_INVALID_ARGUMENT) set_last_cuda_error(cudaGetLastError(), "gemm_tc_prepare");
    delete g;
    return rc;
  }
  *out = g;
  return CHPIR_OK;
}

// C_panel[rows x n] += A_panel . B  for the panel held in ring buffer `buf` (rows <= 128).  C must have been zeroed.
int gemm_tc_panel(const GemmTcB *g, int buf, uint32_t rows, uint32_t *C_panel, cudaStream_t s) {
  if (g->pair) {
    if (rows == 0) return CHPIR_OK;
    if (rows > uint32_t(BM)) return CHPIR_ERR_INVALID_ARGUMENT;
    const uint32_t mq = (rows + 31) / 32 * 32;  // rows of A per pass: an N = mq MMA of a CTA pair needs mq % 32 == 0
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(g->p_ctas, g->p_splits, 1), cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = g->nb == 1 ? pair_stages<1>() * pair_stage_bytes<1>() + 1024 : pair_stages<2>() * pair_stage_bytes<2>() + 1024;
    cfg.stream = s;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
    cfg.attrs = &at, cfg.numAttrs = 1;
    const CUtensorMap &ah = g->map_ah[buf][mq / 32 - 1];
    const cudaError_t e = g->nb == 1 ? cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel<1>, ah, g->map_d, C_panel, rows, g->n, mq, g->kblocks, g->p_kbps)
                                     : cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel<2>, ah, g->map_d, C_panel, rows, g->n, mq, g->kblocks, g->p_kbps);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
    }
    return CHPIR_OK;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(g->tiles_n, g->splits, 1), cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = STAGES * (4 * A_TILE + int(g->nb) * B_TILE) + 1024;
  cfg.stream = s;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = g->csz, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
  cfg.attrs = &at, cfg.numAttrs = g->csz > 1 ? 1 : 0;
  const cudaError_t e =
      g->nb == 1 ? cudaLaunchKernelEx(&cfg, gemm_tc_kernel<1>, g->map_a[buf], g->map_a64[buf], g->map_b, C_panel, rows, g->n, g->tiles_n, g->bn, g->kblocks, g->kbps, g->csz)
                 : cudaLaunchKernelEx(&cfg, gemm_tc_kernel<2>, g->map_a[buf], g->map_a64[buf], g->map_b, C_panel, rows, g->n, g->tiles_n, g->bn, g->kblocks, g->kbps, g->csz);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
  }
  return CHPIR_OK;
}

// u32 rows [row0, row0 + rows) of A (row-major, k columns) -> limb planes of ring buffer `buf`.
int gemm_tc_load_panel_u32(const GemmTcB *g, int buf, const uint32_t *A_rows, uint32_t rows, cudaStream_t s) {
  const uint64_t total = uint64_t(rows) * (g->kp / 4);
  const uint64_t want = (total + 255) / 256;
  split_a_limbs<<<unsigned(want < 148ull * 16 ? (want ? want : 1) : 148ull * 16), 256, 0, s>>>(A_rows, rows, g->k, g->kp, BM,
                                                                                             g->a_ring + buf * gemm_tc_panel_bytes(g));
  return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}

// Whole product for operands already in device memory as u32 (chpir_matmul and the debug paths).
int launch_gemm_tc(const uint32_t *A, const uint32_t *B, uint32_t ldb, uint32_t *C, uint32_t m, uint64_t k, uint32_t n, uint32_t b_bits,
                   int sm_count, cudaStream_t s, float *kernel_ms) {
  if (kernel_ms) *kernel_ms = 0.f;
  GemmTcB *g = nullptr;
  if (int rc = gemm_tc_prepare(B, ldb, k, n, b_bits, sm_count, s, &g); rc != CHPIR_OK) return rc;
  int rc = CHPIR_OK;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float total_ms = 0.f;
  do {
    if (cudaMemsetAsync(C, 0, uint64_t(m) * n * 4, s) != cudaSuccess) {
      rc = CHPIR_ERR_CUDA_TRANSFER_FAILED;
      break;
    }
    for (uint32_t r0 = 0, p = 0; r0 < m && rc == CHPIR_OK; r0 += BM, p++) {
      const uint32_t rows = m - r0 < uint32_t(BM) ? m - r0 : uint32_t(BM);
      if ((rc = gemm_tc_load_panel_u32(g, p & 1, A + uint64_t(r0) * k, rows, s)) != CHPIR_OK) break;
      cudaEventRecord(e0, s);
      if ((rc = gemm_tc_panel(g, p & 1, rows, C + uint64_t(r0) * n, s)) != CHPIR_OK) break;
      cudaEventRecord(e1, s);
      if (kernel_ms) {
        cudaEventSynchronize(e1);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) total_ms += ms;
      }
    }
    if (rc != CHPIR_OK) break;
    cudaError_t e = cudaStreamSynchronize(s);  // the planes are freed below
    if (e != cudaSuccess) {
      set_last_cuda_error(e, "gemm_tc_kernel");
      rc = CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
    }
  } while (false);
  if (rc != CHPIR_OK && rc != CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED) set_last_cuda_error(cudaGetLastError(), "gemm_tc");
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  gemm_tc_free(g);
  if (kernel_ms) *kernel_ms = total_ms;
  return rc;
}

}  // namespace chpir
