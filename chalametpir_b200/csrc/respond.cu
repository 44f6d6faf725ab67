// respond.cu -- the online path: resp = q . D mod 2^32 as an HBM-bandwidth-bound streaming GEMV over a K-major
// bit-packed copy of D, plus the pack kernel that builds that copy at setup.
//
// Replaces (reference, CPU): Matrix::transpose (matrix.rs:517-527), Matrix::row_wise_compress (matrix.rs:98-205) and
// Matrix::row_vector_x_compressed_transposed_matrix (matrix.rs:328-485).
//
// Layout choice.  The reference keeps D transposed (one row per output column) and packs 2/3/4 elements per u32 along
// K, because each rayon task streams one output's K-long dot product.  On the GPU the opposite orientation is the
// natural one: D stays K-major (as it arrives), a thread owns a fixed group of output columns and walks down K, so
//   * q[k] is one broadcast scalar per row instead of cf scalars per packed word,
//   * the accumulators live in registers for the whole kernel, and
//   * every load is a coalesced 16-byte vector from one contiguous stream.
// Fields are packed floor(64/b) per u64 (7 x 9 bit = 63/64 bits used at b=9, against 27/32 in the reference's
// 3-per-u32 layout), so one query streams 1.27 GB instead of 1.48 GB at 2^20 entries.
// Partial sums of different K-ranges are combined with u32 atomics: addition mod 2^32 is associative and
// commutative, so the result is bit-identical regardless of order.
#include "common.cuh"

namespace chpir {

PackedLayout make_layout(uint32_t b, uint32_t ncols) {
  PackedLayout L{};
  L.b = b;
  L.fpw = 64 / b;
  const uint32_t words = (ncols + L.fpw - 1) / L.fpw;
  L.units = (words + 1) / 2;
  L.ncols = ncols;
  return L;
}

namespace {

constexpr int kMaxThreads = 768;
constexpr int kUnroll = 4;

// Streaming 16-byte load (read-only path, no L1 allocation).  Deliberately NOT volatile: the compiler may hoist and
// batch these so that several are in flight per thread.
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// acc[f] += qk * G_f  with  G_f = (word >> f*B) mod 2^32  -- the field WITHOUT masking off the higher fields.
// Because  field_f = G_f - (G_{f+1} << B)  holds as an integer identity, and everything downstream is linear mod 2^32,
// the mask is applied once per kernel (unmask_fields) instead of once per element: one funnel shift + one IMAD per
// element in the hot loop.
template <int B>
__device__ __forceinline__ void fma_fields(uint32_t lo, uint32_t hi, uint32_t qk, uint32_t *acc) {
  constexpr int FPW = 64 / B;
#pragma unroll
  for (int f = 0; f < FPW; f++) {
    const int o = f * B;
    uint32_t g;
    if (o == 0)
      g = lo;
    else if (o < 32)
      g = __funnelshift_r(lo, hi, o);
    else if (o == 32)
      g = hi;
    else
      g = hi >> (o - 32);
    acc[f] += qk * g;
  }
}

// true_f = raw_f - (raw_{f+1} << B); the last field of a word has nothing above it (padding bits are zero).
template <int B>
__device__ __forceinline__ void unmask_fields(uint32_t *acc) {
  constexpr int FPW = 64 / B;
#pragma unroll
  for (int f = 0; f + 1 < FPW; f++) acc[f] -= acc[f + 1] << B;
}

// One block: unit chunk blockIdx.y (cu units starting at ub), k-rows [blockIdx.x*rows_per_block, ...).
// Thread t -> unit u = t % cu, row lane r = t / cu (< R).  Each thread keeps kUnroll 16-byte loads of the NEXT
// iteration in flight while it multiplies the current ones (register double buffering).
template <int B>
__global__ void __launch_bounds__(kMaxThreads) respond_kernel(const uint4 *__restrict__ packed, const uint32_t *__restrict__ q,
                                                                uint32_t *__restrict__ resp, uint64_t K, uint32_t units, uint32_t cu,
                                                                uint32_t R, uint32_t ncols, uint64_t rows_per_block) {
  constexpr int FPW = 64 / B;
  __shared__ uint32_t red[kMaxThreads];
  const uint32_t t = threadIdx.x;
  const uint32_t u = t % cu, r = t / cu;
  const uint32_t unit = blockIdx.y * cu + u;
  const bool active = (r < R) && (unit < units);

  uint32_t acc[2 * FPW];
#pragma unroll
  for (int i = 0; i < 2 * FPW; i++) acc[i] = 0;

  const uint64_t k0 = uint64_t(blockIdx.x) * rows_per_block;
  uint64_t k1 = k0 + rows_per_block;
  if (k1 > K) k1 = K;

  if (active && k0 + r < k1) {
    const uint64_t my_rows = (k1 - k0 - r + R - 1) / R;  // rows k0+r, k0+r+R, ... owned by this thread
    const uint64_t full = my_rows / kUnroll;
    const uint64_t step = uint64_t(R) * units;            // uint4 stride between my consecutive rows
    const uint4 *p = packed + (k0 + r) * units + unit;
    const uint32_t *pq = q + k0 + r;

    if (full > 0) {
      uint4 w[kUnroll];
      uint32_t qk[kUnroll];
#pragma unroll
      for (int j = 0; j < kUnroll; j++) {
        w[j] = ld_stream(p + j * step);
        qk[j] = __ldg(pq + uint64_t(j) * R);
      }
      for (uint64_t it = 0; it < full; it++) {
        // prefetch the next group; on the last trip re-read the current one (in bounds, result unused)
        const bool more = it + 1 < full;
        const uint4 *pn = more ? p + kUnroll * step : p;
        const uint32_t *pqn = more ? pq + uint64_t(kUnroll) * R : pq;
        uint4 nw[kUnroll];
        uint32_t nq[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
          nw[j] = ld_stream(pn + j * step);
          nq[j] = __ldg(pqn + uint64_t(j) * R);
        }
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
          fma_fields<B>(w[j].x, w[j].y, qk[j], acc);
          fma_fields<B>(w[j].z, w[j].w, qk[j], acc + FPW);
        }
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
          w[j] = nw[j];
          qk[j] = nq[j];
        }
        p = pn;
        pq = pqn;
      }
      p += kUnroll * step;  // p sat on the last full group
      pq += uint64_t(kUnroll) * R;
    }
    for (uint64_t i = full * kUnroll; i < my_rows; i++, p += step, pq += R) {
      const uint4 w = ld_stream(p);
      const uint32_t qk = __ldg(pq);
      fma_fields<B>(w.x, w.y, qk, acc);
      fma_fields<B>(w.z, w.w, qk, acc + FPW);
    }
    unmask_fields<B>(acc);
    unmask_fields<B>(acc + FPW);
  }

  // fold the R row lanes of each unit, then one atomic per (block, column)
#pragma unroll
  for (int i = 0; i < 2 * FPW; i++) {
    red[t] = active ? acc[i] : 0u;
    __syncthreads();
    if (r == 0 && unit < units) {
      uint32_t s = 0;
      for (uint32_t rr = 0; rr < R; rr++) s += red[rr * cu + u];
      const uint32_t col = (2 * unit + i / FPW) * FPW + (i % FPW);
      if (col < ncols && s != 0) atomicAdd(resp + col, s);
    }
    __syncthreads();
  }
}

template <int B>
__global__ void pack_kernel(const uint32_t *__restrict__ d, uint64_t K, uint32_t ld, uint32_t col_begin, uint32_t ncols, uint32_t units,
                            uint4 *__restrict__ packed) {
  constexpr int FPW = 64 / B;
  constexpr uint32_t MASK = (1u << B) - 1u;
  const uint64_t total = K * units;
  for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < total; idx += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t k = idx / units;
    const uint32_t u = uint32_t(idx - k * units);
    const uint32_t *row = d + k * ld + col_begin;
    uint64_t w[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
      for (int f = 0; f < FPW; f++) {
        const uint32_t col = (2 * u + h) * FPW + f;
        if (col < ncols) w[h] |= uint64_t(row[col] & MASK) << (f * B);  // `& mat_elem_mask` as matrix.rs:121-156
      }
    }
    packed[idx] = make_uint4(uint32_t(w[0]), uint32_t(w[0] >> 32), uint32_t(w[1]), uint32_t(w[1] >> 32));
  }
}

template <int B>
int respond_dispatch(const uint8_t *packed, const PackedLayout &L, uint64_t K, const RespondPlan &P, const uint32_t *q, uint32_t *resp,
                     cudaStream_t s) {
  const uint32_t cu = P.threads / P.rows_per_iter;  // units per chunk (plan stores threads = R * cu exactly)
  dim3 grid(P.grid, (L.units + cu - 1) / cu);
  const uint32_t block = ((P.threads + 31) / 32) * 32;
  respond_kernel<B><<<grid, block, 0, s>>>(reinterpret_cast<const uint4 *>(packed), q, resp, K, L.units, cu, P.rows_per_iter, L.ncols,
                                           P.rows_per_block);
  return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}

template <int B>
int pack_dispatch(const uint32_t *d, uint64_t K, uint32_t ld, uint32_t col_begin, const PackedLayout &L, uint8_t *packed, cudaStream_t s) {
  const uint64_t total = K * L.units;
  const int block = 256;
  const uint64_t want = (total + block - 1) / block;
  const int grid = int(want < 148ull * 32 ? (want ? want : 1) : 148ull * 32);
  pack_kernel<B><<<grid, block, 0, s>>>(d, K, ld, col_begin, L.ncols, L.units, reinterpret_cast<uint4 *>(packed));
  return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}

template <int B>
int occupancy_of(int threads) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, respond_kernel<B>, threads, 0) != cudaSuccess) {
    (void)cudaGetLastError();
    occ = 1;
  }
  return occ;
}

}  // namespace

#define CHPIR_DISPATCH_B(bits, CALL)                                        \
  switch (bits) {                                                           \
    case 4: return CALL(4);                                                 \
    case 5: return CALL(5);                                                 \
    case 6: return CALL(6);                                                 \
    case 7: return CALL(7);                                                 \
    case 8: return CALL(8);                                                 \
    case 9: return CALL(9);                                                 \
    case 10: return CALL(10);                                               \
    case 11: return CALL(11);                                               \
    case 12: return CALL(12);                                               \
    case 13: return CALL(13);                                               \
    case 14: return CALL(14);                                               \
    default: return CHPIR_ERR_IMPOSSIBLE_ENCODED_DB_MATRIX_ELEMENT_BIT_LENGTH; \
  }

RespondPlan plan_respond(const PackedLayout &L, uint64_t K, int sm_count) {
  RespondPlan P{};
  // column chunking only when one row has more units than a block has threads
  const uint32_t chunks = (L.units + kMaxThreads - 1) / kMaxThreads;
  const uint32_t cu = (L.units + chunks - 1) / chunks;
  auto occ_of = [&](uint32_t b, int threads) -> int {
#define CHPIR_OCC(B) occupancy_of<B>(threads)
    CHPIR_DISPATCH_B(b, CHPIR_OCC)
#undef CHPIR_OCC
  };
  // rows per iteration R: maximise resident ACTIVE threads per SM (register- and warp-granularity aware), prefer
  // smaller blocks on ties (finer K split, less tail)
  uint32_t best_r = 1;
  int best_occ = 1;
  long best_active = -1;
  for (uint32_t r = 1; r * cu <= uint32_t(kMaxThreads); r++) {
    const int th = int(((r * cu + 31) / 32) * 32);
    int occ = occ_of(L.b, th);
    if (occ < 1) occ = 1;
    const long active = long(occ) * long(r * cu);
    if (active > best_active + best_active / 50) best_active = active, best_r = r, best_occ = occ;
  }
  P.rows_per_iter = best_r;
  P.threads = best_r * cu;
  // one wave of resident blocks; each block owns a contiguous k-range
  uint64_t want_blocks = uint64_t(sm_count) * best_occ / chunks;
  if (want_blocks < 1) want_blocks = 1;
  const uint64_t quantum = uint64_t(P.rows_per_iter) * kUnroll;
  uint64_t rpb = (K + want_blocks - 1) / want_blocks;
  rpb = ((rpb + quantum - 1) / quantum) * quantum;
  if (rpb == 0) rpb = quantum;
  P.rows_per_block = rpb;
  P.grid = uint32_t((K + rpb - 1) / rpb);
  if (P.grid == 0) P.grid = 1;
  return P;
}

int launch_respond(const uint8_t *packed, const PackedLayout &L, uint64_t K, const RespondPlan &P, const uint32_t *q_dev, uint32_t *resp_dev,
                   cudaStream_t s) {
#define CHPIR_RESP(B) respond_dispatch<B>(packed, L, K, P, q_dev, resp_dev, s)
  CHPIR_DISPATCH_B(L.b, CHPIR_RESP)
#undef CHPIR_RESP
}

int launch_pack(const uint32_t *d_dev, uint64_t K, uint32_t ld, uint32_t col_begin, const PackedLayout &L, uint8_t *packed, cudaStream_t s) {
#define CHPIR_PACK(B) pack_dispatch<B>(d_dev, K, ld, col_begin, L, packed, s)
  CHPIR_DISPATCH_B(L.b, CHPIR_PACK)
#undef CHPIR_PACK
}

}  // namespace chpir
