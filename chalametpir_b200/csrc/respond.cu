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
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace chpir {

static uint32_t env_u32_early(const char *name, uint32_t dflt) {
  const char *v = std::getenv(name);
  return v && *v ? uint32_t(std::strtoul(v, nullptr, 10)) : dflt;
}

PackedLayout make_layout(uint32_t b, uint32_t ncols) {
  PackedLayout L{};
  L.b = b;
  L.fpw = 64 / b;
  const uint32_t words = (ncols + L.fpw - 1) / L.fpw;
  L.ncols = ncols;
  // Tight rows (CHPIR_TIGHT_PITCH=1, off by default): an odd word count is not rounded up to 16 bytes -- 17 words instead of 18 at
  // 8-way sharding of N = 940, 5.6 % fewer bytes per query.  Measured on a 118-column slice it does not pay: 27.0-28.4 us per
  // query against 24.9-26.3 us padded (profiles/r1_slice_sweep_tight.txt) -- at that width the kernel is bound by the consumers'
  // per-row work (the same 9 units per row either way, plus a second 8-byte load and 2-way bank conflicts), not by the bytes.
  // Only the ring kernel reads this layout, so it is tied to it.
  const bool tight = (words & 1u) && words <= 64 && env_u32_early("CHPIR_RESPOND_RING", 1) != 0 && env_u32_early("CHPIR_TIGHT_PITCH", 0) != 0;
  L.tight = tight ? 1 : 0;
  L.units = (words + 1) / 2;
  return L;
}

bool make_layout_explicit(uint32_t b, uint32_t ncols, uint32_t units, uint32_t tight, PackedLayout *out) {
  if (b < 4 || b > 14 || ncols == 0 || tight > 1) return false;
  PackedLayout L{};
  L.b = b, L.fpw = 64 / b, L.ncols = ncols, L.units = units, L.tight = tight;
  const uint32_t words = L.words();
  if (units != (words + 1) / 2 || (tight && !((words & 1u) && words <= 64))) return false;
  *out = L;
  return true;
}

namespace {

uint32_t env_u32(const char *name, uint32_t dflt) {
  const char *v = std::getenv(name);
  return v && *v ? uint32_t(std::strtoul(v, nullptr, 10)) : dflt;
}

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

// ------------------------------------------------------------------------------------------------------------------
// Production kernel: persistent CTAs, one per SM, each streaming a contiguous K-range of the packed matrix through a
// shared-memory ring filled by the bulk-copy engine (cp.async.bulk + mbarrier complete_tx), so that the bytes in
// flight per SM are bounded by shared memory (~190 KB) instead of by registers.  One producer lane issues the copies
// (matrix rows and the matching slice of q); the consumer threads own (unit, row-lane) pairs as in respond_kernel and
// read 16-byte vectors from shared memory, conflict-free.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

constexpr int kRingMaxThreads = 1024;
constexpr int kMaxDevices = 64;

__device__ __forceinline__ void bar_sync_consumers(uint32_t nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// One CTA per SM; CTA (x, y) owns k-rows [x*rows_per_cta, ...) of queries y*q_per_cta ... (q_per_cta = 1 in production, see
// respond_ring_dispatch).  Within a CTA the producer lane never stops at a query boundary -- the chunk sequence is (query, chunk)
// flattened -- and the epilogue of a query (fold the R row lanes, one atomic per column) borrows the last stage the consumers read
// as scratch: two consumer-only barriers and a few shuffles instead of 2*FPW block-wide reduction rounds (at 8-way column sharding
// a query is ~25 us of streaming per rank, and the old epilogue was ~3 us of it).
template <int B, int RPT, bool TIGHT>
__global__ void __launch_bounds__(kRingMaxThreads, 1)
    respond_ring_kernel(const uint8_t *__restrict__ packed, const uint32_t *__restrict__ q_all, uint32_t *__restrict__ resp_all, uint64_t K,
                        uint32_t pitch, uint32_t units, uint32_t R, uint32_t stages, uint32_t stage_bytes, uint64_t rows_per_cta, uint32_t ncols, uint32_t q_bulk,
                        uint32_t nq, uint32_t q_per_cta) {
  constexpr int FPW = 64 / B;
  constexpr int NACC = 2 * FPW;
  // accumulators folded per epilogue pass: the scratch is the last stage of the query, S*pitch bytes = at least RPT*16 bytes per
  // consumer thread (RPT*8 when rows are tight: pitch >= 8 * units)
  constexpr int G = TIGHT ? (NACC < RPT * 2 ? NACC : RPT * 2) : (NACC < RPT * 4 ? NACC : RPT * 4);
  extern __shared__ __align__(128) uint8_t ring[];
  const uint32_t S = R * RPT;  // rows per stage
  const uint32_t d_bytes = S * pitch;
  const uint32_t ring_base = smem_addr(ring);
  const uint32_t bar_base = ring_base + stages * stage_bytes;  // full[stages], empty[stages]
  const uint32_t t = threadIdx.x;
  const uint32_t n_cons = R * units;
  const uint32_t cons_warps = (n_cons + 31) / 32;
  const uint32_t warp = t / 32, lane = t % 32;

  const uint32_t q_begin = blockIdx.y * q_per_cta;
  const uint32_t q_end = q_begin + q_per_cta < nq ? q_begin + q_per_cta : nq;
  const uint64_t k0 = uint64_t(blockIdx.x) * rows_per_cta;
  uint64_t k1 = k0 + rows_per_cta;
  if (k1 > K) k1 = K;
  const uint32_t n_chunks = k0 < k1 ? uint32_t((k1 - k0 + S - 1) / S) : 0u;
  if (n_chunks == 0 || q_begin >= q_end) return;  // uniform per CTA

  if (t == 0) {
    for (uint32_t s = 0; s < stages; s++) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (stages + s), cons_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == cons_warps) {
    // ------------------------------------------------------------------ producer (one lane)
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      uint32_t g = 0;  // chunks issued so far, across queries
      for (uint32_t qi = q_begin; qi < q_end; qi++) {
        const uint32_t *q = q_all + uint64_t(qi) * K;
        for (uint32_t c = 0; c < n_chunks; c++, g++) {
          const uint32_t s = g % stages, ph = (g / stages) & 1;
          mbar_wait(bar_base + 8 * (stages + s), ph ^ 1);
          const uint64_t kc = k0 + uint64_t(c) * S;
          const uint32_t rows = uint32_t(k1 - kc < S ? k1 - kc : S);
          const bool qb = q_bulk && (rows & 3u) == 0;  // 16-byte granules; a partial chunk's q slice travels the same way
          const uint32_t full = bar_base + 8 * s;
          // tight rows are 8 (mod 16) bytes long: copy an even number of them (S is even; past row K lies the zeroed pad row)
          const uint32_t rows_cp = TIGHT ? rows + (rows & 1u) : rows;
          mbar_expect_tx(full, rows_cp * pitch + (qb ? rows * 4 : 0));
          const uint32_t dst = ring_base + s * stage_bytes;
          bulk_g2s(dst, packed + kc * pitch, rows_cp * pitch, full, policy);
          if (qb) bulk_g2s(dst + d_bytes, q + kc, rows * 4, full);
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const uint32_t u = t % units, r = t / units;
  const bool active = t < n_cons;
  const uint32_t n_sync = cons_warps * 32;
  uint32_t g = 0;
  for (uint32_t qi = q_begin; qi < q_end; qi++) {
    const uint32_t *q = q_all + uint64_t(qi) * K;
    uint32_t acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0;
    uint32_t last_s = 0;
    for (uint32_t c = 0; c < n_chunks; c++, g++) {
      const uint32_t s = g % stages, ph = (g / stages) & 1;
      const uint64_t kc = k0 + uint64_t(c) * S;
      const uint32_t rows = uint32_t(k1 - kc < S ? k1 - kc : S);
      const bool qb = q_bulk && (rows & 3u) == 0;
      mbar_wait(bar_base + 8 * s, ph);
      if (active) {
        const uint32_t base = ring_base + s * stage_bytes + u * 16 + r * pitch;
        const uint32_t qbase = ring_base + s * stage_bytes + d_bytes + r * 4;
        uint4 w[RPT];
        uint32_t qk[RPT];
        // shared-memory reads are unconditional (rows past the end of a partial chunk hold stale bytes and are
        // multiplied by zero), so that all RPT loads of a stage are in flight together.  Tight rows start on 8-byte boundaries:
        // two 8-byte loads; the second word of a row's last unit is the first word of the next row (or of the q slice behind the
        // last row) and only ever feeds accumulators of columns >= ncols, which are never published.
#pragma unroll
        for (int j = 0; j < RPT; j++) {
          if (!TIGHT) {
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w[j].x), "=r"(w[j].y), "=r"(w[j].z), "=r"(w[j].w)
                         : "r"(base + j * R * pitch));
          } else {
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[j].x), "=r"(w[j].y) : "r"(base + j * R * pitch));
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[j].z), "=r"(w[j].w) : "r"(base + j * R * pitch + 8));
          }
        }
        if (qb) {
#pragma unroll
          for (int j = 0; j < RPT; j++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(qk[j]) : "r"(qbase + j * R * 4));
          if (rows != S) {  // partial chunk (uniform per CTA): the words past its end are stale
#pragma unroll
            for (int j = 0; j < RPT; j++)
              if (r + j * R >= rows) qk[j] = 0u;
          }
        } else {
#pragma unroll
          for (int j = 0; j < RPT; j++) {
            const uint32_t row = r + j * R;
            qk[j] = row < rows ? __ldg(q + kc + row) : 0u;
          }
        }
#pragma unroll
        for (int j = 0; j < RPT; j++) {
          fma_fields<B>(w[j].x, w[j].y, qk[j], acc);
          fma_fields<B>(w[j].z, w[j].w, qk[j], acc + FPW);
        }
        // The stage may be overwritten by the bulk-copy engine as soon as this warp has arrived on `empty` below.  Pin the
        // multiply-adds that consume every word loaded from it in front of that arrive: instructions issue in order, so the arrive
        // then cannot issue before the shared-memory loads have delivered their data (the write-after-read side of the ring; the
        // read-after-write side is the acquire of the `full` barrier wait).  compute-sanitizer's racecheck does not model either
        // mbarrier edge for async-proxy writes and reports the pair (profiles/r2_sanitizer.txt).
        asm volatile("" : "+r"(acc[0]), "+r"(acc[FPW]));
      }
      __syncwarp();
      // the last stage of a query is handed back only after the epilogue, which uses it as scratch
      if (c + 1 < n_chunks) {
        if (lane == 0) mbar_arrive(bar_base + 8 * (stages + s));
      } else {
        last_s = s;
      }
    }
    unmask_fields<B>(acc);
    unmask_fields<B>(acc + FPW);

    // Fold the R row lanes of every (unit, accumulator) and publish one atomic per (CTA, column).  Scratch = the stage just
    // consumed, G accumulators per pass; every (accumulator, unit) item is summed by P
    // adjacent lanes (R/P values each) and finished with shuffles.
    uint32_t *red = reinterpret_cast<uint32_t *>(ring + last_s * stage_bytes);
    uint32_t *resp = resp_all + uint64_t(qi) * ncols;
    const uint32_t P = (R % 4 == 0 && G * units * 4 <= n_sync) ? 4u : (R % 2 == 0 && G * units * 2 <= n_sync) ? 2u : 1u;
    const uint32_t per = R / P;
#pragma unroll
    for (int g0 = 0; g0 < NACC; g0 += G) {
      bar_sync_consumers(n_sync);  // everyone is done reading the stage (first pass) / the previous pass's sums
      if (active) {
#pragma unroll
        for (int i = 0; i < G; i++)
          if (g0 + i < NACC) red[i * n_cons + t] = acc[g0 + i];
      }
      bar_sync_consumers(n_sync);
      const uint32_t n_items = G * units;
      for (uint32_t base = 0; base < n_items; base += n_sync / P) {  // same trip count for every consumer thread (full-warp shuffles)
        const uint32_t it = base + t / P, part = t % P;
        const bool valid = it < n_items;
        const uint32_t i = valid ? it / units : 0u, uu = valid ? it % units : 0u;
        uint32_t sum = 0;
        if (valid)
          for (uint32_t rr = part * per; rr < (part + 1) * per; rr++) sum += red[i * n_cons + rr * units + uu];
        if (P >= 2) sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        if (P == 4) sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const uint32_t a = g0 + i;
        const uint32_t col = (2 * uu + a / FPW) * FPW + (a % FPW);
        if (valid && part == 0 && a < NACC && col < ncols && sum != 0) atomicAdd(resp + col, sum);
      }
    }
    bar_sync_consumers(n_sync);  // the scratch is dead: hand the stage back to the producer
    if (lane == 0) mbar_arrive(bar_base + 8 * (stages + last_s));
  }
}

// One thread per u64 word of the packed matrix (row k, word w -> columns w*FPW ...): works for both row layouts, the words of a
// row are contiguous and rows follow each other at `words` u64.
template <int B>
__global__ void pack_kernel(const uint32_t *__restrict__ d, uint64_t K, uint32_t ld, uint32_t col_begin, uint32_t ncols, uint32_t words,
                            uint64_t *__restrict__ packed) {
  constexpr int FPW = 64 / B;
  constexpr uint32_t MASK = (1u << B) - 1u;
  const uint64_t total = K * words;
  for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < total; idx += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t k = idx / words;
    const uint32_t wi = uint32_t(idx - k * words);
    const uint32_t *row = d + k * ld + col_begin;
    uint64_t w = 0;
#pragma unroll
    for (int f = 0; f < FPW; f++) {
      const uint32_t col = wi * FPW + f;
      if (col < ncols) w |= uint64_t(row[col] & MASK) << (f * B);  // `& mat_elem_mask` as matrix.rs:121-156
    }
    packed[idx] = w;
  }
}

// packed rows -> K x ncols u32 (the inverse of pack_kernel; used to rebuild the GEMM's limb planes for a server loaded from disk)
template <int B>
__global__ void unpack_kernel(const uint64_t *__restrict__ packed, uint64_t K, uint32_t ncols, uint32_t words, uint32_t *__restrict__ d) {
  constexpr int FPW = 64 / B;
  constexpr uint32_t MASK = (1u << B) - 1u;
  const uint64_t total = K * words;
  for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < total; idx += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t k = idx / words;
    const uint32_t wi = uint32_t(idx - k * words);
    const uint64_t w = packed[idx];
    uint32_t *row = d + k * ncols;
#pragma unroll
    for (int f = 0; f < FPW; f++) {
      const uint32_t col = wi * FPW + f;
      if (col < ncols) row[col] = uint32_t(w >> (f * B)) & MASK;
    }
  }
}

template <int B>
int unpack_dispatch(const uint8_t *packed, const PackedLayout &L, uint64_t K, uint32_t *d, cudaStream_t s) {
  const uint32_t wpr = uint32_t(L.pitch_bytes() / 8);  // u64 words per row, padding included
  const uint64_t total = K * wpr;
  const int block = 256;
  const uint64_t want = (total + block - 1) / block;
  const int grid = int(want < 148ull * 32 ? (want ? want : 1) : 148ull * 32);
  unpack_kernel<B><<<grid, block, 0, s>>>(reinterpret_cast<const uint64_t *>(packed), K, L.ncols, wpr, d);
  return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
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
int respond_ring_dispatch(const uint8_t *packed, const PackedLayout &L, uint64_t K, const RespondPlan &P, const uint32_t *q, uint32_t *resp,
                          uint32_t nq, cudaStream_t s) {
  // bulk copies of q need 16-byte aligned sources for every query of the batch
  const uint32_t q_bulk = (reinterpret_cast<uintptr_t>(q) % 16 == 0 && (nq == 1 || (K * 4) % 16 == 0)) ? 1u : 0u;
  auto launch = [&](auto kernel) -> int {
    // Function attributes are per device: a thread that serves several GPUs of one process (the cluster of csrc/cluster.cu, or two
    // servers on two devices) must raise the limit on each of them.  Cache = last kernel configured per device ordinal.
    static thread_local const void *configured[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
    if (dev >= kMaxDevices || configured[dev] != reinterpret_cast<const void *>(kernel)) {
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
      if (dev < kMaxDevices) configured[dev] = reinterpret_cast<const void *>(kernel);
    }
    // One grid row per query (q_per_cta = 1): the hardware scheduler hands the next query's CTAs to whichever SMs finish first.
    // Looping over several queries inside one CTA lifetime (CHPIR_RING_Q_PER_CTA > 1) saves the per-query pipeline fill but fixes
    // the K-range -> SM assignment for the whole batch, and SMs do not stream at equal rates: measured on a 118-column slice
    // (8-way sharding) 24.7 us per query at 1, 26.0 at 4, 27.3 at 16 (profiles/r1_slice_sweep.txt).
    const uint32_t q_per_cta = std::max(1u, std::min(env_u32("CHPIR_RING_Q_PER_CTA", 1), nq));
    kernel<<<dim3(P.ring_grid, (nq + q_per_cta - 1) / q_per_cta), P.ring_block, P.ring_smem_bytes, s>>>(
        packed, q, resp, K, uint32_t(L.pitch_bytes()), L.units, P.ring_R, P.ring_stages, P.ring_stage_bytes, P.ring_rows_per_cta, L.ncols, q_bulk, nq, q_per_cta);
    return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
  };
  if (L.tight) {
    switch (P.ring_rpt) {
      case 1: return launch(respond_ring_kernel<B, 1, true>);
      case 2: return launch(respond_ring_kernel<B, 2, true>);
      default: return launch(respond_ring_kernel<B, 4, true>);
    }
  }
  switch (P.ring_rpt) {
    case 1: return launch(respond_ring_kernel<B, 1, false>);
    case 2: return launch(respond_ring_kernel<B, 2, false>);
    default: return launch(respond_ring_kernel<B, 4, false>);
  }
}

template <int B>
int pack_dispatch(const uint32_t *d, uint64_t K, uint32_t ld, uint32_t col_begin, const PackedLayout &L, uint8_t *packed, cudaStream_t s) {
  const uint32_t wpr = uint32_t(L.pitch_bytes() / 8);  // u64 words per row, padding included (written as zero)
  const uint64_t total = K * wpr;
  const int block = 256;
  const uint64_t want = (total + block - 1) / block;
  const int grid = int(want < 148ull * 64 ? (want ? want : 1) : 148ull * 64);
  pack_kernel<B><<<grid, block, 0, s>>>(d, K, ld, col_begin, L.ncols, wpr, reinterpret_cast<uint64_t *>(packed));
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


// Geometry of the bulk-copy ring kernel.  Tunables can be overridden from the environment for experiments:
// CHPIR_RESPOND_RING=0 disables it, CHPIR_RING_R / CHPIR_RING_RPT / CHPIR_RING_STAGES / CHPIR_RING_BUDGET_KB (shared memory per CTA) / CHPIR_RING_GRID_MULT (CTAs per SM) / CHPIR_RING_Q_PER_CTA.
static void plan_ring(const PackedLayout &L, uint64_t K, int sm_count, RespondPlan *P) {
  P->ring = 0;
  if (env_u32("CHPIR_RESPOND_RING", 1) == 0) return;
  const uint32_t units = L.units;
  if (units == 0 || 4 * units > uint32_t(kRingMaxThreads) - 32) return;  // a row must fit the consumer threads of one CTA
  uint32_t R = env_u32("CHPIR_RING_R", 0);
  if (R == 0) {
    R = (640 / units) & ~3u;
    if (R < 4) R = 4;
    if (R > 64) R = 64;
  }
  R = (R + 3) & ~3u;
  while (R > 4 && R * units > uint32_t(kRingMaxThreads) - 32) R -= 4;
  uint32_t rpt = env_u32("CHPIR_RING_RPT", 4);
  if (rpt != 1 && rpt != 2) rpt = 4;
  const uint32_t pitch = uint32_t(L.pitch_bytes());
  const uint32_t budget = std::min(env_u32("CHPIR_RING_BUDGET_KB", 200), 200u) * 1024;
  while (rpt > 1 && 3 * R * rpt * (pitch + 4) > budget) rpt /= 2;
  const uint32_t S = R * rpt;
  const uint32_t stage_bytes = S * (pitch + 4);
  uint32_t stages = budget / stage_bytes;
  const uint32_t want = env_u32("CHPIR_RING_STAGES", 0);
  if (want && want < stages) stages = want;
  if (stages > 32) stages = 32;
  if (stages < 2) return;
  P->ring = 1;
  P->ring_R = R;
  P->ring_rpt = rpt;
  P->ring_stages = stages;
  P->ring_stage_bytes = stage_bytes;
  P->ring_smem_bytes = std::max<uint32_t>(stages * stage_bytes + 16 * stages, 4096 + 64);
  P->ring_grid = uint32_t(sm_count) * std::max(1u, env_u32("CHPIR_RING_GRID_MULT", 1));
  const uint64_t per = (K + P->ring_grid - 1) / P->ring_grid;
  P->ring_rows_per_cta = (per + S - 1) / S * S;
  // Short K-ranges (the row blocks of an 8-way cluster: ~1000 rows per CTA) lose whole SMs to that rounding -- 147 456 rows in
  // units of 32 are 144 CTAs of 1024 -- so when it costs more than 1 % the range is rounded to 4 rows instead (16-byte granules of
  // q, an even row count for tight rows) and every CTA ends on a partial stage.
  if (P->ring_rows_per_cta * 100 > per * 101 && env_u32("CHPIR_RING_FINE_SPLIT", 1) != 0) P->ring_rows_per_cta = (per + 3) / 4 * 4;
  if (P->ring_rows_per_cta == 0) P->ring_rows_per_cta = S;
  P->ring_block = ((R * units + 31) / 32 + 1) * 32;
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
  plan_ring(L, K, sm_count, &P);
  return P;
}

static int launch_respond_one(const uint8_t *packed, const PackedLayout &L, uint64_t K, const RespondPlan &P, const uint32_t *q_dev,
                              uint32_t *resp_dev, cudaStream_t s) {
#define CHPIR_RESP(B) respond_dispatch<B>(packed, L, K, P, q_dev, resp_dev, s)
  CHPIR_DISPATCH_B(L.b, CHPIR_RESP)
#undef CHPIR_RESP
}

int launch_respond(const uint8_t *packed, const PackedLayout &L, uint64_t K, const RespondPlan &P, const uint32_t *q_dev, uint32_t *resp_dev,
                   uint32_t nq, cudaStream_t s) {
  if (nq == 0) return CHPIR_OK;
  if (P.ring) {
    for (uint32_t i0 = 0; i0 < nq; i0 += 65535) {  // grid.y limit
      const uint32_t n = std::min(nq - i0, 65535u);
      const uint32_t *qd = q_dev + uint64_t(i0) * K;
      uint32_t *rd = resp_dev + uint64_t(i0) * L.ncols;
      const int rc = [&]() -> int {
#define CHPIR_RING(B) respond_ring_dispatch<B>(packed, L, K, P, qd, rd, n, s)
        CHPIR_DISPATCH_B(L.b, CHPIR_RING)
#undef CHPIR_RING
      }();
      if (rc != CHPIR_OK) return rc;
    }
    return CHPIR_OK;
  }
  if (L.tight) return CHPIR_ERR_INVALID_ARGUMENT;  // tight rows are only ever planned together with the ring kernel
  for (uint32_t i = 0; i < nq; i++)
    if (int rc = launch_respond_one(packed, L, K, P, q_dev + uint64_t(i) * K, resp_dev + uint64_t(i) * L.ncols, s); rc != CHPIR_OK) return rc;
  return CHPIR_OK;
}

int launch_pack(const uint32_t *d_dev, uint64_t K, uint32_t ld, uint32_t col_begin, const PackedLayout &L, uint8_t *packed, cudaStream_t s) {
#define CHPIR_PACK(B) pack_dispatch<B>(d_dev, K, ld, col_begin, L, packed, s)
  CHPIR_DISPATCH_B(L.b, CHPIR_PACK)
#undef CHPIR_PACK
}

int launch_unpack(const uint8_t *packed, const PackedLayout &L, uint64_t K, uint32_t *d_dev, cudaStream_t s) {
#define CHPIR_UNPACK(B) unpack_dispatch<B>(packed, L, K, d_dev, s)
  CHPIR_DISPATCH_B(L.b, CHPIR_UNPACK)
#undef CHPIR_UNPACK
}

}  // namespace chpir
