// encode_dev.cu -- the row fill of Matrix::from_kv_database_with_{3,4}_wise_xor_filter (chalametpir_common/src/matrix.rs:707-746,
// :839-885) and encode_kv_as_row (serialization.rs:22-116) on the GPU.
//
// The host keeps what is inherently serial or tiny -- key digests, filter construction (peeling), the wave plan
// (host_encode.cpp: digest_and_peel, plan_fill_levels) -- and uploads the raw values (1.07 GB at 2^20 x 1 kB) instead of the
// encoded matrix D (4.4 GB).  One CTA builds one row of D:
//     own[e] = ( field_e(digest || value || 0x81) - D[o1][e] - D[o2][e] (- D[o3][e]) - mix(hash, e) ) & (2^b - 1)
// where field_e is the e-th b-bit field, LSB first, of the byte string (the reference's bit packing), o1..o3 are the key's other
// filter slots and mix is the reference's murmur finaliser (binary_fuse_filter.rs:553-560).  The rows a key reads belong to keys
// peeled later, so the keys are processed in dependency waves (a wave's members are mutually independent).
// D is written row-major u32, exactly the matrix the reference would hold, so everything downstream (pack, limb split, GEMM)
// is unchanged and the bytes can be compared with the host encoder one to one.
//
// Two schedules of the same arithmetic:
//   * column ownership (default): the part of a row that depends on the key alone (field_e - mix) is written for every key at once
//     (encode_rows_kernel), then ONE launch solves the dependent part.  The dependency is per column -- own[e] needs o1[e], o2[e] of
//     the SAME column e -- so a thread that owns a column and walks ALL keys in wave order only ever reads what it wrote itself:
//     program order replaces every inter-thread synchronisation.  N / 4 single-warp CTAs (235 at N = 940), each keeping 128 keys x
//     3 rows of loads in flight inside a wave; the 28 600 waves of a 2^20-entry filter cost one L2 round trip each instead of one
//     kernel launch each.
//   * one launch per wave, one CTA per key (round 1; CHPIR_FILL=waves): kept as the cross-check.
#include <cstdlib>

#include "common.cuh"

namespace chpir {
namespace {

__device__ __forceinline__ uint64_t fmix64_dev(uint64_t h) {
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdULL;
  h ^= h >> 33;
  h *= 0xc4ceb9fe1a85ec53ULL;
  h ^= h >> 33;
  return h;
}

struct FillArgs {
  const uint32_t *members;     // peel-order indices of this wave
  const uint64_t *order;       // hash per peel-order index
  const uint8_t *found;        // which slot the key owns
  const uint32_t *key_of_order;
  const uint8_t *digests;      // 32 bytes per key
  const uint8_t *values;       // value blob
  const uint64_t *val_off;     // n + 1 offsets
  uint32_t *D;                 // K x nc: columns [c0, c0 + nc) of the K x N matrix, row pitch nc
  uint64_t N;
  uint32_t c0, nc;
  uint32_t b;
  uint32_t segment_length, segment_count_length;
};

constexpr int kFillThreads = 256;

// binary_fuse_filter.rs:576-635 hash_batch_for_{3,4}_wise_xor_filter
template <int ARITY>
__device__ __forceinline__ void slots_dev(uint64_t hash, uint32_t sl, uint32_t scl, uint32_t h[4]) {
  const uint32_t m = sl - 1;
  h[0] = static_cast<uint32_t>(__umul64hi(hash, static_cast<uint64_t>(scl)));
  if (ARITY == 3) {
    h[1] = (h[0] + sl) ^ (static_cast<uint32_t>(hash >> 18) & m);
    h[2] = (h[0] + 2 * sl) ^ (static_cast<uint32_t>(hash) & m);
    h[3] = 0;
  } else {
    h[1] = (h[0] + sl) ^ (static_cast<uint32_t>(hash) & m);
    h[2] = (h[0] + 2 * sl) ^ (static_cast<uint32_t>(hash >> 16) & m);
    h[3] = (h[0] + 3 * sl) ^ (static_cast<uint32_t>(hash >> 32) & m);
  }
}

// dynamic shared memory: the key's byte string, zero padded to cover every field of the row (+4 bytes of slack for the 32-bit window)
template <int ARITY>
__global__ void __launch_bounds__(kFillThreads) fill_wave_kernel(FillArgs a, uint32_t stream_bytes) {
  extern __shared__ __align__(4) uint8_t sbytes[];
  const uint32_t i = a.members[blockIdx.x];
  const uint64_t hash = a.order[i];
  const uint32_t which = a.found[i];
  const uint32_t key = a.key_of_order[i];
  const uint64_t v0 = a.val_off[key], vlen = a.val_off[key + 1] - v0;
  // digest || value || 0x81 || 0...
  for (uint32_t t = threadIdx.x; t < stream_bytes; t += kFillThreads) {
    uint8_t v = 0;
    if (t < 32)
      v = a.digests[32ull * key + t];
    else if (t < 32 + vlen)
      v = a.values[v0 + (t - 32)];
    else if (t == 32 + vlen)
      v = 0x81;
    sbytes[t] = v;
  }
  uint32_t h[4];
  slots_dev<ARITY>(hash, a.segment_length, a.segment_count_length, h);
  uint32_t *own = a.D + uint64_t(h[which]) * a.nc;
  const uint32_t *o1 = a.D + uint64_t(h[(which + 1) % ARITY]) * a.nc;
  const uint32_t *o2 = a.D + uint64_t(h[(which + 2) % ARITY]) * a.nc;
  const uint32_t *o3 = a.D + uint64_t(h[(which + 3) % ARITY]) * a.nc;  // ARITY == 4 only
  const uint32_t mask = (1u << a.b) - 1;
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < a.nc; j += kFillThreads) {
    const uint64_t e = uint64_t(a.c0) + j;  // column of the whole matrix
    const uint32_t bit = uint32_t(e) * a.b, byte = bit >> 3, sh = bit & 7;
    const uint32_t w = uint32_t(sbytes[byte]) | uint32_t(sbytes[byte + 1]) << 8 | uint32_t(sbytes[byte + 2]) << 16;
    uint32_t v = (w >> sh) & mask;
    v -= __ldcg(o1 + j) + __ldcg(o2 + j) + static_cast<uint32_t>(fmix64_dev(hash + e));
    if (ARITY == 4) v -= __ldcg(o3 + j);
    own[j] = v & mask;
  }
}

// ---- column ownership ----------------------------------------------------------------------------------------------------------------
// The recurrence is linear:  own = ( t - D[o1] - D[o2] (- D[o3]) ) mod 2^b  with  t = field_e(digest || value || 0x81) - mix(hash, e),
// and t depends on nothing but the key.  So the fill runs as two kernels:
//   encode_rows_kernel    every key at once, no dependencies: its byte string goes through shared memory, D[own] = t & mask.  This is
//                         encode_kv_as_row + the hash mix, HBM bound (reads the values once, writes every owned row once: ~2 ms at 2^20).
//   solve_columns_kernel  the dependent part, D[own] -= D[o1] + D[o2] (+ D[o3]) in wave order: per key four row numbers (16 bytes),
//                         three or four loads, two or three subtractions, one store -- nothing else is left on the dependency chain.
struct FillRec {  // one key, in wave order: what the row encoding needs
  uint64_t hash, v0;
  uint32_t vlen, key, own, pad;
};
static_assert(sizeof(FillRec) == 32, "eight words per record");
static_assert(sizeof(FillRec) + sizeof(uint4) == kFillRecordBytes, "scratch per key: one record + one set of row numbers");

template <int ARITY>
__global__ void __launch_bounds__(256) fill_prep_kernel(FillArgs a, uint64_t count, FillRec *__restrict__ rec, uint4 *__restrict__ rows) {
  const uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (j >= count) return;
  const uint32_t i = a.members[j];
  const uint64_t hash = a.order[i];
  const uint32_t which = a.found[i], key = a.key_of_order[i];
  uint32_t h[4];
  slots_dev<ARITY>(hash, a.segment_length, a.segment_count_length, h);
  FillRec r;
  r.hash = hash, r.v0 = a.val_off[key], r.vlen = uint32_t(a.val_off[key + 1] - r.v0), r.key = key, r.own = h[which], r.pad = 0;
  rec[j] = r;
  // the row the key owns and the rows it reads; for 3-wise filters the fourth entry repeats the own row (never read)
  rows[j] = make_uint4(h[which], h[(which + 1) % ARITY], h[(which + 2) % ARITY], ARITY == 4 ? h[(which + 3) % ARITY] : h[which]);
}

// dynamic shared memory: a key's byte string (as in fill_wave_kernel); CTAs stride over the keys
__global__ void __launch_bounds__(kFillThreads) encode_rows_kernel(const FillRec *__restrict__ rec, uint64_t count, const uint8_t *__restrict__ digests,
                                                                   const uint8_t *__restrict__ values, uint32_t *__restrict__ D, uint32_t c0, uint32_t nc,
                                                                   uint32_t b, uint32_t stream_bytes) {
  extern __shared__ __align__(4) uint8_t sbytes[];
  const uint32_t mask = (1u << b) - 1;
  for (uint64_t j = blockIdx.x; j < count; j += gridDim.x) {
    const FillRec r = rec[j];
    for (uint32_t t = threadIdx.x; t < stream_bytes; t += kFillThreads) {
      uint8_t v = 0;
      if (t < 32)
        v = __ldg(digests + 32ull * r.key + t);
      else if (t - 32 < r.vlen)
        v = __ldg(values + r.v0 + (t - 32));
      else if (t - 32 == r.vlen)
        v = 0x81;
      sbytes[t] = v;
    }
    __syncthreads();
    uint32_t *own = D + uint64_t(r.own) * nc;
    for (uint32_t j = threadIdx.x; j < nc; j += kFillThreads) {
      const uint64_t e = uint64_t(c0) + j;  // column of the whole matrix
      const uint32_t bit = uint32_t(e) * b, byte = bit >> 3, sh = bit & 7;
      const uint32_t w = uint32_t(sbytes[byte]) | uint32_t(sbytes[byte + 1]) << 8 | uint32_t(sbytes[byte + 2]) << 16;
      own[j] = ((w >> sh) - static_cast<uint32_t>(fmix64_dev(r.hash + e))) & mask;
    }
    __syncthreads();
  }
}

constexpr int kFillCols = 4;                            // columns owned by a warp
constexpr int kFillGroups = 32 / kFillCols;             // keys a warp works on per slot: one per group of kFillCols lanes
constexpr int kFillUnroll = 16;                         // slots in flight
constexpr int kFillChunk = kFillGroups * kFillUnroll;   // keys in flight per warp (128)
constexpr int kFillRing = 4 * kFillChunk;               // row-number records staged in shared memory (8 KB)
static_assert(kFillChunk % 32 == 0, "the ring is refilled in whole warp loads");

// One warp per CTA owns kFillCols columns and walks ALL keys in wave order, so every value it reads was written by itself: program
// order replaces all synchronisation.  Its eight lane groups work on eight keys of the same wave per slot, sixteen slots (128 keys x
// 3-4 loads) are in flight inside a wave; a chunk never crosses a wave boundary, and a short wave only issues the slots it has.
// The wall time is (dependent steps) x (one round trip to L2 / HBM): 28 600 waves at 2^20 entries plus keys / 128 chunk steps --
// more warps do not shorten it, less work per step does.  Row numbers reach the lanes through a shared-memory ring that the warp
// refills 128 records at a time, two refills ahead of the walk (the loads are in flight for a whole step before they are stored).
template <int ARITY>
__global__ void __launch_bounds__(32) solve_columns_kernel(const uint4 *__restrict__ rows, const uint32_t *__restrict__ level_start, uint32_t waves,
                                                           uint32_t *D, uint64_t N /* columns of D as stored = its row pitch */, uint32_t b) {
  __shared__ uint4 ring[kFillRing];
  const uint32_t lane = threadIdx.x, group = lane / kFillCols;
  const uint64_t e = uint64_t(blockIdx.x) * kFillCols + (lane % kFillCols);
  const bool act = e < N;
  const uint64_t ec = act ? e : N - 1;  // inactive lanes of the last warp shadow a real column (loads only)
  const uint32_t mask = (1u << b) - 1;
  const uint64_t base = level_start[0], total = level_start[waves] - base;  // rows[0] is the first key of the first wave
  constexpr int kPer = kFillChunk / 32;

  auto fetch = [&](uint64_t j0, uint4 r[kPer]) {  // lane t: records j0 + t, j0 + 32 + t, ... (row 0 past the end: a safe address)
#pragma unroll
    for (int s = 0; s < kPer; s++) {
      const uint64_t j = j0 + 32 * s + lane;
      r[s] = j < total ? __ldg(rows + j) : make_uint4(0, 0, 0, 0);
    }
  };
  auto stash = [&](uint64_t j0, const uint4 r[kPer]) {
#pragma unroll
    for (int s = 0; s < kPer; s++) ring[(j0 + 32 * s + lane) % kFillRing] = r[s];
  };

  uint4 nx[kPer];
  for (int i = lane; i < kFillRing; i += 32) ring[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  uint64_t filled = 0;
  for (int i = 0; i < 2; i++) {
    fetch(filled, nx);
    stash(filled, nx);
    filled += kFillChunk;
  }
  __syncwarp();

  uint64_t j = 0;
  uint64_t end_next = level_start[waves > 1 ? 1 : waves] - base;  // wave ends are read one wave ahead (28 600 short waves at 2^20)
  for (uint32_t l = 0; l < waves; l++) {
    const uint64_t wave_end = l == 0 ? level_start[1] - base : end_next;
    end_next = level_start[l + 2 <= waves ? l + 2 : waves] - base;
    while (j < wave_end) {
      const uint32_t chunk = uint32_t(wave_end - j < kFillChunk ? wave_end - j : kFillChunk);
      const bool refill = filled < total && filled - j < 2 * kFillChunk;  // warp uniform
      if (refill) fetch(filled, nx);
      uint32_t v0[kFillUnroll], v1[kFillUnroll], v2[kFillUnroll], v3[kFillUnroll];
#pragma unroll
      for (int u = 0; u < kFillUnroll; u++) {
        if (kFillGroups * u < chunk) {  // warp uniform: a short wave issues only the slots it has
          // groups past the end of the chunk read a later key's (or an all-zero) record: valid rows, loads only
          const uint4 r = ring[(j + kFillGroups * u + group) % kFillRing];
          v0[u] = __ldcg(D + uint64_t(r.x) * N + ec);
          v1[u] = __ldcg(D + uint64_t(r.y) * N + ec);
          v2[u] = __ldcg(D + uint64_t(r.z) * N + ec);
          v3[u] = ARITY == 4 ? __ldcg(D + uint64_t(r.w) * N + ec) : 0u;
        }
      }
#pragma unroll
      for (int u = 0; u < kFillUnroll; u++) {
        if (kFillGroups * u < chunk) {
          const uint32_t own = ring[(j + kFillGroups * u + group) % kFillRing].x;
          if (act && kFillGroups * u + group < chunk) D[uint64_t(own) * N + e] = (v0[u] - v1[u] - v2[u] - v3[u]) & mask;
        }
      }
      j += chunk;
      if (refill) {
        __syncwarp();  // every lane has read the records this refill overwrites (all of them lie before j)
        stash(filled, nx);
        filled += kFillChunk;
        __syncwarp();
      }
    }
  }
}

}  // namespace

// All pointers except level_start_host are device pointers; D (K x nc u32: columns [c0, c0 + nc) of the K x N matrix) must be zeroed.  scratch_records: kFillRecordBytes per key, scratch_levels:
// waves + 1 words (both device memory; NULL selects the wave-per-launch schedule).  level_start_host must stay valid until the work
// enqueued on s has been synchronised (it is copied with cudaMemcpyAsync from pageable memory, i.e. during the call).
int launch_device_row_fill(uint32_t arity, const uint32_t *members, const uint32_t *level_start_host, uint32_t waves, const uint64_t *order,
                           const uint8_t *found, const uint32_t *key_of_order, const uint8_t *digests, const uint8_t *values,
                           const uint64_t *val_off, uint32_t *D, uint64_t N, uint32_t c0, uint32_t nc, uint32_t b, uint32_t segment_length,
                           uint32_t segment_count_length, void *scratch_records, uint32_t *scratch_levels, cudaStream_t s) {
  if (nc == 0 || uint64_t(c0) + nc > N) return CHPIR_ERR_INVALID_ARGUMENT;
  FillArgs a{};
  a.order = order;
  a.found = found;
  a.key_of_order = key_of_order;
  a.digests = digests;
  a.values = values;
  a.val_off = val_off;
  a.D = D;
  a.N = N;
  a.c0 = c0, a.nc = nc;
  a.b = b;
  a.segment_length = segment_length;
  a.segment_count_length = segment_count_length;
  uint64_t count = 0;
  for (uint32_t l = 0; l < waves; l++) count += level_start_host[l + 1] - level_start_host[l];
  // every field e < N reads bytes [e*b/8, e*b/8 + 2]
  const uint32_t stream_bytes = uint32_t(((N - 1) * b) / 8 + 3 + 3) & ~3u;
  if (stream_bytes > 200 * 1024) return CHPIR_ERR_INVALID_ARGUMENT;
  const char *mode = std::getenv("CHPIR_FILL");
  if (count > 0 && scratch_records && scratch_levels && !(mode && std::strcmp(mode, "waves") == 0)) {
    // column ownership: records in wave order, the independent row encoding, then one launch of N / 4 single-warp CTAs for the
    // dependent part.  The scratch (48 bytes per key + the wave table) comes from the caller, allocated before the host-side peeling:
    // a cudaMalloc here would queue behind whatever large allocation another thread of the setup is making (the 8.4 GB panel ring of
    // the XOF pipeline) with the GPU standing idle.
    FillRec *rec = static_cast<FillRec *>(scratch_records);
    uint4 *rows = reinterpret_cast<uint4 *>(rec + count);
    a.members = members + level_start_host[0];
    if (cudaMemcpyAsync(scratch_levels, level_start_host, (size_t(waves) + 1) * 4, cudaMemcpyHostToDevice, s) != cudaSuccess)
      return CHPIR_ERR_CUDA_TRANSFER_FAILED;
    if (stream_bytes > 48 * 1024 &&
        cudaFuncSetAttribute(encode_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(stream_bytes)) != cudaSuccess)
      return CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
    const unsigned pg = unsigned((count + 255) / 256), fg = unsigned((nc + kFillCols - 1) / kFillCols);
    const unsigned eg = unsigned(count < 148ull * 32 ? count : 148ull * 32);
    if (arity == 3)
      fill_prep_kernel<3><<<pg, 256, 0, s>>>(a, count, rec, rows);
    else
      fill_prep_kernel<4><<<pg, 256, 0, s>>>(a, count, rec, rows);
    encode_rows_kernel<<<eg, kFillThreads, stream_bytes, s>>>(rec, count, digests, values, D, c0, nc, b, stream_bytes);
    if (arity == 3)
      solve_columns_kernel<3><<<fg, 32, 0, s>>>(rows, scratch_levels, waves, D, nc, b);
    else
      solve_columns_kernel<4><<<fg, 32, 0, s>>>(rows, scratch_levels, waves, D, nc, b);
    return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
  }
  if (stream_bytes > 48 * 1024) {
    cudaError_t e = arity == 3 ? cudaFuncSetAttribute(fill_wave_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(stream_bytes))
                               : cudaFuncSetAttribute(fill_wave_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(stream_bytes));
    if (e != cudaSuccess) return CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
  }
  for (uint32_t l = 0; l < waves; l++) {
    const uint32_t cnt = level_start_host[l + 1] - level_start_host[l];
    if (!cnt) continue;
    a.members = members + level_start_host[l];
    if (arity == 3)
      fill_wave_kernel<3><<<cnt, kFillThreads, stream_bytes, s>>>(a, stream_bytes);
    else
      fill_wave_kernel<4><<<cnt, kFillThreads, stream_bytes, s>>>(a, stream_bytes);
  }
  return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}

}  // namespace chpir
