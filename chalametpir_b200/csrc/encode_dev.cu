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
//   * column ownership (default): ONE launch.  The dependency is per column -- own[e] needs o1[e], o2[e] of the SAME column e -- so
//     a thread that owns a column and walks ALL keys in wave order only ever reads what it wrote itself: program order replaces every
//     inter-thread synchronisation.  N / 4 single-warp CTAs (235 at N = 940), each keeping 64 keys x 3 rows of loads in flight
//     inside a wave; the 28 600 waves of a 2^20-entry filter cost one L2 round trip each instead of one kernel launch each.
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
  uint32_t *D;                 // K x N
  uint64_t N;
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
  uint32_t *own = a.D + uint64_t(h[which]) * a.N;
  const uint32_t *o1 = a.D + uint64_t(h[(which + 1) % ARITY]) * a.N;
  const uint32_t *o2 = a.D + uint64_t(h[(which + 2) % ARITY]) * a.N;
  const uint32_t *o3 = a.D + uint64_t(h[(which + 3) % ARITY]) * a.N;  // ARITY == 4 only
  const uint32_t mask = (1u << a.b) - 1;
  __syncthreads();
  for (uint64_t e = threadIdx.x; e < a.N; e += kFillThreads) {
    const uint32_t bit = uint32_t(e) * a.b, byte = bit >> 3, sh = bit & 7;
    const uint32_t w = uint32_t(sbytes[byte]) | uint32_t(sbytes[byte + 1]) << 8 | uint32_t(sbytes[byte + 2]) << 16;
    uint32_t v = (w >> sh) & mask;
    v -= __ldcg(o1 + e) + __ldcg(o2 + e) + static_cast<uint32_t>(fmix64_dev(hash + e));
    if (ARITY == 4) v -= __ldcg(o3 + e);
    own[e] = v & mask;
  }
}

// ---- column ownership ----------------------------------------------------------------------------------------------------------------
struct FillRec {  // one key, in wave order: everything the fill needs that does not depend on the column
  uint64_t hash, v0;
  uint32_t own, o1, o2, o3;  // rows of D: the slot the key owns and its other slots
  uint32_t vlen, key;
};
static_assert(sizeof(FillRec) == 40, "ten words per record");

template <int ARITY>
__global__ void __launch_bounds__(256) fill_prep_kernel(FillArgs a, uint64_t count, FillRec *__restrict__ rec) {
  const uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (j >= count) return;
  const uint32_t i = a.members[j];
  const uint64_t hash = a.order[i];
  const uint32_t which = a.found[i], key = a.key_of_order[i];
  uint32_t h[4];
  slots_dev<ARITY>(hash, a.segment_length, a.segment_count_length, h);
  FillRec r;
  r.hash = hash, r.v0 = a.val_off[key], r.vlen = uint32_t(a.val_off[key + 1] - r.v0), r.key = key;
  r.own = h[which], r.o1 = h[(which + 1) % ARITY], r.o2 = h[(which + 2) % ARITY], r.o3 = h[(which + 3) % ARITY];
  rec[j] = r;
}

constexpr int kFillCols = 4;                    // columns owned by a warp
constexpr int kFillGroups = 32 / kFillCols;     // keys a warp works on per slot: one per group of kFillCols lanes
constexpr int kFillUnroll = 8;                  // slots in flight
constexpr int kFillChunk = kFillGroups * kFillUnroll;  // keys in flight per warp: a whole chunk at once
constexpr int kFillRpl = kFillChunk / 32;       // records held per lane
static_assert(kFillChunk % 32 == 0 && 32 % kFillGroups == 0, "a slot's keys come from one record set");

// One warp per CTA owns kFillCols columns; its eight lane groups work on eight different keys of the same wave at a time (same
// columns, different rows), so a warp keeps 64 keys x (3 rows + 3 value bytes) of loads in flight.  Every warp has to walk ALL keys, so
// the wall time is (keys / keys in flight per warp) x the latency of those scattered reads (DRAM + TLB misses over a 4.4 GB matrix) --
// more warps do not shorten it, more keys per warp do.  Measured at 2^20 entries: 1.1-1.7 s with 8 keys in flight (32 columns per
// warp, value bytes combined inside the issue loop), 0.29 s with 16 keys (16 columns), 0.16-0.25 s with 32 keys (8 columns); narrower
// ownership costs sector efficiency on the row reads (16 of every 32-byte sector are used), which at ~0.2 TB/s is irrelevant.
// Records are fetched a chunk ahead (kFillRpl per lane, coalesced) and handed round by shuffles; a chunk never crosses a wave boundary.
// Every load is issued before the first result is used (the values of a slot are combined in the second loop only).
template <int ARITY>
__global__ void __launch_bounds__(32) fill_columns_kernel(const FillRec *__restrict__ rec, const uint32_t *__restrict__ level_start, uint32_t waves,
                                                          const uint8_t *__restrict__ digests, const uint8_t *__restrict__ values, uint32_t *D,
                                                          uint64_t N, uint32_t b) {
  const uint32_t lane = threadIdx.x, group = lane / kFillCols;
  const uint64_t e = uint64_t(blockIdx.x) * kFillCols + (lane % kFillCols);
  const bool act = e < N;
  const uint64_t ec = act ? e : N - 1;  // inactive lanes of the last warp shadow a real column (loads only)
  const uint32_t bit = uint32_t(ec) * b, byte0 = bit >> 3, sh = bit & 7, mask = (1u << b) - 1;
  const uint32_t *recw = reinterpret_cast<const uint32_t *>(rec);
  const uint64_t base = level_start[0], total = level_start[waves] - base;  // rec[0] is the first key of the first wave

  auto load_rec = [&](uint64_t j0, uint32_t r[kFillRpl][10]) {  // lane t: records j0 + t, j0 + 32 + t, ... (zeros past the end)
#pragma unroll
    for (int s = 0; s < kFillRpl; s++) {
      const uint64_t j = j0 + 32 * s + lane;
#pragma unroll
      for (int w = 0; w < 10; w++) r[s][w] = j < total ? __ldg(recw + j * 10 + w) : 0u;
    }
  };

  uint32_t cur[kFillRpl][10], nxt[kFillRpl][10];
  uint64_t j = 0;
  load_rec(j, cur);
  uint64_t end_next = level_start[waves > 1 ? 1 : waves] - base;  // wave ends are read one wave ahead (28 600 short waves at 2^20)
  for (uint32_t l = 0; l < waves; l++) {
    const uint64_t wave_end = l == 0 ? level_start[1] - base : end_next;
    end_next = level_start[l + 2 <= waves ? l + 2 : waves] - base;
    while (j < wave_end) {
      const uint32_t chunk = uint32_t(wave_end - j < kFillChunk ? wave_end - j : kFillChunk);
      load_rec(j + chunk, nxt);  // the records follow each other whatever the waves are: always one chunk ahead
      uint32_t own[kFillUnroll], d1[kFillUnroll], d2[kFillUnroll], d3[kFillUnroll], by[kFillUnroll][3], hl[kFillUnroll], hh[kFillUnroll];
#pragma unroll
      for (int u = 0; u < kFillUnroll; u++) {
        constexpr int kSlotsPerSet = 32 / kFillGroups;
        const int set = u / kSlotsPerSet;                                    // which of the lane's records (compile time)
        const uint32_t idx = kFillGroups * u + group;                        // key of the chunk this lane group works on
        const uint32_t src = idx < chunk ? idx - 32 * set : 0;               // past the end of the chunk: lane 0's record of the set --
        const uint32_t(&r)[10] = cur[set];                                   // a later key or all zeros, either way safe addresses; unused
        hl[u] = __shfl_sync(0xffffffffu, r[0], src), hh[u] = __shfl_sync(0xffffffffu, r[1], src);
        const uint32_t vl = __shfl_sync(0xffffffffu, r[2], src), vh = __shfl_sync(0xffffffffu, r[3], src);
        own[u] = __shfl_sync(0xffffffffu, r[4], src);
        const uint32_t o1 = __shfl_sync(0xffffffffu, r[5], src), o2 = __shfl_sync(0xffffffffu, r[6], src);
        const uint32_t o3 = __shfl_sync(0xffffffffu, r[7], src);
        const uint32_t vlen = __shfl_sync(0xffffffffu, r[8], src), key = __shfl_sync(0xffffffffu, r[9], src);
        const uint64_t v0 = (uint64_t(vh) << 32) | vl;
        d1[u] = __ldcg(D + uint64_t(o1) * N + ec);
        d2[u] = __ldcg(D + uint64_t(o2) * N + ec);
        d3[u] = ARITY == 4 ? __ldcg(D + uint64_t(o3) * N + ec) : 0u;
        // bytes byte0 .. byte0+2 of  digest || value || 0x81 || 0...  : one predicated load each, no branches
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const uint32_t t = byte0 + k;
          const bool in_digest = t < 32, in_value = !in_digest && t - 32 < vlen;
          const uint8_t *p = in_digest ? digests + 32ull * key + t : values + v0 + (t - 32);
          uint32_t x = (!in_digest && t - 32 == vlen) ? 0x81u : 0u;
          if (in_digest || in_value) x = __ldg(p);
          by[u][k] = x;
        }
      }
#pragma unroll
      for (int u = 0; u < kFillUnroll; u++) {
        const uint32_t v = by[u][0] | by[u][1] << 8 | by[u][2] << 16;
        const uint64_t hash = (uint64_t(hh[u]) << 32) | hl[u];
        uint32_t x = (v >> sh) & mask;
        x -= d1[u] + d2[u] + d3[u] + static_cast<uint32_t>(fmix64_dev(hash + ec));
        if (act && kFillGroups * u + group < chunk) D[uint64_t(own[u]) * N + e] = x & mask;
      }
      j += chunk;
#pragma unroll
      for (int s = 0; s < kFillRpl; s++)
#pragma unroll
        for (int w = 0; w < 10; w++) cur[s][w] = nxt[s][w];
    }
  }
}

}  // namespace

// All pointers except level_start_host are device pointers; D must be zeroed.  scratch_records: 40 bytes per key, scratch_levels:
// waves + 1 words (both device memory; NULL selects the wave-per-launch schedule).  level_start_host must stay valid until the work
// enqueued on s has been synchronised (it is copied with cudaMemcpyAsync from pageable memory, i.e. during the call).
int launch_device_row_fill(uint32_t arity, const uint32_t *members, const uint32_t *level_start_host, uint32_t waves, const uint64_t *order,
                           const uint8_t *found, const uint32_t *key_of_order, const uint8_t *digests, const uint8_t *values,
                           const uint64_t *val_off, uint32_t *D, uint64_t N, uint32_t b, uint32_t segment_length,
                           uint32_t segment_count_length, void *scratch_records, uint32_t *scratch_levels, cudaStream_t s) {
  FillArgs a{};
  a.order = order;
  a.found = found;
  a.key_of_order = key_of_order;
  a.digests = digests;
  a.values = values;
  a.val_off = val_off;
  a.D = D;
  a.N = N;
  a.b = b;
  a.segment_length = segment_length;
  a.segment_count_length = segment_count_length;
  uint64_t count = 0;
  for (uint32_t l = 0; l < waves; l++) count += level_start_host[l + 1] - level_start_host[l];
  const char *mode = std::getenv("CHPIR_FILL");
  if (count > 0 && scratch_records && scratch_levels && !(mode && std::strcmp(mode, "waves") == 0)) {
    // column ownership: records in wave order, then one launch of N / 4 single-warp CTAs.  The scratch (40 bytes per key + the wave
    // table) comes from the caller, allocated before the host-side peeling: a cudaMalloc here would queue behind whatever large
    // allocation another thread of the setup is making (the 8.4 GB panel ring of the XOF pipeline) with the GPU standing idle.
    FillRec *rec = static_cast<FillRec *>(scratch_records);
    a.members = members + level_start_host[0];
    if (cudaMemcpyAsync(scratch_levels, level_start_host, (size_t(waves) + 1) * 4, cudaMemcpyHostToDevice, s) != cudaSuccess)
      return CHPIR_ERR_CUDA_TRANSFER_FAILED;
    const unsigned pg = unsigned((count + 255) / 256), fg = unsigned((N + kFillCols - 1) / kFillCols);
    if (arity == 3) {
      fill_prep_kernel<3><<<pg, 256, 0, s>>>(a, count, rec);
      fill_columns_kernel<3><<<fg, 32, 0, s>>>(rec, scratch_levels, waves, digests, values, D, N, b);
    } else {
      fill_prep_kernel<4><<<pg, 256, 0, s>>>(a, count, rec);
      fill_columns_kernel<4><<<fg, 32, 0, s>>>(rec, scratch_levels, waves, digests, values, D, N, b);
    }
    return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
  }
  // every field e < N reads bytes [e*b/8, e*b/8 + 2]
  const uint32_t stream_bytes = uint32_t(((N - 1) * b) / 8 + 3 + 3) & ~3u;
  if (stream_bytes > 200 * 1024) return CHPIR_ERR_INVALID_ARGUMENT;
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
