// encode_dev.cu -- the row fill of Matrix::from_kv_database_with_{3,4}_wise_xor_filter (chalametpir_common/src/matrix.rs:707-746,
// :839-885) and encode_kv_as_row (serialization.rs:22-116) on the GPU.
//
// The host keeps what is inherently serial or tiny -- key digests, filter construction (peeling), the wave plan
// (host_encode.cpp: digest_and_peel, plan_fill_levels) -- and uploads the raw values (1.07 GB at 2^20 x 1 kB) instead of the
// encoded matrix D (4.4 GB).  One CTA builds one row of D:
//     own[e] = ( field_e(digest || value || 0x81) - D[o1][e] - D[o2][e] (- D[o3][e]) - mix(hash, e) ) & (2^b - 1)
// where field_e is the e-th b-bit field, LSB first, of the byte string (the reference's bit packing), o1..o3 are the key's other
// filter slots and mix is the reference's murmur finaliser (binary_fuse_filter.rs:553-560).  The rows a key reads belong to keys
// peeled later, so the keys are processed in dependency waves (one launch per wave; a wave's members are mutually independent).
// D is written row-major u32, exactly the matrix the reference would hold, so everything downstream (pack, limb split, GEMM)
// is unchanged and the bytes can be compared with the host encoder one to one.
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

}  // namespace

// All pointers are device pointers; D must be zeroed.  One launch per wave, in wave order, on stream s.
int launch_device_row_fill(uint32_t arity, const uint32_t *members, const uint32_t *level_start_host, uint32_t waves, const uint64_t *order,
                           const uint8_t *found, const uint32_t *key_of_order, const uint8_t *digests, const uint8_t *values,
                           const uint64_t *val_off, uint32_t *D, uint64_t N, uint32_t b, uint32_t segment_length,
                           uint32_t segment_count_length, cudaStream_t s) {
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
