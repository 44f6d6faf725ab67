// cluster.cu -- the sharded server on 1..8 GPUs of ONE process, behind the reference's two calls.
//
// The reference's public surface is `Server::setup(seed, db)` (chalametpir_server/src/server.rs:103) and `Server::respond(&self, query)`
// (server.rs:184); neither can grow a rank argument, so the sharding of north_star (3) lives behind the handle: this translation
// unit owns one chpir_ctx per GPU, the peer mappings between them, the per-GPU streams and the NCCL communicators, and exposes
// chpir_cluster_server_{setup*,respond*} (include/chalamet_b200.h).
//
// Two cuts of D = K x N are used, each where it moves the fewest bytes:
//   * setup (hint M = A.D): COLUMN slices.  M[:, n0:n1] = A.D[:, n0:n1] needs no cross-rank arithmetic; the
//     slices of the hint are gathered on rank 0 over NCCL (ncclSend/ncclRecv, communicators from ncclCommInitAll) and downloaded as
//     one wire-format matrix.
//   * respond (resp = q.D): ROW blocks (the K-sharded alternative of SURVEY.md section 8e, default for n > 1).  Rank r keeps rows
//     [k0_r, k0_r + kn_r) of D at full width and needs only words [k0_r, k0_r + kn_r) of a query -- exactly the words it ingests over
//     ITS OWN PCIe link.  No query word ever crosses NVLink; what crosses is each rank's N-word partial sum, added on rank 0 by one
//     small kernel that reads the peers' partials (exact: addition mod 2^32 is order independent, matrix.rs:328-485).  Full-width rows
//     also keep the packed row pitch of the one-GPU layout (1088 B at N = 940) instead of 118-column slivers padded to 144 B.
//     After the hint gather every rank re-cuts its share of D (peer copies / one more upload) and frees the column slice.
//   * CHPIR_CLUSTER_SHARD=cols keeps the column cut for respond as well (the round-1 design, kept as the measured comparison): every
//     rank then needs the WHOLE query, gathered from the ranks that ingested its slices by NVLink peer reads fused with the limb
//     split (gather_q_kernel<true>) or by copy-engine peer copies (GEMV route), and writes its columns straight into the result row.
//   * ingest: a query buffer in page-locked memory (chpir_host_alloc, or anything cudaHostRegister'ed) is read by the GPUs
//     themselves -- one pull kernel per GPU per coalesced batch fetches that GPU's words of up to 128 queries over its PCIe link --
//     instead of one DMA call per query and GPU; pageable buffers take the per-query DMA route.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <memory>
#include <string>
#include <thread>

#include "common.cuh"
#include "host_encode.hpp"
#include "device_fill.cuh"
#include "host_pipe.cuh"
#include "server_state.cuh"

namespace chpir {
namespace {

constexpr uint32_t kMaxRanks = 16;
constexpr uint32_t kMaxBatch = Coalescer::kMaxBatch;       // queries per coalesced batch = one M tile of the limb GEMM
constexpr uint32_t kTcFrom = Coalescer::kTensorCoreFrom;   // below this many queries the streaming GEMV is cheaper
constexpr uint32_t kGemvChunk = 32;                        // device-resident GEMV path, column cut: queries per all-gather / launch
constexpr uint32_t kIngestDepth = 2;                       // batches whose queries may be crossing PCIe at once: the second one's copies
                                                           // are queued while the first one's run, so the links never wait for a host thread
constexpr uint32_t kSecondMin = 16;                        // ... but a second batch joins one that is still crossing only if it has at least this
                                                           // many members: a tensor-core pass costs the same for 8 queries as for 128, so
                                                           // with few callers fewer, larger batches keep the exec stage off the critical path
constexpr uint32_t kSlots = 2 + kIngestDepth;              // coalescing slots: one collecting callers, kIngestDepth on PCIe, one on the SMs

// ---- NCCL, resolved at run time (the library has no link-time dependency on it: single-GPU users never load it) -----------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl_api() {
  static std::mutex mu;
  static NcclApi api;
  static bool tried = false;
  std::lock_guard<std::mutex> g(mu);
  if (!tried) {
    tried = true;
    // a copy the process already holds (e.g. the one a host framework loaded) wins over the system one: two NCCLs in one process
    // would each build their own topology and proxy threads
    // (the loader shares one instance per SONAME, so the FIRST libnccl.so.2 a process loads is the one everybody gets: a host
    // framework that ships its own NCCL must be given the chance to name it -- CHPIR_NCCL_LIB, set by chalametpir_b200/cluster.py)
    void *h = nullptr;
    if (const char *path = std::getenv("CHPIR_NCCL_LIB"); path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (h) {
      api.handle = h;
      api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(h, "ncclGetVersion"));
      api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
      api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(h, "ncclGroupStart"));
      api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
      api.Send = reinterpret_cast<decltype(api.Send)>(dlsym(h, "ncclSend"));
      api.Recv = reinterpret_cast<decltype(api.Recv)>(dlsym(h, "ncclRecv"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
      if (!api.GetVersion || !api.CommInitAll || !api.CommDestroy || !api.GroupStart || !api.GroupEnd || !api.Send || !api.Recv) api.handle = nullptr;
    }
  }
  return api.handle ? &api : nullptr;
}

// The cluster calls switch the calling thread's current device; a host framework that tracks "its" device must find it unchanged.
struct DeviceRestore {
  int dev = -1;
  DeviceRestore() {
    if (cudaGetDevice(&dev) != cudaSuccess) {
      dev = -1;
      (void)cudaGetLastError();
    }
  }
  ~DeviceRestore() {
    if (dev >= 0) cudaSetDevice(dev);
  }
};

uint32_t env_u32(const char *name, uint32_t dflt) {
  const char *v = std::getenv(name);
  return v && *v ? uint32_t(std::strtoul(v, nullptr, 10)) : dflt;
}
bool env_is(const char *name, const char *want) {
  const char *v = std::getenv(name);
  return v && std::strcmp(v, want) == 0;
}

// ---- the slice plan ---------------------------------------------------------------------------------------------------------------
struct Plan {
  uint32_t c0, nc;
  uint64_t k0, kn, ks;
};
Plan plan_of(uint32_t n, uint32_t r, uint64_t K, uint32_t N) {
  Plan p{};
  const uint32_t base = N / n, rem = N % n;
  p.c0 = r * base + std::min(r, rem);
  p.nc = base + (r < rem ? 1u : 0u);
  // query slices: equal padded length, a multiple of 32 words so that every slice starts 128-byte aligned in the whole query
  p.ks = ((K + n - 1) / n + 31) / 32 * 32;
  p.k0 = std::min<uint64_t>(K, uint64_t(r) * p.ks);
  p.kn = std::min<uint64_t>(K, p.k0 + p.ks) - p.k0;
  return p;
}

// ---- fused all-gather + limb split (tensor-core path) / all-gather (GEMV path) over peer memory -------------------------------------
struct SrcTable {
  const uint32_t *p[kMaxRanks];  // p[s] = rank s's slice buffer (rows x ks u32, row pitch ks), a peer-mapped address for s != self
};

// Word k of query row r lives on rank s = k / ks at p[s][r * ks + (k - s * ks)].  One thread handles 4 consecutive words (one 16-byte
// peer load: a group never straddles a slice because ks % 32 == 0); kUnr groups are in flight per thread, which with every thread
// of the GPU resident covers the ~2 us NVLink round trip at full link rate.
//   SPLIT: writes the K-major byte planes the limb GEMM reads (planes[l][row][k], row pitch kp, plane pitch plane_rows * kp) -- the
//          split_a_limbs of gemm_tc.cu with the loads going to whichever GPU ingested those words;
//   else:  writes whole u32 query rows (row pitch K) for the streaming GEMV.
template <bool SPLIT>
__global__ void __launch_bounds__(256) gather_q_kernel(SrcTable src, uint32_t rows, uint64_t ks, uint64_t K, uint64_t kp, uint32_t plane_rows,
                                                       uint8_t *__restrict__ planes, uint32_t *__restrict__ out_rows) {
  constexpr int kUnr = 4;
  const uint64_t groups = kp / 4;  // kp = K rounded up to 16: the tail groups are zero-filled
  const uint64_t total = uint64_t(rows) * groups;
  const uint64_t plane = uint64_t(plane_rows) * kp;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t base = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; base < total; base += stride * kUnr) {
    uint4 v[kUnr];
    uint64_t rr[kUnr], kk[kUnr];
#pragma unroll
    for (int u = 0; u < kUnr; u++) {
      const uint64_t idx = base + u * stride;
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      rr[u] = 0, kk[u] = 0;
      if (idx < total) {
        const uint64_t r = idx / groups, k = 4 * (idx - r * groups);
        rr[u] = r, kk[u] = k;
        if (k < K) {
          const uint64_t s = k / ks;
          const uint32_t *p = src.p[s] + r * ks + (k - s * ks);
          if (k + 3 < K) {
            v[u] = *reinterpret_cast<const uint4 *>(p);
          } else {  // the ragged end of the query: K % 4 words
            v[u].x = p[0];
            if (k + 1 < K) v[u].y = p[1];
            if (k + 2 < K) v[u].z = p[2];
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnr; u++) {
      const uint64_t idx = base + u * stride;
      if (idx >= total) continue;
      const uint64_t r = rr[u], k = kk[u];
      if (SPLIT) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int l = 0; l < 4; l++) {
          const uint32_t b = ((w[0] >> (8 * l)) & 0xffu) | (((w[1] >> (8 * l)) & 0xffu) << 8) | (((w[2] >> (8 * l)) & 0xffu) << 16) |
                             (((w[3] >> (8 * l)) & 0xffu) << 24);
          *reinterpret_cast<uint32_t *>(planes + l * plane + r * kp + k) = b;
        }
      } else {
        uint32_t *o = out_rows + r * K + k;
        if (k < K) o[0] = v[u].x;
        if (k + 1 < K) o[1] = v[u].y;
        if (k + 2 < K) o[2] = v[u].z;
        if (k + 3 < K) o[3] = v[u].w;
      }
    }
  }
}

int launch_gather(bool split, const SrcTable &src, uint32_t rows, uint64_t ks, uint64_t K, uint64_t kp, uint8_t *planes, uint32_t *out_rows,
                  int sm_count, cudaStream_t st) {
  if (rows == 0) return CHPIR_OK;
  const uint64_t total = uint64_t(rows) * (kp / 4);
  const uint64_t want = (total + 256ull * 4 - 1) / (256ull * 4);
  const unsigned grid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>(want, uint64_t(sm_count) * 8)));
  if (split)
    gather_q_kernel<true><<<grid, 256, 0, st>>>(src, rows, ks, K, kp, 128, planes, nullptr);
  else
    gather_q_kernel<false><<<grid, 256, 0, st>>>(src, rows, ks, K, kp, 0, nullptr, out_rows);
  return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}


// ---- ingest: the GPU fetches its words of a batch of queries from page-locked host memory ------------------------------------------
struct PullTable {
  const uint8_t *src[kMaxBatch];  // query i's wire bytes (8-byte header + K u32), device-readable host memory; nullptr = not this way
};

template <class V>
__device__ __forceinline__ void pull_row(const uint8_t *s, uint32_t *d, uint64_t words, uint64_t tid, uint64_t stride) {
  constexpr uint64_t W = sizeof(V) / 4;
  constexpr int kUnr = 4;
  const V *sv = reinterpret_cast<const V *>(s);
  V *dv = reinterpret_cast<V *>(d);
  const uint64_t nv = words / W;
  for (uint64_t i = tid; i < nv; i += stride * kUnr) {
    V v[kUnr];
#pragma unroll
    for (int u = 0; u < kUnr; u++)
      if (i + u * stride < nv) v[u] = __ldcv(sv + i + u * stride);  // host memory: never served from a cache line of an earlier batch
#pragma unroll
    for (int u = 0; u < kUnr; u++)
      if (i + u * stride < nv) dv[i + u * stride] = v[u];
  }
  for (uint64_t i = nv * W + tid; i < words; i += stride) d[i] = __ldcv(reinterpret_cast<const uint32_t *>(s) + i);
}

// A fixed, small grid walks the (row, segment) pairs of the batch: words [0, words) of row r = bytes [byte_off, byte_off + 4 * words)
// of t.src[r]  ->  dst + r * ks.  The kernel is bound by the PCIe round trip, not by threads: one block per SM with four 16-byte loads
// in flight per thread keeps ~2.4 MB outstanding, far more than the link needs, and leaves the SMs' thread slots to the previous
// batch's kernels that run beside it.  byte_off = 8 + 4 * k0 is a multiple of 8, so the widest load the source allows depends only on
// the caller's buffer alignment; 4-byte alignment of the buffer is the caller-side condition for taking this route at all.
constexpr uint32_t kPullSegWords = 256 * 4 * 4;  // one block-wide sweep of 16-byte loads, 4 per thread
__global__ void __launch_bounds__(256) pull_slices_kernel(PullTable t, uint32_t rows, uint64_t byte_off, uint64_t words, uint64_t ks,
                                                          uint32_t *__restrict__ dst) {
  const uint64_t segs = (words + kPullSegWords - 1) / kPullSegWords;
  for (uint64_t item = blockIdx.x; item < uint64_t(rows) * segs; item += gridDim.x) {
    const uint32_t row = uint32_t(item / segs);
    const uint64_t w0 = (item - uint64_t(row) * segs) * kPullSegWords;
    const uint8_t *s = t.src[row];
    if (!s) continue;
    s += byte_off + 4 * w0;
    uint32_t *d = dst + uint64_t(row) * ks + w0;
    const uint64_t n = words - w0 < kPullSegWords ? words - w0 : kPullSegWords;
    const uint32_t a = uint32_t(reinterpret_cast<uintptr_t>(s) & 15u);
    if (a == 0)
      pull_row<uint4>(s, d, n, threadIdx.x, blockDim.x);
    else if ((a & 7u) == 0)
      pull_row<uint2>(s, d, n, threadIdx.x, blockDim.x);
    else
      pull_row<uint32_t>(s, d, n, threadIdx.x, blockDim.x);
  }
}

int launch_pull(const PullTable &t, uint32_t rows, uint64_t k0, uint64_t kn, uint64_t ks, uint32_t *dst, int sm_count, cudaStream_t st) {
  if (rows == 0 || kn == 0) return CHPIR_OK;
  const uint64_t items = uint64_t(rows) * ((kn + kPullSegWords - 1) / kPullSegWords);
  const unsigned grid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>(items, uint64_t(sm_count) * env_u32("CHPIR_PULL_BLOCKS_PER_SM", 1))));
  pull_slices_kernel<<<grid, 256, 0, st>>>(t, rows, 8 + 4 * k0, kn, ks, dst);
  return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}

// ---- row cut: rank 0 adds the ranks' partial responses -------------------------------------------------------------------------------
struct PartTable {
  const uint32_t *p[kMaxRanks];  // p[d] = rank d's partial sums (words u32), a peer-mapped address for d != 0
};

template <class V>
__global__ void __launch_bounds__(256) reduce_parts_kernel(PartTable t, uint32_t n, uint64_t count, V *__restrict__ out) {
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < count; i += uint64_t(gridDim.x) * blockDim.x) {
    V acc = reinterpret_cast<const V *>(t.p[0])[i];
    for (uint32_t d = 1; d < n; d++) {
      const V v = reinterpret_cast<const V *>(t.p[d])[i];
      if constexpr (sizeof(V) == 16) {
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      } else {
        acc += v;
      }
    }
    out[i] = acc;
  }
}

int launch_reduce(const PartTable &t, uint32_t n, uint64_t words, uint32_t *out, int sm_count, cudaStream_t st) {
  if (words == 0) return CHPIR_OK;
  bool vec = words % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  for (uint32_t d = 0; d < n; d++) vec = vec && (reinterpret_cast<uintptr_t>(t.p[d]) & 15u) == 0;
  const uint64_t count = vec ? words / 4 : words;
  const unsigned grid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>((count + 255) / 256, uint64_t(sm_count) * 4)));
  if (vec)
    reduce_parts_kernel<uint4><<<grid, 256, 0, st>>>(t, n, count, reinterpret_cast<uint4 *>(out));
  else
    reduce_parts_kernel<uint32_t><<<grid, 256, 0, st>>>(t, n, count, out);
  return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}

// Is [p, p + bytes) page-locked host memory a kernel may read?  The library's own allocations are known; anything else is asked of
// the driver (cudaHostRegister'ed or another allocator's pinned memory).
bool device_readable_host(const void *p, size_t bytes) {
  if (pinned_registry_contains(p, bytes)) return true;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost && at.devicePointer == p;
}

}  // namespace
}  // namespace chpir

using namespace chpir;

// ------------------------------------------------------------------------------------------------------------------- handles
struct chpir_cluster {
  int n = 0;
  std::vector<int> dev;
  std::vector<chpir_ctx *> ctx;
  std::mutex mu;  // NCCL communicator creation / use (setup-time only)
  std::vector<ncclComm_t> comms;
  int nccl_version = 0;

  int ensure_comms() {
    if (!comms.empty()) return CHPIR_OK;
    NcclApi *api = nccl_api();
    if (!api) return CHPIR_ERR_NCCL_FAILED;
    api->GetVersion(&nccl_version);
    comms.assign(size_t(n), nullptr);
    const ncclResult_t r = api->CommInitAll(comms.data(), n, dev.data());
    if (r != ncclSuccess) {
      comms.clear();
      return CHPIR_ERR_NCCL_FAILED;
    }
    return CHPIR_OK;
  }
  // The first send/recv between two ranks sets up their channels (~0.2 s at n = 8); the hint gather is the only NCCL traffic of a
  // setup and sits on its critical path after the XOF chain, so the same rank d -> rank 0 pattern is run once with a few bytes
  // right after the communicators exist (on the helper thread, beside the chain).
  bool comms_warm = false;
  int warm_comms() {
    if (comms_warm || n < 2 || comms.empty()) return CHPIR_OK;
    NcclApi *api = nccl_api();
    if (!api) return CHPIR_ERR_NCCL_FAILED;
    std::vector<void *> buf(size_t(n), nullptr);
    std::vector<cudaStream_t> st(size_t(n), nullptr);
    bool ok = true;
    for (int d = 0; d < n && ok; d++)
      ok = cudaSetDevice(dev[size_t(d)]) == cudaSuccess && cudaMalloc(&buf[size_t(d)], 256 * size_t(n)) == cudaSuccess &&
           cudaStreamCreateWithFlags(&st[size_t(d)], cudaStreamNonBlocking) == cudaSuccess;
    if (ok) {
      ncclResult_t nr = api->GroupStart();
      for (int d = 1; d < n && nr == ncclSuccess; d++) {
        nr = api->Send(buf[size_t(d)], 64, ncclUint32, 0, comms[size_t(d)], st[size_t(d)]);
        if (nr == ncclSuccess) nr = api->Recv(static_cast<char *>(buf[0]) + 256 * d, 64, ncclUint32, d, comms[0], st[0]);
      }
      ok = api->GroupEnd() == ncclSuccess && nr == ncclSuccess;
    }
    for (int d = 0; d < n; d++) {
      cudaSetDevice(dev[size_t(d)]);
      if (st[size_t(d)]) {
        if (cudaStreamSynchronize(st[size_t(d)]) != cudaSuccess) ok = false;
        cudaStreamDestroy(st[size_t(d)]);
      }
      if (buf[size_t(d)]) cudaFree(buf[size_t(d)]);
    }
    (void)cudaGetLastError();
    comms_warm = ok;
    return ok ? CHPIR_OK : CHPIR_ERR_NCCL_FAILED;
  }
  ~chpir_cluster() {
    if (!comms.empty()) {
      NcclApi *api = nccl_api();
      if (api)
        for (ncclComm_t c : comms)
          if (c) api->CommDestroy(c);
    }
    for (chpir_ctx *c : ctx)
      if (c) chpir_ctx_destroy(c);
  }
};

namespace {

struct Rank {
  int dev = 0;
  chpir_ctx *ctx = nullptr;
  chpir_server *srv = nullptr;  // the shard respond runs on: the row block (row cut) or the column slice (column cut, n = 1)
  chpir_server *col = nullptr;  // row cut, during setup only: the column slice the hint was computed on (freed after the gather)
  Plan pl{};
  cudaStream_t compute = nullptr, gather = nullptr;
  // device-resident path (chpir_cluster_server_respond_device): double-buffered scratch, allocated on first use
  uint32_t *q_full[2] = {nullptr, nullptr};
  uint32_t *resp[2] = {nullptr, nullptr};  // column cut: this rank's columns of a chunk; row cut: its partial sums of a whole pass
  uint32_t q_rows = 0, resp_rows = 0;
  cudaEvent_t gathered[2] = {nullptr, nullptr}, computed[2] = {nullptr, nullptr}, reduced[2] = {nullptr, nullptr};
  uint32_t tc_buf = 0;  // next buffer of this rank's GEMM operand ring
};

// One coalescing slot: the members' slices on every rank, every rank's share of the result, and the pinned rows they land in.
struct CBatch {
  uint32_t *q_slice[kMaxRanks] = {};  // [kMaxBatch][ks] on rank d
  uint32_t *q_full[kMaxRanks] = {};   // column cut: [gemv_rows][K] on rank d, whole queries for the GEMV route
  uint32_t *resp[kMaxRanks] = {};     // column cut: [kMaxBatch][nc_d], rank d's columns; row cut: [kMaxBatch][N], rank d's partial sums
  uint32_t *resp0 = nullptr;          // row cut: [kMaxBatch][N] on rank 0, the sum of the partials
  cudaStream_t copy[kMaxRanks] = {};
  cudaEvent_t uploaded[kMaxRanks] = {}, done[kMaxRanks] = {};
  uint32_t *h_resp = nullptr;  // pinned [kMaxBatch][N]
  PullTable pull{};            // members whose query the GPUs fetch themselves (page-locked source)
  uint32_t n_pull = 0;
  uint32_t count = 0, issued = 0, picked = 0;
  bool closed = false, finished = false;
  int rc = CHPIR_OK;
};

}  // namespace

struct chpir_cluster_server {
  chpir_cluster *cl = nullptr;
  uint32_t n = 0;
  uint64_t K = 0, ks = 0;
  uint32_t N = 0, b = 0, lwe = 0;
  std::vector<Rank> r;
  bool rows = false;        // respond runs on row blocks of D (see the head of this file); false: column slices
  bool pipeline = false;    // chpir_cluster_server_respond goes through the coalescing pipeline below: always for n > 1, for one GPU
                            // when chpir_setup_opts.respond_coalesce asks for it (otherwise a lone GPU serves every call on its own slot)
  bool tc = false;          // every rank keeps its limb planes: batches of >= kTcFrom queries take the tensor-core route
  uint32_t gemv_rows = 0;   // queries one GEMV-route batch may hold (column cut: capacity of CBatch::q_full)
  // coalescer.  A batch passes two stages, each held by one batch at a time:
  // ingest (its queries cross PCIe) and exec (kernels + result download); a third slot collects callers meanwhile.
  std::mutex mu;
  // one condition per reason to wait, so that a wake-up reaches only threads it concerns (hundreds of callers share this object):
  // callers waiting for a slot that takes members, a leader waiting for its members' uploads / for the next slot to be vacated,
  // members waiting for their batch's responses
  std::condition_variable cv_open, cv_free, cv_issued[kSlots], cv_done[kSlots], cv_ingest;
  std::mutex exec_mu;
  // the ingest stage admits up to kIngestDepth batches, in arrival order (guarded by mu, waited for on cv_ingest)
  uint32_t ingest_busy = 0;
  uint64_t ingest_next = 0, ingest_serving = 0;
  CBatch cb[kSlots];
  int open = 0;
  bool co_ready = false;
  uint64_t batches = 0, queries = 0, tc_batches = 0, pulled = 0;
  double ingest_wait_s = 0, ingest_s = 0, exec_wait_s = 0, exec_s = 0;  // leaders' wall time per pipeline stage, summed over batches
  // chpir_cluster_server_respond_batch
  std::mutex batch_mu;
  std::unique_ptr<CBatch> bb;
  // device-resident path
  std::mutex dev_mu;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  double setup_total_s = 0, hint_gather_s = 0, reshard_s = 0;
  uint32_t gather_uses_nccl = 0;

  void free_batch(CBatch &B) {
    for (uint32_t d = 0; d < n; d++) {
      cudaSetDevice(r[d].dev);
      if (B.copy[d]) {
        cudaStreamSynchronize(B.copy[d]);
        cudaStreamDestroy(B.copy[d]);
      }
      if (B.uploaded[d]) cudaEventDestroy(B.uploaded[d]);
      if (B.done[d]) cudaEventDestroy(B.done[d]);
      if (B.q_slice[d]) cudaFree(B.q_slice[d]);
      if (B.q_full[d]) cudaFree(B.q_full[d]);
      if (B.resp[d]) cudaFree(B.resp[d]);
    }
    if (B.resp0) {
      cudaSetDevice(r[0].dev);
      cudaFree(B.resp0);
    }
    if (B.h_resp) cudaFreeHost(B.h_resp);
    B = CBatch{};
  }

  int init_batch(CBatch &B) {
    for (uint32_t d = 0; d < n; d++) {
      if (cudaSetDevice(r[d].dev) != cudaSuccess) return CHPIR_ERR_CUDA_DEVICE_NOT_FOUND;
      const size_t resp_words = size_t(kMaxBatch) * (rows ? N : r[d].pl.nc);
      if (cudaMalloc(&B.q_slice[d], size_t(kMaxBatch) * ks * 4) != cudaSuccess ||
          (!rows && cudaMalloc(&B.q_full[d], size_t(gemv_rows) * K * 4) != cudaSuccess) ||
          cudaMalloc(&B.resp[d], resp_words * 4) != cudaSuccess ||
          cudaStreamCreateWithFlags(&B.copy[d], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&B.uploaded[d], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&B.done[d], cudaEventDisableTiming) != cudaSuccess) {
        set_last_cuda_error(cudaGetLastError(), "cluster respond batch allocation");
        return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
      }
      // the words behind a short last slice (and whole rows of ranks that ingest nothing) face zero rows of D, but must be defined
      if (cudaMemset(B.q_slice[d], 0, size_t(kMaxBatch) * ks * 4) != cudaSuccess) return CHPIR_ERR_CUDA_TRANSFER_FAILED;
    }
    if (rows) {
      if (cudaSetDevice(r[0].dev) != cudaSuccess) return CHPIR_ERR_CUDA_DEVICE_NOT_FOUND;
      if (cudaMalloc(&B.resp0, size_t(kMaxBatch) * N * 4) != cudaSuccess) {
        set_last_cuda_error(cudaGetLastError(), "cluster respond batch allocation");
        return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
      }
    }
    if (cudaHostAlloc(&B.h_resp, size_t(kMaxBatch) * N * 4, cudaHostAllocPortable) != cudaSuccess) {
      set_last_cuda_error(cudaGetLastError(), "cluster respond batch pinned rows");
      return CHPIR_ERR_HOST_ALLOCATION_FAILED;
    }
    return CHPIR_OK;
  }

  ~chpir_cluster_server() {
    for (uint32_t d = 0; d < r.size(); d++) {
      cudaSetDevice(r[d].dev);
      if (r[d].compute) cudaStreamSynchronize(r[d].compute);
      if (r[d].gather) cudaStreamSynchronize(r[d].gather);
    }
    if (n > 0 && r.size() == n) {
      for (CBatch &B : cb) free_batch(B);
      if (bb) free_batch(*bb);
    }
    for (uint32_t d = 0; d < r.size(); d++) {
      cudaSetDevice(r[d].dev);
      for (int p = 0; p < 2; p++) {
        if (r[d].q_full[p]) cudaFree(r[d].q_full[p]);
        if (r[d].resp[p]) cudaFree(r[d].resp[p]);
        if (r[d].gathered[p]) cudaEventDestroy(r[d].gathered[p]);
        if (r[d].computed[p]) cudaEventDestroy(r[d].computed[p]);
        if (r[d].reduced[p]) cudaEventDestroy(r[d].reduced[p]);
      }
      if (r[d].compute) cudaStreamDestroy(r[d].compute);
      if (r[d].gather) cudaStreamDestroy(r[d].gather);
      if (r[d].srv) chpir_server_destroy(r[d].srv);
      if (r[d].col) chpir_server_destroy(r[d].col);
    }
    if (!r.empty()) cudaSetDevice(r[0].dev);
    if (t0) cudaEventDestroy(t0);
    if (t1) cudaEventDestroy(t1);
  }
};

namespace {

#define CHPIR_GUARD_BEGIN try {
#define CHPIR_GUARD_END                      \
  }                                          \
  catch (const std::bad_alloc &) {           \
    return CHPIR_ERR_HOST_ALLOCATION_FAILED; \
  }                                          \
  catch (...) {                              \
    return CHPIR_ERR_INVALID_ARGUMENT;       \
  }

// Streams and events of every rank: needed by the hint gather, so they come first.
int make_streams(chpir_cluster_server *S) {
  for (uint32_t d = 0; d < S->n; d++) {
    Rank &R = S->r[d];
    CHPIR_CUDA(cudaSetDevice(R.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    // the kernels of the batch on the SMs outrank the next batch's ingest kernel, which shares the GPU with them
    int prio_lo = 0, prio_hi = 0;
    CHPIR_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    CHPIR_CUDA(cudaStreamCreateWithPriority(&R.compute, cudaStreamNonBlocking, prio_hi), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
    CHPIR_CUDA(cudaStreamCreateWithPriority(&R.gather, cudaStreamNonBlocking, prio_hi), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
    for (int p = 0; p < 2; p++) {
      CHPIR_CUDA(cudaEventCreateWithFlags(&R.gathered[p], cudaEventDisableTiming), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
      CHPIR_CUDA(cudaEventCreateWithFlags(&R.computed[p], cudaEventDisableTiming), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
      CHPIR_CUDA(cudaEventCreateWithFlags(&R.reduced[p], cudaEventDisableTiming), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
    }
  }
  CHPIR_CUDA(cudaSetDevice(S->r[0].dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  CHPIR_CUDA(cudaEventCreate(&S->t0), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
  CHPIR_CUDA(cudaEventCreate(&S->t1), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
  return CHPIR_OK;
}

// The coalescing slots, once the shards respond runs on exist.
int make_slots(chpir_cluster_server *S) {
  S->tc = true;
  for (uint32_t d = 0; d < S->n; d++) S->tc = S->tc && S->r[d].srv->gemm != nullptr;
  S->gemv_rows = S->rows ? kMaxBatch : (S->tc ? kTcFrom - 1 : kMaxBatch);
  if (S->pipeline) {
    for (CBatch &B : S->cb)
      if (int rc = S->init_batch(B); rc != CHPIR_OK) return rc;
    S->co_ready = true;
  }
  return CHPIR_OK;
}

// Column cut: every rank pulls the slices it did not ingest, answers for its columns and writes them into the pinned rows.
int enqueue_batch_cols(chpir_cluster_server *S, CBatch &B, uint32_t nq, bool tc) {
  SrcTable src{};
  for (uint32_t d = 0; d < S->n; d++) src.p[d] = B.q_slice[d];
  int rc = CHPIR_OK;
  for (uint32_t d = 0; d < S->n && rc == CHPIR_OK; d++) {
    Rank &R = S->r[d];
    chpir_server *sv = R.srv;
    CHPIR_CUDA(cudaSetDevice(R.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    cudaStream_t st = R.compute;
    for (uint32_t s = 0; s < S->n; s++) CHPIR_CUDA(cudaStreamWaitEvent(st, B.uploaded[s], 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
    CHPIR_CUDA(cudaMemsetAsync(B.resp[d], 0, size_t(nq) * R.pl.nc * 4, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
    if (tc) {
      std::lock_guard<std::mutex> g(sv->gemm_mu);
      const int buf = int(R.tc_buf++ & 1);
      if ((rc = gemm_tc_buf_acquire(sv->gemm, buf, st)) != CHPIR_OK) break;
      uint8_t *planes = gemm_tc_ring(sv->gemm) + uint64_t(buf) * gemm_tc_panel_bytes(sv->gemm);
      if ((rc = launch_gather(true, src, nq, S->ks, S->K, gemm_tc_kp(sv->gemm), planes, nullptr, R.ctx->sm_count, st)) != CHPIR_OK) break;
      if ((rc = gemm_tc_panel(sv->gemm, buf, nq, B.resp[d], st)) != CHPIR_OK) break;
      if ((rc = gemm_tc_buf_release(sv->gemm, buf, st)) != CHPIR_OK) break;
    } else {
      if ((rc = launch_gather(false, src, nq, S->ks, S->K, (S->K + 15) / 16 * 16, nullptr, B.q_full[d], R.ctx->sm_count, st)) != CHPIR_OK) break;
      if ((rc = launch_respond(sv->d_packed, sv->layout, sv->K, sv->plan, B.q_full[d], B.resp[d], nq, st)) != CHPIR_OK) break;
    }
    // this rank's columns of every row, straight into the rows the callers read: no gather step
    CHPIR_CUDA(cudaMemcpy2DAsync(B.h_resp + R.pl.c0, size_t(S->N) * 4, B.resp[d], size_t(R.pl.nc) * 4, size_t(R.pl.nc) * 4, nq, cudaMemcpyDeviceToHost, st),
               CHPIR_ERR_CUDA_TRANSFER_FAILED);
    CHPIR_CUDA(cudaEventRecord(B.done[d], st), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
  }
  return rc;
}

// Row cut: every rank multiplies the words it ingested itself with its rows of D (no query word crosses NVLink); rank 0 adds the n
// partial sums and downloads the finished rows.
int enqueue_batch_rows(chpir_cluster_server *S, CBatch &B, uint32_t nq, bool tc) {
  int rc = CHPIR_OK;
  PartTable parts{};
  for (uint32_t d = 0; d < S->n && rc == CHPIR_OK; d++) {
    Rank &R = S->r[d];
    chpir_server *sv = R.srv;
    parts.p[d] = B.resp[d];
    CHPIR_CUDA(cudaSetDevice(R.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    cudaStream_t st = R.compute;
    CHPIR_CUDA(cudaStreamWaitEvent(st, B.uploaded[d], 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
    CHPIR_CUDA(cudaMemsetAsync(B.resp[d], 0, size_t(nq) * S->N * 4, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
    if (tc) {
      std::lock_guard<std::mutex> g(sv->gemm_mu);
      const int buf = int(R.tc_buf++ & 1);
      if ((rc = gemm_tc_buf_acquire(sv->gemm, buf, st)) != CHPIR_OK) break;
      if ((rc = gemm_tc_load_panel_u32(sv->gemm, buf, B.q_slice[d], nq, st)) != CHPIR_OK) break;
      if ((rc = gemm_tc_panel(sv->gemm, buf, nq, B.resp[d], st)) != CHPIR_OK) break;
      if ((rc = gemm_tc_buf_release(sv->gemm, buf, st)) != CHPIR_OK) break;
    } else {
      if ((rc = launch_respond(sv->d_packed, sv->layout, sv->K, sv->plan, B.q_slice[d], B.resp[d], nq, st)) != CHPIR_OK) break;
    }
    CHPIR_CUDA(cudaEventRecord(B.done[d], st), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
  }
  if (rc != CHPIR_OK) return rc;
  Rank &R0 = S->r[0];
  CHPIR_CUDA(cudaSetDevice(R0.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  for (uint32_t d = 1; d < S->n; d++) CHPIR_CUDA(cudaStreamWaitEvent(R0.compute, B.done[d], 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
  if ((rc = launch_reduce(parts, S->n, uint64_t(nq) * S->N, B.resp0, R0.ctx->sm_count, R0.compute)) != CHPIR_OK) return rc;
  CHPIR_CUDA(cudaMemcpyAsync(B.h_resp, B.resp0, size_t(nq) * S->N * 4, cudaMemcpyDeviceToHost, R0.compute), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  return CHPIR_OK;
}

// The GPU half of one batch of `nq` queries whose slices are on their way into B.q_slice (stream-ordered behind B.uploaded[d]);
// returns when the pinned rows hold the responses.
int run_batch(chpir_cluster_server *S, CBatch &B, uint32_t nq, bool *used_tc) {
  const bool tc = S->tc && nq >= kTcFrom;
  if (used_tc) *used_tc = tc;
  if (!tc && nq > S->gemv_rows) return CHPIR_ERR_INVALID_ARGUMENT;
  const int rc = S->rows ? enqueue_batch_rows(S, B, nq, tc) : enqueue_batch_cols(S, B, nq, tc);
  // wait for every rank that was given work, also on the error path: the slot must be quiet before it is reused
  int out = rc;
  for (uint32_t d = 0; d < S->n; d++) {
    cudaSetDevice(S->r[d].dev);
    const cudaError_t e = cudaStreamSynchronize(S->r[d].compute);
    if (e != cudaSuccess && out == CHPIR_OK) {
      set_last_cuda_error(e, "cluster respond");
      out = CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
    }
  }
  return out;
}

// A member's (or the batch call's) upload of one query from pageable memory: its K/n words go to each rank over that rank's own
// PCIe link, one DMA per rank.
int upload_query(chpir_cluster_server *S, CBatch &B, uint32_t row, const uint8_t *query) {
  for (uint32_t d = 0; d < S->n; d++) {
    const Plan &pl = S->r[d].pl;
    if (pl.kn == 0) continue;
    if (cudaSetDevice(S->r[d].dev) != cudaSuccess ||
        cudaMemcpyAsync(B.q_slice[d] + size_t(row) * S->ks, query + 8 + pl.k0 * 4, pl.kn * 4, cudaMemcpyHostToDevice, B.copy[d]) != cudaSuccess) {
      set_last_cuda_error(cudaGetLastError(), "cluster respond: query slice upload");
      return CHPIR_ERR_CUDA_TRANSFER_FAILED;
    }
  }
  return CHPIR_OK;
}

// How the queries of a batch cross PCIe (CHPIR_CLUSTER_INGEST):
//   batch (default)  page-locked sources are moved by the copy engines, ONE cudaMemcpyBatchAsync per GPU and batch (CUDA 12.8+);
//   pull             ... by one kernel per GPU and batch that reads the host memory itself (measured: ~36 GB/s per link against
//                    ~50 GB/s for the copy engines, profiles/r2_e2e_probe_n2.txt; kept for drivers without the batch call);
//   dma              one cudaMemcpyAsync per query and GPU, issued by the calling thread -- also the route of pageable sources.
enum IngestRoute { kIngestBatch, kIngestPull, kIngestDma };
std::atomic<bool> g_batch_copy_broken{false};  // cudaMemcpyBatchAsync refused once: use the pull kernel from then on
IngestRoute ingest_route() {
  if (env_is("CHPIR_CLUSTER_INGEST", "dma")) return kIngestDma;
  if (env_is("CHPIR_CLUSTER_INGEST", "pull") || g_batch_copy_broken.load(std::memory_order_relaxed)) return kIngestPull;
  return kIngestBatch;
}

// Can the GPUs / copy engines fetch this query from where it lies?  (page-locked, device-readable, 4-byte aligned)
bool pullable(const uint8_t *query, size_t len) {
  return ingest_route() != kIngestDma && (reinterpret_cast<uintptr_t>(query) & 3u) == 0 && device_readable_host(query, len);
}

// Ingest stage of a closed batch: the members that registered a page-locked source are moved now, one call per GPU; the PCIe stage is
// over when every copy stream (these transfers and the other members' own DMAs) has drained; B.uploaded[d] marks that point for
// the kernels.
int ingest_batch(chpir_cluster_server *S, CBatch &B, uint32_t nq) {
  int rc = CHPIR_OK;
  const IngestRoute route = ingest_route();
  std::vector<void *> dsts, srcs;
  std::vector<size_t> sizes;
  for (uint32_t d = 0; d < S->n; d++) {
    const Plan &pl = S->r[d].pl;
    if (cudaSetDevice(S->r[d].dev) != cudaSuccess) return CHPIR_ERR_CUDA_DEVICE_NOT_FOUND;
    if (B.n_pull && pl.kn && rc == CHPIR_OK) {
      bool moved = false;
      if (route == kIngestBatch) {
        dsts.clear(), srcs.clear(), sizes.clear();
        for (uint32_t i = 0; i < nq; i++)
          if (B.pull.src[i]) {
            dsts.push_back(B.q_slice[d] + size_t(i) * S->ks);
            srcs.push_back(const_cast<uint8_t *>(B.pull.src[i]) + 8 + pl.k0 * 4);
            sizes.push_back(size_t(pl.kn) * 4);
          }
        cudaMemcpyAttributes at{};
        at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;  // the callers' buffers stay put until their call returns
        at.flags = cudaMemcpyFlagPreferOverlapWithCompute;
        size_t first = 0, fail = 0;
        const cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &at, &first, 1, &fail, B.copy[d]);
        if (e == cudaSuccess) {
          moved = true;
        } else {
          (void)cudaGetLastError();
          g_batch_copy_broken.store(true, std::memory_order_relaxed);
        }
      }
      if (!moved) rc = launch_pull(B.pull, nq, pl.k0, pl.kn, S->ks, B.q_slice[d], S->r[d].ctx->sm_count, B.copy[d]);
    }
    if (cudaEventRecord(B.uploaded[d], B.copy[d]) != cudaSuccess && rc == CHPIR_OK) rc = CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
  }
  for (uint32_t d = 0; d < S->n; d++) {
    cudaSetDevice(S->r[d].dev);
    const cudaError_t e = cudaStreamSynchronize(B.copy[d]);
    if (e != cudaSuccess && rc == CHPIR_OK) {
      set_last_cuda_error(e, "cluster respond: query ingest");
      rc = CHPIR_ERR_CUDA_TRANSFER_FAILED;
    }
  }
  return rc;
}

// One caller's share of a coalesced batch: the protocol of api.cu's respond_coalesced with n GPUs behind it and two pipeline
// stages.  The first caller of a batch is its leader.  It waits for a place in the ingest stage -- that wait IS the batching window:
// an idle server means a batch of one and no added latency, a busy one means everybody who arrives while the previous batches are
// crossing PCIe shares this one -- closes the batch, moves its queries, waits for the exec stage and runs the kernels while the next
// batches are already being ingested.
int respond_coalesced(chpir_cluster_server *S, const uint8_t *query, size_t query_len, uint8_t *resp_out) {
  const bool pull = pullable(query, query_len);
  CBatch *B = nullptr;
  uint32_t row = 0;
  int si = 0;
  {
    std::unique_lock<std::mutex> lk(S->mu);
    S->cv_open.wait(lk, [&] { return !S->cb[S->open].closed && S->cb[S->open].count < kMaxBatch; });
    si = S->open;
    B = &S->cb[si];
    row = B->count++;
    B->pull.src[row] = pull ? query : nullptr;
    if (pull) B->n_pull++, B->issued++;  // nothing to upload from this thread: the leader moves it
    if (B->count == kSecondMin) S->cv_ingest.notify_all();  // the batch has become worth a second place on PCIe
  }
  const bool leader = row == 0;
  if (!pull) {
    const int up = upload_query(S, *B, row, query);
    bool last = false;
    {
      std::lock_guard<std::mutex> lk(S->mu);
      B->issued++;
      if (up != CHPIR_OK && B->rc == CHPIR_OK) B->rc = up;
      last = B->closed && B->issued == B->count;
    }
    if (last) S->cv_issued[si].notify_one();
  }
  if (leader) {
    const double w0 = now_s();
    double w1 = w0;
    uint32_t nq = 0;
    int rc = CHPIR_OK;
    {
      std::unique_lock<std::mutex> lk(S->mu);
      const uint64_t ticket = S->ingest_next++;
      S->cv_ingest.wait(lk, [&] {
        return S->ingest_serving == ticket && (S->ingest_busy == 0 || (S->ingest_busy < kIngestDepth && B->count >= kSecondMin));
      });
      S->ingest_serving++, S->ingest_busy++;
      w1 = now_s();
      B->closed = true;
      nq = B->count;
      S->cv_issued[si].wait(lk, [&] { return B->issued == nq; });
      rc = B->rc;
      const int next = (S->open + 1) % int(kSlots);
      S->cv_free.wait(lk, [&] { return S->cb[next].count == 0 && !S->cb[next].closed; });
      S->open = next;
    }
    S->cv_ingest.notify_all();  // the next leader in line may now be at the head
    S->cv_open.notify_all();
    const int in = ingest_batch(S, *B, nq);  // also on the error path: the members' DMAs must have drained
    if (rc == CHPIR_OK) rc = in;
    const double w2 = now_s();
    double w3 = w2;
    bool tc = false;
    {
      std::lock_guard<std::mutex> ex(S->exec_mu);
      {
        std::lock_guard<std::mutex> lk(S->mu);
        S->ingest_busy--;
      }
      S->cv_ingest.notify_all();
      w3 = now_s();
      if (rc == CHPIR_OK) rc = run_batch(S, *B, nq, &tc);
    }
    const double w4 = now_s();
    {
      std::lock_guard<std::mutex> lk(S->mu);
      S->batches++, S->queries += nq, S->tc_batches += tc ? 1 : 0, S->pulled += B->n_pull;
      S->ingest_wait_s += w1 - w0, S->ingest_s += w2 - w1, S->exec_wait_s += w3 - w2, S->exec_s += w4 - w3;
      B->rc = rc;
      B->finished = true;
    }
    S->cv_done[si].notify_all();
  }
  int rc = CHPIR_OK;
  {
    std::unique_lock<std::mutex> lk(S->mu);
    S->cv_done[si].wait(lk, [&] { return B->finished; });
    rc = B->rc;
  }
  if (rc == CHPIR_OK) {
    const uint32_t hdr[2] = {1u, S->N};
    std::memcpy(resp_out, hdr, 8);
    std::memcpy(resp_out + 8, B->h_resp + size_t(row) * S->N, size_t(S->N) * 4);
  }
  bool vacated = false;
  {
    std::lock_guard<std::mutex> lk(S->mu);
    if (++B->picked == B->count) {  // last one out resets the slot
      B->count = B->issued = B->picked = B->n_pull = 0;
      B->closed = B->finished = false;
      B->rc = CHPIR_OK;
      vacated = true;
    }
  }
  if (vacated) S->cv_free.notify_all();
  return rc;
}

// Gather of the hint column slices on rank 0 and download as ONE wire-format matrix (Matrix::to_bytes of lwe x N).
int gather_hint(chpir_cluster_server *S, uint8_t *hint_out, size_t hint_cap, size_t *hint_len) {
  const uint32_t m = S->lwe;
  const size_t need = 8 + size_t(m) * S->N * 4;
  if (!hint_out || hint_cap < need) return CHPIR_ERR_BUFFER_TOO_SMALL;
  const double t0 = now_s();
  Rank &R0 = S->r[0];
  CHPIR_CUDA(cudaSetDevice(R0.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  DevBuf full, stage;
  if (int rc = full.alloc(size_t(m) * S->N * 4); rc != CHPIR_OK) return rc;
  const bool use_nccl = S->n > 1 && !env_is("CHPIR_CLUSTER_GATHER", "p2p");
  if (use_nccl) {
    std::lock_guard<std::mutex> g(S->cl->mu);
    if (int rc = S->cl->ensure_comms(); rc != CHPIR_OK) return rc;
    NcclApi *api = nccl_api();
    uint64_t stage_words = 0;
    for (uint32_t d = 1; d < S->n; d++) stage_words += uint64_t(m) * S->r[d].pl.nc;
    if (int rc = stage.alloc(stage_words * 4); rc != CHPIR_OK) return rc;
    // rank d sends its contiguous lwe x nc_d block, rank 0 receives the blocks one behind the other
    ncclResult_t nr = api->GroupStart();
    uint64_t off = 0;
    for (uint32_t d = 1; d < S->n && nr == ncclSuccess; d++) {
      const uint64_t words = uint64_t(m) * S->r[d].pl.nc;
      nr = api->Send(S->r[d].srv->d_hint, words, ncclUint32, 0, S->cl->comms[d], S->r[d].compute);
      if (nr == ncclSuccess) nr = api->Recv(stage.as<uint32_t>() + off, words, ncclUint32, int(d), S->cl->comms[0], R0.compute);
      off += words;
    }
    const ncclResult_t ne = api->GroupEnd();
    if (nr != ncclSuccess || ne != ncclSuccess) return CHPIR_ERR_NCCL_FAILED;
    CHPIR_CUDA(cudaSetDevice(R0.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    // re-interleave on rank 0: block d -> columns [c0_d, c0_d + nc_d) of the row-major lwe x N matrix
    off = 0;
    for (uint32_t d = 0; d < S->n; d++) {
      const Plan &pl = S->r[d].pl;
      const uint32_t *blk = d == 0 ? R0.srv->d_hint : stage.as<uint32_t>() + off;
      CHPIR_CUDA(cudaMemcpy2DAsync(full.as<uint32_t>() + pl.c0, size_t(S->N) * 4, blk, size_t(pl.nc) * 4, size_t(pl.nc) * 4, m, cudaMemcpyDeviceToDevice, R0.compute),
                 CHPIR_ERR_CUDA_TRANSFER_FAILED);
      if (d > 0) off += uint64_t(m) * pl.nc;
    }
    S->gather_uses_nccl = 1;
  } else {
    // peer copies: rank 0 pulls every block straight into its columns (strided copy-engine transfers over NVLink)
    for (uint32_t d = 0; d < S->n; d++) {
      const Plan &pl = S->r[d].pl;
      CHPIR_CUDA(cudaMemcpy2DAsync(full.as<uint32_t>() + pl.c0, size_t(S->N) * 4, S->r[d].srv->d_hint, size_t(pl.nc) * 4, size_t(pl.nc) * 4, m, cudaMemcpyDefault,
                                   R0.compute),
                 CHPIR_ERR_CUDA_TRANSFER_FAILED);
    }
  }
  const uint32_t hdr[2] = {m, S->N};
  std::memcpy(hint_out, hdr, 8);
  CHPIR_CUDA(cudaMemcpyAsync(hint_out + 8, full.p, size_t(m) * S->N * 4, cudaMemcpyDeviceToHost, R0.compute), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  for (uint32_t d = 0; d < S->n; d++) {
    cudaSetDevice(S->r[d].dev);
    const cudaError_t e = cudaStreamSynchronize(S->r[d].compute);
    if (e != cudaSuccess) {
      set_last_cuda_error(e, "cluster setup: hint gather");
      return CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
    }
  }
  cudaSetDevice(R0.dev);
  if (hint_len) *hint_len = need;
  S->hint_gather_s = now_s() - t0;
  return CHPIR_OK;
}

int check_opts(const chpir_cluster *cl, const chpir_setup_opts *opts, chpir_setup_opts *o) {
  *o = chpir_setup_opts{};
  if (opts) *o = *opts;
  if (o->col_begin != 0 || o->col_count != 0 || o->hint_on_device != 0) return CHPIR_ERR_INVALID_ARGUMENT;
  return CHPIR_OK;
}

template <class F>
int for_each_rank_parallel(uint32_t n, F f) {
  std::vector<int> rcs(n, CHPIR_OK);
  std::vector<std::thread> th;
  for (uint32_t d = 1; d < n; d++)
    th.emplace_back([&, d] {
      try {
        rcs[d] = f(d);
      } catch (...) {
        rcs[d] = CHPIR_ERR_INVALID_ARGUMENT;
      }
    });
  try {
    rcs[0] = f(0);
  } catch (...) {
    rcs[0] = CHPIR_ERR_INVALID_ARGUMENT;
  }
  for (auto &t : th) t.join();
  for (int rc : rcs)
    if (rc != CHPIR_OK) return rc;
  return CHPIR_OK;
}

// ncclCommInitAll takes about a second; it depends on nothing the setup computes, so it runs beside the XOF / GEMM phase and the hint
// gather finds the communicators ready.
struct EarlyNccl {
  std::thread th;
  EarlyNccl(chpir_cluster *cl, uint32_t n, const chpir_setup_opts &o) {
    if (n > 1 && !o.skip_hint && !env_is("CHPIR_CLUSTER_GATHER", "p2p"))
      th = std::thread([cl] {
        DeviceRestore restore_device;
        std::lock_guard<std::mutex> g(cl->mu);
        if (cl->ensure_comms() == CHPIR_OK) (void)cl->warm_comms();  // a failure is reported by the gather, which asks again
      });
  }
  void wait() {
    if (th.joinable()) th.join();
  }
  ~EarlyNccl() { wait(); }
};

// ONE XOF chain for the whole cluster.  Every rank needs all of A = generate_from_seed(lwe, K, seed) and the squeeze is serial, so
// the first rank that has no cached A runs the host pipeline (one producer core, one upload over its PCIe link) and its uploader
// forwards every finished 128-row panel from its HBM to the other ranks' rings over NVLink (HostAPipe::start_mirror): 8.4 GB cross
// PCIe once instead of n times, n - 1 host cores and their memory traffic are free for the filter / row encoding, and the chain runs
// at its solo speed (n producers side by side: 105-107 ns per permutation and up to 0.5 s of stalls at n = 8 instead of 104 ns).
// CHPIR_CLUSTER_XOF=per_rank keeps one chain per rank (the measured comparison).  all_panels: rings as deep as A (the chain starts
// before the consumers exist, or A is to be kept); otherwise two panels.
struct SharedChain {
  std::vector<std::unique_ptr<HostAPipe>> pipes;  // per rank; null where A comes from the ctx cache or no host chain is wanted
  HostAPipe *leader = nullptr;
  ~SharedChain() {
    if (leader) leader->shutdown();  // its uploader writes into the mirrors: it goes first
  }
  HostAPipe *of(uint32_t d) const { return pipes[d].get(); }
  int start(chpir_cluster_server *S, const chpir_setup_opts &o, const uint8_t *seed, uint64_t K, bool all_panels) {
    pipes.resize(S->n);
    if (o.a_expand == CHPIR_A_EXPAND_DEVICE || o.skip_hint || o.gemm_variant != 0) return CHPIR_OK;
    const uint32_t panels = (S->lwe + 127) / 128, depth = all_panels || o.a_cache ? panels : 2;
    const bool per_rank = env_is("CHPIR_CLUSTER_XOF", "per_rank");
    std::vector<uint32_t> need;
    for (uint32_t d = 0; d < S->n; d++) {
      bool cached = false;
      if (o.a_cache) {
        std::lock_guard<std::mutex> g(S->r[d].ctx->mu);
        cached = S->r[d].ctx->a_cache.matches(seed, S->lwe, K);
      }
      if (!cached) need.push_back(d);
    }
    if (need.empty()) return CHPIR_OK;
    if (per_rank) {
      for (uint32_t d : need) {
        pipes[d].reset(new HostAPipe());
        if (int rc = pipes[d]->start(S->r[d].dev, seed, S->lwe, K, o.host_chunk_rows, depth); rc != CHPIR_OK) return rc;
      }
      return CHPIR_OK;
    }
    // the chain first (it is the critical path of the whole setup), the mirrors' rings while it is already squeezing: the leader
    // starts forwarding with its first complete panel, ~0.4 s in
    std::vector<HostAPipe *> mirrors;
    for (size_t i = 1; i < need.size(); i++) {
      pipes[need[i]].reset(new HostAPipe());
      mirrors.push_back(pipes[need[i]].get());
    }
    pipes[need[0]].reset(new HostAPipe());
    leader = pipes[need[0]].get();
    // (the leader's own start-up first: the mirrors' allocations would queue in front of its ring and hold its uploader back while
    // the chunk ring fills; its forwarder thread waits for `mirrors_ready`, the uploader does not)
    std::vector<int> rcs(need.size(), CHPIR_OK);
    rcs[0] = leader->start(S->r[need[0]].dev, seed, S->lwe, K, o.host_chunk_rows, depth, mirrors);
    std::vector<std::thread> th;
    if (rcs[0] == CHPIR_OK)
      for (size_t i = 1; i < need.size(); i++)
        th.emplace_back([&, i] { rcs[i] = pipes[need[i]]->start_mirror(S->r[need[i]].dev, S->lwe, K, depth); });
    for (auto &t : th) t.join();
    for (int rc : rcs)
      if (rc != CHPIR_OK) {
        leader->shutdown();  // before the first forward: a mirror without a ring must never be written to
        return rc;
      }
    leader->mirrors_ready();
    return CHPIR_OK;
  }
};

// Where a setup flavour left D: the whole matrix in host memory, or the ranks' compact column slices in their HBM.
struct DSource {
  const uint32_t *host = nullptr;
  const uint32_t *const *dev_slices = nullptr;
};

// Row cut: every rank assembles rows [k0, k0 + kn) of D at full width (zero rows up to the common pitch ks, so that every shard is
// ks x N and a query slice is a whole row of it), packs them (and splits the limb planes when the column slice had them), and the
// column slices go away.  The shards are ordinary single-GPU servers without a hint.
int reshard_rows(chpir_cluster_server *S, const uint8_t *seed, const DSource &src) {
  const double t0 = now_s();
  const uint64_t N = S->N, ks = S->ks;
  int rc = for_each_rank_parallel(S->n, [&](uint32_t d) -> int {
    Rank &R = S->r[d];
    const Plan &pl = R.pl;
    CHPIR_CUDA(cudaSetDevice(R.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    DevBuf blk;
    if (int rc = blk.alloc(ks * N * 4); rc != CHPIR_OK) return rc;
    cudaStream_t st = R.compute;
    uint32_t *rows = blk.as<uint32_t>();
    if (pl.kn < ks) CHPIR_CUDA(cudaMemsetAsync(rows + pl.kn * N, 0, (ks - pl.kn) * N * 4, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
    if (pl.kn > 0) {
      if (src.host) {
        CHPIR_CUDA(cudaMemcpyAsync(rows, src.host + pl.k0 * N, pl.kn * N * 4, cudaMemcpyHostToDevice, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
      } else {
        for (uint32_t s0 = 0; s0 < S->n; s0++) {
          const uint32_t s = (d + s0) % S->n;  // every rank starts with a different peer
          const Plan &ps = S->r[s].pl;
          CHPIR_CUDA(cudaMemcpy2DAsync(rows + ps.c0, N * 4, src.dev_slices[s] + pl.k0 * ps.nc, size_t(ps.nc) * 4, size_t(ps.nc) * 4, pl.kn, cudaMemcpyDefault, st),
                     CHPIR_ERR_CUDA_TRANSFER_FAILED);
        }
      }
    }
    CHPIR_CUDA(cudaStreamSynchronize(st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
    chpir_setup_opts ro{};
    ro.skip_hint = 1;
    ro.batch_tc = R.srv->gemm ? 1 : 2;
    chpir_server *row = nullptr;
    if (int rc = chpir_server_setup_device(R.ctx, seed, rows, ks, uint32_t(N), S->b, &ro, nullptr, 0, nullptr, &row); rc != CHPIR_OK) return rc;
    // what the caller reads through chpir_cluster_server_shard: the setup phases of this rank (measured on the column slice) and the
    // resident row block
    const double pack_s = row->timing.pack_s;
    row->timing = R.srv->timing;
    row->timing.pack_s += pack_s;
    row->last_gemm_ms = R.srv->last_gemm_ms, row->last_expand_ms = R.srv->last_expand_ms;
    row->shard_k_total = S->K;
    R.col = R.srv;
    R.srv = row;
    return CHPIR_OK;
  });
  for (uint32_t d = 0; d < S->n; d++) {
    if (!S->r[d].col) continue;
    cudaSetDevice(S->r[d].dev);
    chpir_server_destroy(S->r[d].col);
    S->r[d].col = nullptr;
  }
  S->reshard_s = now_s() - t0;
  return rc;
}

// Common tail of every setup flavour: per-rank column-slice servers exist in S->r[*].srv.
int complete_setup(chpir_cluster_server *S, const chpir_setup_opts &o, const uint8_t *seed, const DSource &src, uint8_t *hint_out, size_t hint_cap,
                   size_t *hint_len) {
  for (uint32_t d = 0; d < S->n; d++) S->r[d].srv->col_begin = S->r[d].pl.c0;  // logical position of a compact slice
  if (int rc = make_streams(S); rc != CHPIR_OK) return rc;
  if (hint_len) *hint_len = 0;
  if (!o.skip_hint) {
    if (int rc = gather_hint(S, hint_out, hint_cap, hint_len); rc != CHPIR_OK) return rc;
  }
  if (S->rows) {
    if (int rc = reshard_rows(S, seed, src); rc != CHPIR_OK) return rc;
  }
  return make_slots(S);
}

chpir_cluster_server *new_server(chpir_cluster *cl, uint64_t K, uint32_t N, uint32_t b, const chpir_setup_opts &o) {
  chpir_cluster_server *S = new chpir_cluster_server();
  S->cl = cl, S->n = uint32_t(cl->n), S->K = K, S->N = N, S->b = b;
  S->lwe = o.lwe_rows ? o.lwe_rows : CHPIR_LWE_DIMENSION;
  S->r.resize(S->n);
  for (uint32_t d = 0; d < S->n; d++) {
    S->r[d].dev = cl->dev[d], S->r[d].ctx = cl->ctx[d];
    S->r[d].pl = plan_of(S->n, d, K, N);
  }
  S->ks = S->r[0].pl.ks;
  S->rows = S->n > 1 && !env_is("CHPIR_CLUSTER_SHARD", "cols");
  S->pipeline = S->n > 1 || o.respond_coalesce != 0;
  return S;
}

// per-rank options: the rank's columns, hint slice kept in HBM for the gather, no per-shard coalescer (the cluster has its own)
chpir_setup_opts rank_opts(const chpir_setup_opts &o, const Plan &pl, bool compact) {
  chpir_setup_opts ro = o;
  ro.col_begin = compact ? 0 : pl.c0;
  ro.col_count = compact ? 0 : pl.nc;
  ro.hint_on_device = 1;
  ro.respond_coalesce = 0;
  return ro;
}

// Device-resident respond on the row cut: every rank streams its own rows against the query words resident on it, nothing is
// gathered; rank 0's second stream adds the partial sums of pass i while the compute streams are already in pass i + 1.
int respond_device_rows(chpir_cluster_server *S, const uint32_t *const *q_slices, uint32_t nq, uint32_t *resp_device0, bool tc, uint32_t repeats,
                        float *device_ms) {
  for (uint32_t d = 0; d < S->n; d++) {
    Rank &R = S->r[d];
    CHPIR_CUDA(cudaSetDevice(R.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    if (R.resp_rows < nq) {
      for (int p = 0; p < 2; p++) {
        if (R.resp[p]) cudaFree(R.resp[p]);
        R.resp[p] = nullptr;
      }
      R.resp_rows = 0;
      for (int p = 0; p < 2; p++)
        if (cudaMalloc(&R.resp[p], size_t(nq) * S->N * 4) != cudaSuccess) {
          set_last_cuda_error(cudaGetLastError(), "cluster device-path scratch (partial sums)");
          return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
        }
      R.resp_rows = nq;
    }
  }
  Rank &R0 = S->r[0];
  CHPIR_CUDA(cudaSetDevice(R0.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  CHPIR_CUDA(cudaEventRecord(S->t0, R0.compute), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
  CHPIR_CUDA(cudaStreamWaitEvent(R0.gather, S->t0, 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
  for (uint32_t d = 1; d < S->n; d++) {
    CHPIR_CUDA(cudaSetDevice(S->r[d].dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    CHPIR_CUDA(cudaStreamWaitEvent(S->r[d].compute, S->t0, 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
  }
  int rc = CHPIR_OK;
  for (uint32_t rep = 0; rep < repeats && rc == CHPIR_OK; rep++) {
    const int p = int(rep & 1);
    PartTable parts{};
    for (uint32_t d = 0; d < S->n && rc == CHPIR_OK; d++) {
      Rank &R = S->r[d];
      chpir_server *sv = R.srv;
      parts.p[d] = R.resp[p];
      CHPIR_CUDA(cudaSetDevice(R.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
      cudaStream_t st = R.compute;
      CHPIR_CUDA(cudaStreamWaitEvent(st, R0.reduced[p], 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);  // pass rep-2 has been summed out of this buffer
      CHPIR_CUDA(cudaMemsetAsync(R.resp[p], 0, size_t(nq) * S->N * 4, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
      if (tc) {
        std::lock_guard<std::mutex> gg(sv->gemm_mu);
        for (uint32_t row0 = 0; row0 < nq && rc == CHPIR_OK; row0 += kMaxBatch) {
          const uint32_t rows = std::min(kMaxBatch, nq - row0);
          const int buf = int(R.tc_buf++ & 1);
          if ((rc = gemm_tc_buf_acquire(sv->gemm, buf, st)) != CHPIR_OK) break;
          if ((rc = gemm_tc_load_panel_u32(sv->gemm, buf, q_slices[d] + size_t(row0) * S->ks, rows, st)) != CHPIR_OK) break;
          if ((rc = gemm_tc_panel(sv->gemm, buf, rows, R.resp[p] + size_t(row0) * S->N, st)) != CHPIR_OK) break;
          if ((rc = gemm_tc_buf_release(sv->gemm, buf, st)) != CHPIR_OK) break;
        }
      } else {
        rc = launch_respond(sv->d_packed, sv->layout, sv->K, sv->plan, q_slices[d], R.resp[p], nq, st);
      }
      if (rc != CHPIR_OK) break;
      CHPIR_CUDA(cudaEventRecord(R.computed[p], st), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
    }
    if (rc != CHPIR_OK) break;
    CHPIR_CUDA(cudaSetDevice(R0.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    for (uint32_t d = 0; d < S->n; d++) CHPIR_CUDA(cudaStreamWaitEvent(R0.gather, S->r[d].computed[p], 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
    if ((rc = launch_reduce(parts, S->n, uint64_t(nq) * S->N, resp_device0, R0.ctx->sm_count, R0.gather)) != CHPIR_OK) break;
    CHPIR_CUDA(cudaEventRecord(R0.reduced[p], R0.gather), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
  }
  if (rc == CHPIR_OK) {
    cudaSetDevice(R0.dev);
    for (int p = 0; p < 2; p++) cudaStreamWaitEvent(R0.compute, R0.reduced[p], 0);
    cudaEventRecord(S->t1, R0.compute);
  }
  for (uint32_t d = 0; d < S->n; d++) {
    cudaSetDevice(S->r[d].dev);
    cudaError_t e = cudaStreamSynchronize(S->r[d].gather);
    if (e == cudaSuccess) e = cudaStreamSynchronize(S->r[d].compute);
    if (e != cudaSuccess && rc == CHPIR_OK) {
      set_last_cuda_error(e, "cluster respond (device-resident, row cut)");
      rc = CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
    }
  }
  cudaSetDevice(R0.dev);
  if (rc == CHPIR_OK && device_ms) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, S->t0, S->t1) == cudaSuccess) *device_ms = ms;
  }
  return rc;
}

}  // namespace

extern "C" {

int chpir_cluster_plan(uint32_t n_ranks, uint32_t rank, uint64_t rows_k, uint32_t cols_n, uint32_t *col_begin, uint32_t *col_count, uint64_t *k_begin,
                       uint64_t *k_count, uint64_t *k_pitch) {
  if (n_ranks == 0 || n_ranks > kMaxRanks || rank >= n_ranks || rows_k == 0 || cols_n == 0) return CHPIR_ERR_INVALID_ARGUMENT;
  const Plan p = plan_of(n_ranks, rank, rows_k, cols_n);
  if (col_begin) *col_begin = p.c0;
  if (col_count) *col_count = p.nc;
  if (k_begin) *k_begin = p.k0;
  if (k_count) *k_count = p.kn;
  if (k_pitch) *k_pitch = p.ks;
  return CHPIR_OK;
}

int chpir_cluster_create(int n_gpus, const int *device_ordinals, chpir_cluster **out) {
  CHPIR_GUARD_BEGIN
  DeviceRestore restore_device;
  if (!out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (n_gpus == 0) n_gpus = int(env_u32("CHPIR_GPUS", 1));
  if (n_gpus < 1 || n_gpus > int(kMaxRanks)) return CHPIR_ERR_INVALID_ARGUMENT;
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) {
    (void)cudaGetLastError();
    return CHPIR_ERR_CUDA_DEVICE_NOT_FOUND;
  }
  std::unique_ptr<chpir_cluster> cl(new chpir_cluster());
  cl->n = n_gpus;
  for (int i = 0; i < n_gpus; i++) {
    const int d = device_ordinals ? device_ordinals[i] : i;
    if (d < 0 || d >= have) return CHPIR_ERR_CUDA_DEVICE_NOT_FOUND;
    for (int j : cl->dev)
      if (j == d) return CHPIR_ERR_INVALID_ARGUMENT;  // a GPU can hold one rank
    cl->dev.push_back(d);
  }
  for (int i = 0; i < n_gpus; i++) {
    chpir_ctx *c = nullptr;
    if (int rc = chpir_ctx_create(cl->dev[i], &c); rc != CHPIR_OK) return rc;
    cl->ctx.push_back(c);
  }
  // every rank reads every other rank's query slices (and rank 0 the hint slices): full peer access, both directions
  for (int i = 0; i < n_gpus; i++)
    for (int j = 0; j < n_gpus; j++) {
      if (i == j) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, cl->dev[i], cl->dev[j]) != cudaSuccess || !can) {
        (void)cudaGetLastError();
        return CHPIR_ERR_CUDA_PEER_ACCESS_UNAVAILABLE;
      }
      CHPIR_CUDA(cudaSetDevice(cl->dev[i]), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
      const cudaError_t e = cudaDeviceEnablePeerAccess(cl->dev[j], 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        (void)cudaGetLastError();
      } else if (e != cudaSuccess) {
        set_last_cuda_error(e, "cudaDeviceEnablePeerAccess");
        return CHPIR_ERR_CUDA_PEER_ACCESS_UNAVAILABLE;
      }
    }
  *out = cl.release();
  return CHPIR_OK;
  CHPIR_GUARD_END
}

void chpir_cluster_destroy(chpir_cluster *cluster) {
  DeviceRestore restore_device;
  delete cluster;
}

int chpir_cluster_size(const chpir_cluster *cluster, int *n_gpus) {
  if (!cluster || !n_gpus) return CHPIR_ERR_INVALID_ARGUMENT;
  *n_gpus = cluster->n;
  return CHPIR_OK;
}

int chpir_cluster_ctx(const chpir_cluster *cluster, int rank, chpir_ctx **ctx, int *device_ordinal) {
  if (!cluster || rank < 0 || rank >= cluster->n) return CHPIR_ERR_INVALID_ARGUMENT;
  if (ctx) *ctx = cluster->ctx[size_t(rank)];
  if (device_ordinal) *device_ordinal = cluster->dev[size_t(rank)];
  return CHPIR_OK;
}

int chpir_cluster_server_setup_device(chpir_cluster *cl, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint32_t *const *d_slices, uint64_t rows_k,
                                      uint32_t cols_n, uint32_t b, const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap, size_t *hint_len,
                                      chpir_cluster_server **out) {
  CHPIR_GUARD_BEGIN
  DeviceRestore restore_device;
  if (!out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (!cl || !seed || !d_slices) return CHPIR_ERR_INVALID_ARGUMENT;
  if (rows_k == 0 || cols_n == 0) return CHPIR_ERR_INVALID_MATRIX_DIMENSION;
  if (cols_n < uint32_t(cl->n)) return CHPIR_ERR_INVALID_ARGUMENT;  // every rank needs at least one column
  chpir_setup_opts o;
  if (int rc = check_opts(cl, opts, &o); rc != CHPIR_OK) return rc;
  const double t0 = now_s();
  std::unique_ptr<chpir_cluster_server> S(new_server(cl, rows_k, cols_n, b, o));
  EarlyNccl nccl(cl, S->n, o);
  SharedChain chain;
  if (S->n > 1)
    if (int rc = chain.start(S.get(), o, seed, rows_k, false); rc != CHPIR_OK) return rc;
  int rc = for_each_rank_parallel(S->n, [&](uint32_t d) -> int {
    if (!d_slices[d]) return CHPIR_ERR_INVALID_ARGUMENT;
    const chpir_setup_opts ro = rank_opts(o, S->r[d].pl, true);
    return server_setup_from_device_matrix(S->r[d].ctx, seed, d_slices[d], rows_k, S->r[d].pl.nc, b, &ro, nullptr, 0, nullptr, &S->r[d].srv,
                                           S->n > 1 ? chain.of(d) : nullptr);
  });
  if (rc != CHPIR_OK) return rc;
  nccl.wait();
  DSource dsrc;
  dsrc.dev_slices = d_slices;
  if ((rc = complete_setup(S.get(), o, seed, dsrc, hint_out, hint_cap, hint_len)) != CHPIR_OK) return rc;
  S->setup_total_s = now_s() - t0;
  *out = S.release();
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_cluster_server_setup(chpir_cluster *cl, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint32_t *d_host, uint64_t rows_k, uint32_t cols_n,
                               uint32_t b, const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap, size_t *hint_len,
                               chpir_cluster_server **out) {
  CHPIR_GUARD_BEGIN
  DeviceRestore restore_device;
  if (!out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (!cl || !seed || !d_host) return CHPIR_ERR_INVALID_ARGUMENT;
  if (rows_k == 0 || cols_n == 0) return CHPIR_ERR_INVALID_MATRIX_DIMENSION;
  if (cols_n < uint32_t(cl->n)) return CHPIR_ERR_INVALID_ARGUMENT;
  chpir_setup_opts o;
  if (int rc = check_opts(cl, opts, &o); rc != CHPIR_OK) return rc;
  const double t0 = now_s();
  std::unique_ptr<chpir_cluster_server> S(new_server(cl, rows_k, cols_n, b, o));
  EarlyNccl nccl(cl, S->n, o);
  SharedChain chain;
  if (S->n > 1)
    if (int rc = chain.start(S.get(), o, seed, rows_k, false); rc != CHPIR_OK) return rc;
  int rc = for_each_rank_parallel(S->n, [&](uint32_t d) -> int {
    const chpir_setup_opts ro = rank_opts(o, S->r[d].pl, false);
    return server_setup_from_host_matrix(S->r[d].ctx, seed, d_host, rows_k, cols_n, b, &ro, nullptr, 0, nullptr, &S->r[d].srv,
                                         S->n > 1 ? chain.of(d) : nullptr);
  });
  if (rc != CHPIR_OK) return rc;
  nccl.wait();
  DSource dsrc;
  dsrc.host = d_host;
  if ((rc = complete_setup(S.get(), o, seed, dsrc, hint_out, hint_cap, hint_len)) != CHPIR_OK) return rc;
  S->setup_total_s = now_s() - t0;
  *out = S.release();
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_cluster_server_setup_from_db(chpir_cluster *cl, uint32_t arity, const uint8_t seed[CHPIR_SEED_BYTE_LEN], uint64_t n, const uint8_t *key_blob,
                                       const uint64_t *key_offsets, const uint8_t *value_blob, const uint64_t *value_offsets,
                                       const uint64_t *filter_seed_rng, const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap,
                                       size_t *hint_len, uint8_t filter_params_out[CHPIR_FILTER_PARAM_BYTE_LEN], chpir_cluster_server **out) {
  CHPIR_GUARD_BEGIN
  DeviceRestore restore_device;
  if (!out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (n == 0) return CHPIR_ERR_EMPTY_KV_DATABASE;  // server.rs:104-107
  if (!cl || !seed || !key_blob || !key_offsets || !value_blob || !value_offsets || !filter_params_out) return CHPIR_ERR_INVALID_ARGUMENT;
  if (arity != 3 && arity != 4) return CHPIR_ERR_UNSUPPORTED_ARITY_FOR_BINARY_FUSE_FILTER;
  chpir_setup_opts o;
  if (int rc = check_opts(cl, opts, &o); rc != CHPIR_OK) return rc;
  uint32_t b = 0;
  if (int rc = find_mat_elem_bit_len(n, &b); rc != CHPIR_OK) return rc;
  uint64_t max_vlen = 0;
  for (uint64_t i = 0; i < n; i++) max_vlen = std::max<uint64_t>(max_vlen, value_offsets[i + 1] - value_offsets[i]);
  uint64_t K = 0, N = 0;
  if (int rc = db_matrix_shape(arity, n, max_vlen, b, &K, &N); rc != CHPIR_OK) return rc;
  if (N > 0xffffffffull) return CHPIR_ERR_KV_DATABASE_SIZE_TOO_LARGE;
  if (N < uint64_t(cl->n)) return CHPIR_ERR_INVALID_ARGUMENT;
  const double t0 = now_s();
  std::unique_ptr<chpir_cluster_server> S(new_server(cl, K, uint32_t(N), b, o));
  EarlyNccl nccl(cl, S->n, o);
  std::unique_ptr<uint32_t[]> d_store;  // n > 1: D, encoded once on the host; lives until the row blocks have been cut from it
  SharedChain chain;
  std::vector<DevBuf> d_dev;              // n > 1, db_encode = device: every rank's columns of D, built in its HBM
  std::vector<const uint32_t *> d_ptrs;
  if (S->n == 1) {
    // one GPU: the single-GPU call as it is (device row fill, its own early XOF start), hint slice = whole hint
    chpir_setup_opts ro = o;
    ro.hint_on_device = 1;
    ro.respond_coalesce = 0;
    if (int rc = chpir_server_setup_from_db(S->r[0].ctx, arity, seed, n, key_blob, key_offsets, value_blob, value_offsets, filter_seed_rng, &ro, nullptr, 0,
                                            nullptr, filter_params_out, &S->r[0].srv);
        rc != CHPIR_OK)
      return rc;
  } else if (o.db_encode == CHPIR_DB_ENCODE_DEVICE) {
    // Row encoding + dependent fill on the GPUs: the host does the key digests, the peeling and the wave plan ONCE; every rank
    // receives the raw values over its own PCIe link (beside the peeling) and builds ITS columns of D in its HBM -- the recurrence
    // couples rows, never columns.  The host never holds D, and the row blocks are cut from the slices over NVLink afterwards.
    if (int rc = chain.start(S.get(), o, seed, K, true); rc != CHPIR_OK) return rc;
    const double t_enc0 = now_s();
    std::vector<std::unique_ptr<DeviceFillRank>> fill(S->n);
    d_dev.resize(S->n);
    int rc = for_each_rank_parallel(S->n, [&](uint32_t d) -> int {
      fill[d].reset(new DeviceFillRank());
      return fill[d]->begin(S->r[d].ctx, n, value_blob, value_offsets, K, S->r[d].pl.nc, &d_dev[d]);
    });
    if (rc != CHPIR_OK) return rc;
    DeviceFillHost fh;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (chain.leader) set_encode_threads(hw > 3 ? hw - 2 : 1);  // the producer core stays free for the chain
    rc = fh.prepare(arity, n, key_blob, key_offsets, b, CHPIR_SERVER_SETUP_MAX_ATTEMPT_COUNT, filter_seed_rng);
    set_encode_threads(0);
    if (rc != CHPIR_OK) return rc;
    std::memcpy(filter_params_out, fh.filter_bytes, CHPIR_FILTER_PARAM_BYTE_LEN);
    std::vector<double> fill_s(S->n, 0.0);
    rc = for_each_rank_parallel(S->n, [&](uint32_t d) -> int { return fill[d]->finish(arity, fh, N, S->r[d].pl.c0, b, &fill_s[d]); });
    fill.clear();  // values, records: gone before the hint phase allocates
    if (rc != CHPIR_OK) return rc;
    const double t_enc1 = now_s();
    rc = for_each_rank_parallel(S->n, [&](uint32_t d) -> int {
      const chpir_setup_opts ro = rank_opts(o, S->r[d].pl, true);
      return server_setup_from_device_matrix(S->r[d].ctx, seed, d_dev[d].as<uint32_t>(), K, S->r[d].pl.nc, b, &ro, nullptr, 0, nullptr, &S->r[d].srv,
                                             chain.of(d));
    });
    if (rc != CHPIR_OK) return rc;
    for (uint32_t d = 0; d < S->n; d++) {
      S->r[d].srv->timing.host_encode_s = t_enc1 - t_enc0;
      S->r[d].srv->timing.device_encode_s = fill_s[d];
    }
    for (uint32_t d = 0; d < S->n; d++) d_ptrs.push_back(d_dev[d].as<uint32_t>());
  } else {
    // Every rank needs all of A = generate_from_seed(lwe, K, seed), and the XOF chain is serial: it starts NOW (one producer core,
    // rings as deep as A on every GPU, panels forwarded over NVLink) and squeezes beside the host filter/encode phase below, exactly
    // as the single-GPU call does; D is encoded once on the host and every rank uploads its own columns.
    const bool host_a = o.a_expand != CHPIR_A_EXPAND_DEVICE && !o.skip_hint && o.gemm_variant == 0;
    if (int rc = chain.start(S.get(), o, seed, K, true); rc != CHPIR_OK) return rc;
    d_store.reset(new (std::nothrow) uint32_t[K * N]);
    if (!d_store) return CHPIR_ERR_HOST_ALLOCATION_FAILED;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned chains = chain.leader ? 1u : S->n;
    if (host_a) set_encode_threads(hw > chains + 2 ? hw - chains - 1 : 1);  // the producer cores stay free for the chains
    int rc = encode_kv_database(arity, n, key_blob, key_offsets, value_blob, value_offsets, b, CHPIR_SERVER_SETUP_MAX_ATTEMPT_COUNT, filter_seed_rng,
                                d_store.get(), filter_params_out);
    set_encode_threads(0);
    if (rc != CHPIR_OK) return rc;
    const double t1 = now_s();
    rc = for_each_rank_parallel(S->n, [&](uint32_t d) -> int {
      const chpir_setup_opts ro = rank_opts(o, S->r[d].pl, false);
      return server_setup_from_host_matrix(S->r[d].ctx, seed, d_store.get(), K, uint32_t(N), b, &ro, nullptr, 0, nullptr, &S->r[d].srv, chain.of(d));
    });
    if (rc != CHPIR_OK) return rc;
    for (uint32_t d = 0; d < S->n; d++) S->r[d].srv->timing.host_encode_s = t1 - t0;
  }
  nccl.wait();
  DSource dsrc;
  dsrc.host = d_store.get();
  if (!d_ptrs.empty()) dsrc.dev_slices = d_ptrs.data();
  if (int rc = complete_setup(S.get(), o, seed, dsrc, hint_out, hint_cap, hint_len); rc != CHPIR_OK) return rc;
  d_store.reset();
  d_dev.clear();
  S->setup_total_s = now_s() - t0;
  *out = S.release();
  return CHPIR_OK;
  CHPIR_GUARD_END
}

void chpir_cluster_server_destroy(chpir_cluster_server *srv) {
  DeviceRestore restore_device;
  delete srv;
}

int chpir_cluster_server_shard(const chpir_cluster_server *srv, int rank, chpir_server **shard) {
  if (!srv || !shard || rank < 0 || uint32_t(rank) >= srv->n) return CHPIR_ERR_INVALID_ARGUMENT;
  *shard = srv->r[size_t(rank)].srv;
  return CHPIR_OK;
}

int chpir_cluster_server_save(chpir_cluster_server *srv, const char *path_prefix) {
  CHPIR_GUARD_BEGIN
  DeviceRestore restore_device;
  if (!srv || !path_prefix) return CHPIR_ERR_INVALID_ARGUMENT;
  for (uint32_t d = 0; d < srv->n; d++) {
    const std::string p = std::string(path_prefix) + ".rank" + std::to_string(d) + "of" + std::to_string(srv->n);
    if (int rc = chpir_server_save(srv->r[d].srv, p.c_str()); rc != CHPIR_OK) return rc;
  }
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_cluster_server_load(chpir_cluster *cl, const char *path_prefix, const chpir_setup_opts *opts, chpir_cluster_server **out) {
  CHPIR_GUARD_BEGIN
  DeviceRestore restore_device;
  if (!out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (!cl || !path_prefix) return CHPIR_ERR_INVALID_ARGUMENT;
  chpir_setup_opts o;
  if (int rc = check_opts(cl, opts, &o); rc != CHPIR_OK) return rc;
  const uint32_t n = uint32_t(cl->n);
  std::vector<chpir_server *> shards(n, nullptr);
  struct Cleanup {
    std::vector<chpir_server *> *v;
    ~Cleanup() {
      for (chpir_server *s : *v)
        if (s) chpir_server_destroy(s);
    }
  } cleanup{&shards};
  chpir_setup_opts ro = o;
  ro.respond_coalesce = 0;  // the cluster has its own coalescer
  int rc = for_each_rank_parallel(n, [&](uint32_t d) -> int {
    const std::string p = std::string(path_prefix) + ".rank" + std::to_string(d) + "of" + std::to_string(n);
    return chpir_server_load(cl->ctx[d], p.c_str(), &ro, &shards[d]);
  });
  if (rc != CHPIR_OK) return rc;
  // the files must describe ONE matrix cut by this cluster's plan: row blocks (written by an n > 1 cluster in the default cut)
  // or column slices
  const bool by_rows = shards[0]->shard_k_total != 0;
  uint64_t K = 0, ncols = 0;
  if (by_rows) {
    K = shards[0]->shard_k_total, ncols = shards[0]->ncols;
    if (n < 2) return CHPIR_ERR_INVALID_SAVED_SERVER;
    for (uint32_t d = 0; d < n; d++) {
      const Plan pl = plan_of(n, d, K, uint32_t(ncols));
      if (shards[d]->shard_k_total != K || shards[d]->K != pl.ks || shards[d]->b != shards[0]->b || shards[d]->ncols != ncols)
        return CHPIR_ERR_INVALID_SAVED_SERVER;
    }
  } else {
    K = shards[0]->K;
    for (uint32_t d = 0; d < n; d++) ncols += shards[d]->ncols;
    if (ncols > 0xffffffffull) return CHPIR_ERR_INVALID_SAVED_SERVER;
    for (uint32_t d = 0; d < n; d++) {
      const Plan pl = plan_of(n, d, K, uint32_t(ncols));
      if (shards[d]->shard_k_total != 0 || shards[d]->K != K || shards[d]->b != shards[0]->b || shards[d]->ncols != pl.nc || shards[d]->col_begin != pl.c0)
        return CHPIR_ERR_INVALID_SAVED_SERVER;
    }
  }
  std::unique_ptr<chpir_cluster_server> S(new_server(cl, K, uint32_t(ncols), shards[0]->b, o));
  S->rows = by_rows;  // the files decide, not the environment
  S->lwe = 0;         // no hint travels with a saved server
  for (uint32_t d = 0; d < n; d++) {
    S->r[d].srv = shards[d];
    shards[d] = nullptr;
  }
  if ((rc = make_streams(S.get())) != CHPIR_OK) return rc;
  if ((rc = make_slots(S.get())) != CHPIR_OK) return rc;
  *out = S.release();
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_cluster_server_respond(chpir_cluster_server *S, const uint8_t *query, size_t query_len, uint8_t *resp_out, size_t resp_cap, size_t *resp_len) {
  CHPIR_GUARD_BEGIN
  DeviceRestore restore_device;
  if (!S) return CHPIR_ERR_INVALID_ARGUMENT;
  if (!S->pipeline) return chpir_server_respond(S->r[0].srv, query, query_len, resp_out, resp_cap, resp_len);
  if (int rc = validate_query_bytes(S->K, query, query_len); rc != CHPIR_OK) return rc;
  const size_t need = 8 + size_t(S->N) * 4;
  if (!resp_out || resp_cap < need) return CHPIR_ERR_BUFFER_TOO_SMALL;
  if (!S->co_ready) return CHPIR_ERR_INVALID_ARGUMENT;
  const int rc = respond_coalesced(S, query, query_len, resp_out);
  if (rc == CHPIR_OK && resp_len) *resp_len = need;
  return rc;
  CHPIR_GUARD_END
}

int chpir_cluster_server_respond_batch(chpir_cluster_server *S, const uint8_t *const *queries, const size_t *query_lens, uint32_t nq, uint8_t *resp_out,
                                       size_t resp_stride) {
  CHPIR_GUARD_BEGIN
  DeviceRestore restore_device;
  if (!S || !queries || !query_lens || !resp_out) return CHPIR_ERR_INVALID_ARGUMENT;
  if (!S->pipeline) return chpir_server_respond_batch(S->r[0].srv, queries, query_lens, nq, resp_out, resp_stride);
  const size_t need = 8 + size_t(S->N) * 4;
  if (resp_stride < need) return CHPIR_ERR_BUFFER_TOO_SMALL;
  for (uint32_t i = 0; i < nq; i++)
    if (int rc = validate_query_bytes(S->K, queries[i], query_lens[i]); rc != CHPIR_OK) return rc;
  if (nq == 0) return CHPIR_OK;
  std::lock_guard<std::mutex> g(S->batch_mu);
  if (!S->bb) {
    S->bb.reset(new CBatch());
    if (int rc = S->init_batch(*S->bb); rc != CHPIR_OK) {
      S->free_batch(*S->bb);
      S->bb.reset();
      return rc;
    }
  }
  CBatch &B = *S->bb;
  // groups of up to 128 queries (one M tile); without limb planes the GEMV route takes gemv_rows at a time
  const uint32_t group = S->tc ? kMaxBatch : S->gemv_rows;
  const uint32_t hdr[2] = {1u, S->N};
  for (uint32_t q0 = 0; q0 < nq; q0 += group) {
    const uint32_t cnt = std::min(group, nq - q0);
    B.n_pull = 0;
    int rc = CHPIR_OK;
    for (uint32_t i = 0; i < cnt && rc == CHPIR_OK; i++) {
      const bool pull = pullable(queries[q0 + i], query_lens[q0 + i]);
      B.pull.src[i] = pull ? queries[q0 + i] : nullptr;
      B.n_pull += pull ? 1 : 0;
      if (!pull) rc = upload_query(S, B, i, queries[q0 + i]);
    }
    const int in = ingest_batch(S, B, cnt);
    if (rc == CHPIR_OK) rc = in;
    if (rc == CHPIR_OK) {
      std::lock_guard<std::mutex> ex(S->exec_mu);
      rc = run_batch(S, B, cnt, nullptr);
    }
    if (rc != CHPIR_OK) return rc;
    for (uint32_t i = 0; i < cnt; i++) {
      uint8_t *o = resp_out + size_t(q0 + i) * resp_stride;
      std::memcpy(o, hdr, 8);
      std::memcpy(o + 8, B.h_resp + size_t(i) * S->N, size_t(S->N) * 4);
    }
  }
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_cluster_server_respond_device(chpir_cluster_server *S, const uint32_t *const *q_slices, uint32_t nq, uint32_t *resp_device0, uint32_t mode,
                                        uint32_t repeats, float *device_ms) {
  CHPIR_GUARD_BEGIN
  DeviceRestore restore_device;
  if (!S || !q_slices || !resp_device0 || mode > CHPIR_RESPOND_TC) return CHPIR_ERR_INVALID_ARGUMENT;
  if (device_ms) *device_ms = 0.f;
  if (nq == 0 || repeats == 0) return CHPIR_OK;
  const bool tc = mode == CHPIR_RESPOND_TC;
  if (tc && !S->tc) return CHPIR_ERR_INVALID_ARGUMENT;  // set up without limb planes
  for (uint32_t d = 0; d < S->n; d++)
    if (!q_slices[d] || (reinterpret_cast<uintptr_t>(q_slices[d]) & 15u)) return CHPIR_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> g(S->dev_mu);
  std::lock_guard<std::mutex> ex(S->exec_mu);  // shares the compute streams and the operand rings with the coalesced route
  if (S->rows) return respond_device_rows(S, q_slices, nq, resp_device0, tc, repeats, device_ms);
  const uint32_t chunk = tc ? kMaxBatch : std::max(1u, std::min(env_u32("CHPIR_CLUSTER_GEMV_CHUNK", kGemvChunk), 4096u));
  // one rank whose slice IS the whole query (K a multiple of 32): the rows are already what the GEMV reads, nothing to gather
  const bool direct = !tc && S->n == 1 && S->ks == S->K;
  // scratch: whole-query rows (GEMV route only) and the rank's response columns, double-buffered
  for (uint32_t d = 0; d < S->n; d++) {
    Rank &R = S->r[d];
    CHPIR_CUDA(cudaSetDevice(R.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    if (!tc && !direct && R.q_rows < chunk) {
      for (int p = 0; p < 2; p++) {
        if (R.q_full[p]) cudaFree(R.q_full[p]);
        R.q_full[p] = nullptr;
      }
      R.q_rows = 0;
      for (int p = 0; p < 2; p++)
        if (cudaMalloc(&R.q_full[p], size_t(chunk) * S->K * 4) != cudaSuccess) {
          set_last_cuda_error(cudaGetLastError(), "cluster device-path scratch (query rows)");
          return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
        }
      R.q_rows = chunk;
    }
    if (R.resp_rows < chunk) {
      for (int p = 0; p < 2; p++) {
        if (R.resp[p]) cudaFree(R.resp[p]);
        R.resp[p] = nullptr;
      }
      R.resp_rows = 0;
      for (int p = 0; p < 2; p++)
        if (cudaMalloc(&R.resp[p], size_t(chunk) * R.pl.nc * 4) != cudaSuccess) {
          set_last_cuda_error(cudaGetLastError(), "cluster device-path scratch (response columns)");
          return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
        }
      R.resp_rows = chunk;
    }
  }
  Rank &R0 = S->r[0];
  // t0 on rank 0 before anything moves; every other stream starts behind it
  CHPIR_CUDA(cudaSetDevice(R0.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  CHPIR_CUDA(cudaEventRecord(S->t0, R0.compute), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
  for (uint32_t d = 0; d < S->n; d++) {
    CHPIR_CUDA(cudaSetDevice(S->r[d].dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    CHPIR_CUDA(cudaStreamWaitEvent(S->r[d].gather, S->t0, 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
    if (d != 0) CHPIR_CUDA(cudaStreamWaitEvent(S->r[d].compute, S->t0, 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
  }
  const bool ce_gather = !env_is("CHPIR_CLUSTER_QGATHER", "kernel");  // copy engines (default) or the pull kernel
  int rc = CHPIR_OK;
  uint64_t it = 0;
  for (uint32_t rep = 0; rep < repeats && rc == CHPIR_OK; rep++) {
    for (uint32_t row0 = 0; row0 < nq && rc == CHPIR_OK; row0 += chunk, it++) {
      const uint32_t rows = std::min(chunk, nq - row0);
      const int p = int(it & 1);
      SrcTable src{};
      for (uint32_t s = 0; s < S->n; s++) src.p[s] = q_slices[s] + size_t(row0) * S->ks;
      if (!tc && !direct) {
        // all-gather of the query rows on the gather streams: runs beside the previous chunk's GEMVs
        for (uint32_t d = 0; d < S->n && rc == CHPIR_OK; d++) {
          Rank &R = S->r[d];
          CHPIR_CUDA(cudaSetDevice(R.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
          CHPIR_CUDA(cudaStreamWaitEvent(R.gather, R.computed[p], 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);  // chunk it-2 has left this buffer
          if (ce_gather) {
            for (uint32_t s0 = 0; s0 < S->n; s0++) {
              const uint32_t s = (d + s0) % S->n;  // every rank starts with a different peer: no hot source
              const Plan &ps = S->r[s].pl;
              if (ps.kn == 0) continue;
              CHPIR_CUDA(cudaMemcpy2DAsync(R.q_full[p] + ps.k0, S->K * 4, src.p[s], S->ks * 4, ps.kn * 4, rows, cudaMemcpyDefault, R.gather),
                         CHPIR_ERR_CUDA_TRANSFER_FAILED);
            }
          } else {
            rc = launch_gather(false, src, rows, S->ks, S->K, (S->K + 15) / 16 * 16, nullptr, R.q_full[p], R.ctx->sm_count, R.gather);
          }
          CHPIR_CUDA(cudaEventRecord(R.gathered[p], R.gather), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
        }
      }
      for (uint32_t d = 0; d < S->n && rc == CHPIR_OK; d++) {
        Rank &R = S->r[d];
        chpir_server *sv = R.srv;
        CHPIR_CUDA(cudaSetDevice(R.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
        cudaStream_t st = R.compute;
        CHPIR_CUDA(cudaMemsetAsync(R.resp[p], 0, size_t(rows) * R.pl.nc * 4, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
        if (tc) {
          std::lock_guard<std::mutex> gg(sv->gemm_mu);
          const int buf = int(R.tc_buf++ & 1);
          if ((rc = gemm_tc_buf_acquire(sv->gemm, buf, st)) != CHPIR_OK) break;
          uint8_t *planes = gemm_tc_ring(sv->gemm) + uint64_t(buf) * gemm_tc_panel_bytes(sv->gemm);
          if ((rc = launch_gather(true, src, rows, S->ks, S->K, gemm_tc_kp(sv->gemm), planes, nullptr, R.ctx->sm_count, st)) != CHPIR_OK) break;
          if ((rc = gemm_tc_panel(sv->gemm, buf, rows, R.resp[p], st)) != CHPIR_OK) break;
          if ((rc = gemm_tc_buf_release(sv->gemm, buf, st)) != CHPIR_OK) break;
        } else if (direct) {
          if ((rc = launch_respond(sv->d_packed, sv->layout, sv->K, sv->plan, src.p[0], R.resp[p], rows, st)) != CHPIR_OK) break;
        } else {
          CHPIR_CUDA(cudaStreamWaitEvent(st, R.gathered[p], 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
          if ((rc = launch_respond(sv->d_packed, sv->layout, sv->K, sv->plan, R.q_full[p], R.resp[p], rows, st)) != CHPIR_OK) break;
        }
        // the rank's columns go straight into rank 0's row-major nq x N result (strided peer copy, copy engine)
        CHPIR_CUDA(cudaMemcpy2DAsync(resp_device0 + size_t(row0) * S->N + R.pl.c0, size_t(S->N) * 4, R.resp[p], size_t(R.pl.nc) * 4, size_t(R.pl.nc) * 4, rows,
                                     cudaMemcpyDefault, st),
                   CHPIR_ERR_CUDA_TRANSFER_FAILED);
        CHPIR_CUDA(cudaEventRecord(R.computed[p], st), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
      }
    }
  }
  // rank 0 waits for the last columns of every rank, then t1
  if (rc == CHPIR_OK) {
    cudaSetDevice(R0.dev);
    for (uint32_t d = 1; d < S->n; d++)
      for (int p = 0; p < 2; p++) cudaStreamWaitEvent(R0.compute, S->r[d].computed[p], 0);
    cudaEventRecord(S->t1, R0.compute);
  }
  for (uint32_t d = 0; d < S->n; d++) {
    cudaSetDevice(S->r[d].dev);
    cudaError_t e = cudaStreamSynchronize(S->r[d].gather);
    if (e == cudaSuccess) e = cudaStreamSynchronize(S->r[d].compute);
    if (e != cudaSuccess && rc == CHPIR_OK) {
      set_last_cuda_error(e, "cluster respond (device-resident)");
      rc = CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
    }
  }
  cudaSetDevice(R0.dev);
  if (rc == CHPIR_OK && device_ms) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, S->t0, S->t1) == cudaSuccess) *device_ms = ms;
  }
  return rc;
  CHPIR_GUARD_END
}

int chpir_cluster_server_respond_concurrent(chpir_cluster_server *S, const uint8_t *const *queries, const size_t *query_lens, uint32_t n_distinct,
                                            uint64_t total_calls, uint8_t *resp_out, size_t resp_stride, uint32_t n_threads, double *seconds) {
  CHPIR_GUARD_BEGIN
  if (!S || !queries || !query_lens || !resp_out || n_distinct == 0 || n_threads == 0) return CHPIR_ERR_INVALID_ARGUMENT;
  if (seconds) *seconds = 0.0;
  if (total_calls == 0) return CHPIR_OK;
  n_threads = uint32_t(std::min<uint64_t>(n_threads, total_calls));
  std::vector<int> rcs(n_threads, CHPIR_OK);
  std::vector<std::thread> th;
  th.reserve(n_threads);
  const double t0 = now_s();
  for (uint32_t t = 0; t < n_threads; t++)
    th.emplace_back([&, t] {
      for (uint64_t j = t; j < total_calls; j += n_threads) {
        const uint32_t i = uint32_t(j % n_distinct);
        size_t len = 0;
        const int rc = chpir_cluster_server_respond(S, queries[i], query_lens[i], resp_out + size_t(i) * resp_stride, resp_stride, &len);
        if (rc != CHPIR_OK) {
          rcs[t] = rc;
          return;
        }
      }
    });
  for (auto &t : th) t.join();
  if (seconds) *seconds = now_s() - t0;
  for (int rc : rcs)
    if (rc != CHPIR_OK) return rc;
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_cluster_server_get_info(const chpir_cluster_server *S, chpir_cluster_server_info *out) {
  if (!S || !out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = chpir_cluster_server_info{};
  out->n_gpus = S->n, out->cols_n = S->N, out->rows_k = S->K, out->mat_elem_bit_len = S->b, out->lwe_rows = S->lwe, out->k_pitch = S->ks;
  for (uint32_t d = 0; d < S->n; d++) {
    out->packed_bytes_total += S->r[d].srv->packed_bytes;
    out->packed_bytes_max_rank = std::max(out->packed_bytes_max_rank, S->r[d].srv->packed_bytes);
  }
  out->setup_total_s = S->setup_total_s, out->hint_gather_s = S->hint_gather_s;
  out->gather_uses_nccl = S->gather_uses_nccl;
  out->nccl_version = uint32_t(S->cl->nccl_version);
  out->respond_by_rows = S->rows ? 1u : 0u;
  out->reshard_s = S->reshard_s;
  out->pulled_queries = S->pulled;
  out->ingest_wait_s = S->ingest_wait_s, out->ingest_s = S->ingest_s, out->exec_wait_s = S->exec_wait_s, out->exec_s = S->exec_s;
  out->batches = S->batches, out->queries = S->queries, out->tc_batches = S->tc_batches;
  return CHPIR_OK;
}

}  // extern "C"
