// cluster.cu -- the column-sharded server on 1..8 GPUs of ONE process, behind the reference's two calls.
//
// The reference's public surface is `Server::setup(seed, db)` (chalametpir_server/src/server.rs:103) and `Server::respond(&self, query)`
// (server.rs:184); neither can grow a rank argument, so the sharding of north_star (3) lives behind the handle: this translation
// unit owns one chpir_ctx per GPU, the peer mappings between them, the per-GPU streams and the NCCL communicators, and exposes
// chpir_cluster_server_{setup*,respond*} (include/chalamet_b200.h).  Column slices need no cross-rank arithmetic
// (M[:, n0:n1] = A.D[:, n0:n1], resp[n0:n1] = q.D[:, n0:n1]); what has to move is
//   * the query: every rank needs all K words of it.  Rank r ingests words [k0_r, k0_r + kn_r) over ITS OWN PCIe link (8 links in
//     parallel) and the other ranks read them from its HBM over NVLink -- inside the limb-split kernel that builds the tensor-core
//     operand (gather_q_kernel<true>: peer loads fused with the split) or, for the streaming GEMV, with copy-engine peer copies that
//     run beside the previous chunk's kernels and use no SM;
//   * the response: every rank writes its columns straight into the caller-visible row (strided D2H / peer copy), so there is no
//     gather step at all;
//   * the hint, once per setup: the slices are gathered on rank 0 over NCCL (ncclSend/ncclRecv, communicators from ncclCommInitAll)
//     and downloaded as one wire-format matrix.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <memory>
#include <string>
#include <thread>

#include "common.cuh"
#include "host_encode.hpp"
#include "host_pipe.cuh"
#include "server_state.cuh"

namespace chpir {
namespace {

constexpr uint32_t kMaxRanks = 16;
constexpr uint32_t kMaxBatch = Coalescer::kMaxBatch;       // queries per coalesced batch = one M tile of the limb GEMM
constexpr uint32_t kTcFrom = Coalescer::kTensorCoreFrom;   // below this many queries the streaming GEMV is cheaper
constexpr uint32_t kGemvChunk = 32;                        // device-resident GEMV path: queries per all-gather / launch

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
  chpir_server *srv = nullptr;
  Plan pl{};
  cudaStream_t compute = nullptr, gather = nullptr;
  // device-resident path (chpir_cluster_server_respond_device): double-buffered scratch, allocated on first use
  uint32_t *q_full[2] = {nullptr, nullptr};
  uint32_t *resp[2] = {nullptr, nullptr};
  uint32_t q_rows = 0, resp_rows = 0;
  cudaEvent_t gathered[2] = {nullptr, nullptr}, computed[2] = {nullptr, nullptr};
  uint32_t tc_buf = 0;  // next buffer of this rank's GEMM operand ring
};

// One coalescing slot: the members' slices on every rank, every rank's response columns, and the pinned rows they land in.
struct CBatch {
  uint32_t *q_slice[kMaxRanks] = {};  // [kMaxBatch][ks] on rank d
  uint32_t *q_full[kMaxRanks] = {};   // [gemv_rows][K] on rank d: whole queries for the GEMV route
  uint32_t *resp[kMaxRanks] = {};     // [kMaxBatch][nc_d] on rank d
  cudaStream_t copy[kMaxRanks] = {};
  cudaEvent_t uploaded[kMaxRanks] = {}, done[kMaxRanks] = {};
  uint32_t *h_resp = nullptr;  // pinned [kMaxBatch][N]
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
  bool tc = false;          // every rank keeps its limb planes: batches of >= kTcFrom queries take the tensor-core route
  uint32_t gemv_rows = 0;   // capacity of CBatch::q_full
  // coalescer (n > 1; a one-GPU cluster delegates to the shard's own)
  std::mutex mu;
  std::condition_variable cv;
  std::mutex exec_mu;  // one batch on the GPUs at a time
  CBatch cb[2];
  int open = 0;
  bool co_ready = false;
  uint64_t batches = 0, queries = 0, tc_batches = 0;
  // chpir_cluster_server_respond_batch
  std::mutex batch_mu;
  std::unique_ptr<CBatch> bb;
  // device-resident path
  std::mutex dev_mu;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  double setup_total_s = 0, hint_gather_s = 0;
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
    if (B.h_resp) cudaFreeHost(B.h_resp);
    B = CBatch{};
  }

  int init_batch(CBatch &B) {
    for (uint32_t d = 0; d < n; d++) {
      if (cudaSetDevice(r[d].dev) != cudaSuccess) return CHPIR_ERR_CUDA_DEVICE_NOT_FOUND;
      if (cudaMalloc(&B.q_slice[d], size_t(kMaxBatch) * ks * 4) != cudaSuccess ||
          cudaMalloc(&B.q_full[d], size_t(gemv_rows) * K * 4) != cudaSuccess ||
          cudaMalloc(&B.resp[d], size_t(kMaxBatch) * r[d].pl.nc * 4) != cudaSuccess ||
          cudaStreamCreateWithFlags(&B.copy[d], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&B.uploaded[d], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&B.done[d], cudaEventDisableTiming) != cudaSuccess) {
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
      }
      if (r[d].compute) cudaStreamDestroy(r[d].compute);
      if (r[d].gather) cudaStreamDestroy(r[d].gather);
      if (r[d].srv) chpir_server_destroy(r[d].srv);
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

// Streams, events and (for n > 1) the coalescing slots, once every shard exists.
int finish_server(chpir_cluster_server *S) {
  S->tc = true;
  for (uint32_t d = 0; d < S->n; d++) S->tc = S->tc && S->r[d].srv->gemm != nullptr;
  S->gemv_rows = S->tc ? kTcFrom - 1 : kMaxBatch;
  for (uint32_t d = 0; d < S->n; d++) {
    Rank &R = S->r[d];
    CHPIR_CUDA(cudaSetDevice(R.dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    CHPIR_CUDA(cudaStreamCreateWithFlags(&R.compute, cudaStreamNonBlocking), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
    CHPIR_CUDA(cudaStreamCreateWithFlags(&R.gather, cudaStreamNonBlocking), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
    for (int p = 0; p < 2; p++) {
      CHPIR_CUDA(cudaEventCreateWithFlags(&R.gathered[p], cudaEventDisableTiming), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
      CHPIR_CUDA(cudaEventCreateWithFlags(&R.computed[p], cudaEventDisableTiming), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
    }
  }
  CHPIR_CUDA(cudaSetDevice(S->r[0].dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  CHPIR_CUDA(cudaEventCreate(&S->t0), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
  CHPIR_CUDA(cudaEventCreate(&S->t1), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
  if (S->n > 1) {
    for (CBatch &B : S->cb)
      if (int rc = S->init_batch(B); rc != CHPIR_OK) return rc;
    S->co_ready = true;
  }
  return CHPIR_OK;
}

// The GPU half of one batch of `nq` queries whose slices sit in B.q_slice: every rank pulls the slices it did not ingest, answers
// for its columns and writes them into the pinned rows; returns when all ranks have.
int run_batch(chpir_cluster_server *S, CBatch &B, uint32_t nq, bool *used_tc) {
  const bool tc = S->tc && nq >= kTcFrom;
  if (used_tc) *used_tc = tc;
  if (!tc && nq > S->gemv_rows) return CHPIR_ERR_INVALID_ARGUMENT;
  SrcTable src{};
  for (uint32_t d = 0; d < S->n; d++) src.p[d] = B.q_slice[d];
  for (uint32_t d = 0; d < S->n; d++) {
    CHPIR_CUDA(cudaSetDevice(S->r[d].dev), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    CHPIR_CUDA(cudaEventRecord(B.uploaded[d], B.copy[d]), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
  }
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
  // wait for every rank that was given work, also on the error path: the slot must be quiet before it is reused
  for (uint32_t d = 0; d < S->n; d++) {
    cudaSetDevice(S->r[d].dev);
    const cudaError_t e = cudaStreamSynchronize(S->r[d].compute);
    if (e != cudaSuccess && rc == CHPIR_OK) {
      set_last_cuda_error(e, "cluster respond");
      rc = CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
    }
  }
  return rc;
}

// A member's (or the batch call's) upload of one query: its K/n words go to each rank over that rank's own PCIe link.
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

// One caller's share of a coalesced batch: the protocol of api.cu's respond_coalesced with n GPUs behind it.
int respond_coalesced(chpir_cluster_server *S, const uint8_t *query, uint8_t *resp_out) {
  CBatch *B = nullptr;
  uint32_t row = 0;
  {
    std::unique_lock<std::mutex> lk(S->mu);
    S->cv.wait(lk, [&] { return !S->cb[S->open].closed && S->cb[S->open].count < kMaxBatch; });
    B = &S->cb[S->open];
    row = B->count++;
  }
  const bool leader = row == 0;
  const int up = upload_query(S, *B, row, query);
  {
    std::lock_guard<std::mutex> lk(S->mu);
    B->issued++;
    if (up != CHPIR_OK && B->rc == CHPIR_OK) B->rc = up;
  }
  S->cv.notify_all();
  if (leader) {
    std::lock_guard<std::mutex> ex(S->exec_mu);  // the previous batch leaves the GPUs: everyone who arrived meanwhile is in this one
    uint32_t nq = 0;
    int rc = CHPIR_OK;
    {
      std::unique_lock<std::mutex> lk(S->mu);
      B->closed = true;
      nq = B->count;
      S->cv.wait(lk, [&] { return B->issued == nq; });
      rc = B->rc;
      S->cv.wait(lk, [&] { return S->cb[S->open ^ 1].count == 0 && !S->cb[S->open ^ 1].closed; });
      S->open ^= 1;
    }
    S->cv.notify_all();
    bool tc = false;
    if (rc == CHPIR_OK) {
      rc = run_batch(S, *B, nq, &tc);
    } else {
      for (uint32_t d = 0; d < S->n; d++) {
        cudaSetDevice(S->r[d].dev);
        cudaStreamSynchronize(B->copy[d]);
      }
    }
    {
      std::lock_guard<std::mutex> lk(S->mu);
      S->batches++, S->queries += nq, S->tc_batches += tc ? 1 : 0;
      B->rc = rc;
      B->finished = true;
    }
    S->cv.notify_all();
  }
  int rc = CHPIR_OK;
  {
    std::unique_lock<std::mutex> lk(S->mu);
    S->cv.wait(lk, [&] { return B->finished; });
    rc = B->rc;
  }
  if (rc == CHPIR_OK) {
    const uint32_t hdr[2] = {1u, S->N};
    std::memcpy(resp_out, hdr, 8);
    std::memcpy(resp_out + 8, B->h_resp + size_t(row) * S->N, size_t(S->N) * 4);
  }
  {
    std::lock_guard<std::mutex> lk(S->mu);
    if (++B->picked == B->count) {  // last one out resets the slot
      B->count = B->issued = B->picked = 0;
      B->closed = B->finished = false;
      B->rc = CHPIR_OK;
    }
  }
  S->cv.notify_all();
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
  if (cl->n > 1 && o->db_encode == CHPIR_DB_ENCODE_DEVICE) return CHPIR_ERR_INVALID_ARGUMENT;
  return CHPIR_OK;
}

// Common tail of every setup flavour: per-rank servers exist in S->r[*].srv.
int complete_setup(chpir_cluster_server *S, const chpir_setup_opts &o, uint8_t *hint_out, size_t hint_cap, size_t *hint_len) {
  for (uint32_t d = 0; d < S->n; d++) S->r[d].srv->col_begin = S->r[d].pl.c0;  // logical position of a compact slice
  if (int rc = finish_server(S); rc != CHPIR_OK) return rc;
  if (hint_len) *hint_len = 0;
  if (!o.skip_hint) {
    if (int rc = gather_hint(S, hint_out, hint_cap, hint_len); rc != CHPIR_OK) return rc;
  }
  return CHPIR_OK;
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
  return S;
}

// per-rank options: the rank's columns, hint slice kept in HBM for the gather, no per-shard coalescer (the cluster has its own)
chpir_setup_opts rank_opts(const chpir_setup_opts &o, const Plan &pl, bool compact, uint32_t n) {
  chpir_setup_opts ro = o;
  ro.col_begin = compact ? 0 : pl.c0;
  ro.col_count = compact ? 0 : pl.nc;
  ro.hint_on_device = 1;
  if (n > 1) ro.respond_coalesce = 0;
  return ro;
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
  int rc = for_each_rank_parallel(S->n, [&](uint32_t d) -> int {
    if (!d_slices[d]) return CHPIR_ERR_INVALID_ARGUMENT;
    const chpir_setup_opts ro = rank_opts(o, S->r[d].pl, true, S->n);
    return chpir_server_setup_device(S->r[d].ctx, seed, d_slices[d], rows_k, S->r[d].pl.nc, b, &ro, nullptr, 0, nullptr, &S->r[d].srv);
  });
  if (rc != CHPIR_OK) return rc;
  if ((rc = complete_setup(S.get(), o, hint_out, hint_cap, hint_len)) != CHPIR_OK) return rc;
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
  int rc = for_each_rank_parallel(S->n, [&](uint32_t d) -> int {
    const chpir_setup_opts ro = rank_opts(o, S->r[d].pl, false, S->n);
    return server_setup_from_host_matrix(S->r[d].ctx, seed, d_host, rows_k, cols_n, b, &ro, nullptr, 0, nullptr, &S->r[d].srv, nullptr);
  });
  if (rc != CHPIR_OK) return rc;
  if ((rc = complete_setup(S.get(), o, hint_out, hint_cap, hint_len)) != CHPIR_OK) return rc;
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
  if (S->n == 1) {
    // one GPU: the single-GPU call as it is (device row fill, its own early XOF start, its coalescer), hint slice = whole hint
    chpir_setup_opts ro = o;
    ro.hint_on_device = 1;
    if (int rc = chpir_server_setup_from_db(S->r[0].ctx, arity, seed, n, key_blob, key_offsets, value_blob, value_offsets, filter_seed_rng, &ro, nullptr, 0,
                                            nullptr, filter_params_out, &S->r[0].srv);
        rc != CHPIR_OK)
      return rc;
  } else {
    // Every rank needs all of A = generate_from_seed(lwe, K, seed), and the XOF chain is serial: each rank's host pipeline starts
    // NOW (its own producer core, ring as deep as A) and squeezes beside the host filter/encode phase below, exactly as the
    // single-GPU call does; D is encoded once on the host and every rank uploads its own columns.
    const bool host_a = o.a_expand != CHPIR_A_EXPAND_DEVICE && !o.skip_hint && o.gemm_variant == 0;
    std::vector<std::unique_ptr<HostAPipe>> pipes(S->n);
    if (host_a) {
      for (uint32_t d = 0; d < S->n; d++) {
        bool cached = false;
        if (o.a_cache) {
          std::lock_guard<std::mutex> g(S->r[d].ctx->mu);
          cached = S->r[d].ctx->a_cache.matches(seed, S->lwe, K);
        }
        if (cached) continue;
        pipes[d].reset(new HostAPipe());
        if (int rc = pipes[d]->start(S->r[d].dev, seed, S->lwe, K, o.host_chunk_rows, (S->lwe + 127) / 128); rc != CHPIR_OK) return rc;
      }
    }
    std::unique_ptr<uint32_t[]> d_store(new (std::nothrow) uint32_t[K * N]);
    if (!d_store) return CHPIR_ERR_HOST_ALLOCATION_FAILED;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (host_a) set_encode_threads(hw > S->n + 2 ? hw - S->n - 1 : 1);  // the producer cores stay free for the chains
    int rc = encode_kv_database(arity, n, key_blob, key_offsets, value_blob, value_offsets, b, CHPIR_SERVER_SETUP_MAX_ATTEMPT_COUNT, filter_seed_rng,
                                d_store.get(), filter_params_out);
    set_encode_threads(0);
    if (rc != CHPIR_OK) return rc;
    const double t1 = now_s();
    rc = for_each_rank_parallel(S->n, [&](uint32_t d) -> int {
      const chpir_setup_opts ro = rank_opts(o, S->r[d].pl, false, S->n);
      return server_setup_from_host_matrix(S->r[d].ctx, seed, d_store.get(), K, uint32_t(N), b, &ro, nullptr, 0, nullptr, &S->r[d].srv, pipes[d].get());
    });
    if (rc != CHPIR_OK) return rc;
    for (uint32_t d = 0; d < S->n; d++) S->r[d].srv->timing.host_encode_s = t1 - t0;
  }
  if (int rc = complete_setup(S.get(), o, hint_out, hint_cap, hint_len); rc != CHPIR_OK) return rc;
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
  if (n > 1) ro.respond_coalesce = 0;
  int rc = for_each_rank_parallel(n, [&](uint32_t d) -> int {
    const std::string p = std::string(path_prefix) + ".rank" + std::to_string(d) + "of" + std::to_string(n);
    return chpir_server_load(cl->ctx[d], p.c_str(), &ro, &shards[d]);
  });
  if (rc != CHPIR_OK) return rc;
  // the files must describe ONE matrix cut by this cluster's plan
  uint64_t ncols = 0;
  for (uint32_t d = 0; d < n; d++) ncols += shards[d]->ncols;
  if (ncols > 0xffffffffull) return CHPIR_ERR_INVALID_SAVED_SERVER;
  for (uint32_t d = 0; d < n; d++) {
    const Plan pl = plan_of(n, d, shards[0]->K, uint32_t(ncols));
    if (shards[d]->K != shards[0]->K || shards[d]->b != shards[0]->b || shards[d]->ncols != pl.nc || shards[d]->col_begin != pl.c0)
      return CHPIR_ERR_INVALID_SAVED_SERVER;
  }
  std::unique_ptr<chpir_cluster_server> S(new_server(cl, shards[0]->K, uint32_t(ncols), shards[0]->b, o));
  S->lwe = 0;  // no hint travels with a saved server
  for (uint32_t d = 0; d < n; d++) {
    S->r[d].srv = shards[d];
    shards[d] = nullptr;
  }
  if ((rc = finish_server(S.get())) != CHPIR_OK) return rc;
  *out = S.release();
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_cluster_server_respond(chpir_cluster_server *S, const uint8_t *query, size_t query_len, uint8_t *resp_out, size_t resp_cap, size_t *resp_len) {
  CHPIR_GUARD_BEGIN
  DeviceRestore restore_device;
  if (!S) return CHPIR_ERR_INVALID_ARGUMENT;
  if (S->n == 1) return chpir_server_respond(S->r[0].srv, query, query_len, resp_out, resp_cap, resp_len);
  if (int rc = validate_query_bytes(S->K, query, query_len); rc != CHPIR_OK) return rc;
  const size_t need = 8 + size_t(S->N) * 4;
  if (!resp_out || resp_cap < need) return CHPIR_ERR_BUFFER_TOO_SMALL;
  if (!S->co_ready) return CHPIR_ERR_INVALID_ARGUMENT;
  const int rc = respond_coalesced(S, query, resp_out);
  if (rc == CHPIR_OK && resp_len) *resp_len = need;
  return rc;
  CHPIR_GUARD_END
}

int chpir_cluster_server_respond_batch(chpir_cluster_server *S, const uint8_t *const *queries, const size_t *query_lens, uint32_t nq, uint8_t *resp_out,
                                       size_t resp_stride) {
  CHPIR_GUARD_BEGIN
  DeviceRestore restore_device;
  if (!S || !queries || !query_lens || !resp_out) return CHPIR_ERR_INVALID_ARGUMENT;
  if (S->n == 1) return chpir_server_respond_batch(S->r[0].srv, queries, query_lens, nq, resp_out, resp_stride);
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
    for (uint32_t i = 0; i < cnt; i++)
      if (int rc = upload_query(S, B, i, queries[q0 + i]); rc != CHPIR_OK) return rc;
    int rc;
    {
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
  if (S->n == 1) {
    const Coalescer &co = S->r[0].srv->co;
    out->batches = co.batches, out->queries = co.queries, out->tc_batches = co.tc_batches;
  } else {
    out->batches = S->batches, out->queries = S->queries, out->tc_batches = S->tc_batches;
  }
  return CHPIR_OK;
}

}  // extern "C"
