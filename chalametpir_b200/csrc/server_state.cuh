// server_state.cuh -- the state behind the opaque chpir_server handle, shared by api.cu (single GPU) and cluster.cu (1-8 GPUs).
#pragma once
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"

namespace chpir {

// One in-flight respond: its own stream and buffers, so concurrent callers never share state.
struct RespondSlot {
  cudaStream_t stream = nullptr;
  uint32_t *d_q = nullptr;
  uint32_t *d_resp = nullptr;
  uint32_t *h_resp = nullptr;  // pinned
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};

// Transparent coalescing of concurrent chpir_server_respond calls (chpir_setup_opts.respond_coalesce).  Callers join the open
// batch, each uploading its own query straight into its row of the batch's device buffer; the first caller of a batch is its
// leader: it waits for the previous batch to leave the GPU (that wait IS the batching window -- an idle GPU means a batch of
// one and no added latency), closes the batch, runs one launch for all of it (grid.y GEMV for a handful of queries, the
// tensor-core limb GEMM beyond that, where one pass over D serves up to 128 queries) and reads the responses back.  Two batches
// ping-pong, so the uploads of the next batch overlap the compute of the current one.
struct CoalesceBatch {
  uint32_t *d_q = nullptr, *d_resp = nullptr, *h_resp = nullptr;
  cudaStream_t copy = nullptr;
  cudaEvent_t uploaded = nullptr;
  uint32_t count = 0, issued = 0, picked = 0;
  bool closed = false, done = false;
  int rc = CHPIR_OK;
};

struct Coalescer {
  static constexpr uint32_t kMaxBatch = 128;
  static constexpr uint32_t kTensorCoreFrom = 6;  // a tensor-core pass costs about as much as five streaming GEMVs
  std::mutex mu;
  std::condition_variable cv;
  std::mutex exec_mu;  // one batch on the GPU at a time
  CoalesceBatch b[2];
  int open = 0;
  cudaStream_t compute = nullptr;
  bool ready = false;
  uint64_t batches = 0, queries = 0, tc_batches = 0;  // statistics
};

class HostAPipe;
// api.cu internals the cluster layer builds on
int server_setup_from_host_matrix(chpir_ctx *ctx, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint32_t *d_host, uint64_t rows_k, uint32_t cols_n,
                                  uint32_t b, const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap, size_t *hint_len, chpir_server **out,
                                  HostAPipe *pipe);
int server_setup_from_device_matrix(chpir_ctx *ctx, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint32_t *d_device, uint64_t rows_k, uint32_t cols_n,
                                    uint32_t b, const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap, size_t *hint_len, chpir_server **out,
                                    HostAPipe *pipe);
// registry of the page-locked ranges chpir_host_alloc handed out (device-readable from every GPU of the process)
void pinned_registry_add(const void *p, size_t bytes);
void pinned_registry_remove(const void *p);
bool pinned_registry_contains(const void *p, size_t bytes);
// Matrix::from_bytes validation + the 1 x K dimension check, in the reference's order (matrix.rs:973-1010, :329-331)
int validate_query_bytes(uint64_t K, const uint8_t *query, size_t len);

}  // namespace chpir

using chpir::CoalesceBatch;
using chpir::Coalescer;
using chpir::RespondSlot;

struct chpir_server {
  chpir_ctx *ctx = nullptr;
  uint64_t K = 0;
  uint32_t ncols = 0, col_begin = 0, b = 0;
  uint64_t shard_k_total = 0;  // != 0: this server is a ROW block of a cluster's D (csrc/cluster.cu); the rows of the whole matrix
  chpir::PackedLayout layout{};
  chpir::RespondPlan plan{};
  uint8_t *d_packed = nullptr;
  uint64_t packed_bytes = 0;
  uint32_t *d_hint = nullptr;  // chpir_setup_opts.hint_on_device: the hint slice (hint_rows x ncols u32) left in HBM
  uint32_t hint_rows = 0;
  chpir::GemmTcB *gemm = nullptr;  // D's byte-limb planes + operand ring, kept for the tensor-core batched respond
  std::mutex gemm_mu;
  bool coalesce = false;
  chpir::Coalescer co;
  chpir_setup_timing timing{};
  float last_respond_ms = 0.f, last_gemm_ms = 0.f, last_expand_ms = 0.f;
  std::mutex pool_mu;
  std::vector<RespondSlot *> free_slots;
  std::vector<RespondSlot *> all_slots;
  // chpir_server_respond_batch: one batch in flight per server, buffers grown on demand
  std::mutex batch_mu;
  cudaStream_t batch_stream = nullptr;
  uint32_t *batch_q = nullptr, *batch_resp = nullptr, *batch_h_resp = nullptr;
  uint32_t batch_cap = 0;

  ~chpir_server() {
    if (ctx) cudaSetDevice(ctx->device);
    for (RespondSlot *s : all_slots) {
      if (s->stream) cudaStreamSynchronize(s->stream);
      if (s->d_q) cudaFree(s->d_q);
      if (s->d_resp) cudaFree(s->d_resp);
      if (s->h_resp) cudaFreeHost(s->h_resp);
      if (s->e0) cudaEventDestroy(s->e0);
      if (s->e1) cudaEventDestroy(s->e1);
      if (s->stream) cudaStreamDestroy(s->stream);
      delete s;
    }
    if (batch_stream) {
      cudaStreamSynchronize(batch_stream);
      cudaStreamDestroy(batch_stream);
    }
    if (batch_q) cudaFree(batch_q);
    if (batch_resp) cudaFree(batch_resp);
    if (batch_h_resp) cudaFreeHost(batch_h_resp);
    if (d_packed) cudaFree(d_packed);
    if (d_hint) cudaFree(d_hint);
    if (gemm) gemm_tc_free(gemm);
    for (CoalesceBatch &cb : co.b) {
      if (cb.copy) {
        cudaStreamSynchronize(cb.copy);
        cudaStreamDestroy(cb.copy);
      }
      if (cb.uploaded) cudaEventDestroy(cb.uploaded);
      if (cb.d_q) cudaFree(cb.d_q);
      if (cb.d_resp) cudaFree(cb.d_resp);
      if (cb.h_resp) cudaFreeHost(cb.h_resp);
    }
    if (co.compute) {
      cudaStreamSynchronize(co.compute);
      cudaStreamDestroy(co.compute);
    }
  }

  int init_coalescer() {
    for (CoalesceBatch &cb : co.b) {
      if (cudaMalloc(&cb.d_q, size_t(Coalescer::kMaxBatch) * K * 4) != cudaSuccess ||
          cudaMalloc(&cb.d_resp, size_t(Coalescer::kMaxBatch) * ncols * 4) != cudaSuccess ||
          cudaMallocHost(&cb.h_resp, size_t(Coalescer::kMaxBatch) * ncols * 4) != cudaSuccess ||
          cudaStreamCreateWithFlags(&cb.copy, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&cb.uploaded, cudaEventDisableTiming) != cudaSuccess) {
        chpir::set_last_cuda_error(cudaGetLastError(), "respond coalescer allocation");
        return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
      }
    }
    if (cudaStreamCreateWithFlags(&co.compute, cudaStreamNonBlocking) != cudaSuccess) return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
    co.ready = true;
    return CHPIR_OK;
  }

  int reserve_batch(uint32_t nq) {
    if (!batch_stream && cudaStreamCreateWithFlags(&batch_stream, cudaStreamNonBlocking) != cudaSuccess) return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
    if (nq <= batch_cap) return CHPIR_OK;
    if (batch_q) cudaFree(batch_q);
    if (batch_resp) cudaFree(batch_resp);
    if (batch_h_resp) cudaFreeHost(batch_h_resp);
    batch_q = batch_resp = batch_h_resp = nullptr;
    batch_cap = 0;
    if (cudaMalloc(&batch_q, size_t(nq) * K * 4) != cudaSuccess || cudaMalloc(&batch_resp, size_t(nq) * ncols * 4) != cudaSuccess ||
        cudaMallocHost(&batch_h_resp, size_t(nq) * ncols * 4) != cudaSuccess) {
      chpir::set_last_cuda_error(cudaGetLastError(), "respond batch allocation");
      return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
    }
    batch_cap = nq;
    return CHPIR_OK;
  }

  int acquire(RespondSlot **out) {
    {
      std::lock_guard<std::mutex> g(pool_mu);
      if (!free_slots.empty()) {
        *out = free_slots.back();
        free_slots.pop_back();
        return CHPIR_OK;
      }
    }
    RespondSlot *s = new (std::nothrow) RespondSlot();
    if (!s) return CHPIR_ERR_HOST_ALLOCATION_FAILED;
    bool ok = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&s->d_q, K * 4) == cudaSuccess && cudaMalloc(&s->d_resp, size_t(ncols) * 4) == cudaSuccess &&
              cudaMallocHost(&s->h_resp, size_t(ncols) * 4) == cudaSuccess && cudaEventCreate(&s->e0) == cudaSuccess &&
              cudaEventCreate(&s->e1) == cudaSuccess;
    {
      std::lock_guard<std::mutex> g(pool_mu);
      all_slots.push_back(s);
    }
    if (!ok) {
      chpir::set_last_cuda_error(cudaGetLastError(), "respond slot allocation");
      return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
    }
    *out = s;
    return CHPIR_OK;
  }
  void release(RespondSlot *s) {
    std::lock_guard<std::mutex> g(pool_mu);
    free_slots.push_back(s);
  }
};

