// staged_upload.cuh -- a large PAGEABLE host buffer into HBM with several host threads.
//
// cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread: one core copying into one bounce buffer,
// 2.6-3.5 GB/s on the measured box while the filter construction keeps the other cores busy.  The values of a database (1.07 GB at
// 2^20 x 1 kB, the input of the device row fill, encode_dev.cu) are the largest thing `Server::setup` moves when A is cached, and
// page-locking the caller's buffer costs more than the copy.  Here T helper threads each copy 4 MB chunks into two page-locked
// bounce buffers of their own and send them on with cudaMemcpyAsync on their own stream; the bounce buffers belong to the ctx and
// are allocated once (page-locking under a busy host costs tens of milliseconds per buffer).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace chpir {

struct StagePool {  // owned by a chpir_ctx, used under its setup mutex
  static constexpr size_t kChunk = 4ull << 20;
  static constexpr int kThreads = 4, kBufs = 2 * kThreads;
  uint8_t *buf[kBufs] = {};
  cudaEvent_t sent[kBufs] = {};  // the DMA out of the bounce buffer has finished
  cudaEvent_t done[kThreads] = {};
  cudaStream_t stream[kThreads] = {};
  bool ready = false;

  bool init() {  // device already current
    if (ready) return true;
    for (int i = 0; i < kBufs; i++)
      if (cudaMallocHost(&buf[i], kChunk) != cudaSuccess || cudaEventCreateWithFlags(&sent[i], cudaEventDisableTiming) != cudaSuccess) return fail();
    for (int t = 0; t < kThreads; t++)
      if (cudaStreamCreateWithFlags(&stream[t], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&done[t], cudaEventDisableTiming) != cudaSuccess)
        return fail();
    return ready = true;
  }
  bool fail() {
    (void)cudaGetLastError();
    release();
    return false;
  }
  void release() {
    for (int i = 0; i < kBufs; i++) {
      if (buf[i]) cudaFreeHost(buf[i]);
      if (sent[i]) cudaEventDestroy(sent[i]);
      buf[i] = nullptr, sent[i] = nullptr;
    }
    for (int t = 0; t < kThreads; t++) {
      if (stream[t]) cudaStreamDestroy(stream[t]);
      if (done[t]) cudaEventDestroy(done[t]);
      stream[t] = nullptr, done[t] = nullptr;
    }
    ready = false;
  }
};

class StagedUpload {
 public:
  // Starts copying src[0, bytes) to dst (device memory of `device`) and returns at once.  Falls back to one pageable cudaMemcpyAsync
  // on `consumer` when the pool cannot be set up.  `src` must stay valid until finish() has returned.
  bool start(StagePool *pool, int device, void *dst, const void *src, size_t bytes, cudaStream_t consumer) {
    pool_ = pool;
    size_t min_bytes = 8 * StagePool::kChunk;  // below this the driver's own staging is as good
    if (const char *v = std::getenv("CHPIR_STAGE_MIN_BYTES"); v && *v) min_bytes = size_t(std::strtoull(v, nullptr, 10));
    if (bytes == 0 || bytes < min_bytes || !pool->init()) {
      pool_ = nullptr;
      return bytes == 0 || cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, consumer) == cudaSuccess;
    }
    const size_t chunks = (bytes + StagePool::kChunk - 1) / StagePool::kChunk;
    for (int t = 0; t < StagePool::kThreads; t++) {
      ok_[t] = true;
      workers_.emplace_back([=] {
        if (cudaSetDevice(device) != cudaSuccess) {
          ok_[t] = false;
          return;
        }
        size_t turn = 0;
        for (size_t c = size_t(t); c < chunks; c += StagePool::kThreads, turn++) {
          const int b = 2 * t + int(turn & 1);
          const size_t off = c * StagePool::kChunk, len = bytes - off < StagePool::kChunk ? bytes - off : StagePool::kChunk;
          if (cudaEventSynchronize(pool->sent[b]) != cudaSuccess) ok_[t] = false;  // the DMA that last read this buffer (none: returns at once)
          std::memcpy(pool->buf[b], static_cast<const uint8_t *>(src) + off, len);
          if (cudaMemcpyAsync(static_cast<uint8_t *>(dst) + off, pool->buf[b], len, cudaMemcpyHostToDevice, pool->stream[t]) != cudaSuccess ||
              cudaEventRecord(pool->sent[b], pool->stream[t]) != cudaSuccess)
            ok_[t] = false;
        }
        if (cudaEventRecord(pool->done[t], pool->stream[t]) != cudaSuccess) ok_[t] = false;
      });
    }
    return true;
  }
  // Joins the helpers and makes `consumer` wait for their copies; false if any of them failed.
  bool finish(cudaStream_t consumer) {
    bool ok = true;
    for (auto &w : workers_) w.join();
    const bool had_workers = !workers_.empty();
    workers_.clear();
    if (!pool_ || !had_workers) return true;
    for (int t = 0; t < StagePool::kThreads; t++) {
      ok = ok && ok_[t];
      if (cudaStreamWaitEvent(consumer, pool_->done[t], 0) != cudaSuccess) ok = false;
    }
    if (!ok) (void)cudaGetLastError();
    return ok;
  }
  ~StagedUpload() {
    for (auto &w : workers_) w.join();
    // an abandoned upload (error path of the caller): the bounce buffers must be idle before anybody reuses or frees them
    if (pool_ && pool_->ready)
      for (int t = 0; t < StagePool::kThreads; t++) cudaStreamSynchronize(pool_->stream[t]);
  }

 private:
  StagePool *pool_ = nullptr;
  std::vector<std::thread> workers_;
  bool ok_[StagePool::kThreads] = {};
};

}  // namespace chpir
