// host_pipe.cuh -- small RAII helpers shared by the API layer and the host-pipelined producer of the LWE matrix A (HostAPipe).
#pragma once
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

#include "common.cuh"
#include "host_xof.hpp"

namespace chpir {

inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct DevBuf {
  void *p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int alloc(size_t bytes) {
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
      set_last_cuda_error(e, "cudaMalloc");
      p = nullptr;
      return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
    }
    return CHPIR_OK;
  }
  template <class T>
  T *as() const {
    return static_cast<T *>(p);
  }
  void *release() {
    void *q = p;
    p = nullptr;
    return q;
  }
};

struct EventTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  EventTimer() {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
  }
  ~EventTimer() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
  void start(cudaStream_t s) { cudaEventRecord(a, s); }
  void stop(cudaStream_t s) { cudaEventRecord(b, s); }
  float ms() {
    float v = 0.f;
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&v, a, b);
    return v;
  }
};


// ---- host-pipelined A (chpir_setup_opts.a_expand = CHPIR_A_EXPAND_HOST_PIPELINED) ----------------------------------------
// The XOF squeeze is one serial chain; a CPU core walks it several times faster than a GPU warp.  A producer thread squeezes
// the stream (= the row-major u32 bytes of A, matrix.rs:552-555) into a small ring of pinned chunks of whole rows; this thread
// uploads every chunk into a 128-row u32 staging panel and, once a panel is complete, splits it into byte planes and runs the
// tensor-core GEMM for it -- all on one stream, so uploads and GEMMs hide entirely behind the producer.
struct HostXofStream {
  HostXof x;
  uint8_t carry[kXofRate];
  uint32_t carry_off = 0, carry_len = 0;
  int impl = 0;
  void fill(uint8_t *dst, uint64_t n) {
    const uint64_t take = std::min<uint64_t>(carry_len, n);
    std::memcpy(dst, carry + carry_off, take);
    carry_off += uint32_t(take), carry_len -= uint32_t(take);
    dst += take, n -= take;
    const uint64_t blocks = n / kXofRate;
    host_xof_squeeze_blocks(&x, dst, blocks, impl);
    dst += blocks * kXofRate, n -= blocks * kXofRate;
    if (n) {
      host_xof_squeeze_blocks(&x, carry, 1, impl);
      std::memcpy(dst, carry, n);
      carry_off = uint32_t(n), carry_len = uint32_t(kXofRate - n);
    }
  }
};

// Producer (XOF) + uploader threads feeding a ring of `depth` 128-row u32 panel buffers in HBM.  The consumer (setup_core)
// takes panels in order: acquire_panel -> split into byte planes -> release_panel -> GEMM.  With depth = all panels the
// pipeline never waits for the consumer, which lets Server::setup(seed, db) start it BEFORE the host filter/encode phase.
class HostAPipe {
 public:
  HostAPipe() = default;
  HostAPipe(const HostAPipe &) = delete;
  HostAPipe &operator=(const HostAPipe &) = delete;
  ~HostAPipe() {
    shutdown();
    free_chunks();
    for (auto e : copied_)
      if (e) cudaEventDestroy(e);
    for (auto e : ready_)
      if (e) cudaEventDestroy(e);
    for (auto e : consumed_)
      if (e) cudaEventDestroy(e);
    if (up_) cudaStreamDestroy(up_);
  }

  // A follower of another pipe's chain (n GPUs of one process need the same A, and the chain is serial): no producer, no host
  // buffers -- the leader's forwarder thread sends every finished panel from its own HBM over NVLink (copy engines) into this pipe's
  // ring and signals it exactly as an uploader of its own would.  The leader touches a mirror's ring only after mirrors_ready();
  // a mirror must outlive the leader's threads (shut the leader down first).
  int start_mirror(int device, uint32_t m, uint64_t K, uint32_t depth) {
    device_ = device, m_ = m, K_ = K, mirror_ = true;
    row_bytes_ = K * 4;
    panels_ = (m + 127) / 128;
    panel_bytes_ = uint64_t(std::min(m, 128u)) * row_bytes_;
    depth_ = std::max(1u, std::min(depth, panels_));
    while (depth_ > 2 && uint64_t(depth_) * panel_bytes_ > (48ull << 30)) depth_--;
    if (cudaSetDevice(device) != cudaSuccess) return CHPIR_ERR_CUDA_DEVICE_NOT_FOUND;
    if (int rc = panels_dev_.alloc(uint64_t(depth_) * panel_bytes_); rc != CHPIR_OK) return rc;
    ready_.assign(panels_, nullptr);
    consumed_.assign(panels_, nullptr);
    for (uint32_t p = 0; p < panels_; p++)
      if (cudaEventCreateWithFlags(&ready_[p], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&consumed_[p], cudaEventDisableTiming) != cudaSuccess)
        return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
    if (cudaStreamCreateWithFlags(&up_, cudaStreamNonBlocking) != cudaSuccess) return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
    return CHPIR_OK;
  }
  bool is_mirror() const { return mirror_; }
  // Leader: the mirrors handed to start() have finished start_mirror() (their rings are allocated beside the running chain).
  void mirrors_ready() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      mirrors_ready_ = true;
    }
    cv_.notify_all();
  }

  int start(int device, const uint8_t seed[32], uint32_t m, uint64_t K, uint32_t chunk_rows_opt, uint32_t depth,
            std::vector<HostAPipe *> mirrors = {}) {
    mirrors_ = std::move(mirrors);
    device_ = device, m_ = m, K_ = K;
    std::memcpy(seed_, seed, 32);
    row_bytes_ = K * 4;
    panels_ = (m + 127) / 128;
    uint32_t chunk_rows = chunk_rows_opt ? chunk_rows_opt : uint32_t(std::max<uint64_t>(1, (32ull << 20) / row_bytes_));
    chunk_rows = std::min(chunk_rows, 128u);
    // chunks never straddle a panel: (first row, row count) in production order
    for (uint32_t p0 = 0; p0 < m; p0 += 128)
      for (uint32_t r = p0; r < std::min(m, p0 + 128); r += chunk_rows) chunks_.push_back({r, std::min(chunk_rows, std::min(m, p0 + 128) - r)});
    panel_bytes_ = uint64_t(std::min(m, 128u)) * row_bytes_;
    depth_ = std::max(1u, std::min(depth, panels_));
    while (depth_ > 2 && uint64_t(depth_) * panel_bytes_ > (48ull << 30)) depth_--;
    if (cudaSetDevice(device) != cudaSuccess) return CHPIR_ERR_CUDA_DEVICE_NOT_FOUND;
    // The chain is the critical path of the whole setup, so it starts the moment ONE pinned chunk exists; the other chunks, the
    // events, the panel ring in HBM (8.4 GB at 2^20 entries: a few tenths of a second of cudaMalloc) and the upload stream are
    // created while the producer is already squeezing: kBufs chunks of ~32 MB are ~0.3 s of chain to fill before it needs the
    // uploader to have started.
    const uint64_t chunk_bytes = uint64_t(chunk_rows) * row_bytes_;
    // The chunks are PAGEABLE by default: page-locking them cost 50-140 ms per 32 MB chunk on the measured box whenever the host was
    // busy touching memory (the filter/encode phase next door), which stalled the chain for up to 2 s in its first lap, while the
    // rate needed (8.4 GB in 5.8 s) is a fraction of what a staged pageable upload sustains.  CHPIR_XOF_PINNED=1 page-locks them.
    const char *pe = std::getenv("CHPIR_XOF_PINNED");
    pinned_alloc_ = pe && *pe && *pe != '0';
    auto pin = [&](int i) {
      if (pinned_alloc_) {
        if (cudaMallocHost(&pinned_[i], chunk_bytes) != cudaSuccess) {
          set_last_cuda_error(cudaGetLastError(), "pinned XOF chunk ring");
          pinned_[i] = nullptr;
          return false;
        }
      } else {
        void *m = nullptr;
        if (posix_memalign(&m, 4096, chunk_bytes ? chunk_bytes : 1) != 0) return false;
        std::memset(m, 0, chunk_bytes);  // first touch here, not in the producer
        pinned_[i] = static_cast<uint8_t *>(m);
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        pinned_ready_ = i + 1;
      }
      cv_.notify_all();
      return true;
    };
    if (!pin(0)) return CHPIR_ERR_HOST_ALLOCATION_FAILED;
    producer_ = std::thread([this] { produce(); });
    for (int i = 1; i < kBufs; i++)  // ~10 ms each; the producer needs ~20 ms per chunk, so it never catches up with this loop
      if (!pin(i)) return CHPIR_ERR_HOST_ALLOCATION_FAILED;  // (on any failure below the destructor stops the threads)
    for (int i = 0; i < kBufs; i++)
      if (cudaEventCreateWithFlags(&copied_[i], cudaEventDisableTiming) != cudaSuccess) return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
    if (int rc = panels_dev_.alloc(uint64_t(depth_) * panel_bytes_); rc != CHPIR_OK) return rc;
    ready_.assign(panels_, nullptr);
    consumed_.assign(panels_, nullptr);
    for (uint32_t p = 0; p < panels_; p++)
      if (cudaEventCreateWithFlags(&ready_[p], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&consumed_[p], cudaEventDisableTiming) != cudaSuccess)
        return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
    if (cudaStreamCreateWithFlags(&up_, cudaStreamNonBlocking) != cudaSuccess) return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
    uploader_ = std::thread([this] { upload(); });  // the first chunks are already waiting for it
    if (!mirrors_.empty()) forwarder_ = std::thread([this] { forward_loop(); });
    return CHPIR_OK;
  }

  // Blocks until panel p is completely uploaded, makes `st` wait for that upload, returns the panel's u32 rows.
  int acquire_panel(uint32_t p, cudaStream_t st, const uint32_t **rows) {
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return uploaded_panels_ > p || rc_ != CHPIR_OK; });
      if (rc_ != CHPIR_OK) return rc_;
    }
    CHPIR_CUDA(cudaStreamWaitEvent(st, ready_[p], 0), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
    *rows = reinterpret_cast<const uint32_t *>(panels_dev_.as<uint8_t>() + uint64_t(p % depth_) * panel_bytes_);
    return CHPIR_OK;
  }
  // Call once everything that reads panel p has been enqueued on `st`.
  int release_panel(uint32_t p, cudaStream_t st) {
    CHPIR_CUDA(cudaEventRecord(consumed_[p], st), CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
    {
      std::lock_guard<std::mutex> lk(mu_);
      released_panels_ = p + 1;
    }
    cv_.notify_all();
    return CHPIR_OK;
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      abort_ = true;
    }
    cv_.notify_all();
    if (producer_.joinable()) producer_.join();
    if (uploader_.joinable()) uploader_.join();
    if (forwarder_.joinable()) forwarder_.join();
    if (up_) cudaStreamSynchronize(up_);
    free_chunks();  // the chunk ring is only needed while the chain runs (a client keeps the pipe alive as the owner of A)
  }
  double busy_s() const { return busy_; }            // time the producer core spent inside the XOF
  double wait_s() const { return wait_; }            // time the producer core waited for a free pinned chunk (the chain stood still)
  uint32_t panels() const { return panels_; }
  uint32_t depth() const { return depth_; }
  // the panel ring; with depth() == panels() this is A itself, row-major u32
  const uint32_t *base() const { return panels_dev_.as<uint32_t>(); }
  // hands the ring over to the caller (chpir_setup_opts.a_cache); only meaningful after shutdown() and with depth() == panels()
  uint32_t *take_panels() { return static_cast<uint32_t *>(panels_dev_.release()); }

 private:
  static constexpr int kBufs = 16;

  void free_chunks() {
    for (auto &p : pinned_)
      if (p) {
        if (pinned_alloc_)
          cudaFreeHost(p);
        else
          std::free(p);
        p = nullptr;
      }
  }

  void fail(int rc) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (rc_ == CHPIR_OK) rc_ = rc;
      abort_ = true;
    }
    cv_.notify_all();
    fail_mirrors(rc);
  }

  // Leader's forwarder: panel p is complete in this pipe's ring (ready_[p] recorded on up_) -- send it on to every mirror.
  bool forward_panel(uint32_t p) {
    const uint32_t rows = std::min(m_, (p + 1) * 128) - p * 128;
    const uint8_t *src = panels_dev_.as<uint8_t>() + uint64_t(p % depth_) * panel_bytes_;
    bool ok = true;
    for (HostAPipe *m : mirrors_) {
      if (p >= m->depth_) {  // the ring slot's previous tenant must have been read by the mirror's consumer
        std::unique_lock<std::mutex> lk(m->mu_);
        // (not broken by this pipe's own shutdown(): the leader's consumer is done as soon as IT has the last panel, the mirrors
        // still need theirs; shutdown() joins this thread)
        m->cv_.wait(lk, [&] { return m->released_panels_ > p - m->depth_ || m->abort_ || m->rc_ != CHPIR_OK; });
        if (m->abort_ || m->rc_ != CHPIR_OK) continue;  // that rank gave up: nobody reads its ring any more
      }
      uint8_t *dst = m->panels_dev_.as<uint8_t>() + uint64_t(p % m->depth_) * m->panel_bytes_;
      ok = ok && cudaSetDevice(m->device_) == cudaSuccess;
      if (p >= m->depth_) ok = ok && cudaStreamWaitEvent(m->up_, m->consumed_[p - m->depth_], 0) == cudaSuccess;
      ok = ok && cudaStreamWaitEvent(m->up_, ready_[p], 0) == cudaSuccess &&
           cudaMemcpyPeerAsync(dst, m->device_, src, device_, uint64_t(rows) * row_bytes_, m->up_) == cudaSuccess &&
           cudaEventRecord(m->ready_[p], m->up_) == cudaSuccess;
      if (!ok) break;
      {
        std::lock_guard<std::mutex> lk(m->mu_);
        m->uploaded_panels_ = p + 1;
      }
      m->cv_.notify_all();
    }
    cudaSetDevice(device_);
    return ok;
  }
  // Leader's forwarder thread: panels go on to the mirrors in order, as soon as they are complete here; the uploader never waits for
  // a mirror except to reuse a ring slot.  Keeps going after shutdown() until everything uploaded has been forwarded (the leader's
  // consumer is done as soon as IT has the last panel; shutdown() joins this thread).
  void forward_loop() {
    cudaSetDevice(device_);
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return mirrors_ready_ || abort_; });
      if (!mirrors_ready_) return;
    }
    for (uint32_t p = 0; p < panels_; p++) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return uploaded_panels_ > p || abort_; });
        if (uploaded_panels_ <= p) return;  // stopped early; upload() tells the mirrors
      }
      if (!forward_panel(p)) {
        set_last_cuda_error(cudaGetLastError(), "XOF panel forward over NVLink");
        return fail(CHPIR_ERR_CUDA_TRANSFER_FAILED);
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        forwarded_panels_ = p + 1;
      }
      cv_.notify_all();
    }
  }

  void fail_mirrors(int rc) {
    for (HostAPipe *m : mirrors_) {  // their consumers wait for panels that will never come
      {
        std::lock_guard<std::mutex> lk(m->mu_);
        if (m->rc_ == CHPIR_OK) m->rc_ = rc;
      }
      m->cv_.notify_all();
    }
  }

  void produce() {
    cudaSetDevice(device_);
    HostXofStream xs;
    host_xof_init(&xs.x, seed_);
    for (uint64_t i = 0; i < chunks_.size(); i++) {
      const int b = int(i % kBufs);
      if (i < uint64_t(kBufs)) {  // first lap: the chunk may still be on its way through the allocation loop in start()
        const double w0 = now_s();
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return pinned_ready_ > b || abort_; });
        if (abort_) return;
        wait_ += now_s() - w0;
      }
      if (i >= uint64_t(kBufs)) {
        const double w0 = now_s();
        {
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [&] { return issued_ > i - kBufs || abort_; });
          if (abort_) return;
        }
        cudaEventSynchronize(copied_[b]);  // the upload of chunk i - kBufs has left this buffer
        wait_ += now_s() - w0;
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (abort_) return;
      }
      const double t0 = now_s();
      xs.fill(pinned_[b], uint64_t(chunks_[i].second) * row_bytes_);
      busy_ += now_s() - t0;
      {
        std::lock_guard<std::mutex> lk(mu_);
        filled_ = i + 1;
      }
      cv_.notify_all();
    }
  }

  void upload() {
    upload_chunks();
    if (mirrors_.empty()) return;
    bool complete;
    int rc;
    {
      std::lock_guard<std::mutex> lk(mu_);
      complete = issued_ == chunks_.size();
      rc = rc_;
    }
    if (!complete) fail_mirrors(rc != CHPIR_OK ? rc : CHPIR_ERR_CUDA_TRANSFER_FAILED);  // the chain was stopped early: no more panels
  }

  void upload_chunks() {
    cudaSetDevice(device_);
    for (uint64_t i = 0; i < chunks_.size(); i++) {
      const uint32_t r0 = chunks_[i].first, nr = chunks_[i].second, p = r0 / 128, panel_end = std::min(m_, (p + 1) * 128);
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return filled_ > i || abort_; });
        if (abort_) return;
        if (r0 == p * 128 && p >= depth_) {  // first chunk of a panel that reuses a ring slot: its previous tenant must have been read
          cv_.wait(lk, [&] { return released_panels_ > p - depth_ || abort_; });
          if (abort_) return;
        }
      }
      if (r0 == p * 128 && p >= depth_ && cudaStreamWaitEvent(up_, consumed_[p - depth_], 0) != cudaSuccess) return fail(CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
      if (r0 == p * 128 && p >= depth_ && !mirrors_.empty()) {  // ... and copied out by every mirror (their ready event of that panel)
        {
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [&] { return forwarded_panels_ > p - depth_ || abort_; });
          if (forwarded_panels_ <= p - depth_) return;
        }
        for (HostAPipe *m : mirrors_)
          if (cudaStreamWaitEvent(up_, m->ready_[p - depth_], 0) != cudaSuccess) return fail(CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED);
      }
      const int b = int(i % kBufs);
      uint8_t *dst = panels_dev_.as<uint8_t>() + uint64_t(p % depth_) * panel_bytes_ + uint64_t(r0 - p * 128) * row_bytes_;
      if (cudaMemcpyAsync(dst, pinned_[b], uint64_t(nr) * row_bytes_, cudaMemcpyHostToDevice, up_) != cudaSuccess ||
          cudaEventRecord(copied_[b], up_) != cudaSuccess) {
        set_last_cuda_error(cudaGetLastError(), "XOF chunk upload");
        return fail(CHPIR_ERR_CUDA_TRANSFER_FAILED);
      }
      const bool last = r0 + nr == panel_end;
      if (last && cudaEventRecord(ready_[p], up_) != cudaSuccess) return fail(CHPIR_ERR_CUDA_TRANSFER_FAILED);
      {
        std::lock_guard<std::mutex> lk(mu_);
        issued_ = i + 1;
        if (last) uploaded_panels_ = p + 1;
      }
      cv_.notify_all();
    }
  }

  bool mirror_ = false, mirrors_ready_ = false;
  std::vector<HostAPipe *> mirrors_;                       // leader only
  uint32_t forwarded_panels_ = 0;                          // leader only: panels handed to every mirror's stream
  int device_ = 0;
  uint32_t m_ = 0, panels_ = 0, depth_ = 0;
  uint64_t K_ = 0, row_bytes_ = 0, panel_bytes_ = 0;
  uint8_t seed_[32] = {};
  std::vector<std::pair<uint32_t, uint32_t>> chunks_;
  uint8_t *pinned_[kBufs] = {};
  cudaEvent_t copied_[kBufs] = {};
  std::vector<cudaEvent_t> ready_, consumed_;
  DevBuf panels_dev_;
  cudaStream_t up_ = nullptr;
  std::mutex mu_;
  std::condition_variable cv_;
  uint64_t filled_ = 0, issued_ = 0;
  int pinned_ready_ = 0;
  bool pinned_alloc_ = false;
  uint32_t uploaded_panels_ = 0, released_panels_ = 0;
  bool abort_ = false;
  int rc_ = CHPIR_OK;
  std::thread producer_, uploader_, forwarder_;
  double busy_ = 0.0, wait_ = 0.0;
};

}  // namespace chpir
