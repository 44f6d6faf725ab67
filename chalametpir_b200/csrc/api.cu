// api.cu -- the extern "C" surface declared in include/chalamet_b200.h.
// No exceptions cross this boundary: every entry point is wrapped in a catch-all and returns a chpir_status.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>

#include <unistd.h>

#include "common.cuh"
#include "device_fill.cuh"
#include "host_encode.hpp"
#include "host_pipe.cuh"
#include "host_xof.hpp"
#include "server_state.cuh"

namespace chpir {

static thread_local std::string g_last_cuda_error;
void set_last_cuda_error(cudaError_t e, const char *what) {
  g_last_cuda_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " at " + what;
  (void)cudaGetLastError();  // clear the sticky-less error state
}

// Page-locked ranges handed out by chpir_host_alloc: memory every GPU of the process can read directly (portable + mapped), which
// the cluster's query ingest uses to pull a whole batch with one kernel per GPU instead of one DMA call per query and GPU.
namespace {
std::mutex g_pinned_mu;
std::vector<std::pair<uintptr_t, uintptr_t>> g_pinned;  // [begin, end), few entries
}  // namespace
void pinned_registry_add(const void *p, size_t bytes) {
  std::lock_guard<std::mutex> g(g_pinned_mu);
  g_pinned.emplace_back(reinterpret_cast<uintptr_t>(p), reinterpret_cast<uintptr_t>(p) + bytes);
}
void pinned_registry_remove(const void *p) {
  std::lock_guard<std::mutex> g(g_pinned_mu);
  for (size_t i = 0; i < g_pinned.size(); i++)
    if (g_pinned[i].first == reinterpret_cast<uintptr_t>(p)) {
      g_pinned.erase(g_pinned.begin() + long(i));
      return;
    }
}
bool pinned_registry_contains(const void *p, size_t bytes) {
  const uintptr_t b = reinterpret_cast<uintptr_t>(p), e = b + bytes;
  std::lock_guard<std::mutex> g(g_pinned_mu);
  for (const auto &r : g_pinned)
    if (b >= r.first && e <= r.second) return true;
  return false;
}

}  // namespace chpir

using namespace chpir;

namespace {

int validate_bits(uint32_t b) { return (b >= 4 && b <= 14) ? CHPIR_OK : CHPIR_ERR_IMPOSSIBLE_ENCODED_DB_MATRIX_ELEMENT_BIT_LENGTH; }

// Hint GEMM fed by a HostAPipe (started here with a two-panel ring unless the caller started one earlier).
// keep_a: the ring is made as deep as A and handed to the ctx afterwards (chpir_setup_opts.a_cache).
int hint_host_pipelined(chpir_ctx *ctx, const uint8_t *seed, const GemmTcB *g, uint32_t m, uint64_t K, uint32_t ncols, uint32_t *c_dev,
                        uint32_t chunk_rows_opt, HostAPipe *pipe, cudaStream_t st, float *gemm_ms_out, double *xof_busy_s, double *xof_wait_s, bool keep_a) {
  HostAPipe own;
  if (!pipe) {
    pipe = &own;
    if (int rc = own.start(ctx->device, seed, m, K, chunk_rows_opt, keep_a ? (m + 127) / 128 : 2); rc != CHPIR_OK) return rc;
  }
  const uint32_t panels = pipe->panels();
  struct Ev {
    std::vector<cudaEvent_t> ev;
    ~Ev() {
      for (auto e : ev)
        if (e) cudaEventDestroy(e);
    }
  } evs;
  evs.ev.resize(2 * panels, nullptr);
  for (auto &e : evs.ev)
    if (cudaEventCreate(&e) != cudaSuccess) return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
  int rc = CHPIR_OK;
  for (uint32_t p = 0; p < panels && rc == CHPIR_OK; p++) {
    const uint32_t rows = std::min(m, (p + 1) * 128) - p * 128;
    const uint32_t *a_rows = nullptr;
    if ((rc = pipe->acquire_panel(p, st, &a_rows)) != CHPIR_OK) break;
    if ((rc = gemm_tc_load_panel_u32(g, int(p & 1), a_rows, rows, st)) != CHPIR_OK) break;
    if ((rc = pipe->release_panel(p, st)) != CHPIR_OK) break;
    cudaEventRecord(evs.ev[2 * p], st);
    rc = gemm_tc_panel(g, int(p & 1), rows, c_dev + size_t(p) * 128u * ncols, st);
    cudaEventRecord(evs.ev[2 * p + 1], st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  pipe->shutdown();
  if (rc == CHPIR_OK && e != cudaSuccess) {
    set_last_cuda_error(e, "setup: host-pipelined A + hint GEMM");
    rc = CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
  }
  float gemm_ms = 0.f;
  if (rc == CHPIR_OK)
    for (uint32_t p = 0; p < panels; p++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, evs.ev[2 * p], evs.ev[2 * p + 1]) == cudaSuccess) gemm_ms += ms;
    }
  *gemm_ms_out = gemm_ms;
  *xof_busy_s = pipe->busy_s();
  *xof_wait_s = pipe->wait_s();
  if (rc == CHPIR_OK && keep_a && pipe->depth() == panels) {  // the ring holds every panel: it IS A, row-major u32
    if (ctx->a_cache.a) cudaFree(ctx->a_cache.a);
    ctx->a_cache.a = pipe->take_panels();
    ctx->a_cache.m = m, ctx->a_cache.K = K;
    std::memcpy(ctx->a_cache.seed, seed, 32);
  }
  return rc;
}

// Hint GEMM from the A a previous setup left resident in the ctx: per 128-row panel, split the u32 rows into byte planes and multiply.
int hint_from_cached_a(chpir_ctx *ctx, const GemmTcB *g, uint32_t m, uint64_t K, uint32_t ncols, uint32_t *c_dev, cudaStream_t st, float *gemm_ms_out) {
  const uint32_t panels = (m + 127) / 128;
  struct Ev {
    std::vector<cudaEvent_t> ev;
    ~Ev() {
      for (auto e : ev)
        if (e) cudaEventDestroy(e);
    }
  } evs;
  evs.ev.resize(2 * panels, nullptr);
  for (auto &e : evs.ev)
    if (cudaEventCreate(&e) != cudaSuccess) return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
  for (uint32_t p = 0; p < panels; p++) {
    const uint32_t rows = std::min(m, (p + 1) * 128) - p * 128;
    if (int rc = gemm_tc_load_panel_u32(g, int(p & 1), ctx->a_cache.a + uint64_t(p) * 128u * K, rows, st); rc != CHPIR_OK) return rc;
    cudaEventRecord(evs.ev[2 * p], st);
    if (int rc = gemm_tc_panel(g, int(p & 1), rows, c_dev + size_t(p) * 128u * ncols, st); rc != CHPIR_OK) return rc;
    cudaEventRecord(evs.ev[2 * p + 1], st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_last_cuda_error(e, "setup: hint GEMM from the cached A");
    return CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
  }
  float gemm_ms = 0.f;
  for (uint32_t p = 0; p < panels; p++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, evs.ev[2 * p], evs.ev[2 * p + 1]) == cudaSuccess) gemm_ms += ms;
  }
  *gemm_ms_out = gemm_ms;
  return CHPIR_OK;
}

// Core of setup once D (K x ld u32, device) is resident.  d_dev columns [col0, col0+ncols) form this server's slice.
int setup_core(chpir_ctx *ctx, const uint8_t *seed, const uint32_t *d_dev, uint64_t K, uint32_t ld, uint32_t col0, uint32_t ncols,
               uint32_t col_begin_logical, uint32_t b, const chpir_setup_opts &o, uint8_t *hint_out, size_t hint_cap, size_t *hint_len,
               chpir_server *srv, HostAPipe *pipe = nullptr) {
  cudaStream_t st = ctx->stream;
  srv->ctx = ctx;
  srv->K = K;
  srv->ncols = ncols;
  srv->col_begin = col_begin_logical;
  srv->b = b;
  srv->layout = make_layout(b, ncols);
  srv->packed_bytes = K * srv->layout.pitch_bytes();

  // -- respond layout
  EventTimer t_pack;
  {
    void *p = nullptr;
    CHPIR_CUDA(cudaMalloc(&p, srv->layout.alloc_bytes(K)), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
    srv->d_packed = static_cast<uint8_t *>(p);
    if (srv->layout.alloc_bytes(K) > srv->packed_bytes)  // the zeroed pad row behind tight rows (PackedLayout::alloc_bytes)
      CHPIR_CUDA(cudaMemsetAsync(srv->d_packed + srv->packed_bytes, 0, srv->layout.alloc_bytes(K) - srv->packed_bytes, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  }
  t_pack.start(st);
  if (int rc = launch_pack(d_dev, K, ld, col0, srv->layout, srv->d_packed, st); rc != CHPIR_OK) return rc;
  t_pack.stop(st);
  srv->plan = plan_respond(srv->layout, K, ctx->sm_count);
  double tt = trace_now();
  trace_phase("core: plan_respond", tt);

  const uint32_t m = o.lwe_rows ? o.lwe_rows : CHPIR_LWE_DIMENSION;
  if (!o.skip_hint) {
    const size_t need = 8 + size_t(m) * ncols * 4;
    if (!o.hint_on_device && (!hint_out || hint_cap < need)) return CHPIR_ERR_BUFFER_TOO_SMALL;
    DevBuf scratch, c;
    if (int rc = scratch.alloc(512); rc != CHPIR_OK) return rc;
    if (int rc = c.alloc(size_t(m) * ncols * 4); rc != CHPIR_OK) return rc;
    EventTimer t_all;
    float gemm_ms = 0.f, host_wall_ms = 0.f;
    if (o.gemm_variant == 1) {
      // debug path: A = generate_from_seed(m, K, seed) as u32 in HBM, then the SIMT u32 GEMM
      DevBuf a;
      if (int rc = a.alloc(size_t(m) * K * 4); rc != CHPIR_OK) return rc;
      EventTimer t_gemm;
      t_all.start(st);
      if (int rc = launch_expand(seed, a.as<uint8_t>(), uint64_t(m) * K * 4, scratch.as<uint8_t>(), st); rc != CHPIR_OK) return rc;
      t_gemm.start(st);
      if (int rc = launch_gemm_simt(a.as<uint32_t>(), d_dev + col0, ld, c.as<uint32_t>(), m, K, ncols, st); rc != CHPIR_OK) return rc;
      t_gemm.stop(st);
      t_all.stop(st);
      CHPIR_CUDA(cudaStreamSynchronize(st), CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED);
      gemm_ms = t_gemm.ms();
    } else {
      // production path: A is squeezed out of the XOF 128 rows at a time, directly as the byte planes the tensor-core
      // GEMM reads, into a two-buffer ring; each panel is multiplied as soon as it exists.  One stream: the XOF chain is
      // serial and ~1000x longer than a panel GEMM, so there is nothing to gain from overlapping them.
      GemmTcB *g = nullptr;
      trace_phase("core: scratch + hint alloc", tt);
      if (int rc = gemm_tc_prepare(d_dev + col0, ld, K, ncols, b, ctx->sm_count, st, &g); rc != CHPIR_OK) return rc;
      trace_phase("core: gemm_tc_prepare", tt);
      struct Guard {
        GemmTcB *g;
        std::vector<cudaEvent_t> ev;
        ~Guard() {
          for (cudaEvent_t e : ev) cudaEventDestroy(e);
          if (g) gemm_tc_free(g);
        }
      } guard{g, {}};
      const uint32_t panels = (m + 127) / 128;
      guard.ev.resize(2 * panels, nullptr);
      for (auto &e : guard.ev)
        if (cudaEventCreate(&e) != cudaSuccess) return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
      CHPIR_CUDA(cudaMemsetAsync(c.p, 0, size_t(m) * ncols * 4, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
      t_all.start(st);
      bool a_filled_now = false;
      if (o.a_cache && !ctx->a_cache.matches(seed, m, K) && o.a_expand == CHPIR_A_EXPAND_DEVICE) {
        a_filled_now = true;
        // device expansion with a_cache: squeeze A once as row-major u32 (what the cache holds), then take the cached route below
        if (ctx->a_cache.a) cudaFree(ctx->a_cache.a);
        ctx->a_cache = chpir_ctx::ACache{};
        DevBuf a;
        if (int rc = a.alloc(size_t(m) * K * 4); rc != CHPIR_OK) return rc;
        if (int rc = launch_expand(seed, a.as<uint8_t>(), uint64_t(m) * K * 4, scratch.as<uint8_t>(), st); rc != CHPIR_OK) return rc;
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
          set_last_cuda_error(e, "setup: expand A for the cache");
          return CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
        }
        ctx->a_cache.a = static_cast<uint32_t *>(a.release());
        ctx->a_cache.m = m, ctx->a_cache.K = K;
        std::memcpy(ctx->a_cache.seed, seed, 32);
      }
      if (o.a_cache && ctx->a_cache.matches(seed, m, K)) {
        if (pipe) pipe->shutdown();  // started by the caller before it could know: not needed
        if (int rc = hint_from_cached_a(ctx, g, m, K, ncols, c.as<uint32_t>(), st, &gemm_ms); rc != CHPIR_OK) return rc;
        t_all.stop(st);
        srv->timing.a_cache_hit = a_filled_now ? 0.0 : 1.0;
      } else if (o.a_expand != CHPIR_A_EXPAND_DEVICE) {
        const double w0 = now_s();
        if (int rc = hint_host_pipelined(ctx, seed, g, m, K, ncols, c.as<uint32_t>(), o.host_chunk_rows, pipe, st, &gemm_ms, &srv->timing.xof_host_busy_s,
                                         &srv->timing.xof_host_wait_s, o.a_cache != 0);
            rc != CHPIR_OK)
          return rc;
        host_wall_ms = float((now_s() - w0) * 1e3);
      } else {
        if (int rc = expand_begin(seed, scratch.as<uint8_t>(), st); rc != CHPIR_OK) return rc;
        const uint64_t total_blocks = (uint64_t(m) * K * 4 + 167) / 168;
        uint64_t blk = 0;
        for (uint32_t p = 0; p < panels; p++) {
          const uint32_t r1 = std::min<uint32_t>(m, (p + 1) * 128u);
          const uint64_t blk_end = std::min<uint64_t>(total_blocks, (uint64_t(r1) * K * 4 + 167) / 168);
          if (int rc = launch_expand_planes(gemm_tc_ring(g), m, K, gemm_tc_kp(g), scratch.as<uint8_t>(), blk, blk_end - blk, st); rc != CHPIR_OK)
            return rc;
          blk = blk_end;
          cudaEventRecord(guard.ev[2 * p], st);
          if (int rc = gemm_tc_panel(g, p & 1, r1 - p * 128u, c.as<uint32_t>() + size_t(p) * 128u * ncols, st); rc != CHPIR_OK) return rc;
          cudaEventRecord(guard.ev[2 * p + 1], st);
        }
        t_all.stop(st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
          set_last_cuda_error(e, "setup: expand + hint GEMM");
          return CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
        }
        for (uint32_t p = 0; p < panels; p++) {
          float ms = 0.f;
          if (cudaEventElapsedTime(&ms, guard.ev[2 * p], guard.ev[2 * p + 1]) == cudaSuccess) gemm_ms += ms;
        }
      }
      if (o.batch_tc != 2) {  // the planes exist: keep them for the batched respond
        srv->gemm = g;
        guard.g = nullptr;
      }
    }
    trace_phase("core: A + hint GEMM", tt);
    const float all_ms = host_wall_ms > 0.f ? host_wall_ms : t_all.ms();
    const double t0 = now_s();
    if (o.hint_on_device) {
      // the slice stays in HBM for whoever gathers the slices of all ranks (csrc/cluster.cu); no download here
      srv->d_hint = static_cast<uint32_t *>(c.release());
      srv->hint_rows = m;
    } else {
      const uint32_t hdr[2] = {m, ncols};
      std::memcpy(hint_out, hdr, 8);
      CHPIR_CUDA(cudaMemcpyAsync(hint_out + 8, c.p, size_t(m) * ncols * 4, cudaMemcpyDeviceToHost, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
    }
    CHPIR_CUDA(cudaStreamSynchronize(st), CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED);
    srv->timing.d2h_s = now_s() - t0;
    trace_phase("core: hint download", tt);
    srv->timing.gemm_s = gemm_ms * 1e-3;
    srv->timing.expand_a_s = (all_ms - gemm_ms) * 1e-3;
    srv->last_gemm_ms = gemm_ms;
    srv->last_expand_ms = all_ms - gemm_ms;
    if (hint_len) *hint_len = o.hint_on_device ? 0 : need;
  } else {
    CHPIR_CUDA(cudaStreamSynchronize(st), CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED);
    if (hint_len) *hint_len = 0;
  }
  if (o.batch_tc == 1 && !srv->gemm) {
    if (int rc = gemm_tc_prepare(d_dev + col0, ld, K, ncols, b, ctx->sm_count, st, &srv->gemm); rc != CHPIR_OK) return rc;
    CHPIR_CUDA(cudaStreamSynchronize(st), CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED);
  }
  srv->timing.pack_s = t_pack.ms() * 1e-3;
  if (o.respond_coalesce) {
    if (int rc = srv->init_coalescer(); rc != CHPIR_OK) return rc;
    srv->coalesce = true;
  }
  return CHPIR_OK;
}

// Matrix::from_kv_database::<ARITY> with the row fill on the GPU: host = key digests + filter construction + wave plan (once, whatever
// the number of GPUs), device = row encoding + dependent fill of a COLUMN RANGE of D (the recurrence couples rows, never columns, so
// every GPU of a cluster builds its own columns from the same plan).  d_out receives columns [c0, c0 + nc) as a K x nc u32 matrix.
}  // namespace

int DeviceFillHost::prepare(uint32_t arity, uint64_t n, const uint8_t *key_blob, const uint64_t *key_off, uint32_t b, uint32_t max_attempts,
                            const uint64_t *seed_rng) {
  if (int rc = digest_and_peel(arity, n, key_blob, key_off, b, max_attempts, seed_rng, &digests, &pr); rc != CHPIR_OK) return rc;
  double tt = trace_now();
  plan_fill_levels(arity, pr, &plan);
  trace_phase("fill wave plan", tt);
  pr.params.to_bytes(filter_bytes);
  return CHPIR_OK;
}

// Step 1 (before the host-side peeling, so that the memset and the values upload run under it): allocations, D cleared, values on
// their way through the ctx's bounce buffers.  Holds the ctx's setup mutex until finish() or destruction.
int DeviceFillRank::begin(chpir_ctx *ctx, uint64_t n, const uint8_t *val_blob, const uint64_t *val_off, uint64_t K, uint32_t nc, DevBuf *d_out) {
  ctx_ = ctx, n_ = n, K_ = K, nc_ = nc, d_out_ = d_out;
  lock_ = std::unique_lock<std::mutex>(ctx->mu);
  CHPIR_CUDA(cudaSetDevice(ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  cudaStream_t st = ctx->stream;
  double tt = trace_now();
  if (int rc = d_out->alloc(K * nc * 4); rc != CHPIR_OK) return rc;
  CHPIR_CUDA(cudaMemsetAsync(d_out->p, 0, K * nc * 4, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  const uint64_t val_bytes = val_off[n];
  if (int rc = d_values_.alloc(val_bytes); rc != CHPIR_OK) return rc;
  if (int rc = d_valoff_.alloc((n + 1) * 8); rc != CHPIR_OK) return rc;
  // scratch of the row fill (per-key records, wave table: at most one wave per key), allocated up front with the rest
  if (int rc = d_fill_rec_.alloc(n * kFillRecordBytes); rc != CHPIR_OK) return rc;
  if (int rc = d_fill_levels_.alloc((n + 2) * 4); rc != CHPIR_OK) return rc;
  if (int rc = d_members_.alloc(n * 4); rc != CHPIR_OK) return rc;
  if (int rc = d_order_.alloc(n * 8); rc != CHPIR_OK) return rc;
  if (int rc = d_found_.alloc(n); rc != CHPIR_OK) return rc;
  if (int rc = d_koo_.alloc(n * 4); rc != CHPIR_OK) return rc;
  if (int rc = d_digests_.alloc(n * 32); rc != CHPIR_OK) return rc;
  trace_phase("device allocations", tt);
  // values do not depend on the filter: their upload (pageable memory) also precedes the peeling
  CHPIR_CUDA(cudaMemcpyAsync(d_valoff_.p, val_off, (n + 1) * 8, cudaMemcpyHostToDevice, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  // (several helper threads with page-locked bounce buffers: the driver's own staging of a pageable copy is one core, ~3 GB/s here)
  if (!values_up_.start(&ctx->stage, ctx->device, d_values_.p, val_blob, val_bytes, st)) {
    set_last_cuda_error(cudaGetLastError(), "values upload");
    return CHPIR_ERR_CUDA_TRANSFER_FAILED;
  }
  trace_phase("values upload (enqueue)", tt);
  return CHPIR_OK;
}

// Step 2 (after DeviceFillHost::prepare): the plan goes up, the fill runs; the ctx's stream is synchronised on return.
int DeviceFillRank::finish(uint32_t arity, const DeviceFillHost &h, uint64_t N, uint32_t c0, uint32_t b, double *device_s) {
  struct Unlock {
    std::unique_lock<std::mutex> &l;
    ~Unlock() {
      if (l.owns_lock()) l.unlock();
    }
  } unlock{lock_};
  chpir_ctx *ctx = ctx_;
  CHPIR_CUDA(cudaSetDevice(ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  cudaStream_t st = ctx->stream;
  const uint64_t n = n_;
  const uint32_t waves = uint32_t(h.plan.level_start.size() - 1);
  double tt = trace_now();
  CHPIR_CUDA(cudaMemcpyAsync(d_members_.p, h.plan.members.data(), n * 4, cudaMemcpyHostToDevice, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  CHPIR_CUDA(cudaMemcpyAsync(d_order_.p, h.pr.order.data(), n * 8, cudaMemcpyHostToDevice, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  CHPIR_CUDA(cudaMemcpyAsync(d_found_.p, h.pr.found.data(), n, cudaMemcpyHostToDevice, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  CHPIR_CUDA(cudaMemcpyAsync(d_koo_.p, h.pr.key_of_order.data(), n * 4, cudaMemcpyHostToDevice, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  CHPIR_CUDA(cudaMemcpyAsync(d_digests_.p, h.digests.data(), n * 32, cudaMemcpyHostToDevice, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  trace_phase("plan upload (enqueue)", tt);
  if (!values_up_.finish(st)) {
    set_last_cuda_error(cudaGetLastError(), "values upload");
    return CHPIR_ERR_CUDA_TRANSFER_FAILED;
  }
  trace_phase("values upload (rest)", tt);
  EventTimer t_fill;
  t_fill.start(st);
  if (int rc = launch_device_row_fill(arity, d_members_.as<uint32_t>(), h.plan.level_start.data(), waves, d_order_.as<uint64_t>(),
                                      d_found_.as<uint8_t>(), d_koo_.as<uint32_t>(), d_digests_.as<uint8_t>(), d_values_.as<uint8_t>(),
                                      d_valoff_.as<uint64_t>(), d_out_->as<uint32_t>(), N, c0, nc_, b, h.pr.params.segment_length,
                                      h.pr.params.segment_count_length, d_fill_rec_.p, d_fill_levels_.as<uint32_t>(), st);
      rc != CHPIR_OK)
    return rc;
  t_fill.stop(st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_last_cuda_error(e, "device row fill");
    return CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
  }
  trace_phase("device row fill", tt);
  if (device_s) *device_s = t_fill.ms() * 1e-3;
  return CHPIR_OK;
}

namespace {

// One GPU, all columns (the caller holds no lock: DeviceFillRank takes the ctx's setup mutex and releases it before returning).
int encode_kv_database_device(chpir_ctx *ctx, uint32_t arity, uint64_t n, const uint8_t *key_blob, const uint64_t *key_off,
                              const uint8_t *val_blob, const uint64_t *val_off, uint32_t b, uint32_t max_attempts, const uint64_t *seed_rng,
                              uint64_t K, uint64_t N, DevBuf *d_out, uint8_t filter_bytes[68], double *host_s, double *device_s) {
  const double t0 = now_s();
  DeviceFillRank rank;
  if (int rc = rank.begin(ctx, n, val_blob, val_off, K, uint32_t(N), d_out); rc != CHPIR_OK) return rc;
  DeviceFillHost h;
  if (int rc = h.prepare(arity, n, key_blob, key_off, b, max_attempts, seed_rng); rc != CHPIR_OK) return rc;
  const double t1 = now_s();
  if (int rc = rank.finish(arity, h, N, 0, b, device_s); rc != CHPIR_OK) return rc;
  std::memcpy(filter_bytes, h.filter_bytes, 68);
  if (host_s) *host_s = t1 - t0;
  return CHPIR_OK;
}

int resolve_slice(const chpir_setup_opts &o, uint32_t N, uint32_t *c0, uint32_t *nc) {
  *c0 = o.col_begin;
  *nc = o.col_count ? o.col_count : (N > o.col_begin ? N - o.col_begin : 0);
  if (*nc == 0 || uint64_t(*c0) + *nc > N) return CHPIR_ERR_INVALID_ARGUMENT;
  return CHPIR_OK;
}

}  // namespace

#define CHPIR_GUARD_BEGIN try {
#define CHPIR_GUARD_END                          \
  }                                              \
  catch (const std::bad_alloc &) {               \
    return CHPIR_ERR_HOST_ALLOCATION_FAILED;     \
  }                                              \
  catch (...) {                                  \
    return CHPIR_ERR_INVALID_ARGUMENT;           \
  }

extern "C" {

const char *chpir_strerror(int status) {
  switch (status) {
    case CHPIR_OK: return "ok";
    case CHPIR_ERR_INVALID_MATRIX_DIMENSION: return "InvalidMatrixDimension";
    case CHPIR_ERR_INCOMPATIBLE_DIMENSION_FOR_MATRIX_MULTIPLICATION: return "IncompatibleDimensionForMatrixMultiplication";
    case CHPIR_ERR_INCOMPATIBLE_DIMENSION_FOR_ROW_VECTOR_TRANSPOSED_MATRIX_MULTIPLICATION:
      return "IncompatibleDimensionForRowVectorTransposedMatrixMultiplication";
    case CHPIR_ERR_FAILED_TO_DESERIALIZE_MATRIX_FROM_BYTES: return "FailedToDeserializeMatrixFromBytes";
    case CHPIR_ERR_EMPTY_KV_DATABASE: return "EmptyKVDatabase";
    case CHPIR_ERR_EXHAUSTED_ALL_ATTEMPTS_TO_BUILD_3_WISE_XOR_FILTER: return "ExhaustedAllAttemptsToBuild3WiseXorFilter";
    case CHPIR_ERR_EXHAUSTED_ALL_ATTEMPTS_TO_BUILD_4_WISE_XOR_FILTER: return "ExhaustedAllAttemptsToBuild4WiseXorFilter";
    case CHPIR_ERR_ROW_NOT_DECODABLE: return "RowNotDecodable";
    case CHPIR_ERR_DECODED_ROW_NOT_PREPENDED_WITH_DIGEST_OF_KEY: return "DecodedRowNotPrependedWithDigestOfKey";
    case CHPIR_ERR_FAILED_TO_DESERIALIZE_FILTER_FROM_BYTES: return "FailedToDeserializeFilterFromBytes";
    case CHPIR_ERR_KV_DATABASE_SIZE_TOO_LARGE: return "KVDatabaseSizeTooLarge";
    case CHPIR_ERR_INVALID_HINT_MATRIX: return "InvalidHintMatrix";
    case CHPIR_ERR_ARITHMETIC_OVERFLOW_ADDING_QUERY_INDICATOR: return "ArithmeticOverflowAddingQueryIndicator";
    case CHPIR_ERR_INVALID_RESPONSE_VECTOR: return "InvalidResponseVector";
    case CHPIR_ERR_PENDING_QUERY_EXISTS_FOR_KEY: return "PendingQueryExistsForKey";
    case CHPIR_ERR_PENDING_QUERY_DOES_NOT_EXIST_FOR_KEY: return "PendingQueryDoesNotExistForKey";
    case CHPIR_ERR_UNSUPPORTED_ARITY_FOR_BINARY_FUSE_FILTER: return "UnsupportedArityForBinaryFuseFilter";
    case CHPIR_ERR_IMPOSSIBLE_ENCODED_DB_MATRIX_ELEMENT_BIT_LENGTH: return "ImpossibleEncodedDBMatrixElementBitLength";
    case CHPIR_ERR_INVALID_ARGUMENT: return "InvalidArgument";
    case CHPIR_ERR_IO_FAILED: return "IoFailed";
    case CHPIR_ERR_INVALID_SAVED_SERVER: return "InvalidSavedServer";
    case CHPIR_ERR_BUFFER_TOO_SMALL: return "BufferTooSmall";
    case CHPIR_ERR_CUDA_DEVICE_NOT_FOUND: return "CudaDeviceNotFound";
    case CHPIR_ERR_CUDA_ALLOCATION_FAILED: return "CudaAllocationFailed";
    case CHPIR_ERR_CUDA_TRANSFER_FAILED: return "CudaTransferFailed";
    case CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED: return "CudaKernelLaunchFailed";
    case CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED: return "CudaKernelExecutionFailed";
    case CHPIR_ERR_CUDA_UNSUPPORTED_DEVICE: return "CudaUnsupportedDevice";
    case CHPIR_ERR_CUDA_PEER_ACCESS_UNAVAILABLE: return "CudaPeerAccessUnavailable";
    case CHPIR_ERR_NCCL_FAILED: return "NcclFailed";
    case CHPIR_ERR_HOST_ALLOCATION_FAILED: return "HostAllocationFailed";
    default: return "UnknownStatus";
  }
}

const char *chpir_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

int chpir_device_count(int *count) {
  if (!count) return CHPIR_ERR_INVALID_ARGUMENT;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_last_cuda_error(e, "cudaGetDeviceCount");
    *count = 0;
    return CHPIR_ERR_CUDA_DEVICE_NOT_FOUND;
  }
  *count = n;
  return CHPIR_OK;
}

int chpir_ctx_create(int device_ordinal, chpir_ctx **out) {
  CHPIR_GUARD_BEGIN
  if (!out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device_ordinal < 0 || device_ordinal >= n) {
    (void)cudaGetLastError();
    return CHPIR_ERR_CUDA_DEVICE_NOT_FOUND;
  }
  CHPIR_CUDA(cudaSetDevice(device_ordinal), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  cudaDeviceProp prop{};
  CHPIR_CUDA(cudaGetDeviceProperties(&prop, device_ordinal), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  // sm_100a-only binary: no fallback, fail loudly on anything else
  if (prop.major != 10) return CHPIR_ERR_CUDA_UNSUPPORTED_DEVICE;
  chpir_ctx *c = new chpir_ctx();
  c->device = device_ordinal;
  c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
  }
  *out = c;
  return CHPIR_OK;
  CHPIR_GUARD_END
}

void chpir_ctx_destroy(chpir_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
  }
  if (ctx->a_cache.a) cudaFree(ctx->a_cache.a);
  ctx->stage.release();
  delete ctx;
}

int chpir_ctx_drop_a_cache(chpir_ctx *ctx, uint64_t *bytes_freed) {
  CHPIR_GUARD_BEGIN
  if (!ctx) return CHPIR_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (bytes_freed) *bytes_freed = ctx->a_cache.a ? uint64_t(ctx->a_cache.m) * ctx->a_cache.K * 4 : 0;
  if (ctx->a_cache.a) {
    CHPIR_CUDA(cudaSetDevice(ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    cudaFree(ctx->a_cache.a);
  }
  ctx->a_cache = chpir_ctx::ACache{};
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_host_alloc(size_t bytes, void **out) {
  if (!out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
    set_last_cuda_error(cudaGetLastError(), "cudaHostAlloc");
    return CHPIR_ERR_HOST_ALLOCATION_FAILED;
  }
  pinned_registry_add(*out, bytes ? bytes : 1);
  return CHPIR_OK;
}

void chpir_host_free(void *p) {
  if (!p) return;
  pinned_registry_remove(p);
  cudaFreeHost(p);
}

int chpir_upload_rows(void *dst_device, size_t dst_pitch, const void *src_host, size_t src_pitch, size_t width_bytes, size_t rows, void *cuda_stream) {
  CHPIR_GUARD_BEGIN
  if (!dst_device || !src_host || width_bytes > dst_pitch || width_bytes > src_pitch) return CHPIR_ERR_INVALID_ARGUMENT;
  if (rows == 0 || width_bytes == 0) return CHPIR_OK;
  CHPIR_CUDA(cudaMemcpy2DAsync(dst_device, dst_pitch, src_host, src_pitch, width_bytes, rows, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(cuda_stream)),
             CHPIR_ERR_CUDA_TRANSFER_FAILED);
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_find_mat_elem_bit_len(uint64_t n, uint32_t *b) {
  if (!b) return CHPIR_ERR_INVALID_ARGUMENT;
  return find_mat_elem_bit_len(n, b);
}

int chpir_db_matrix_shape(uint32_t arity, uint64_t n, uint64_t max_value_byte_len, uint32_t b, uint64_t *rows_k, uint64_t *cols_n) {
  if (!rows_k || !cols_n) return CHPIR_ERR_INVALID_ARGUMENT;
  return db_matrix_shape(arity, n, max_value_byte_len, b, rows_k, cols_n);
}

int chpir_encode_kv_database(uint32_t arity, uint64_t n, const uint8_t *key_blob, const uint64_t *key_offsets, const uint8_t *value_blob,
                             const uint64_t *value_offsets, uint32_t b, uint32_t max_attempt_count, const uint64_t *filter_seed_rng,
                             uint32_t *d_out, uint8_t filter_params_out[CHPIR_FILTER_PARAM_BYTE_LEN]) {
  CHPIR_GUARD_BEGIN
  if (n == 0) return CHPIR_ERR_EMPTY_KV_DATABASE;
  if (!key_blob || !key_offsets || !value_blob || !value_offsets || !d_out || !filter_params_out) return CHPIR_ERR_INVALID_ARGUMENT;
  return encode_kv_database(arity, n, key_blob, key_offsets, value_blob, value_offsets, b, max_attempt_count, filter_seed_rng, d_out,
                            filter_params_out);
  CHPIR_GUARD_END
}

int chpir_encode_kv_database_device(chpir_ctx *ctx, uint32_t arity, uint64_t n, const uint8_t *key_blob, const uint64_t *key_offsets,
                                    const uint8_t *value_blob, const uint64_t *value_offsets, uint32_t b, uint32_t max_attempt_count,
                                    const uint64_t *filter_seed_rng, uint32_t *d_out, uint8_t filter_params_out[CHPIR_FILTER_PARAM_BYTE_LEN]) {
  CHPIR_GUARD_BEGIN
  if (n == 0) return CHPIR_ERR_EMPTY_KV_DATABASE;
  if (!ctx || !key_blob || !key_offsets || !value_blob || !value_offsets || !d_out || !filter_params_out) return CHPIR_ERR_INVALID_ARGUMENT;
  uint64_t max_vlen = 0, K = 0, N = 0;
  for (uint64_t i = 0; i < n; i++) max_vlen = std::max<uint64_t>(max_vlen, value_offsets[i + 1] - value_offsets[i]);
  if (int rc = db_matrix_shape(arity, n, max_vlen, b, &K, &N); rc != CHPIR_OK) return rc;
  DevBuf d;
  if (int rc = encode_kv_database_device(ctx, arity, n, key_blob, key_offsets, value_blob, value_offsets, b, max_attempt_count, filter_seed_rng, K, N,
                                         &d, filter_params_out, nullptr, nullptr);
      rc != CHPIR_OK)
    return rc;
  std::lock_guard<std::mutex> g(ctx->mu);
  CHPIR_CUDA(cudaSetDevice(ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  CHPIR_CUDA(cudaMemcpy(d_out, d.p, K * N * 4, cudaMemcpyDeviceToHost), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_server_setup_device(chpir_ctx *ctx, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint32_t *d_device, uint64_t rows_k,
                              uint32_t cols_n, uint32_t b, const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap,
                              size_t *hint_len, chpir_server **out) {
  return server_setup_from_device_matrix(ctx, seed, d_device, rows_k, cols_n, b, opts, hint_out, hint_cap, hint_len, out, nullptr);
}

extern "C++" {
int chpir::server_setup_from_device_matrix(chpir_ctx *ctx, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint32_t *d_device, uint64_t rows_k,
                                           uint32_t cols_n, uint32_t b, const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap,
                                           size_t *hint_len, chpir_server **out, HostAPipe *pipe) {
  CHPIR_GUARD_BEGIN
  if (!ctx || !seed || !d_device || !out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (rows_k == 0 || cols_n == 0) return CHPIR_ERR_INVALID_MATRIX_DIMENSION;
  if (int rc = validate_bits(b); rc != CHPIR_OK) return rc;
  chpir_setup_opts o{};
  if (opts) o = *opts;
  uint32_t c0, nc;
  if (int rc = resolve_slice(o, cols_n, &c0, &nc); rc != CHPIR_OK) return rc;
  std::lock_guard<std::mutex> g(ctx->mu);
  CHPIR_CUDA(cudaSetDevice(ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  // D was produced by the caller on some other stream: setup is a one-off, so order against everything in flight
  CHPIR_CUDA(cudaDeviceSynchronize(), CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED);
  const double t0 = now_s();
  chpir_server *srv = new chpir_server();
  int rc = setup_core(ctx, seed, d_device, rows_k, cols_n, c0, nc, c0, b, o, hint_out, hint_cap, hint_len, srv, pipe);
  if (rc != CHPIR_OK) {
    delete srv;
    return rc;
  }
  srv->timing.total_s = now_s() - t0;
  *out = srv;
  return CHPIR_OK;
  CHPIR_GUARD_END
}
}  // extern "C++"

extern "C++" {
int chpir::server_setup_from_host_matrix(chpir_ctx *ctx, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint32_t *d_host, uint64_t rows_k,
                                         uint32_t cols_n, uint32_t b, const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap,
                                         size_t *hint_len, chpir_server **out, HostAPipe *pipe) {
  CHPIR_GUARD_BEGIN
  if (!ctx || !seed || !d_host || !out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (rows_k == 0 || cols_n == 0) return CHPIR_ERR_INVALID_MATRIX_DIMENSION;
  if (int rc = validate_bits(b); rc != CHPIR_OK) return rc;
  chpir_setup_opts o{};
  if (opts) o = *opts;
  uint32_t c0, nc;
  if (int rc = resolve_slice(o, cols_n, &c0, &nc); rc != CHPIR_OK) return rc;
  std::lock_guard<std::mutex> g(ctx->mu);
  CHPIR_CUDA(cudaSetDevice(ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  const double t0 = now_s();
  // upload only this server's column slice, compacted to K x nc
  DevBuf d;
  if (int rc = d.alloc(rows_k * nc * 4); rc != CHPIR_OK) return rc;
  CHPIR_CUDA(cudaMemcpy2DAsync(d.p, size_t(nc) * 4, d_host + c0, size_t(cols_n) * 4, size_t(nc) * 4, rows_k, cudaMemcpyHostToDevice, ctx->stream),
             CHPIR_ERR_CUDA_TRANSFER_FAILED);
  CHPIR_CUDA(cudaStreamSynchronize(ctx->stream), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  const double t1 = now_s();
  {
    double tt = t0;
    trace_phase("D upload", tt);
  }
  chpir_server *srv = new chpir_server();
  int rc = setup_core(ctx, seed, d.as<uint32_t>(), rows_k, nc, 0, nc, c0, b, o, hint_out, hint_cap, hint_len, srv, pipe);
  if (rc != CHPIR_OK) {
    delete srv;
    return rc;
  }
  srv->timing.h2d_s = t1 - t0;
  srv->timing.total_s = now_s() - t0;
  *out = srv;
  return CHPIR_OK;
  CHPIR_GUARD_END
}
}  // extern "C++"

int chpir_server_setup(chpir_ctx *ctx, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint32_t *d_host, uint64_t rows_k, uint32_t cols_n,
                       uint32_t b, const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap, size_t *hint_len, chpir_server **out) {
  return server_setup_from_host_matrix(ctx, seed, d_host, rows_k, cols_n, b, opts, hint_out, hint_cap, hint_len, out, nullptr);
}

int chpir_server_setup_from_db(chpir_ctx *ctx, uint32_t arity, const uint8_t seed[CHPIR_SEED_BYTE_LEN], uint64_t n, const uint8_t *key_blob,
                               const uint64_t *key_offsets, const uint8_t *value_blob, const uint64_t *value_offsets,
                               const uint64_t *filter_seed_rng, const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap,
                               size_t *hint_len, uint8_t filter_params_out[CHPIR_FILTER_PARAM_BYTE_LEN], chpir_server **out) {
  CHPIR_GUARD_BEGIN
  if (!out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  // server.rs:104-107
  if (n == 0) return CHPIR_ERR_EMPTY_KV_DATABASE;
  if (!ctx || !seed || !key_blob || !key_offsets || !value_blob || !value_offsets || !filter_params_out) return CHPIR_ERR_INVALID_ARGUMENT;
  if (arity != 3 && arity != 4) return CHPIR_ERR_UNSUPPORTED_ARITY_FOR_BINARY_FUSE_FILTER;
  uint32_t b = 0;
  if (int rc = find_mat_elem_bit_len(n, &b); rc != CHPIR_OK) return rc;
  uint64_t max_vlen = 0;
  for (uint64_t i = 0; i < n; i++) max_vlen = std::max<uint64_t>(max_vlen, value_offsets[i + 1] - value_offsets[i]);
  uint64_t K = 0, N = 0;
  if (int rc = db_matrix_shape(arity, n, max_vlen, b, &K, &N); rc != CHPIR_OK) return rc;
  if (N > 0xffffffffull) return CHPIR_ERR_KV_DATABASE_SIZE_TOO_LARGE;
  const double t0 = now_s();
  // host-pipelined A: the XOF chain depends only on the seed and K, so it starts NOW and runs beside the filter/encode phase;
  // its panel ring is as deep as A itself (lwe_rows x K u32 in HBM) so that it never has to wait for D
  HostAPipe pipe;
  HostAPipe *pipe_p = nullptr;
  bool a_cached = false;
  if (opts && opts->a_cache) {  // A from an earlier setup with this seed: no chain to walk
    std::lock_guard<std::mutex> g(ctx->mu);
    a_cached = ctx->a_cache.matches(seed, opts->lwe_rows ? opts->lwe_rows : CHPIR_LWE_DIMENSION, K);
  }
  chpir_setup_opts o0{};
  if (opts) o0 = *opts;
  if (o0.a_expand != CHPIR_A_EXPAND_DEVICE && !o0.skip_hint && o0.gemm_variant == 0 && !a_cached) {
    const uint32_t m = o0.lwe_rows ? o0.lwe_rows : CHPIR_LWE_DIMENSION;
    if (int rc = pipe.start(ctx->device, seed, m, K, o0.host_chunk_rows, (m + 127) / 128); rc != CHPIR_OK) return rc;
    pipe_p = &pipe;
  }
  if (opts && opts->db_encode == CHPIR_DB_ENCODE_DEVICE) {
    // row encoding + dependent fill on the GPU: D is born in HBM, the host never holds it
    chpir_setup_opts o = *opts;
    uint32_t c0, nc;
    if (int rc = resolve_slice(o, uint32_t(N), &c0, &nc); rc != CHPIR_OK) return rc;
    if (pipe_p) set_encode_threads(std::max(1u, std::thread::hardware_concurrency()) > 3 ? std::thread::hardware_concurrency() - 2 : 1);
    DevBuf d;
    double enc_host_s = 0, enc_dev_s = 0;
    int rc = encode_kv_database_device(ctx, arity, n, key_blob, key_offsets, value_blob, value_offsets, b, CHPIR_SERVER_SETUP_MAX_ATTEMPT_COUNT,
                                       filter_seed_rng, K, N, &d, filter_params_out, &enc_host_s, &enc_dev_s);
    set_encode_threads(0);
    if (rc != CHPIR_OK) return rc;
    std::lock_guard<std::mutex> g(ctx->mu);  // (the fill took and released it itself)
    CHPIR_CUDA(cudaSetDevice(ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
    const double t1 = now_s();
    {
      double tt = t0;
      trace_phase("from_db: start pipe + device encode", tt);
    }
    chpir_server *srv = new chpir_server();
    rc = setup_core(ctx, seed, d.as<uint32_t>(), K, uint32_t(N), c0, nc, c0, b, o, hint_out, hint_cap, hint_len, srv, pipe_p);
    {
      double tt = t1;
      trace_phase("from_db: setup_core", tt);
    }
    if (rc != CHPIR_OK) {
      delete srv;
      return rc;
    }
    srv->timing.host_encode_s = t1 - t0;
    srv->timing.device_encode_s = enc_dev_s;
    srv->timing.total_s = now_s() - t0;
    *out = srv;
    return CHPIR_OK;
  }
  // D lives in pageable memory: pinning 4.4 GB costs ~2 s in cudaMallocHost + cudaFreeHost at 2^20 entries and holds the driver
  // lock the XOF uploader needs meanwhile; the staged pageable upload is a few tenths of a second
  std::unique_ptr<uint32_t[]> d_store(new (std::nothrow) uint32_t[K * N]);
  uint32_t *D = d_store.get();
  if (!D) return CHPIR_ERR_HOST_ALLOCATION_FAILED;
  if (pipe_p) set_encode_threads(std::max(1u, std::thread::hardware_concurrency()) > 3 ? std::thread::hardware_concurrency() - 2 : 1);
  int rc = encode_kv_database(arity, n, key_blob, key_offsets, value_blob, value_offsets, b, CHPIR_SERVER_SETUP_MAX_ATTEMPT_COUNT,
                              filter_seed_rng, D, filter_params_out);
  set_encode_threads(0);
  const double t1 = now_s();
  if (rc == CHPIR_OK) rc = server_setup_from_host_matrix(ctx, seed, D, K, uint32_t(N), b, opts, hint_out, hint_cap, hint_len, out, pipe_p);
  if (rc == CHPIR_OK) {
    (*out)->timing.host_encode_s = t1 - t0;
    (*out)->timing.total_s += t1 - t0;
  }
  return rc;
  CHPIR_GUARD_END
}

void chpir_server_destroy(chpir_server *srv) { delete srv; }

// ---- persisted server state (the reference's Server lives in memory only, server.rs:16-21) ----
namespace {
struct SavedHeader {
  char magic[8];  // "CHPIRSV1"
  uint32_t version, b;
  uint64_t K;
  uint32_t ncols, col_begin, fpw, units;
  uint64_t packed_bytes, checksum;
  uint8_t reserved[8];
};
static_assert(sizeof(SavedHeader) == 64, "on-disk header is 64 bytes");
constexpr char kSavedMagic[8] = {'C', 'H', 'P', 'I', 'R', 'S', 'V', '1'};
constexpr size_t kIoChunk = 64ull << 20;

// FNV-1a over 64-bit words (the payload is a whole number of u64 words)
inline uint64_t fnv1a64(uint64_t h, const uint8_t *p, size_t n) {
  for (size_t i = 0; i + 8 <= n; i += 8) {
    uint64_t w;
    std::memcpy(&w, p + i, 8);
    h = (h ^ w) * 0x100000001b3ull;
  }
  return h;
}
struct FileCloser {
  void operator()(FILE *f) const {
    if (f) std::fclose(f);
  }
};
struct PinnedPair {
  uint8_t *p[2] = {nullptr, nullptr};
  ~PinnedPair() {
    for (auto q : p)
      if (q) cudaFreeHost(q);
  }
};
}  // namespace

int chpir_server_save(chpir_server *srv, const char *path) {
  CHPIR_GUARD_BEGIN
  if (!srv || !path) return CHPIR_ERR_INVALID_ARGUMENT;
  CHPIR_CUDA(cudaSetDevice(srv->ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  // written beside the target and renamed over it once complete and flushed: a crash never leaves a damaged file under `path`,
  // and a good older snapshot survives a failed save
  const std::string tmp = std::string(path) + ".tmp";
  struct TmpGuard {
    const std::string &name;
    bool keep = false;
    ~TmpGuard() {
      if (!keep) std::remove(name.c_str());
    }
  } tmp_guard{tmp};
  std::unique_ptr<FILE, FileCloser> f(std::fopen(tmp.c_str(), "wb"));
  if (!f) return CHPIR_ERR_IO_FAILED;
  SavedHeader h{};
  std::memcpy(h.magic, kSavedMagic, 8);
  h.version = 1, h.b = srv->b, h.K = srv->K, h.ncols = srv->ncols, h.col_begin = srv->col_begin;
  h.fpw = srv->layout.fpw, h.units = srv->layout.units, h.packed_bytes = srv->packed_bytes;
  h.reserved[0] = uint8_t(srv->layout.tight);  // 1 = tight rows (PackedLayout); 0 also in files written before they existed
  if (srv->shard_k_total) {                     // a row block of a cluster's matrix: the rows of the whole matrix, 48 bits LE
    h.reserved[1] = 1;
    for (int i = 0; i < 6; i++) h.reserved[2 + i] = uint8_t(srv->shard_k_total >> (8 * i));
  }
  if (std::fwrite(&h, sizeof h, 1, f.get()) != 1) return CHPIR_ERR_IO_FAILED;
  PinnedPair buf;
  if (cudaMallocHost(&buf.p[0], kIoChunk) != cudaSuccess) return CHPIR_ERR_HOST_ALLOCATION_FAILED;
  uint64_t sum = 0xcbf29ce484222325ull;
  {
    std::lock_guard<std::mutex> g(srv->ctx->mu);
    for (uint64_t off = 0; off < srv->packed_bytes; off += kIoChunk) {
      const size_t n = size_t(std::min<uint64_t>(kIoChunk, srv->packed_bytes - off));
      CHPIR_CUDA(cudaMemcpy(buf.p[0], srv->d_packed + off, n, cudaMemcpyDeviceToHost), CHPIR_ERR_CUDA_TRANSFER_FAILED);
      sum = fnv1a64(sum, buf.p[0], n);
      if (std::fwrite(buf.p[0], 1, n, f.get()) != n) return CHPIR_ERR_IO_FAILED;
    }
  }
  h.checksum = sum;
  if (std::fseek(f.get(), 0, SEEK_SET) != 0 || std::fwrite(&h, sizeof h, 1, f.get()) != 1 || std::fflush(f.get()) != 0) return CHPIR_ERR_IO_FAILED;
  if (fsync(fileno(f.get())) != 0) return CHPIR_ERR_IO_FAILED;
  f.reset();
  if (std::rename(tmp.c_str(), path) != 0) return CHPIR_ERR_IO_FAILED;
  tmp_guard.keep = true;
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_server_load(chpir_ctx *ctx, const char *path, const chpir_setup_opts *opts, chpir_server **out) {
  CHPIR_GUARD_BEGIN
  if (!out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (!ctx || !path) return CHPIR_ERR_INVALID_ARGUMENT;
  chpir_setup_opts o{};
  if (opts) o = *opts;
  std::unique_ptr<FILE, FileCloser> f(std::fopen(path, "rb"));
  if (!f) return CHPIR_ERR_IO_FAILED;
  SavedHeader h{};
  if (std::fread(&h, sizeof h, 1, f.get()) != 1) return CHPIR_ERR_INVALID_SAVED_SERVER;
  if (std::memcmp(h.magic, kSavedMagic, 8) != 0 || h.version != 1) return CHPIR_ERR_INVALID_SAVED_SERVER;
  // K is bounded before it meets any multiplication (TMA coordinates are int32 anyway): a hostile header cannot wrap the size checks
  if (validate_bits(h.b) != CHPIR_OK || h.K == 0 || h.K > 0x7fffffffull || h.ncols == 0 || h.reserved[0] > 1 || h.reserved[1] > 1)
    return CHPIR_ERR_INVALID_SAVED_SERVER;
  uint64_t k_total = 0;
  if (h.reserved[1]) {
    for (int i = 0; i < 6; i++) k_total |= uint64_t(h.reserved[2 + i]) << (8 * i);
    if (k_total == 0 || k_total > 0x7fffffffull) return CHPIR_ERR_INVALID_SAVED_SERVER;
  }
  // the file's own row layout (not whatever make_layout would choose today): the kernels read either
  PackedLayout L{};
  if (!make_layout_explicit(h.b, h.ncols, h.units, h.reserved[0], &L)) return CHPIR_ERR_INVALID_SAVED_SERVER;
  if (L.fpw != h.fpw || h.packed_bytes != h.K * L.pitch_bytes()) return CHPIR_ERR_INVALID_SAVED_SERVER;
  std::lock_guard<std::mutex> g(ctx->mu);
  CHPIR_CUDA(cudaSetDevice(ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  const double t0 = now_s();
  std::unique_ptr<chpir_server> srv(new chpir_server());
  srv->ctx = ctx, srv->K = h.K, srv->ncols = h.ncols, srv->col_begin = h.col_begin, srv->b = h.b;
  srv->layout = L, srv->packed_bytes = h.packed_bytes;
  srv->shard_k_total = k_total;
  {
    void *p = nullptr;
    CHPIR_CUDA(cudaMalloc(&p, L.alloc_bytes(h.K)), CHPIR_ERR_CUDA_ALLOCATION_FAILED);
    srv->d_packed = static_cast<uint8_t *>(p);
    if (L.alloc_bytes(h.K) > srv->packed_bytes)
      CHPIR_CUDA(cudaMemsetAsync(srv->d_packed + srv->packed_bytes, 0, L.alloc_bytes(h.K) - srv->packed_bytes, ctx->stream), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  }
  PinnedPair buf;
  if (cudaMallocHost(&buf.p[0], kIoChunk) != cudaSuccess || cudaMallocHost(&buf.p[1], kIoChunk) != cudaSuccess) return CHPIR_ERR_HOST_ALLOCATION_FAILED;
  cudaStream_t st = ctx->stream;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  struct EvGuard {
    cudaEvent_t *e;
    ~EvGuard() {
      for (int i = 0; i < 2; i++)
        if (e[i]) cudaEventDestroy(e[i]);
    }
  } evg{ev};
  for (auto &e : ev)
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return CHPIR_ERR_CUDA_ALLOCATION_FAILED;
  uint64_t sum = 0xcbf29ce484222325ull;
  int i = 0;
  for (uint64_t off = 0; off < srv->packed_bytes; off += kIoChunk, i ^= 1) {
    const size_t n = size_t(std::min<uint64_t>(kIoChunk, srv->packed_bytes - off));
    cudaEventSynchronize(ev[i]);  // the upload that last used this buffer (no-op the first time round)
    if (std::fread(buf.p[i], 1, n, f.get()) != n) return CHPIR_ERR_INVALID_SAVED_SERVER;  // truncated
    sum = fnv1a64(sum, buf.p[i], n);
    CHPIR_CUDA(cudaMemcpyAsync(srv->d_packed + off, buf.p[i], n, cudaMemcpyHostToDevice, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
    cudaEventRecord(ev[i], st);
  }
  CHPIR_CUDA(cudaStreamSynchronize(st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  if (sum != h.checksum || std::fgetc(f.get()) != EOF) return CHPIR_ERR_INVALID_SAVED_SERVER;
  srv->plan = plan_respond(srv->layout, srv->K, ctx->sm_count);
  if (o.batch_tc == 1) {
    // the limb planes are derived data: rebuild them from the packed rows
    DevBuf d;
    if (int rc = d.alloc(srv->K * srv->ncols * 4); rc != CHPIR_OK) return rc;
    if (int rc = launch_unpack(srv->d_packed, srv->layout, srv->K, d.as<uint32_t>(), st); rc != CHPIR_OK) return rc;
    if (int rc = gemm_tc_prepare(d.as<uint32_t>(), srv->ncols, srv->K, srv->ncols, srv->b, ctx->sm_count, st, &srv->gemm); rc != CHPIR_OK) return rc;
    CHPIR_CUDA(cudaStreamSynchronize(st), CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED);
  }
  if (o.respond_coalesce) {
    if (int rc = srv->init_coalescer(); rc != CHPIR_OK) return rc;
    srv->coalesce = true;
  }
  srv->timing.total_s = now_s() - t0;
  *out = srv.release();
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_server_setup_timing(const chpir_server *srv, chpir_setup_timing *out) {
  if (!srv || !out) return CHPIR_ERR_INVALID_ARGUMENT;
  *out = srv->timing;
  return CHPIR_OK;
}

int chpir_server_get_info(const chpir_server *srv, chpir_server_info *out) {
  if (!srv || !out) return CHPIR_ERR_INVALID_ARGUMENT;
  out->rows_k = srv->K;
  out->cols_n = srv->ncols;
  out->col_begin = srv->col_begin;
  out->mat_elem_bit_len = srv->b;
  out->fields_per_word = srv->layout.fpw;
  out->row_pitch_bytes = srv->layout.pitch_bytes();
  out->packed_bytes = srv->packed_bytes;
  return CHPIR_OK;
}

int chpir_server_hint_device(const chpir_server *srv, const uint32_t **hint_device, uint32_t *rows) {
  if (!srv || !hint_device) return CHPIR_ERR_INVALID_ARGUMENT;
  if (!srv->d_hint) return CHPIR_ERR_INVALID_ARGUMENT;  // set up without chpir_setup_opts.hint_on_device
  *hint_device = srv->d_hint;
  if (rows) *rows = srv->hint_rows;
  return CHPIR_OK;
}

extern "C++" {
// Matrix::from_bytes validation (matrix.rs:973-1010) + the dimension check of
// row_vector_x_compressed_transposed_matrix (matrix.rs:329-331), in that order.
int chpir::validate_query_bytes(uint64_t K, const uint8_t *query, size_t len) {
  if (!query || len <= 8) return CHPIR_ERR_FAILED_TO_DESERIALIZE_MATRIX_FROM_BYTES;
  uint32_t rows, cols;
  std::memcpy(&rows, query, 4);
  std::memcpy(&cols, query + 4, 4);
  const uint64_t n = uint64_t(rows) * cols;
  if (n == 0) return CHPIR_ERR_FAILED_TO_DESERIALIZE_MATRIX_FROM_BYTES;
  if (n * 4 != uint64_t(len - 8)) return CHPIR_ERR_FAILED_TO_DESERIALIZE_MATRIX_FROM_BYTES;
  if (!(rows == 1 && cols == K)) return CHPIR_ERR_INCOMPATIBLE_DIMENSION_FOR_ROW_VECTOR_TRANSPOSED_MATRIX_MULTIPLICATION;
  return CHPIR_OK;
}
static int validate_query(const chpir_server *srv, const uint8_t *query, size_t len) { return chpir::validate_query_bytes(srv->K, query, len); }
}  // extern "C++"

// One caller's share of a coalesced batch (see Coalescer).  query has been validated; resp_out holds 8 + 4*ncols bytes.
static int respond_coalesced(chpir_server *srv, const uint8_t *query, uint8_t *resp_out) {
  Coalescer &co = srv->co;
  CoalesceBatch *B = nullptr;
  uint32_t row = 0;
  {
    std::unique_lock<std::mutex> lk(co.mu);
    co.cv.wait(lk, [&] { return !co.b[co.open].closed && co.b[co.open].count < Coalescer::kMaxBatch; });
    B = &co.b[co.open];
    row = B->count++;
  }
  const bool leader = row == 0;
  // Between joining the batch (count++) and reporting the upload (issued++) a member runs this one C call: nothing here can throw or
  // return early, so a batch cannot be left waiting for a member that never reports; a failed enqueue is reported through B->rc and
  // fails the whole batch.  (A thread killed from outside between the two points would wedge any lock-based protocol the same way.)
  const bool sent = cudaMemcpyAsync(B->d_q + size_t(row) * srv->K, query + 8, srv->K * 4, cudaMemcpyHostToDevice, B->copy) == cudaSuccess;
  {
    std::lock_guard<std::mutex> lk(co.mu);
    B->issued++;
    if (!sent && B->rc == CHPIR_OK) B->rc = CHPIR_ERR_CUDA_TRANSFER_FAILED;
  }
  co.cv.notify_all();
  if (leader) {
    std::lock_guard<std::mutex> ex(co.exec_mu);  // the previous batch leaves the GPU: everyone who arrived meanwhile is in this one
    uint32_t n = 0;
    int rc = CHPIR_OK;
    {
      std::unique_lock<std::mutex> lk(co.mu);
      B->closed = true;
      n = B->count;
      co.cv.wait(lk, [&] { return B->issued == n; });  // every member has enqueued its upload
      rc = B->rc;
      // the other batch becomes the open one as soon as its previous members have all collected their responses
      co.cv.wait(lk, [&] { return co.b[co.open ^ 1].count == 0 && !co.b[co.open ^ 1].closed; });
      co.open ^= 1;
    }
    co.cv.notify_all();
    if (rc == CHPIR_OK) {
      const bool tc = srv->gemm && n >= Coalescer::kTensorCoreFrom;
      cudaEventRecord(B->uploaded, B->copy);
      cudaStreamWaitEvent(co.compute, B->uploaded, 0);
      if (tc) {
        rc = chpir_server_respond_device_tc(srv, B->d_q, n, B->d_resp, co.compute);
      } else {
        if (cudaMemsetAsync(B->d_resp, 0, size_t(n) * srv->ncols * 4, co.compute) != cudaSuccess) rc = CHPIR_ERR_CUDA_TRANSFER_FAILED;
        if (rc == CHPIR_OK) rc = launch_respond(srv->d_packed, srv->layout, srv->K, srv->plan, B->d_q, B->d_resp, n, co.compute);
      }
      if (rc == CHPIR_OK &&
          cudaMemcpyAsync(B->h_resp, B->d_resp, size_t(n) * srv->ncols * 4, cudaMemcpyDeviceToHost, co.compute) != cudaSuccess)
        rc = CHPIR_ERR_CUDA_TRANSFER_FAILED;
      cudaError_t e = cudaStreamSynchronize(co.compute);
      if (rc == CHPIR_OK && e != cudaSuccess) {
        set_last_cuda_error(e, "coalesced respond");
        rc = CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
      }
      co.batches++, co.queries += n, co.tc_batches += tc ? 1 : 0;
    } else {
      cudaStreamSynchronize(B->copy);
    }
    {
      std::lock_guard<std::mutex> lk(co.mu);
      B->rc = rc;
      B->done = true;
    }
    co.cv.notify_all();
  }
  int rc = CHPIR_OK;
  {
    std::unique_lock<std::mutex> lk(co.mu);
    co.cv.wait(lk, [&] { return B->done; });
    rc = B->rc;
  }
  if (rc == CHPIR_OK) {
    const uint32_t hdr[2] = {1u, srv->ncols};
    std::memcpy(resp_out, hdr, 8);
    std::memcpy(resp_out + 8, B->h_resp + size_t(row) * srv->ncols, size_t(srv->ncols) * 4);
  }
  {
    std::lock_guard<std::mutex> lk(co.mu);
    if (++B->picked == B->count) {  // last one out resets the batch for reuse
      B->count = B->issued = B->picked = 0;
      B->closed = B->done = false;
      B->rc = CHPIR_OK;
    }
  }
  co.cv.notify_all();
  return rc;
}

int chpir_server_respond(chpir_server *srv, const uint8_t *query, size_t query_len, uint8_t *resp_out, size_t resp_cap, size_t *resp_len) {
  CHPIR_GUARD_BEGIN
  if (!srv) return CHPIR_ERR_INVALID_ARGUMENT;
  if (int rc = validate_query(srv, query, query_len); rc != CHPIR_OK) return rc;
  const size_t need = 8 + size_t(srv->ncols) * 4;
  if (!resp_out || resp_cap < need) return CHPIR_ERR_BUFFER_TOO_SMALL;
  CHPIR_CUDA(cudaSetDevice(srv->ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  if (srv->coalesce && srv->co.ready) {
    const int rc = respond_coalesced(srv, query, resp_out);
    if (rc == CHPIR_OK && resp_len) *resp_len = need;
    return rc;
  }
  RespondSlot *s = nullptr;
  if (int rc = srv->acquire(&s); rc != CHPIR_OK) return rc;
  int rc = CHPIR_OK;
  do {
    if (cudaMemcpyAsync(s->d_q, query + 8, srv->K * 4, cudaMemcpyHostToDevice, s->stream) != cudaSuccess ||
        cudaMemsetAsync(s->d_resp, 0, size_t(srv->ncols) * 4, s->stream) != cudaSuccess) {
      rc = CHPIR_ERR_CUDA_TRANSFER_FAILED;
      break;
    }
    cudaEventRecord(s->e0, s->stream);
    rc = launch_respond(srv->d_packed, srv->layout, srv->K, srv->plan, s->d_q, s->d_resp, 1, s->stream);
    if (rc != CHPIR_OK) break;
    cudaEventRecord(s->e1, s->stream);
    if (cudaMemcpyAsync(s->h_resp, s->d_resp, size_t(srv->ncols) * 4, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) {
      rc = CHPIR_ERR_CUDA_TRANSFER_FAILED;
      break;
    }
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) {
      set_last_cuda_error(e, "respond");
      rc = CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
      break;
    }
    const uint32_t hdr[2] = {1u, srv->ncols};
    std::memcpy(resp_out, hdr, 8);
    std::memcpy(resp_out + 8, s->h_resp, size_t(srv->ncols) * 4);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->e0, s->e1) == cudaSuccess) srv->last_respond_ms = ms;
    if (resp_len) *resp_len = need;
  } while (false);
  if (rc != CHPIR_OK && rc != CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED) set_last_cuda_error(cudaGetLastError(), "respond enqueue");
  srv->release(s);
  return rc;
  CHPIR_GUARD_END
}

int chpir_server_respond_batch(chpir_server *srv, const uint8_t *const *queries, const size_t *query_lens, uint32_t nq, uint8_t *resp_out,
                               size_t resp_stride) {
  CHPIR_GUARD_BEGIN
  if (!srv || !queries || !query_lens || !resp_out) return CHPIR_ERR_INVALID_ARGUMENT;
  const size_t need = 8 + size_t(srv->ncols) * 4;
  if (resp_stride < need) return CHPIR_ERR_BUFFER_TOO_SMALL;
  for (uint32_t i = 0; i < nq; i++)
    if (int rc = validate_query(srv, queries[i], query_lens[i]); rc != CHPIR_OK) return rc;
  if (nq == 0) return CHPIR_OK;
  CHPIR_CUDA(cudaSetDevice(srv->ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  std::lock_guard<std::mutex> g(srv->batch_mu);
  if (int rc = srv->reserve_batch(nq); rc != CHPIR_OK) return rc;
  cudaStream_t st = srv->batch_stream;
  // all uploads, one launch over the whole batch (grid.y = query), one download
  for (uint32_t i = 0; i < nq; i++)
    CHPIR_CUDA(cudaMemcpyAsync(srv->batch_q + size_t(i) * srv->K, queries[i] + 8, srv->K * 4, cudaMemcpyHostToDevice, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  if (srv->gemm && nq >= Coalescer::kTensorCoreFrom) {
    // one pass over D's limb planes per 128 queries beats nq passes over the packed D from about six queries up
    if (int rc = chpir_server_respond_device_tc(srv, srv->batch_q, nq, srv->batch_resp, st); rc != CHPIR_OK) return rc;
  } else {
    CHPIR_CUDA(cudaMemsetAsync(srv->batch_resp, 0, size_t(nq) * srv->ncols * 4, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
    if (int rc = launch_respond(srv->d_packed, srv->layout, srv->K, srv->plan, srv->batch_q, srv->batch_resp, nq, st); rc != CHPIR_OK) return rc;
  }
  CHPIR_CUDA(cudaMemcpyAsync(srv->batch_h_resp, srv->batch_resp, size_t(nq) * srv->ncols * 4, cudaMemcpyDeviceToHost, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_last_cuda_error(e, "respond batch");
    return CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED;
  }
  const uint32_t hdr[2] = {1u, srv->ncols};
  for (uint32_t i = 0; i < nq; i++) {
    uint8_t *o = resp_out + size_t(i) * resp_stride;
    std::memcpy(o, hdr, 8);
    std::memcpy(o + 8, srv->batch_h_resp + size_t(i) * srv->ncols, size_t(srv->ncols) * 4);
  }
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_server_respond_device(chpir_server *srv, const uint32_t *q_device, uint32_t nq, uint32_t *resp_device, void *cuda_stream) {
  CHPIR_GUARD_BEGIN
  if (!srv || !q_device || !resp_device) return CHPIR_ERR_INVALID_ARGUMENT;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);  // NULL is the CUDA default stream, as for any launch
  CHPIR_CUDA(cudaMemsetAsync(resp_device, 0, size_t(nq) * srv->ncols * 4, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  return launch_respond(srv->d_packed, srv->layout, srv->K, srv->plan, q_device, resp_device, nq, st);
  CHPIR_GUARD_END
}

int chpir_server_respond_device_tc(chpir_server *srv, const uint32_t *q_device, uint32_t nq, uint32_t *resp_device, void *cuda_stream) {
  CHPIR_GUARD_BEGIN
  if (!srv || !q_device || !resp_device) return CHPIR_ERR_INVALID_ARGUMENT;
  if (!srv->gemm) return CHPIR_ERR_INVALID_ARGUMENT;  // set up without limb planes
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  std::lock_guard<std::mutex> g(srv->gemm_mu);
  CHPIR_CUDA(cudaMemsetAsync(resp_device, 0, size_t(nq) * srv->ncols * 4, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  // 128 queries per pass over D: the query block is the "A" operand of the hint GEMM, D's planes the "B" operand
  // (gemm_mu serialises the enqueue; the buffer events order the EXECUTION of calls that arrive on different streams)
  for (uint32_t q0 = 0, p = 0; q0 < nq; q0 += 128, p++) {
    const uint32_t rows = std::min<uint32_t>(128u, nq - q0);
    if (int rc = gemm_tc_buf_acquire(srv->gemm, p & 1, st); rc != CHPIR_OK) return rc;
    if (int rc = gemm_tc_load_panel_u32(srv->gemm, p & 1, q_device + size_t(q0) * srv->K, rows, st); rc != CHPIR_OK) return rc;
    if (int rc = gemm_tc_panel(srv->gemm, p & 1, rows, resp_device + size_t(q0) * srv->ncols, st); rc != CHPIR_OK) return rc;
    if (int rc = gemm_tc_buf_release(srv->gemm, p & 1, st); rc != CHPIR_OK) return rc;
  }
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_generate_from_seed(chpir_ctx *ctx, const uint8_t seed[CHPIR_SEED_BYTE_LEN], uint64_t rows, uint64_t cols, uint64_t row_begin,
                             uint64_t row_count, uint32_t *out_host) {
  CHPIR_GUARD_BEGIN
  if (!ctx || !seed || !out_host) return CHPIR_ERR_INVALID_ARGUMENT;
  if (rows == 0 || cols == 0) return CHPIR_ERR_INVALID_MATRIX_DIMENSION;
  if (row_begin + row_count > rows || row_count == 0) return CHPIR_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> g(ctx->mu);
  CHPIR_CUDA(cudaSetDevice(ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  // only the prefix up to the last requested row has to be produced
  const uint64_t total = (row_begin + row_count) * cols * 4;
  DevBuf a, scratch;
  if (int rc = a.alloc(total); rc != CHPIR_OK) return rc;
  if (int rc = scratch.alloc(512); rc != CHPIR_OK) return rc;
  if (int rc = launch_expand(seed, a.as<uint8_t>(), total, scratch.as<uint8_t>(), ctx->stream); rc != CHPIR_OK) return rc;
  CHPIR_CUDA(cudaMemcpyAsync(out_host, a.as<uint8_t>() + row_begin * cols * 4, row_count * cols * 4, cudaMemcpyDeviceToHost, ctx->stream),
             CHPIR_ERR_CUDA_TRANSFER_FAILED);
  CHPIR_CUDA(cudaStreamSynchronize(ctx->stream), CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED);
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_matmul(chpir_ctx *ctx, const uint32_t *a_host, uint64_t a_rows, uint64_t a_cols, const uint32_t *b_host, uint64_t b_rows,
                 uint64_t b_cols, uint32_t b_elem_bit_len, uint32_t variant, uint32_t *out_host) {
  CHPIR_GUARD_BEGIN
  if (!ctx || !a_host || !b_host || !out_host) return CHPIR_ERR_INVALID_ARGUMENT;
  if (a_rows == 0 || a_cols == 0 || b_rows == 0 || b_cols == 0) return CHPIR_ERR_INVALID_MATRIX_DIMENSION;
  if (a_cols != b_rows) return CHPIR_ERR_INCOMPATIBLE_DIMENSION_FOR_MATRIX_MULTIPLICATION;
  if (a_rows > 0xffffffffull || b_cols > 0xffffffffull) return CHPIR_ERR_INVALID_ARGUMENT;
  if (b_elem_bit_len == 0 || b_elem_bit_len > 32) return CHPIR_ERR_INVALID_ARGUMENT;
  // The limb GEMM splits B into two byte limbs: wider entries go to the u32 SIMT kernel (exact for any operand), and entries that
  // do not fit the width the caller declared are refused instead of being truncated silently by the split.
  if (variant == 0 && b_elem_bit_len > 16) variant = 1;
  if (b_elem_bit_len < 32) {
    const uint32_t limit = 1u << b_elem_bit_len;
    for (uint64_t i = 0; i < b_rows * b_cols; i++)
      if (b_host[i] >= limit) return CHPIR_ERR_INVALID_ARGUMENT;
  }
  std::lock_guard<std::mutex> g(ctx->mu);
  CHPIR_CUDA(cudaSetDevice(ctx->device), CHPIR_ERR_CUDA_DEVICE_NOT_FOUND);
  DevBuf a, b, c;
  if (int rc = a.alloc(a_rows * a_cols * 4); rc != CHPIR_OK) return rc;
  if (int rc = b.alloc(b_rows * b_cols * 4); rc != CHPIR_OK) return rc;
  if (int rc = c.alloc(a_rows * b_cols * 4); rc != CHPIR_OK) return rc;
  cudaStream_t st = ctx->stream;
  CHPIR_CUDA(cudaMemcpyAsync(a.p, a_host, a_rows * a_cols * 4, cudaMemcpyHostToDevice, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  CHPIR_CUDA(cudaMemcpyAsync(b.p, b_host, b_rows * b_cols * 4, cudaMemcpyHostToDevice, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  int rc;
  float ms = 0.f;
  if (variant == 1)
    rc = launch_gemm_simt(a.as<uint32_t>(), b.as<uint32_t>(), uint32_t(b_cols), c.as<uint32_t>(), uint32_t(a_rows), a_cols, uint32_t(b_cols), st);
  else
    rc = launch_gemm_tc(a.as<uint32_t>(), b.as<uint32_t>(), uint32_t(b_cols), c.as<uint32_t>(), uint32_t(a_rows), a_cols, uint32_t(b_cols),
                        b_elem_bit_len, ctx->sm_count, st, &ms);
  if (rc != CHPIR_OK) return rc;
  CHPIR_CUDA(cudaMemcpyAsync(out_host, c.p, a_rows * b_cols * 4, cudaMemcpyDeviceToHost, st), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  CHPIR_CUDA(cudaStreamSynchronize(st), CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED);
  return CHPIR_OK;
  CHPIR_GUARD_END
}

int chpir_host_generate_from_seed(const uint8_t seed[CHPIR_SEED_BYTE_LEN], uint64_t rows, uint64_t cols, uint64_t row_begin,
                                  uint64_t row_count, uint32_t impl, uint32_t *out_host) {
  CHPIR_GUARD_BEGIN
  if (!seed || !out_host || impl > 4) return CHPIR_ERR_INVALID_ARGUMENT;
  if (rows == 0 || cols == 0 || row_begin + row_count > rows) return CHPIR_ERR_INVALID_MATRIX_DIMENSION;
  HostXofStream xs;
  xs.impl = int(impl);
  host_xof_init(&xs.x, seed);
  uint8_t probe[kXofRate];
  HostXof tmp = xs.x;
  if (!host_xof_squeeze_blocks(&tmp, probe, 1, int(impl))) return CHPIR_ERR_INVALID_ARGUMENT;  // implementation not on this CPU
  // walk to the first requested byte: whole blocks are skipped, the partial one is squeezed and dropped
  const uint64_t skip = row_begin * cols * 4;
  host_xof_skip_blocks(&xs.x, skip / kXofRate);
  if (skip % kXofRate) {
    uint8_t drop[kXofRate];
    xs.fill(drop, skip % kXofRate);
  }
  xs.fill(reinterpret_cast<uint8_t *>(out_host), row_count * cols * 4);
  return CHPIR_OK;
  CHPIR_GUARD_END
}

const char *chpir_host_xof_impl(void) { return host_xof_impl_name(); }

int chpir_server_last_kernel_ms(const chpir_server *srv, float *respond_ms, float *gemm_ms, float *expand_ms) {
  if (!srv) return CHPIR_ERR_INVALID_ARGUMENT;
  if (respond_ms) *respond_ms = srv->last_respond_ms;
  if (gemm_ms) *gemm_ms = srv->last_gemm_ms;
  if (expand_ms) *expand_ms = srv->last_expand_ms;
  return CHPIR_OK;
}

}  // extern "C"
