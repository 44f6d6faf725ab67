// common.cuh -- shared declarations for the CUDA translation units of libchalamet_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/chalamet_b200.h"
#include "staged_upload.cuh"

namespace chpir {

void set_last_cuda_error(cudaError_t e, const char *what);

#define CHPIR_CUDA(expr, code)                      \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) {                        \
      ::chpir::set_last_cuda_error(_e, #expr);      \
      return (code);                                \
    }                                               \
  } while (0)

// ---- packed layout of D for respond ------------------------------------------------------------------------
// K-major: row k holds the N columns of D[k][*] as b-bit fields, floor(64/b) fields per little-endian u64 word
// (field f of word j = column j*fpw + f at bit offset f*b), rows padded to a whole number of 16-byte units.
struct PackedLayout {
  uint32_t b;      // bits per element
  uint32_t fpw;    // fields per u64 word
  uint32_t units;  // 16-byte units (two u64 words) per row = columns owned by one thread of the respond kernels
  uint32_t ncols;  // logical columns
  uint32_t tight;  // 1 = rows with an odd number of words are NOT padded to 16 bytes (pitch = 8 * words).  Exists for the narrow
                   // column slices of 8-way sharding: 118 columns x 9 bit = 17 words = 136 bytes instead of 144
  __host__ __device__ uint32_t words() const { return (ncols + fpw - 1) / fpw; }
  __host__ __device__ uint64_t pitch_bytes() const { return tight ? uint64_t(words()) * 8 : uint64_t(units) * 16; }
  // rows to allocate for K logical rows: tight rows are fetched in even-sized groups by the bulk-copy engine (16-byte granules),
  // and the last unit of a row reads 8 bytes into the next one, so one zeroed pad row follows the last
  __host__ __device__ uint64_t alloc_bytes(uint64_t K) const { return (K + (tight ? 1 : 0)) * pitch_bytes(); }
};
PackedLayout make_layout(uint32_t b, uint32_t ncols);
// the layout a saved server was written with (units / tight from its header); false if they do not describe (b, ncols)
bool make_layout_explicit(uint32_t b, uint32_t ncols, uint32_t units, uint32_t tight, PackedLayout *out);

// ---- kernels (defined in respond.cu / expand.cu / gemm_simt.cu / gemm_tc.cu) --------------------------------
// D (u32, K x ld, columns [col_begin, col_begin+ncols)) -> packed rows.
int launch_pack(const uint32_t *d_dev, uint64_t K, uint32_t ld, uint32_t col_begin, const PackedLayout &L, uint8_t *packed,
                cudaStream_t s);
// packed rows -> K x ncols u32 row-major (inverse of launch_pack for a whole slice)
int launch_unpack(const uint8_t *packed, const PackedLayout &L, uint64_t K, uint32_t *d_dev, cudaStream_t s);
struct RespondPlan {
  // register-pipelined kernel (respond_kernel): generic shapes
  uint32_t threads;         // block size = rows_per_iter * units (+ idle lanes)
  uint32_t rows_per_iter;   // k-rows one block covers per loop iteration
  uint32_t grid;            // blocks (K split)
  uint64_t rows_per_block;  // multiple of rows_per_iter
  // bulk-copy ring kernel (respond_ring_kernel): the production path whenever a packed row fits one CTA
  uint32_t ring;            // 1 = use it
  uint32_t ring_R;          // row lanes (multiple of 4)
  uint32_t ring_rpt;        // rows per thread per stage (1, 2 or 4)
  uint32_t ring_stages;
  uint32_t ring_grid, ring_block;
  uint32_t ring_stage_bytes, ring_smem_bytes;
  uint64_t ring_rows_per_cta;
};
RespondPlan plan_respond(const PackedLayout &L, uint64_t K, int sm_count);
// nq queries (q_dev: nq x K, resp_dev: nq x ncols).  resp must be zeroed by the caller on the same stream (the kernel
// accumulates with atomics, exact mod 2^32).
int launch_respond(const uint8_t *packed, const PackedLayout &L, uint64_t K, const RespondPlan &P, const uint32_t *q_dev,
                   uint32_t *resp_dev, uint32_t nq, cudaStream_t s);

// TurboSHAKE128(seed) squeezed into out_dev[0 .. total_bytes) (total_bytes % 4 == 0); serial chain on one warp.
// state_scratch_dev: 512 bytes of device scratch (seed copy + sponge state carried between launches).
int launch_expand(const uint8_t seed[32], uint8_t *out_dev, uint64_t total_bytes, uint8_t *state_scratch_dev, cudaStream_t s);
// The same stream written as K-major byte planes into a ring of two 128-row panel buffers ([4][128][kp] each), for XOF
// blocks [block_begin, block_begin + block_count) of a rows x cols u32 matrix.  expand_begin() first, once.
int expand_begin(const uint8_t seed[32], uint8_t *state_scratch_dev, cudaStream_t s);
int launch_expand_planes(uint8_t *ring_dev, uint64_t rows, uint64_t cols, uint64_t kp, uint8_t *state_scratch_dev, uint64_t block_begin,
                         uint64_t block_count, cudaStream_t s);

// Row fill of Matrix::from_kv_database on the GPU (encode_dev.cu): columns [c0, c0 + nc) of the K x N matrix, stored K x nc u32 (zeroed),
// are filled in the dependency order of the peeling; every pointer except level_start_host is a device pointer.
int launch_device_row_fill(uint32_t arity, const uint32_t *members, const uint32_t *level_start_host, uint32_t waves, const uint64_t *order,
                           const uint8_t *found, const uint32_t *key_of_order, const uint8_t *digests, const uint8_t *values,
                           const uint64_t *val_off, uint32_t *D, uint64_t N, uint32_t c0, uint32_t nc, uint32_t b, uint32_t segment_length,
                           uint32_t segment_count_length, void *scratch_records, uint32_t *scratch_levels, cudaStream_t s);
constexpr size_t kFillRecordBytes = 48;  // scratch_records: this many bytes per key

// C[m x n] = A[m x k] * B[k x n] mod 2^32, all u32 row-major in device memory (B with leading dimension ldb).
int launch_gemm_simt(const uint32_t *A, const uint32_t *B, uint32_t ldb, uint32_t *C, uint32_t m, uint64_t k, uint32_t n, cudaStream_t s);

// tcgen05 int8-limb GEMM (gemm_tc.cu), one 128-row panel of A at a time.
//   gemm_tc_prepare: limb-split + transpose B (k x n u32, entries < 2^b_bits, ld = ldb) once, allocate the A panel ring;
//   gemm_tc_panel:   C_panel[rows x n] += A_panel . B for the panel in ring buffer `buf` (C zeroed by the caller);
//   gemm_tc_load_panel_u32: fill ring buffer `buf` from u32 rows (when A is not produced by launch_expand_planes).
struct GemmTcB;
int gemm_tc_prepare(const uint32_t *B, uint32_t ldb, uint64_t k, uint32_t n, uint32_t b_bits, int sm_count, cudaStream_t s, GemmTcB **out);
int gemm_tc_panel(const GemmTcB *g, int buf, uint32_t rows, uint32_t *C_panel, cudaStream_t s);
int gemm_tc_load_panel_u32(const GemmTcB *g, int buf, const uint32_t *A_rows, uint32_t rows, cudaStream_t s);
uint8_t *gemm_tc_ring(const GemmTcB *g);
uint64_t gemm_tc_panel_bytes(const GemmTcB *g);  // bytes of one ring buffer: [4 limbs][128 rows][kp]
// Cross-stream ordering of the two-buffer operand ring: acquire before refilling `buf` on stream s, release after the panel GEMM
// that read it has been enqueued on s.
int gemm_tc_buf_acquire(const GemmTcB *g, int buf, cudaStream_t s);
int gemm_tc_buf_release(const GemmTcB *g, int buf, cudaStream_t s);
uint64_t gemm_tc_kp(const GemmTcB *g);
void gemm_tc_free(GemmTcB *g);
// Whole product, A: m x k u32 (device), C: m x n u32 (ldc = n), overwritten.  Workspace is allocated/freed internally.
int launch_gemm_tc(const uint32_t *A, const uint32_t *B, uint32_t ldb, uint32_t *C, uint32_t m, uint64_t k, uint32_t n, uint32_t b_bits,
                   int sm_count, cudaStream_t s, float *kernel_ms);

}  // namespace chpir

struct chpir_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;  // serialises setup-type work on `stream`
  // chpir_setup_opts.a_cache: A = generate_from_seed(m, K, seed) as row-major u32, kept between setups (guarded by mu)
  struct ACache {
    uint8_t seed[32] = {};
    uint32_t m = 0;
    uint64_t K = 0;
    uint32_t *a = nullptr;
    bool matches(const uint8_t *s, uint32_t m_, uint64_t K_) const { return a && m == m_ && K == K_ && std::memcmp(seed, s, 32) == 0; }
  } a_cache;
  chpir::StagePool stage;  // page-locked bounce buffers of the multi-threaded pageable upload (staged_upload.cuh), allocated on first use
};
