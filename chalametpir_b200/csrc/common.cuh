// common.cuh -- shared declarations for the CUDA translation units of libchalamet_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <vector>

#include "../../include/chalamet_b200.h"

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
  uint32_t units;  // 16-byte units per row
  uint32_t ncols;  // logical columns
  __host__ __device__ uint64_t pitch_bytes() const { return uint64_t(units) * 16; }
};
PackedLayout make_layout(uint32_t b, uint32_t ncols);

// ---- kernels (defined in respond.cu / expand.cu / gemm_simt.cu / gemm_tc.cu) --------------------------------
// D (u32, K x ld, columns [col_begin, col_begin+ncols)) -> packed rows.
int launch_pack(const uint32_t *d_dev, uint64_t K, uint32_t ld, uint32_t col_begin, const PackedLayout &L, uint8_t *packed,
                cudaStream_t s);
struct RespondPlan {
  uint32_t threads;         // block size = rows_per_iter * units (+ idle lanes)
  uint32_t rows_per_iter;   // k-rows one block covers per loop iteration
  uint32_t grid;            // blocks (K split)
  uint64_t rows_per_block;  // multiple of rows_per_iter
};
RespondPlan plan_respond(const PackedLayout &L, uint64_t K, int sm_count);
// resp must be zeroed by the caller on the same stream (the kernel accumulates with atomics, exact mod 2^32).
int launch_respond(const uint8_t *packed, const PackedLayout &L, uint64_t K, const RespondPlan &P, const uint32_t *q_dev,
                   uint32_t *resp_dev, cudaStream_t s);

// TurboSHAKE128(seed) squeezed into out_dev[0 .. total_bytes) (total_bytes % 4 == 0); serial chain on one warp.
int launch_expand(const uint8_t seed[32], uint8_t *out_dev, uint64_t total_bytes, uint8_t *state_scratch_dev, cudaStream_t s);

// C[m x n] = A[m x k] * B[k x n] mod 2^32, all u32 row-major in device memory (B with leading dimension ldb).
int launch_gemm_simt(const uint32_t *A, const uint32_t *B, uint32_t ldb, uint32_t *C, uint32_t m, uint64_t k, uint32_t n, cudaStream_t s);

// tcgen05 int8-limb GEMM (gemm_tc.cu).  A: m x k u32 (device), B: k x n u32 with entries < 2^b_bits (device, ld = ldb),
// C: m x n u32 (ldc = n), overwritten.  workspace is allocated/freed internally.
int launch_gemm_tc(const uint32_t *A, const uint32_t *B, uint32_t ldb, uint32_t *C, uint32_t m, uint64_t k, uint32_t n, uint32_t b_bits,
                   int sm_count, cudaStream_t s, float *kernel_ms);

}  // namespace chpir

struct chpir_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;  // serialises setup-type work on `stream`
};
