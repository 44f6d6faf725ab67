// gemm_simt.cu -- plain u32 wrapping GEMM on CUDA cores (shared-memory tiled, split-K).
// On-device reference point for the tensor-core hint GEMM (gemm_tc.cu): same result as the reference's
// `impl Mul<&Matrix> for &Matrix` (chalametpir_common/src/matrix.rs:1040-1059) and as shaders/mat_x_mat.glsl:27-47,
// selected with chpir_setup_opts.gemm_variant = 1.  Not the production path.
#include "common.cuh"

namespace chpir {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const uint32_t *__restrict__ A, const uint32_t *__restrict__ B,
                                                         uint32_t *__restrict__ C, uint32_t m, uint64_t k, uint32_t n,
                                                         uint32_t ldb, uint64_t k_per_split) {
  __shared__ uint32_t sa[TK][TM + 1];
  __shared__ uint32_t sb[TK][TN];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const uint32_t m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const uint64_t kb = uint64_t(blockIdx.z) * k_per_split;
  uint64_t ke = kb + k_per_split;
  if (ke > k) ke = k;
  uint32_t acc[4][4] = {};
  for (uint64_t kk = kb; kk < ke; kk += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      const int r = i / TK, c = i % TK;
      const uint32_t gr = m0 + r;
      const uint64_t gc = kk + c;
      sa[c][r] = (gr < m && gc < ke) ? A[uint64_t(gr) * k + gc] : 0u;
    }
    for (int i = threadIdx.x; i < TK * TN; i += 256) {
      const int r = i / TN, c = i % TN;
      const uint64_t gr = kk + r;
      const uint32_t gc = n0 + c;
      sb[r][c] = (gr < ke && gc < n) ? B[gr * ldb + gc] : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < TK; q++) {
      uint32_t av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; i++) av[i] = sa[q][ty * 4 + i], bv[i] = sb[q][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] += av[i] * bv[j];
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      const uint32_t r = m0 + ty * 4 + i, c = n0 + tx * 4 + j;
      if (r < m && c < n) atomicAdd(&C[uint64_t(r) * n + c], acc[i][j]);
    }
}

}  // namespace

int launch_gemm_simt(const uint32_t *A, const uint32_t *B, uint32_t ldb, uint32_t *C, uint32_t m, uint64_t k, uint32_t n, cudaStream_t s) {
  CHPIR_CUDA(cudaMemsetAsync(C, 0, uint64_t(m) * n * 4, s), CHPIR_ERR_CUDA_TRANSFER_FAILED);
  const uint32_t mt = (m + TM - 1) / TM, nt = (n + TN - 1) / TN;
  uint64_t splits = (148ull * 4 + uint64_t(mt) * nt - 1) / (uint64_t(mt) * nt);
  const uint64_t max_splits = (k + 4095) / 4096;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  uint64_t kps = (k + splits - 1) / splits;
  kps = ((kps + TK - 1) / TK) * TK;
  splits = (k + kps - 1) / kps;
  dim3 grid(nt, mt, uint32_t(splits));
  gemm_simt_kernel<<<grid, 256, 0, s>>>(A, B, C, m, k, n, ldb, kps);
  return cudaGetLastError() == cudaSuccess ? CHPIR_OK : CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED;
}

}  // namespace chpir
