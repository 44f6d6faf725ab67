// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace chpir {
int launch_gemm_tc(const uint32_t *A, const uint32_t *B, uint32_t ldb, uint32_t *C, uint32_t m, uint64_t k, uint32_t n, uint32_t, int,
                   cudaStream_t s, float *ms) {
  if (ms) *ms = 0.f;
  return launch_gemm_simt(A, B, ldb, C, m, k, n, s);
}
}  // namespace chpir
