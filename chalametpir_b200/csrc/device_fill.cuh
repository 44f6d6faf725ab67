// device_fill.cuh -- Matrix::from_kv_database with the row fill on the GPU, split so that a cluster can share the host half.
//   DeviceFillHost  key digests, filter construction (peeling), wave plan, 68-byte filter parameters: once per database;
//   DeviceFillRank  one GPU's columns [c0, c0 + nc) of D: begin() before the host half (allocations, memset, values upload),
//                   finish() after it (plan upload, encode_rows + solve_columns, csrc/encode_dev.cu).
#pragma once
#include <mutex>
#include <vector>

#include "common.cuh"
#include "host_encode.hpp"
#include "host_pipe.cuh"

namespace chpir {

struct DeviceFillHost {
  std::vector<uint8_t> digests;
  PeelResult pr;
  FillPlan plan;
  uint8_t filter_bytes[68] = {};
  int prepare(uint32_t arity, uint64_t n, const uint8_t *key_blob, const uint64_t *key_off, uint32_t b, uint32_t max_attempts, const uint64_t *seed_rng);
};

class DeviceFillRank {
 public:
  int begin(chpir_ctx *ctx, uint64_t n, const uint8_t *val_blob, const uint64_t *val_off, uint64_t K, uint32_t nc, DevBuf *d_out);
  int finish(uint32_t arity, const DeviceFillHost &h, uint64_t N, uint32_t c0, uint32_t b, double *device_s);

 private:
  chpir_ctx *ctx_ = nullptr;
  uint64_t n_ = 0, K_ = 0;
  uint32_t nc_ = 0;
  DevBuf *d_out_ = nullptr;
  std::unique_lock<std::mutex> lock_;  // the ctx's setup mutex, from begin() to the end of finish()
  DevBuf d_values_, d_valoff_, d_fill_rec_, d_fill_levels_, d_members_, d_order_, d_found_, d_koo_, d_digests_;
  StagedUpload values_up_;  // declared last: joins its helpers and drains their streams before the buffers above are freed
};

}  // namespace chpir
