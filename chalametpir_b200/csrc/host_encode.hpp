// host_encode.hpp -- host-side filter construction and row encoding (see host_encode.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "../../include/chalamet_b200.h"

namespace chpir {

struct Slots {
  uint32_t h[4];
};

struct FilterShape {
  uint32_t segment_length;
  uint32_t segment_count;
  uint32_t segment_count_length;
  uint64_t num_fingerprints;
};

// binary_fuse_filter.rs:15-23 BinaryFuseFilter (usize fields fixed at 64 bit)
struct FilterParams {
  uint8_t seed[32];
  uint32_t arity;
  uint32_t segment_length;
  uint32_t segment_count_length;
  uint64_t num_fingerprints;
  uint64_t filter_size;
  uint64_t mat_elem_bit_len;
  void to_bytes(uint8_t out[68]) const;
};

struct PeelResult {
  FilterParams params;
  std::vector<uint64_t> order;         // reverse_order: hashes in peel order
  std::vector<uint8_t> found;          // reverse_h: which of the key's slots it owns
  std::vector<uint32_t> key_of_order;  // index of the key behind order[i]
};

// CHPIR_TRACE=1: phase split of the host side of setup on stderr (diagnostics only); t is advanced to now
double trace_now();
void trace_phase(const char *name, double &t);
// Waves of mutually independent keys for the dependent row fill: members[level_start[l] .. level_start[l+1]) are the peel-order
// indices i of wave l (0-based); wave l only reads rows written by waves < l.
struct FillPlan {
  std::vector<uint32_t> level_start;
  std::vector<uint32_t> members;
};
void plan_fill_levels(uint32_t arity, const PeelResult &pr, FillPlan *plan);
int digest_and_peel(uint32_t arity, uint64_t n, const uint8_t *key_blob, const uint64_t *key_off, uint32_t b, uint32_t max_attempts,
                    const uint64_t *seed_rng, std::vector<uint8_t> *digests, PeelResult *pr);
void set_encode_threads(unsigned n);  // cap on worker threads for this thread's encode calls; 0 = all hardware threads
void key_digest(const uint8_t *key, size_t len, uint8_t out[32]);
uint64_t mix(uint64_t key, uint64_t seed);
uint64_t mix256(const uint8_t digest[32], const uint8_t seed[32]);
Slots slots_of(uint32_t arity, uint64_t hash, uint32_t segment_length, uint32_t segment_count_length);
int find_mat_elem_bit_len(uint64_t n, uint32_t *out);
FilterShape filter_shape(uint32_t arity, uint64_t n);
int db_matrix_shape(uint32_t arity, uint64_t n, uint64_t max_value_len, uint32_t b, uint64_t *rows, uint64_t *cols);
int peel(uint32_t arity, const std::vector<uint8_t> &digests, uint64_t n, uint32_t b, uint32_t max_attempts, const uint64_t *seed_rng,
         PeelResult *res);
void encode_row(const uint8_t digest[32], const uint8_t *value, size_t vlen, uint32_t b, uint32_t *row, uint64_t cols);
int encode_kv_database(uint32_t arity, uint64_t n, const uint8_t *key_blob, const uint64_t *key_off, const uint8_t *val_blob,
                       const uint64_t *val_off, uint32_t b, uint32_t max_attempts, const uint64_t *seed_rng, uint32_t *D,
                       uint8_t filter_bytes[68]);

}  // namespace chpir
