// host_encode.cpp -- the part of Server::setup that stays on the host (north_star): binary fuse filter
// construction and key/value row encoding, producing the database matrix D (K x N, u32 row-major) and the
// 68-byte filter parameters.  Product code: it never touches oracle/.
//
// Reference behaviour mirrored (byte-identical D for a given DB and filter seed):
//   chalametpir_common/src/binary_fuse_filter.rs:40-235 (3-wise), :249-456 (4-wise), :462-486 (to_bytes), :519-635
//   chalametpir_common/src/serialization.rs:22-116 (encode_kv_as_row)
//   chalametpir_common/src/matrix.rs:633-648, :687-755, :819-894 (from_kv_database)
//   chalametpir_server/src/server.rs:193-218 (element bit length)
//
// Structure differs from the reference where it helps a 2^20-entry build: key digests are computed once per key
// (the reference hashes every key twice: binary_fuse_filter.rs:113 and serialization.rs:24), in parallel; rows are
// bit-packed in parallel straight into their own slot of D, and only the dependent "subtract the other slots" pass
// runs in reverse peel order.  The peel itself is the reference's serial stack algorithm.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "host_encode.hpp"

namespace chpir {

// CHPIR_TRACE=1 prints the phase split of the host encode to stderr (diagnostics only).
double trace_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
void trace_phase(const char *name, double &t) {
  static const bool on = [] {
    const char *e = std::getenv("CHPIR_TRACE");
    return e && *e && *e != '0';
  }();
  const double now = trace_now();
  if (on) std::fprintf(stderr, "[chpir] setup(host): %-28s %8.3f ms\n", name, (now - t) * 1e3);
  t = now;
}


// ------------------------------------------------------------------------------------------------------------
// TurboSHAKE128 (RFC 9861) for short messages: what the `turboshake` crate =0.4.1 computes at the reference's
// call sites binary_fuse_filter.rs:569-574 and serialization.rs:24-29 (domain separator 0x1F, 32-byte output).
// ------------------------------------------------------------------------------------------------------------
namespace {

constexpr uint64_t kRoundConstants[12] = {  // rounds 12..23 of Keccak-f[1600]
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

inline uint64_t rol(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }

// Keccak-p[1600,12], fully unrolled lane schedule (pi folded into the variable naming).
void keccak_p12(uint64_t s[25]) {
  uint64_t a00 = s[0], a01 = s[1], a02 = s[2], a03 = s[3], a04 = s[4];
  uint64_t a05 = s[5], a06 = s[6], a07 = s[7], a08 = s[8], a09 = s[9];
  uint64_t a10 = s[10], a11 = s[11], a12 = s[12], a13 = s[13], a14 = s[14];
  uint64_t a15 = s[15], a16 = s[16], a17 = s[17], a18 = s[18], a19 = s[19];
  uint64_t a20 = s[20], a21 = s[21], a22 = s[22], a23 = s[23], a24 = s[24];
  for (int r = 0; r < 12; r++) {
    const uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21, c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22,
                   c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23, c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
    const uint64_t d0 = c4 ^ rol(c1, 1), d1 = c0 ^ rol(c2, 1), d2 = c1 ^ rol(c3, 1), d3 = c2 ^ rol(c4, 1), d4 = c3 ^ rol(c0, 1);
    // theta + rho + pi: b[y][2x+3y] = rol(a[x][y] ^ d[x], rho[x][y])
    const uint64_t b00 = a00 ^ d0, b01 = rol(a06 ^ d1, 44), b02 = rol(a12 ^ d2, 43), b03 = rol(a18 ^ d3, 21), b04 = rol(a24 ^ d4, 14);
    const uint64_t b05 = rol(a03 ^ d3, 28), b06 = rol(a09 ^ d4, 20), b07 = rol(a10 ^ d0, 3), b08 = rol(a16 ^ d1, 45),
                   b09 = rol(a22 ^ d2, 61);
    const uint64_t b10 = rol(a01 ^ d1, 1), b11 = rol(a07 ^ d2, 6), b12 = rol(a13 ^ d3, 25), b13 = rol(a19 ^ d4, 8),
                   b14 = rol(a20 ^ d0, 18);
    const uint64_t b15 = rol(a04 ^ d4, 27), b16 = rol(a05 ^ d0, 36), b17 = rol(a11 ^ d1, 10), b18 = rol(a17 ^ d2, 15),
                   b19 = rol(a23 ^ d3, 56);
    const uint64_t b20 = rol(a02 ^ d2, 62), b21 = rol(a08 ^ d3, 55), b22 = rol(a14 ^ d4, 39), b23 = rol(a15 ^ d0, 41),
                   b24 = rol(a21 ^ d1, 2);
    // chi (+ iota on lane 0)
    a00 = b00 ^ (~b01 & b02) ^ kRoundConstants[r];
    a01 = b01 ^ (~b02 & b03);
    a02 = b02 ^ (~b03 & b04);
    a03 = b03 ^ (~b04 & b00);
    a04 = b04 ^ (~b00 & b01);
    a05 = b05 ^ (~b06 & b07);
    a06 = b06 ^ (~b07 & b08);
    a07 = b07 ^ (~b08 & b09);
    a08 = b08 ^ (~b09 & b05);
    a09 = b09 ^ (~b05 & b06);
    a10 = b10 ^ (~b11 & b12);
    a11 = b11 ^ (~b12 & b13);
    a12 = b12 ^ (~b13 & b14);
    a13 = b13 ^ (~b14 & b10);
    a14 = b14 ^ (~b10 & b11);
    a15 = b15 ^ (~b16 & b17);
    a16 = b16 ^ (~b17 & b18);
    a17 = b17 ^ (~b18 & b19);
    a18 = b18 ^ (~b19 & b15);
    a19 = b19 ^ (~b15 & b16);
    a20 = b20 ^ (~b21 & b22);
    a21 = b21 ^ (~b22 & b23);
    a22 = b22 ^ (~b23 & b24);
    a23 = b23 ^ (~b24 & b20);
    a24 = b24 ^ (~b20 & b21);
  }
  s[0] = a00, s[1] = a01, s[2] = a02, s[3] = a03, s[4] = a04, s[5] = a05, s[6] = a06, s[7] = a07, s[8] = a08, s[9] = a09;
  s[10] = a10, s[11] = a11, s[12] = a12, s[13] = a13, s[14] = a14, s[15] = a15, s[16] = a16, s[17] = a17, s[18] = a18;
  s[19] = a19, s[20] = a20, s[21] = a21, s[22] = a22, s[23] = a23, s[24] = a24;
}

constexpr size_t kRate = 168;

}  // namespace

void key_digest(const uint8_t *key, size_t len, uint8_t out[32]) {
  uint64_t st[25] = {0};
  uint8_t *sb = reinterpret_cast<uint8_t *>(st);
  while (len >= kRate) {
    for (size_t i = 0; i < kRate; i++) sb[i] ^= key[i];
    keccak_p12(st);
    key += kRate;
    len -= kRate;
  }
  for (size_t i = 0; i < len; i++) sb[i] ^= key[i];
  sb[len] ^= 0x1f;
  sb[kRate - 1] ^= 0x80;
  keccak_p12(st);
  std::memcpy(out, sb, 32);
}

// ------------------------------------------------------------------------------------------------------------
// hashing helpers (binary_fuse_filter.rs:553-635)
// ------------------------------------------------------------------------------------------------------------
static inline uint64_t fmix64(uint64_t h) {
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdULL;
  h ^= h >> 33;
  h *= 0xc4ceb9fe1a85ec53ULL;
  h ^= h >> 33;
  return h;
}
uint64_t mix(uint64_t key, uint64_t seed) { return fmix64(key + seed); }

uint64_t mix256(const uint8_t digest[32], const uint8_t seed[32]) {
  uint64_t kw[4], sw[4];
  std::memcpy(kw, digest, 32);
  std::memcpy(sw, seed, 32);
  uint64_t sum = 0;
  for (uint64_t k : kw) {
    uint64_t acc = 0;
    for (uint64_t s : sw) acc = fmix64(acc + fmix64(k + s));
    sum += acc;
  }
  return sum;
}

Slots slots_of(uint32_t arity, uint64_t hash, uint32_t segment_length, uint32_t segment_count_length) {
  Slots r{};
  const uint32_t m = segment_length - 1;
  r.h[0] = static_cast<uint32_t>((static_cast<unsigned __int128>(hash) * segment_count_length) >> 64);
  if (arity == 3) {
    r.h[1] = (r.h[0] + segment_length) ^ (static_cast<uint32_t>(hash >> 18) & m);
    r.h[2] = (r.h[0] + 2 * segment_length) ^ (static_cast<uint32_t>(hash) & m);
  } else {
    r.h[1] = (r.h[0] + segment_length) ^ (static_cast<uint32_t>(hash) & m);
    r.h[2] = (r.h[0] + 2 * segment_length) ^ (static_cast<uint32_t>(hash >> 16) & m);
    r.h[3] = (r.h[0] + 3 * segment_length) ^ (static_cast<uint32_t>(hash >> 32) & m);
  }
  return r;
}

// ------------------------------------------------------------------------------------------------------------
// shapes
// ------------------------------------------------------------------------------------------------------------
int find_mat_elem_bit_len(uint64_t n, uint32_t *out) {
  if (n == 0) return CHPIR_ERR_EMPTY_KV_DATABASE;
  uint64_t s = static_cast<uint64_t>(std::sqrt(static_cast<long double>(n)));
  while (s * s > n) --s;
  while ((s + 1) * (s + 1) <= n) ++s;
  // largest b with 2^32 >= 8 * (2^b)^2 * isqrt(n)
  int b = -1;
  for (int cand = 0; cand <= 16; cand++) {
    const unsigned __int128 rhs = static_cast<unsigned __int128>(8) * (1ULL << (2 * cand)) * s;
    if (rhs <= (static_cast<unsigned __int128>(1) << 32))
      b = cand;
    else
      break;
  }
  if (b < 4) return CHPIR_ERR_KV_DATABASE_SIZE_TOO_LARGE;
  *out = static_cast<uint32_t>(b);
  return CHPIR_OK;
}

FilterShape filter_shape(uint32_t arity, uint64_t n) {
  FilterShape fs{};
  const double ln_n = std::log(static_cast<double>(static_cast<uint32_t>(n)));
  double expo = arity == 3 ? std::floor(ln_n / std::log(3.33) + 2.25) : std::floor(ln_n / std::log(2.91) - 0.5);
  if (!(expo > 0)) expo = 0;
  uint32_t sl = n == 0 ? 4u : (1u << static_cast<unsigned>(expo));
  sl = std::min(sl, 1u << 18);
  const double factor = arity == 3 ? std::max(1.125, 0.875 + 0.25 * std::log(1e6) / ln_n) : std::max(1.075, 0.77 + 0.305 * std::log(6e5) / ln_n);
  const uint32_t capacity = n > 1 ? static_cast<uint32_t>(std::round(static_cast<double>(n) * factor)) : 0u;
  const uint32_t init_segments = (capacity + sl - 1) / sl;
  const uint32_t segments = init_segments < arity ? 1u : init_segments - (arity - 1);
  fs.segment_length = sl;
  fs.segment_count = segments;
  fs.segment_count_length = segments * sl;
  fs.num_fingerprints = static_cast<uint64_t>(segments + arity - 1) * sl;
  return fs;
}

int db_matrix_shape(uint32_t arity, uint64_t n, uint64_t max_value_len, uint32_t b, uint64_t *rows, uint64_t *cols) {
  if (arity != 3 && arity != 4) return CHPIR_ERR_UNSUPPORTED_ARITY_FOR_BINARY_FUSE_FILTER;
  if (n == 0) return CHPIR_ERR_EMPTY_KV_DATABASE;
  if (b < 4 || b > 14) return CHPIR_ERR_IMPOSSIBLE_ENCODED_DB_MATRIX_ELEMENT_BIT_LENGTH;
  *rows = filter_shape(arity, n).num_fingerprints;
  *cols = (256 + 8 * max_value_len + 8 + b - 1) / b;
  return CHPIR_OK;
}

void FilterParams::to_bytes(uint8_t out[68]) const {
  std::memcpy(out, seed, 32);
  std::memcpy(out + 32, &arity, 4);
  std::memcpy(out + 36, &segment_length, 4);
  std::memcpy(out + 40, &segment_count_length, 4);
  std::memcpy(out + 44, &num_fingerprints, 8);
  std::memcpy(out + 52, &filter_size, 8);
  std::memcpy(out + 60, &mat_elem_bit_len, 8);
}

// ------------------------------------------------------------------------------------------------------------
// parallel helper
// ------------------------------------------------------------------------------------------------------------
// worker-thread cap for the calling thread's parallel_for calls (0 = one per hardware thread); Server::setup lowers it while
// the host XOF producer is running beside the encode so that the serial chain keeps a core to itself
static thread_local unsigned t_max_threads = 0;
void set_encode_threads(unsigned n) { t_max_threads = n; }

template <class F>
static void parallel_for(uint64_t n, uint64_t grain, F &&fn) {
  unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  if (t_max_threads) nt = std::min(nt, t_max_threads);
  nt = static_cast<unsigned>(std::min<uint64_t>(nt, (n + grain - 1) / std::max<uint64_t>(grain, 1)));
  if (nt <= 1) {
    fn(0, n);
    return;
  }
  std::atomic<uint64_t> next{0};
  std::vector<std::thread> pool;
  pool.reserve(nt);
  for (unsigned t = 0; t < nt; t++)
    pool.emplace_back([&] {
      for (;;) {
        const uint64_t lo = next.fetch_add(grain);
        if (lo >= n) break;
        fn(lo, std::min(n, lo + grain));
      }
    });
  for (auto &th : pool) th.join();
}

// ------------------------------------------------------------------------------------------------------------
// peeling (binary_fuse_filter.rs:102-215 / :311-436)
// ------------------------------------------------------------------------------------------------------------
namespace {

struct CandidateSeeds {
  bool deterministic;
  uint64_t state;
  std::random_device os;
  explicit CandidateSeeds(const uint64_t *seed) : deterministic(seed != nullptr), state(seed ? *seed : 0) {}
  void next(uint8_t out[32]) {
    for (int w = 0; w < 4; w++) {
      uint64_t v;
      if (deterministic) {
        uint64_t z = (state += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        v = z ^ (z >> 31);
      } else {
        v = (static_cast<uint64_t>(os()) << 32) | os();
      }
      std::memcpy(out + 8 * w, &v, 8);
    }
  }
};

}  // namespace

int peel(uint32_t arity, const std::vector<uint8_t> &digests, uint64_t n, uint32_t b, uint32_t max_attempts,
         const uint64_t *seed_rng, PeelResult *res) {
  const FilterShape fs = filter_shape(arity, n);
  const uint64_t slots = fs.num_fingerprints;
  std::vector<uint8_t> count(slots);
  std::vector<uint64_t> xored(slots);
  std::vector<uint32_t> xidx(slots);  // XOR of the key indices on a slot: the index of the last key left, next to its hash in xored
  std::vector<uint32_t> alone(slots);
  std::vector<uint64_t> hashes(n);
  res->order.assign(n, 0);
  res->found.assign(n, 0);
  res->key_of_order.assign(n, 0);

  // The reference first bucket-sorts the hashes by their top bits (binary_fuse_filter.rs:108-126) purely for cache
  // locality of the counting pass; the count/xor tables, and hence the peel order, do not depend on that order.
  // We keep the locality trick with an index sort so each hash still knows its key.
  unsigned block_bits = 1;
  while ((1u << block_bits) < fs.segment_count) block_bits++;
  std::vector<uint32_t> bucket_start((size_t(1) << block_bits) + 1);
  std::vector<uint32_t> sorted_idx(n);

  CandidateSeeds seeds(seed_rng);
  for (uint32_t attempt = 0; attempt < max_attempts; attempt++) {
    uint8_t seed[32];
    seeds.next(seed);
    parallel_for(n, 1 << 14, [&](uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; i++) hashes[i] = mix256(&digests[32 * i], seed);
    });
    std::fill(bucket_start.begin(), bucket_start.end(), 0u);
    for (uint64_t i = 0; i < n; i++) bucket_start[(hashes[i] >> (64 - block_bits)) + 1]++;
    for (size_t k = 1; k < bucket_start.size(); k++) bucket_start[k] += bucket_start[k - 1];
    for (uint64_t i = 0; i < n; i++) sorted_idx[bucket_start[hashes[i] >> (64 - block_bits)]++] = static_cast<uint32_t>(i);

    std::fill(count.begin(), count.end(), 0);
    std::fill(xored.begin(), xored.end(), 0);
    std::fill(xidx.begin(), xidx.end(), 0u);
    uint8_t seen = 0;
    bool wrapped = false;
    for (uint64_t j = 0; j < n; j++) {
      const uint32_t key = sorted_idx[j];
      const uint64_t h = hashes[key];
      const Slots s = slots_of(arity, h, fs.segment_length, fs.segment_count_length);
      for (uint32_t a = 0; a < arity; a++) {
        wrapped |= count[s.h[a]] >= 252;
        count[s.h[a]] = static_cast<uint8_t>((count[s.h[a]] + 4) ^ a);
        xored[s.h[a]] ^= h;
        xidx[s.h[a]] ^= key;
        seen |= count[s.h[a]];
      }
    }
    // Occupancy overflow.  4-wise: the reference retries once any slot holds >= 32 keys (:337-368).  3-wise: it
    // carries on up to the 6-bit counter's limit and only a wrap (>= 64 keys on a slot) spoils the attempt (:144).
    if (wrapped || (arity == 4 && seen >= 0x80)) continue;

    uint64_t q = 0;
    for (uint64_t i = 0; i < slots; i++) {
      alone[q] = static_cast<uint32_t>(i);
      q += (count[i] >> 2) == 1;
    }
    uint64_t top = 0;
    while (q > 0) {
      const uint32_t slot = alone[--q];
      if ((count[slot] >> 2) != 1) continue;
      const uint64_t h = xored[slot];
      const uint32_t key = xidx[slot];
      const uint8_t which = count[slot] & 3;
      res->found[top] = which;
      res->order[top] = h;
      res->key_of_order[top] = key;  // hash -> key (the reference's HashMap<u64,&[u8]> hash_to_key) without a lookup
      top++;
      const Slots s = slots_of(arity, h, fs.segment_length, fs.segment_count_length);
      for (uint32_t step = 1; step < arity; step++) {
        const uint32_t a = (which + step) % arity;
        const uint32_t other = s.h[a];
        alone[q] = other;
        q += (count[other] >> 2) == 2;
        count[other] = static_cast<uint8_t>((count[other] - 4) ^ a);
        xored[other] ^= h;
        xidx[other] ^= key;
      }
    }
    if (top != n) continue;

    std::memcpy(res->params.seed, seed, 32);
    res->params.arity = arity;
    res->params.segment_length = fs.segment_length;
    res->params.segment_count_length = fs.segment_count_length;
    res->params.num_fingerprints = slots;
    res->params.filter_size = n;
    res->params.mat_elem_bit_len = b;
    return CHPIR_OK;
  }
  return arity == 3 ? CHPIR_ERR_EXHAUSTED_ALL_ATTEMPTS_TO_BUILD_3_WISE_XOR_FILTER
                    : CHPIR_ERR_EXHAUSTED_ALL_ATTEMPTS_TO_BUILD_4_WISE_XOR_FILTER;
}

// ------------------------------------------------------------------------------------------------------------
// dependency levels of the row fill (matrix.rs:707-746 / :839-885 walk the keys in reverse peel order)
// ------------------------------------------------------------------------------------------------------------
// Key i's row is  own = enc(i) - sum(rows of its other slots) - mix(hash, e).  A slot it reads is either never owned (all zero)
// or owned by a key that was peeled LATER (a key is peeled only when it is alone on its slot), so the reverse peel order is one
// valid schedule -- but not the only one: level(i) = 1 + max(level of the owners of i's other slots) groups the keys into waves
// whose members are independent of each other.  The waves are what a parallel (GPU) row fill executes.
void plan_fill_levels(uint32_t arity, const PeelResult &pr, FillPlan *plan) {
  const FilterParams &fp = pr.params;
  const uint64_t n = pr.order.size();
  std::vector<uint32_t> slot_level(fp.num_fingerprints, 0);  // level of the key that owns the slot, 0 = not owned (yet)
  std::vector<uint32_t> level(n);
  uint32_t max_level = 0;
  for (uint64_t i = n; i-- > 0;) {
    const Slots s = slots_of(arity, pr.order[i], fp.segment_length, fp.segment_count_length);
    const uint32_t which = pr.found[i];
    uint32_t l = 0;
    for (uint32_t step = 1; step < arity; step++) l = std::max(l, slot_level[s.h[(which + step) % arity]]);
    level[i] = l + 1;
    slot_level[s.h[which]] = l + 1;
    max_level = std::max(max_level, l + 1);
  }
  std::vector<uint32_t> counts(size_t(max_level) + 1, 0);
  for (uint64_t i = 0; i < n; i++) counts[level[i]]++;
  plan->level_start.assign(size_t(max_level) + 1, 0);
  for (uint32_t l = 1; l <= max_level; l++) plan->level_start[l] = plan->level_start[l - 1] + counts[l];
  plan->members.resize(n);
  std::vector<uint32_t> cursor(plan->level_start.begin(), plan->level_start.end() - 1);
  for (uint64_t i = 0; i < n; i++) plan->members[cursor[level[i] - 1]++] = static_cast<uint32_t>(i);
}

// ------------------------------------------------------------------------------------------------------------
// row codec (serialization.rs:22-116): digest || value || 0x81, LSB-first b-bit fields
// ------------------------------------------------------------------------------------------------------------
// The byte stream digest || value || 0x81 is laid out once in a zero-padded scratch buffer; field f is then bits [f*b, f*b + b) of
// it, read with one unaligned 64-bit load (b <= 14, so a field never spans more than 3 bytes).  Same fields as pushing the bytes
// through a bit accumulator LSB-first and flushing the remainder (serialization.rs:31-113), a few times faster.
void encode_row(const uint8_t digest[32], const uint8_t *value, size_t vlen, uint32_t b, uint32_t *row, uint64_t cols) {
  static thread_local std::vector<uint8_t> scratch;
  const size_t stream = 32 + vlen + 1;
  const size_t need = std::max<size_t>(stream, (cols * b + 7) / 8) + 8;
  if (scratch.size() < need) scratch.resize(need);
  uint8_t *buf = scratch.data();
  std::memcpy(buf, digest, 32);
  if (vlen) std::memcpy(buf + 32, value, vlen);
  buf[32 + vlen] = 0x81;
  std::memset(buf + stream, 0, need - stream);
  const uint32_t mask = (1u << b) - 1u;
  // fields past the end of the stream read zeros; a row narrower than the stream (cannot happen for shapes from db_matrix_shape)
  // would simply be truncated, as the reference's writer would overrun -- callers size cols from the longest value
  for (uint64_t f = 0; f < cols; f++) {
    const uint64_t bit = f * b;
    uint64_t w;
    std::memcpy(&w, buf + (bit >> 3), 8);
    row[f] = static_cast<uint32_t>(w >> (bit & 7)) & mask;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Matrix::from_kv_database
// ------------------------------------------------------------------------------------------------------------
// key digests (serialization.rs:24-29 / binary_fuse_filter.rs:569-574) + filter construction: the part of the encode that
// stays on the host in every mode
int digest_and_peel(uint32_t arity, uint64_t n, const uint8_t *key_blob, const uint64_t *key_off, uint32_t b, uint32_t max_attempts,
                    const uint64_t *seed_rng, std::vector<uint8_t> *digests, PeelResult *pr) {
  if (arity != 3 && arity != 4) return CHPIR_ERR_UNSUPPORTED_ARITY_FOR_BINARY_FUSE_FILTER;
  if (n == 0) return CHPIR_ERR_EMPTY_KV_DATABASE;
  if (b < 4 || b > 14) return CHPIR_ERR_IMPOSSIBLE_ENCODED_DB_MATRIX_ELEMENT_BIT_LENGTH;
  if (n > 0xffffffffULL) return CHPIR_ERR_KV_DATABASE_SIZE_TOO_LARGE;
  double tt = trace_now();
  digests->resize(32 * n);
  parallel_for(n, 1 << 12, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) key_digest(key_blob + key_off[i], key_off[i + 1] - key_off[i], &(*digests)[32 * i]);
  });
  trace_phase("key digests", tt);
  const int rc = peel(arity, *digests, n, b, max_attempts, seed_rng, pr);
  trace_phase("peel", tt);
  return rc;
}

int encode_kv_database(uint32_t arity, uint64_t n, const uint8_t *key_blob, const uint64_t *key_off, const uint8_t *val_blob,
                       const uint64_t *val_off, uint32_t b, uint32_t max_attempts, const uint64_t *seed_rng, uint32_t *D,
                       uint8_t filter_bytes[68]) {
  if (arity != 3 && arity != 4) return CHPIR_ERR_UNSUPPORTED_ARITY_FOR_BINARY_FUSE_FILTER;
  if (n == 0) return CHPIR_ERR_EMPTY_KV_DATABASE;
  if (b < 4 || b > 14) return CHPIR_ERR_IMPOSSIBLE_ENCODED_DB_MATRIX_ELEMENT_BIT_LENGTH;
  if (n > 0xffffffffULL) return CHPIR_ERR_KV_DATABASE_SIZE_TOO_LARGE;

  double tt = trace_now();
  std::vector<uint8_t> digests;
  uint64_t max_vlen = 0;
  for (uint64_t i = 0; i < n; i++) max_vlen = std::max<uint64_t>(max_vlen, val_off[i + 1] - val_off[i]);
  PeelResult pr;
  if (int rc = digest_and_peel(arity, n, key_blob, key_off, b, max_attempts, seed_rng, &digests, &pr); rc != CHPIR_OK) return rc;

  uint64_t K = 0, N = 0;
  db_matrix_shape(arity, n, max_vlen, b, &K, &N);
  const uint32_t mask = (1u << b) - 1;
  const FilterParams &fp = pr.params;

  // the slot rows of every key, own row first, in peel order (computed once, in parallel)
  std::vector<uint32_t> rows(n * arity);
  std::vector<uint8_t> owned(K, 0);
  parallel_for(n, 1 << 12, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) {
      const Slots s = slots_of(arity, pr.order[i], fp.segment_length, fp.segment_count_length);
      for (uint32_t j = 0; j < arity; j++) rows[i * arity + j] = s.h[(pr.found[i] + j) % arity];
      owned[rows[i * arity]] = 1;  // distinct keys own distinct rows: no two threads write the same byte
    }
  });
  // 1. write every key's packed row into its own slot and zero the rows no key owns -- independent, parallel.
  parallel_for(K, 256, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t r = lo; r < hi; r++)
      if (!owned[r]) std::memset(D + r * N, 0, N * sizeof(uint32_t));
  });
  parallel_for(n, 256, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) {
      const uint64_t key = pr.key_of_order[i];
      encode_row(&digests[32 * key], val_blob + val_off[key], val_off[key + 1] - val_off[key], b, D + uint64_t(rows[i * arity]) * N, N);
    }
  });
  trace_phase("zero + own rows", tt);
  // 2. dependent pass in reverse peel order (matrix.rs:707-746 / :839-885).  A slot read here is either final
  //    (its key was peeled later, i.e. handled earlier in this loop) or never owned by any key (all zero), because a
  //    key is peeled only when it is the last one left on its own slot.
  //    The dependency runs from key to key, never from column to column, so the columns are split among the worker threads
  //    and every thread walks ALL keys in the reference's order over its own column range: no synchronisation, the same
  //    arithmetic per element, hence the same D.
  unsigned workers = std::max(1u, std::thread::hardware_concurrency());
  if (t_max_threads) workers = std::min(workers, t_max_threads);
  // column ranges in multiples of 16 (one cache line of u32), at least 32 columns each
  const uint64_t col_grain = std::max<uint64_t>(32, ((N + workers - 1) / workers + 15) / 16 * 16);
  constexpr uint64_t kAhead = 6;  // rows of the key this many steps ahead are prefetched (random 3.8 KB rows: DRAM latency bound otherwise)
  parallel_for(N, col_grain, [&](uint64_t c0, uint64_t c1) {
    for (uint64_t i = n; i-- > 0;) {
      if (i >= kAhead) {
        const uint32_t *r = &rows[(i - kAhead) * arity];
        for (uint32_t j = 0; j < arity; j++) {
          const char *p = reinterpret_cast<const char *>(D + uint64_t(r[j]) * N + c0);
          for (uint64_t off = 0; off < (c1 - c0) * 4; off += 64) __builtin_prefetch(p + off, 0, 1);
        }
      }
      const uint64_t hash = pr.order[i];
      const uint32_t *r = &rows[i * arity];
      uint32_t *own = D + uint64_t(r[0]) * N;
      const uint32_t *o1 = D + uint64_t(r[1]) * N;
      const uint32_t *o2 = D + uint64_t(r[2]) * N;
      if (arity == 3) {
        for (uint64_t e = c0; e < c1; e++) own[e] = (own[e] - o1[e] - o2[e] - static_cast<uint32_t>(mix(hash, e))) & mask;
      } else {
        const uint32_t *o3 = D + uint64_t(r[3]) * N;
        for (uint64_t e = c0; e < c1; e++) own[e] = (own[e] - o1[e] - o2[e] - o3[e] - static_cast<uint32_t>(mix(hash, e))) & mask;
      }
    }
  });
  trace_phase("dependent row fill", tt);
  if (std::getenv("CHPIR_TRACE_LEVELS")) {
    FillPlan plan;
    plan_fill_levels(arity, pr, &plan);
    trace_phase("plan_fill_levels", tt);
    const size_t L = plan.level_start.size() - 1;
    uint32_t big = 0;
    for (size_t l = 0; l < L; l++) big = std::max(big, plan.level_start[l + 1] - plan.level_start[l]);
    std::fprintf(stderr, "[chpir] fill waves: %zu, largest %u, mean %.1f\n", L, big, double(n) / double(L));
  }
  fp.to_bytes(filter_bytes);
  return CHPIR_OK;
}

}  // namespace chpir
