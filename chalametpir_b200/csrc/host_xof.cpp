// host_xof.cpp -- the TurboSHAKE128 stream of Matrix::generate_from_seed (chalametpir_common/src/matrix.rs:541-558) produced on a HOST
// core, for the host-pipelined setup mode (chpir_setup_opts.a_expand = CHPIR_A_EXPAND_HOST_PIPELINED).
//
// The squeeze is one serial chain of Keccak-p[1600,12] permutations (RFC 9861), so what matters is the latency of one
// permutation on one core.  A CPU core runs that chain several times faster than one GPU warp can (csrc/expand.cu); the
// reference's own `gpu` feature also expands A on the CPU (chalametpir_server/src/server.rs:115) and uploads it.  Three
// formulations (four builds), picked at run time:
//   * AVX-512 (F + VL): one 64-bit lane per qword slot, a plane A[0..4][y] per zmm register; theta and chi are three-input
//     vpternlogq, rho is vprolvq, pi is the slot permutation that makes chi register-wise, and the only cross-register data
//     movement per round is one 5x5 qword transposition (two shuffle levels + one trip through the store buffer);
//   * scalar 64-bit code (two rounds unrolled, lanes in locals), built twice: baseline x86-64 and BMI2 (rorx / andn);
//   * the same lane-per-register formulation on xmm registers with AVX-512VL instructions (EVEX-128): three-operand
//     vpternlogq / vprolq, 32 registers -- 101 ns per permutation on the measured Xeon against 108.5 (BMI2) and 113.7 (AVX-512 planes).
// impl 0 times each available variant (a few ms, after a warm-up long enough for the core's 512-bit frequency licence to settle)
// and keeps the fastest.
#include "host_xof.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace chpir {
namespace {

constexpr uint64_t kRC[12] = {0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
                              0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
                              0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};

static inline uint64_t rol(uint64_t v, unsigned r) { return (v << r) | (v >> ((64 - r) & 63)); }

// One round, reading lanes A.. and writing lanes E.. (the classic two-array formulation: pi is a renaming).
#define CHPIR_ROUND(A, E, RC)                                                                  \
  do {                                                                                         \
    const uint64_t c0 = A[0] ^ A[5] ^ A[10] ^ A[15] ^ A[20];                                   \
    const uint64_t c1 = A[1] ^ A[6] ^ A[11] ^ A[16] ^ A[21];                                   \
    const uint64_t c2 = A[2] ^ A[7] ^ A[12] ^ A[17] ^ A[22];                                   \
    const uint64_t c3 = A[3] ^ A[8] ^ A[13] ^ A[18] ^ A[23];                                   \
    const uint64_t c4 = A[4] ^ A[9] ^ A[14] ^ A[19] ^ A[24];                                   \
    const uint64_t d0 = c4 ^ rol(c1, 1), d1 = c0 ^ rol(c2, 1), d2 = c1 ^ rol(c3, 1);           \
    const uint64_t d3 = c2 ^ rol(c4, 1), d4 = c3 ^ rol(c0, 1);                                 \
    uint64_t b0, b1, b2, b3, b4;                                                               \
    b0 = A[0] ^ d0, b1 = rol(A[6] ^ d1, 44), b2 = rol(A[12] ^ d2, 43), b3 = rol(A[18] ^ d3, 21), b4 = rol(A[24] ^ d4, 14); \
    E[0] = b0 ^ (~b1 & b2) ^ (RC), E[1] = b1 ^ (~b2 & b3), E[2] = b2 ^ (~b3 & b4), E[3] = b3 ^ (~b4 & b0), E[4] = b4 ^ (~b0 & b1); \
    b0 = rol(A[3] ^ d3, 28), b1 = rol(A[9] ^ d4, 20), b2 = rol(A[10] ^ d0, 3), b3 = rol(A[16] ^ d1, 45), b4 = rol(A[22] ^ d2, 61); \
    E[5] = b0 ^ (~b1 & b2), E[6] = b1 ^ (~b2 & b3), E[7] = b2 ^ (~b3 & b4), E[8] = b3 ^ (~b4 & b0), E[9] = b4 ^ (~b0 & b1); \
    b0 = rol(A[1] ^ d1, 1), b1 = rol(A[7] ^ d2, 6), b2 = rol(A[13] ^ d3, 25), b3 = rol(A[19] ^ d4, 8), b4 = rol(A[20] ^ d0, 18); \
    E[10] = b0 ^ (~b1 & b2), E[11] = b1 ^ (~b2 & b3), E[12] = b2 ^ (~b3 & b4), E[13] = b3 ^ (~b4 & b0), E[14] = b4 ^ (~b0 & b1); \
    b0 = rol(A[4] ^ d4, 27), b1 = rol(A[5] ^ d0, 36), b2 = rol(A[11] ^ d1, 10), b3 = rol(A[17] ^ d2, 15), b4 = rol(A[23] ^ d3, 56); \
    E[15] = b0 ^ (~b1 & b2), E[16] = b1 ^ (~b2 & b3), E[17] = b2 ^ (~b3 & b4), E[18] = b3 ^ (~b4 & b0), E[19] = b4 ^ (~b0 & b1); \
    b0 = rol(A[2] ^ d2, 62), b1 = rol(A[8] ^ d3, 55), b2 = rol(A[14] ^ d4, 39), b3 = rol(A[15] ^ d0, 41), b4 = rol(A[21] ^ d1, 2); \
    E[20] = b0 ^ (~b1 & b2), E[21] = b1 ^ (~b2 & b3), E[22] = b2 ^ (~b3 & b4), E[23] = b3 ^ (~b4 & b0), E[24] = b4 ^ (~b0 & b1); \
  } while (0)

#define CHPIR_SQUEEZE_BODY                                  \
  uint64_t a[25], e[25];                                    \
  std::memcpy(a, s, sizeof a);                              \
  for (uint64_t i = 0; i < nblocks; i++) {                  \
    for (int r = 0; r < 12; r += 2) {                       \
      CHPIR_ROUND(a, e, kRC[r]);                            \
      CHPIR_ROUND(e, a, kRC[r + 1]);                        \
    }                                                       \
    std::memcpy(out + i * kXofRate, a, kXofRate);           \
  }                                                         \
  std::memcpy(s, a, sizeof a);

void squeeze_scalar(uint64_t s[25], uint8_t *out, uint64_t nblocks) { CHPIR_SQUEEZE_BODY }

#if defined(__x86_64__)
// the same code with rorx / andn available: about twice as fast as the baseline x86-64 build of it
__attribute__((target("bmi,bmi2"))) void squeeze_bmi2(uint64_t s[25], uint8_t *out, uint64_t nblocks) { CHPIR_SQUEEZE_BODY }
#endif

#if defined(__x86_64__)
// ---- AVX-512.  Plane representation: register P[y] holds A[0..4][y], lane x in qword slot kTau[x] (slots 3, 6, 7 are don't-care).
// Per round: theta is two three-input XORs, two slot permutes of the parity vector and five vpternlogq; rho is vprolvq; pi is one
// slot permute per plane, after which chi is register-wise (R[X] slot Y = A'[X][Y]).  The 5x5 transposition back to planes is
// two levels of two-source shuffles (in-lane unpack, then 128-bit lane select) for X = 0..3; the fifth source R[4] goes through
// memory: one 64-byte store and five merge-masked 8-byte broadcast loads, which run on the load ports instead of the one
// shuffle port every vpermq needs (16 shuffle uops per round instead of 21, and a shorter dependent chain).
constexpr unsigned kRho[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
constexpr int kTau[5] = {0, 1, 4, 5, 2};  // slot of lane x: what unpack + lane select produce for x = 0..3; x = 4 is merged into slot 2

struct Avx512Consts {
  alignas(64) uint64_t rho[5][8];  // rho[y][kTau[x]] = rotation of lane (x, y)
  alignas(64) uint64_t pi[5][8];   // pi[X][Y] = kTau[(3Y + X) % 5]: slot of plane X that lands at position (X, Y) after pi
  alignas(64) uint64_t rot_m1[8], rot_p1[8];  // parity vector C -> C[x-1], C[x+1] in plane slot order
  alignas(64) uint64_t natural[8];            // plane slot order -> x = 0..4 in slots 0..4 (for the 168-byte output)
  alignas(64) uint64_t rc[12][8];
};

const Avx512Consts &avx512_consts() {
  static const Avx512Consts c = [] {
    Avx512Consts k{};
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) k.rho[y][kTau[x]] = kRho[x + 5 * y];
    for (int X = 0; X < 5; X++)
      for (int Y = 0; Y < 5; Y++) k.pi[X][Y] = uint64_t(kTau[(3 * Y + X) % 5]);
    for (int s = 0; s < 8; s++) k.rot_m1[s] = k.rot_p1[s] = uint64_t(s);
    for (int x = 0; x < 5; x++) {
      k.rot_m1[kTau[x]] = uint64_t(kTau[(x + 4) % 5]);
      k.rot_p1[kTau[x]] = uint64_t(kTau[(x + 1) % 5]);
      k.natural[x] = uint64_t(kTau[x]);
    }
    for (int r = 0; r < 12; r++) k.rc[r][kTau[0]] = kRC[r];
    return k;
  }();
  return c;
}

__attribute__((target("avx512f,avx512vl"))) void squeeze_avx512(uint64_t s[25], uint8_t *out, uint64_t nblocks) {
  const Avx512Consts &k = avx512_consts();
  const __mmask8 m5 = 0x1f, mx4 = uint8_t(1u << kTau[4]);
  alignas(64) uint64_t tmp[5][8] = {};
  for (int y = 0; y < 5; y++)
    for (int x = 0; x < 5; x++) tmp[y][kTau[x]] = s[x + 5 * y];
  __m512i P0 = _mm512_load_si512(tmp[0]), P1 = _mm512_load_si512(tmp[1]), P2 = _mm512_load_si512(tmp[2]), P3 = _mm512_load_si512(tmp[3]),
          P4 = _mm512_load_si512(tmp[4]);
  const __m512i rho0 = _mm512_load_si512(k.rho[0]), rho1 = _mm512_load_si512(k.rho[1]), rho2 = _mm512_load_si512(k.rho[2]),
                rho3 = _mm512_load_si512(k.rho[3]), rho4 = _mm512_load_si512(k.rho[4]);
  const __m512i pi0 = _mm512_load_si512(k.pi[0]), pi1 = _mm512_load_si512(k.pi[1]), pi2 = _mm512_load_si512(k.pi[2]),
                pi3 = _mm512_load_si512(k.pi[3]), pi4 = _mm512_load_si512(k.pi[4]);
  const __m512i im1 = _mm512_load_si512(k.rot_m1), ip1 = _mm512_load_si512(k.rot_p1), nat = _mm512_load_si512(k.natural);
  alignas(64) uint64_t r4[8];
  // slot kTau[4] of plane P <- r4[Y], straight from memory (written as asm so that the compiler cannot turn the store + loads back
  // into register shuffles, which is exactly the port pressure this avoids)
#define CHPIR_MERGE_X4(P, Y) asm("vpbroadcastq %2, %0 %{%1%}" : "+v"(P) : "Yk"(mx4), "m"(r4[Y]), "m"(*(const __m512i *)r4))
  for (uint64_t blk = 0; blk < nblocks; blk++) {
    for (int r = 0; r < 12; r++) {
      // theta
      const __m512i C = _mm512_ternarylogic_epi64(_mm512_ternarylogic_epi64(P0, P1, P2, 0x96), P3, P4, 0x96);
      const __m512i Cm = _mm512_permutexvar_epi64(im1, C);
      const __m512i Cp = _mm512_rol_epi64(_mm512_permutexvar_epi64(ip1, C), 1);
      P0 = _mm512_ternarylogic_epi64(P0, Cm, Cp, 0x96);
      P1 = _mm512_ternarylogic_epi64(P1, Cm, Cp, 0x96);
      P2 = _mm512_ternarylogic_epi64(P2, Cm, Cp, 0x96);
      P3 = _mm512_ternarylogic_epi64(P3, Cm, Cp, 0x96);
      P4 = _mm512_ternarylogic_epi64(P4, Cm, Cp, 0x96);
      // rho, then pi as a slot permutation inside each plane: Q[X] slot Y = B[X][Y]
      const __m512i Q0 = _mm512_permutexvar_epi64(pi0, _mm512_rolv_epi64(P0, rho0));
      const __m512i Q1 = _mm512_permutexvar_epi64(pi1, _mm512_rolv_epi64(P1, rho1));
      const __m512i Q2 = _mm512_permutexvar_epi64(pi2, _mm512_rolv_epi64(P2, rho2));
      const __m512i Q3 = _mm512_permutexvar_epi64(pi3, _mm512_rolv_epi64(P3, rho3));
      const __m512i Q4 = _mm512_permutexvar_epi64(pi4, _mm512_rolv_epi64(P4, rho4));
      // chi is register-wise in this representation: R[X] slot Y = A'[X][Y]
      const __m512i R4 = _mm512_ternarylogic_epi64(Q4, Q0, Q1, 0xD2);
      asm("vmovdqa64 %1, %0" : "=m"(*(__m512i *)r4) : "v"(R4));
      const __m512i R0 = _mm512_ternarylogic_epi64(Q0, Q1, Q2, 0xD2);
      const __m512i R1 = _mm512_ternarylogic_epi64(Q1, Q2, Q3, 0xD2);
      const __m512i R2 = _mm512_ternarylogic_epi64(Q2, Q3, Q4, 0xD2);
      const __m512i R3 = _mm512_ternarylogic_epi64(Q3, Q4, Q0, 0xD2);
      // transpose back to planes: 128-bit lane j of lo = (R[even][2j], R[odd][2j]), of hi = (R[even][2j+1], R[odd][2j+1])
      const __m512i lo01 = _mm512_unpacklo_epi64(R0, R1), hi01 = _mm512_unpackhi_epi64(R0, R1);
      const __m512i lo23 = _mm512_unpacklo_epi64(R2, R3), hi23 = _mm512_unpackhi_epi64(R2, R3);
      P0 = _mm512_shuffle_i64x2(lo01, lo23, 0x00);
      P1 = _mm512_shuffle_i64x2(hi01, hi23, 0x00);
      P2 = _mm512_shuffle_i64x2(lo01, lo23, 0x11);
      P3 = _mm512_shuffle_i64x2(hi01, hi23, 0x11);
      P4 = _mm512_shuffle_i64x2(lo01, lo23, 0x22);
      CHPIR_MERGE_X4(P0, 0);
      P0 = _mm512_xor_si512(P0, _mm512_load_si512(k.rc[r]));  // iota on lane (0, 0); P0 is the plane that is ready first
      CHPIR_MERGE_X4(P1, 1);
      CHPIR_MERGE_X4(P2, 2);
      CHPIR_MERGE_X4(P3, 3);
      CHPIR_MERGE_X4(P4, 4);
    }
    // 168 bytes = lanes 0..20: planes 0..3 whole, lane (0, 4) (slot kTau[0] = 0 of plane 4)
    uint8_t *o = out + blk * kXofRate;
    _mm512_mask_storeu_epi64(o, m5, _mm512_permutexvar_epi64(nat, P0));
    _mm512_mask_storeu_epi64(o + 40, m5, _mm512_permutexvar_epi64(nat, P1));
    _mm512_mask_storeu_epi64(o + 80, m5, _mm512_permutexvar_epi64(nat, P2));
    _mm512_mask_storeu_epi64(o + 120, m5, _mm512_permutexvar_epi64(nat, P3));
    _mm512_mask_storeu_epi64(o + 160, 0x01, P4);
  }
#undef CHPIR_MERGE_X4
  _mm512_mask_storeu_epi64(s + 0, m5, _mm512_permutexvar_epi64(nat, P0));
  _mm512_mask_storeu_epi64(s + 5, m5, _mm512_permutexvar_epi64(nat, P1));
  _mm512_mask_storeu_epi64(s + 10, m5, _mm512_permutexvar_epi64(nat, P2));
  _mm512_mask_storeu_epi64(s + 15, m5, _mm512_permutexvar_epi64(nat, P3));
  _mm512_mask_storeu_epi64(s + 20, m5, _mm512_permutexvar_epi64(nat, P4));
}

// ---- EVEX-128: the scalar formulation (one lane per register, pi as a renaming) on xmm registers with AVX-512VL instructions.
// The BMI2 code is front-end bound on the measured Xeon (~225 instructions per round at ~6 per cycle, almost half of them moves and
// spills: 16 GPRs for 25 lanes, two-operand XOR).  Here every operation is three-operand, there are 32 registers, a five-input
// parity is two vpternlogq, theta's  a ^ c[x-1] ^ rol(c[x+1])  is one, chi's  a ^ (~b & c)  is one, and rho is one vprolq:
// ~90 instructions per round, and 128-bit operations do not lower the core clock the way 512-bit ones do.
#define CHPIR_X3(a, b, c) _mm_ternarylogic_epi64(a, b, c, 0x96)
#define CHPIR_CHI(a, b, c) _mm_ternarylogic_epi64(a, b, c, 0xD2)
#define CHPIR_ROUND_X(A, E, RC)                                                                                             \
  do {                                                                                                                      \
    const __m128i c0 = CHPIR_X3(CHPIR_X3(A[0], A[5], A[10]), A[15], A[20]);                                                 \
    const __m128i c1 = CHPIR_X3(CHPIR_X3(A[1], A[6], A[11]), A[16], A[21]);                                                 \
    const __m128i c2 = CHPIR_X3(CHPIR_X3(A[2], A[7], A[12]), A[17], A[22]);                                                 \
    const __m128i c3 = CHPIR_X3(CHPIR_X3(A[3], A[8], A[13]), A[18], A[23]);                                                 \
    const __m128i c4 = CHPIR_X3(CHPIR_X3(A[4], A[9], A[14]), A[19], A[24]);                                                 \
    const __m128i r0 = _mm_rol_epi64(c0, 1), r1 = _mm_rol_epi64(c1, 1), r2 = _mm_rol_epi64(c2, 1), r3 = _mm_rol_epi64(c3, 1), \
                  r4 = _mm_rol_epi64(c4, 1);                                                                                \
    /* lane (x, y) ^ d[x] with d[x] = c[x-1] ^ rol(c[x+1], 1) */                                                           \
    __m128i b0, b1, b2, b3, b4;                                                                                             \
    b0 = CHPIR_X3(A[0], c4, r1), b1 = _mm_rol_epi64(CHPIR_X3(A[6], c0, r2), 44), b2 = _mm_rol_epi64(CHPIR_X3(A[12], c1, r3), 43); \
    b3 = _mm_rol_epi64(CHPIR_X3(A[18], c2, r4), 21), b4 = _mm_rol_epi64(CHPIR_X3(A[24], c3, r0), 14);                       \
    E[0] = _mm_xor_si128(CHPIR_CHI(b0, b1, b2), _mm_cvtsi64_si128((long long)(RC)));                                        \
    E[1] = CHPIR_CHI(b1, b2, b3), E[2] = CHPIR_CHI(b2, b3, b4), E[3] = CHPIR_CHI(b3, b4, b0), E[4] = CHPIR_CHI(b4, b0, b1); \
    b0 = _mm_rol_epi64(CHPIR_X3(A[3], c2, r4), 28), b1 = _mm_rol_epi64(CHPIR_X3(A[9], c3, r0), 20);                         \
    b2 = _mm_rol_epi64(CHPIR_X3(A[10], c4, r1), 3), b3 = _mm_rol_epi64(CHPIR_X3(A[16], c0, r2), 45);                        \
    b4 = _mm_rol_epi64(CHPIR_X3(A[22], c1, r3), 61);                                                                        \
    E[5] = CHPIR_CHI(b0, b1, b2), E[6] = CHPIR_CHI(b1, b2, b3), E[7] = CHPIR_CHI(b2, b3, b4), E[8] = CHPIR_CHI(b3, b4, b0); \
    E[9] = CHPIR_CHI(b4, b0, b1);                                                                                           \
    b0 = _mm_rol_epi64(CHPIR_X3(A[1], c0, r2), 1), b1 = _mm_rol_epi64(CHPIR_X3(A[7], c1, r3), 6);                           \
    b2 = _mm_rol_epi64(CHPIR_X3(A[13], c2, r4), 25), b3 = _mm_rol_epi64(CHPIR_X3(A[19], c3, r0), 8);                        \
    b4 = _mm_rol_epi64(CHPIR_X3(A[20], c4, r1), 18);                                                                        \
    E[10] = CHPIR_CHI(b0, b1, b2), E[11] = CHPIR_CHI(b1, b2, b3), E[12] = CHPIR_CHI(b2, b3, b4), E[13] = CHPIR_CHI(b3, b4, b0); \
    E[14] = CHPIR_CHI(b4, b0, b1);                                                                                          \
    b0 = _mm_rol_epi64(CHPIR_X3(A[4], c3, r0), 27), b1 = _mm_rol_epi64(CHPIR_X3(A[5], c4, r1), 36);                         \
    b2 = _mm_rol_epi64(CHPIR_X3(A[11], c0, r2), 10), b3 = _mm_rol_epi64(CHPIR_X3(A[17], c1, r3), 15);                       \
    b4 = _mm_rol_epi64(CHPIR_X3(A[23], c2, r4), 56);                                                                        \
    E[15] = CHPIR_CHI(b0, b1, b2), E[16] = CHPIR_CHI(b1, b2, b3), E[17] = CHPIR_CHI(b2, b3, b4), E[18] = CHPIR_CHI(b3, b4, b0); \
    E[19] = CHPIR_CHI(b4, b0, b1);                                                                                          \
    b0 = _mm_rol_epi64(CHPIR_X3(A[2], c1, r3), 62), b1 = _mm_rol_epi64(CHPIR_X3(A[8], c2, r4), 55);                         \
    b2 = _mm_rol_epi64(CHPIR_X3(A[14], c3, r0), 39), b3 = _mm_rol_epi64(CHPIR_X3(A[15], c4, r1), 41);                       \
    b4 = _mm_rol_epi64(CHPIR_X3(A[21], c0, r2), 2);                                                                         \
    E[20] = CHPIR_CHI(b0, b1, b2), E[21] = CHPIR_CHI(b1, b2, b3), E[22] = CHPIR_CHI(b2, b3, b4), E[23] = CHPIR_CHI(b3, b4, b0); \
    E[24] = CHPIR_CHI(b4, b0, b1);                                                                                          \
  } while (0)

__attribute__((target("avx512f,avx512vl"))) void squeeze_evex128(uint64_t s[25], uint8_t *out, uint64_t nblocks) {
  __m128i a[25], e[25];
  for (int i = 0; i < 25; i++) a[i] = _mm_cvtsi64_si128((long long)s[i]);
  for (uint64_t blk = 0; blk < nblocks; blk++) {
    for (int r = 0; r < 12; r += 2) {
      CHPIR_ROUND_X(a, e, kRC[r]);
      CHPIR_ROUND_X(e, a, kRC[r + 1]);
    }
    uint8_t *o = out + blk * kXofRate;
    for (int i = 0; i < 21; i++) _mm_storel_epi64(reinterpret_cast<__m128i *>(o + 8 * i), a[i]);
  }
  for (int i = 0; i < 25; i++) s[i] = (uint64_t)_mm_cvtsi128_si64(a[i]);
}
#undef CHPIR_X3
#undef CHPIR_CHI

bool have_avx512() {
  static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl");
  return ok;
}
bool have_bmi2() {
  static const bool ok = __builtin_cpu_supports("bmi") && __builtin_cpu_supports("bmi2");
  return ok;
}
#else
bool have_avx512() { return false; }
bool have_bmi2() { return false; }
#endif

bool run_impl(int impl, uint64_t s[25], uint8_t *out, uint64_t nblocks) {
  switch (impl) {
    case kXofScalar: squeeze_scalar(s, out, nblocks); return true;
#if defined(__x86_64__)
    case kXofBmi2:
      if (!have_bmi2()) return false;
      squeeze_bmi2(s, out, nblocks);
      return true;
    case kXofAvx512:
      if (!have_avx512()) return false;
      squeeze_avx512(s, out, nblocks);
      return true;
    case kXofEvex128:
      if (!have_avx512()) return false;
      squeeze_evex128(s, out, nblocks);
      return true;
#endif
    default: return false;
  }
}

// impl 0: whichever implementation walks the chain fastest on this CPU (measured once, ~1 ms each)
int best_impl() {
  static const int best = [] {
    int win = kXofScalar;
    double win_t = 1e30;
    constexpr uint64_t kCal = 4096;
    std::vector<uint8_t> buf(kCal * kXofRate);
    for (int impl : {kXofScalar, kXofBmi2, kXofAvx512, kXofEvex128}) {
      uint64_t st[25] = {1, 2, 3};
      if (!run_impl(impl, st, buf.data(), kCal)) continue;  // availability + warm-up (~0.5 ms: past the frequency-licence transition)
      double t = 1e30;
      for (int rep = 0; rep < 5; rep++) {
        const auto t0 = std::chrono::steady_clock::now();
        run_impl(impl, st, buf.data(), kCal);
        t = std::min(t, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
      }
      if (t < win_t) win_t = t, win = impl;
    }
    return win;
  }();
  return best;
}

}  // namespace

void host_xof_init(HostXof *x, const uint8_t seed[32]) {
  // absorb(seed) ; finalize::<0x1F>() : pad10*1 over the 168-byte rate; the first permutation happens in the squeeze
  std::memset(x->s, 0, sizeof x->s);
  std::memcpy(x->s, seed, 32);
  x->s[4] ^= 0x1full;
  x->s[20] ^= 0x80ull << 56;
}

bool host_xof_squeeze_blocks(HostXof *x, uint8_t *out, uint64_t nblocks, int impl) {
  return run_impl(impl == kXofAuto ? best_impl() : impl, x->s, out, nblocks);
}

void host_xof_skip_blocks(HostXof *x, uint64_t nblocks) {
  // the permutation without the 168-byte copy is not worth a second code path: squeeze into a scratch block
  uint8_t scratch[64 * kXofRate];
  const int impl = best_impl();
  while (nblocks) {
    const uint64_t n = std::min<uint64_t>(nblocks, 64);
    run_impl(impl, x->s, scratch, n);
    nblocks -= n;
  }
}

const char *host_xof_impl_name() {
  switch (best_impl()) {
    case kXofAvx512: return "avx512";
    case kXofEvex128: return "evex128";
    case kXofBmi2: return "bmi2";
    default: return "scalar";
  }
}

}  // namespace chpir
