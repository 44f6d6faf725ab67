/* dropin_test.c -- a plain-C caller of include/chalamet_b200.h, linked against libchalamet_b200.so: what the reference's
 * `gpu`-feature build would do through its extern "C" block, without Python or ctypes in between.
 *
 *   Server::setup(seed, db)  (chalametpir_server/src/server.rs:103)  -> chpir_cluster_server_setup_from_db
 *   Server::respond(query)   (server.rs:184)                         -> chpir_cluster_server_respond
 *
 * Checks, with arithmetic done here in C (no oracle involved):
 *   - the hint header and row 0 of the hint against  A[0] . D  with A[0] from chpir_host_generate_from_seed (the head of the
 *     TurboSHAKE128 stream, matrix.rs:541-558) and D from chpir_encode_kv_database with the same filter seed;
 *   - a response against the exact combination of rows of D the query selects;
 *   - the reference's error variants for malformed queries.
 * usage: dropin_test [n_gpus]      (compiled by __graft_entry__.build(); run by tests/test_gpu_cluster.py) */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "chalamet_b200.h"

#define CHECK(call)                                                                             \
  do {                                                                                          \
    int rc_ = (call);                                                                           \
    if (rc_ != CHPIR_OK) {                                                                      \
      fprintf(stderr, "%s:%d: %s -> %s (%s)\n", __FILE__, __LINE__, #call, chpir_strerror(rc_), \
              chpir_last_cuda_error());                                                         \
      return 1;                                                                                 \
    }                                                                                           \
  } while (0)

static uint64_t splitmix(uint64_t *s) {
  uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

int main(int argc, char **argv) {
  const int n_gpus = argc > 1 ? atoi(argv[1]) : 1;
  enum { N_ENTRIES = 700, KEY_LEN = 16, VAL_LEN = 40, LWE = 64, ARITY = 3 };
  uint8_t seed[CHPIR_SEED_BYTE_LEN];
  for (int i = 0; i < (int)CHPIR_SEED_BYTE_LEN; i++) seed[i] = (uint8_t)(3 * i + 11);

  /* a synthetic key-value database, flattened as blob + offsets */
  uint8_t *keys = malloc((size_t)N_ENTRIES * KEY_LEN), *vals = malloc((size_t)N_ENTRIES * VAL_LEN);
  uint64_t *koff = malloc((N_ENTRIES + 1) * sizeof(uint64_t)), *voff = malloc((N_ENTRIES + 1) * sizeof(uint64_t));
  uint64_t s = 42;
  for (int i = 0; i < N_ENTRIES; i++) {
    for (int j = 0; j < KEY_LEN; j++) keys[i * KEY_LEN + j] = (uint8_t)splitmix(&s);
    memcpy(keys + i * KEY_LEN, &i, sizeof i); /* distinct by construction */
    for (int j = 0; j < VAL_LEN; j++) vals[i * VAL_LEN + j] = (uint8_t)splitmix(&s);
  }
  for (int i = 0; i <= N_ENTRIES; i++) koff[i] = (uint64_t)i * KEY_LEN, voff[i] = (uint64_t)i * VAL_LEN;

  uint32_t b = 0;
  uint64_t K = 0, N = 0;
  CHECK(chpir_find_mat_elem_bit_len(N_ENTRIES, &b));
  CHECK(chpir_db_matrix_shape(ARITY, N_ENTRIES, VAL_LEN, b, &K, &N));
  const uint64_t filter_rng = 7;
  uint32_t *D = malloc(K * N * sizeof(uint32_t));
  uint8_t fparams_host[CHPIR_FILTER_PARAM_BYTE_LEN], fparams[CHPIR_FILTER_PARAM_BYTE_LEN];
  CHECK(chpir_encode_kv_database(ARITY, N_ENTRIES, keys, koff, vals, voff, b, CHPIR_SERVER_SETUP_MAX_ATTEMPT_COUNT, &filter_rng, D, fparams_host));

  chpir_cluster *cluster = NULL;
  CHECK(chpir_cluster_create(n_gpus, NULL, &cluster));
  int have = 0;
  CHECK(chpir_cluster_size(cluster, &have));
  if (have != n_gpus) return fprintf(stderr, "cluster size %d != %d\n", have, n_gpus), 1;

  chpir_setup_opts opts;
  memset(&opts, 0, sizeof opts);
  opts.lwe_rows = LWE;
  opts.batch_tc = 1;
  const size_t hint_cap = 8 + (size_t)LWE * N * 4;
  uint8_t *hint = malloc(hint_cap);
  size_t hint_len = 0;
  chpir_cluster_server *srv = NULL;
  CHECK(chpir_cluster_server_setup_from_db(cluster, ARITY, seed, N_ENTRIES, keys, koff, vals, voff, &filter_rng, &opts, hint, hint_cap, &hint_len,
                                           fparams, &srv));
  if (hint_len != hint_cap) return fprintf(stderr, "hint length %zu != %zu\n", hint_len, hint_cap), 1;
  if (memcmp(fparams, fparams_host, sizeof fparams) != 0) return fprintf(stderr, "filter parameter bytes differ\n"), 1;
  uint32_t hdr[2];
  memcpy(hdr, hint, 8);
  if (hdr[0] != LWE || hdr[1] != N) return fprintf(stderr, "hint header %u x %u\n", hdr[0], hdr[1]), 1;

  /* hint row 0 == A[0] . D mod 2^32 */
  uint32_t *a0 = malloc(K * sizeof(uint32_t));
  CHECK(chpir_host_generate_from_seed(seed, LWE, K, 0, 1, 0, a0));
  for (uint64_t c = 0; c < N; c++) {
    uint32_t acc = 0;
    for (uint64_t k = 0; k < K; k++) acc += a0[k] * D[k * N + c];
    uint32_t got;
    memcpy(&got, hint + 8 + 4 * c, 4);
    if (got != acc) return fprintf(stderr, "hint[0][%llu] = %u, want %u\n", (unsigned long long)c, got, acc), 1;
  }

  /* respond: a sparse full-range query, checked against the rows of D it selects */
  const size_t qlen = 8 + 4 * K, rcap = 8 + 4 * N;
  uint8_t *query = calloc(1, qlen), *resp = malloc(rcap);
  uint32_t *q = malloc(K * sizeof(uint32_t));
  memset(q, 0, K * sizeof(uint32_t));
  for (int i = 0; i < 9; i++) q[splitmix(&s) % K] = (uint32_t)splitmix(&s);
  q[K - 1] = 0xffffffffu;
  const uint32_t qhdr[2] = {1u, (uint32_t)K};
  memcpy(query, qhdr, 8);
  memcpy(query + 8, q, 4 * K);
  size_t rlen = 0;
  for (int rep = 0; rep < 3; rep++) {
    memset(resp, 0xee, rcap);
    CHECK(chpir_cluster_server_respond(srv, query, qlen, resp, rcap, &rlen));
    if (rlen != rcap) return fprintf(stderr, "response length %zu != %zu\n", rlen, rcap), 1;
    memcpy(hdr, resp, 8);
    if (hdr[0] != 1 || hdr[1] != N) return fprintf(stderr, "response header %u x %u\n", hdr[0], hdr[1]), 1;
    for (uint64_t c = 0; c < N; c++) {
      uint32_t acc = 0;
      for (uint64_t k = 0; k < K; k++)
        if (q[k]) acc += q[k] * D[k * N + c];
      uint32_t got;
      memcpy(&got, resp + 8 + 4 * c, 4);
      if (got != acc) return fprintf(stderr, "resp[%llu] = %u, want %u\n", (unsigned long long)c, got, acc), 1;
    }
  }
  /* the batch call: 8 copies of the query take the tensor-core route; every response must equal the single one */
  {
    enum { NQ = 8 };
    const uint8_t *qs[NQ];
    size_t lens[NQ];
    for (int i = 0; i < NQ; i++) qs[i] = query, lens[i] = qlen;
    uint8_t *many = malloc(NQ * rcap);
    CHECK(chpir_cluster_server_respond_batch(srv, qs, lens, NQ, many, rcap));
    for (int i = 0; i < NQ; i++)
      if (memcmp(many + i * rcap, resp, rcap) != 0) return fprintf(stderr, "batched response %d differs\n", i), 1;
    free(many);
  }
  /* error behaviour of Matrix::from_bytes / the dimension check (matrix.rs:973-1010, :329-331) */
  if (chpir_cluster_server_respond(srv, query, 8, resp, rcap, &rlen) != CHPIR_ERR_FAILED_TO_DESERIALIZE_MATRIX_FROM_BYTES) return fprintf(stderr, "short query accepted\n"), 1;
  if (chpir_cluster_server_respond(srv, query, qlen - 4, resp, rcap, &rlen) != CHPIR_ERR_FAILED_TO_DESERIALIZE_MATRIX_FROM_BYTES) return fprintf(stderr, "truncated query accepted\n"), 1;
  {
    uint32_t bad[2] = {1u, (uint32_t)K - 1};
    memcpy(query, bad, 8);
    if (chpir_cluster_server_respond(srv, query, qlen - 4, resp, rcap, &rlen) != CHPIR_ERR_INCOMPATIBLE_DIMENSION_FOR_ROW_VECTOR_TRANSPOSED_MATRIX_MULTIPLICATION)
      return fprintf(stderr, "wrong-dimension query accepted\n"), 1;
  }
  if (chpir_cluster_server_respond(srv, query, qlen, resp, 4, &rlen) == CHPIR_OK) return fprintf(stderr, "tiny response buffer accepted\n"), 1;

  chpir_cluster_server_info info;
  CHECK(chpir_cluster_server_get_info(srv, &info));
  printf("dropin_test ok: %u GPU(s), K=%llu N=%u b=%u, hint %zu B, setup %.3f s (hint gather %.4f s, nccl=%u v%u)\n", info.n_gpus,
         (unsigned long long)info.rows_k, info.cols_n, info.mat_elem_bit_len, hint_len, info.setup_total_s, info.hint_gather_s, info.gather_uses_nccl,
         info.nccl_version);
  chpir_cluster_server_destroy(srv);
  chpir_cluster_destroy(cluster);
  free(keys), free(vals), free(koff), free(voff), free(D), free(hint), free(a0), free(query), free(resp), free(q);
  return 0;
}
