/*
 * chalamet_b200.h -- C ABI of the B200-native ChalametPIR server hot path.
 *
 * This is the drop-in boundary behind the reference crate's `gpu` cargo feature
 * (chalametpir_server/Cargo.toml:31-32, chalametpir_server/src/lib.rs:75-76).  The reference has no FFI:
 * its plugin point is five Rust free functions over vulkano types
 * (chalametpir_server/src/gpu/gpu_utils.rs:25,81,140,156,222) called from the `gpu` variant of
 * `Server::setup` (chalametpir_server/src/server.rs:102-167).  The entry points below are what a Rust
 * `extern "C"` block in that module would bind instead (see INTEGRATION.md for the stub), plus the host-side
 * helpers the Rust crate already owns (filter construction / row encoding) so that C++ / Python callers can
 * drive the complete `Server::setup(seed, db)` / `Server::respond(query)` flow.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns a `chpir_status` value (0 = ok) and never throws/aborts;
 *   - all matrices are row-major u32; "wire format" = Matrix::to_bytes (chalametpir_common/src/matrix.rs:947-971):
 *     rows LE32 | cols LE32 | elems LE32 row-major;
 *   - caller owns every output buffer; opaque handles own device memory and streams;
 *   - chpir_server_respond* are thread-safe and re-entrant on one handle (the reference shares `Arc<Server>`
 *     across tasks, chalametpir_server/examples/server.rs:45,55,85).
 */
#ifndef CHALAMET_B200_H
#define CHALAMET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CHPIR_API __attribute__((visibility("default")))
#else
#define CHPIR_API
#endif

#define CHPIR_LWE_DIMENSION 1774u     /* chalametpir_common/src/params.rs:1  */
#define CHPIR_SEED_BYTE_LEN 32u       /* chalametpir_common/src/params.rs:5  */
#define CHPIR_FILTER_PARAM_BYTE_LEN 68u /* chalametpir_common/src/binary_fuse_filter.rs:462-486 (64-bit usize) */
#define CHPIR_SERVER_SETUP_MAX_ATTEMPT_COUNT 100u /* chalametpir_common/src/params.rs:10 */

/* Status codes.  1..49 map one-to-one onto existing ChalametPIRError variants
 * (chalametpir_common/src/error.rs:8-50); 100+ are new device-side failures a Rust shim would map onto
 * new `Cuda*` variants (the reference maps every Vulkan failure to a unit variant the same way,
 * gpu_utils.rs:26-72). */
typedef enum chpir_status {
  CHPIR_OK = 0,
  CHPIR_ERR_INVALID_MATRIX_DIMENSION = 1,
  CHPIR_ERR_INCOMPATIBLE_DIMENSION_FOR_MATRIX_MULTIPLICATION = 2,
  CHPIR_ERR_INCOMPATIBLE_DIMENSION_FOR_ROW_VECTOR_TRANSPOSED_MATRIX_MULTIPLICATION = 3,
  CHPIR_ERR_FAILED_TO_DESERIALIZE_MATRIX_FROM_BYTES = 4,
  CHPIR_ERR_EMPTY_KV_DATABASE = 5,
  CHPIR_ERR_EXHAUSTED_ALL_ATTEMPTS_TO_BUILD_3_WISE_XOR_FILTER = 6,
  CHPIR_ERR_EXHAUSTED_ALL_ATTEMPTS_TO_BUILD_4_WISE_XOR_FILTER = 7,
  CHPIR_ERR_ROW_NOT_DECODABLE = 8,
  CHPIR_ERR_DECODED_ROW_NOT_PREPENDED_WITH_DIGEST_OF_KEY = 9,
  CHPIR_ERR_FAILED_TO_DESERIALIZE_FILTER_FROM_BYTES = 10,
  CHPIR_ERR_KV_DATABASE_SIZE_TOO_LARGE = 11,
  CHPIR_ERR_INVALID_HINT_MATRIX = 12,
  CHPIR_ERR_ARITHMETIC_OVERFLOW_ADDING_QUERY_INDICATOR = 13,
  CHPIR_ERR_UNSUPPORTED_ARITY_FOR_BINARY_FUSE_FILTER = 14,
  CHPIR_ERR_INVALID_RESPONSE_VECTOR = 15,
  CHPIR_ERR_IMPOSSIBLE_ENCODED_DB_MATRIX_ELEMENT_BIT_LENGTH = 16,
  CHPIR_ERR_PENDING_QUERY_EXISTS_FOR_KEY = 17,
  CHPIR_ERR_PENDING_QUERY_DOES_NOT_EXIST_FOR_KEY = 18,
  CHPIR_ERR_INVALID_ARGUMENT = 50,
  CHPIR_ERR_BUFFER_TOO_SMALL = 51,
  CHPIR_ERR_IO_FAILED = 52,
  CHPIR_ERR_INVALID_SAVED_SERVER = 53,
  CHPIR_ERR_CUDA_DEVICE_NOT_FOUND = 100,
  CHPIR_ERR_CUDA_ALLOCATION_FAILED = 101,
  CHPIR_ERR_CUDA_TRANSFER_FAILED = 102,
  CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED = 103,
  CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED = 104,
  CHPIR_ERR_CUDA_UNSUPPORTED_DEVICE = 105,
  CHPIR_ERR_CUDA_PEER_ACCESS_UNAVAILABLE = 106, /* a cluster needs peer access (NVLink / NVSwitch) between all of its GPUs */
  CHPIR_ERR_NCCL_FAILED = 107,                  /* libnccl.so.2 could not be loaded, or an NCCL call failed */
  CHPIR_ERR_HOST_ALLOCATION_FAILED = 110
} chpir_status;

typedef struct chpir_ctx chpir_ctx;       /* one GPU: device ordinal, streams, scratch pools           */
typedef struct chpir_server chpir_server; /* server.rs:15-21 `Server`: packed D resident in HBM (a column slice) */

CHPIR_API const char *chpir_strerror(int status);
/* Last CUDA error string seen on this thread (diagnostics only). */
CHPIR_API const char *chpir_last_cuda_error(void);

/* ---- device context: replaces gpu_utils::setup_gpu (gpu_utils.rs:25-79) ------------------------------ */
CHPIR_API int chpir_device_count(int *count);
CHPIR_API int chpir_ctx_create(int device_ordinal, chpir_ctx **out);
CHPIR_API void chpir_ctx_destroy(chpir_ctx *ctx);
/* Frees the LWE matrix a setup with chpir_setup_opts.a_cache = 1 left resident in the ctx (no-op if there is none);
 * *bytes_freed (may be NULL) receives its size. */
CHPIR_API int chpir_ctx_drop_a_cache(chpir_ctx *ctx, uint64_t *bytes_freed);

/* Page-locked host buffers for queries and responses: chpir_server_respond* accept any host pointer, but only page-locked memory
 * is copied by DMA at the full PCIe rate (a 4.7 MB query moves in 90 us instead of ~500 us from pageable memory). */
CHPIR_API int chpir_host_alloc(size_t bytes, void **out);
CHPIR_API void chpir_host_free(void *p);
/* Column-sharded serving: rank r uploads only words [k0, k1) of every query of a batch over its own PCIe link
 * before the slices are all-gathered over NVLink.  One strided DMA for the whole batch: `rows` rows of `width_bytes` from page-locked
 * host memory (row pitch src_pitch) into device memory (row pitch dst_pitch), asynchronously on `cuda_stream`. */
CHPIR_API int chpir_upload_rows(void *dst_device, size_t dst_pitch, const void *src_host, size_t src_pitch, size_t width_bytes, size_t rows,
                      void *cuda_stream);

/* ---- host side that stays on the host (north_star); mirrors chalametpir_common --------------------- */
/* server.rs:193-218 find_encoded_db_matrix_element_bit_length */
CHPIR_API int chpir_find_mat_elem_bit_len(uint64_t db_entry_count, uint32_t *mat_elem_bit_len);
/* Shape of D for a DB: rows K = num_fingerprints (binary_fuse_filter.rs:52-67 / :261-276),
 * cols N = ceil((256 + 8*max_value_byte_len + 8) / b) (matrix.rs:699-700). */
CHPIR_API int chpir_db_matrix_shape(uint32_t arity, uint64_t db_entry_count, uint64_t max_value_byte_len, uint32_t mat_elem_bit_len,
                          uint64_t *rows_k, uint64_t *cols_n);
/* Matrix::from_kv_database::<ARITY> (matrix.rs:633-648, :687-755, :819-894) + BinaryFuseFilter::to_bytes.
 * Keys/values are passed flattened: blob + (n+1) offsets.  d_out (rows_k*cols_n u32, caller allocated) receives D.
 * filter_seed_rng: NULL -> OS entropy (reference behaviour: ChaCha20Rng::from_os_rng, binary_fuse_filter.rs:100,309);
 * otherwise a 64-bit seed for a deterministic stream of candidate filter seeds (tests / reproducible vectors). */
CHPIR_API int chpir_encode_kv_database(uint32_t arity, uint64_t db_entry_count, const uint8_t *key_blob, const uint64_t *key_offsets,
                             const uint8_t *value_blob, const uint64_t *value_offsets, uint32_t mat_elem_bit_len,
                             uint32_t max_attempt_count, const uint64_t *filter_seed_rng, uint32_t *d_out,
                             uint8_t filter_params_out[CHPIR_FILTER_PARAM_BYTE_LEN]);

/* The same matrix and filter parameters with the row encoding and the dependent fill done on the GPU (csrc/encode_dev.cu); D is
 * downloaded into d_out so that the two encoders can be compared byte for byte. */
CHPIR_API int chpir_encode_kv_database_device(chpir_ctx *ctx, uint32_t arity, uint64_t db_entry_count, const uint8_t *key_blob,
                                    const uint64_t *key_offsets, const uint8_t *value_blob, const uint64_t *value_offsets,
                                    uint32_t mat_elem_bit_len, uint32_t max_attempt_count, const uint64_t *filter_seed_rng, uint32_t *d_out,
                                    uint8_t filter_params_out[CHPIR_FILTER_PARAM_BYTE_LEN]);

/* ---- setup: replaces generate_from_seed + transfer_mat_to_device + mat_x_mat + mat_transpose + row_wise_compress
 *      (server.rs:115-156, gpu_utils.rs:81-281, shaders/mat_x_mat.glsl, shaders/mat_transpose.glsl,
 *      matrix.rs:98-205, :517-558, :1040-1059) ------------------------------------------------------------- */
typedef struct chpir_setup_opts {
  uint32_t lwe_rows;     /* 0 = CHPIR_LWE_DIMENSION; tests may shrink it                                   */
  uint32_t col_begin;    /* column slice [col_begin, col_begin+col_count) of D owned by this server (rank) */
  uint32_t col_count;    /* 0 = all columns                                                               */
  uint32_t gemm_variant; /* 0 = default (tcgen05 int8-limb GEMM); 1 = SIMT u32 reference kernel (debug)    */
  uint32_t skip_hint;    /* 1 = only make D resident (pack), no A expansion / GEMM                        */
  uint32_t batch_tc;     /* batched respond on the tensor cores (chpir_server_respond_device_tc): 0 = keep D's byte-limb
                            planes resident iff the hint GEMM built them anyway, 1 = always build and keep them,
                            2 = never keep them (saves 2*K*N bytes of HBM)                                */
  uint32_t a_expand;     /* where the LWE matrix A = generate_from_seed(lwe_rows, K, seed) (matrix.rs:541-558) is squeezed out of
                            TurboSHAKE128: CHPIR_A_EXPAND_AUTO (default), _HOST_PIPELINED or _DEVICE                  */
  uint32_t host_chunk_rows; /* host-pipelined mode: rows of A per pinned upload chunk; 0 = about 32 MB worth (tests shrink it) */
  uint32_t respond_coalesce; /* 1 = concurrent chpir_server_respond calls on this server are coalesced: whoever arrives while the
                            previous batch is on the GPU is answered by ONE launch (a grid.y GEMV for up to 5 queries, the
                            tensor-core limb GEMM for 6..128 when the limb planes are resident, see batch_tc).  A lone caller
                            still gets a batch of one with no added wait.  Costs 2 x 128 x K x 4 bytes of HBM staging.  */
  uint32_t db_encode;    /* chpir_server_setup_from_db only: where the rows of D are encoded and filled, CHPIR_DB_ENCODE_HOST
                            (default, north_star) or CHPIR_DB_ENCODE_DEVICE                                            */
  uint32_t a_cache;      /* 1 = keep A = generate_from_seed(lwe_rows, K, seed) resident in the ctx (row-major u32, 8.4 GB of HBM at
                            2^20 entries) and reuse it: A depends on the seed and on K only, not on the database, so a server that
                            re-runs setup after a database update with the same seed skips the serial XOF chain -- the whole of
                            the hint computation is then the tensor-core GEMM (milliseconds).  A setup with another seed,
                            lwe_rows or K replaces the cached matrix.  Same hint bytes either way.                      */
  uint32_t hint_on_device; /* 1 = leave this server's hint slice (lwe_rows x col_count u32, row-major) in HBM instead of downloading it:
                            hint_out may be NULL, *hint_len is 0, and chpir_server_hint_device returns the pointer.  This is what the
                            cluster layer gathers over NCCL / NVLink into the wire-format hint.                          */
} chpir_setup_opts;

/* chpir_setup_opts.db_encode.  Key digests and filter construction (peeling) always run on the host.
 *   HOST:   encode_kv_as_row + the dependent row fill of Matrix::from_kv_database run on the host (csrc/host_encode.cpp) and the
 *           K x N u32 matrix D (4.4 GB at 2^20 entries) is uploaded;
 *   DEVICE: the raw values (1.1 GB) are uploaded and D is built in HBM by csrc/encode_dev.cu in the dependency order the peeling
 *           implies -- the same D, byte for byte; on a cluster every GPU builds its own columns. */
#define CHPIR_DB_ENCODE_HOST 0u
#define CHPIR_DB_ENCODE_DEVICE 1u

/* chpir_setup_opts.a_expand.  The squeeze is ONE serial chain of Keccak-p[1600,12] permutations (49.8 M of them at
 * 2^20 entries); no mode changes a single byte of A or of the hint.
 *   AUTO (0, the default): whichever walker is faster -- HOST_PIPELINED on every machine measured so far (one host core: ~100 ns
 *                   per permutation, 5.3 s at 2^20 entries; one GPU warp: 1146 ns, 57 s), so that the stock call is the
 *                   seconds-scale setup.
 *   HOST_PIPELINED: one host core walks the chain (csrc/host_xof.cpp) into a ring of chunks that are uploaded and
 *                   multiplied panel by panel while the core keeps squeezing -- what the reference's own `gpu` feature does
 *                   with its CPU-side generate_from_seed + upload (server.rs:115-123), but overlapped.
 *   DEVICE:         one warp walks the chain on the GPU and writes the GEMM's byte planes directly (csrc/expand.cu): A never
 *                   leaves the device (north_star's "A expanded on device"), at the latency of a GPU warp. */
#define CHPIR_A_EXPAND_AUTO 0u
#define CHPIR_A_EXPAND_HOST_PIPELINED 1u
#define CHPIR_A_EXPAND_DEVICE 2u

/* d_host: K x N row-major u32 (Matrix elems), values < 2^mat_elem_bit_len.
 * hint_out receives the wire-format hint slice: header (lwe_rows, col_count) + lwe_rows*col_count u32; with the
 * default full slice this is byte-identical to the reference's `hint_bytes`.  hint_out may be NULL iff skip_hint. */
CHPIR_API int chpir_server_setup(chpir_ctx *ctx, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint32_t *d_host, uint64_t rows_k,
                       uint32_t cols_n, uint32_t mat_elem_bit_len, const chpir_setup_opts *opts, uint8_t *hint_out,
                       size_t hint_cap, size_t *hint_len, chpir_server **out);
/* Same, with D already resident in device memory (K x N row-major u32) -- used for device-generated synthetic D. */
CHPIR_API int chpir_server_setup_device(chpir_ctx *ctx, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint32_t *d_device, uint64_t rows_k,
                              uint32_t cols_n, uint32_t mat_elem_bit_len, const chpir_setup_opts *opts, uint8_t *hint_out,
                              size_t hint_cap, size_t *hint_len, chpir_server **out);
/* Complete Server::setup::<ARITY>(seed, db) (server.rs:103): host filter + encode, then chpir_server_setup. */
CHPIR_API int chpir_server_setup_from_db(chpir_ctx *ctx, uint32_t arity, const uint8_t seed[CHPIR_SEED_BYTE_LEN], uint64_t db_entry_count,
                               const uint8_t *key_blob, const uint64_t *key_offsets, const uint8_t *value_blob,
                               const uint64_t *value_offsets, const uint64_t *filter_seed_rng, const chpir_setup_opts *opts,
                               uint8_t *hint_out, size_t hint_cap, size_t *hint_len,
                               uint8_t filter_params_out[CHPIR_FILTER_PARAM_BYTE_LEN], chpir_server **out);
CHPIR_API void chpir_server_destroy(chpir_server *srv);

/* Persisted server state (the reference's Server exists in memory only, server.rs:16-21): the resident
 * packed column slice with its shape, so that respond can start without re-running setup -- one file per rank under column
 * sharding.  64-byte header (magic "CHPIRSV1", version, b, K, columns, first column, layout, payload bytes, FNV-1a checksum) followed
 * by the packed rows exactly as they sit in HBM.  The hint and the filter parameters are the caller's bytes already (setup returned
 * them); derived data (the limb planes of batch_tc = 1) is rebuilt on load.  chpir_server_load honours opts->batch_tc and
 * opts->respond_coalesce and ignores the setup-only fields. */
CHPIR_API int chpir_server_save(chpir_server *srv, const char *path);
CHPIR_API int chpir_server_load(chpir_ctx *ctx, const char *path, const chpir_setup_opts *opts, chpir_server **out);

/* Phase split of the last setup on this server, seconds. */
typedef struct chpir_setup_timing {
  double host_encode_s; /* filter construction + row encoding (0 if D was supplied)   */
  double h2d_s;         /* D upload                                                    */
  double pack_s;        /* respond layout pack + GEMM operand prep (limb split/transpose) */
  double expand_a_s;    /* TurboSHAKE128 expansion of A on device                      */
  double gemm_s;        /* hint GEMM                                                   */
  double d2h_s;         /* hint download                                               */
  double total_s;
  double device_encode_s; /* db_encode = DEVICE: device time of the row-fill waves (part of host_encode_s's wall time) */
  double xof_host_busy_s; /* host-pipelined mode: time the producer core spent inside the XOF (0 in device mode); expand_a_s is the
                             wall time of the whole expansion phase in either mode            */
  double a_cache_hit;     /* 1.0 if A came from the ctx cache (chpir_setup_opts.a_cache): no XOF chain ran; expand_a_s is then the
                             time of the u32 -> byte-plane splits around the panel GEMMs            */
  double xof_host_wait_s; /* host-pipelined mode: time the producer core stood still waiting for a free pinned chunk (uploader behind) */
} chpir_setup_timing;
CHPIR_API int chpir_server_setup_timing(const chpir_server *srv, chpir_setup_timing *out);

/* Shape / footprint of the resident server. */
typedef struct chpir_server_info {
  uint64_t rows_k;
  uint32_t cols_n;     /* columns in this slice */
  uint32_t col_begin;
  uint32_t mat_elem_bit_len;
  uint32_t fields_per_word; /* packed layout: floor(64 / b) fields per u64 word */
  uint64_t row_pitch_bytes; /* bytes per k-row of the packed slice */
  uint64_t packed_bytes;    /* total resident packed bytes = what one respond streams */
} chpir_server_info;
CHPIR_API int chpir_server_get_info(const chpir_server *srv, chpir_server_info *out);
/* The hint slice a setup with chpir_setup_opts.hint_on_device = 1 left in HBM (*rows x cols_n u32; owned by the server). */
CHPIR_API int chpir_server_hint_device(const chpir_server *srv, const uint32_t **hint_device, uint32_t *rows);

/* ---- respond: replaces Server::respond (server.rs:184-190) -> Matrix::from_bytes (matrix.rs:973-1010) +
 *      row_vector_x_compressed_transposed_matrix (matrix.rs:328-485) + to_bytes --------------------------- */
/* query: wire format 1 x K.  resp_out: wire format 1 x col_count (3768 B at N=940). */
CHPIR_API int chpir_server_respond(chpir_server *srv, const uint8_t *query, size_t query_len, uint8_t *resp_out, size_t resp_cap,
                         size_t *resp_len);
/* nq queries, each wire format 1 x K, concatenated responses each of `resp_stride` bytes.  One launch for the whole batch: the
 * streaming GEMV with grid.y = query, or -- from 6 queries up when the limb planes are resident (batch_tc) -- the tensor-core
 * limb GEMM, which passes over D once per 128 queries. */
CHPIR_API int chpir_server_respond_batch(chpir_server *srv, const uint8_t *const *queries, const size_t *query_lens, uint32_t nq,
                               uint8_t *resp_out, size_t resp_stride);
/* Device-resident variants (inputs already in HBM): q_device = nq x K u32, resp_device = nq x col_count u32,
 * enqueued on `cuda_stream` (a cudaStream_t passed as void*; NULL = the CUDA default stream); no synchronisation. */
CHPIR_API int chpir_server_respond_device(chpir_server *srv, const uint32_t *q_device, uint32_t nq, uint32_t *resp_device, void *cuda_stream);

/* Batched respond as a limb-decomposed int8 GEMM on the tensor cores (north_star (2), BASELINE.json configs[3]): the nq x K
 * query matrix plays the role A plays in setup, D's resident byte-limb planes are the B operand, so D is streamed once per
 * 128 queries instead of once per query.  Same inputs/outputs as chpir_server_respond_device; resp_device is overwritten.
 * Calls on one server must be stream-ordered with respect to each other (they share the operand staging ring).
 * Returns CHPIR_ERR_INVALID_ARGUMENT if the server was set up without limb planes (chpir_setup_opts.batch_tc). */
CHPIR_API int chpir_server_respond_device_tc(chpir_server *srv, const uint32_t *q_device, uint32_t nq, uint32_t *resp_device, void *cuda_stream);

/* ---- cluster: the same server on 1..8 GPUs of ONE process (north_star (3)) ---------------------------
 * The reference's public surface cannot grow a rank argument -- `Server::setup(seed, db)` (server.rs:103) and
 * `Server::respond(&self, query)` (server.rs:184) -- so the sharding lives behind the handle: the library owns the per-device
 * contexts, peer mappings, streams and NCCL communicators, and the Rust stub in INTEGRATION.md reaches all GPUs through the
 * unchanged two calls.  Two cuts of D, each where it moves the fewest bytes:
 *   setup   -- COLUMN slices: rank r computes hint columns chpir_cluster_plan gives it (no cross-rank arithmetic); the slices are
 *              gathered on GPU 0 (NCCL, libnccl.so.2 resolved at run time; CHPIR_CLUSTER_GATHER=p2p: peer copies) and downloaded once;
 *   respond -- ROW blocks (default for n > 1): rank r keeps rows [k_begin, +k_count) of D at full width and needs exactly the query
 *              words it ingests over its own PCIe link; what crosses NVLink is each rank's N-word partial response, summed on GPU 0
 *              (exact: addition mod 2^32 is order independent).  CHPIR_CLUSTER_SHARD=cols keeps the column cut for respond as well
 *              (the whole query is then gathered on every GPU by NVLink peer reads). */
typedef struct chpir_cluster chpir_cluster;
typedef struct chpir_cluster_server chpir_cluster_server;

/* n_gpus GPUs: device_ordinals[0..n_gpus) or, if NULL, ordinals 0..n_gpus-1.  n_gpus = 0 reads the environment variable CHPIR_GPUS
 * (default 1) -- the runtime knob a Rust `Server::setup` uses, since its signature has no room for one.  Every pair of GPUs must
 * have peer access (CHPIR_ERR_CUDA_PEER_ACCESS_UNAVAILABLE otherwise). */
CHPIR_API int chpir_cluster_create(int n_gpus, const int *device_ordinals, chpir_cluster **out);
CHPIR_API void chpir_cluster_destroy(chpir_cluster *cluster);
CHPIR_API int chpir_cluster_size(const chpir_cluster *cluster, int *n_gpus);
/* The per-GPU context of rank `rank` (borrowed; owned by the cluster). */
CHPIR_API int chpir_cluster_ctx(const chpir_cluster *cluster, int rank, chpir_ctx **ctx, int *device_ordinal);
/* The slice plan, pure host arithmetic: rank r of n owns columns [*col_begin, +*col_count) of D / hint / response (sizes differ by
 * at most one, earlier ranks take the remainder) and ingests words [*k_begin, +*k_count) of every query over its own PCIe link;
 * *k_pitch (may be NULL) receives the padded slice length in words (a multiple of 32, the same on every rank). */
CHPIR_API int chpir_cluster_plan(uint32_t n_ranks, uint32_t rank, uint64_t rows_k, uint32_t cols_n, uint32_t *col_begin, uint32_t *col_count,
                                 uint64_t *k_begin, uint64_t *k_count, uint64_t *k_pitch);

/* Server::setup on the cluster.  opts as for the single-GPU calls (col_begin / col_count / hint_on_device must be 0: the cluster
 * slices; respond_coalesce is implied for n_gpus > 1 and, for n_gpus = 1, selects the cluster's batching pipeline instead of one
 * slot per call; with db_encode = DEVICE every GPU builds its own columns of D in its HBM from one host-side peeling).  hint_out receives the COMPLETE
 * wire-format hint (8 + 4 * lwe_rows * N bytes), byte-identical to a single-GPU setup: each rank computes its column slice and
 * the slices are gathered on GPU 0 (NCCL by default, see CHPIR_CLUSTER_GATHER below) and downloaded once. */
CHPIR_API int chpir_cluster_server_setup_from_db(chpir_cluster *cluster, uint32_t arity, const uint8_t seed[CHPIR_SEED_BYTE_LEN],
                                                 uint64_t db_entry_count, const uint8_t *key_blob, const uint64_t *key_offsets,
                                                 const uint8_t *value_blob, const uint64_t *value_offsets, const uint64_t *filter_seed_rng,
                                                 const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap, size_t *hint_len,
                                                 uint8_t filter_params_out[CHPIR_FILTER_PARAM_BYTE_LEN], chpir_cluster_server **out);
/* d_host: the whole K x N matrix in host memory; every rank uploads its own columns. */
CHPIR_API int chpir_cluster_server_setup(chpir_cluster *cluster, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint32_t *d_host,
                                         uint64_t rows_k, uint32_t cols_n, uint32_t mat_elem_bit_len, const chpir_setup_opts *opts,
                                         uint8_t *hint_out, size_t hint_cap, size_t *hint_len, chpir_cluster_server **out);
/* d_slices[r]: rank r's columns as a compact K x col_count(r) u32 matrix already resident on rank r's GPU (synthetic D). */
CHPIR_API int chpir_cluster_server_setup_device(chpir_cluster *cluster, const uint8_t seed[CHPIR_SEED_BYTE_LEN],
                                                const uint32_t *const *d_slices, uint64_t rows_k, uint32_t cols_n, uint32_t mat_elem_bit_len,
                                                const chpir_setup_opts *opts, uint8_t *hint_out, size_t hint_cap, size_t *hint_len,
                                                chpir_cluster_server **out);
CHPIR_API void chpir_cluster_server_destroy(chpir_cluster_server *srv);
/* Rank r's resident slice as a plain single-GPU server handle, borrowed (info, timing, save); do not destroy it. */
CHPIR_API int chpir_cluster_server_shard(const chpir_cluster_server *srv, int rank, chpir_server **shard);
/* One file per rank, `<path_prefix>.rank<r>of<n>` in chpir_server_save's format; load needs a cluster of the same size. */
CHPIR_API int chpir_cluster_server_save(chpir_cluster_server *srv, const char *path_prefix);
CHPIR_API int chpir_cluster_server_load(chpir_cluster *cluster, const char *path_prefix, const chpir_setup_opts *opts, chpir_cluster_server **out);

/* Server::respond on the cluster: the same contract as chpir_server_respond (validation order, wire formats, thread-safe and
 * re-entrant), N = all columns.  Concurrent callers are coalesced: each caller DMAs its query's K/n slices to the n GPUs itself,
 * whoever arrives while the previous batch is on the GPUs shares one launch per GPU (GEMV for up to 5 queries, the tensor-core
 * limb GEMM for 6..128 when batch_tc planes are resident). */
CHPIR_API int chpir_cluster_server_respond(chpir_cluster_server *srv, const uint8_t *query, size_t query_len, uint8_t *resp_out,
                                           size_t resp_cap, size_t *resp_len);
CHPIR_API int chpir_cluster_server_respond_batch(chpir_cluster_server *srv, const uint8_t *const *queries, const size_t *query_lens,
                                                 uint32_t nq, uint8_t *resp_out, size_t resp_stride);

/* Load generator: n_threads native threads call chpir_cluster_server_respond concurrently -- what the reference's example server does
 * with one task per connection sharing an Arc<Server> (chalametpir_server/examples/server.rs:45-85).  Call j (0 <= j < total_calls)
 * sends queries[j % n_distinct] and writes its response to resp_out + (j % n_distinct) * resp_stride; *seconds (may be NULL) is the
 * wall time from the first thread's start to the last one's return.  Used by bench.py and the tests so that the caller side of the
 * end-to-end measurement is native code too. */
CHPIR_API int chpir_cluster_server_respond_concurrent(chpir_cluster_server *srv, const uint8_t *const *queries, const size_t *query_lens,
                                                      uint32_t n_distinct, uint64_t total_calls, uint8_t *resp_out, size_t resp_stride,
                                                      uint32_t n_threads, double *seconds);

/* Device-resident respond (inputs already in the cluster's HBM): q_slices[r] = nq x k_pitch u32 on rank r's GPU holding words
 * [k_begin(r), +k_count(r)) of every query -- exactly what the PCIe ingest above leaves behind -- and resp_device0 = nq x N u32 on
 * rank 0's GPU.  mode CHPIR_RESPOND_GEMV: every query streams every rank's packed slice once (the HBM-bound path); mode
 * CHPIR_RESPOND_TC: the limb GEMM, one pass over the byte planes per 128 queries.  `repeats` passes are enqueued back to back and
 * the call returns when they have finished; *device_ms (may be NULL) receives the device time of all of them, measured with CUDA
 * events on rank 0's GPU from before the first byte moves until the last rank's columns have landed in resp_device0. */
#define CHPIR_RESPOND_GEMV 0u
#define CHPIR_RESPOND_TC 1u
CHPIR_API int chpir_cluster_server_respond_device(chpir_cluster_server *srv, const uint32_t *const *q_slices, uint32_t nq,
                                                  uint32_t *resp_device0, uint32_t mode, uint32_t repeats, float *device_ms);

typedef struct chpir_cluster_server_info {
  uint32_t n_gpus;
  uint32_t cols_n;  /* all columns */
  uint64_t rows_k;
  uint32_t mat_elem_bit_len;
  uint32_t lwe_rows;
  uint64_t k_pitch;               /* padded query-slice length (words) */
  uint64_t packed_bytes_total;    /* resident packed bytes over all ranks = what one query streams in total */
  uint64_t packed_bytes_max_rank; /* the largest rank's share (bounds the per-query time) */
  double setup_total_s;           /* wall time of the whole cluster setup */
  double hint_gather_s;           /* of which: gathering the hint slices + download */
  uint32_t gather_uses_nccl;      /* 1 = the hint slices moved over NCCL, 0 = peer copies */
  uint32_t nccl_version;          /* ncclGetVersion, 0 if NCCL was not needed */
  /* coalescer statistics of chpir_cluster_server_respond since setup */
  uint64_t batches, queries, tc_batches;
  uint64_t pulled_queries;        /* of `queries`: fetched by the GPUs themselves from page-locked caller memory (no per-query DMA call) */
  uint32_t respond_by_rows;       /* 1 = respond runs on ROW blocks of D (default for n_gpus > 1), 0 = on column slices */
  uint32_t reserved0;
  double reshard_s;               /* of setup_total_s: re-cutting D from column slices into row blocks */
  /* wall time the batch leaders spent per pipeline stage of chpir_cluster_server_respond, summed over all batches since setup:
   * waiting for the ingest stage (= collecting callers), moving the queries over PCIe, waiting for the SMs, kernels + download */
  double ingest_wait_s, ingest_s, exec_wait_s, exec_s;
} chpir_cluster_server_info;
CHPIR_API int chpir_cluster_server_get_info(const chpir_cluster_server *srv, chpir_cluster_server_info *out);

/* ---- client: chalametpir_client::Client (chalametpir_client/src/client.rs:21-283) with the public
 *      matrix A resident in HBM and b = s*A + e computed there.  Not on the server hot path; it lets a complete PIR round be run and
 *      checked at the full 2^20-entry shape. ------------------------------------------------------------------------------------ */
typedef struct chpir_client chpir_client;
typedef struct chpir_client_opts {
  uint32_t lwe_rows;        /* 0 = CHPIR_LWE_DIMENSION; must equal the hint's row count (else InvalidHintMatrix) */
  uint32_t a_expand;        /* CHPIR_A_EXPAND_AUTO / _HOST_PIPELINED / _DEVICE, as in chpir_setup_opts             */
  uint32_t host_chunk_rows; /* as in chpir_setup_opts                                                              */
} chpir_client_opts;
typedef struct chpir_client_info {
  uint64_t rows_k;
  uint32_t cols_n, lwe_rows, mat_elem_bit_len, arity;
  uint64_t pub_mat_a_bytes;   /* resident A: lwe_rows * K * 4 */
  double setup_expand_s;      /* wall time of the A expansion in chpir_client_setup */
  float last_query_kernel_ms; /* device time of the s*A kernel of the last query */
} chpir_client_info;
/* Client::setup (client.rs:39-57).  filter_params: BinaryFuseFilter::to_bytes (68 bytes), hint: Matrix::to_bytes of 1774 x N. */
CHPIR_API int chpir_client_setup(chpir_ctx *ctx, const uint8_t seed[CHPIR_SEED_BYTE_LEN], const uint8_t *hint, size_t hint_len,
                       const uint8_t *filter_params, size_t filter_params_len, const chpir_client_opts *opts, chpir_client **out);
CHPIR_API void chpir_client_destroy(chpir_client *client);
/* Client::query (client.rs:95-194): query_out receives the wire-format 1 x K query (8 + 4K bytes).  The LWE secret and error are
 * drawn from ChaCha8 keyed by the OS (getrandom), as the reference's ChaCha8Rng::from_os_rng (matrix.rs:583); rng_seed != NULL
 * replaces the OS key by a reproducible one and is a TEST-ONLY hook (queries made with it are not private).
 * ArithmeticOverflowAddingQueryIndicator: retry (fresh randomness). */
CHPIR_API int chpir_client_query(chpir_client *client, const uint8_t *key, size_t key_len, const uint64_t *rng_seed, uint8_t *query_out,
                       size_t query_cap, size_t *query_len);
/* The same with the secret vector s (lwe_rows words) and the error vector e (K words) supplied: the deterministic core, used
 * to check the device arithmetic against the oracle word for word. */
CHPIR_API int chpir_client_query_with(chpir_client *client, const uint8_t *key, size_t key_len, const uint32_t *secret_s,
                            const uint32_t *error_e, uint8_t *query_out, size_t query_cap, size_t *query_len);
/* Client::process_response (client.rs:209-275): value_out receives the value bytes (at most N*b/8 - 33). */
CHPIR_API int chpir_client_process_response(chpir_client *client, const uint8_t *key, size_t key_len, const uint8_t *response,
                                  size_t response_len, uint8_t *value_out, size_t value_cap, size_t *value_len);
CHPIR_API int chpir_client_get_info(const chpir_client *client, chpir_client_info *out);

/* ---- building blocks exposed for parity tests and for composing other paths ------------------------ */
/* Matrix::generate_from_seed (matrix.rs:541-558) on device; rows [row_begin, row_begin+row_count) of the rows x cols
 * matrix are copied to out_host (row_count*cols u32).  The whole prefix of the XOF stream is walked on device. */
CHPIR_API int chpir_generate_from_seed(chpir_ctx *ctx, const uint8_t seed[CHPIR_SEED_BYTE_LEN], uint64_t rows, uint64_t cols,
                             uint64_t row_begin, uint64_t row_count, uint32_t *out_host);
/* &A * &D (matrix.rs:1040-1059) on device for host operands; variant as in chpir_setup_opts.gemm_variant.  Every entry of B must be
 * < 2^b_elem_bit_len (CHPIR_ERR_INVALID_ARGUMENT otherwise); the tensor-core variant covers b_elem_bit_len <= 16 (two byte limbs),
 * wider B is multiplied by the u32 SIMT kernel whatever `variant` says. */
CHPIR_API int chpir_matmul(chpir_ctx *ctx, const uint32_t *a_host, uint64_t a_rows, uint64_t a_cols, const uint32_t *b_host, uint64_t b_rows,
                 uint64_t b_cols, uint32_t b_elem_bit_len, uint32_t variant, uint32_t *out_host);
/* Matrix::generate_from_seed (matrix.rs:541-558) on a host core (csrc/host_xof.cpp; no GPU involved): rows
 * [row_begin, row_begin+row_count) of the rows x cols matrix.  impl: 0 = fastest available on this CPU, 1 = portable scalar,
 * 2 = BMI2 scalar, 3 = AVX-512 (one plane per zmm register), 4 = EVEX-128 (one lane per xmm register, AVX-512VL)
 * (CHPIR_ERR_INVALID_ARGUMENT if the CPU lacks it).  This is the producer of the
 * host-pipelined setup mode, exposed so that it can be checked against the device expander and the oracle byte for byte. */
CHPIR_API int chpir_host_generate_from_seed(const uint8_t seed[CHPIR_SEED_BYTE_LEN], uint64_t rows, uint64_t cols, uint64_t row_begin,
                                  uint64_t row_count, uint32_t impl, uint32_t *out_host);
/* Name of the implementation impl = 0 resolves to on this CPU ("evex128", "avx512", "bmi2" or "scalar"). */
CHPIR_API const char *chpir_host_xof_impl(void);
/* Device time (ms) of the dominant kernels in the last call on this ctx/server, for bench.py's roofline block. */
CHPIR_API int chpir_server_last_kernel_ms(const chpir_server *srv, float *respond_ms, float *gemm_ms, float *expand_ms);

#ifdef __cplusplus
}
#endif
#endif /* CHALAMET_B200_H */
