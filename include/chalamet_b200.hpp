// chalamet_b200.hpp -- the host side above the C ABI, in C++17, header only.
//
// The reference is a Rust crate; no Rust toolchain exists where this library is built, so the types a user of the reference works
// with are mirrored here in the other compiled language of the image, over include/chalamet_b200.h:
//
//   chalametpir::Server::setup<ARITY>(seed, db) -> Result<(Server, hint_bytes, filter_param_bytes)>   chalametpir_server/src/server.rs:103
//   chalametpir::Server::respond(query_bytes)   -> Result<response_bytes>                              chalametpir_server/src/server.rs:184
//   chalametpir::Client::setup(seed, hint_bytes, filter_param_bytes) -> Result<Client>                 chalametpir_client/src/client.rs:39
//   chalametpir::Client::query(key) -> Result<query_bytes>                                             chalametpir_client/src/client.rs:95
//   chalametpir::Client::process_response(key, response_bytes) -> Result<value>                        chalametpir_client/src/client.rs:209
//   chalametpir::ChalametPIRError                                                                      chalametpir_common/src/error.rs:8-50
//
// Same names, same argument meaning, same error variants in the same situations, so that a test written against the reference
// (integrations/src/test_pir.rs) reads the same against this header (tests/cpp/test_pir.cpp).  As in the reference nothing throws on
// a protocol error: every call returns a Result; only unwrap()/expect() on an error abort (Rust's panic).
//
// The number of GPUs is a property of the process, not of a call (the two signatures have no room for it): the first Server or
// Client creates one chpir_cluster over $CHPIR_GPUS GPUs (default 1) that all later ones share.
#ifndef CHALAMET_B200_HPP
#define CHALAMET_B200_HPP

#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <utility>
#include <variant>
#include <vector>

#include "chalamet_b200.h"

namespace chalametpir {

using Bytes = std::vector<uint8_t>;
constexpr size_t SEED_BYTE_LEN = CHPIR_SEED_BYTE_LEN;      // chalametpir_common/src/params.rs:5
constexpr uint32_t LWE_DIMENSION = CHPIR_LWE_DIMENSION;    // chalametpir_common/src/params.rs:1
using Seed = std::array<uint8_t, SEED_BYTE_LEN>;

// chalametpir_common/src/error.rs:8-50, variant for variant; the Vulkan* variants of the reference's `gpu` feature are replaced by
// the Cuda* / Nccl ones (the reference maps every device failure to a unit variant the same way, gpu_utils.rs:26-72).
enum class ChalametPIRError : int {
  InvalidMatrixDimension = CHPIR_ERR_INVALID_MATRIX_DIMENSION,
  IncompatibleDimensionForMatrixMultiplication = CHPIR_ERR_INCOMPATIBLE_DIMENSION_FOR_MATRIX_MULTIPLICATION,
  IncompatibleDimensionForRowVectorTransposedMatrixMultiplication = CHPIR_ERR_INCOMPATIBLE_DIMENSION_FOR_ROW_VECTOR_TRANSPOSED_MATRIX_MULTIPLICATION,
  FailedToDeserializeMatrixFromBytes = CHPIR_ERR_FAILED_TO_DESERIALIZE_MATRIX_FROM_BYTES,
  EmptyKVDatabase = CHPIR_ERR_EMPTY_KV_DATABASE,
  ExhaustedAllAttemptsToBuild3WiseXorFilter = CHPIR_ERR_EXHAUSTED_ALL_ATTEMPTS_TO_BUILD_3_WISE_XOR_FILTER,
  ExhaustedAllAttemptsToBuild4WiseXorFilter = CHPIR_ERR_EXHAUSTED_ALL_ATTEMPTS_TO_BUILD_4_WISE_XOR_FILTER,
  RowNotDecodable = CHPIR_ERR_ROW_NOT_DECODABLE,
  DecodedRowNotPrependedWithDigestOfKey = CHPIR_ERR_DECODED_ROW_NOT_PREPENDED_WITH_DIGEST_OF_KEY,
  FailedToDeserializeFilterFromBytes = CHPIR_ERR_FAILED_TO_DESERIALIZE_FILTER_FROM_BYTES,
  KVDatabaseSizeTooLarge = CHPIR_ERR_KV_DATABASE_SIZE_TOO_LARGE,
  InvalidHintMatrix = CHPIR_ERR_INVALID_HINT_MATRIX,
  ArithmeticOverflowAddingQueryIndicator = CHPIR_ERR_ARITHMETIC_OVERFLOW_ADDING_QUERY_INDICATOR,
  UnsupportedArityForBinaryFuseFilter = CHPIR_ERR_UNSUPPORTED_ARITY_FOR_BINARY_FUSE_FILTER,
  InvalidResponseVector = CHPIR_ERR_INVALID_RESPONSE_VECTOR,
  ImpossibleEncodedDBMatrixElementBitLength = CHPIR_ERR_IMPOSSIBLE_ENCODED_DB_MATRIX_ELEMENT_BIT_LENGTH,
  PendingQueryExistsForKey = CHPIR_ERR_PENDING_QUERY_EXISTS_FOR_KEY,
  PendingQueryDoesNotExistForKey = CHPIR_ERR_PENDING_QUERY_DOES_NOT_EXIST_FOR_KEY,
  InvalidArgument = CHPIR_ERR_INVALID_ARGUMENT,
  BufferTooSmall = CHPIR_ERR_BUFFER_TOO_SMALL,
  IoFailed = CHPIR_ERR_IO_FAILED,
  InvalidSavedServer = CHPIR_ERR_INVALID_SAVED_SERVER,
  CudaDeviceNotFound = CHPIR_ERR_CUDA_DEVICE_NOT_FOUND,
  CudaAllocationFailed = CHPIR_ERR_CUDA_ALLOCATION_FAILED,
  CudaTransferFailed = CHPIR_ERR_CUDA_TRANSFER_FAILED,
  CudaKernelLaunchFailed = CHPIR_ERR_CUDA_KERNEL_LAUNCH_FAILED,
  CudaKernelExecutionFailed = CHPIR_ERR_CUDA_KERNEL_EXECUTION_FAILED,
  CudaUnsupportedDevice = CHPIR_ERR_CUDA_UNSUPPORTED_DEVICE,
  CudaPeerAccessUnavailable = CHPIR_ERR_CUDA_PEER_ACCESS_UNAVAILABLE,
  NcclFailed = CHPIR_ERR_NCCL_FAILED,
  HostAllocationFailed = CHPIR_ERR_HOST_ALLOCATION_FAILED,
};
inline const char *to_string(ChalametPIRError e) { return chpir_strerror(static_cast<int>(e)); }

// Result<T, ChalametPIRError>
template <class T>
class Result {
 public:
  Result(T v) : v_(std::move(v)) {}                // NOLINT: Ok(v)
  Result(ChalametPIRError e) : v_(e) {}            // NOLINT: Err(e)
  bool is_ok() const { return v_.index() == 0; }
  bool is_err() const { return !is_ok(); }
  explicit operator bool() const { return is_ok(); }
  ChalametPIRError error() const { return std::get<1>(v_); }
  T &value() & { return std::get<0>(v_); }
  const T &value() const & { return std::get<0>(v_); }
  T unwrap() && { return std::move(*this).expect("called unwrap() on an error"); }
  T expect(const char *msg) && {
    if (is_err()) {
      std::fprintf(stderr, "%s: %s\n", msg, to_string(error()));
      std::abort();
    }
    return std::move(std::get<0>(v_));
  }

 private:
  std::variant<T, ChalametPIRError> v_;
};

namespace detail {

// One cluster per process, created by the first user, destroyed with the last.
struct Cluster {
  chpir_cluster *handle = nullptr;
  ~Cluster() {
    if (handle) chpir_cluster_destroy(handle);
  }
  static Result<std::shared_ptr<Cluster>> shared() {
    static std::mutex mu;
    static std::weak_ptr<Cluster> cached;
    std::lock_guard<std::mutex> g(mu);
    if (auto c = cached.lock()) return c;
    auto c = std::make_shared<Cluster>();
    if (int rc = chpir_cluster_create(0, nullptr, &c->handle); rc != CHPIR_OK) return static_cast<ChalametPIRError>(rc);
    cached = c;
    return c;
  }
};

// HashMap<&[u8], &[u8]> -> blob + offsets, in the container's iteration order
template <class Db>
void flatten(const Db &db, Bytes *kb, std::vector<uint64_t> *ko, Bytes *vb, std::vector<uint64_t> *vo) {
  ko->assign(1, 0);
  vo->assign(1, 0);
  for (const auto &kv : db) {
    const auto *k = reinterpret_cast<const uint8_t *>(kv.first.data());
    const auto *v = reinterpret_cast<const uint8_t *>(kv.second.data());
    kb->insert(kb->end(), k, k + kv.first.size());
    vb->insert(vb->end(), v, v + kv.second.size());
    ko->push_back(kb->size());
    vo->push_back(vb->size());
  }
  if (kb->empty()) kb->push_back(0);  // the C ABI wants non-null blobs
  if (vb->empty()) vb->push_back(0);
}

}  // namespace detail

// Not in the reference: what its signature has no room for.  The defaults are the reference's behaviour.
struct SetupOptions {
  bool device_row_fill = false;          // rows of D encoded and filled on the GPU instead of the host (same bytes)
  bool keep_a = false;                   // keep A = generate_from_seed(1774, K, seed) in HBM for the next setup with this seed
  uint32_t lwe_rows = 0;                 // 0 = LWE_DIMENSION; tests may shrink it (the Client must be told the same)
  const uint64_t *filter_seed_rng = nullptr;  // nullptr = OS entropy (binary_fuse_filter.rs:100,309); else a reproducible stream
};

// chalametpir_server::Server (server.rs:15-21).  Copies share the resident database, like clones of an Arc<Server>
// (examples/server.rs:45); respond() may be called from any number of threads at once.
class Server {
 public:
  // Server::setup::<ARITY>(seed_mu, db)  (server.rs:103).  Db: any range of (key, value) pairs whose members have data()/size() over
  // bytes -- std::unordered_map<std::string, std::string>, std::map<Bytes, Bytes>, a vector of pairs of string_views ...
  template <uint32_t ARITY, class Db>
  static Result<std::tuple<Server, Bytes, Bytes>> setup(const Seed &seed, const Db &db, const SetupOptions &opt = {}) {
    if (db.size() == 0) return ChalametPIRError::EmptyKVDatabase;  // server.rs:105-107
    auto cluster = detail::Cluster::shared();
    if (cluster.is_err()) return cluster.error();
    Bytes kb, vb;
    std::vector<uint64_t> ko, vo;
    detail::flatten(db, &kb, &ko, &vb, &vo);
    const uint64_t n = ko.size() - 1;
    uint32_t b = 0;
    if (int rc = chpir_find_mat_elem_bit_len(n, &b); rc != CHPIR_OK) return static_cast<ChalametPIRError>(rc);
    uint64_t max_vlen = 0, K = 0, N = 0;
    for (uint64_t i = 0; i < n; i++) max_vlen = vo[i + 1] - vo[i] > max_vlen ? vo[i + 1] - vo[i] : max_vlen;
    if (int rc = chpir_db_matrix_shape(ARITY, n, max_vlen, b, &K, &N); rc != CHPIR_OK) return static_cast<ChalametPIRError>(rc);
    chpir_setup_opts o{};
    o.lwe_rows = opt.lwe_rows;
    o.db_encode = opt.device_row_fill ? CHPIR_DB_ENCODE_DEVICE : CHPIR_DB_ENCODE_HOST;
    o.a_cache = opt.keep_a ? 1u : 0u;
    o.respond_coalesce = 1;  // concurrent respond() calls share launches
    const uint64_t lwe = opt.lwe_rows ? opt.lwe_rows : LWE_DIMENSION;
    Bytes hint(8 + 4 * lwe * N), filter(CHPIR_FILTER_PARAM_BYTE_LEN);
    size_t hint_len = 0;
    auto st = std::make_shared<State>();
    st->cluster = cluster.value();
    if (int rc = chpir_cluster_server_setup_from_db(st->cluster->handle, ARITY, seed.data(), n, kb.data(), ko.data(), vb.data(), vo.data(),
                                                    opt.filter_seed_rng, &o, hint.data(), hint.size(), &hint_len, filter.data(), &st->srv);
        rc != CHPIR_OK)
      return static_cast<ChalametPIRError>(rc);
    hint.resize(hint_len);
    st->resp_len = 8 + 4 * N;
    return std::make_tuple(Server(std::move(st)), std::move(hint), std::move(filter));
  }

  // Server::respond(&self, query)  (server.rs:184-190)
  Result<Bytes> respond(const uint8_t *query, size_t query_len) const {
    Bytes resp(st_->resp_len);
    size_t len = 0;
    if (int rc = chpir_cluster_server_respond(st_->srv, query, query_len, resp.data(), resp.size(), &len); rc != CHPIR_OK)
      return static_cast<ChalametPIRError>(rc);
    resp.resize(len);
    return resp;
  }
  Result<Bytes> respond(const Bytes &query) const { return respond(query.data(), query.size()); }

  // server.rs:193-218
  static Result<uint32_t> find_encoded_db_matrix_element_bit_length(uint64_t db_entry_count) {
    uint32_t b = 0;
    if (int rc = chpir_find_mat_elem_bit_len(db_entry_count, &b); rc != CHPIR_OK) return static_cast<ChalametPIRError>(rc);
    return b;
  }

  chpir_cluster_server *native_handle() const { return st_->srv; }

 private:
  struct State {
    std::shared_ptr<detail::Cluster> cluster;
    chpir_cluster_server *srv = nullptr;
    size_t resp_len = 0;
    ~State() {
      if (srv) chpir_cluster_server_destroy(srv);
    }
  };
  explicit Server(std::shared_ptr<State> st) : st_(std::move(st)) {}
  std::shared_ptr<State> st_;
};

// chalametpir_client::Client (client.rs:21-283) with A resident in HBM.  Not on the server hot path; it is what lets a complete
// PIR round be run against the Server above.
class Client {
 public:
  // Client::setup(seed_mu, hint_bytes, filter_param_bytes)  (client.rs:39-57)
  static Result<Client> setup(const Seed &seed, const Bytes &hint_bytes, const Bytes &filter_param_bytes, uint32_t lwe_rows = 0) {
    auto cluster = detail::Cluster::shared();
    if (cluster.is_err()) return cluster.error();
    chpir_ctx *ctx = nullptr;
    int dev = 0;
    if (int rc = chpir_cluster_ctx(cluster.value()->handle, 0, &ctx, &dev); rc != CHPIR_OK) return static_cast<ChalametPIRError>(rc);
    chpir_client_opts o{};
    o.lwe_rows = lwe_rows;
    auto st = std::make_shared<State>();
    st->cluster = cluster.value();
    if (int rc = chpir_client_setup(ctx, seed.data(), hint_bytes.data(), hint_bytes.size(), filter_param_bytes.data(), filter_param_bytes.size(), &o,
                                    &st->cl);
        rc != CHPIR_OK)
      return static_cast<ChalametPIRError>(rc);
    if (int rc = chpir_client_get_info(st->cl, &st->info); rc != CHPIR_OK) return static_cast<ChalametPIRError>(rc);
    return Client(std::move(st));
  }

  // Client::query(&mut self, key)  (client.rs:95-194)
  Result<Bytes> query(const uint8_t *key, size_t key_len) {
    Bytes q(8 + 4 * st_->info.rows_k);
    size_t len = 0;
    if (int rc = chpir_client_query(st_->cl, key, key_len, nullptr, q.data(), q.size(), &len); rc != CHPIR_OK)
      return static_cast<ChalametPIRError>(rc);
    q.resize(len);
    return q;
  }
  template <class Key>
  Result<Bytes> query(const Key &key) {
    return query(reinterpret_cast<const uint8_t *>(key.data()), key.size());
  }

  // Client::process_response(&mut self, key, response_bytes)  (client.rs:209-275)
  Result<Bytes> process_response(const uint8_t *key, size_t key_len, const Bytes &response_bytes) {
    Bytes value(size_t(st_->info.cols_n) * st_->info.mat_elem_bit_len / 8 + 8);
    size_t len = 0;
    if (int rc = chpir_client_process_response(st_->cl, key, key_len, response_bytes.data(), response_bytes.size(), value.data(), value.size(), &len);
        rc != CHPIR_OK)
      return static_cast<ChalametPIRError>(rc);
    value.resize(len);
    return value;
  }
  template <class Key>
  Result<Bytes> process_response(const Key &key, const Bytes &response_bytes) {
    return process_response(reinterpret_cast<const uint8_t *>(key.data()), key.size(), response_bytes);
  }

 private:
  struct State {
    std::shared_ptr<detail::Cluster> cluster;
    chpir_client *cl = nullptr;
    chpir_client_info info{};
    ~State() {
      if (cl) chpir_client_destroy(cl);
    }
  };
  explicit Client(std::shared_ptr<State> st) : st_(std::move(st)) {}
  std::shared_ptr<State> st_;
};

}  // namespace chalametpir

#endif  // CHALAMET_B200_HPP
