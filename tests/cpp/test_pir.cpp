// test_pir.cpp -- the reference's end-to-end tests (integrations/src/test_pir.rs:12-142) and the error behaviour of
// Server::setup / Server::respond (chalametpir_server/src/server.rs:103-107,184-190), written against include/chalamet_b200.hpp:
// plain C++17, no Python, no oracle.  Built by __graft_entry__.build(), run by tests/test_gpu_cluster.py on the GPU box.
//
//   build/test_pir_cpp [iterations=3] [log2 of the largest database=13]
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>

#include "chalamet_b200.hpp"

using namespace chalametpir;

#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      std::fprintf(stderr, "%s:%d: check failed: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                          \
    }                                                                        \
  } while (0)

using Database = std::unordered_map<std::string, std::string>;

// utils::generate_random_kv_database (chalametpir_common/src/utils.rs:22-45): keys of 16..32 bytes, values of 1..512 bytes
static Database generate_random_kv_database(size_t num_kv_pairs, std::mt19937_64 &rng) {
  Database kv;
  kv.reserve(num_kv_pairs);
  auto bytes = [&](size_t n) {
    std::string s(n, '\0');
    for (auto &c : s) c = char(rng() & 0xff);
    return s;
  };
  while (kv.size() < num_kv_pairs) kv.emplace(bytes(16 + rng() % 17), bytes(1 + rng() % 512));
  return kv;
}

template <uint32_t ARITY>
static void test_keyword_pir(size_t iterations, unsigned max_log2, std::mt19937_64 &rng) {
  constexpr size_t NUMBER_OF_PIR_QUERIES = 10;
  for (size_t it = 0; it < iterations; it++) {
    const size_t lo = size_t(1) << 8, hi = size_t(1) << max_log2;
    const size_t num_kv_pairs_in_db = lo + rng() % (hi - lo + 1);
    const Database kv_db = generate_random_kv_database(num_kv_pairs_in_db, rng);
    Seed seed_mu;
    for (auto &b : seed_mu) b = uint8_t(rng());

    auto [server, hint_bytes, filter_param_bytes] = Server::setup<ARITY>(seed_mu, kv_db).expect("Server setup failed");
    Client client = Client::setup(seed_mu, hint_bytes, filter_param_bytes).expect("Client setup failed");

    std::vector<const std::string *> keys;
    for (const auto &kv : kv_db) keys.push_back(&kv.first);
    std::shuffle(keys.begin(), keys.end(), rng);
    keys.resize(std::min(keys.size(), NUMBER_OF_PIR_QUERIES));
    size_t retries = 0;
    for (size_t i = 0; i < keys.size();) {
      const std::string &key = *keys[i];
      auto query = client.query(key);
      if (query.is_err()) {  // the only error a well-formed query may meet: draw fresh randomness (test_pir.rs:66-70)
        CHECK(query.error() == ChalametPIRError::ArithmeticOverflowAddingQueryIndicator);
        CHECK(++retries < 1000);
        continue;
      }
      const Bytes response_bytes = server.respond(query.value()).expect("Server can't respond");
      const Bytes received_value = client.process_response(key, response_bytes).expect("Client can't extract value from response");
      const std::string &value = kv_db.at(key);
      CHECK(received_value.size() == value.size() && std::memcmp(received_value.data(), value.data(), value.size()) == 0);
      i++;
    }
    std::printf("  %u-wise, %zu entries: %zu values recovered (hint %zu bytes)\n", ARITY, num_kv_pairs_in_db, keys.size(), hint_bytes.size());
  }
}

static void test_error_behaviour(std::mt19937_64 &rng) {
  Seed seed_mu{};
  // server.rs:105-107
  {
    auto r = Server::setup<3>(seed_mu, Database{});
    CHECK(r.is_err() && r.error() == ChalametPIRError::EmptyKVDatabase);
  }
  const Database kv_db = generate_random_kv_database(300, rng);
  auto [server, hint_bytes, filter_param_bytes] = Server::setup<3>(seed_mu, kv_db).expect("Server setup failed");
  Client client = Client::setup(seed_mu, hint_bytes, filter_param_bytes).expect("Client setup failed");
  const std::string &key = kv_db.begin()->first;
  Result<Bytes> q = client.query(key);
  while (q.is_err()) {
    CHECK(q.error() == ChalametPIRError::ArithmeticOverflowAddingQueryIndicator);
    q = client.query(key);
  }
  // a second query for a key whose response is still outstanding (client.rs:96-98)
  {
    auto again = client.query(key);
    CHECK(again.is_err() && again.error() == ChalametPIRError::PendingQueryExistsForKey);
  }
  const Bytes query = q.value();
  // Matrix::from_bytes (matrix.rs:973-1010): too short, length not matching the header
  CHECK(server.respond(query.data(), 8).error() == ChalametPIRError::FailedToDeserializeMatrixFromBytes);
  CHECK(server.respond(query.data(), query.size() - 4).error() == ChalametPIRError::FailedToDeserializeMatrixFromBytes);
  // a well-formed 1 x (K - 1) matrix: wrong dimension for the product (matrix.rs:336-338)
  {
    Bytes shorter(query.begin(), query.end() - 4);
    const uint32_t k = uint32_t((shorter.size() - 8) / 4);
    std::memcpy(shorter.data() + 4, &k, 4);
    CHECK(server.respond(shorter).error() == ChalametPIRError::IncompatibleDimensionForRowVectorTransposedMatrixMultiplication);
  }
  // a 2 x K/2 matrix of the right byte length: not a row vector
  if (((query.size() - 8) / 4) % 2 == 0) {
    Bytes two_rows = query;
    const uint32_t rows = 2, cols = uint32_t((query.size() - 8) / 8);
    std::memcpy(two_rows.data(), &rows, 4);
    std::memcpy(two_rows.data() + 4, &cols, 4);
    CHECK(server.respond(two_rows).error() == ChalametPIRError::IncompatibleDimensionForRowVectorTransposedMatrixMultiplication);
  }
  // the untouched query still works, and a response for a key nobody asked about is refused (client.rs:213-216)
  const Bytes response = server.respond(query).expect("Server can't respond");
  {
    auto r = client.process_response(std::string("no such pending key"), response);
    CHECK(r.is_err() && r.error() == ChalametPIRError::PendingQueryDoesNotExistForKey);
  }
  const Bytes value = client.process_response(key, response).expect("Client can't extract value from response");
  CHECK(std::string(value.begin(), value.end()) == kv_db.at(key));
  // hint / filter bytes that are not what Server::setup returned (client.rs:40-51)
  {
    Bytes bad_filter(filter_param_bytes.begin(), filter_param_bytes.end() - 1);
    auto r = Client::setup(seed_mu, hint_bytes, bad_filter);
    CHECK(r.is_err() && r.error() == ChalametPIRError::FailedToDeserializeFilterFromBytes);
    Bytes bad_hint(hint_bytes.begin(), hint_bytes.end() - 4);
    auto r2 = Client::setup(seed_mu, bad_hint, filter_param_bytes);
    CHECK(r2.is_err());
  }
  std::printf("  error variants ok\n");
}

// examples/server.rs:45-85: one Server shared by concurrent tasks
static void test_shared_server(std::mt19937_64 &rng) {
  Seed seed_mu;
  for (auto &b : seed_mu) b = uint8_t(rng());
  const Database kv_db = generate_random_kv_database(2000, rng);
  auto [server, hint_bytes, filter_param_bytes] = Server::setup<4>(seed_mu, kv_db).expect("Server setup failed");
  Client client = Client::setup(seed_mu, hint_bytes, filter_param_bytes).expect("Client setup failed");
  constexpr int kTasks = 16;
  std::vector<const std::string *> keys;
  std::vector<Bytes> queries;
  for (const auto &kv : kv_db) {
    if (int(keys.size()) == kTasks) break;
    auto q = client.query(kv.first);
    if (q.is_err()) continue;
    keys.push_back(&kv.first);
    queries.push_back(q.value());
  }
  std::vector<Bytes> responses(keys.size());
  std::atomic<int> failures{0};
  std::vector<std::thread> tasks;
  for (size_t t = 0; t < keys.size(); t++)
    tasks.emplace_back([&, t, srv = server] {  // a clone of the Arc<Server> per task
      for (int rep = 0; rep < 4; rep++) {
        auto r = srv.respond(queries[t]);
        if (r.is_err()) {
          failures++;
          return;
        }
        responses[t] = r.value();
      }
    });
  for (auto &t : tasks) t.join();
  CHECK(failures == 0);
  for (size_t t = 0; t < keys.size(); t++) {
    const Bytes v = client.process_response(*keys[t], responses[t]).expect("Client can't extract value from response");
    CHECK(std::string(v.begin(), v.end()) == kv_db.at(*keys[t]));
  }
  std::printf("  %zu concurrent tasks on one Server ok\n", keys.size());
}

int main(int argc, char **argv) {
  const size_t iterations = argc > 1 ? size_t(std::atoi(argv[1])) : 3;
  const unsigned max_log2 = argc > 2 ? unsigned(std::atoi(argv[2])) : 13;
  std::mt19937_64 rng(std::random_device{}());  // like the reference's tests: fresh entropy every run
  std::printf("test_keyword_pir_with_3_wise_xor_filter\n");
  test_keyword_pir<3>(iterations, max_log2, rng);
  std::printf("test_keyword_pir_with_4_wise_xor_filter\n");
  test_keyword_pir<4>(iterations, max_log2, rng);
  std::printf("test_error_behaviour\n");
  test_error_behaviour(rng);
  std::printf("test_shared_server\n");
  test_shared_server(rng);
  std::printf("test_pir_cpp ok\n");
  return 0;
}
