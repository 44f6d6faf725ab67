// host_xof_bench.cpp -- ns per Keccak-p[1600,12] permutation of every host XOF implementation on THIS machine's cores
// (csrc/host_xof.cpp: 1 = portable scalar, 2 = BMI2 scalar, 3 = AVX-512 planes, 4 = EVEX-128 lane per register), each checked against the scalar stream.
//   g++ -O3 -std=c++17 -I chalametpir_b200/csrc tools/host_xof_bench.cpp chalametpir_b200/csrc/host_xof.cpp -o tools/host_xof_bench
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include "host_xof.hpp"

using namespace chpir;

int main() {
  const uint64_t kBlocks = 200000, kCheck = 5000;
  std::vector<uint8_t> buf(kBlocks * kXofRate), ref;
  uint8_t seed[32];
  for (int i = 0; i < 32; i++) seed[i] = uint8_t(i * 7 + 1);
  for (int impl : {1, 2, 3, 4}) {
    HostXof x;
    host_xof_init(&x, seed);
    if (!host_xof_squeeze_blocks(&x, buf.data(), kCheck, impl)) {
      std::printf("impl %d: not available on this CPU\n", impl);
      continue;
    }
    if (ref.empty()) ref.assign(buf.begin(), buf.begin() + kCheck * kXofRate);
    const bool same = !std::memcmp(ref.data(), buf.data(), kCheck * kXofRate);
    double best = 1e30, worst = 0;
    for (int rep = 0; rep < 7; rep++) {
      const auto t0 = std::chrono::steady_clock::now();
      host_xof_squeeze_blocks(&x, buf.data(), kBlocks, impl);
      const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      best = t < best ? t : best, worst = t > worst ? t : worst;
    }
    std::printf("impl %d: %.1f ns/permutation (best of 7 x %lu blocks; worst %.1f)  stream identical to scalar: %s\n", impl, best / kBlocks * 1e9,
                (unsigned long)kBlocks, worst / kBlocks * 1e9, same ? "yes" : "NO");
  }
  std::printf("auto picks: %s\n", host_xof_impl_name());
  return 0;
}
