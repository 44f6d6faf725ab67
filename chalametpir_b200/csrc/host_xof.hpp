// host_xof.hpp -- host-side TurboSHAKE128 squeeze of the seed (see host_xof.cpp).
#pragma once
#include <cstddef>
#include <cstdint>

namespace chpir {

constexpr uint64_t kXofRate = 168;  // TurboSHAKE128 rate in bytes (RFC 9861)

struct HostXof {
  uint64_t s[25];
};

// TurboShake128::default(); absorb(seed[32]); finalize::<0x1F>()  (matrix.rs:542-551)
void host_xof_init(HostXof *x, const uint8_t seed[32]);
enum : int { kXofAuto = 0, kXofScalar = 1, kXofBmi2 = 2, kXofAvx512 = 3, kXofEvex128 = 4 };
// squeeze the next nblocks whole rate blocks (168 bytes each) into out; false if this CPU lacks the requested implementation
bool host_xof_squeeze_blocks(HostXof *x, uint8_t *out, uint64_t nblocks, int impl);
// advance the stream by nblocks blocks without producing output
void host_xof_skip_blocks(HostXof *x, uint64_t nblocks);
const char *host_xof_impl_name();

}  // namespace chpir
