// Counter-based dropout masks (Philox4x32-10, Salmon et al. SC'11): a keep decision is a pure function of
// (seed, step offset, site, element coordinates), so the backward kernels regenerate the forward's mask instead of
// reading a stored one, forward and backward may tile the same tensor differently, and a CUDA-graph replay gets fresh
// masks by bumping the offset in device memory.  The reference draws its masks from torch's generator
// (nn.Dropout at models/layers.py:440-441,469,485, models/decoder.py:236-243; SDPA dropout_p at models/layers.py:465;
// nn.MultiheadAttention dropout at :537-542): the stream differs, the distribution (Bernoulli(1-p), survivors scaled by
// 1/(1-p)) is the same.  tests/helpers.py restates this generator in numpy; masks are compared bit for bit.
#pragma once
#include <stdint.h>

namespace i2t {

// what a kernel needs to drop: thr == 0 means "no dropout" (the branch is uniform over the grid)
struct DropArgs {
  const unsigned long long* state;   // device: [0] = seed, [1] = step offset
  uint32_t site;                     // which dropout call of the forward pass this is
  uint32_t thr;                      // keep iff random u32 >= thr; thr = round(p * 2^32)
  uint32_t thr16;                    // attention-probability sites spend 16 random bits per key: keep iff u16 >= thr16 = round(p * 2^16)
  float inv_keep;                    // 1 / (1 - p)
};

inline DropArgs make_drop(float p, const void* state, int64_t site) {
  DropArgs d;
  d.state = (const unsigned long long*)state;
  d.site = (uint32_t)site;
  d.thr = 0u;
  d.thr16 = 0u;
  d.inv_keep = 1.0f;
  if (p > 0.f && state != nullptr) {
    double t = (double)p * 4294967296.0 + 0.5;
    d.thr = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    const double t16 = (double)p * 65536.0 + 0.5;
    d.thr16 = t16 >= 65535.0 ? 65535u : (uint32_t)t16;
    d.inv_keep = 1.0f / (1.0f - p);
  }
  return d;
}

struct Philox4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// per-launch constants read once from device memory
struct DropKey {
  uint32_t k0, k1, off;
};
__device__ __forceinline__ DropKey drop_key(const DropArgs& d) {
  const unsigned long long seed = d.state[0], off = d.state[1];
  return DropKey{(uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(off >> 32), (uint32_t)off};
}

// elementwise sites: elements 4*e4 .. 4*e4+3 of the flattened tensor
__device__ __forceinline__ Philox4 drop_elem4(const DropArgs& d, const DropKey& k, uint64_t e4) {
  return philox4x32_10((uint32_t)e4, (uint32_t)(e4 >> 32), d.site, k.off, k.k0, k.k1);
}
// attention-probability sites: row = (b*H + h)*Tq + qi.  One Philox call yields EIGHT 16-bit values (the Philox rounds, not the
// softmax, dominated the dropout attention kernels: 16 bits per key halve the calls; p is resolved to 2^-16).  Call (blk, h) covers,
// in the 16-key block blk = kj >> 4, the pairs t = 2h and 2h + 1 (t = (kj >> 1) & 3), i.e. keys 16 blk + {4h .. 4h+3, 4h+8 .. 4h+11};
// half-word j = 4 (t & 1) + 2 ((kj >> 3) & 1) + (kj & 1) of the call belongs to key kj: half-words 4 (t & 1) + {0, 1, 2, 3} are the
// four key columns {2t, 2t+1, 8+2t, 9+2t} one thread of an m16n8k16 accumulator pair owns in a query row.
__device__ __forceinline__ Philox4 drop_attn8(const DropArgs& d, const DropKey& k, uint32_t row, uint32_t blk, uint32_t h) {
  return philox4x32_10(row, blk * 2u + h, d.site | 0x80000000u, k.off, k.k0, k.k1);
}
__device__ __forceinline__ uint32_t philox_word(const Philox4& r, int lane) {
  return lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
}
__device__ __forceinline__ uint32_t philox_half(const Philox4& r, int j) {           // half-word j = 0..7
  const uint32_t w = philox_word(r, j >> 1);
  return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
}
// single element (slow path: the fp32 parity kernels)
__device__ __forceinline__ bool drop_attn_keep(const DropArgs& d, const DropKey& k, uint32_t row, int kj) {
  const uint32_t t = ((uint32_t)kj >> 1) & 3u;
  const Philox4 r = drop_attn8(d, k, row, (uint32_t)kj >> 4, t >> 1);
  return philox_half(r, (int)((t & 1u) * 4u + (((uint32_t)kj >> 3) & 1u) * 2u + ((uint32_t)kj & 1u))) >= d.thr16;
}

}  // namespace i2t
