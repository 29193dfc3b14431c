// Counter-based dropout: Philox-4x32-10 keyed by (seed), counter = (element offset, site).
// Masks are regenerated in the backward pass from the same (seed, site, offset) triple, so no
// mask tensor is ever stored.  RNG-stream parity with torch is not a goal.
#pragma once
#include <stdint.h>

namespace msr3d {

// Philox-4x32-10: four 32-bit words per (key, counter).  Element `offset` uses word (offset & 3) of
// block (offset >> 2), so vectorised callers pay one evaluation per four consecutive elements.
struct Philox4 { uint32_t w[4]; };
__host__ __device__ __forceinline__ Philox4 philox4(uint64_t seed, uint64_t block, uint32_t site) {
  uint32_t c0 = (uint32_t)block, c1 = (uint32_t)(block >> 32), c2 = site, c3 = 0x6d737233u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox4 r;
  r.w[0] = c0; r.w[1] = c1; r.w[2] = c2; r.w[3] = c3;
  return r;
}
__host__ __device__ __forceinline__ uint32_t philox_word(uint64_t seed, uint64_t offset, uint32_t site) {
  const Philox4 p = philox4(seed, offset >> 2, site);
  const uint32_t i = (uint32_t)offset & 3u;
  return i == 0 ? p.w[0] : (i == 1 ? p.w[1] : (i == 2 ? p.w[2] : p.w[3]));
}

struct Dropout {
  uint64_t seed;
  const unsigned long long *seed_ptr;  // optional device counter added to `seed` (CUDA-graph replays)
  uint32_t threshold;  // drop iff word < threshold  (threshold = p * 2^32)
  float scale;         // 1 / (1 - p); p == 0 => threshold 0, scale 1
  // call once per thread before apply()/mask(): folds the device-side step counter into the seed
  __device__ __forceinline__ void resolve() {
    if (threshold != 0u && seed_ptr != nullptr) seed += *seed_ptr;
  }
  __host__ __device__ __forceinline__ float apply(float v, uint64_t offset, uint32_t site) const {
    if (threshold == 0u) return v;
    return philox_word(seed, offset, site) < threshold ? 0.f : v * scale;
  }
  __host__ __device__ __forceinline__ float mask(uint64_t offset, uint32_t site) const {
    if (threshold == 0u) return 1.f;
    return philox_word(seed, offset, site) < threshold ? 0.f : scale;
  }
  // masks of the four elements offset4 .. offset4+3 (offset4 % 4 == 0)
  __host__ __device__ __forceinline__ void mask4(uint64_t offset4, uint32_t site, float (&m)[4]) const {
    if (threshold == 0u) { m[0] = m[1] = m[2] = m[3] = 1.f; return; }
    const Philox4 p = philox4(seed, offset4 >> 2, site);
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = p.w[i] < threshold ? 0.f : scale;
  }
};

inline Dropout make_dropout(uint64_t seed, float p, const unsigned long long *seed_ptr = nullptr) {
  Dropout d;
  d.seed = seed;
  d.seed_ptr = seed_ptr;
  if (p <= 0.f) { d.threshold = 0u; d.scale = 1.f; }
  else {
    double t = (double)p * 4294967296.0;
    d.threshold = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
    d.scale = 1.f / (1.f - p);
  }
  return d;
}

__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float v) {
  const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * v * v);
  return cdf + v * pdf;
}

}  // namespace msr3d
