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

// Eight keep/drop decisions per Philox block: element `offset` uses 16-bit half (offset & 1) of word
// ((offset >> 1) & 3) of block (offset >> 3); drop iff half < threshold (= p * 2^16, i.e. p is honoured to
// 1/65536).  Round 1 spent one 32-bit word per element: the ten Philox rounds per four elements were a third of
// the instructions of the FFN epilogue (tools/trace_kernels.py).
struct Dropout {
  uint64_t seed;
  const unsigned long long *seed_ptr;  // optional device counter added to `seed` (CUDA-graph replays)
  uint32_t threshold;  // 16-bit: drop iff half-word < threshold; 0 = dropout off
  float scale;         // 1 / (1 - p); p == 0 => threshold 0, scale 1
  // call once per thread before apply()/mask(): folds the device-side step counter into the seed
  __device__ __forceinline__ void resolve() {
    if (threshold != 0u && seed_ptr != nullptr) seed += *seed_ptr;
  }
  __host__ __device__ __forceinline__ static uint32_t half_of(const Philox4 &p, uint32_t sub) {
    const uint32_t wi = sub >> 1;
    const uint32_t w = wi == 0 ? p.w[0] : (wi == 1 ? p.w[1] : (wi == 2 ? p.w[2] : p.w[3]));
    return (sub & 1u) ? (w >> 16) : (w & 0xffffu);
  }
  __host__ __device__ __forceinline__ float mask(uint64_t offset, uint32_t site) const {
    if (threshold == 0u) return 1.f;
    const Philox4 p = philox4(seed, offset >> 3, site);
    return half_of(p, (uint32_t)offset & 7u) < threshold ? 0.f : scale;
  }
  __host__ __device__ __forceinline__ float apply(float v, uint64_t offset, uint32_t site) const {
    return v * mask(offset, site);
  }
  // masks of the eight elements offset8 .. offset8+7 (offset8 % 8 == 0): ONE Philox evaluation
  __host__ __device__ __forceinline__ void mask8(uint64_t offset8, uint32_t site, float (&m)[8]) const {
    if (threshold == 0u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = 1.f;
      return;
    }
    const Philox4 p = philox4(seed, offset8 >> 3, site);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      m[2 * i] = (p.w[i] & 0xffffu) < threshold ? 0.f : scale;
      m[2 * i + 1] = (p.w[i] >> 16) < threshold ? 0.f : scale;
    }
  }
  // masks of the four elements offset4 .. offset4+3 (offset4 % 4 == 0): half a block
  __host__ __device__ __forceinline__ void mask4(uint64_t offset4, uint32_t site, float (&m)[4]) const {
    if (threshold == 0u) { m[0] = m[1] = m[2] = m[3] = 1.f; return; }
    const Philox4 p = philox4(seed, offset4 >> 3, site);
    const bool hi = (offset4 & 4u) != 0;
    const uint32_t w0 = hi ? p.w[2] : p.w[0], w1 = hi ? p.w[3] : p.w[1];
    m[0] = (w0 & 0xffffu) < threshold ? 0.f : scale;
    m[1] = (w0 >> 16) < threshold ? 0.f : scale;
    m[2] = (w1 & 0xffffu) < threshold ? 0.f : scale;
    m[3] = (w1 >> 16) < threshold ? 0.f : scale;
  }
};

inline Dropout make_dropout(uint64_t seed, float p, const unsigned long long *seed_ptr = nullptr) {
  Dropout d;
  d.seed = seed;
  d.seed_ptr = seed_ptr;
  if (p <= 0.f) { d.threshold = 0u; d.scale = 1.f; }
  else {
    double t = (double)p * 65536.0 + 0.5;
    d.threshold = t >= 65535.0 ? 0xffffu : (t < 1.0 ? 1u : (uint32_t)t);
    d.scale = 1.f / (1.f - p);
  }
  return d;
}

// GELU(erf) and its derivative from ONE exponential:  E = exp(-x^2 / 2) is both the Gaussian density
// (up to 1/sqrt(2 pi)) and the tail factor of erf(x / sqrt 2) = 1 - poly(t) E, t = 1 / (1 + p |x| / sqrt 2)
// (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7: fp32 rounding level).  ~15 instructions for both values;
// erff + expf (two branches, ~60 instructions) made the FFN epilogue the longest phase of the kernel.
__device__ __forceinline__ void gelu_cdf_pdf(float x, float &cdf, float &pdf) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.f));
  const float e = __expf(-z * z);
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float tail = 0.5f * poly * t * e;          // 0.5 * (1 - erf(z))
  cdf = x >= 0.f ? 1.f - tail : tail;
  pdf = 0.39894228040143267794f * e;
}

__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float v) {
  const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * v * v);
  return cdf + v * pdf;
}

}  // namespace msr3d
