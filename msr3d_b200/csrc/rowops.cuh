// Row-wise helpers shared by the non-GEMM kernels of the spatial transformer (spatial.cu) and the fused
// attention-block kernels (attn_block.cu): one warp per row of kD = 256 values, LayerNorm forward / backward,
// per-column accumulators for parameter gradients.  (transformers.py:250-251, 324-328.)
#pragma once
#include "common.cuh"
#include "rng.cuh"

namespace msr3d {

constexpr int kD = 256;        // hidden size of the spatial encoder (configs/msr3d.yaml:181)
constexpr int kMaxL = 128;     // tokens per scene supported by the attention kernels
constexpr float kLnEps = 1e-5f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------
// LayerNorm helpers: one warp per row of kD = 256 values; lane l owns the 8 CONSECUTIVE columns
// 8l .. 8l+7 (two 16 B accesses per row and, above all, two Philox blocks per dropout site instead
// of eight: the mask of element `offset` is word offset & 3 of Philox block offset >> 2, rng.cuh —
// with the lane-strided mapping every element paid its own ten-round Philox evaluation, which made
// the tail kernels 9-10 us of pure integer arithmetic).
// ---------------------------------------------------------------------------
struct Row8 { float v[8]; };
__device__ __forceinline__ int col8(int lane, int i) { return lane * 8 + i; }
__device__ __forceinline__ Row8 row_load(const float *p, int lane) {
  Row8 r;
  const float4 a = *reinterpret_cast<const float4 *>(p + lane * 8);
  const float4 b = *reinterpret_cast<const float4 *>(p + lane * 8 + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void row_store(float *p, int lane, const Row8 &r) {
  *reinterpret_cast<float4 *>(p + lane * 8) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4 *>(p + lane * 8 + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
// dropout masks of this lane's 8 columns of the row starting at element offset `base` (base % 8 == 0)
__device__ __forceinline__ Row8 drop_mask8(const Dropout &d, size_t base, int lane, unsigned site) {
  Row8 m;
  d.mask8((uint64_t)base + lane * 8, site, m.v);   // one Philox block per eight decisions
  return m;
}
// y = gamma * xhat + beta; returns xhat in `x`, rstd via out param
__device__ __forceinline__ Row8 ln_forward(Row8 &x, const float *gamma, const float *beta, int lane,
                                           float &rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x.v[i];
  const float mean = warp_sum(s) * (1.f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { x.v[i] -= mean; q += x.v[i] * x.v[i]; }
  rstd = rsqrtf(warp_sum(q) * (1.f / kD) + kLnEps);
  Row8 y;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x.v[i] *= rstd;
    y.v[i] = x.v[i] * gamma[col8(lane, i)] + beta[col8(lane, i)];
  }
  return y;
}
// dx = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy * gamma
__device__ __forceinline__ Row8 ln_backward(const Row8 &dy, const Row8 &xhat, const float *gamma,
                                            int lane, float rstd) {
  Row8 g;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    g.v[i] = dy.v[i] * gamma[col8(lane, i)];
    s1 += g.v[i];
    s2 += g.v[i] * xhat.v[i];
  }
  s1 = warp_sum(s1) * (1.f / kD);
  s2 = warp_sum(s2) * (1.f / kD);
  Row8 dx;
#pragma unroll
  for (int i = 0; i < 8; ++i) dx.v[i] = rstd * (g.v[i] - s1 - xhat.v[i] * s2);
  return dx;
}

constexpr int kLnWarps = 8;
// Block-level accumulation of per-column sums (for dgamma/dbeta/bias grads): each warp adds its
// row's contribution into shared partials, the CTA flushes them with one atomic per column.
struct ColAcc {
  float *s;  // [n_acc][kD] shared
  __device__ void zero(int n_acc, int tid) {
    for (int i = tid; i < n_acc * kD; i += kLnWarps * 32) s[i] = 0.f;
  }
  // column c = 8 * lane + i lives at slot i * 32 + lane: lanes hit distinct banks
  __device__ static int slot(int c) { return (c & 7) * 32 + (c >> 3); }
  __device__ void add(int which, int lane, const Row8 &r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&s[which * kD + i * 32 + lane], r.v[i]);
  }
  __device__ float get(int which, int c) const { return s[which * kD + slot(c)]; }
  __device__ void flush(int which, float *dst, int tid) {
    if (dst == nullptr) return;
    for (int c = tid; c < kD; c += kLnWarps * 32) atomicAdd(dst + c, get(which, c));
  }
};

struct LnParams { const float *gamma, *beta; };

}  // namespace msr3d
