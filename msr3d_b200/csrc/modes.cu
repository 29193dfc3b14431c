// Situation modes with an agent ("anchor") token in front of the objects: as_object and as_object_add_loc
// (model/ose3d_situation.py:334-363, 385-395, 429-431).  The transformer layers are the ones of the
// shipped mode with L + 1 tokens per scene; what differs is the input assembly and its backward:
//   all_loc   = [ (anchor_locs, anchor_size) ; obj_locs ]                       (B, L+1, 6)
//   key_pad   = [ 0 ; !obj_masks ]                                             (B, L+1)
//   x_in0 row = projected embedding + object_orientation_feat + type_embedding[0] + qpos        (object rows)
//             = anchor_feat + orientation_encoder(fourier(anchor_orientation)) + type_embedding[1] + qpos   (row 0)
//   qpos      = LN(Linear_6(all_loc))  (as_object: loc_layers[0])  |  the Fourier + size encoders without the
//               agent-frame transform (as_object_add_loc: geometry_kernel use_agent_frame = 0 + qpos_fwd_kernel)
// All kernels: one warp per token row of kD = 256 values (rowops.cuh), deterministic parameter gradients apart
// from the per-column shared-memory accumulators the shipped mode uses too.
#include "common.cuh"
#include "rng.cuh"
#include "rowops.cuh"

namespace msr3d {

constexpr int kOriFeat = 84;   // generate_fourier_features of a quaternion: 4 + 4 * 10 * 2 (ose3d_situation.py:31-59)

// Fourier features of the scene's anchor quaternion, [q(4) | sin(pi q_d f_j) d-major (40) | cos (40)], by one warp.
__device__ __forceinline__ void anchor_fourier(const float *q, int lane, float *f /* shared [84] */) {
  for (int e = lane; e < 40; e += 32) {
    const int d = e / 10, j = e % 10;
    const float kPi = 3.14159265358979323846f;
    const float step = 14.f / 9.f;     // torch.linspace(1, 15, 10), as in geometry_kernel
    const float fr = j < 5 ? 1.f + j * step : 15.f - (9 - j) * step;
    float sn, cs;
    sincosf(kPi * (q[d] * fr), &sn, &cs);
    f[4 + e] = sn;
    f[44 + e] = cs;
  }
  if (lane < 4) f[lane] = q[lane];
  __syncwarp();
}

__global__ void __launch_bounds__(256)
anchor_inputs_kernel(const float *__restrict__ obj_locs, const unsigned char *__restrict__ obj_masks,
                     const float *__restrict__ anchor_locs, const float *__restrict__ anchor_size, int B, int L,
                     float *__restrict__ all_loc, unsigned char *__restrict__ key_pad) {
  pdl_sync();
  const int Lt = L + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;    // token
  if (i >= B * Lt) return;
  const int b = i / Lt, t = i % Lt;
  float *o = all_loc + (size_t)i * 6;
  if (t == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { o[d] = anchor_locs[b * 3 + d]; o[3 + d] = anchor_size[d]; }
    key_pad[i] = 0;
  } else {
    const float *p = obj_locs + ((size_t)b * L + t - 1) * 6;
#pragma unroll
    for (int d = 0; d < 6; ++d) o[d] = p[d];
    key_pad[i] = obj_masks[b * L + t - 1] ? 0 : 1;
  }
}

// qpos = LN(W6 loc + b6) for every token (as_object: loc_layers[0], ose3d_situation.py:396); x0 (zero-filled layer-0
// input) receives it as its first term.
__global__ void __launch_bounds__(kLnWarps * 32)
loc_layer_fwd_kernel(const float *__restrict__ all_loc, int M, const float *__restrict__ W6,
                     const float *__restrict__ b6, LnParams ln, float *__restrict__ xhat, float *__restrict__ rstd,
                     float *__restrict__ qpos, float *__restrict__ x0, const float *__restrict__ extra_a,
                     const float *__restrict__ extra_b) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  const size_t base = (size_t)row * kD;
  float l[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) l[d] = all_loc[(size_t)row * 6 + d];
  Row8 a;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = col8(lane, i);
    float s = b6[c];
#pragma unroll
    for (int d = 0; d < 6; ++d) s += W6[c * 6 + d] * l[d];
    a.v[i] = s;
  }
  float r;
  const Row8 y = ln_forward(a, ln.gamma, ln.beta, lane, r);
  row_store(xhat + base, lane, a);
  row_store(qpos + base, lane, y);
  Row8 y0 = y;            // layer-0 input: + the constant terms of the objects when there is no agent row to tell apart
#pragma unroll
  for (int i = 0; i < 8; ++i)
    y0.v[i] += (extra_a ? extra_a[col8(lane, i)] : 0.f) + (extra_b ? extra_b[col8(lane, i)] : 0.f);
  row_store(x0 + base, lane, y0);
  if (lane == 0) rstd[row] = r;
}

// Backward: g = sum of the per-layer gradients wrt qpos; dgamma, dbeta, dW6 (256 x 6), db6.
__global__ void __launch_bounds__(kLnWarps * 32)
loc_layer_bwd_kernel(const float *__restrict__ dq, const float *__restrict__ dq1, const float *__restrict__ dq2,
                     const float *__restrict__ all_loc, int M, LnParams ln, const float *__restrict__ xhat,
                     const float *__restrict__ rstd, float *__restrict__ dgamma, float *__restrict__ dbeta,
                     float *__restrict__ dW6, float *__restrict__ db6) {
  pdl_sync();
  __shared__ float s_acc[9 * kD];
  ColAcc acc{s_acc};
  acc.zero(9, threadIdx.x);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row < M) {
    const size_t base = (size_t)row * kD;
    Row8 g = row_load(dq + base, lane);
    if (dq1 != nullptr) {
      const Row8 e = row_load(dq1 + base, lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) g.v[i] += e.v[i];
    }
    if (dq2 != nullptr) {
      const Row8 e = row_load(dq2 + base, lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) g.v[i] += e.v[i];
    }
    const Row8 h = row_load(xhat + base, lane);
    Row8 t;
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = g.v[i] * h.v[i];
    acc.add(0, lane, t);
    acc.add(1, lane, g);
    const Row8 d1 = ln_backward(g, h, ln.gamma, lane, rstd[row]);
    acc.add(2, lane, d1);
#pragma unroll
    for (int d = 0; d < 6; ++d) {
      const float ld = all_loc[(size_t)row * 6 + d];
#pragma unroll
      for (int i = 0; i < 8; ++i) t.v[i] = d1.v[i] * ld;
      acc.add(3 + d, lane, t);
    }
  }
  __syncthreads();
  acc.flush(0, dgamma, threadIdx.x);
  acc.flush(1, dbeta, threadIdx.x);
  acc.flush(2, db6, threadIdx.x);
  for (int c = threadIdx.x; c < kD; c += kLnWarps * 32)
#pragma unroll
    for (int d = 0; d < 6; ++d) atomicAdd(dW6 + c * 6 + d, acc.get(3 + d, c));
}

// x0[token] += the mode's constant terms (+ the projected embedding for object rows).  Runs after the projection
// GEMM and the query-position kernels (stream order), so a plain read-modify-write is enough.
__global__ void __launch_bounds__(kLnWarps * 32)
anchor_assemble_kernel(float *__restrict__ x0, const float *__restrict__ proj, const float *__restrict__ ori_feat,
                       const float *__restrict__ type_w, const float *__restrict__ anchor_feat,
                       const float *__restrict__ W_ori, const float *__restrict__ b_ori,
                       const float *__restrict__ anchor_ori, int B, int L) {
  pdl_sync();
  __shared__ float s_f[kLnWarps][kOriFeat];
  const int Lt = L + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * kLnWarps + warp;
  if (row >= B * Lt) return;
  const int b = row / Lt, t = row % Lt;
  const size_t base = (size_t)row * kD;
  Row8 x = row_load(x0 + base, lane);
  if (t > 0) {
    const Row8 p = row_load(proj + ((size_t)b * L + t - 1) * kD, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = col8(lane, i);
      x.v[i] += (p.v[i] + ori_feat[c]) + type_w[c];
    }
  } else {
    float *f = s_f[warp];
    anchor_fourier(anchor_ori + (size_t)b * 4, lane, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = col8(lane, i);
      float s = b_ori[c];
      for (int j = 0; j < kOriFeat; ++j) s += W_ori[(size_t)c * kOriFeat + j] * f[j];
      x.v[i] += (anchor_feat[c] + s) + type_w[kD + c];
    }
  }
  row_store(x0 + base, lane, x);
}

// Backward of the assembly.  CTAs 0 .. n-1: object rows of dx0 -> compact dproj (B*L, 256) (operand of the projection's
// weight-gradient GEMM and of the column sums behind its bias / object_orientation_feat / type_embedding[0]).
// Last CTA (thread = column): anchor rows -> d anchor_feat, d type_embedding[1], d orientation_encoder (weight, bias),
// summed over the scenes in a fixed order.
__global__ void __launch_bounds__(kLnWarps * 32)
anchor_disassemble_kernel(const float *__restrict__ dx0, const float *__restrict__ anchor_ori, int B, int L,
                          float *__restrict__ dproj, float *__restrict__ d_anchor_feat, float *__restrict__ d_type1,
                          float *__restrict__ dW_ori, float *__restrict__ db_ori) {
  pdl_sync();
  const int Lt = L + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x + 1 < gridDim.x) {
    const int r = blockIdx.x * kLnWarps + warp;      // object row
    if (r >= B * L) return;
    const int b = r / L, o = r % L;
    row_store(dproj + (size_t)r * kD, lane, row_load(dx0 + ((size_t)b * Lt + 1 + o) * kD, lane));
    return;
  }
  __shared__ float s_f[kOriFeat];
  const int c = threadIdx.x;             // kLnWarps * 32 = 256 = kD
  float gsum = 0.f;
  for (int b = 0; b < B; ++b) {
    __syncthreads();
    if (warp == 0) anchor_fourier(anchor_ori + (size_t)b * 4, lane, s_f);
    __syncthreads();
    const float g = dx0[(size_t)b * Lt * kD + c];
    gsum += g;
    for (int j = 0; j < kOriFeat; ++j) dW_ori[(size_t)c * kOriFeat + j] += g * s_f[j];
  }
  d_anchor_feat[c] += gsum;
  d_type1[c] += gsum;
  db_ori[c] += gsum;
}

// ---------------------------------------------------------------------------
// as_embedding (ose3d_situation.py:385-393): the situation enters as one embedding per scene,
//   s_b = loc_embedding_encoder(fourier(anchor_loc_b)) + orientation_encoder(fourier(anchor_orientation_b)),
// added to the query position of EVERY object of the scene (no agent token, no agent-frame transform).
// Forward: one CTA per scene, thread = column; s_b goes into qpos (read again by the later layers) and, with
// red.add, into the layer-0 input that the projection GEMM is reducing into at the same time.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float block_sum256(float v, float *s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += s_red[i];
  return t;
}

// [p(3) | sin(pi p_d f_j) d-major (30) | cos (30)] of a position, by one warp (geometry_kernel's layout)
__device__ __forceinline__ void loc_fourier(const float *p, int lane, float *f /* shared [64] */) {
  if (lane < 30) {
    const int d = lane / 10, j = lane % 10;
    const float kPi = 3.14159265358979323846f;
    const float step = 14.f / 9.f;
    const float fr = j < 5 ? 1.f + j * step : 15.f - (9 - j) * step;
    float sn, cs;
    sincosf(kPi * (p[d] * fr), &sn, &cs);
    f[3 + lane] = sn;
    f[33 + lane] = cs;
  }
  if (lane < 3) f[lane] = p[lane];
  if (lane == 31) f[63] = 0.f;
  __syncwarp();
}

__global__ void __launch_bounds__(256)
situation_embed_fwd_kernel(const float *__restrict__ anchor_locs, const float *__restrict__ anchor_ori, int L,
                           const float *__restrict__ Wl, const float *__restrict__ bl, LnParams lnl,
                           const float *__restrict__ W_ori, const float *__restrict__ b_ori,
                           float *__restrict__ xhat_sit, float *__restrict__ rstd_sit, float *__restrict__ qpos,
                           float *__restrict__ x0, float *__restrict__ s_out) {
  pdl_sync();
  __shared__ float f_loc[64], f_ori[kOriFeat], s_red[8];
  const int b = blockIdx.x, c = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) loc_fourier(anchor_locs + (size_t)b * 3, lane, f_loc);
  if (warp == 1) anchor_fourier(anchor_ori + (size_t)b * 4, lane, f_ori);
  __syncthreads();
  float h = bl[c];
  for (int j = 0; j < 63; ++j) h += Wl[(size_t)c * 63 + j] * f_loc[j];
  const float mean = block_sum256(h, s_red) * (1.f / kD);
  const float xc = h - mean;
  const float rstd = rsqrtf(block_sum256(xc * xc, s_red) * (1.f / kD) + kLnEps);
  const float xh = xc * rstd;
  float s = xh * lnl.gamma[c] + lnl.beta[c] + b_ori[c];
  for (int j = 0; j < kOriFeat; ++j) s += W_ori[(size_t)c * kOriFeat + j] * f_ori[j];
  xhat_sit[(size_t)b * kD + c] = xh;
  if (c == 0) rstd_sit[b] = rstd;
  if (s_out != nullptr) s_out[(size_t)b * kD + c] = s;      // as_cross_attention: the vector itself is the product
  if (qpos == nullptr) return;
  for (int t = 0; t < L; ++t) {
    const size_t i = ((size_t)b * L + t) * kD + c;
    qpos[i] += s;
    atomicAdd(x0 + i, s);
  }
}

// Backward: ONE CTA walks the scenes in order (deterministic); parameter gradients accumulate with atomics because
// loc_embedding_encoder also receives the objects' contributions (qpos_bwd_kernel, the dWloc GEMM).
__global__ void __launch_bounds__(256)
situation_embed_bwd_kernel(const float *__restrict__ dq, const float *__restrict__ dq1, const float *__restrict__ dq2,
                           const float *__restrict__ anchor_locs, const float *__restrict__ anchor_ori, int B, int L,
                           const float *__restrict__ gamma_l, const float *__restrict__ xhat_sit,
                           const float *__restrict__ rstd_sit, float *__restrict__ dWl, float *__restrict__ dbl,
                           float *__restrict__ dgamma_l, float *__restrict__ dbeta_l, float *__restrict__ dW_ori,
                           float *__restrict__ db_ori) {
  pdl_sync();
  __shared__ float f_loc[64], f_ori[kOriFeat], s_red[8];
  const int c = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float a_gamma = 0.f, a_beta = 0.f, a_bl = 0.f;
  for (int b = 0; b < B; ++b) {
    __syncthreads();
    if (warp == 0) loc_fourier(anchor_locs + (size_t)b * 3, lane, f_loc);
    if (warp == 1) anchor_fourier(anchor_ori + (size_t)b * 4, lane, f_ori);
    __syncthreads();
    float g = 0.f;
    for (int t = 0; t < L; ++t) {
      const size_t i = ((size_t)b * L + t) * kD + c;
      g += dq[i];
      if (dq1 != nullptr) g += dq1[i];
      if (dq2 != nullptr) g += dq2[i];
    }
    const float xh = xhat_sit[(size_t)b * kD + c];
    a_gamma += g * xh;
    a_beta += g;
    for (int j = 0; j < kOriFeat; ++j) atomicAdd(dW_ori + (size_t)c * kOriFeat + j, g * f_ori[j]);
    const float gg = g * gamma_l[c];
    const float s1 = block_sum256(gg, s_red) * (1.f / kD);
    const float s2 = block_sum256(gg * xh, s_red) * (1.f / kD);
    const float dh = rstd_sit[b] * (gg - s1 - xh * s2);
    a_bl += dh;
    for (int j = 0; j < 63; ++j) atomicAdd(dWl + (size_t)c * 63 + j, dh * f_loc[j]);
  }
  atomicAdd(dgamma_l + c, a_gamma);
  atomicAdd(dbeta_l + c, a_beta);
  atomicAdd(db_ori + c, a_beta);
  atomicAdd(dbl + c, a_bl);
}

// ---------------------------------------------------------------------------
// as_cross_attention (ose3d_situation.py:411-418, CrossAttentionLayer transformers.py:22-73, prenorm): in front of every
// spatial layer the objects cross-attend to the situation embedding REPEATED for every object — all N keys and
// values of a scene are the same vector s_b.  The softmax is therefore uniform whatever the queries are, the
// attention output of head h is (sum of its dropped-out probabilities) * v_h with v = W_v s_b + b_v, and the
// gradients of W_q, W_k, b_q, b_k and norm1 are exactly zero (the reference's autograd yields 1e-7 noise).  What is
// left per layer:
//   u[b, n, :] = b_o + sum_h scale[b, h, n] * P[b, h, :],   P[b, h, c] = sum_{k in head h} W_o[c, k] v_b[k]
//   xa = x + drop_2(u);   t3 = LN_3(xa);   x' = xa + drop_3(W_2 drop(gelu(W_1 t3 + b_1)) + b_2)
// scale[b, h, n] = mean over the keys of the attention-dropout mask (1 in eval mode).  The FFN runs on the fused
// pair kernels of csrc/gemm3x.cu.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cross_vec_fwd_kernel(const float *__restrict__ s, const float *__restrict__ Wv, const float *__restrict__ bv,
                     const float *__restrict__ Wo, float *__restrict__ v_out, float *__restrict__ P) {
  pdl_sync();
  __shared__ float s_s[kD], s_v[kD];
  const int b = blockIdx.x, c = threadIdx.x;
  s_s[c] = s[(size_t)b * kD + c];
  __syncthreads();
  float v = bv[c];
  for (int j = 0; j < kD; ++j) v += Wv[(size_t)c * kD + j] * s_s[j];
  s_v[c] = v;
  v_out[(size_t)b * kD + c] = v;
  __syncthreads();
  for (int h = 0; h < 8; ++h) {
    float a = 0.f;
    for (int k = h * 32; k < h * 32 + 32; ++k) a += Wo[(size_t)c * kD + k] * s_v[k];
    P[((size_t)b * 8 + h) * kD + c] = a;
  }
}

__global__ void __launch_bounds__(kLnWarps * 32)
cross_residual_ln_fwd_kernel(const float *__restrict__ x, const float *__restrict__ P, const float *__restrict__ bo,
                             int M, int L, LnParams ln3, Dropout drop, unsigned site_attn, unsigned site_2,
                             float *__restrict__ scale_out, float *__restrict__ xa, float *__restrict__ xhat3,
                             float *__restrict__ rstd3, float *__restrict__ t3) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  drop.resolve();
  const int b = row / L;
  const size_t base = (size_t)row * kD;
  float sc[8];
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    float a = 0.f;
    for (int t = lane; t < L; t += 32) a += drop.mask((uint64_t)((size_t)row * 8 + h) * L + t, site_attn);
    sc[h] = warp_sum(a) / (float)L;
  }
  if (lane < 8) scale_out[(size_t)row * 8 + lane] = sc[lane];
  const Row8 xr = row_load(x + base, lane);
  const Row8 m2 = drop_mask8(drop, base, lane, site_2);
  Row8 u;
#pragma unroll
  for (int i = 0; i < 8; ++i) u.v[i] = bo[col8(lane, i)];
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    const Row8 p = row_load(P + ((size_t)b * 8 + h) * kD, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) u.v[i] += sc[h] * p.v[i];
  }
  Row8 w;
#pragma unroll
  for (int i = 0; i < 8; ++i) w.v[i] = xr.v[i] + u.v[i] * m2.v[i];
  row_store(xa + base, lane, w);
  float r;
  const Row8 y = ln_forward(w, ln3.gamma, ln3.beta, lane, r);
  row_store(xhat3 + base, lane, w);
  row_store(t3 + base, lane, y);
  if (lane == 0) rstd3[row] = r;
}

// x' = xa + drop_3(f)
__global__ void __launch_bounds__(kLnWarps * 32)
cross_ffn_residual_fwd_kernel(const float *__restrict__ xa, const float *__restrict__ f, int M, Dropout drop,
                              unsigned site_3, float *__restrict__ out) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  drop.resolve();
  const size_t base = (size_t)row * kD;
  const Row8 a = row_load(xa + base, lane), ff = row_load(f + base, lane), m3 = drop_mask8(drop, base, lane, site_3);
  Row8 o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = a.v[i] + ff.v[i] * m3.v[i];
  row_store(out + base, lane, o);
}

// df = drop_3 mask * dx' (+ column sums -> gradient of linear2.bias)
__global__ void __launch_bounds__(kLnWarps * 32)
cross_ffn_residual_bwd_kernel(const float *__restrict__ dx, int M, Dropout drop, unsigned site_3, float *__restrict__ df,
                              float *__restrict__ db2) {
  pdl_sync();
  drop.resolve();
  __shared__ float s_acc[kD];
  ColAcc acc{s_acc};
  acc.zero(1, threadIdx.x);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row < M) {
    const size_t base = (size_t)row * kD;
    const Row8 g = row_load(dx + base, lane), m3 = drop_mask8(drop, base, lane, site_3);
    Row8 o;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = g.v[i] * m3.v[i];
    acc.add(0, lane, o);
    row_store(df + base, lane, o);
  }
  __syncthreads();
  acc.flush(0, db2, threadIdx.x);
}

// dxa = dx' + LN_3 backward(dt3);  g_attn = drop_2 mask * dxa;  dgamma_3, dbeta_3, d b_o (column sums of g_attn)
__global__ void __launch_bounds__(kLnWarps * 32)
cross_ln_bwd_kernel(const float *__restrict__ dt3, const float *__restrict__ dx, int M, LnParams ln3, Dropout drop,
                    unsigned site_2, const float *__restrict__ xhat3, const float *__restrict__ rstd3,
                    float *__restrict__ dxa, float *__restrict__ g_attn, float *__restrict__ dgamma3,
                    float *__restrict__ dbeta3, float *__restrict__ dbo) {
  pdl_sync();
  drop.resolve();
  __shared__ float s_acc[3 * kD];
  ColAcc acc{s_acc};
  acc.zero(3, threadIdx.x);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row < M) {
    const size_t base = (size_t)row * kD;
    const Row8 g = row_load(dt3 + base, lane), h = row_load(xhat3 + base, lane), d = row_load(dx + base, lane);
    const Row8 m2 = drop_mask8(drop, base, lane, site_2);
    Row8 t;
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = g.v[i] * h.v[i];
    acc.add(0, lane, t);
    acc.add(1, lane, g);
    const Row8 dw = ln_backward(g, h, ln3.gamma, lane, rstd3[row]);
    Row8 da, ga;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      da.v[i] = d.v[i] + dw.v[i];
      ga.v[i] = da.v[i] * m2.v[i];
    }
    acc.add(2, lane, ga);
    row_store(dxa + base, lane, da);
    row_store(g_attn + base, lane, ga);
  }
  __syncthreads();
  acc.flush(0, dgamma3, threadIdx.x);
  acc.flush(1, dbeta3, threadIdx.x);
  acc.flush(2, dbo, threadIdx.x);
}

// One CTA per scene, thread = column.  dP[h, c] = sum_n scale[n, h] g_attn[n, c];  dW_o[c, k] += dP[h(k), c] v[k];
// dv[k] = sum_c dP[h(k), c] W_o[c, k];  db_v += dv;  dW_v[k, j] += dv[k] s[j];  ds[b, j] += sum_k dv[k] W_v[k, j].
__global__ void __launch_bounds__(256)
cross_vec_bwd_kernel(const float *__restrict__ g_attn, const float *__restrict__ scale, int L,
                     const float *__restrict__ v, const float *__restrict__ s, const float *__restrict__ Wv,
                     const float *__restrict__ Wo, float *__restrict__ dWo, float *__restrict__ dWv,
                     float *__restrict__ dbv, float *__restrict__ ds) {
  pdl_sync();
  __shared__ float s_dP[8][kD], s_v[kD], s_s[kD], s_dv[kD];
  const int b = blockIdx.x, c = threadIdx.x;
  float dP[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int n = 0; n < L; ++n) {
    const size_t row = (size_t)b * L + n;
    const float g = g_attn[row * kD + c];
#pragma unroll
    for (int h = 0; h < 8; ++h) dP[h] += scale[row * 8 + h] * g;
  }
#pragma unroll
  for (int h = 0; h < 8; ++h) s_dP[h][c] = dP[h];
  s_v[c] = v[(size_t)b * kD + c];
  s_s[c] = s[(size_t)b * kD + c];
  __syncthreads();
  for (int k = 0; k < kD; ++k) atomicAdd(dWo + (size_t)c * kD + k, dP[k >> 5] * s_v[k]);
  float dv = 0.f;                                   // thread = k from here on
  for (int cc = 0; cc < kD; ++cc) dv += s_dP[c >> 5][cc] * Wo[(size_t)cc * kD + c];
  s_dv[c] = dv;
  atomicAdd(dbv + c, dv);
  for (int j = 0; j < kD; ++j) atomicAdd(dWv + (size_t)c * kD + j, dv * s_s[j]);
  __syncthreads();
  float d = 0.f;                                    // thread = j
  for (int k = 0; k < kD; ++k) d += s_dv[k] * Wv[(size_t)k * kD + c];
  ds[(size_t)b * kD + c] += d;
}

}  // namespace msr3d

using namespace msr3d;

MSR3D_API int msr3d_anchor_inputs(const float *obj_locs, const unsigned char *obj_masks, const float *anchor_locs,
                                  const float *anchor_size, int B, int L, float *all_loc, unsigned char *key_pad,
                                  msr3d_stream_t stream) {
  MSR3D_REQUIRE(obj_locs && obj_masks && anchor_locs && anchor_size && all_loc && key_pad, "anchor_inputs: null pointer");
  MSR3D_REQUIRE(B >= 0 && L >= 1 && L + 1 <= kMaxL, "anchor_inputs: L=%d must be in [1,%d]", L, kMaxL - 1);
  if (B == 0) return MSR3D_OK;
  launch_pdl(anchor_inputs_kernel, dim3(ceil_div(B * (L + 1), 256)), dim3(256), 0, (cudaStream_t)stream, obj_locs,
             obj_masks, anchor_locs, anchor_size, B, L, all_loc, key_pad);
  return check_launch("anchor_inputs");
}

MSR3D_API int msr3d_loc_layer_fwd_f32(const float *all_loc, int M, const float *W6, const float *b6,
                                      const float *gamma, const float *beta, float *xhat, float *rstd, float *qpos,
                                      float *x0, const float *extra_a, const float *extra_b, msr3d_stream_t stream) {
  MSR3D_REQUIRE(all_loc && W6 && b6 && gamma && beta && xhat && rstd && qpos && x0, "loc_layer_fwd: null pointer");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(loc_layer_fwd_kernel, dim3(ceil_div(M, kLnWarps)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, all_loc, M,
             W6, b6, LnParams{gamma, beta}, xhat, rstd, qpos, x0, extra_a, extra_b);
  return check_launch("loc_layer_fwd");
}

MSR3D_API int msr3d_loc_layer_bwd_f32(const float *dq, const float *dq1, const float *dq2, const float *all_loc, int M,
                                      const float *gamma, const float *xhat, const float *rstd, float *dgamma,
                                      float *dbeta, float *dW6, float *db6, msr3d_stream_t stream) {
  MSR3D_REQUIRE(dq && all_loc && gamma && xhat && rstd && dgamma && dbeta && dW6 && db6, "loc_layer_bwd: null pointer");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(loc_layer_bwd_kernel, dim3(ceil_div(M, kLnWarps)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, dq, dq1,
             dq2, all_loc, M, LnParams{gamma, nullptr}, xhat, rstd, dgamma, dbeta, dW6, db6);
  return check_launch("loc_layer_bwd");
}

MSR3D_API int msr3d_anchor_assemble_f32(float *x0, const float *proj, const float *ori_feat, const float *type_w,
                                        const float *anchor_feat, const float *W_ori, const float *b_ori,
                                        const float *anchor_ori, int B, int L, msr3d_stream_t stream) {
  MSR3D_REQUIRE(x0 && proj && ori_feat && type_w && anchor_feat && W_ori && b_ori && anchor_ori,
                "anchor_assemble: null pointer");
  if (B <= 0) return MSR3D_OK;
  launch_pdl(anchor_assemble_kernel, dim3(ceil_div(B * (L + 1), kLnWarps)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream,
             x0, proj, ori_feat, type_w, anchor_feat, W_ori, b_ori, anchor_ori, B, L);
  return check_launch("anchor_assemble");
}

MSR3D_API int msr3d_anchor_disassemble_f32(const float *dx0, const float *anchor_ori, int B, int L, float *dproj,
                                           float *d_anchor_feat, float *d_type1, float *dW_ori, float *db_ori,
                                           msr3d_stream_t stream) {
  MSR3D_REQUIRE(dx0 && anchor_ori && dproj && d_anchor_feat && d_type1 && dW_ori && db_ori,
                "anchor_disassemble: null pointer");
  if (B <= 0) return MSR3D_OK;
  launch_pdl(anchor_disassemble_kernel, dim3(ceil_div(B * L, kLnWarps) + 1), dim3(kLnWarps * 32), 0,
             (cudaStream_t)stream, dx0, anchor_ori, B, L, dproj, d_anchor_feat, d_type1, dW_ori, db_ori);
  return check_launch("anchor_disassemble");
}

MSR3D_API int msr3d_situation_embed_fwd_f32(const float *anchor_locs, const float *anchor_ori, int B, int L,
                                            const float *Wl, const float *bl, const float *gamma_l, const float *beta_l,
                                            const float *W_ori, const float *b_ori, float *xhat_sit, float *rstd_sit,
                                            float *qpos, float *x0, float *s_out, msr3d_stream_t stream) {
  MSR3D_REQUIRE(anchor_locs && anchor_ori && Wl && bl && gamma_l && beta_l && W_ori && b_ori && xhat_sit && rstd_sit,
                "situation_embed_fwd: null pointer");
  MSR3D_REQUIRE((qpos != nullptr) == (x0 != nullptr) && (qpos != nullptr || s_out != nullptr),
                "situation_embed_fwd: qpos and x0 come together; without them s_out is the output");
  if (B <= 0 || L <= 0) return MSR3D_OK;
  launch_pdl(situation_embed_fwd_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, anchor_locs, anchor_ori, L, Wl, bl,
             LnParams{gamma_l, beta_l}, W_ori, b_ori, xhat_sit, rstd_sit, qpos, x0, s_out);
  return check_launch("situation_embed_fwd");
}

MSR3D_API int msr3d_situation_embed_bwd_f32(const float *dq, const float *dq1, const float *dq2, const float *anchor_locs,
                                            const float *anchor_ori, int B, int L, const float *gamma_l,
                                            const float *xhat_sit, const float *rstd_sit, float *dWl, float *dbl,
                                            float *dgamma_l, float *dbeta_l, float *dW_ori, float *db_ori,
                                            msr3d_stream_t stream) {
  MSR3D_REQUIRE(dq && anchor_locs && anchor_ori && gamma_l && xhat_sit && rstd_sit && dWl && dbl && dgamma_l && dbeta_l &&
                    dW_ori && db_ori, "situation_embed_bwd: null pointer");
  if (B <= 0 || L <= 0) return MSR3D_OK;
  launch_pdl(situation_embed_bwd_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, dq, dq1, dq2, anchor_locs, anchor_ori,
             B, L, gamma_l, xhat_sit, rstd_sit, dWl, dbl, dgamma_l, dbeta_l, dW_ori, db_ori);
  return check_launch("situation_embed_bwd");
}

static inline Dropout mk_drop(unsigned long long seed, float p, const unsigned long long *sp) { return make_dropout(seed, p, sp); }

MSR3D_API int msr3d_cross_vec_fwd_f32(const float *s, const float *Wv, const float *bv, const float *Wo, int B, float *v,
                                      float *P, msr3d_stream_t stream) {
  MSR3D_REQUIRE(s && Wv && bv && Wo && v && P, "cross_vec_fwd: null pointer");
  if (B <= 0) return MSR3D_OK;
  launch_pdl(cross_vec_fwd_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, s, Wv, bv, Wo, v, P);
  return check_launch("cross_vec_fwd");
}

MSR3D_API int msr3d_cross_residual_ln_fwd_f32(const float *x, const float *P, const float *bo, int M, int L,
                                              const float *gamma3, const float *beta3, unsigned long long seed,
                                              const unsigned long long *seed_ptr, float p, unsigned site_attn,
                                              unsigned site_2, float *scale, float *xa, float *xhat3, float *rstd3,
                                              float *t3, msr3d_stream_t stream) {
  MSR3D_REQUIRE(x && P && bo && gamma3 && beta3 && scale && xa && xhat3 && rstd3 && t3, "cross_residual_ln_fwd: null pointer");
  MSR3D_REQUIRE(L >= 1 && M % L == 0 && p >= 0.f && p < 1.f, "cross_residual_ln_fwd: bad sizes / p");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(cross_residual_ln_fwd_kernel, dim3(ceil_div(M, kLnWarps)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, x, P, bo,
             M, L, LnParams{gamma3, beta3}, mk_drop(seed, p, seed_ptr), site_attn, site_2, scale, xa, xhat3, rstd3, t3);
  return check_launch("cross_residual_ln_fwd");
}

MSR3D_API int msr3d_cross_ffn_residual_fwd_f32(const float *xa, const float *f, int M, unsigned long long seed,
                                               const unsigned long long *seed_ptr, float p, unsigned site_3, float *out,
                                               msr3d_stream_t stream) {
  MSR3D_REQUIRE(xa && f && out && p >= 0.f && p < 1.f, "cross_ffn_residual_fwd: bad args");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(cross_ffn_residual_fwd_kernel, dim3(ceil_div(M, kLnWarps)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, xa, f,
             M, mk_drop(seed, p, seed_ptr), site_3, out);
  return check_launch("cross_ffn_residual_fwd");
}

MSR3D_API int msr3d_cross_ffn_residual_bwd_f32(const float *dx, int M, unsigned long long seed,
                                               const unsigned long long *seed_ptr, float p, unsigned site_3, float *df,
                                               float *db2, msr3d_stream_t stream) {
  MSR3D_REQUIRE(dx && df && db2 && p >= 0.f && p < 1.f, "cross_ffn_residual_bwd: bad args");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(cross_ffn_residual_bwd_kernel, dim3(ceil_div(M, kLnWarps)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, dx, M,
             mk_drop(seed, p, seed_ptr), site_3, df, db2);
  return check_launch("cross_ffn_residual_bwd");
}

MSR3D_API int msr3d_cross_ln_bwd_f32(const float *dt3, const float *dx, int M, const float *gamma3,
                                     unsigned long long seed, const unsigned long long *seed_ptr, float p, unsigned site_2,
                                     const float *xhat3, const float *rstd3, float *dxa, float *g_attn, float *dgamma3,
                                     float *dbeta3, float *dbo, msr3d_stream_t stream) {
  MSR3D_REQUIRE(dt3 && dx && gamma3 && xhat3 && rstd3 && dxa && g_attn && dgamma3 && dbeta3 && dbo, "cross_ln_bwd: null pointer");
  MSR3D_REQUIRE(p >= 0.f && p < 1.f, "cross_ln_bwd: bad p");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(cross_ln_bwd_kernel, dim3(ceil_div(M, kLnWarps)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, dt3, dx, M,
             LnParams{gamma3, nullptr}, mk_drop(seed, p, seed_ptr), site_2, xhat3, rstd3, dxa, g_attn, dgamma3, dbeta3, dbo);
  return check_launch("cross_ln_bwd");
}

MSR3D_API int msr3d_cross_vec_bwd_f32(const float *g_attn, const float *scale, int B, int L, const float *v, const float *s,
                                      const float *Wv, const float *Wo, float *dWo, float *dWv, float *dbv, float *ds,
                                      msr3d_stream_t stream) {
  MSR3D_REQUIRE(g_attn && scale && v && s && Wv && Wo && dWo && dWv && dbv && ds, "cross_vec_bwd: null pointer");
  if (B <= 0 || L <= 0) return MSR3D_OK;
  launch_pdl(cross_vec_bwd_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, g_attn, scale, L, v, s, Wv, Wo, dWo, dWv, dbv,
             ds);
  return check_launch("cross_vec_bwd");
}
