// Group C: the pieces of the situated spatial-attention transformer that are not GEMMs.
// (The GEMMs — QKV+cond, out-proj, FFN and every weight/data gradient — run on the tensor
// cores through msr3d_gemm_tc_f32, linear_tc.cu.)  All fp32, all tiny (M = B*L ~ 240 tokens):
// the point is to collapse the ~45 elementwise/reduction launches per layer of the reference
// (modules/layers/transformers.py:200-252,314-329; modules/utils.py:88-137) into a handful and
// to drop its per-layer host sync.
#include "common.cuh"
#include "rng.cuh"
#include "rowops.cuh"

#include <cooperative_groups.h>

namespace msr3d {


// ---------------------------------------------------------------------------
// Geometry: pairwise relative locations (calc_pairwise_locs, 'center', spatial_dim 5,
// modules/utils.py:88-137) and Fourier features of agent-frame centres
// (transform_to_agent_coor :60-82 + generate_fourier_features, ose3d_situation.py:31-59).
// grid (scene, row block): every CTA recomputes the scene's maximum pairwise distance (L*L/256
// evaluations per thread — cheaper than a second launch), then one thread per (row, key) pair of
// its row block writes the 5 pairwise features and one thread per (row, feature) the Fourier block.
// The first version used one CTA per scene and one thread per row (50 us for 4 scenes, a serial
// loop over the keys and 30 serial sincosf per thread).
//   loc   (B, L, 6)   centre xyz + size whl
//   anchor(B, 3), quat (B, 4) (x,y,z,w); use_agent_frame = 0 -> features of the raw centres
//   pw    (B, L, L, 5), ff (B*L, 64) [p'(3) | sin(pi p'_d f_j) d-major (30) | cos (30) | 0], sz (B*L, 3)
// ---------------------------------------------------------------------------
constexpr int kGeoThreads = 256;
constexpr int kGeoRows = 8;    // rows of the pairwise matrix per CTA

__global__ void __launch_bounds__(kGeoThreads)
geometry_kernel(const float *__restrict__ loc, const float *__restrict__ anchor,
                const float *__restrict__ quat, int L, int dist_norm, int use_agent_frame,
                float *__restrict__ pw, float *__restrict__ ff, float *__restrict__ sz) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  __shared__ float cx[kMaxL], cy[kMaxL], cz[kMaxL];
  __shared__ float s_red[kGeoThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float eps = 1e-10f;
  for (int l = tid; l < L; l += kGeoThreads) {
    const float *p = loc + ((size_t)b * L + l) * 6;
    cx[l] = p[0]; cy[l] = p[1]; cz[l] = p[2];
  }
  __syncthreads();
  float gmax = 1.f;
  if (dist_norm) {
    float mymax = 0.f;
    for (int i = tid; i < L * L; i += kGeoThreads) {
      const int l = i / L, t = i - l * L;
      const float dx = cx[l] - cx[t], dy = cy[l] - cy[t], dz = cz[l] - cz[t];
      mymax = fmaxf(mymax, sqrtf((dx * dx + dy * dy) + dz * dz + eps));
    }
    mymax = warp_max(mymax);
    if ((tid & 31) == 0) s_red[tid >> 5] = mymax;
    __syncthreads();
    gmax = 0.f;
#pragma unroll
    for (int w = 0; w < kGeoThreads / 32; ++w) gmax = fmaxf(gmax, s_red[w]);
  }
  const int l0 = blockIdx.y * kGeoRows;
  const int nrows = min(kGeoRows, L - l0);
  for (int i = tid; i < nrows * L; i += kGeoThreads) {
    const int l = l0 + i / L, t = i % L;
    const float dx = cx[l] - cx[t], dy = cy[l] - cy[t], dz = cz[l] - cz[t];
    const float d = sqrtf((dx * dx + dy * dy) + dz * dz + eps);
    const float d2 = sqrtf(dx * dx + dy * dy + eps);
    float *o = pw + (((size_t)b * L + l) * L + t) * 5;
    o[0] = dist_norm ? d / gmax : d;
    o[1] = dz / d;
    o[2] = d2 / d;
    o[3] = dy / d2;
    o[4] = dx / d2;
  }
  // Fourier block: one thread per (row, slot) with slot = 0..31: slots 0..29 -> (sin, cos) pair e,
  // slot 30 -> the three raw coordinates + zero pad, slot 31 -> the size triple
  for (int i = tid; i < nrows * 32; i += kGeoThreads) {
    const int l = l0 + (i >> 5), e = i & 31;
    // agent frame: p' = (p - anchor) R, R from the conjugate quaternion
    float px = cx[l], py = cy[l], pz = cz[l];
    if (use_agent_frame) {
      const float *a = anchor + (size_t)b * 3;
      const float *qq = quat + (size_t)b * 4;
      const float x = -qq[0], y = -qq[1], z = -qq[2], w = qq[3];
      const float xx = x * x, yy = y * y, zz = z * z;
      const float xy = x * y, xz = x * z, xw = x * w, yz = y * z, yw = y * w, zw = z * w;
      const float r00 = 1 - 2 * (yy + zz), r01 = 2 * (xy + zw), r02 = 2 * (xz - yw);
      const float r10 = 2 * (xy - zw), r11 = 1 - 2 * (xx + zz), r12 = 2 * (yz + xw);
      const float r20 = 2 * (xz + yw), r21 = 2 * (yz - xw), r22 = 1 - 2 * (xx + yy);
      const float ux = px - a[0], uy = py - a[1], uz = pz - a[2];
      px = (ux * r00 + uy * r10) + uz * r20;
      py = (ux * r01 + uy * r11) + uz * r21;
      pz = (ux * r02 + uy * r12) + uz * r22;
    }
    float *f = ff + ((size_t)b * L + l) * 64;
    if (e < 30) {
      const int d = e / 10, j = e % 10;
      const float kPi = 3.14159265358979323846f;
      // torch.linspace(1, 15, 10): start + j*step below the midpoint, end - (9-j)*step above
      const float step = 14.f / 9.f;
      const float fr = j < 5 ? 1.f + j * step : 15.f - (9 - j) * step;
      const float arg = kPi * ((d == 0 ? px : (d == 1 ? py : pz)) * fr);
      float sn, cs;
      sincosf(arg, &sn, &cs);
      f[3 + e] = sn;
      f[33 + e] = cs;
    } else if (e == 30) {
      f[0] = px; f[1] = py; f[2] = pz;
      f[63] = 0.f;
    } else {
      const float *p = loc + ((size_t)b * L + l) * 6;
      float *sp = sz + ((size_t)b * L + l) * 3;
      sp[0] = p[3]; sp[1] = p[4]; sp[2] = p[5];
    }
  }
}


// ---------------------------------------------------------------------------
// Attention block tail (transformers.py:250-251 then :324-325), fused:
//   v  = dropout_fc(o) + x          y1 = LN_a(v)
//   w  = x + dropout_1(y1)          y2 = LN_1(w)
// saves xhat/rstd of both norms for the backward pass.
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(kLnWarps * 32)
attn_tail_fwd_kernel(const float *__restrict__ o, const float *__restrict__ x, int M, LnParams lna,
                     LnParams ln1, Dropout drop, unsigned site_fc, unsigned site_1,
                     float *__restrict__ xhat_a, float *__restrict__ rstd_a,
                     float *__restrict__ xhat_1, float *__restrict__ rstd_1,
                     float *__restrict__ y2) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  const size_t base = (size_t)row * kD;
  Row8 xo = row_load(o + base, lane), xr = row_load(x + base, lane), v;   // rows in flight ...
  drop.resolve();                                                         // ... with the seed counter
  const Row8 mfc = drop_mask8(drop, base, lane, site_fc), m1 = drop_mask8(drop, base, lane, site_1);
#pragma unroll
  for (int i = 0; i < 8; ++i) v.v[i] = xo.v[i] * mfc.v[i] + xr.v[i];
  float ra, r1;
  Row8 y1 = ln_forward(v, lna.gamma, lna.beta, lane, ra);
  row_store(xhat_a + base, lane, v);
  Row8 w;
#pragma unroll
  for (int i = 0; i < 8; ++i) w.v[i] = xr.v[i] + y1.v[i] * m1.v[i];
  Row8 out = ln_forward(w, ln1.gamma, ln1.beta, lane, r1);
  row_store(xhat_1 + base, lane, w);
  row_store(y2 + base, lane, out);
  if (lane == 0) { rstd_a[row] = ra; rstd_1[row] = r1; }
}

// backward of the above.  In: dy2 (grad wrt y2, already including the FFN data gradient).
// Out: dx (grad wrt x through both residual branches), do_ (grad wrt the out-proj output),
// atomically accumulated: dgamma/dbeta of both norms and the out-proj bias gradient (colsum do_).
__global__ void __launch_bounds__(kLnWarps * 32)
attn_tail_bwd_kernel(const float *__restrict__ dy2, int M, LnParams lna, LnParams ln1, Dropout drop,
                     unsigned site_fc, unsigned site_1, const float *__restrict__ xhat_a,
                     const float *__restrict__ rstd_a, const float *__restrict__ xhat_1,
                     const float *__restrict__ rstd_1, float *__restrict__ dx,
                     float *__restrict__ do_, float *__restrict__ dgamma_a,
                     float *__restrict__ dbeta_a, float *__restrict__ dgamma_1,
                     float *__restrict__ dbeta_1, float *__restrict__ dbias_o) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  drop.resolve();
  __shared__ float s_acc[5 * kD];
  ColAcc acc{s_acc};
  acc.zero(5, threadIdx.x);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row < M) {
    const size_t base = (size_t)row * kD;
    const Row8 g2 = row_load(dy2 + base, lane);
    const Row8 h1 = row_load(xhat_1 + base, lane), ha = row_load(xhat_a + base, lane);
    const Row8 m1 = drop_mask8(drop, base, lane, site_1), mfc = drop_mask8(drop, base, lane, site_fc);
    Row8 t;
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = g2.v[i] * h1.v[i];
    acc.add(2, lane, t);   // dgamma_1
    acc.add(3, lane, g2);  // dbeta_1
    const Row8 dw = ln_backward(g2, h1, ln1.gamma, lane, rstd_1[row]);
    Row8 dy1;
#pragma unroll
    for (int i = 0; i < 8; ++i) dy1.v[i] = dw.v[i] * m1.v[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = dy1.v[i] * ha.v[i];
    acc.add(0, lane, t);    // dgamma_a
    acc.add(1, lane, dy1);  // dbeta_a
    const Row8 dv = ln_backward(dy1, ha, lna.gamma, lane, rstd_a[row]);
    Row8 dxr, dor;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dxr.v[i] = dw.v[i] + dv.v[i];
      dor.v[i] = dv.v[i] * mfc.v[i];
    }
    acc.add(4, lane, dor);  // bias gradient of the out projection
    row_store(dx + base, lane, dxr);
    row_store(do_ + base, lane, dor);
  }
  __syncthreads();
  acc.flush(0, dgamma_a, threadIdx.x);
  acc.flush(1, dbeta_a, threadIdx.x);
  acc.flush(2, dgamma_1, threadIdx.x);
  acc.flush(3, dbeta_1, threadIdx.x);
  acc.flush(4, dbias_o, threadIdx.x);
}

// ---------------------------------------------------------------------------
// FFN tail:  w = y2 + dropout_2(f);  y3 = LN_2(w);  x_next = y3 + add (add = qpos of the next
// layer, nullable).  Backward: dw = LNbwd(dy3); df = mask_2 * dw (+ colsum -> dbias_2).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kLnWarps * 32)
ffn_tail_fwd_kernel(const float *__restrict__ f, const float *__restrict__ y2, int M, LnParams ln2,
                    Dropout drop, unsigned site_2, const float *__restrict__ add,
                    float *__restrict__ xhat_2, float *__restrict__ rstd_2, float *__restrict__ y3,
                    float *__restrict__ x_next) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  const size_t base = (size_t)row * kD;
  Row8 xf = row_load(f + base, lane), xr = row_load(y2 + base, lane), w;
  drop.resolve();
  const Row8 m2 = drop_mask8(drop, base, lane, site_2);
#pragma unroll
  for (int i = 0; i < 8; ++i) w.v[i] = xr.v[i] + xf.v[i] * m2.v[i];
  float r2;
  Row8 out = ln_forward(w, ln2.gamma, ln2.beta, lane, r2);
  row_store(xhat_2 + base, lane, w);
  if (y3 != nullptr) row_store(y3 + base, lane, out);
  if (x_next != nullptr) {
    if (add != nullptr) {
      const Row8 a = row_load(add + base, lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) out.v[i] += a.v[i];
    }
    row_store(x_next + base, lane, out);
  }
  if (lane == 0) rstd_2[row] = r2;
}

__global__ void __launch_bounds__(kLnWarps * 32)
ffn_tail_bwd_kernel(const float *__restrict__ dy3, int M, LnParams ln2, Dropout drop, unsigned site_2,
                    const float *__restrict__ xhat_2, const float *__restrict__ rstd_2,
                    float *__restrict__ dy2, float *__restrict__ df, float *__restrict__ dgamma_2,
                    float *__restrict__ dbeta_2, float *__restrict__ dbias_2, float *__restrict__ df_big,
                    float *__restrict__ df_small) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  drop.resolve();
  __shared__ float s_acc[3 * kD];
  ColAcc acc{s_acc};
  acc.zero(3, threadIdx.x);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row < M) {
    const size_t base = (size_t)row * kD;
    const Row8 g = row_load(dy3 + base, lane), h = row_load(xhat_2 + base, lane);
    const Row8 m2 = drop_mask8(drop, base, lane, site_2);
    Row8 t;
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = g.v[i] * h.v[i];
    acc.add(0, lane, t);
    acc.add(1, lane, g);
    const Row8 dw = ln_backward(g, h, ln2.gamma, lane, rstd_2[row]);
    Row8 dfr;
#pragma unroll
    for (int i = 0; i < 8; ++i) dfr.v[i] = dw.v[i] * m2.v[i];
    acc.add(2, lane, dfr);
    row_store(dy2 + base, lane, dw);
    row_store(df + base, lane, dfr);
    if (df_big != nullptr) {     // operand images of df for the pre-split FFN pair (big = tf32-rounded, small = rest)
      Row8 bg, sm;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t ub = (__float_as_uint(dfr.v[i]) + 0x1000u) & 0xffffe000u;
        bg.v[i] = __uint_as_float(ub);
        sm.v[i] = dfr.v[i] - bg.v[i];
      }
      row_store(df_big + base, lane, bg);
      row_store(df_small + base, lane, sm);
    }
  }
  __syncthreads();
  acc.flush(0, dgamma_2, threadIdx.x);
  acc.flush(1, dbeta_2, threadIdx.x);
  acc.flush(2, dbias_2, threadIdx.x);
}

// ---------------------------------------------------------------------------
// Query position (ose3d_situation.py:399-404):
//   qpos = LN_l(h1) + LN_s(Ws sz + bs),  h1 = Wl ff + bl comes from the tensor-core GEMM.
//   qpos0 = qpos + extra   (extra = object_orientation_feat + type_embedding[0], layer-0 input)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kLnWarps * 32)
qpos_fwd_kernel(const float *__restrict__ h1, const float *__restrict__ sz, int M,
                const float *__restrict__ Ws, const float *__restrict__ bs, LnParams lnl, LnParams lns,
                const float *__restrict__ extra_a, const float *__restrict__ extra_b,
                float *__restrict__ xhat_l, float *__restrict__ rstd_l, float *__restrict__ xhat_s,
                float *__restrict__ rstd_s, float *__restrict__ qpos, float *__restrict__ qpos0,
                float *__restrict__ x0_accum) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  const size_t base = (size_t)row * kD;
  Row8 a = row_load(h1 + base, lane), b;
  const float s0 = sz[row * 3 + 0], s1 = sz[row * 3 + 1], s2 = sz[row * 3 + 2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = col8(lane, i);
    b.v[i] = (Ws[c * 3 + 0] * s0 + Ws[c * 3 + 1] * s1) + Ws[c * 3 + 2] * s2 + bs[c];
  }
  float rl, rs;
  Row8 ya = ln_forward(a, lnl.gamma, lnl.beta, lane, rl);
  Row8 yb = ln_forward(b, lns.gamma, lns.beta, lane, rs);
  row_store(xhat_l + base, lane, a);
  row_store(xhat_s + base, lane, b);
  Row8 q, q0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = col8(lane, i);
    q.v[i] = ya.v[i] + yb.v[i];
    q0.v[i] = q.v[i] + (extra_a ? extra_a[c] : 0.f) + (extra_b ? extra_b[c] : 0.f);
  }
  row_store(qpos + base, lane, q);
  row_store(qpos0 + base, lane, q0);
  if (x0_accum != nullptr) {
    // layer-0 input = projected object embedding (split-K partial sums reduced with red.add into a zeroed buffer by the
    // projection GEMM, which runs beside this kernel) + qpos0: one more term of the same reduction
    float *d = x0_accum + base + lane * 8;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(q0.v[0]), "f"(q0.v[1]), "f"(q0.v[2]), "f"(q0.v[3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(q0.v[4]), "f"(q0.v[5]), "f"(q0.v[6]), "f"(q0.v[7]) : "memory");
  }
  if (lane == 0) { rstd_l[row] = rl; rstd_s[row] = rs; }
}

// dqpos -> dh1 (for the Wl GEMMs), and in place: dgamma/dbeta of both norms, dWs (256x3), dbs,
// and the bias gradient of the loc linear (colsum dh1).
__global__ void __launch_bounds__(kLnWarps * 32)
qpos_bwd_kernel(const float *__restrict__ dq, const float *__restrict__ dq1,
                const float *__restrict__ dq2, const float *__restrict__ sz, int M, LnParams lnl,
                LnParams lns, const float *__restrict__ xhat_l, const float *__restrict__ rstd_l,
                const float *__restrict__ xhat_s, const float *__restrict__ rstd_s,
                float *__restrict__ dh1, float *__restrict__ dgamma_l, float *__restrict__ dbeta_l,
                float *__restrict__ dgamma_s, float *__restrict__ dbeta_s, float *__restrict__ dWs,
                float *__restrict__ dbs, float *__restrict__ dbl) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  __shared__ float s_acc[8 * kD];
  ColAcc acc{s_acc};
  acc.zero(8, threadIdx.x);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row < M) {
    const size_t base = (size_t)row * kD;
    Row8 g = row_load(dq + base, lane);   // qpos is added at every layer: sum the per-layer grads
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
    const Row8 hl = row_load(xhat_l + base, lane), hs = row_load(xhat_s + base, lane);
    Row8 t;
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = g.v[i] * hl.v[i];
    acc.add(0, lane, t);
    acc.add(1, lane, g);   // dbeta_l == dbeta_s == colsum(dq)
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = g.v[i] * hs.v[i];
    acc.add(2, lane, t);
    const Row8 d1 = ln_backward(g, hl, lnl.gamma, lane, rstd_l[row]);
    const Row8 d2 = ln_backward(g, hs, lns.gamma, lane, rstd_s[row]);
    row_store(dh1 + base, lane, d1);
    acc.add(3, lane, d1);  // dbl
    acc.add(4, lane, d2);  // dbs
    const float s0 = sz[row * 3 + 0], s1 = sz[row * 3 + 1], s2 = sz[row * 3 + 2];
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = d2.v[i] * s0;
    acc.add(5, lane, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = d2.v[i] * s1;
    acc.add(6, lane, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = d2.v[i] * s2;
    acc.add(7, lane, t);
  }
  __syncthreads();
  acc.flush(0, dgamma_l, threadIdx.x);
  acc.flush(1, dbeta_l, threadIdx.x);
  acc.flush(2, dgamma_s, threadIdx.x);
  acc.flush(1, dbeta_s, threadIdx.x);
  acc.flush(3, dbl, threadIdx.x);
  acc.flush(4, dbs, threadIdx.x);
  if (dWs != nullptr)
    for (int c = threadIdx.x; c < kD; c += kLnWarps * 32) {
      atomicAdd(dWs + c * 3 + 0, acc.get(5, c));
      atomicAdd(dWs + c * 3 + 1, acc.get(6, c));
      atomicAdd(dWs + c * 3 + 2, acc.get(7, c));
    }
}

// ---------------------------------------------------------------------------
// dst[n] += sum_m src[m, n]   (bias gradients, d(orientation_feat), d(type_embedding[0]), ...)
// up to three destinations receive the same column sums.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const float *__restrict__ src, long long ld, int M, int N, float *__restrict__ d0,
              float *__restrict__ d1, float *__restrict__ d2, int rows_per_cta) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m0 = blockIdx.y * rows_per_cta;
  const int m1 = min(M, m0 + rows_per_cta);
  float s = 0.f;
  for (int m = m0; m < m1; ++m) s += src[(long long)m * ld + n];
  if (d0) atomicAdd(d0 + n, s);
  if (d1) atomicAdd(d1 + n, s);
  if (d2) atomicAdd(d2 + n, s);
}

// ---------------------------------------------------------------------------
// Attention core, `cond` fusion (transformers.py:205-248).  qkvc rows:
// [q(256) | k(256) | v(256) | cond(8 heads x 6)].
//   S = q k^T / sqrt(dh);  loc = sigmoid(sw[1:] . pw[l,t,:] + sw[0]);
//   masked key: S = -inf, loc = 0;  P = softmax(log(max(loc, 1e-6)) + S);  ctx = P v
// One CTA per (head, scene, row block); ONE WARP PER QUERY ROW with the keys spread over the
// lanes (key t = lane + 32 j), so every global access is coalesced and nothing is serial over
// the keys: the first version (one thread per row) was a chain of ~60 dependent L2 round trips.
// K is staged transposed in shared memory (lane = key -> conflict-free), V row-major (lane = dim).
// Saves P and loc (B,H,L,L) for the backward pass.
// ---------------------------------------------------------------------------
constexpr int kDh = 32;
constexpr size_t kOnePerSmBytes = 116 * 1024;   // more than half of the 227 KB an SM can give its CTAs
constexpr int kAttnMaxWarps = 16;    // launch bound; the warp count itself is a launch-time choice
constexpr int kKeysPerLane = kMaxL / 32;  // 4

// FUSION selects how the pairwise-location term z = sw[0] + sw[1:] . pw[l,t,:] enters (transformers.py:207-244):
//   kFuseCond  logits = log(clamp(sigmoid(z), 1e-6)) + S      (sw = lang_cond_fc(x), per token and head)
//   kFuseMul   logits = log(clamp(relu(z), 1e-6)) + S         (sw = pairwise_loc_fc rows, the same for every token)
//   kFuseBias  logits = z + S                                 (`bias`; `ctx` too, with sw derived from q: spatial_weights_*)
//   kFuseAdd   P = (softmax(S) + softmax(z)) / 2              (both softmaxes saved: P buffer <- softmax(S), LOC <- softmax(z))
// masked keys: S = -inf; loc = 0 (cond, mul) / z = -inf (bias, ctx, add).
enum { kFuseCond = 0, kFuseMul = 1, kFuseBias = 2, kFuseAdd = 3 };

template <int FUSION>
__global__ void __launch_bounds__(kAttnMaxWarps * 32)
attn_fwd_kernel(const float *__restrict__ qkvc, int ldq, const float *__restrict__ pw,
                const unsigned char *__restrict__ key_pad, int L, int H, float *__restrict__ P,
                float *__restrict__ LOC, float *__restrict__ ctx, int ldc) {
  if (threadIdx.x == 0) trace_mark(12);
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  if (threadIdx.x == 0) trace_mark(13);
  extern __shared__ float sm[];
  const int LT = ((L + 31) / 32) * 32;
  float *kT = sm;                 // [32][LT]
  float *vs = sm + kDh * LT;      // [L][32]
  const int h = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = H * kDh;
  for (int i = threadIdx.x; i < L * kDh; i += blockDim.x) {
    const int t = i / kDh, d = i % kDh;
    const float *r = qkvc + ((size_t)b * L + t) * ldq;
    kT[d * LT + t] = r[D + h * kDh + d];
    vs[i] = r[2 * D + h * kDh + d];
  }
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(14);
  const float scale = rsqrtf((float)kDh);
  const int nj = LT / 32;
  const int nwarps = blockDim.x >> 5;
  for (int l = blockIdx.z * nwarps + warp; l < L; l += gridDim.z * nwarps) {
    const float *r = qkvc + ((size_t)b * L + l) * ldq;
    const float qd = r[h * kDh + lane];
    const float swl = lane < 6 ? r[3 * D + h * 6 + lane] : 0.f;
    float sw[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) sw[d] = __shfl_sync(0xffffffffu, swl, d);
    float s[kKeysPerLane], loc[kKeysPerLane];
#pragma unroll
    for (int j = 0; j < kKeysPerLane; ++j) s[j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < kDh; ++d) {
      const float q = __shfl_sync(0xffffffffu, qd, d);
#pragma unroll
      for (int j = 0; j < kKeysPerLane; ++j)
        if (j < nj) s[j] += q * kT[d * LT + lane + 32 * j];
    }
    const float *pwr = pw + ((size_t)b * L + l) * L * 5;
    float mx = -INFINITY, mz = -INFINITY;
#pragma unroll
    for (int j = 0; j < kKeysPerLane; ++j) {
      const int t = lane + 32 * j;
      if (j < nj && t < L) {
        const float *g = pwr + t * 5;
        float z = sw[0];
#pragma unroll
        for (int d = 0; d < 5; ++d) z += sw[1 + d] * g[d];
        const bool masked = key_pad != nullptr && key_pad[(size_t)b * L + t];
        float sc = masked ? -INFINITY : s[j] * scale;
        if (FUSION == kFuseCond || FUSION == kFuseMul) {
          float lc = FUSION == kFuseCond ? 1.f / (1.f + __expf(-z)) : fmaxf(z, 0.f);
          if (masked) lc = 0.f;
          s[j] = __logf(fmaxf(lc, 1e-6f)) + sc;
          loc[j] = lc;
        } else if (FUSION == kFuseBias) {
          s[j] = z + sc;
          loc[j] = 0.f;
        } else {
          s[j] = sc;
          loc[j] = masked ? -INFINITY : z;
          mz = fmaxf(mz, loc[j]);
        }
        mx = fmaxf(mx, s[j]);
      } else {
        s[j] = -INFINITY;
        loc[j] = FUSION == kFuseAdd ? -INFINITY : 0.f;
      }
    }
    mx = warp_max(mx);
    if (threadIdx.x == 0) trace_mark(15);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kKeysPerLane; ++j) {
      s[j] = __expf(s[j] - mx);   // exp(-inf) = 0 for masked / out-of-range keys
      sum += s[j];
    }
    const float inv = 1.f / warp_sum(sum);
    float invz = 0.f;
    if (FUSION == kFuseAdd) {       // second softmax, over the location logits alone
      mz = warp_max(mz);
      float sumz = 0.f;
#pragma unroll
      for (int j = 0; j < kKeysPerLane; ++j) {
        loc[j] = __expf(loc[j] - mz);
        sumz += loc[j];
      }
      invz = 1.f / warp_sum(sumz);
    }
    float *Pr = P + (((size_t)b * H + h) * L + l) * L;
    float *Lr = LOC + (((size_t)b * H + h) * L + l) * L;
#pragma unroll
    for (int j = 0; j < kKeysPerLane; ++j) {
      s[j] *= inv;
      if (FUSION == kFuseAdd) loc[j] *= invz;
      const int t = lane + 32 * j;
      if (j < nj && t < L) {
        Pr[t] = s[j];
        if (FUSION != kFuseBias) Lr[t] = loc[j];
      }
      if (FUSION == kFuseAdd) s[j] = 0.5f * (s[j] + loc[j]);   // the fused attention that weights v
    }
    float acc = 0.f;  // lane = output dim
#pragma unroll
    for (int j = 0; j < kKeysPerLane; ++j) {
      if (j < nj) {
        const int tmax = min(32, L - 32 * j);
        for (int tt = 0; tt < tmax; ++tt)
          acc += __shfl_sync(0xffffffffu, s[j], tt) * vs[(32 * j + tt) * kDh + lane];
      }
    }
    ctx[((size_t)b * L + l) * ldc + h * kDh + lane] = acc;
    if (threadIdx.x == 0) trace_mark(8);
  }
}

// Backward of the attention core.  In: dctx (M,256), P, LOC, qkvc, pw.  Out: dqkvc (M,816),
// every column written exactly once (the cond gradient goes to columns 768..815).
// One thread-block CLUSTER per (head, scene); the CTAs of the cluster split the rows.
//   Phase 1: warp per query row (keys over lanes) -> dlogits and P rows into this CTA's shared
//            memory, dq and d(cond) to global.
//   Phase 2: warp per key row (dims over lanes) -> dk, dv.  These contract over EVERY query row, so
//            the dlogits / P rows of the other CTAs are read through distributed shared memory
//            (no atomics, no zero-fill, deterministic).
// With 32 (head, scene) pairs a one-CTA-per-pair kernel filled 32 SMs and walked 8 rows per warp in
// each phase (40 us); a cluster of 4 x 16 warps does one row per warp per phase.
template <int FUSION>
__global__ void __launch_bounds__(kAttnMaxWarps * 32)
attn_bwd_kernel(const float *__restrict__ dctx, int ldc, const float *__restrict__ qkvc, int ldq,
                const float *__restrict__ pw, const float *__restrict__ P,
                const float *__restrict__ LOC, int L, int H, float *__restrict__ dqkvc) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int nz = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  extern __shared__ float sm[];
  const int LT = ((L + 31) / 32) * 32;
  const int LP = LT + 1;
  const int rpc = (L + nz - 1) / nz;   // rows per CTA (query rows in phase 1, key rows in phase 2)
  float *qs = sm;                 // [L][32]
  float *ks = qs + L * kDh;       // [L][32]
  float *gs = ks + L * kDh;       // dctx [L][32]
  float *vT = gs + L * kDh;       // [32][LT]
  float *dS = vT + kDh * LT;      // [rpc][LP]   rows rank*rpc ..
  float *Ps = dS + rpc * LP;      // [rpc][LP]
  const int h = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int D = H * kDh;
  for (int i = threadIdx.x; i < L * kDh; i += blockDim.x) {
    const int t = i / kDh, d = i % kDh;
    const float *r = qkvc + ((size_t)b * L + t) * ldq;
    qs[i] = r[h * kDh + d];
    ks[i] = r[D + h * kDh + d];
    vT[d * LT + t] = r[2 * D + h * kDh + d];
    gs[i] = dctx[((size_t)b * L + t) * ldc + h * kDh + d];
  }
  __syncthreads();
  const float scale = rsqrtf((float)kDh);
  const int nj = LT / 32;
  const int row0 = rank * rpc, row1 = min(L, row0 + rpc);
  for (int l = row0 + warp; l < row1; l += nwarps) {
    const int ll = l - row0;
    const float *r = qkvc + ((size_t)b * L + l) * ldq;
    const float swl = lane < 6 ? r[3 * D + h * 6 + lane] : 0.f;
    float sw[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) sw[d] = __shfl_sync(0xffffffffu, swl, d);
    const float gd = gs[l * kDh + lane];
    float dp[kKeysPerLane], p[kKeysPerLane];
#pragma unroll
    for (int j = 0; j < kKeysPerLane; ++j) dp[j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < kDh; ++d) {
      const float g = __shfl_sync(0xffffffffu, gd, d);
#pragma unroll
      for (int j = 0; j < kKeysPerLane; ++j)
        if (j < nj) dp[j] += g * vT[d * LT + lane + 32 * j];
    }
    const float *Pr = P + (((size_t)b * H + h) * L + l) * L;
    const float *Lr = LOC + (((size_t)b * H + h) * L + l) * L;
    const float *pwr = pw + ((size_t)b * L + l) * L * 5;
    float rowdot = 0.f, rowdot_z = 0.f;
    float pz[kKeysPerLane];   // kFuseAdd: softmax of the location logits (LOC buffer)
#pragma unroll
    for (int j = 0; j < kKeysPerLane; ++j) {
      const int t = lane + 32 * j;
      p[j] = (j < nj && t < L) ? Pr[t] : 0.f;
      rowdot += dp[j] * p[j];
      pz[j] = 0.f;
      if (FUSION == kFuseAdd) {
        pz[j] = (j < nj && t < L) ? Lr[t] : 0.f;
        rowdot_z += dp[j] * pz[j];
      }
    }
    rowdot = warp_sum(rowdot);
    if (FUSION == kFuseAdd) rowdot_z = warp_sum(rowdot_z);
    float dsw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < kKeysPerLane; ++j) {
      const int t = lane + 32 * j;
      // d logits of the content term (kFuseAdd: of softmax(S), which carries half of the fused attention)
      const float dl = (FUSION == kFuseAdd ? 0.5f : 1.f) * p[j] * (dp[j] - rowdot);
      const float dz_add = 0.5f * pz[j] * (dp[j] - rowdot_z);
      dp[j] = dl;
      if (j < nj && t < L) {
        dS[ll * LP + t] = dl;
        Ps[ll * LP + t] = FUSION == kFuseAdd ? 0.5f * (p[j] + pz[j]) : p[j];
        const float *g = pwr + t * 5;
        float gz;
        if (FUSION == kFuseCond) {
          // log(clamp(sigmoid(z), 1e-6)): d/dz = (loc >= 1e-6) * (1 - loc); 1 - sigmoid(z) is evaluated as
          // sigmoid(-z) from the recomputed z (no cancellation near loc ~ 1); masked keys have loc = 0
          const float lc = Lr[t];
          float z = sw[0];
#pragma unroll
          for (int d = 0; d < 5; ++d) z += sw[1 + d] * g[d];
          gz = lc >= 1e-6f ? dl / (1.f + __expf(z)) : 0.f;
        } else if (FUSION == kFuseMul) {
          // log(clamp(relu(z), 1e-6)): d/dz = 1 / loc where loc = relu(z) >= 1e-6 (which implies z > 0)
          const float lc = Lr[t];
          gz = lc >= 1e-6f ? dl / lc : 0.f;
        } else if (FUSION == kFuseBias) {
          gz = dl;                  // logits = z + S; masked keys have P = 0, hence dl = 0
        } else {
          gz = dz_add;
        }
        dsw[0] += gz;
#pragma unroll
        for (int d = 0; d < 5; ++d) dsw[1 + d] += gz * g[d];
      }
    }
    float *o = dqkvc + ((size_t)b * L + l) * ldq;
#pragma unroll
    for (int d = 0; d < 6; ++d) {
      const float tsum = warp_sum(dsw[d]);
      if (lane == d) o[3 * D + h * 6 + d] = tsum;
    }
    float dq = 0.f;  // lane = dim
#pragma unroll
    for (int j = 0; j < kKeysPerLane; ++j) {
      if (j < nj) {
        const int tmax = min(32, L - 32 * j);
        for (int tt = 0; tt < tmax; ++tt)
          dq += __shfl_sync(0xffffffffu, dp[j], tt) * ks[(32 * j + tt) * kDh + lane];
      }
    }
    o[h * kDh + lane] = dq * scale;
  }
  cluster.sync();   // every CTA's dS / Ps rows are complete and visible cluster-wide
  for (int t = row0 + warp; t < row1; t += nwarps) {  // lane = dim
    float dk = 0.f, dv = 0.f;
    for (int r = 0; r < nz; ++r) {
      const float *dS_r = cluster.map_shared_rank(dS, r);
      const float *Ps_r = cluster.map_shared_rank(Ps, r);
      const int i0 = r * rpc, cnt = min(L, i0 + rpc) - i0;
      for (int il = 0; il < cnt; ++il) {
        dk += dS_r[il * LP + t] * qs[(i0 + il) * kDh + lane];
        dv += Ps_r[il * LP + t] * gs[(i0 + il) * kDh + lane];
      }
    }
    float *o = dqkvc + ((size_t)b * L + t) * ldq;
    o[D + h * kDh + lane] = dk * scale;
    o[2 * D + h * kDh + lane] = dv;
  }
  cluster.sync();   // nobody exits while a peer may still be reading its shared memory
}

// ---------------------------------------------------------------------------
// Spatial weights of the fusions whose location term does not come out of the QKV+cond GEMM
// (transformers.py:207-221).  They are written into the `cond` columns of the qkvc rows,
// [bias, w0..w4] per (token, head), so that the attention kernels above read every fusion the same way:
//   mul / bias / add   sw[m, h, :] = [b[hs], W[hs, 0..4]],  pairwise_loc_fc = Linear(5, spatial_n_head), hs = h or 0
//   ctx                loc[l,t] = q_h(l) . (Wc_h pw[l,t] + bc_h) / sqrt(dh)   (pairwise_loc_fc = Linear(5, d_model))
//                      = sw . [1, pw]  with  sw[m, h, 0] = q_h(m) . bc_h / sqrt(dh),  sw[m, h, 1+d] = q_h(m) . Wc_h[:, d] / sqrt(dh)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spatial_weights_fwd_kernel(const float *__restrict__ W, const float *__restrict__ bvec, int ctx_mode, int multihead,
                           float *__restrict__ qkvc, int ldq, int M, int H) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (token, head)
  if (i >= M * H) return;
  const int m = i / H, h = i % H;
  float *o = qkvc + (size_t)m * ldq + 3 * H * kDh + h * 6;
  if (!ctx_mode) {
    const int hs = multihead ? h : 0;
    o[0] = bvec[hs];
#pragma unroll
    for (int d = 0; d < 5; ++d) o[1 + d] = W[hs * 5 + d];
    return;
  }
  const float *q = qkvc + (size_t)m * ldq + h * kDh;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < kDh; ++k) {
    const float qk = q[k];
    const float *w = W + (size_t)(h * kDh + k) * 5;
    acc[0] += qk * bvec[h * kDh + k];
#pragma unroll
    for (int d = 0; d < 5; ++d) acc[1 + d] += qk * w[d];
  }
  const float scale = rsqrtf((float)kDh);
#pragma unroll
  for (int d = 0; d < 6; ++d) o[d] = acc[d] * scale;
}

// Backward of the above from d(sw) = dqkvc[:, 3 D + 6 h + j] (written by attn_bwd_kernel).  Deterministic: fixed
// summation order, no atomics; parameter gradients ACCUMULATE into dW / db (flat gradient buffer).
//   mul / bias / add: one CTA; column sums over the tokens, heads folded when the layer is not multi-head.
//   ctx: CTA m < M adds the q gradient of token m (thread = (head, k)); CTA M reduces dWc, dbc over the tokens.
__global__ void __launch_bounds__(256)
spatial_weights_bwd_kernel(const float *__restrict__ W, const float *__restrict__ bvec, int ctx_mode, int multihead,
                           const float *__restrict__ qkvc, float *__restrict__ dqkvc, int ldq, int M, int H,
                           float *__restrict__ dW, float *__restrict__ db) {
  pdl_sync();
  const int D = H * kDh;
  if (!ctx_mode) {
    __shared__ float part[4][64];
    const int c = threadIdx.x & 63, r = threadIdx.x >> 6;     // column of d(sw), row group
    float s = 0.f;
    if (c < H * 6)
      for (int m = r; m < M; m += 4) s += dqkvc[(size_t)m * ldq + 3 * D + c];
    part[r][c] = s;
    __syncthreads();
    const int nsh = multihead ? H : 1;
    if (threadIdx.x < nsh * 6) {
      const int hs = threadIdx.x / 6, j = threadIdx.x % 6;
      float t = 0.f;
      for (int h = multihead ? hs : 0; h < (multihead ? hs + 1 : H); ++h)
        t += (part[0][h * 6 + j] + part[1][h * 6 + j]) + (part[2][h * 6 + j] + part[3][h * 6 + j]);
      if (j == 0) db[hs] += t;
      else dW[hs * 5 + j - 1] += t;
    }
    return;
  }
  const float scale = rsqrtf((float)kDh);
  const int hk = threadIdx.x;            // blockDim = D = 256
  if (hk >= D) return;
  const int h = hk / kDh;
  if ((int)blockIdx.x < M) {
    const int m = blockIdx.x;
    const float *ds = dqkvc + (size_t)m * ldq + 3 * D + h * 6;
    float g = ds[0] * bvec[hk];
#pragma unroll
    for (int d = 0; d < 5; ++d) g += ds[1 + d] * W[(size_t)hk * 5 + d];
    dqkvc[(size_t)m * ldq + hk] += g * scale;
    return;
  }
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int m = 0; m < M; ++m) {
    const float q = qkvc[(size_t)m * ldq + hk];
    const float *ds = dqkvc + (size_t)m * ldq + 3 * D + h * 6;
#pragma unroll
    for (int d = 0; d < 6; ++d) acc[d] += q * ds[d];
  }
  db[hk] += acc[0] * scale;
#pragma unroll
  for (int d = 0; d < 5; ++d) dW[(size_t)hk * 5 + d] += acc[1 + d] * scale;
}

__global__ void dropout_mask_kernel(Dropout drop, unsigned site, long long n, float *out) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  drop.resolve();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = drop.mask((uint64_t)i, site);
}

}  // namespace msr3d

namespace msr3d {
int set_trace_spatial(long long *buf) {
  MSR3D_CUDA(cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)));
  return MSR3D_OK;
}
}  // namespace msr3d

using namespace msr3d;

static inline Dropout mk(unsigned long long seed, float p, const unsigned long long *sp) {
  return make_dropout(seed, p, sp);
}
static inline int ln_grid(int M) { return ceil_div(M, kLnWarps); }

MSR3D_API int msr3d_spatial_geometry_f32(const float *loc, const float *anchor, const float *quat,
                                         int B, int L, int dist_norm, int use_agent_frame, float *pw,
                                         float *ff, float *sz, msr3d_stream_t stream) {
  MSR3D_REQUIRE(loc && pw && ff && sz, "geometry: null pointer");
  MSR3D_REQUIRE(!use_agent_frame || (anchor && quat), "geometry: agent frame needs anchor + quaternion");
  MSR3D_REQUIRE(B >= 0 && L >= 1 && L <= kMaxL, "geometry: L=%d must be in [1,%d]", L, kMaxL);
  if (B == 0) return MSR3D_OK;
  launch_pdl(geometry_kernel, dim3(B, ceil_div(L, kGeoRows)), dim3(kGeoThreads), 0, (cudaStream_t)stream, loc, anchor, quat, L, dist_norm,
                                                        use_agent_frame, pw, ff, sz);
  return check_launch("geometry");
}

MSR3D_API int msr3d_attn_tail_fwd_f32(const float *o, const float *x, int M, const float *gamma_a,
                                      const float *beta_a, const float *gamma_1, const float *beta_1,
                                      unsigned long long seed, const unsigned long long *seed_ptr,
                                      float p, unsigned site_fc, unsigned site_1, float *xhat_a, float *rstd_a, float *xhat_1,
                                      float *rstd_1, float *y2, msr3d_stream_t stream) {
  MSR3D_REQUIRE(o && x && gamma_a && beta_a && gamma_1 && beta_1 && xhat_a && rstd_a && xhat_1 && rstd_1 && y2,
                "attn_tail_fwd: null pointer");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(attn_tail_fwd_kernel, dim3(ln_grid(M)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, 
      o, x, M, LnParams{gamma_a, beta_a}, LnParams{gamma_1, beta_1}, mk(seed, p, seed_ptr), site_fc, site_1,
      xhat_a, rstd_a, xhat_1, rstd_1, y2);
  return check_launch("attn_tail_fwd");
}

MSR3D_API int msr3d_attn_tail_bwd_f32(const float *dy2, int M, const float *gamma_a,
                                      const float *gamma_1, unsigned long long seed,
                                      const unsigned long long *seed_ptr, float p,
                                      unsigned site_fc, unsigned site_1, const float *xhat_a,
                                      const float *rstd_a, const float *xhat_1, const float *rstd_1,
                                      float *dx, float *do_, float *dgamma_a, float *dbeta_a,
                                      float *dgamma_1, float *dbeta_1, float *dbias_o,
                                      msr3d_stream_t stream) {
  MSR3D_REQUIRE(dy2 && gamma_a && gamma_1 && xhat_a && rstd_a && xhat_1 && rstd_1 && dx && do_,
                "attn_tail_bwd: null pointer");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(attn_tail_bwd_kernel, dim3(ln_grid(M)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, 
      dy2, M, LnParams{gamma_a, nullptr}, LnParams{gamma_1, nullptr}, mk(seed, p, seed_ptr), site_fc, site_1,
      xhat_a, rstd_a, xhat_1, rstd_1, dx, do_, dgamma_a, dbeta_a, dgamma_1, dbeta_1, dbias_o);
  return check_launch("attn_tail_bwd");
}

MSR3D_API int msr3d_ffn_tail_fwd_f32(const float *f, const float *y2, int M, const float *gamma_2,
                                     const float *beta_2, unsigned long long seed,
                                     const unsigned long long *seed_ptr, float p,
                                     unsigned site_2, const float *add, float *xhat_2, float *rstd_2,
                                     float *y3, float *x_next, msr3d_stream_t stream) {
  MSR3D_REQUIRE(f && y2 && gamma_2 && beta_2 && xhat_2 && rstd_2, "ffn_tail_fwd: null pointer");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(ffn_tail_fwd_kernel, dim3(ln_grid(M)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, 
      f, y2, M, LnParams{gamma_2, beta_2}, mk(seed, p, seed_ptr), site_2, add, xhat_2, rstd_2, y3, x_next);
  return check_launch("ffn_tail_fwd");
}

MSR3D_API int msr3d_ffn_tail_bwd_ps_f32(const float *dy3, int M, const float *gamma_2,
                                        unsigned long long seed, const unsigned long long *seed_ptr, float p,
                                        unsigned site_2, const float *xhat_2, const float *rstd_2, float *dy2, float *df,
                                        float *dgamma_2, float *dbeta_2, float *dbias_2, float *df_big, float *df_small,
                                        msr3d_stream_t stream) {
  MSR3D_REQUIRE(dy3 && gamma_2 && xhat_2 && rstd_2 && dy2 && df, "ffn_tail_bwd: null pointer");
  MSR3D_REQUIRE((df_big == nullptr) == (df_small == nullptr), "ffn_tail_bwd: both images or none");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(ffn_tail_bwd_kernel, dim3(ln_grid(M)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, 
      dy3, M, LnParams{gamma_2, nullptr}, mk(seed, p, seed_ptr), site_2, xhat_2, rstd_2, dy2, df, dgamma_2,
      dbeta_2, dbias_2, df_big, df_small);
  return check_launch("ffn_tail_bwd");
}

MSR3D_API int msr3d_ffn_tail_bwd_f32(const float *dy3, int M, const float *gamma_2,
                                     unsigned long long seed, const unsigned long long *seed_ptr, float p,
                                     unsigned site_2, const float *xhat_2, const float *rstd_2, float *dy2, float *df,
                                     float *dgamma_2, float *dbeta_2, float *dbias_2,
                                     msr3d_stream_t stream) {
  return msr3d_ffn_tail_bwd_ps_f32(dy3, M, gamma_2, seed, seed_ptr, p, site_2, xhat_2, rstd_2, dy2, df, dgamma_2, dbeta_2, dbias_2,
                                   nullptr, nullptr, stream);
}

static int qpos_fwd_impl(const float *h1, const float *sz, int M, const float *Ws,
                                 const float *bs, const float *gamma_l, const float *beta_l,
                                 const float *gamma_s, const float *beta_s, const float *extra_a,
                                 const float *extra_b, float *xhat_l, float *rstd_l, float *xhat_s,
                                 float *rstd_s, float *qpos, float *qpos0, float *x0_accum, msr3d_stream_t stream) {
  MSR3D_REQUIRE(h1 && sz && Ws && bs && gamma_l && beta_l && gamma_s && beta_s && xhat_l && rstd_l &&
                    xhat_s && rstd_s && qpos && qpos0, "qpos_fwd: null pointer");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(qpos_fwd_kernel, dim3(ln_grid(M)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, 
      h1, sz, M, Ws, bs, LnParams{gamma_l, beta_l}, LnParams{gamma_s, beta_s}, extra_a, extra_b,
      xhat_l, rstd_l, xhat_s, rstd_s, qpos, qpos0, x0_accum);
  return check_launch("qpos_fwd");
}

MSR3D_API int msr3d_qpos_fwd_f32(const float *h1, const float *sz, int M, const float *Ws, const float *bs,
                                 const float *gamma_l, const float *beta_l, const float *gamma_s, const float *beta_s,
                                 const float *extra_a, const float *extra_b, float *xhat_l, float *rstd_l, float *xhat_s,
                                 float *rstd_s, float *qpos, float *qpos0, msr3d_stream_t stream) {
  return qpos_fwd_impl(h1, sz, M, Ws, bs, gamma_l, beta_l, gamma_s, beta_s, extra_a, extra_b, xhat_l, rstd_l, xhat_s, rstd_s,
                       qpos, qpos0, nullptr, stream);
}

/* Same; additionally qpos0 is ADDED (red.add) into x0_accum (M, 256): the layer-0 input is assembled by this kernel and
 * the object-projection GEMM running side by side, both reducing into one zero-filled buffer. */
MSR3D_API int msr3d_qpos_fwd_acc_f32(const float *h1, const float *sz, int M, const float *Ws, const float *bs,
                                     const float *gamma_l, const float *beta_l, const float *gamma_s, const float *beta_s,
                                     const float *extra_a, const float *extra_b, float *xhat_l, float *rstd_l,
                                     float *xhat_s, float *rstd_s, float *qpos, float *qpos0, float *x0_accum,
                                     msr3d_stream_t stream) {
  MSR3D_REQUIRE(x0_accum != nullptr && (reinterpret_cast<uintptr_t>(x0_accum) & 15) == 0, "qpos_fwd_acc: x0_accum must be 16 B aligned");
  return qpos_fwd_impl(h1, sz, M, Ws, bs, gamma_l, beta_l, gamma_s, beta_s, extra_a, extra_b, xhat_l, rstd_l, xhat_s, rstd_s,
                       qpos, qpos0, x0_accum, stream);
}

MSR3D_API int msr3d_qpos_bwd_f32(const float *dq, const float *dq1, const float *dq2,
                                 const float *sz, int M, const float *gamma_l,
                                 const float *gamma_s, const float *xhat_l, const float *rstd_l,
                                 const float *xhat_s, const float *rstd_s, float *dh1,
                                 float *dgamma_l, float *dbeta_l, float *dgamma_s, float *dbeta_s,
                                 float *dWs, float *dbs, float *dbl, msr3d_stream_t stream) {
  MSR3D_REQUIRE(dq && sz && gamma_l && gamma_s && xhat_l && rstd_l && xhat_s && rstd_s && dh1,
                "qpos_bwd: null pointer");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(qpos_bwd_kernel, dim3(ln_grid(M)), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, 
      dq, dq1, dq2, sz, M, LnParams{gamma_l, nullptr}, LnParams{gamma_s, nullptr}, xhat_l, rstd_l, xhat_s,
      rstd_s, dh1, dgamma_l, dbeta_l, dgamma_s, dbeta_s, dWs, dbs, dbl);
  return check_launch("qpos_bwd");
}

MSR3D_API int msr3d_colsum_add_f32(const float *src, long long ld, int M, int N, float *dst0,
                                   float *dst1, float *dst2, msr3d_stream_t stream) {
  MSR3D_REQUIRE(src && dst0 && M >= 0 && N >= 0, "colsum: bad args");
  if (M == 0 || N == 0) return MSR3D_OK;
  const int rows = 32;
  dim3 grid(ceil_div(N, 256), ceil_div(M, rows));
  launch_pdl(colsum_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, src, ld, M, N, dst0, dst1, dst2, rows);
  return check_launch("colsum");
}

template <int FUSION>
static int launch_attn_fwd(const float *qkvc, int ldq, const float *pw, const unsigned char *key_pad, int B, int L,
                           int H, float *P, float *LOC, float *ctx, int ldc, msr3d_stream_t stream) {
  const int LT = ((L + 31) / 32) * 32;
  size_t smem = ((size_t)kDh * LT + (size_t)L * kDh) * sizeof(float);
  // ONE CTA PER SM: with programmatic dependent launch the CTAs of this kernel are placed while the GEMM in front
  // of it still owns most SMs (register file: they cannot co-reside), so the block scheduler packs them two per
  // SM onto the few free SMs and the kernel runs at half speed there — 14.6 us instead of 8.1 us in the forward
  // chain (tools/chain_profile.py).  Asking for more than half of the shared memory forces one CTA per SM.
  static bool configured = false;
  if (!configured) {
    MSR3D_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<FUSION>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOnePerSmBytes));
    configured = true;
  }
  if (g_tune[TUNE_ATTN_PACK] == 2 && smem < kOnePerSmBytes) smem = kOnePerSmBytes;   // A/B knob; default: pack (r2 sweep, SM partition)
  // few (head, scene) pairs and L ~ 60 rows: one warp per query row, rows spread over `zr` CTAs
  int warps = g_tune[TUNE_ATTN_FWD_WARPS] > 0 ? g_tune[TUNE_ATTN_FWD_WARPS] : 16;
  warps = max(1, min(kAttnMaxWarps, warps));
  int zr = g_tune[TUNE_ATTN_FWD_ZSPLIT] > 0 ? g_tune[TUNE_ATTN_FWD_ZSPLIT] : ceil_div(L, warps);
  zr = max(1, min(zr, 8));
  launch_pdl(attn_fwd_kernel<FUSION>, dim3(dim3(H, B, zr)), dim3(warps * 32), smem, (cudaStream_t)stream,
      qkvc, ldq, pw, key_pad, L, H, P, LOC, ctx, ldc);
  return check_launch("attn_fwd");
}

MSR3D_API int msr3d_attn_fusion_fwd_f32(const float *qkvc, int ldq, const float *pw,
                                        const unsigned char *key_pad, int B, int L, int H, int fusion, float *P,
                                        float *LOC, float *ctx, int ldc, msr3d_stream_t stream) {
  MSR3D_REQUIRE(qkvc && pw && P && LOC && ctx, "attn_fwd: null pointer");
  MSR3D_REQUIRE(L >= 1 && L <= kMaxL && H >= 1 && ldq >= 3 * H * kDh + H * 6 && ldc >= H * kDh,
                "attn_fwd: bad sizes L=%d H=%d", L, H);
  MSR3D_REQUIRE(fusion >= 0 && fusion <= 3, "attn_fwd: fusion %d (0 cond, 1 mul, 2 bias / ctx, 3 add)", fusion);
  if (B == 0) return MSR3D_OK;
  switch (fusion) {
    case kFuseMul: return launch_attn_fwd<kFuseMul>(qkvc, ldq, pw, key_pad, B, L, H, P, LOC, ctx, ldc, stream);
    case kFuseBias: return launch_attn_fwd<kFuseBias>(qkvc, ldq, pw, key_pad, B, L, H, P, LOC, ctx, ldc, stream);
    case kFuseAdd: return launch_attn_fwd<kFuseAdd>(qkvc, ldq, pw, key_pad, B, L, H, P, LOC, ctx, ldc, stream);
    default: return launch_attn_fwd<kFuseCond>(qkvc, ldq, pw, key_pad, B, L, H, P, LOC, ctx, ldc, stream);
  }
}

MSR3D_API int msr3d_attn_fwd_f32(const float *qkvc, int ldq, const float *pw,
                                 const unsigned char *key_pad, int B, int L, int H, float *P,
                                 float *LOC, float *ctx, int ldc, msr3d_stream_t stream) {
  return msr3d_attn_fusion_fwd_f32(qkvc, ldq, pw, key_pad, B, L, H, kFuseCond, P, LOC, ctx, ldc, stream);
}

template <int FUSION>
static int launch_attn_bwd(const float *dctx, int ldc, const float *qkvc, int ldq, const float *pw, const float *P,
                           const float *LOC, int B, int L, int H, float *dqkvc, msr3d_stream_t stream) {
  // cluster of `nz` CTAs per (head, scene): rows per CTA = ceil(L / nz); `warps` rows in flight per CTA
  int nz = g_tune[TUNE_ATTN_BWD_CLUSTER] > 0 ? g_tune[TUNE_ATTN_BWD_CLUSTER] : 4;
  nz = max(1, min(nz, 8));
  int warps = g_tune[TUNE_ATTN_BWD_WARPS] > 0 ? g_tune[TUNE_ATTN_BWD_WARPS] : 16;
  warps = max(1, min(kAttnMaxWarps, warps));
  auto bytes = [](int l, int z) {
    const int lt = ((l + 31) / 32) * 32;
    const int rpc = (l + z - 1) / z;
    return ((size_t)3 * l * kDh + (size_t)kDh * lt + (size_t)2 * rpc * (lt + 1)) * sizeof(float);
  };
  static bool configured = false;
  if (!configured) {
    const size_t mx = bytes(kMaxL, 1) > kOnePerSmBytes ? bytes(kMaxL, 1) : kOnePerSmBytes;
    MSR3D_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<FUSION>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
    configured = true;
  }
  size_t smem_bwd = bytes(L, nz);
  if (g_tune[TUNE_ATTN_PACK] == 2 && smem_bwd < kOnePerSmBytes) smem_bwd = kOnePerSmBytes;   // one CTA per SM, see attn_fwd
  launch_pdl_cluster(attn_bwd_kernel<FUSION>, dim3(H, B, nz), dim3(warps * 32), smem_bwd, (cudaStream_t)stream,
                     dim3(1, 1, nz), dctx, ldc, qkvc, ldq, pw, P, LOC, L, H, dqkvc);
  return check_launch("attn_bwd");
}

MSR3D_API int msr3d_attn_fusion_bwd_f32(const float *dctx, int ldc, const float *qkvc, int ldq,
                                        const float *pw, const float *P, const float *LOC, int B, int L,
                                        int H, int fusion, float *dqkvc, msr3d_stream_t stream) {
  MSR3D_REQUIRE(dctx && qkvc && pw && P && LOC && dqkvc, "attn_bwd: null pointer");
  MSR3D_REQUIRE(L >= 1 && L <= kMaxL && H >= 1, "attn_bwd: bad sizes");
  MSR3D_REQUIRE(fusion >= 0 && fusion <= 3, "attn_bwd: fusion %d (0 cond, 1 mul, 2 bias / ctx, 3 add)", fusion);
  if (B == 0) return MSR3D_OK;
  switch (fusion) {
    case kFuseMul: return launch_attn_bwd<kFuseMul>(dctx, ldc, qkvc, ldq, pw, P, LOC, B, L, H, dqkvc, stream);
    case kFuseBias: return launch_attn_bwd<kFuseBias>(dctx, ldc, qkvc, ldq, pw, P, LOC, B, L, H, dqkvc, stream);
    case kFuseAdd: return launch_attn_bwd<kFuseAdd>(dctx, ldc, qkvc, ldq, pw, P, LOC, B, L, H, dqkvc, stream);
    default: return launch_attn_bwd<kFuseCond>(dctx, ldc, qkvc, ldq, pw, P, LOC, B, L, H, dqkvc, stream);
  }
}

MSR3D_API int msr3d_attn_bwd_f32(const float *dctx, int ldc, const float *qkvc, int ldq,
                                 const float *pw, const float *P, const float *LOC, int B, int L,
                                 int H, float *dqkvc, msr3d_stream_t stream) {
  return msr3d_attn_fusion_bwd_f32(dctx, ldc, qkvc, ldq, pw, P, LOC, B, L, H, kFuseCond, dqkvc, stream);
}

MSR3D_API int msr3d_spatial_weights_fwd_f32(const float *W, const float *b, int ctx_mode, int multihead, float *qkvc,
                                            int ldq, int M, int H, msr3d_stream_t stream) {
  MSR3D_REQUIRE(W && b && qkvc, "spatial_weights_fwd: null pointer");
  MSR3D_REQUIRE(H >= 1 && H <= 10 && ldq >= 3 * H * kDh + H * 6, "spatial_weights_fwd: bad sizes H=%d ldq=%d", H, ldq);
  if (M <= 0) return MSR3D_OK;
  launch_pdl(spatial_weights_fwd_kernel, dim3(ceil_div(M * H, 256)), dim3(256), 0, (cudaStream_t)stream, W, b, ctx_mode,
             multihead, qkvc, ldq, M, H);
  return check_launch("spatial_weights_fwd");
}

MSR3D_API int msr3d_spatial_weights_bwd_f32(const float *W, const float *b, int ctx_mode, int multihead,
                                            const float *qkvc, float *dqkvc, int ldq, int M, int H, float *dW,
                                            float *db, msr3d_stream_t stream) {
  MSR3D_REQUIRE(W && b && qkvc && dqkvc && dW && db, "spatial_weights_bwd: null pointer");
  MSR3D_REQUIRE(H >= 1 && H <= 10 && ldq >= 3 * H * kDh + H * 6, "spatial_weights_bwd: bad sizes H=%d ldq=%d", H, ldq);
  MSR3D_REQUIRE(!ctx_mode || H * kDh == 256, "spatial_weights_bwd: ctx fusion needs d_model = 256");
  if (M <= 0) return MSR3D_OK;
  launch_pdl(spatial_weights_bwd_kernel, dim3(ctx_mode ? M + 1 : 1), dim3(256), 0, (cudaStream_t)stream, W, b,
             ctx_mode, multihead, qkvc, dqkvc, ldq, M, H, dW, db);
  return check_launch("spatial_weights_bwd");
}

/* Test hook: the multiplicative mask (0 or 1/(1-p)) every kernel applies at `site` to element
 * offsets [0, n). */
MSR3D_API int msr3d_dropout_mask_f32(unsigned long long seed, const unsigned long long *seed_ptr,
                                     float p, unsigned site, long long n, float *out,
                                     msr3d_stream_t stream) {
  MSR3D_REQUIRE(out && n >= 0 && p >= 0.f && p < 1.f, "dropout_mask: bad args");
  if (n == 0) return MSR3D_OK;
  launch_pdl(dropout_mask_kernel, dim3(kNumSMs * 2), dim3(256), 0, (cudaStream_t)stream, mk(seed, p, seed_ptr), site, n, out);
  return check_launch("dropout_mask");
}
