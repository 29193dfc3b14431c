// Unfrozen-encoder training path (modules/vision/pcd_pointnet_encoder.py:58-62): the pieces of a
// set-abstraction level around its tensor-core GEMMs when BatchNorm runs in TRAINING mode and gradients flow.
//
// Layout: everything is ROW-major [row][channel] with row = (cloud, centre, sample) — the K-major operand the GEMM
// kernels read directly (msr3d_gemm_tc_f32: Z = X W^T, dX = dZ W, dW = dZ^T X).  The reference's (B, C, npoint,
// nsample) tensors (pointnet2_utils.py:314-373, pytorch_utils.py:11-58) never exist.
//
//   group_rows        X[row] = [xyz[idx] - centre (3) | feat[idx] (C)]          QueryAndGroup / GroupAll forward
//   bn_stats (+final) per-channel batch mean / biased variance over all rows, two-stage and deterministic;
//                     running statistics updated as nn.BatchNorm2d does (momentum 0.1, unbiased variance)
//   bn_relu           Y = relu(gamma * (Z - mean) * rstd + beta)
//   maxpool / _bwd    max over the nsample rows of a centre (first index wins ties) / routing of the gradient
//   bn_bwd_stats (+final), bn_bwd_apply
//                     g = dY * (Y > 0);  dgamma += sum g xhat;  dbeta += sum g;
//                     dZ = gamma rstd (g - mean(g) - xhat mean(g xhat))   (eval-mode BN: dZ = gamma rstd g)
//   group_rows_bwd    d feat[b, j, :] = sum over the (centre, sample) pairs with idx == j — one CTA per cloud,
//                     thread = channel, pairs walked in order: deterministic, no atomics
//                     (replaces group_points_grad's atomicAdd, _ext_src/src/group_points_gpu.cu:43-75)
#include "common.cuh"

#include <algorithm>

namespace msr3d {

constexpr int kStatCols = 32;      // columns per CTA of the statistics kernels (8 threads x one float4)
constexpr int kStatLanes = 32;     // row lanes per CTA (256 threads)

__global__ void __launch_bounds__(256)
group_rows_kernel(const float *__restrict__ xyz, int ld_xyz, const float *__restrict__ feat, int ld_feat, int C,
                  const float *__restrict__ new_xyz, const int *__restrict__ idx, int N, int groups_per_cloud,
                  int nsample, long long rows, float *__restrict__ X, int ldx) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  long long b, src;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (idx != nullptr) {
    const long long grp = row / nsample;
    b = grp / groups_per_cloud;
    src = b * N + idx[row];
    const float *c = new_xyz + grp * 3;
    cx = c[0]; cy = c[1]; cz = c[2];
  } else {            // GroupAll: row = point, xyz not centred (pointnet2_utils.py:407)
    src = row;
  }
  float *o = X + row * ldx;
  if (lane < 3) {
    const float v = xyz[src * ld_xyz + lane];
    o[lane] = __fsub_rn(v, lane == 0 ? cx : (lane == 1 ? cy : cz));
  }
  const float *f = feat + src * ld_feat;
  for (int c = lane; c < C; c += 32) o[3 + c] = f[c];
  if (lane < ldx - (3 + C)) o[3 + C + lane] = 0.f;      // pad columns (row pitch rounded up for TMA): zero
}

// Column sums over a chunk of rows.  Thread = (row lane 0..31, column quad 0..7): one 16 B load per row, four rows of
// 128 B per warp instruction, the row loop unrolled so that several loads are in flight; the 32 row lanes meet in
// shared memory.  (The first version read one float per thread with 8 row lanes and 64 chunks: 15-25 % of the HBM
// rate on the 126 MB activations of SA1, tools/profile_enc_train.py.)
template <int NARR>
__device__ __forceinline__ void stat_reduce_store(const float4 (&acc)[NARR], float *__restrict__ partial, int C, int c0) {
  __shared__ float s[NARR][kStatLanes][kStatCols + 1];
  const int cq = threadIdx.x & 7, rl = threadIdx.x >> 3;
#pragma unroll
  for (int a = 0; a < NARR; ++a) {
    s[a][rl][cq * 4 + 0] = acc[a].x; s[a][rl][cq * 4 + 1] = acc[a].y;
    s[a][rl][cq * 4 + 2] = acc[a].z; s[a][rl][cq * 4 + 3] = acc[a].w;
  }
  __syncthreads();
  if (threadIdx.x < kStatCols * NARR) {
    const int a = threadIdx.x / kStatCols, col = threadIdx.x % kStatCols;
    if (c0 + col < C) {
      float t = 0.f;
#pragma unroll 8
      for (int i = 0; i < kStatLanes; ++i) t += s[a][i][col];      // fixed order: deterministic
      partial[((size_t)blockIdx.y * NARR + a) * C + c0 + col] = t;
    }
  }
}

// partial[chunk][0][c] = sum_rows Z, partial[chunk][1][c] = sum_rows Z^2 over the rows of the chunk
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float *__restrict__ Z, long long rows, int C, int rows_per_chunk, float *__restrict__ partial) {
  pdl_sync();
  const int cq = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * kStatCols, c = c0 + cq * 4;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  if (c < C) {
#pragma unroll 4
    for (long long r = r0 + rl; r < r1; r += kStatLanes) {
      const float4 v = *reinterpret_cast<const float4 *>(Z + r * C + c);
      acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
      acc[1].x += v.x * v.x; acc[1].y += v.y * v.y; acc[1].z += v.z * v.z; acc[1].w += v.w * v.w;
    }
  }
  stat_reduce_store<2>(acc, partial, C, c0);
}

// Sum of the chunk partials of 32 columns: thread = (column, lane of 8); lane l adds chunks l, l + 8, ... in double
// (independent loads), the eight lanes meet in shared memory in a fixed order.  (One thread per column walking all
// chunks was a chain of dependent L2 round trips: 68 us per BatchNorm layer for a kernel that moves 0.5 MB, ncu.)
__device__ __forceinline__ void sum_partials(const float *__restrict__ partial, int chunks, int C, int c, double &a,
                                             double &q) {
  __shared__ double s_a[8][kStatCols], s_q[8][kStatCols];
  const int col = threadIdx.x & 31, l = threadIdx.x >> 5;
  double ta = 0.0, tq = 0.0;
  if (c < C) {
#pragma unroll 4
    for (int k = l; k < chunks; k += 8) {
      ta += (double)partial[((size_t)k * 2 + 0) * C + c];
      tq += (double)partial[((size_t)k * 2 + 1) * C + c];
    }
  }
  s_a[l][col] = ta;
  s_q[l][col] = tq;
  __syncthreads();
  a = 0.0; q = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { a += s_a[i][col]; q += s_q[i][col]; }
}

// stats[0][c] = mean, stats[1][c] = rstd; running statistics as nn.BatchNorm2d (momentum m, unbiased variance)
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float *__restrict__ partial, int chunks, int C, long long rows, float eps,
                   float momentum, float *__restrict__ stats, float *__restrict__ running_mean,
                   float *__restrict__ running_var) {
  pdl_sync();
  const int c = blockIdx.x * kStatCols + (threadIdx.x & 31);
  double a, q;
  sum_partials(partial, chunks, C, c, a, q);
  if (c >= C || threadIdx.x >= kStatCols) return;
  const double mean = a / (double)rows;
  double var = q / (double)rows - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[c] = (float)mean;
  stats[C + c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

// eval-mode BatchNorm: stats from the running statistics
__global__ void bn_running_stats_kernel(const float *__restrict__ running_mean, const float *__restrict__ running_var,
                                        int C, float eps, float *__restrict__ stats) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  stats[c] = running_mean[c];
  stats[C + c] = 1.f / sqrtf(running_var[c] + eps);
}

__global__ void __launch_bounds__(256)
bn_relu_kernel(const float *__restrict__ Z, const float *__restrict__ stats, const float *__restrict__ gamma,
               const float *__restrict__ beta, long long n4_, int C, float *__restrict__ Y) {
  pdl_sync();
  const unsigned n4 = (unsigned)n4_, cols4 = (unsigned)C >> 2;       // 32-bit index arithmetic (n4 < 2^31 checked on the host)
#pragma unroll 4
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const int c = (int)((i % cols4) << 2);
    const float4 z = reinterpret_cast<const float4 *>(Z)[i];
    const float zz[4] = {z.x, z.y, z.z, z.w};
    float y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (zz[k] - stats[c + k]) * stats[C + c + k];
      y[k] = fmaxf(xh * gamma[c + k] + beta[c + k], 0.f);
    }
    reinterpret_cast<float4 *>(Y)[i] = make_float4(y[0], y[1], y[2], y[3]);
  }
}

__global__ void __launch_bounds__(256)
maxpool_rows_kernel(const float *__restrict__ Y, long long groups, int S, int C, float *__restrict__ out,
                    long long ld_out, int *__restrict__ arg) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * C) return;
  const long long g = i / C;
  const int c = (int)(i % C);
  const float *p = Y + (g * S) * C + c;
  float best = p[0];
  int bi = 0;
  for (int s = 1; s < S; ++s) {
    const float v = p[(long long)s * C];
    if (v > best) { best = v; bi = s; }
  }
  out[g * ld_out + c] = best;
  arg[i] = bi;
}

// dY[row, c] = dOut[group, c] if this row won the max (and the ReLU in front of the pool was active) else 0.
// Thread = (row, column quad): one 16 B arg-max load, one 16 B store; 32-bit index arithmetic (the first version did
// four 64-bit divisions per element and ran at a quarter of the store rate).
__global__ void __launch_bounds__(256)
maxpool_rows_bwd_kernel(const float *__restrict__ dOut, long long ld_dout, const int *__restrict__ arg,
                        long long rows_, int S, int C, float *__restrict__ dY) {
  pdl_sync();
  const unsigned cols4 = (unsigned)C >> 2, n4 = (unsigned)rows_ * cols4, uS = (unsigned)S;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const unsigned row = i / cols4, c = (i - row * cols4) << 2;
    const unsigned g = row / uS;
    const int s = (int)(row - g * uS);
    const int4 a = *reinterpret_cast<const int4 *>(arg + (size_t)g * C + c);
    const float *d = dOut + (size_t)g * ld_dout + c;          // may be a column slice of a wider matrix: scalar loads
    float4 o;
    o.x = a.x == s ? d[0] : 0.f;
    o.y = a.y == s ? d[1] : 0.f;
    o.z = a.z == s ? d[2] : 0.f;
    o.w = a.w == s ? d[3] : 0.f;
    *reinterpret_cast<float4 *>(dY + (size_t)row * C + c) = o;
  }
}

// partial[chunk][0][c] = sum g, partial[chunk][1][c] = sum g * xhat, g = dY * (Y > 0)
__global__ void __launch_bounds__(256)
bn_bwd_stats_kernel(const float *__restrict__ dY, const float *__restrict__ Y, const float *__restrict__ Z,
                    const float *__restrict__ stats, long long rows, int C, int rows_per_chunk,
                    float *__restrict__ partial) {
  pdl_sync();
  const int cq = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * kStatCols, c = c0 + cq * 4;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  if (c < C) {
    const float4 mean = *reinterpret_cast<const float4 *>(stats + c), rstd = *reinterpret_cast<const float4 *>(stats + C + c);
#pragma unroll 2
    for (long long r = r0 + rl; r < r1; r += kStatLanes) {
      const float4 d = *reinterpret_cast<const float4 *>(dY + r * C + c);
      const float4 y = *reinterpret_cast<const float4 *>(Y + r * C + c);
      const float4 z = *reinterpret_cast<const float4 *>(Z + r * C + c);
      const float gx = y.x > 0.f ? d.x : 0.f, gy = y.y > 0.f ? d.y : 0.f, gz = y.z > 0.f ? d.z : 0.f, gw = y.w > 0.f ? d.w : 0.f;
      acc[0].x += gx; acc[0].y += gy; acc[0].z += gz; acc[0].w += gw;
      acc[1].x += gx * ((z.x - mean.x) * rstd.x); acc[1].y += gy * ((z.y - mean.y) * rstd.y);
      acc[1].z += gz * ((z.z - mean.z) * rstd.z); acc[1].w += gw * ((z.w - mean.w) * rstd.w);
    }
  }
  stat_reduce_store<2>(acc, partial, C, c0);
}

// bstat[0][c] = mean(g), bstat[1][c] = mean(g xhat) (zero for eval-mode BN); dgamma += sum g xhat, dbeta += sum g
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float *__restrict__ partial, int chunks, int C, long long rows,
                       int batch_stats, float *__restrict__ bstat, float *__restrict__ dgamma,
                       float *__restrict__ dbeta) {
  pdl_sync();
  const int c = blockIdx.x * kStatCols + (threadIdx.x & 31);
  double a, q;
  sum_partials(partial, chunks, C, c, a, q);
  if (c >= C || threadIdx.x >= kStatCols) return;
  dbeta[c] += (float)a;
  dgamma[c] += (float)q;
  bstat[c] = batch_stats ? (float)(a / (double)rows) : 0.f;
  bstat[C + c] = batch_stats ? (float)(q / (double)rows) : 0.f;
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float *dY /* may alias dZ */, const float *__restrict__ Y, const float *__restrict__ Z,
                    const float *__restrict__ stats, const float *__restrict__ bstat,
                    const float *__restrict__ gamma, long long n4_, int C, float *dZ) {
  pdl_sync();
  const unsigned n4 = (unsigned)n4_, cols4 = (unsigned)C >> 2;
#pragma unroll 2
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const int c = (int)((i % cols4) << 2);
    const float4 d4 = reinterpret_cast<const float4 *>(dY)[i];
    const float4 y4 = reinterpret_cast<const float4 *>(Y)[i];
    const float4 z4 = reinterpret_cast<const float4 *>(Z)[i];
    const float d[4] = {d4.x, d4.y, d4.z, d4.w}, y[4] = {y4.x, y4.y, y4.z, y4.w}, z[4] = {z4.x, z4.y, z4.z, z4.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float rstd = stats[C + c + k];
      const float xh = (z[k] - stats[c + k]) * rstd;
      const float g = y[k] > 0.f ? d[k] : 0.f;
      o[k] = gamma[c + k] * rstd * (g - bstat[c + k] - xh * bstat[C + c + k]);
    }
    reinterpret_cast<float4 *>(dZ)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// d feat[b, j, 0:C] = sum over pairs p of cloud b with idx[b, p] == j of dX[b, p, col0 : col0 + C].
// One CTA per cloud.  The pairs are cut into `nsl` consecutive slices, each with its own [N][C] accumulator in shared
// memory; a thread owns (slice, channel), walks its slice in order with eight loads in flight, and the slices are added
// in a fixed order at the end: deterministic, no atomics.  (One slice and one load in flight: 214 us for 240 clouds.)
__global__ void __launch_bounds__(512)
group_rows_bwd_kernel(const float *__restrict__ dX, int ldx, int col0, const int *__restrict__ idx, int pairs, int N,
                      int C, int nsl, float *__restrict__ dfeat) {
  pdl_sync();
  extern __shared__ float acc[];     // [nsl][N][C], then int idx[pairs]
  int *s_idx = reinterpret_cast<int *>(acc + (size_t)nsl * N * C);
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < nsl * N * C; i += blockDim.x) acc[i] = 0.f;
  for (int i = threadIdx.x; i < pairs; i += blockDim.x) s_idx[i] = idx[(size_t)b * pairs + i];
  __syncthreads();
  const int tps = blockDim.x / nsl;                   // threads per slice
  const int sl = threadIdx.x / tps, t = threadIdx.x % tps;
  const int pps = (pairs + nsl - 1) / nsl;
  const int p0 = sl * pps, p1 = min(pairs, p0 + pps);
  const float *xb = dX + (size_t)b * pairs * ldx + col0;
  float *a = acc + (size_t)sl * N * C;
  for (int c = t; c < C; c += tps) {
    for (int p = p0; p < p1; p += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = p + u < p1 ? xb[(size_t)(p + u) * ldx + c] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (p + u < p1) a[s_idx[p + u] * C + c] += v[u];
    }
  }
  __syncthreads();
  float *o = dfeat + (size_t)b * N * C;
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
    float tsum = acc[i];
    for (int q = 1; q < nsl; ++q) tsum += acc[(size_t)q * N * C + i];
    o[i] = tsum;
  }
}

static inline int stat_chunks(long long rows, int *rows_per_chunk) {
  int chunks = (int)std::min<long long>(256, std::max<long long>(1, rows / 256));
  *rows_per_chunk = (int)((rows + chunks - 1) / chunks);
  return chunks;
}

}  // namespace msr3d

using namespace msr3d;

MSR3D_API int msr3d_sa_train_stat_chunks(long long rows) {
  int rpc;
  return stat_chunks(rows, &rpc);
}

MSR3D_API int msr3d_sa_group_rows_f32(const float *xyz, int ld_xyz, const float *feat, int ld_feat, int C,
                                      const float *new_xyz, const int32_t *idx, int N, int groups_per_cloud,
                                      int nsample, long long rows, float *X, int ldx, msr3d_stream_t stream) {
  MSR3D_REQUIRE(xyz && feat && X && C >= 1 && ldx >= 3 + C && ldx < 3 + C + 32, "sa_group_rows: null pointer / bad C / ldx");
  MSR3D_REQUIRE((idx == nullptr) == (new_xyz == nullptr), "sa_group_rows: idx and new_xyz come together");
  MSR3D_REQUIRE(idx == nullptr || (N >= 1 && groups_per_cloud >= 1 && nsample >= 1), "sa_group_rows: bad sizes");
  if (rows <= 0) return MSR3D_OK;
  launch_pdl(group_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, xyz, ld_xyz, feat,
             ld_feat, C, new_xyz, (const int *)idx, N, groups_per_cloud, nsample, rows, X, ldx);
  return check_launch("sa_group_rows");
}

MSR3D_API int msr3d_bn_train_fwd_f32(const float *Z, long long rows, int C, const float *gamma, const float *beta,
                                     float *running_mean, float *running_var, int batch_stats, float eps,
                                     float momentum, float *partial, float *stats, float *Y, msr3d_stream_t stream) {
  MSR3D_REQUIRE(Z && gamma && beta && partial && stats && Y, "bn_train_fwd: null pointer");
  MSR3D_REQUIRE(rows >= 1 && C >= 4 && C % 4 == 0 && rows * C / 4 < (1ll << 31), "bn_train_fwd: rows=%lld C=%d (C %% 4, rows * C < 2^33)", rows, C);
  MSR3D_REQUIRE(batch_stats || (running_mean && running_var), "bn_train_fwd: eval mode needs running statistics");
  cudaStream_t st = (cudaStream_t)stream;
  if (batch_stats) {
    int rpc;
    const int chunks = stat_chunks(rows, &rpc);
    launch_pdl(bn_stats_kernel, dim3(ceil_div(C, kStatCols), chunks), dim3(256), 0, st, Z, rows, C, rpc, partial);
    launch_pdl(bn_finalize_kernel, dim3(ceil_div(C, kStatCols)), dim3(256), 0, st, (const float *)partial, chunks, C, rows, eps,
               momentum, stats, running_mean, running_var);
  } else {
    launch_pdl(bn_running_stats_kernel, dim3(ceil_div(C, 128)), dim3(128), 0, st, (const float *)running_mean,
               (const float *)running_var, C, eps, stats);
  }
  const long long n4 = rows * C / 4;
  const int grid = (int)std::min<long long>((n4 + 255) / 256, kNumSMs * 8);
  launch_pdl(bn_relu_kernel, dim3(grid), dim3(256), 0, st, Z, (const float *)stats, gamma, beta, n4, C, Y);
  return check_launch("bn_train_fwd");
}

MSR3D_API int msr3d_bn_train_bwd_f32(const float *dY, const float *Y, const float *Z, long long rows, int C,
                                     const float *gamma, const float *stats, int batch_stats, float *partial,
                                     float *bstat, float *dgamma, float *dbeta, float *dZ, msr3d_stream_t stream) {
  MSR3D_REQUIRE(dY && Y && Z && gamma && stats && partial && bstat && dgamma && dbeta && dZ, "bn_train_bwd: null pointer");
  MSR3D_REQUIRE(rows >= 1 && C >= 4 && C % 4 == 0 && rows * C / 4 < (1ll << 31), "bn_train_bwd: rows=%lld C=%d (C %% 4, rows * C < 2^33)", rows, C);
  cudaStream_t st = (cudaStream_t)stream;
  int rpc;
  const int chunks = stat_chunks(rows, &rpc);
  launch_pdl(bn_bwd_stats_kernel, dim3(ceil_div(C, kStatCols), chunks), dim3(256), 0, st, dY, Y, Z, stats, rows, C, rpc,
             partial);
  launch_pdl(bn_bwd_finalize_kernel, dim3(ceil_div(C, kStatCols)), dim3(256), 0, st, (const float *)partial, chunks, C, rows,
             batch_stats, bstat, dgamma, dbeta);
  const long long n4 = rows * C / 4;
  const int grid = (int)std::min<long long>((n4 + 255) / 256, kNumSMs * 8);
  launch_pdl(bn_bwd_apply_kernel, dim3(grid), dim3(256), 0, st, dY, Y, Z, stats, (const float *)bstat, gamma, n4, C, dZ);
  return check_launch("bn_train_bwd");
}

MSR3D_API int msr3d_maxpool_rows_f32(const float *Y, long long groups, int S, int C, float *out, long long ld_out,
                                     int32_t *arg, msr3d_stream_t stream) {
  MSR3D_REQUIRE(Y && out && arg && S >= 1 && C >= 1 && ld_out >= C, "maxpool_rows: bad args");
  if (groups <= 0) return MSR3D_OK;
  launch_pdl(maxpool_rows_kernel, dim3((unsigned)((groups * C + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, Y,
             groups, S, C, out, ld_out, (int *)arg);
  return check_launch("maxpool_rows");
}

MSR3D_API int msr3d_maxpool_rows_bwd_f32(const float *dOut, long long ld_dout, const int32_t *arg, long long groups,
                                         int S, int C, float *dY, msr3d_stream_t stream) {
  MSR3D_REQUIRE(dOut && arg && dY && S >= 1 && C >= 4 && C % 4 == 0 && ld_dout >= C && groups * S * C / 4 < (1ll << 31),
                "maxpool_rows_bwd: bad args (C %% 4, rows * C < 2^33)");
  if (groups <= 0) return MSR3D_OK;
  const long long n = groups * S * C;
  const int grid = (int)std::min<long long>((n + 255) / 256, kNumSMs * 8);
  launch_pdl(maxpool_rows_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, dOut, ld_dout, (const int *)arg,
             groups * S, S, C, dY);
  return check_launch("maxpool_rows_bwd");
}

MSR3D_API int msr3d_sa_group_rows_bwd_f32(const float *dX, int ldx, int col0, const int32_t *idx, int B, int pairs,
                                          int N, int C, float *dfeat, msr3d_stream_t stream) {
  MSR3D_REQUIRE(dX && idx && dfeat && pairs >= 1 && N >= 1 && C >= 1 && col0 >= 0 && col0 + C <= ldx,
                "sa_group_rows_bwd: bad args");
  const size_t one = (size_t)N * C * sizeof(float), idx_bytes = (size_t)pairs * sizeof(int);
  MSR3D_REQUIRE(one + idx_bytes <= 200 * 1024, "sa_group_rows_bwd: N*C = %d floats do not fit in shared memory", N * C);
  if (B <= 0) return MSR3D_OK;
  int nsl = 4;
  while (nsl > 1 && nsl * one + idx_bytes > 200 * 1024) nsl >>= 1;
  const size_t smem = nsl * one + idx_bytes;
  static bool configured = false;
  if (!configured) {
    MSR3D_CUDA(cudaFuncSetAttribute(group_rows_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  launch_pdl(group_rows_bwd_kernel, dim3(B), dim3(512), smem, (cudaStream_t)stream, dX, ldx, col0, (const int *)idx, pairs,
             N, C, nsl, dfeat);
  return check_launch("sa_group_rows_bwd");
}
