// Group F: per-object point-cloud preprocessing on the device — what the reference
// does on the CPU in the dataset, data/datasets/msr3d.py:181-215 (`preprocess_pcd`) followed by the
// padding of dataset_wrapper.py:151-158.  The random draws (axis rotation, subsample indices) stay on
// the host and are inputs, so results can be compared with the reference draw for draw.
#include "common.cuh"

#include <cooperative_groups.h>

namespace msr3d {

constexpr int kPreThreads = 256;
constexpr int kPreCluster = 8;     // CTAs per object slot (thread-block cluster): pass 1 walks up to ~100 k points

__device__ __forceinline__ float block_reduce(float v, float *s_red, int op) {  // 0 sum, 1 min, 2 max
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = op == 0 ? v + u : (op == 1 ? fminf(v, u) : fmaxf(v, u));
  }
  __syncthreads();   // s_red may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s_red[0];
#pragma unroll
  for (int w = 1; w < kPreThreads / 32; ++w) r = op == 0 ? r + s_red[w] : (op == 1 ? fminf(r, s_red[w]) : fmaxf(r, s_red[w]));
  return r;
}

// One CLUSTER of kPreCluster CTAs per output object slot.  Slots >= n_obj are padding: points of ones, zero
// location, mask 0.
//   pass 1 (all points of the object, split over the CTAs of the cluster; partial sums / bounds meet in rank 0
//          through distributed shared memory, fixed order): rotate, centre = mean, size = max - min   -> obj_locs
//   pass 2 (the P sampled points, rank 0): rotate, subtract their mean, divide by the largest norm
//          (1 when it is < 1e-6)                                                        -> obj_fts
// Sums are accumulated in double per thread (numpy's pairwise float32 sums are within 1 ulp of that).
// Round 1 ran one CTA per slot: 60 CTAs on 148 SMs, the largest object (tens of thousands of points) on one SM —
// 10 % of the HBM roof (profiles/r1_next_rows.md).
__global__ void __launch_bounds__(kPreThreads)
preprocess_objects_kernel(const float *__restrict__ points, const int *__restrict__ offsets,
                          const int *__restrict__ sample_idx, const float *__restrict__ rot, int n_obj, int P,
                          float *__restrict__ obj_fts, float *__restrict__ obj_locs,
                          unsigned char *__restrict__ obj_masks) {
  pdl_sync();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), nz = (int)cluster.num_blocks();
  extern __shared__ float s_xyz[];      // [P][3] rotated coordinates of the sampled points
  __shared__ float s_red[kPreThreads / 32];
  __shared__ double s_sum[3];           // this CTA's partial sums of pass 1
  __shared__ float s_mm[6];             // ... and bounds (min xyz, max xyz)
  const int o = blockIdx.x, tid = threadIdx.x;
  float *fo = obj_fts + (size_t)o * P * 6;
  if (o >= n_obj) {                     // the whole cluster takes this branch (same slot)
    for (int i = rank * kPreThreads + tid; i < P * 6; i += nz * kPreThreads) fo[i] = 1.f;
    if (rank == 0 && tid < 6) obj_locs[(size_t)o * 6 + tid] = 0.f;
    if (rank == 0 && tid == 0) obj_masks[o] = 0;
    return;
  }
  float r[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  const bool has_rot = rot != nullptr;
  if (has_rot) {
#pragma unroll
    for (int i = 0; i < 9; ++i) r[i] = rot[i];
  }
  // p' = p R^T: p'_j = sum_k p_k R[j][k]  (np.matmul(xyz, rot.T), float32 products summed left to right)
  auto rotate = [&](float x, float y, float z, float &ox, float &oy, float &oz) {
    if (has_rot) {
      ox = __fadd_rn(__fadd_rn(__fmul_rn(x, r[0]), __fmul_rn(y, r[1])), __fmul_rn(z, r[2]));
      oy = __fadd_rn(__fadd_rn(__fmul_rn(x, r[3]), __fmul_rn(y, r[4])), __fmul_rn(z, r[5]));
      oz = __fadd_rn(__fadd_rn(__fmul_rn(x, r[6]), __fmul_rn(y, r[7])), __fmul_rn(z, r[8]));
    } else { ox = x; oy = y; oz = z; }
  };
  const int p0 = offsets[o], n = offsets[o + 1] - p0;
  const float *src = points + (size_t)p0 * 6;
  double sx = 0.0, sy = 0.0, sz = 0.0;
  float mnx = INFINITY, mny = INFINITY, mnz = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mxz = -INFINITY;
  for (int i = rank * kPreThreads + tid; i < n; i += nz * kPreThreads) {
    float x, y, z;
    rotate(src[(size_t)i * 6], src[(size_t)i * 6 + 1], src[(size_t)i * 6 + 2], x, y, z);
    sx += x; sy += y; sz += z;
    mnx = fminf(mnx, x); mny = fminf(mny, y); mnz = fminf(mnz, z);
    mxx = fmaxf(mxx, x); mxy = fmaxf(mxy, y); mxz = fmaxf(mxz, z);
  }
  // double sums through two float reductions would lose the point: reduce hi and lo parts
  auto reduce_d = [&](double v) -> double {
    const float hi = (float)v, lo = (float)(v - (double)hi);
    return (double)block_reduce(hi, s_red, 0) + (double)block_reduce(lo, s_red, 0);
  };
  {
    const double tx = reduce_d(sx), ty = reduce_d(sy), tz = reduce_d(sz);
    const float lox = block_reduce(mnx, s_red, 1), loy = block_reduce(mny, s_red, 1), loz = block_reduce(mnz, s_red, 1);
    const float hix = block_reduce(mxx, s_red, 2), hiy = block_reduce(mxy, s_red, 2), hiz = block_reduce(mxz, s_red, 2);
    if (tid == 0) {
      s_sum[0] = tx; s_sum[1] = ty; s_sum[2] = tz;
      s_mm[0] = lox; s_mm[1] = loy; s_mm[2] = loz; s_mm[3] = hix; s_mm[4] = hiy; s_mm[5] = hiz;
    }
  }
  cluster.sync();                       // every CTA's partials are visible cluster-wide
  if (rank == 0 && tid == 0) {
    double t[3] = {0.0, 0.0, 0.0};
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int q = 0; q < nz; ++q) {      // fixed order: deterministic
      const double *ps = cluster.map_shared_rank(s_sum, q);
      const float *pm = cluster.map_shared_rank(s_mm, q);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        t[d] += ps[d];
        lo[d] = fminf(lo[d], pm[d]);
        hi[d] = fmaxf(hi[d], pm[3 + d]);
      }
    }
    float *l = obj_locs + (size_t)o * 6;
    l[0] = (float)(t[0] / n); l[1] = (float)(t[1] / n); l[2] = (float)(t[2] / n);
    l[3] = hi[0] - lo[0]; l[4] = hi[1] - lo[1]; l[5] = hi[2] - lo[2];
    obj_masks[o] = 1;
  }
  cluster.sync();                       // nobody exits while rank 0 may still be reading its shared memory
  if (rank != 0) return;                // the P sampled points are one CTA's work
  // ---- the sampled points
  const int *si = sample_idx + (size_t)o * P;
  double ax = 0.0, ay = 0.0, az = 0.0;
  for (int j = tid; j < P; j += kPreThreads) {
    const float *q = src + (size_t)si[j] * 6;
    float x, y, z;
    rotate(q[0], q[1], q[2], x, y, z);
    s_xyz[j * 3 + 0] = x; s_xyz[j * 3 + 1] = y; s_xyz[j * 3 + 2] = z;
    ax += x; ay += y; az += z;
    fo[(size_t)j * 6 + 3] = q[3]; fo[(size_t)j * 6 + 4] = q[4]; fo[(size_t)j * 6 + 5] = q[5];
  }
  const float mx = (float)(reduce_d(ax) / P), my = (float)(reduce_d(ay) / P), mz = (float)(reduce_d(az) / P);
  float far = 0.f;
  for (int j = tid; j < P; j += kPreThreads) {
    const float x = __fsub_rn(s_xyz[j * 3 + 0], mx), y = __fsub_rn(s_xyz[j * 3 + 1], my), z = __fsub_rn(s_xyz[j * 3 + 2], mz);
    s_xyz[j * 3 + 0] = x; s_xyz[j * 3 + 1] = y; s_xyz[j * 3 + 2] = z;
    far = fmaxf(far, sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))));
  }
  float md = block_reduce(far, s_red, 2);
  if (md < 1e-6f) md = 1.f;      // tiny point clouds, i.e. padding (msr3d.py:207-208)
  for (int j = tid; j < P; j += kPreThreads) {
    fo[(size_t)j * 6 + 0] = s_xyz[j * 3 + 0] / md;
    fo[(size_t)j * 6 + 1] = s_xyz[j * 3 + 1] / md;
    fo[(size_t)j * 6 + 2] = s_xyz[j * 3 + 2] / md;
  }
}

}  // namespace msr3d

using namespace msr3d;

MSR3D_API int msr3d_preprocess_objects_f32(const float *points, const int32_t *offsets, const int32_t *sample_idx,
                                           const float *rot, int n_obj, int o_max, int P, float *obj_fts,
                                           float *obj_locs, uint8_t *obj_masks, msr3d_stream_t stream) {
  MSR3D_REQUIRE(obj_fts && obj_locs && obj_masks, "preprocess_objects: null output");
  MSR3D_REQUIRE(n_obj >= 0 && o_max >= n_obj && P >= 1, "preprocess_objects: bad sizes n_obj=%d o_max=%d P=%d", n_obj, o_max, P);
  MSR3D_REQUIRE(n_obj == 0 || (points && offsets && sample_idx), "preprocess_objects: null input");
  MSR3D_REQUIRE((size_t)P * 12 <= 96 * 1024, "preprocess_objects: P=%d sampled points do not fit in shared memory", P);
  if (o_max == 0) return MSR3D_OK;
  const size_t smem = (size_t)P * 3 * sizeof(float);
  static bool configured = false;
  if (!configured) {
    MSR3D_CUDA(cudaFuncSetAttribute(preprocess_objects_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  launch_pdl_cluster(preprocess_objects_kernel, dim3(o_max, 1, kPreCluster), dim3(kPreThreads), smem, (cudaStream_t)stream,
                     dim3(1, 1, kPreCluster), points, offsets, sample_idx, rot, n_obj, P, obj_fts, obj_locs, obj_masks);
  return check_launch("preprocess_objects");
}
