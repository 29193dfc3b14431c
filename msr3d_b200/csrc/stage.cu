// Input staging of the pipelined training step as ONE launch.
//
// The prefetch schedule (msr3d_b200/fused.py) keeps three generations of the small per-scene inputs in flight
// (consumed now / encoded last step / just copied from the host) and hands the frozen encoder's output and the
// pairwise geometry from the step that produced them to the step that consumes them.  As torch copies that was
// eleven dependent ~1.5 us memcpy nodes plus a fill and a logical_not at the head of every CUDA-graph replay
// (tools/step_timeline.py: 24 us before the first useful kernel).  Here: every element is moved by one thread,
//   rotate:  a[i] <- b[i];  b[i] <- c[i]      (three generations; same thread reads b before it overwrites it)
//   copy:    dst[i] <- src[i]
//   masks:   rotate on bytes, and key_pad[i] <- !a[i]   (transformers.py:228: padded keys)
#include "common.cuh"

namespace msr3d {

struct StageParams {
  msr3d_stage_desc d;
  long long start4[2 * MSR3D_STAGE_MAX + 1];   // flat kernel: first float4 index of each segment (rotates, then copies)
};

__device__ __forceinline__ void stage_rotate(float *a, float *b, const float *c, long long n, long long i0, long long stride) {
  const bool vec = (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0) && (n & 3) == 0;
  if (vec) {
    float4 *a4 = reinterpret_cast<float4 *>(a), *b4 = reinterpret_cast<float4 *>(b);
    const float4 *c4 = reinterpret_cast<const float4 *>(c);
    for (long long i = i0; i < (n >> 2); i += stride) {
      const float4 t = b4[i];
      a4[i] = t;
      if (c4 != nullptr) b4[i] = c4[i];
    }
  } else {
    for (long long i = i0; i < n; i += stride) {
      const float t = b[i];
      a[i] = t;
      if (c != nullptr) b[i] = c[i];
    }
  }
}

__global__ void __launch_bounds__(256)
stage_step_kernel(const __grid_constant__ StageParams p) {
  pdl_sync();
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const msr3d_stage_desc &d = p.d;
  for (int k = 0; k < d.n_rotate; ++k) stage_rotate(d.rot_a[k], d.rot_b[k], d.rot_c[k], d.rot_n[k], i0, stride);
  for (int k = 0; k < d.n_copy; ++k) stage_rotate(d.cp_dst[k], const_cast<float *>(d.cp_src[k]), nullptr, d.cp_n[k], i0, stride);
  for (long long i = i0; i < d.n_mask; i += stride) {
    const unsigned char t = d.mask_b[i];
    d.mask_a[i] = t;
    if (d.mask_c != nullptr) d.mask_b[i] = d.mask_c[i];
    if (d.key_pad != nullptr) d.key_pad[i] = t ? 0 : 1;
  }
}

// Flat variant (every segment 16 B aligned, length a multiple of 4 floats): ONE float4 per thread, so a thread pays one
// memory round trip instead of one per segment (the looping kernel above took 8-12 us for 1.4 MB: nine dependent
// load -> store round trips per thread).
__global__ void __launch_bounds__(256)
stage_step_flat_kernel(const __grid_constant__ StageParams p) {
  pdl_sync();
  const msr3d_stage_desc &d = p.d;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nseg = d.n_rotate + d.n_copy;
  if (i < p.start4[nseg]) {
    int k = 0;
    while (k + 1 < nseg && i >= p.start4[k + 1]) ++k;
    const long long j = i - p.start4[k];
    if (k < d.n_rotate) {
      float4 *a4 = reinterpret_cast<float4 *>(d.rot_a[k]), *b4 = reinterpret_cast<float4 *>(d.rot_b[k]);
      const float4 *c4 = reinterpret_cast<const float4 *>(d.rot_c[k]);
      const float4 t = b4[j];
      if (c4 != nullptr) b4[j] = c4[j];
      a4[j] = t;
    } else {
      const int c = k - d.n_rotate;
      reinterpret_cast<float4 *>(d.cp_dst[c])[j] = reinterpret_cast<const float4 *>(d.cp_src[c])[j];
    }
  }
  if (i < d.n_mask) {
    const unsigned char t = d.mask_b[i];
    d.mask_a[i] = t;
    if (d.mask_c != nullptr) d.mask_b[i] = d.mask_c[i];
    if (d.key_pad != nullptr) d.key_pad[i] = t ? 0 : 1;
  }
}

}  // namespace msr3d

using namespace msr3d;

/* see include/msr3d_b200.h */
MSR3D_API int msr3d_stage_step(const msr3d_stage_desc *d, msr3d_stream_t stream) {
  MSR3D_REQUIRE(d != nullptr && d->n_rotate >= 0 && d->n_rotate <= MSR3D_STAGE_MAX && d->n_copy >= 0 && d->n_copy <= MSR3D_STAGE_MAX,
                "stage_step: bad descriptor");
  long long total = d->n_mask;
  for (int k = 0; k < d->n_rotate; ++k) {
    MSR3D_REQUIRE(d->rot_a[k] && d->rot_b[k] && d->rot_n[k] >= 0, "stage_step: rotate %d", k);
    total += d->rot_n[k] / 4;
  }
  for (int k = 0; k < d->n_copy; ++k) {
    MSR3D_REQUIRE(d->cp_dst[k] && d->cp_src[k] && d->cp_n[k] >= 0, "stage_step: copy %d", k);
    total += d->cp_n[k] / 4;
  }
  MSR3D_REQUIRE(d->n_mask == 0 || (d->mask_a && d->mask_b), "stage_step: masks");
  if (total == 0) return MSR3D_OK;
  StageParams p;
  p.d = *d;
  // flat kernel when every segment can be moved as float4s
  bool flat = true;
  long long n4 = 0;
  auto ok = [](const void *q, long long n) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (n & 3) == 0; };
  for (int k = 0; k < d->n_rotate; ++k) {
    flat = flat && ok(d->rot_a[k], d->rot_n[k]) && ok(d->rot_b[k], 0) && ok(d->rot_c[k], 0);
    p.start4[k] = n4;
    n4 += d->rot_n[k] >> 2;
  }
  for (int k = 0; k < d->n_copy; ++k) {
    flat = flat && ok(d->cp_dst[k], d->cp_n[k]) && ok(d->cp_src[k], 0);
    p.start4[d->n_rotate + k] = n4;
    n4 += d->cp_n[k] >> 2;
  }
  p.start4[d->n_rotate + d->n_copy] = n4;
  if (flat) {
    const long long work = n4 > d->n_mask ? n4 : d->n_mask;
    launch_pdl(stage_step_flat_kernel, dim3((unsigned)((work + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, p);
    return check_launch("stage_step");
  }
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 2) grid = kNumSMs * 2;
  if (grid < 1) grid = 1;
  launch_pdl(stage_step_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p);
  return check_launch("stage_step");
}
