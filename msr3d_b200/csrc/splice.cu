// Group E: index side and backward gather of the scene-token splice into the LLM input
// (MSR3D.build_embeds, model/msr3d/msr3d.py:277-287).  The projection itself is the tensor-core GEMM
// of linear_tc.cu with a scattered fp16 / bf16 epilogue.
#include "common.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace msr3d {

constexpr int kScanThreads = 1024;

// One CTA walks the B*T ids in row-major order, kScanThreads at a time: ballot + popc inside a warp,
// a shared-memory scan over the warps, a running base across the passes.  B*T is a few thousand.
__global__ void __launch_bounds__(kScanThreads)
scene_rows_kernel(const long long *__restrict__ ids, int n, long long sp, const long long *__restrict__ am,
                  const unsigned char *__restrict__ obj_masks, int *__restrict__ rows, int max_rows,
                  int *__restrict__ count, unsigned char *__restrict__ mask_out) {
  pdl_sync();
  __shared__ int warp_tot[kScanThreads / 32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += kScanThreads) {
    const int i = i0 + tid;
    const bool hit = i < n && ids[i] == sp;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const int before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
    const int base = s_base;
    const int k = base + wbase + before;
    if (i < n && mask_out != nullptr && am != nullptr) {
      unsigned char m = am[i] != 0;
      if (hit && obj_masks != nullptr && k < max_rows) m = obj_masks[k] != 0;
      mask_out[i] = m;
    }
    if (hit && k < max_rows) rows[k] = i;
    __syncthreads();
    if (tid == kScanThreads - 1) s_base = k + (hit ? 1 : 0);
    __syncthreads();
  }
  const int total = s_base;
  for (int k = total + tid; k < max_rows; k += kScanThreads) rows[k] = -1;
  if (tid == 0) *count = total;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// one warp per (row, 256-column block): coalesced reads of the source row, coalesced fp32 writes
template <typename T>
__global__ void __launch_bounds__(256)
splice_gather_kernel(const T *__restrict__ src, long long ld, const int *__restrict__ rows, int M, int D,
                     float *__restrict__ dY) {
  pdl_sync();
  const int m = blockIdx.y;
  const int r = rows[m];
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < D; c += gridDim.x * blockDim.x)
    dY[(size_t)m * D + c] = r >= 0 ? to_f32<T>(src[(long long)r * ld + c]) : 0.f;
}

}  // namespace msr3d

using namespace msr3d;

MSR3D_API int msr3d_scene_rows_i64(const int64_t *input_ids, int B, int T, long long sp_token,
                                   const int64_t *attention_mask, const uint8_t *obj_masks, int32_t *rows,
                                   int max_rows, int32_t *count, uint8_t *mask_out, msr3d_stream_t stream) {
  MSR3D_REQUIRE(input_ids && rows && count, "scene_rows: null pointer");
  MSR3D_REQUIRE(B >= 0 && T >= 0 && max_rows >= 0, "scene_rows: bad sizes B=%d T=%d max_rows=%d", B, T, max_rows);
  MSR3D_REQUIRE((long long)B * T < (1ll << 31), "scene_rows: B*T too large");
  launch_pdl(scene_rows_kernel, dim3(1), dim3(kScanThreads), 0, (cudaStream_t)stream,
             (const long long *)input_ids, B * T, sp_token, (const long long *)attention_mask, obj_masks, rows,
             max_rows, count, mask_out);
  return check_launch("scene_rows");
}

MSR3D_API int msr3d_splice_gather_grad(const void *d_embeds, int dtype, long long ld, const int32_t *rows, int M,
                                       int D, float *dY, msr3d_stream_t stream) {
  MSR3D_REQUIRE(d_embeds && rows && dY, "splice_gather_grad: null pointer");
  MSR3D_REQUIRE(M >= 0 && D > 0 && dtype >= 0 && dtype <= 2, "splice_gather_grad: bad args M=%d D=%d dtype=%d", M, D, dtype);
  if (M == 0) return MSR3D_OK;
  dim3 grid(ceil_div(D, 256 * 4) > 0 ? ceil_div(D, 256 * 4) : 1, M);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 0) launch_pdl(splice_gather_kernel<float>, grid, dim3(256), 0, st, (const float *)d_embeds, ld, rows, M, D, dY);
  else if (dtype == 1) launch_pdl(splice_gather_kernel<__half>, grid, dim3(256), 0, st, (const __half *)d_embeds, ld, rows, M, D, dY);
  else launch_pdl(splice_gather_kernel<__nv_bfloat16>, grid, dim3(256), 0, st, (const __nv_bfloat16 *)d_embeds, ld, rows, M, D, dY);
  return check_launch("splice_gather_grad");
}
