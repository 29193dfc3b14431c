// Shared host/device helpers for libmsr3d_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <utility>

#include "../../include/msr3d_b200.h"

#define MSR3D_API extern "C" __attribute__((visibility("default")))

namespace msr3d {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

int set_error(int code, const char *fmt, ...);

// Call after every launch: converts a launch failure into a status code.
inline int check_launch(const char *what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MSR3D_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return MSR3D_OK;
}

#define MSR3D_REQUIRE(cond, ...)                                          \
  do {                                                                    \
    if (!(cond)) return msr3d::set_error(MSR3D_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define MSR3D_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess)                                                           \
      return msr3d::set_error(MSR3D_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// The step is a chain of ~60 small dependent kernels; with the ProgrammaticStreamSerialization
// launch attribute the next kernel's CTAs are scheduled (and run their set-up code) while the
// current kernel is still executing, and block in pdl_wait() until it has completed and flushed.
// Every kernel of this library calls pdl_wait() before its first global access, so the
// dependency is never weaker than stream order.  Toggle: msr3d_set_pdl() / env MSR3D_PDL=0.
extern int g_pdl;
// Tuning knobs (msr3d_set_tunable): 0 = library heuristics.  Used by tools/gemm_sweep.py to pick the
// heuristics' constants on the hardware; the product never changes them.
enum { TUNE_GEMM_BN = 0, TUNE_GEMM_KSPLIT = 1, TUNE_ATTN_FWD_WARPS = 2, TUNE_ATTN_BWD_WARPS = 3,
       TUNE_ATTN_FWD_ZSPLIT = 4, TUNE_ATTN_BWD_CLUSTER = 5, TUNE_SA_GROUP_FUSED = 6, TUNE_SA_SM_LIMIT = 7, TUNE_GRAD_ATOMICS = 8, TUNE_GEMM3X_DIRECT_EPI = 9, TUNE_ATTN_PACK = 10, TUNE_LINEAR_TMA_OFF = 11, TUNE_SA_OVERSUB = 12, TUNE_COUNT = 14 };
extern int g_tune[TUNE_COUNT];
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_launch_dependents(); }

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);  // errors surface in check_launch()
}
// same, with a thread-block cluster shape (grid must be a multiple of it)
template <typename... KArgs, typename... Args>
inline void launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                               dim3 cluster, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = cluster.x;
  attr[1].val.clusterDim.y = cluster.y;
  attr[1].val.clusterDim.z = cluster.z;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

#if defined(__CUDACC__)
// Optional per-CTA timeline (tools/trace_kernels.py): slot s of CTA c = SM clock at a phase boundary.  One pointer
// per translation unit (no relocatable device code), set through msr3d_set_trace_buffer; nullptr (default) = off.
constexpr int kTraceSlots = 16;
static __device__ long long *g_trace = nullptr;
__device__ __forceinline__ void trace_mark(int slot) {
  long long *t = g_trace;
  if (t != nullptr) {
    const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    t[(size_t)cta * kTraceSlots + slot] = clock64();
  }
}
#endif

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- squared distance with the contraction the reference binary uses ----
// nvcc 12.9 compiles `(ax-bx)*(ax-bx) + (ay-by)*(ay-by) + (az-bz)*(az-bz)` of the
// reference (ball_query_gpu.cu:31-32, sampling_gpu.cu:103-104, interpolate_gpu.cu:35)
// to  t = dy*dy; t = fma(dx,dx,t); t = fma(dz,dz,t).  Pinned with intrinsics so
// that this translation unit can never re-contract it differently.
__device__ __forceinline__ float sqdist3_ref(float ax, float ay, float az, float bx,
                                             float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  float t = __fmul_rn(dy, dy);
  t = __fmaf_rn(dx, dx, t);
  return __fmaf_rn(dz, dz, t);
}
__device__ __forceinline__ float sqnorm3_ref(float x, float y, float z) {
  float t = __fmul_rn(y, y);
  t = __fmaf_rn(x, x, t);
  return __fmaf_rn(z, z, t);
}

}  // namespace msr3d
