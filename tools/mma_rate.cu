// Microbenchmark (not part of the library): cycles per tcgen05.mma as a function of the issue pattern, the
// operand kind and N.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_rate tools/mma_rate.cu
// One CTA per SM, 128 threads; operands = whatever is in shared memory (zeros), SWIZZLE_128B K-major.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int KIND>   // 0 = tf32, 1 = bf16
__device__ __forceinline__ void mma(uint32_t tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// MODE 0: `if (lane == 0)` issue (divergent);  MODE 1: whole warp + elect.sync
template <int KIND, int N, int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long *out, int n_mma) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += 128) ((uint32_t *)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  const uint32_t fmt = KIND == 0 ? 2u : 1u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t da = desc_k(smem_u32(smem)), db = desc_k(smem_u32(smem + 128 * 128));
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 1) {
    if (MODE == 0) {
      if (lane == 0) {
        t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < n_mma; i += 12) {
#pragma unroll
          for (int j = 0; j < 12; ++j) mma<KIND>(tmem, da + 2 * (j & 3), db + 2 * (j & 3), idesc, 1u);
        }
        t1 = clock64();
        commit(&bar);
        while (!try_wait(&bar, 0)) {}
        t2 = clock64();
      }
    } else {
      t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < n_mma; i += 12) {
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 12; ++j) mma<KIND>(tmem, da + 2 * (j & 3), db + 2 * (j & 3), idesc, 1u);
        }
        __syncwarp();
      }
      t1 = clock64();
      if (elect_one()) commit(&bar);
      __syncwarp();
      while (!try_wait(&bar, 0)) {}
      t2 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

template <int KIND, int N, int MODE>
void run(const char *name, long long *d_out, int grid) {
  const int n_mma = 480;
  const size_t smem = (128 + 256) * 128 + 1024;
  cudaFuncSetAttribute(rate_kernel<KIND, N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) {
    rate_kernel<KIND, N, MODE><<<grid, 128, smem>>>(d_out, n_mma);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s grid %3d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA\n", name, grid, (double)h[0] / n_mma, (double)h[1] / n_mma);
}

int main() {
  long long *d_out;
  cudaMalloc(&d_out, 16);
  for (int grid : {1, 148}) {
    run<0, 64, 0>("tf32 N=64  lane0", d_out, grid);
    run<0, 64, 1>("tf32 N=64  elect", d_out, grid);
    run<0, 128, 0>("tf32 N=128 lane0", d_out, grid);
    run<0, 128, 1>("tf32 N=128 elect", d_out, grid);
    run<0, 256, 0>("tf32 N=256 lane0", d_out, grid);
    run<0, 256, 1>("tf32 N=256 elect", d_out, grid);
    run<1, 64, 0>("bf16 N=64  lane0", d_out, grid);
    run<1, 64, 1>("bf16 N=64  elect", d_out, grid);
    run<1, 128, 1>("bf16 N=128 elect", d_out, grid);
    run<1, 256, 1>("bf16 N=256 elect", d_out, grid);
  }
  return 0;
}
