// Gradient all-reduce of the data-parallel step as ONE kernel per bucket over NVLink peer memory (two-shot:
// reduce-scatter + all-gather by direct peer loads / stores), with the squared-norm partials of the clip fused in.
//
// The reference wraps the trainable prompter in DDP (trainer/leo_trainer.py:50-52,135) and clips the global norm
// afterwards (:192-193).  Round 1 / 2 called ncclAllReduce on the flat gradient buffer: for these sizes (5.8 MB and
// 11 MB) NCCL picks RING_LL and needs 41 us and 59 us on two GPUs (tools/step_timeline.py under torchrun) — the wire
// time is 6 us.  Here every rank's gradient buffer lives in symmetric memory (torch.distributed._symmetric_memory:
// the same allocation mapped into every peer), and
//     rank r owns slice r of the bucket:   s[i] = sum_p grad_p[i]  (peer loads, FIXED order p = 0..W-1: every rank
//                                          receives bit-identical values, replicas stay bit-identical)
//                                          grad_p[i] <- s[i] for every p          (peer stores)
//                                          partial[r][block] = sum_i s[i]^2       (the clip's squared norm, pushed to
//                                                                                  every peer with the data)
// bracketed by two cross-GPU barriers on flags in symmetric memory, PER BLOCK (block b of rank r pairs with block b of
// every peer: no grid-wide synchronisation, no co-residency requirement).  Flags carry a monotonically increasing
// epoch kept in device memory, so the launch is CUDA-graph replayable.  msr3d_grad_norm_finalize_f32 then adds the
// W x blocks partials of both buckets in a fixed order (what sumsq_kernel's pass over all gradients did).
#include "common.cuh"

namespace msr3d {

constexpr int kArMaxWorld = MSR3D_AR_MAX_WORLD;
constexpr int kArMaxBlocks = MSR3D_AR_MAX_BLOCKS;
constexpr int kArThreads = 512;

struct ArParams {
  msr3d_allreduce_desc d;
  long long begin, end;
};

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float *p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// every peer's block b has written `epoch` into my flag (p, phase); bounded: a lost peer traps instead of hanging the GPU
__device__ __forceinline__ void wait_flag(const unsigned *flag, unsigned epoch) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned spins = 0;
  while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
    if ((++spins & 4095u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 5000000000ull) __trap();
    }
  }
}

__global__ void __launch_bounds__(kArThreads)
allreduce_p2p_kernel(const __grid_constant__ ArParams p) {
  pdl_sync();
  const msr3d_allreduce_desc &d = p.d;
  if (d.gate != nullptr && *d.gate == 0) return;   // gradient accumulation: every rank skips the same micro-steps
  const int W = d.world, r = d.rank, b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
  __shared__ unsigned s_epoch;
  __shared__ float s_red[kArThreads / 32];
  unsigned *ep = d.epoch + d.channel * kArMaxBlocks + b;
  if (tid == 0) s_epoch = *ep + 1u;
  __syncthreads();
  const unsigned epoch = s_epoch;
  const size_t fbase = ((size_t)(d.channel * kArMaxBlocks + b) * kArMaxWorld) * 2;   // flags of (channel, block): [peer][phase]
  // ---- barrier A: every peer has reached this kernel, i.e. its gradients of this bucket are final
  if (tid < W) st_release_sys(d.flags[tid] + fbase + (size_t)r * 2 + 0, epoch);
  if (tid < W) wait_flag(d.flags[r] + fbase + (size_t)tid * 2 + 0, epoch);
  __syncthreads();
  // ---- slice of this rank, chunk of this block
  const long long n4 = (p.end - p.begin) >> 2;                 // float4s in the bucket
  const long long per = (n4 + W - 1) / W;
  const long long s0 = (long long)r * per, s1 = s0 + per < n4 ? s0 + per : n4;
  float ss = 0.f;
  // kArUnroll float4s per thread in flight from every peer (a peer load is a ~2 us NVLink round trip: one per
  // iteration made the 11 MB bucket 40 us on two GPUs)
  constexpr int kArUnroll = 4;
  const long long stride = (long long)nb * kArThreads;
  for (long long i0 = s0 + (long long)b * kArThreads + tid; i0 < s1; i0 += stride * kArUnroll) {
    float4 acc[kArUnroll];
#pragma unroll
    for (int u = 0; u < kArUnroll; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < kArMaxWorld; ++q) {
      if (q < W) {
        float4 v[kArUnroll];
#pragma unroll
        for (int u = 0; u < kArUnroll; ++u) {
          const long long i = i0 + u * stride;
          v[u] = i < s1 ? ld_peer_f4(d.buf[q] + p.begin + (i << 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kArUnroll; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
      }
    }
#pragma unroll
    for (int u = 0; u < kArUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < s1) {
        ss += (acc[u].x * acc[u].x + acc[u].y * acc[u].y) + (acc[u].z * acc[u].z + acc[u].w * acc[u].w);
#pragma unroll
        for (int q = 0; q < kArMaxWorld; ++q)
          if (q < W) *reinterpret_cast<float4 *>(d.buf[q] + p.begin + (i << 2)) = acc[u];
      }
    }
  }
  // squared norm of this block's part of the reduced bucket -> every peer (fixed summation order inside the block)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < kArThreads / 32; ++i) t += s_red[i];
    for (int q = 0; q < W; ++q) d.partials[q][((size_t)d.channel * kArMaxWorld + r) * kArMaxBlocks + b] = t;
  }
  __threadfence_system();
  __syncthreads();
  // ---- barrier B: every peer's block b has pushed its slice
  if (tid < W) st_release_sys(d.flags[tid] + fbase + (size_t)r * 2 + 1, epoch);
  if (tid < W) wait_flag(d.flags[r] + fbase + (size_t)tid * 2 + 1, epoch);
  __syncthreads();
  if (tid == 0) *ep = epoch;
}

struct OptStateHead { float sumsq, step; };

// sum of the W x blocks partials of `nch` buckets in a fixed order -> the clip's squared norm (+ step / seed counters:
// what sumsq_kernel's last block does on the single-GPU path).  grad_scale: the 1/world factor of the mean.
__global__ void __launch_bounds__(256)
grad_norm_finalize_kernel(const float *__restrict__ partials, int world, int nb0, int nb1, float grad_scale,
                          const float *__restrict__ extra_sumsq, float *state, unsigned long long *seed_counter,
                          const int *__restrict__ gate) {
  pdl_sync();
  if (gate != nullptr && *gate == 0) return;
  __shared__ float w[8];
  float t = 0.f;
  const int nbs[2] = {nb0, nb1};
  for (int ch = 0; ch < 2; ++ch)
    for (int i = threadIdx.x; i < world * nbs[ch]; i += 256) {
      const int q = i / nbs[ch], b = i - q * nbs[ch];
      t += partials[((size_t)ch * kArMaxWorld + q) * kArMaxBlocks + b];
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += w[i];
    tot *= grad_scale * grad_scale;
    if (extra_sumsq != nullptr) tot += *extra_sumsq;
    OptStateHead *st = reinterpret_cast<OptStateHead *>(state);
    st->sumsq = tot;
    st->step += 1.f;
    if (seed_counter != nullptr) *seed_counter += 1ull;
  }
}

// Squared norm of g[begin, end) as `nblocks` block partials in the layout grad_norm_finalize_kernel reads (rank 0 of
// channel `channel`): the single-GPU counterpart of the partials the exchange kernel pushes along — lets the norm of the
// early gradient bucket be taken on a side stream under the backward of layer 0, so that only the late bucket is left for
// the tail of the step.  Deterministic (fixed order inside a block, fixed block -> slot mapping).
__global__ void __launch_bounds__(256)
sumsq_partials_kernel(const float *__restrict__ g, long long begin, long long end, float *__restrict__ partials,
                      int channel) {
  pdl_sync();
  __shared__ float w[8];
  const long long n4 = (end - begin) >> 2;
  const float4 *g4 = reinterpret_cast<const float4 *>(g + begin);
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 q = g4[i];
    s += (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += w[i];
    partials[((size_t)channel * kArMaxWorld + 0) * kArMaxBlocks + blockIdx.x] = t;
  }
}

}  // namespace msr3d

using namespace msr3d;

/* see include/msr3d_b200.h */
MSR3D_API int msr3d_allreduce_sum_f32(const msr3d_allreduce_desc *d, long long begin, long long end, int nblocks,
                                      msr3d_stream_t stream) {
  MSR3D_REQUIRE(d != nullptr && d->world >= 2 && d->world <= kArMaxWorld && d->rank >= 0 && d->rank < d->world,
                "allreduce: bad world / rank");
  MSR3D_REQUIRE(d->channel >= 0 && d->channel < 2 && nblocks >= 1 && nblocks <= kArMaxBlocks, "allreduce: channel / blocks");
  MSR3D_REQUIRE(begin >= 0 && end >= begin && (begin & 3) == 0 && ((end - begin) & 3) == 0, "allreduce: range must be float4 aligned");
  for (int q = 0; q < d->world; ++q)
    MSR3D_REQUIRE(d->buf[q] && d->flags[q] && d->partials[q] && (reinterpret_cast<uintptr_t>(d->buf[q]) & 15) == 0, "allreduce: peer %d", q);
  MSR3D_REQUIRE(d->epoch != nullptr, "allreduce: epoch counters");
  if (end == begin) return MSR3D_OK;
  ArParams p;
  p.d = *d;
  p.begin = begin; p.end = end;
  launch_pdl(allreduce_p2p_kernel, dim3(nblocks), dim3(kArThreads), 0, (cudaStream_t)stream, p);
  return check_launch("allreduce_p2p");
}

MSR3D_API int msr3d_grad_norm_finalize_f32(const float *partials, int world, int nblocks0, int nblocks1, float grad_scale,
                                           const float *extra_sumsq, float *state, unsigned long long *seed_counter,
                                           const int *gate, msr3d_stream_t stream) {
  MSR3D_REQUIRE(partials && state && world >= 1 && world <= kArMaxWorld && nblocks0 >= 0 && nblocks0 <= kArMaxBlocks &&
                    nblocks1 >= 0 && nblocks1 <= kArMaxBlocks, "grad_norm_finalize: bad arguments");
  launch_pdl(grad_norm_finalize_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, partials, world, nblocks0, nblocks1, grad_scale,
             extra_sumsq, state, seed_counter, gate);
  return check_launch("grad_norm_finalize");
}

MSR3D_API int msr3d_sumsq_partials_f32(const float *g, long long begin, long long end, int nblocks, float *partials,
                                       int channel, msr3d_stream_t stream) {
  MSR3D_REQUIRE(g && partials && channel >= 0 && channel < 2 && nblocks >= 1 && nblocks <= kArMaxBlocks,
                "sumsq_partials: bad arguments");
  MSR3D_REQUIRE(begin >= 0 && end >= begin && (begin & 3) == 0 && ((end - begin) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(g) & 15) == 0, "sumsq_partials: range must be float4 aligned");
  launch_pdl(sumsq_partials_kernel, dim3(nblocks), dim3(256), 0, (cudaStream_t)stream, g, begin, end, partials, channel);
  return check_launch("sumsq_partials");
}
