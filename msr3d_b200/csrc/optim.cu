// Group D: the optimiser step around the collective, on ONE flat fp32 buffer.
// Replaces accelerator.clip_grad_norm_(5.0) + torch.optim.AdamW.step()
// (trainer/leo_trainer.py:192-195, optim/build.py:7-17) — per-tensor launches in the
// reference — with two launches: a sum-of-squares reduction and a fused clip+AdamW update.
// Hyper-parameters live in a small device struct so the whole training step can be replayed
// as a CUDA graph while the host changes the learning rate between replays.
#include "common.cuh"

namespace msr3d {

constexpr int kSumsqBlocks = kNumSMs * 4;
struct OptState {   // device-resident: 8 floats + one partial sum per block of sumsq_kernel
  float sumsq;      // written by sumsq_kernel's last block
  float step;       // incremented once per step by sumsq_kernel's last block
  float lr, beta1, beta2, eps, weight_decay, max_norm;
  float partial[kSumsqBlocks];
};

__global__ void __launch_bounds__(256)
sumsq_kernel(const float *__restrict__ g, long long n, float grad_scale, OptState *st,
             unsigned int *ticket, unsigned long long *seed_counter, const float *__restrict__ extra_sumsq,
             const int *__restrict__ gate) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  if (gate != nullptr && *gate == 0) return;   // gradient accumulation: not a synchronisation step (uniform over the grid)
  float s = 0.f;
  const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {   // 16 B loads; the tail (n % 4) goes to the scalar loop below
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    const long long n4 = n >> 2;
    for (long long i = tid0; i < n4; i += nthr) {
      const float4 q = g4[i];
      const float a = q.x * grad_scale, b = q.y * grad_scale, c = q.z * grad_scale, d = q.w * grad_scale;
      s += (a * a + b * b) + (c * c + d * d);
    }
    for (long long i = (n4 << 2) + tid0; i < n; i += nthr) {
      const float v = g[i] * grad_scale;
      s += v * v;
    }
  } else {
    for (long long i = tid0; i < n; i += nthr) {
      const float v = g[i] * grad_scale;
      s += v * v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float w[8];
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
  __syncthreads();
  // Deterministic: block partials are stored, and the last block to finish adds them up in a fixed
  // order.  (An atomicAdd per block made the norm — hence the clip factor, hence the update — depend on
  // the arrival order: data-parallel replicas that hold the same reduced gradient drifted apart in the
  // last bits, tools/check_ddp_equiv.py.)
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += w[i];
    st->partial[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) t += ((volatile float *)st->partial)[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < 8; ++i) tot += w[i];
      // the reference clips by the norm over ALL trainable parameters of the model
      // (accelerator.clip_grad_norm_(self.model.parameters(), ...), trainer/leo_trainer.py:192-193): the caller
      // hands in the squared norm of whatever it optimises itself (llm_proj, LoRA, ...)
      if (extra_sumsq != nullptr) tot += *extra_sumsq;
      st->sumsq = tot;
      st->step += 1.f;
      if (seed_counter != nullptr) *seed_counter += 1ull;  // next step draws fresh dropout masks
      *ticket = 0u;
    }
  }
}

// p -= lr * ( m_hat / (sqrt(v_hat) + eps) + wd * p ), grads first scaled by grad_scale (1/world)
// and by the global-norm clip coefficient min(1, max_norm / (norm + 1e-6)).
// Elements [decay_begin, decay_end) get weight decay, the rest (biases, LayerNorm) do not.
__global__ void __launch_bounds__(256)
adamw_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m,
             float *__restrict__ v, long long n, long long decay_begin, long long decay_end, float grad_scale,
             OptState *st, int zero_grads, float *__restrict__ big, float *__restrict__ small, const int *__restrict__ gate) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  if (gate != nullptr && *gate == 0) return;
  const float norm = sqrtf(st->sumsq);
  const float clip = st->max_norm > 0.f ? fminf(1.f, st->max_norm / (norm + 1e-6f)) : 1.f;
  const float gs = grad_scale * clip;
  const float b1 = st->beta1, b2 = st->beta2, lr = st->lr, eps = st->eps, wd = st->weight_decay;
  const float bc1 = 1.f - powf(b1, st->step), bc2 = 1.f - powf(b2, st->step);
  const float step_size = lr / bc1;
  const float inv_bc2_sqrt = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    if (zero_grads) g[i] = 0.f;   // the next step accumulates into a clean buffer: no separate 17 MB memset in front of it
    float pi = p[i];
    if (i >= decay_begin && i < decay_end) pi *= 1.f - lr * wd;  // decoupled weight decay (torch AdamW order)
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    const float pn = pi - step_size * (mi / denom);
    p[i] = pn;
    if (big != nullptr) {
      // tf32 operand images of the updated parameter for the next step's 3xTF32 GEMMs (tma.cuh: big = x rounded to
      // tf32, small = x - big): written here so that no separate pass over the parameters sits in front of the step
      const float b = __uint_as_float((__float_as_uint(pn) + 0x1000u) & 0xffffe000u);
      big[i] = b;
      small[i] = pn - b;
    }
  }
}

// out = <a, b>: the surrogate loss <obj_tokens, upstream> read back by the training loop.  One CTA,
// fixed summation order (deterministic); n is a few 10^4.
__global__ void __launch_bounds__(1024)
dot_kernel(const float *__restrict__ a, const float *__restrict__ b, long long n, float *__restrict__ out) {
  pdl_sync();
  float s = 0.f;
  if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) {
    const float4 *a4 = reinterpret_cast<const float4 *>(a), *b4 = reinterpret_cast<const float4 *>(b);
    const long long n4 = n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 1024) {
      const float4 x = a4[i], y = b4[i];
      s += (x.x * y.x + x.y * y.y) + (x.z * y.z + x.w * y.w);
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += 1024) s += a[i] * b[i];
  } else {
    for (long long i = threadIdx.x; i < n; i += 1024) s += a[i] * b[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float w[32];
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = w[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) *out = t;
  }
}

}  // namespace msr3d

using namespace msr3d;

/* state: MSR3D_OPT_STATE_FLOATS floats on the device {sumsq, step, lr, beta1, beta2, eps, weight_decay, max_norm,
 * then one scratch float per block of the norm reduction};
 * ticket: one zero-initialised uint32 on the device. */
static int adamw_impl(float *params, float *grads, float *exp_avg, float *exp_avg_sq, long long n, long long decay_begin,
                      long long decay_end, float grad_scale, float *state, unsigned int *ticket,
                      unsigned long long *seed_counter, const float *extra_sumsq, int zero_grads, float *big, float *small,
                      msr3d_stream_t stream, const int *gate = nullptr) {
  MSR3D_REQUIRE(params && grads && exp_avg && exp_avg_sq && state && ticket, "adamw: null pointer");
  MSR3D_REQUIRE(n >= 0 && decay_begin >= 0 && decay_begin <= decay_end && decay_end <= n, "adamw: bad sizes");
  if (n == 0) return MSR3D_OK;
  cudaStream_t s = (cudaStream_t)stream;
  OptState *st = reinterpret_cast<OptState *>(state);
  int grid = (int)((n + 255) / 256);
  if (grid > kSumsqBlocks) grid = kSumsqBlocks;
  int rc;
  launch_pdl(sumsq_kernel, dim3(grid), dim3(256), 0, s, grads, n, grad_scale, st, ticket, seed_counter, extra_sumsq, gate);
  rc = check_launch("sumsq");
  if (rc) return rc;
  launch_pdl(adamw_kernel, dim3(grid), dim3(256), 0, s, params, grads, exp_avg, exp_avg_sq, n, decay_begin, decay_end, grad_scale, st,
             zero_grads, big, small, gate);
  return check_launch("adamw");
}

MSR3D_API int msr3d_adamw_clip_step_f32(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                                        long long decay_begin, long long decay_end, float grad_scale, float *state,
                                        unsigned int *ticket, unsigned long long *seed_counter, const float *extra_sumsq,
                                        msr3d_stream_t stream) {
  return adamw_impl(params, const_cast<float *>(grads), exp_avg, exp_avg_sq, n, decay_begin, decay_end, grad_scale, state, ticket,
                    seed_counter, extra_sumsq, 0, nullptr, nullptr, stream);
}

/* Same, and the gradient buffer is zeroed as it is consumed (zero_grads != 0): the next step's backward accumulates
 * into it without a memset in front of the step. */
MSR3D_API int msr3d_adamw_clip_step_zg_f32(float *params, float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                                           long long decay_begin, long long decay_end, float grad_scale, float *state,
                                           unsigned int *ticket, unsigned long long *seed_counter,
                                           const float *extra_sumsq, int zero_grads, msr3d_stream_t stream) {
  return adamw_impl(params, grads, exp_avg, exp_avg_sq, n, decay_begin, decay_end, grad_scale, state, ticket, seed_counter,
                    extra_sumsq, zero_grads, nullptr, nullptr, stream);
}

/* Same as msr3d_adamw_clip_step_f32; additionally writes the tf32 big / small operand images of every updated parameter
 * (what msr3d_split_tf32_f32 would produce) into big[0..n) / small[0..n). */
MSR3D_API int msr3d_adamw_clip_step_split_f32(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                                              long long decay_begin, long long decay_end, float grad_scale, float *state,
                                              unsigned int *ticket, unsigned long long *seed_counter,
                                              const float *extra_sumsq, float *big, float *small, msr3d_stream_t stream) {
  MSR3D_REQUIRE(big && small, "adamw_split: null image pointer");
  return adamw_impl(params, const_cast<float *>(grads), exp_avg, exp_avg_sq, n, decay_begin, decay_end, grad_scale, state, ticket,
                    seed_counter, extra_sumsq, 0, big, small, stream);
}

MSR3D_API int msr3d_adamw_apply_f32(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                                    long long decay_begin, long long decay_end, float grad_scale, float *state, float *big,
                                    float *small, const int *gate, msr3d_stream_t stream) {
  MSR3D_REQUIRE(params && grads && exp_avg && exp_avg_sq && state, "adamw_apply: null pointer");
  MSR3D_REQUIRE(n >= 0 && decay_begin >= 0 && decay_begin <= decay_end && decay_end <= n, "adamw_apply: bad sizes");
  MSR3D_REQUIRE((big == nullptr) == (small == nullptr), "adamw_apply: both images or none");
  if (n == 0) return MSR3D_OK;
  int grid = (int)((n + 255) / 256);
  if (grid > kSumsqBlocks) grid = kSumsqBlocks;
  launch_pdl(adamw_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, params, const_cast<float *>(grads), exp_avg, exp_avg_sq, n,
             decay_begin, decay_end, grad_scale, reinterpret_cast<OptState *>(state), 0, big, small, gate);
  return check_launch("adamw");
}

namespace msr3d {
// ---- gradient accumulation (configs/msr3d.yaml:33 gradient_accumulation_steps, trainer/leo_trainer.py:66,211) inside a
// replayable step: a device-side counter decides whether this micro-step clears / exchanges / applies, so ONE CUDA graph
// serves every micro-step.  state = {micro, k, clear_next, sync_now}.
__global__ void accum_tick_kernel(int *st, unsigned long long *seed_counter, int phase) {
  pdl_sync();
  if (threadIdx.x == 0) {
    const int k = st[1] > 0 ? st[1] : 1;
    if (phase == 0) {               // start of a micro-step: what this one does
      st[2] = st[0] == 0 ? 1 : 0;       // first of the window: gradients are cleared
      st[3] = st[0] == k - 1 ? 1 : 0;   // last of the window: exchange + clip + AdamW run
    } else {                        // end of a micro-step
      st[0] = (st[0] + 1) % k;
      if (seed_counter != nullptr) *seed_counter += 1ull;   // fresh dropout masks every micro-step
    }
  }
}
__global__ void __launch_bounds__(256)
zero_gated_kernel(float4 *__restrict__ p, long long n4, const int *__restrict__ gate) {
  pdl_sync();
  if (gate != nullptr && *gate == 0) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace msr3d
using namespace msr3d;

MSR3D_API int msr3d_accum_tick(int *state, unsigned long long *seed_counter, int phase, msr3d_stream_t stream) {
  MSR3D_REQUIRE(state != nullptr && (phase == 0 || phase == 1), "accum_tick: bad arguments");
  launch_pdl(accum_tick_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, state, seed_counter, phase);
  return check_launch("accum_tick");
}
MSR3D_API int msr3d_zero_gated_f32(float *p, long long n, const int *gate, msr3d_stream_t stream) {
  MSR3D_REQUIRE(p != nullptr && n >= 0 && (n & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0, "zero_gated: alignment");
  if (n == 0) return MSR3D_OK;
  launch_pdl(zero_gated_kernel, dim3(kNumSMs * 2), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<float4 *>(p), n >> 2, gate);
  return check_launch("zero_gated");
}
/* msr3d_adamw_clip_step_split_f32 gated by *gate (gradient accumulation; big / small nullable). */
MSR3D_API int msr3d_adamw_clip_step_gated_f32(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                                              long long decay_begin, long long decay_end, float grad_scale, float *state,
                                              unsigned int *ticket, const float *extra_sumsq, float *big, float *small,
                                              const int *gate, msr3d_stream_t stream) {
  return adamw_impl(params, const_cast<float *>(grads), exp_avg, exp_avg_sq, n, decay_begin, decay_end, grad_scale, state, ticket,
                    nullptr, extra_sumsq, 0, big, small, stream, gate);
}


/* out[0] = sum_i a[i] * b[i] (device scalar), deterministic. */
MSR3D_API int msr3d_dot_f32(const float *a, const float *b, long long n, float *out, msr3d_stream_t stream) {
  MSR3D_REQUIRE(a && b && out && n >= 0, "dot: bad args");
  launch_pdl(dot_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, a, b, n, out);
  return check_launch("dot");
}
