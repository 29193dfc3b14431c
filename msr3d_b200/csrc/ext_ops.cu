// Group A of include/msr3d_b200.h: the nine `pointnet2._ext` operators as
// sm_100a kernels.  HBM-bound integer/index work: no tensor cores here; the
// rules that matter are coalescing, keeping the cloud on chip and not
// serialising a cloud onto one thread (the reference scans 1024 points per
// thread in ball query, ball_query_gpu.cu:23-43, and round-trips a `temp`
// array through global memory 31 times in FPS, sampling_gpu.cu:89-172).
//
// Bit-exactness contract: int32 outputs equal the reference kernels' outputs.
// That requires (1) the same float contraction for squared distances
// (common.cuh: sqdist3_ref), (2) the same strictness of every comparison and
// (3) for FPS the reference's thread-major / tree-order tie-break, restated
// here as a total order ("rank") on point indices — see fps_rank().
#include "common.cuh"

namespace msr3d {

// ----------------------------------------------------------------------------
// Furthest point sampling
// ----------------------------------------------------------------------------
// Reference semantics (sampling_gpu.cu:69-173), block_size bs = opt_n_threads(N):
//   thread t scans k = t, t+bs, ... keeping the first strict maximum of
//   temp[k] = min(temp[k], |p_k - p_old|^2) over points with |p_k|^2 > 1e-3;
//   a shared-memory tree (strides bs/2..1) merges threads, a tie keeping the
//   lower slot.  Lower slot at stride s differs from the upper one only in bit
//   log2(s) of the thread id, and lower strides are merged later, so among
//   equal maxima the winner is the thread with the smallest BIT-REVERSED id,
//   then the smallest k inside that thread.  rank(k) encodes exactly that
//   order; arg-max of (value, -rank) reproduces the reference pick for pick.
// Only the ORDER of ranks matters, so the rank is (bit-reversed thread id) * 16 + (k / bs): k / bs < 16
// for every supported size (N <= 8192 with bs = 512, and N < 2 * bs below that), and unranking is a
// shift and a mask instead of an integer division in every thread every round.
constexpr int kFpsRankShift = 4;
__device__ __forceinline__ uint32_t fps_rank(int k, int log2bs) {
  const uint32_t t = (uint32_t)k & ((1u << log2bs) - 1u);
  const uint32_t rev = log2bs ? (__brev(t) >> (32 - log2bs)) : 0u;
  return (rev << kFpsRankShift) | ((uint32_t)k >> log2bs);
}
__device__ __forceinline__ int fps_unrank(uint32_t r, int log2bs) {
  const uint32_t rev = r >> kFpsRankShift, i = r & ((1u << kFpsRankShift) - 1u);
  const uint32_t t = log2bs ? (__brev(rev) >> (32 - log2bs)) : 0u;
  return (int)(t + (i << log2bs));
}

// One CTA per cloud.  The cloud lives in shared memory (for the p_old
// broadcast) and in registers (PPT points per thread, with their running
// min-distance), so a round touches no global memory at all.
// Per round: PPT distance updates, two REDUX per warp, ONE __syncthreads
// (double-buffered partials), a redundant NW-way merge in every thread.
template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS)
fps_kernel(const float *__restrict__ xyz, int ld, int32_t *__restrict__ idx,
           float *__restrict__ new_xyz, int N, int m, int log2bs, int nper,
           int32_t *__restrict__ ball_idx, float radius2, int nsample) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  extern __shared__ float s_xyz[];  // N*3 floats, then m int32 picks
  int32_t *s_pick = reinterpret_cast<int32_t *>(s_xyz + (size_t)N * 3);
  constexpr int NW = THREADS / 32;
  __shared__ uint2 s_part[2][NW];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // `ld` floats per point (3 for a packed xyz tensor, 6 when xyz is the head of an
  // (x,y,z,r,g,b) record: the break_up_pc copy of modules/layers/pointnet.py:6-19 is skipped)
  const float *cloud = xyz + (size_t)blockIdx.x * N * ld;
  if (ld == 3) {
    for (int i = tid; i < N * 3; i += THREADS) s_xyz[i] = cloud[i];
  } else {
    for (int i = tid; i < N; i += THREADS) {
      s_xyz[i * 3 + 0] = cloud[(size_t)i * ld + 0];
      s_xyz[i * 3 + 1] = cloud[(size_t)i * ld + 1];
      s_xyz[i * 3 + 2] = cloud[(size_t)i * ld + 2];
    }
  }
  __syncthreads();

  float px[PPT], py[PPT], pz[PPT], dmin[PPT];
  uint32_t rk[PPT];
  uint32_t valid = 0;
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = tid + i * THREADS;
    px[i] = py[i] = pz[i] = 0.f;
    dmin[i] = 1e10f;  // sampling.cpp:74-76
    rk[i] = 0xffffffffu;
    if (k < N) {
      px[i] = s_xyz[k * 3 + 0];
      py[i] = s_xyz[k * 3 + 1];
      pz[i] = s_xyz[k * 3 + 2];
      const float mag = sqnorm3_ref(px[i], py[i], pz[i]);
      // `mag <= 1e-3` is a double comparison in the reference (DSETP.GTU).
      if (!((double)mag <= 1e-3)) {
        valid |= 1u << i;
        rk[i] = fps_rank(k, log2bs);
      }
    }
  }

  int old = 0;
  if (tid == 0 && m > 0) s_pick[0] = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = s_xyz[old * 3 + 0], y1 = s_xyz[old * 3 + 1], z1 = s_xyz[old * 3 + 2];
    // key = bits(d2)+1 for a candidate (d2 >= +0, so unsigned order == float
    // order), 0 for "this thread has none" (reference: best = -1, besti = 0).
    uint32_t bkey = 0, brank = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      if (valid & (1u << i)) {
        const float d = sqdist3_ref(px[i], py[i], pz[i], x1, y1, z1);
        const float d2 = fminf(d, dmin[i]);
        dmin[i] = d2;
        const uint32_t key = __float_as_uint(d2) + 1u;
        if (key > bkey || (key == bkey && rk[i] < brank)) {
          bkey = key;
          brank = rk[i];
        }
      }
    }
    const uint32_t wkey = __reduce_max_sync(0xffffffffu, bkey);
    const uint32_t wrank = __reduce_min_sync(0xffffffffu, bkey == wkey ? brank : 0xffffffffu);
    if (NW > 1) {
      if (lane == 0) s_part[j & 1][warp] = make_uint2(wkey, wrank);
      __syncthreads();
      // cross-warp merge: lane w holds warp w's partial, two more REDUX (every warp redundantly, so no
      // second barrier); the serial NW-way compare chain in every thread was a third of the round
      const uint2 p = lane < NW ? s_part[j & 1][lane] : make_uint2(0u, 0xffffffffu);
      const uint32_t gkey = __reduce_max_sync(0xffffffffu, p.x);
      const uint32_t grank = __reduce_min_sync(0xffffffffu, p.x == gkey ? p.y : 0xffffffffu);
      old = gkey ? fps_unrank(grank, log2bs) : 0;
    } else {
      old = wkey ? fps_unrank(wrank, log2bs) : 0;
    }
    if (tid == 0) s_pick[j] = old;
  }
  __syncthreads();
  for (int j = tid; j < m; j += THREADS) idx[(size_t)blockIdx.x * m + j] = s_pick[j];
  if (new_xyz != nullptr) {
    // fused gather (pointnet2_modules.py:54-58): new_xyz[b,j,:] = xyz[b,idx[b,j],:]
    for (int i = tid; i < m * 3; i += THREADS)
      new_xyz[(size_t)blockIdx.x * m * 3 + i] = s_xyz[s_pick[i / 3] * 3 + i % 3];
  }
  if (ball_idx != nullptr) {
    // fused ball query (msr3d_sa_group_f32): the cloud is already in shared memory and the centres
    // are the picks, so the second launch and its re-read of the cloud are gone.  Same scan as
    // ball_query_kernel below — a warp per centre, ascending index order, strict `<`, pad with the
    // first hit, zeros for an empty ball.
    for (int j = warp; j < m; j += NW) {
      const int c = s_pick[j];
      const float qx = s_xyz[c * 3 + 0], qy = s_xyz[c * 3 + 1], qz = s_xyz[c * 3 + 2];
      int32_t *row = ball_idx + ((size_t)blockIdx.x * m + j) * nsample;
      int cnt = 0, first = -1;
      for (int base = 0; base < N && cnt < nsample; base += 32) {
        const int k = base + lane;
        bool hit = false;
        if (k < N) {
          const float d2 = sqdist3_ref(qx, qy, qz, s_xyz[k * 3 + 0], s_xyz[k * 3 + 1], s_xyz[k * 3 + 2]);
          hit = d2 < radius2;  // strict, NaN rejects
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal) {
          if (first < 0) first = base + __ffs(bal) - 1;
          const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
          if (hit && pos < nsample) row[pos] = k;
          cnt += __popc(bal);
        }
      }
      if (cnt > nsample) cnt = nsample;
      const int fill = first < 0 ? 0 : first;
      for (int l = cnt + lane; l < nsample; l += 32) row[l] = fill;
    }
  }
}

static int ilog2_floor(int v) {
  int l = 0;
  while ((2 << l) <= v) ++l;
  return l;
}

// Host side of FPS, shared with the fused SA entry points (sa_fused.cu).
int launch_fps(const float *xyz, int ld, int32_t *idx, float *new_xyz, int B, int N, int m,
               cudaStream_t stream, int32_t *ball_idx, float radius, int nsample) {
  MSR3D_REQUIRE(B >= 0 && N >= 1 && m >= 0 && ld >= 3, "fps: bad sizes B=%d N=%d m=%d ld=%d", B, N, m, ld);
  if (B == 0 || m == 0) return MSR3D_OK;
  // reference block size: largest power of two <= N, capped at 512 (cuda_utils.h:13-19)
  int log2bs = ilog2_floor(N);
  if (log2bs > 9) log2bs = 9;
  const int bs = 1 << log2bs;
  const int nper = ceil_div(N, bs);
  MSR3D_REQUIRE(nper <= (1 << kFpsRankShift), "fps: N=%d needs %d points per reference thread (max 16)", N, nper);
  const size_t smem = (size_t)N * 3 * sizeof(float) + (size_t)m * sizeof(int32_t);
  const float radius2 = radius * radius;   // f32 product, as ball_query.cpp / ball_query_gpu.cu:22
  if (nsample <= 0) ball_idx = nullptr;
#define FPS_LAUNCH(T, P)                                                                  \
  do {                                                                                    \
    if (smem > 48 * 1024)                                                                 \
      MSR3D_CUDA(cudaFuncSetAttribute(fps_kernel<T, P>,                                   \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    launch_pdl(fps_kernel<T, P>, dim3(B), dim3(T), smem, stream, xyz, ld, idx, new_xyz, N, m, log2bs, nper,   \
               ball_idx, radius2, nsample);                                                                   \
  } while (0)
  if (N <= 32) FPS_LAUNCH(32, 1);
  else if (N <= 64) FPS_LAUNCH(32, 2);
  else if (N <= 128) FPS_LAUNCH(32, 4);
  else if (N <= 256) FPS_LAUNCH(256, 1);
  else if (N <= 512) FPS_LAUNCH(256, 2);
  else if (N <= 1024) FPS_LAUNCH(256, 4);
  else if (N <= 2048) FPS_LAUNCH(256, 8);
  else if (N <= 4096) FPS_LAUNCH(512, 8);
  else if (N <= 8192) FPS_LAUNCH(1024, 8);
  else
    return set_error(MSR3D_ERR_UNSUPPORTED, "fps: N=%d > 8192 points per cloud not supported", N);
#undef FPS_LAUNCH
  return check_launch("fps_kernel");
}

// ----------------------------------------------------------------------------
// Ball query: one WARP per centre.  Lanes test 32 consecutive points, a ballot
// + popc prefix compacts hits in ascending index order (== the reference's
// serial scan order, ball_query_gpu.cu:27-42), early exit at nsample hits.
// ----------------------------------------------------------------------------
constexpr int kBqWarps = 4;
__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int ld,
                  int32_t *__restrict__ idx, int N, int M, float radius2, int nsample) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * kBqWarps + warp;
  const int b = blockIdx.y;
  if (j >= M) return;
  const float *q = new_xyz + ((size_t)b * M + j) * 3;
  const float qx = q[0], qy = q[1], qz = q[2];
  const float *P = xyz + (size_t)b * N * ld;
  int32_t *row = idx + ((size_t)b * M + j) * nsample;
  int cnt = 0, first = -1;
  for (int base = 0; base < N && cnt < nsample; base += 32) {
    const int k = base + lane;
    bool hit = false;
    if (k < N) {
      const float *pk = P + (size_t)k * ld;
      const float d2 = sqdist3_ref(qx, qy, qz, pk[0], pk[1], pk[2]);
      hit = d2 < radius2;  // strict, NaN rejects
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (bal) {
      if (first < 0) first = base + __ffs(bal) - 1;
      const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
      if (hit && pos < nsample) row[pos] = k;
      cnt += __popc(bal);
    }
  }
  if (cnt > nsample) cnt = nsample;
  // pad with the first hit; an empty ball is all zeros (ball_query.cpp:19-21)
  const int fill = first < 0 ? 0 : first;
  for (int l = cnt + lane; l < nsample; l += 32) row[l] = fill;
}

// Same scan with the cloud staged in shared memory: a CTA = up to 32 centres of ONE cloud (one warp
// each).  The global version above walks the cloud with a chain of dependent L2 round trips (the
// early-exit test sits between consecutive loads): 21 us for 240 clouds x 32 centres x 1024 points;
// here the cloud is fetched once per CTA in one round trip and scanned at shared-memory latency.
__global__ void __launch_bounds__(1024)
ball_query_smem_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int ld,
                       int32_t *__restrict__ idx, int N, int M, float radius2, int nsample) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  extern __shared__ float s_cloud[];  // N * 3
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int b = blockIdx.y;
  const float *P = xyz + (size_t)b * N * ld;
  const int j = blockIdx.x * nwarps + warp;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (j < M) {   // centre coordinates in flight together with the cloud
    const float *q = new_xyz + ((size_t)b * M + j) * 3;
    qx = q[0]; qy = q[1]; qz = q[2];
  }
  if (ld == 3) {
    for (int i = threadIdx.x; i < N * 3; i += blockDim.x) s_cloud[i] = P[i];
  } else {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      s_cloud[i * 3 + 0] = P[(size_t)i * ld + 0];
      s_cloud[i * 3 + 1] = P[(size_t)i * ld + 1];
      s_cloud[i * 3 + 2] = P[(size_t)i * ld + 2];
    }
  }
  __syncthreads();
  if (j >= M) return;
  int32_t *row = idx + ((size_t)b * M + j) * nsample;
  int cnt = 0, first = -1;
  for (int base = 0; base < N && cnt < nsample; base += 32) {
    const int k = base + lane;
    bool hit = false;
    if (k < N) {
      const float d2 = sqdist3_ref(qx, qy, qz, s_cloud[k * 3 + 0], s_cloud[k * 3 + 1], s_cloud[k * 3 + 2]);
      hit = d2 < radius2;  // strict, NaN rejects
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (bal) {
      if (first < 0) first = base + __ffs(bal) - 1;
      const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
      if (hit && pos < nsample) row[pos] = k;
      cnt += __popc(bal);
    }
  }
  if (cnt > nsample) cnt = nsample;
  const int fill = first < 0 ? 0 : first;
  for (int l = cnt + lane; l < nsample; l += 32) row[l] = fill;
}

// ----------------------------------------------------------------------------
// Gather / group (pure index copies) and their scatter-add gradients.
// ----------------------------------------------------------------------------
// out[bc, q] = points[bc, idx[b, q]] with bc = b*C + c, q in [0, Q).
// Covers gather_points (Q = M) and group_points (Q = npoint*nsample).
__global__ void __launch_bounds__(256)
index_gather_kernel(const float *__restrict__ points, const int32_t *__restrict__ idx,
                    float *__restrict__ out, int C, int N, int Q, size_t total) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t bc = i / Q;
    const int q = (int)(i - bc * Q);
    const size_t b = bc / C;
    out[i] = points[bc * N + idx[b * Q + q]];
  }
}
__global__ void __launch_bounds__(256)
index_scatter_add_kernel(const float *__restrict__ grad_out, const int32_t *__restrict__ idx,
                         float *__restrict__ grad_points, int C, int N, int Q, size_t total) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t bc = i / Q;
    const int q = (int)(i - bc * Q);
    const size_t b = bc / C;
    atomicAdd(grad_points + bc * N + idx[b * Q + q], grad_out[i]);
  }
}

// Deterministic, atomic-free form of the two scatter-add gradients:
//   grad_points[b, c, n] = sum over { q : idx[b, q] == n } of grad_out[b, c, q],  q ascending.
// A CTA owns one cloud (and a slice of the channels).  It inverts idx[b, :] into a CSR in shared memory
// with a STABLE counting sort — counts by shared-memory atomics, an exclusive block scan, then one
// warp walks q in order, 32 at a time, and ranks equal keys inside a chunk with __match_any_sync — and
// then every (channel, point) pair adds up its segment in a fixed order.  No global atomics, no
// zero-fill (every output element is written), run-to-run bit-identical; the reference kernels
// (group_points_gpu.cu:43-75, sampling_gpu.cu:34-57) and the fallback above use atomicAdd.
constexpr int kSegThreads = 256;
__global__ void __launch_bounds__(kSegThreads)
index_segment_sum_kernel(const float *__restrict__ grad_out, const int32_t *__restrict__ idx,
                         float *__restrict__ grad_points, int C, int N, int Q, int c_per_cta) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  extern __shared__ int seg_sm[];
  int *start = seg_sm;            // [N + 1] exclusive prefix of the bin sizes
  int *cursor = start + N + 1;    // [N]     counts, then fill cursors
  int *entry = cursor + N;        // [Q]     q values grouped by bin, ascending inside a bin
  __shared__ int s_warp[kSegThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int32_t *ib = idx + (size_t)b * Q;
  for (int n = tid; n < N; n += kSegThreads) cursor[n] = 0;
  __syncthreads();
  for (int q = tid; q < Q; q += kSegThreads) atomicAdd(&cursor[ib[q]], 1);
  __syncthreads();
  // exclusive scan over N bins: a contiguous slab per thread, then a block scan of the slab totals
  const int per = (N + kSegThreads - 1) / kSegThreads;
  const int n0 = min(N, tid * per), n1 = min(N, n0 + per);
  int tot = 0;
  for (int n = n0; n < n1; ++n) tot += cursor[n];
  int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; ++w) base += s_warp[w];
  int run = base + incl - tot;
  for (int n = n0; n < n1; ++n) {
    start[n] = run;
    run += cursor[n];
    cursor[n] = 0;
  }
  if (tid == kSegThreads - 1) start[N] = run;
  __syncthreads();
  if (warp == 0) {   // stable fill: chunks of 32 consecutive q, equal keys ranked by lane
    for (int q0 = 0; q0 < Q; q0 += 32) {
      const int q = q0 + lane;
      const bool ok = q < Q;
      const int key = ok ? ib[q] : -1 - lane;                  // distinct dummies for the tail
      const unsigned m = __match_any_sync(0xffffffffu, key);
      const int rank = __popc(m & ((1u << lane) - 1u));
      const int leader = __ffs(m) - 1;
      int pos = 0;
      if (ok && lane == leader) {
        pos = cursor[key];
        cursor[key] = pos + __popc(m);
      }
      pos = __shfl_sync(0xffffffffu, pos, leader);
      if (ok) entry[start[key] + pos + rank] = q;
      __syncwarp();
    }
  }
  __syncthreads();
  const int c0 = blockIdx.y * c_per_cta, c1 = min(C, c0 + c_per_cta);
  for (int i = tid; i < (c1 - c0) * N; i += kSegThreads) {
    const int c = c0 + i / N, n = i % N;
    const float *g = grad_out + ((size_t)b * C + c) * Q;
    float acc = 0.f;
    for (int e = start[n]; e < start[n + 1]; ++e) acc += g[entry[e]];
    grad_points[((size_t)b * C + c) * N + n] = acc;
  }
}

// shared launcher of the two gradients; returns -1 when the CSR does not fit (caller falls back to atomics)
static int launch_segment_sum(const float *grad_out, const int32_t *idx, float *grad_points, int B, int C, int N, int Q,
                              cudaStream_t stream) {
  const size_t smem = ((size_t)2 * N + 1 + Q) * sizeof(int);
  if (smem > 96 * 1024 || B > 65535 * 32) return -1;
  static bool configured = false;
  if (!configured) {
    MSR3D_CUDA(cudaFuncSetAttribute(index_segment_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  // enough CTAs to fill the GPU: split the channels when there are few clouds
  int csplit = B >= 2 * kNumSMs ? 1 : ceil_div(2 * kNumSMs, B);
  if (csplit > C) csplit = C;
  if (csplit > 65535) csplit = 65535;
  const int c_per_cta = ceil_div(C, csplit);
  launch_pdl(index_segment_sum_kernel, dim3(B, ceil_div(C, c_per_cta)), dim3(kSegThreads), smem, stream, grad_out, idx,
             grad_points, C, N, Q, c_per_cta);
  return check_launch("index_segment_sum");
}

static int flat_grid(size_t total) {
  size_t g = (total + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ----------------------------------------------------------------------------
// three_nn / three_interpolate (feature-propagation ops; not used by
// PointNetPP, exported for binding parity with bindings.cpp:10-12).
// ----------------------------------------------------------------------------
// The reference keeps the running bests as doubles initialised to 1e40 and
// compares an f32 candidate against them (interpolate_gpu.cu:28-50).  Every
// value ever stored is an f32, and (float)1e40 == +inf compares identically
// for all f32 candidates, so f32 bests initialised to +inf are equivalent.
__global__ void __launch_bounds__(128)
three_nn_kernel(const float *__restrict__ unknown, const float *__restrict__ known,
                float *__restrict__ dist2, int32_t *__restrict__ idx, int n, int m) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float *u = unknown + ((size_t)b * n + j) * 3;
  const float ux = u[0], uy = u[1], uz = u[2];
  const float *K = known + (size_t)b * m * 3;
  float best1 = INFINITY, best2 = INFINITY, best3 = INFINITY;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int k = 0; k < m; ++k) {
    const float d = sqdist3_ref(ux, uy, uz, K[k * 3 + 0], K[k * 3 + 1], K[k * 3 + 2]);
    if (d < best1) {
      best3 = best2; i3 = i2;
      best2 = best1; i2 = i1;
      best1 = d;     i1 = k;
    } else if (d < best2) {
      best3 = best2; i3 = i2;
      best2 = d;     i2 = k;
    } else if (d < best3) {
      best3 = d;     i3 = k;
    }
  }
  float *D = dist2 + ((size_t)b * n + j) * 3;
  int32_t *I = idx + ((size_t)b * n + j) * 3;
  D[0] = best1; D[1] = best2; D[2] = best3;
  I[0] = i1; I[1] = i2; I[2] = i3;
}

// out[b,l,j] = sum_t points[b,l,idx[b,j,t]] * weight[b,j,t]; contraction as in the
// reference SASS: r = p2*w2; r = fma(p1,w1,r); r = fma(p3,w3,r).
__global__ void __launch_bounds__(256)
three_interpolate_kernel(const float *__restrict__ points, const int32_t *__restrict__ idx,
                         const float *__restrict__ weight, float *__restrict__ out, int c,
                         int m, int n, size_t total) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t bl = i / n;
    const int j = (int)(i - bl * n);
    const size_t b = bl / c;
    const int32_t *I = idx + (b * n + j) * 3;
    const float *W = weight + (b * n + j) * 3;
    const float *P = points + bl * m;
    float r = __fmul_rn(P[I[1]], W[1]);
    r = __fmaf_rn(P[I[0]], W[0], r);
    r = __fmaf_rn(P[I[2]], W[2], r);
    out[i] = r;
  }
}
__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(const float *__restrict__ grad_out,
                              const int32_t *__restrict__ idx,
                              const float *__restrict__ weight,
                              float *__restrict__ grad_points, int c, int n, int m,
                              size_t total) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t bl = i / n;
    const int j = (int)(i - bl * n);
    const size_t b = bl / c;
    const int32_t *I = idx + (b * n + j) * 3;
    const float *W = weight + (b * n + j) * 3;
    float *G = grad_points + bl * m;
    const float g = grad_out[i];
    atomicAdd(G + I[0], __fmul_rn(g, W[0]));
    atomicAdd(G + I[1], __fmul_rn(g, W[1]));
    atomicAdd(G + I[2], __fmul_rn(g, W[2]));
  }
}

}  // namespace msr3d

using namespace msr3d;

// ------------------------------------------------------------------- C ABI
MSR3D_API int msr3d_furthest_point_sampling_f32(const float *xyz, int32_t *idx, int B, int N,
                                                int m, msr3d_stream_t stream) {
  MSR3D_REQUIRE(xyz && idx, "fps: null pointer");
  return launch_fps(xyz, 3, idx, nullptr, B, N, m, (cudaStream_t)stream, nullptr, 0.f, 0);
}

MSR3D_API int msr3d_gather_points_f32(const float *points, const int32_t *idx, float *out,
                                      int B, int C, int N, int M, msr3d_stream_t stream) {
  MSR3D_REQUIRE(points && idx && out, "gather_points: null pointer");
  MSR3D_REQUIRE(B >= 0 && C >= 0 && N >= 1 && M >= 0, "gather_points: bad sizes");
  const size_t total = (size_t)B * C * M;
  if (total == 0) return MSR3D_OK;
  launch_pdl(index_gather_kernel, dim3(flat_grid(total)), dim3(256), 0, (cudaStream_t)stream, points, idx, out, C,
                                                                           N, M, total);
  return check_launch("gather_points");
}

MSR3D_API int msr3d_gather_points_grad_f32(const float *grad_out, const int32_t *idx,
                                           float *grad_points, int B, int C, int N, int M,
                                           msr3d_stream_t stream) {
  MSR3D_REQUIRE(grad_out && idx && grad_points, "gather_points_grad: null pointer");
  MSR3D_REQUIRE(B >= 0 && C >= 0 && N >= 1 && M >= 0, "gather_points_grad: bad sizes");
  if ((size_t)B * C * N == 0) return MSR3D_OK;
  if (M > 0 && !g_tune[TUNE_GRAD_ATOMICS]) {   // deterministic segmented reduce when the CSR fits in shared memory
    const int rc = launch_segment_sum(grad_out, idx, grad_points, B, C, N, M, (cudaStream_t)stream);
    if (rc >= 0) return rc;
  }
  MSR3D_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)B * C * N,
                             (cudaStream_t)stream));
  const size_t total = (size_t)B * C * M;
  if (total == 0) return MSR3D_OK;
  launch_pdl(index_scatter_add_kernel, dim3(flat_grid(total)), dim3(256), 0, (cudaStream_t)stream, 
      grad_out, idx, grad_points, C, N, M, total);
  return check_launch("gather_points_grad");
}

MSR3D_API int msr3d_ball_query_f32(const float *new_xyz, const float *xyz, int32_t *idx, int B,
                                   int N, int M, float radius, int nsample,
                                   msr3d_stream_t stream) {
  MSR3D_REQUIRE(new_xyz && xyz && idx, "ball_query: null pointer");
  MSR3D_REQUIRE(B >= 0 && N >= 1 && M >= 0 && nsample >= 0, "ball_query: bad sizes");
  MSR3D_REQUIRE(B <= 65535, "ball_query: B=%d exceeds grid.y", B);
  if (B == 0 || M == 0 || nsample == 0) return MSR3D_OK;
  const float radius2 = radius * radius;  // ball_query_gpu.cu:22, f32 product
  dim3 grid(ceil_div(M, kBqWarps), B);
  launch_pdl(ball_query_kernel, dim3(grid), dim3(kBqWarps * 32), 0, (cudaStream_t)stream, new_xyz, xyz, 3, idx, N, M,
                                                                      radius2, nsample);
  return check_launch("ball_query");
}

// Strided variants used by the fused encoder (sa_fused.cu): xyz is the head of a
// record of `ld` floats per point.
namespace msr3d {
int launch_ball_query(const float *new_xyz, const float *xyz, int ld, int32_t *idx, int B, int N,
                      int M, float radius, int nsample, cudaStream_t stream) {
  MSR3D_REQUIRE(B >= 0 && N >= 1 && M >= 0 && nsample >= 0 && ld >= 3 && B <= 65535,
                "ball_query: bad sizes");
  if (B == 0 || M == 0 || nsample == 0) return MSR3D_OK;
  const float radius2 = radius * radius;
  const size_t smem = (size_t)N * 3 * sizeof(float);
  if (smem <= 96 * 1024 && M >= 8) {   // cloud fits in shared memory and there are enough centres to share it
    const int warps = M < 32 ? M : 32;
    static bool configured = false;
    if (!configured) {
      MSR3D_CUDA(cudaFuncSetAttribute(ball_query_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      configured = true;
    }
    launch_pdl(ball_query_smem_kernel, dim3(ceil_div(M, warps), B), dim3(warps * 32), smem, stream, new_xyz, xyz, ld, idx,
               N, M, radius2, nsample);
    return check_launch("ball_query_smem");
  }
  dim3 grid(ceil_div(M, kBqWarps), B);
  launch_pdl(ball_query_kernel, dim3(grid), dim3(kBqWarps * 32), 0, stream, new_xyz, xyz, ld, idx, N, M, radius2, nsample);
  return check_launch("ball_query");
}
}  // namespace msr3d

MSR3D_API int msr3d_group_points_f32(const float *points, const int32_t *idx, float *out, int B,
                                     int C, int N, int npoint, int nsample,
                                     msr3d_stream_t stream) {
  MSR3D_REQUIRE(points && idx && out, "group_points: null pointer");
  MSR3D_REQUIRE(B >= 0 && C >= 0 && N >= 1 && npoint >= 0 && nsample >= 0,
                "group_points: bad sizes");
  const size_t total = (size_t)B * C * npoint * nsample;
  if (total == 0) return MSR3D_OK;
  launch_pdl(index_gather_kernel, dim3(flat_grid(total)), dim3(256), 0, (cudaStream_t)stream, 
      points, idx, out, C, N, npoint * nsample, total);
  return check_launch("group_points");
}

MSR3D_API int msr3d_group_points_grad_f32(const float *grad_out, const int32_t *idx,
                                          float *grad_points, int B, int C, int N, int npoint,
                                          int nsample, msr3d_stream_t stream) {
  MSR3D_REQUIRE(grad_out && idx && grad_points, "group_points_grad: null pointer");
  MSR3D_REQUIRE(B >= 0 && C >= 0 && N >= 1 && npoint >= 0 && nsample >= 0,
                "group_points_grad: bad sizes");
  if ((size_t)B * C * N == 0) return MSR3D_OK;
  if ((size_t)npoint * nsample > 0 && !g_tune[TUNE_GRAD_ATOMICS]) {
    const int rc = launch_segment_sum(grad_out, idx, grad_points, B, C, N, npoint * nsample, (cudaStream_t)stream);
    if (rc >= 0) return rc;
  }
  MSR3D_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)B * C * N,
                             (cudaStream_t)stream));
  const size_t total = (size_t)B * C * npoint * nsample;
  if (total == 0) return MSR3D_OK;
  launch_pdl(index_scatter_add_kernel, dim3(flat_grid(total)), dim3(256), 0, (cudaStream_t)stream, 
      grad_out, idx, grad_points, C, N, npoint * nsample, total);
  return check_launch("group_points_grad");
}

MSR3D_API int msr3d_three_nn_f32(const float *unknown, const float *known, float *dist2,
                                 int32_t *idx, int B, int n, int m, msr3d_stream_t stream) {
  MSR3D_REQUIRE(unknown && known && dist2 && idx, "three_nn: null pointer");
  MSR3D_REQUIRE(B >= 0 && n >= 0 && m >= 0 && B <= 65535, "three_nn: bad sizes");
  if (B == 0 || n == 0) return MSR3D_OK;
  dim3 grid(ceil_div(n, 128), B);
  launch_pdl(three_nn_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, unknown, known, dist2, idx, n, m);
  return check_launch("three_nn");
}

MSR3D_API int msr3d_three_interpolate_f32(const float *points, const int32_t *idx,
                                          const float *weight, float *out, int B, int c, int m,
                                          int n, msr3d_stream_t stream) {
  MSR3D_REQUIRE(points && idx && weight && out, "three_interpolate: null pointer");
  MSR3D_REQUIRE(B >= 0 && c >= 0 && m >= 1 && n >= 0, "three_interpolate: bad sizes");
  const size_t total = (size_t)B * c * n;
  if (total == 0) return MSR3D_OK;
  launch_pdl(three_interpolate_kernel, dim3(flat_grid(total)), dim3(256), 0, (cudaStream_t)stream, 
      points, idx, weight, out, c, m, n, total);
  return check_launch("three_interpolate");
}

MSR3D_API int msr3d_three_interpolate_grad_f32(const float *grad_out, const int32_t *idx,
                                               const float *weight, float *grad_points, int B,
                                               int c, int n, int m, msr3d_stream_t stream) {
  MSR3D_REQUIRE(grad_out && idx && weight && grad_points, "three_interpolate_grad: null pointer");
  MSR3D_REQUIRE(B >= 0 && c >= 0 && m >= 1 && n >= 0, "three_interpolate_grad: bad sizes");
  if ((size_t)B * c * m == 0) return MSR3D_OK;
  MSR3D_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)B * c * m,
                             (cudaStream_t)stream));
  const size_t total = (size_t)B * c * n;
  if (total == 0) return MSR3D_OK;
  launch_pdl(three_interpolate_grad_kernel, dim3(flat_grid(total)), dim3(256), 0, (cudaStream_t)stream, 
      grad_out, idx, weight, grad_points, c, n, m, total);
  return check_launch("three_interpolate_grad");
}
