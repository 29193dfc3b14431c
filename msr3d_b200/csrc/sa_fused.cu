// Group B: one PointNet++ set-abstraction level as ONE tensor-core kernel.
//
// Reference path being replaced (per level, pointnet2_modules.py:34-75):
//   ball_query -> group_points(xyz) -> subtract centre -> group_points(feats) ->
//   cat -> 3 x [conv1x1 -> BatchNorm2d(eval) -> ReLU] -> max_pool over nsample
// i.e. >= 15 kernels that round-trip a (B, C, npoint, nsample) tensor through
// HBM after every op (SA2: 63 MB per intermediate).  Here a CTA owns a tile of
// 128 (centre, sample) rows = 4 centres x 32 samples and never leaves the SM:
//
//   gather : rows are built straight from the ball-query indices into the
//            K-major bf16 hi/lo operand image in shared memory (the grouped
//            tensor is never materialised);
//   layer i: tcgen05.mma (split-bf16, 3 MMAs per K=16 slice) into its own TMEM
//            columns; weights (BatchNorm folded in) arrive by TMA bulk copy;
//   between: TMEM -> registers -> +bias, ReLU -> split -> next layer's operand
//            image in shared memory (aliasing the previous one);
//   pool   : after layer 3 a warp holds the 32 samples of one centre in its 32
//            TMEM lanes: REDUX.MAX per column, then one coalesced row store.
//
// Inter-level tensors are point-major ([cloud][point][channel]) so a gathered
// neighbour is one contiguous row.
#include "common.cuh"
#include "umma.cuh"

namespace msr3d {

using namespace umma;

// ext_ops.cu
int launch_fps(const float *xyz, int ld, int32_t *idx, float *new_xyz, int B, int N, int m,
               cudaStream_t stream, int32_t *ball_idx, float radius, int nsample);
int launch_ball_query(const float *new_xyz, const float *xyz, int ld, int32_t *idx, int B, int N,
                      int M, float radius, int nsample, cudaStream_t stream);

struct SaParams {
  const float *feat;      // [B', Nsrc, ld_feat]  neighbour features
  const float *xyz;       // [B', Nsrc, ld_xyz]   neighbour coordinates (first 3 floats)
  const float *new_xyz;   // [B', npoint, 3]      centres
  const int32_t *ball;    // [B', npoint, 32]     ball-query indices
  const uint8_t *w[3];    // packed, BN-folded weights per layer (blocks of <=128 rows)
  const float *b[3];      // folded bias per layer
  float *out;             // [B', npoint, ld_out] pooled features
  long long ld_out;
  int ld_feat, ld_xyz, nfeat, Nsrc, npoint, n_tiles, split;
};

template <int K0P, int C1, int C2, int C3, bool RESIDENT>
struct SaLayout {
  static constexpr int kRows = 128;
  static constexpr uint32_t a_bytes(int k) { return kRows * k * 4; }  // hi + lo
  static constexpr uint32_t w_bytes(int n, int k) { return n * k * 4; }
  static constexpr uint32_t kA = a_bytes(K0P) > a_bytes(C1) ? (a_bytes(K0P) > a_bytes(C2) ? a_bytes(K0P) : a_bytes(C2))
                                                             : (a_bytes(C1) > a_bytes(C2) ? a_bytes(C1) : a_bytes(C2));
  static constexpr uint32_t kW1 = w_bytes(C1, K0P), kW2 = w_bytes(C2, C1), kW3 = w_bytes(C3, C2);
  static constexpr int kNB3 = C3 > 128 ? C3 / 128 : 1;        // layer-3 weight blocks
  static constexpr uint32_t kW3Block = kW3 / kNB3;
  // resident: all weights side by side.  streaming: two ping-pong buffers.
  static constexpr uint32_t kBuf0 = kW1 > kW3Block ? kW1 : kW3Block;
  static constexpr uint32_t kBuf1 = kW2 > kW3Block ? kW2 : kW3Block;
  static constexpr uint32_t kWTotal = RESIDENT ? (kW1 + kW2 + kW3) : (kBuf0 + kBuf1);
  static constexpr uint32_t kBias = (C1 + C2 + C3) * 4;
  static constexpr uint32_t kSmem = kA + kWTotal + kBias + 128;
  static constexpr int kTmemCols = C1 + C2 + C3;
  static_assert(C1 <= 128 && C2 <= 128, "inner layers are single MMA-N blocks");
  static_assert(RESIDENT || kNB3 == 2, "streaming mode expects a 256-wide last layer");
  static_assert((kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols >= 32 && kTmemCols <= 512, "TMEM cols");
};

// issue one layer: D[:, dcol .. dcol+N) = A[128 x K] * Wblock[N x K]^T
__device__ __forceinline__ void issue_layer(uint32_t a_addr, int K, uint32_t w_addr, int N,
                                            uint32_t tmem_d, int split) {
  // K-slice descriptors advance by a constant (2*rows in 16 B units): the single issuing thread
  // does two 64-bit adds per MMA instead of rebuilding descriptors
  uint64_t dah = kslice_desc(a_addr, 128, 0), dal = kslice_desc(a_addr + 128u * K * 2u, 128, 0);
  uint64_t dbh = kslice_desc(w_addr, N, 0), dbl = kslice_desc(w_addr + (uint32_t)N * K * 2u, N, 0);
  const uint32_t idesc = make_idesc_bf16(128, N);
  const uint64_t step_a = 2 * 128, step_b = 2 * (uint64_t)N;
#pragma unroll 2
  for (int ks = 0; ks < K / 16; ++ks) {
    mma_bf16(tmem_d, dah, dbh, idesc, ks ? 1u : 0u);
    if (split == 3) {
      mma_bf16(tmem_d, dah, dbl, idesc, 1u);
      mma_bf16(tmem_d, dal, dbh, idesc, 1u);
    }
    dah += step_a; dal += step_a; dbh += step_b; dbl += step_b;
  }
}

// TMEM -> (+bias, ReLU) -> split bf16 -> operand image of the next layer.
// Warp w: lanes 32*(w&3).., columns [h*C/2, (h+1)*C/2) with h = w>>2.
template <int C>
__device__ __forceinline__ void epilogue_to_operand(uint32_t tmem_d, const float *s_bias,
                                                    uint8_t *a_next, int warp, int lane) {
  const int q = warp & 3, h = warp >> 2;
  const int row = q * 32 + lane;
  constexpr int kHalf = C / 2;
  static_assert(kHalf % 32 == 0, "column halves are multiples of 32");
#pragma unroll 1
  for (int c0 = h * kHalf; c0 < (h + 1) * kHalf; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = g * 8 + 2 * j;
        const float v0 = fmaxf(__uint_as_float(r[c]) + s_bias[c0 + c], 0.f);
        const float v1 = fmaxf(__uint_as_float(r[c + 1]) + s_bias[c0 + c + 1], 0.f);
        split_pack2(v0, v1, hi[j], lo[j]);
      }
      const uint32_t off = chunk_offset(128, row, (c0 >> 3) + g);
      *reinterpret_cast<uint4 *>(a_next + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4 *>(a_next + 128u * C * 2u + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

template <int K0P, int C1, int C2, int C3, bool RESIDENT>
__global__ void __launch_bounds__(256, RESIDENT ? 2 : 1)
sa_mlp_kernel(const SaParams p) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  using L = SaLayout<K0P, C1, C2, C3, RESIDENT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *s_a = smem;
  uint8_t *s_w = smem + L::kA;
  float *s_bias = reinterpret_cast<float *>(smem + L::kA + L::kWTotal);
  __shared__ __align__(8) uint64_t bar_w[2];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar_w[0], 1);
    mbar_init(&bar_w[1], 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem, L::kTmemCols);
  for (int i = tid; i < C1; i += 256) s_bias[i] = p.b[0][i];
  for (int i = tid; i < C2; i += 256) s_bias[C1 + i] = p.b[1][i];
  for (int i = tid; i < C3; i += 256) s_bias[C1 + C2 + i] = p.b[2][i];
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s_tmem;
  const uint32_t d1 = tmem, d2 = tmem + C1, d3 = tmem + C1 + C2;
  const uint32_t a_addr = smem_u32(s_a), w_addr = smem_u32(s_w);

  uint32_t mma_uses = 0, w_uses0 = 0, w_uses1 = 0;
  // Everything that drives the uniform datapath (TMA bulk copies, tcgen05.mma / commit) is issued by ONE ELECTED
  // LANE OF WARP 0 inside a warp-uniform branch: under `if (tid == 0)` the compiler wraps every such instruction
  // in an ELECT / BRA.U.ANY serialisation loop (53 of them in the round-1 build of this file).
  if (RESIDENT && warp == 0 && elect_one()) {
    mbar_expect_tx(&bar_w[0], L::kW1 + L::kW2 + L::kW3);
    bulk_g2s(s_w, p.w[0], L::kW1, &bar_w[0]);
    bulk_g2s(s_w + L::kW1, p.w[1], L::kW2, &bar_w[0]);
    bulk_g2s(s_w + L::kW1 + L::kW2, p.w[2], L::kW3, &bar_w[0]);
  }

  const int tiles_per_cloud = p.npoint / 4;
  const int row = tid & 127, half = tid >> 7;
  // neighbour index of this thread's row in tile t (the first link of the gather's dependent chain:
  // index -> feature row -> operand image); the one of the NEXT tile is fetched a whole tile ahead
  auto ball_of = [&](int t) -> int {
    const int cl = t / tiles_per_cloud, c0 = (t % tiles_per_cloud) * 4;
    return p.ball[((size_t)cl * p.npoint + c0 + (row >> 5)) * 32 + (row & 31)];
  };
  int nb_next = (int)blockIdx.x < p.n_tiles ? ball_of(blockIdx.x) : 0;
  for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
    const int cloud = t / tiles_per_cloud;
    const int centre0 = (t % tiles_per_cloud) * 4;
    if (!RESIDENT && warp == 0 && elect_one()) {  // all MMAs of the previous tile are complete: both buffers are free
      mbar_expect_tx(&bar_w[0], L::kW1);
      bulk_g2s(s_w, p.w[0], L::kW1, &bar_w[0]);
      mbar_expect_tx(&bar_w[1], L::kW2);
      bulk_g2s(s_w + L::kBuf0, p.w[1], L::kW2, &bar_w[1]);
    }
    // ---------------- gather: build the layer-1 operand image (K order: feats | dxyz | 0)
    {
      const int centre = centre0 + (row >> 5);
      const int nb = nb_next;
      if (t + (int)gridDim.x < p.n_tiles) nb_next = ball_of(t + gridDim.x);
      const float *f = p.feat + ((size_t)cloud * p.Nsrc + nb) * p.ld_feat;
      const float *x = p.xyz + ((size_t)cloud * p.Nsrc + nb) * p.ld_xyz;
      const float *c = p.new_xyz + ((size_t)cloud * p.npoint + centre) * 3;
      const bool vec_ok = (p.ld_feat & 3) == 0 && (reinterpret_cast<uintptr_t>(p.feat) & 15) == 0;
      // chunks ch = half, half+2, ...: every 16 B load of a batch is issued before any conversion.  The
      // streaming variant (one CTA per SM, 255 registers available) takes the whole row in ONE batch
      // (nine chunks = 72 registers, one L2 round trip per tile instead of three); the resident
      // variant has a single batch anyway.
      constexpr int kChunks = K0P / 8;
      constexpr int kIters = (kChunks + 1) / 2;
      constexpr int kBatch = RESIDENT ? 3 : (kIters < 9 ? kIters : 9);
#pragma unroll
      for (int it0 = 0; it0 < kIters; it0 += kBatch) {
        float v[kBatch][8];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int ch = half + 2 * (it0 + u);
          const int k0 = ch * 8;
          if (it0 + u < kIters && ch < kChunks) {
            if (vec_ok && k0 + 8 <= p.nfeat) {
              const float4 a = *reinterpret_cast<const float4 *>(f + k0);
              const float4 b = *reinterpret_cast<const float4 *>(f + k0 + 4);
              v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w;
              v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int k = k0 + j;
                float val = 0.f;
                if (k < p.nfeat) val = f[k];
                else if (k < p.nfeat + 3) val = __fsub_rn(x[k - p.nfeat], c[k - p.nfeat]);  // pointnet2_utils.py:347
                v[u][j] = val;
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int ch = half + 2 * (it0 + u);
          if (it0 + u < kIters && ch < kChunks) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split_pack2(v[u][2 * j], v[u][2 * j + 1], hi[j], lo[j]);
            const uint32_t off = chunk_offset(128, row, ch);
            *reinterpret_cast<uint4 *>(s_a + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(s_a + 128u * K0P * 2u + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
    }
    fence_async_smem();
    __syncthreads();

    // ---------------- layer 1
    if (warp == 0) {
      if (RESIDENT) { if (t == (int)blockIdx.x) mbar_wait(&bar_w[0], 0); }
      else mbar_wait(&bar_w[0], w_uses0 & 1);
      fence_after_sync();
      if (elect_one()) {
        issue_layer(a_addr, K0P, w_addr, C1, d1, p.split);
        mma_commit(&bar_mma);
      }
      __syncwarp();
    }
    if (!RESIDENT) ++w_uses0;
    mbar_wait(&bar_mma, mma_uses & 1); ++mma_uses;
    fence_after_sync();
    if (!RESIDENT && warp == 0 && elect_one()) {  // buffer 0 is free: prefetch first half of layer 3
      mbar_expect_tx(&bar_w[0], L::kW3Block);
      bulk_g2s(s_w, p.w[2], L::kW3Block, &bar_w[0]);
    }
    epilogue_to_operand<C1>(d1, s_bias, s_a, warp, lane);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();

    // ---------------- layer 2
    if (warp == 0) {
      if (!RESIDENT) mbar_wait(&bar_w[1], w_uses1 & 1);
      fence_after_sync();
      if (elect_one()) {
        issue_layer(a_addr, C1, RESIDENT ? w_addr + L::kW1 : w_addr + L::kBuf0, C2, d2, p.split);
        mma_commit(&bar_mma);
      }
      __syncwarp();
    }
    if (!RESIDENT) ++w_uses1;
    mbar_wait(&bar_mma, mma_uses & 1); ++mma_uses;
    fence_after_sync();
    if (!RESIDENT && warp == 0 && elect_one()) {  // buffer 1 is free: second half of layer 3
      mbar_expect_tx(&bar_w[1], L::kW3Block);
      bulk_g2s(s_w + L::kBuf0, p.w[2] + L::kW3Block, L::kW3Block, &bar_w[1]);
    }
    epilogue_to_operand<C2>(d2, s_bias + C1, s_a, warp, lane);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();

    // ---------------- layer 3
    if (warp == 0) {
      fence_after_sync();
      if (RESIDENT) {
        if (elect_one()) {
          constexpr int NB = L::kNB3, BNN = C3 / NB;
          for (int nb = 0; nb < NB; ++nb)
            issue_layer(a_addr, C2, w_addr + L::kW1 + L::kW2 + nb * L::kW3Block, BNN, d3 + nb * BNN, p.split);
          mma_commit(&bar_mma);
        }
      } else {
        mbar_wait(&bar_w[0], w_uses0 & 1);
        fence_after_sync();
        if (elect_one()) issue_layer(a_addr, C2, w_addr, 128, d3, p.split);
        __syncwarp();
        mbar_wait(&bar_w[1], w_uses1 & 1);
        fence_after_sync();
        if (elect_one()) {
          issue_layer(a_addr, C2, w_addr + L::kBuf0, 128, d3 + 128, p.split);
          mma_commit(&bar_mma);
        }
      }
      __syncwarp();
    }
    if (!RESIDENT) { ++w_uses0; ++w_uses1; }
    mbar_wait(&bar_mma, mma_uses & 1); ++mma_uses;
    fence_after_sync();

    // ---------------- pool: max over the 32 samples (= 32 lanes) of each centre
    {
      const int q = warp & 3, h = warp >> 2;
      const int centre = centre0 + q;
      float *orow = p.out + ((size_t)cloud * p.npoint + centre) * p.ld_out;
      const float *b3 = s_bias + C1 + C2;
      constexpr int kHalf = C3 / 2;
#pragma unroll 1
      for (int c0 = h * kHalf; c0 < (h + 1) * kHalf; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(d3 + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        float mine = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = fmaxf(__uint_as_float(r[j]) + b3[c0 + j], 0.f);  // >= +0: uint order == float order
          const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(v));
          if (lane == j) mine = __uint_as_float(m);
        }
        orow[c0 + lane] = mine;
      }
    }
    fence_before_sync();
    __syncthreads();  // TMEM and the operand image are reused by the next tile
    fence_after_sync();
  }
  if (warp == 0) tmem_dealloc(tmem, L::kTmemCols);
}

template <int K0P, int C1, int C2, int C3, bool RESIDENT>
static int launch_sa_mlp(const SaParams &p, cudaStream_t stream) {
  using L = SaLayout<K0P, C1, C2, C3, RESIDENT>;
  auto kern = sa_mlp_kernel<K0P, C1, C2, C3, RESIDENT>;
  static bool configured = false;
  if (!configured) {
    MSR3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kSmem));
    configured = true;
  }
  const int per_sm = RESIDENT ? 2 : 1;
  int sms = kNumSMs;
  if (g_tune[TUNE_SA_SM_LIMIT] > 0 && g_tune[TUNE_SA_SM_LIMIT] < kNumSMs) sms = g_tune[TUNE_SA_SM_LIMIT];  // A/B: leave SMs to a concurrent chain
  // oversubscription (A/B): k x as many CTAs as fit at once, each walking 1/k of the tiles — a CTA then lives a few
  // microseconds and a concurrent high-priority chain gets the SM at the next retirement instead of after the whole level
  const int over = g_tune[TUNE_SA_OVERSUB] > 0 ? g_tune[TUNE_SA_OVERSUB] : 2;   // 2: measured best (profiles/r2_ab_experiments.md)
  int grid = p.n_tiles < sms * per_sm * over ? p.n_tiles : sms * per_sm * over;
  launch_pdl(kern, dim3(grid), dim3(256), L::kSmem, stream, p);
  return check_launch("sa_mlp_kernel");
}

}  // namespace msr3d

using namespace msr3d;

// Front end of one SA level: FPS fused with the centre gather, then ball query
// (pointnet2_modules.py:54-58 + pointnet2_utils.py:331).  xyz is the head of a record
// of `ld` floats per point ((B,N,ld) with ld >= 3).  Outputs: fps_idx (B,m) i32,
// new_xyz (B,m,3) f32, ball_idx (B,m,nsample) i32.
MSR3D_API int msr3d_sa_group_f32(const float *xyz, int ld, int B, int N, int m, float radius,
                                 int nsample, int32_t *fps_idx, float *new_xyz, int32_t *ball_idx,
                                 msr3d_stream_t stream) {
  MSR3D_REQUIRE(xyz && fps_idx && new_xyz && ball_idx, "sa_group: null pointer");
  MSR3D_REQUIRE(nsample >= 0, "sa_group: bad nsample %d", nsample);
  // Two launches by default.  Fusing the ball query into the FPS kernel (the cloud is already in its shared
  // memory) saves a launch and a re-read, but the FPS CTAs then walk 32 centres with 8 warps while the
  // rest of the GPU idles: measured +38 us per pipelined step (A/B with MSR3D_TUNE=6=1).
  if (!g_tune[TUNE_SA_GROUP_FUSED]) {
    int rc = launch_fps(xyz, ld, fps_idx, new_xyz, B, N, m, (cudaStream_t)stream, nullptr, 0.f, 0);
    if (rc != MSR3D_OK) return rc;
    return launch_ball_query(new_xyz, xyz, ld, ball_idx, B, N, m, radius, nsample, (cudaStream_t)stream);
  }
  // one launch: FPS, centre gather and ball query share the cloud staged in shared memory
  return launch_fps(xyz, ld, fps_idx, new_xyz, B, N, m, (cudaStream_t)stream, ball_idx, radius, nsample);
}

namespace msr3d {
// dst[r, col0 + c] = src[r, c] for c < 3 (GroupAll's xyz channels, pointnet2_utils.py:407)
__global__ void __launch_bounds__(256)
scatter_xyz_cols_kernel(const float *__restrict__ src, float *__restrict__ dst, long long ld_dst,
                        int col0, int rows) {
  pdl_sync();  // PDL: wait for the producer grid, then let the consumer grid start scheduling
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * 3) dst[(long long)(i / 3) * ld_dst + col0 + i % 3] = src[i];
}
}  // namespace msr3d

/* Writes xyz (rows,3) into columns [col0, col0+3) of a (rows, ld_dst) matrix. */
MSR3D_API int msr3d_scatter_xyz_cols_f32(const float *xyz, float *dst, long long ld_dst, int col0,
                                         int rows, msr3d_stream_t stream) {
  MSR3D_REQUIRE(xyz && dst && rows >= 0 && col0 >= 0 && ld_dst >= col0 + 3, "scatter_xyz_cols: bad args");
  if (rows == 0) return MSR3D_OK;
  launch_pdl(scatter_xyz_cols_kernel, dim3(ceil_div(rows * 3, 256)), dim3(256), 0, (cudaStream_t)stream, xyz, dst, ld_dst,
                                                                                    col0, rows);
  return check_launch("scatter_xyz_cols");
}

MSR3D_API int msr3d_sa_mlp_fused_f32(const float *feat, int ld_feat, int nfeat, const float *xyz,
                                     int ld_xyz, const float *new_xyz, const int32_t *ball_idx,
                                     int B, int Nsrc, int npoint, int nsample, const void *w1,
                                     const void *w2, const void *w3, const float *b1,
                                     const float *b2, const float *b3, int c1, int c2, int c3,
                                     float *out, long long ld_out, int split,
                                     msr3d_stream_t stream) {
  MSR3D_REQUIRE(feat && xyz && new_xyz && ball_idx && w1 && w2 && w3 && b1 && b2 && b3 && out,
                "sa_mlp_fused: null pointer");
  MSR3D_REQUIRE(nsample == 32, "sa_mlp_fused: nsample must be 32 (got %d)", nsample);
  MSR3D_REQUIRE(npoint > 0 && npoint % 4 == 0, "sa_mlp_fused: npoint must be a multiple of 4");
  MSR3D_REQUIRE(split == 1 || split == 3, "sa_mlp_fused: split must be 1 or 3");
  MSR3D_REQUIRE(ld_out >= c3 && ld_feat >= nfeat && ld_xyz >= 3, "sa_mlp_fused: bad leading dims");
  if (B == 0) return MSR3D_OK;
  SaParams p;
  p.feat = feat; p.xyz = xyz; p.new_xyz = new_xyz; p.ball = ball_idx;
  p.w[0] = (const uint8_t *)w1; p.w[1] = (const uint8_t *)w2; p.w[2] = (const uint8_t *)w3;
  p.b[0] = b1; p.b[1] = b2; p.b[2] = b3;
  p.out = out; p.ld_out = ld_out; p.ld_feat = ld_feat; p.ld_xyz = ld_xyz; p.nfeat = nfeat;
  p.Nsrc = Nsrc; p.npoint = npoint; p.n_tiles = B * (npoint / 4); p.split = split;
  const int k0 = nfeat + 3;
  if (k0 <= 16 && c1 == 64 && c2 == 64 && c3 == 128)
    return launch_sa_mlp<16, 64, 64, 128, true>(p, (cudaStream_t)stream);
  if (k0 <= 144 && k0 > 128 && c1 == 128 && c2 == 128 && c3 == 256)
    return launch_sa_mlp<144, 128, 128, 256, false>(p, (cudaStream_t)stream);
  return set_error(MSR3D_ERR_UNSUPPORTED,
                   "sa_mlp_fused: no fused instance for widths %d->%d->%d->%d (use the layer-wise path)",
                   k0, c1, c2, c3);
}
