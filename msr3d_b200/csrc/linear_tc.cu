// Dense contractions of the hot path on the 5th-gen tensor cores.
//
//   out[M, N] (+)= act( A[M, K] * W[N, K]^T + bias[N] )          (optionally max-pooled
//                                                                  over groups of rows)
//
// tcgen05.mma kind::f16 with bf16 operands and f32 accumulation in TMEM.  The
// reference computes these layers in fp32 (cuDNN/cuBLAS); to stay within the
// 1e-3 budget of the north star with margin, operands are split x = hi + lo
// (two bf16) and three MMAs (hi*hi + hi*lo + lo*hi) accumulate into the same
// TMEM tile: ~2^-16 relative error per product instead of 2^-8 (split = 1 runs
// the single-pass bf16 variant for comparison).
//
// Data movement per CTA (128 x BN output tile, K streamed in chunks of 64):
//   W : pre-packed once per weight update (pack_weight_kernel) into the exact
//       shared-memory image the MMA wants; one TMA bulk copy (UBLKCP) per chunk,
//       completing on an mbarrier.
//   A : fp32 activations read from global with arbitrary strides, split to
//       bf16 hi/lo in registers and stored in the canonical K-major layout
//       (umma.cuh).  Double-buffered: chunk k+1 is loaded while the tensor core
//       works on chunk k; tcgen05.commit releases a stage back to the loaders.
//   D : 128 lanes x BN columns of TMEM; epilogue = tcgen05.ld -> bias -> act
//       (-> row-group max) -> global.
#include "common.cuh"
#include "umma.cuh"

namespace msr3d {

using namespace umma;

constexpr int kTileM = 128;
constexpr int kChunkK = 64;  // K elements per pipeline stage

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

// ---------------------------------------------------------------------------
// Weight packing: W (N x K, element (n,k) at W[n*sn + k*sk], optional per-row
// scale = folded BatchNorm, optional column permutation) -> blocks of
// [BN rows x BK k] in canonical layout, hi tile then lo tile, block order
// (n_tile major, k_block minor).  One thread per (row, 8-wide k chunk).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float *__restrict__ W, int N, int K, long long sn, long long sk,
                   const float *__restrict__ row_scale, const int *__restrict__ col_perm,
                   int BN, int BK, int Np, int Kp, uint8_t *__restrict__ out) {
  const int chunks_per_row = Kp / 8;
  const long long total = (long long)Np * chunks_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % Np);
    const int kc = (int)(i / Np);  // global 8-chunk index along K
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = kc * 8 + j;
      if (col_perm != nullptr) k = (k < Kp) ? col_perm[k] : -1;
      float x = 0.f;
      if (n < N && k >= 0 && k < K) {
        x = W[(long long)n * sn + (long long)k * sk];
        if (row_scale != nullptr) x *= row_scale[n];
      }
      v[j] = x;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_pack2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
    const int nt = n / BN, nl = n % BN;
    const int kb = (kc * 8) / BK, kcl = kc % (BK / 8);
    const size_t block_bytes = (size_t)BN * BK * 4;  // hi + lo
    uint8_t *blk = out + ((size_t)nt * (Kp / BK) + kb) * block_bytes;
    const size_t off = (size_t)kcl * BN * 16 + (size_t)nl * 16;
    *reinterpret_cast<uint4 *>(blk + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4 *>(blk + (size_t)BN * BK * 2 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  return v;
}

// ---------------------------------------------------------------------------
// The GEMM.  256 threads: every thread loads/splits A; thread 0 drives TMA and
// the tensor core; all 8 warps run the epilogue (warp w owns TMEM lanes
// 32*(w%4).., column half w/4).
// ---------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(256, 1)
linear_tc_kernel(const float *__restrict__ A, long long sam, long long sak, int M, int K,
                 const uint8_t *__restrict__ PW, int Kp, const float *__restrict__ bias,
                 float *__restrict__ out, long long ldo, int N, int act, int pool, int split,
                 int accumulate) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t kATile = kTileM * kChunkK * 2;  // one of hi / lo
  constexpr uint32_t kBBlock = BN * kChunkK * 4;     // hi + lo
  constexpr uint32_t kStage = 2 * kATile + kBBlock;
  __shared__ __align__(8) uint64_t b_full[2];
  __shared__ __align__(8) uint64_t mma_done[2];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * kTileM;
  const int nt = blockIdx.y;
  const int nkc = Kp / kChunkK;

  if (tid == 0) {
    mbar_init(&b_full[0], 1);
    mbar_init(&b_full[1], 1);
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem, BN);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = s_tmem;
  const uint32_t idesc = make_idesc_bf16(kTileM, BN);

  const int row = tid & 127;        // A row handled by this thread
  const int khalf = tid >> 7;       // which 32-wide half of the 64-wide chunk
  const bool row_ok = (m0 + row) < M;
  const float *arow = A + (long long)(m0 + row) * sam;
  const bool vec_ok = (sak == 1) && ((sam & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);

  for (int kc = 0; kc < nkc; ++kc) {
    const int s = kc & 1;
    uint8_t *st = smem + (size_t)s * kStage;
    if (kc >= 2) mbar_wait(&mma_done[s], ((kc >> 1) - 1) & 1);  // stage free again
    if (tid == 0) {
      mbar_expect_tx(&b_full[s], kBBlock);
      bulk_g2s(st + 2 * kATile, PW + ((size_t)nt * nkc + kc) * kBBlock, kBBlock, &b_full[s]);
    }
    // ---- A chunk: 4 x (8 floats -> 16 B hi + 16 B lo) per thread
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int kchunk = khalf * 4 + c;  // 0..7 inside the stage
      const int k0 = kc * kChunkK + kchunk * 8;
      float v[8];
      if (row_ok && vec_ok && k0 + 8 <= K) {
        const float4 a = *reinterpret_cast<const float4 *>(arow + k0);
        const float4 b = *reinterpret_cast<const float4 *>(arow + k0 + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          v[j] = (row_ok && k0 + j < K) ? arow[(long long)(k0 + j) * sak] : 0.f;
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_pack2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
      const uint32_t off = chunk_offset(kTileM, row, kchunk);
      *reinterpret_cast<uint4 *>(st + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4 *>(st + kATile + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&b_full[s], (kc >> 1) & 1);
      fence_after_sync();
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + kATile;
      const uint32_t b_hi = a_hi + 2 * kATile, b_lo = b_hi + BN * kChunkK * 2;
#pragma unroll
      for (int ks = 0; ks < kChunkK / 16; ++ks) {
        const uint64_t dah = kslice_desc(a_hi, kTileM, ks), dal = kslice_desc(a_lo, kTileM, ks);
        const uint64_t dbh = kslice_desc(b_hi, BN, ks), dbl = kslice_desc(b_lo, BN, ks);
        mma_bf16(tmem_d, dah, dbh, idesc, (kc | ks) ? 1u : 0u);
        if (split == 3) {
          mma_bf16(tmem_d, dah, dbl, idesc, 1u);
          mma_bf16(tmem_d, dal, dbh, idesc, 1u);
        }
      }
      mma_commit(&mma_done[s]);
    }
  }
  // the last commit covers every MMA issued before it
  {
    const int last = nkc - 1;
    mbar_wait(&mma_done[last & 1], (last >> 1) & 1);
    fence_after_sync();
  }

  // ---- epilogue
  const int q = warp & 3, h = warp >> 2;
  const int grow = m0 + q * 32 + lane;  // global row of this thread's TMEM lane
  constexpr int kColsPerHalf = BN / 2;
#pragma unroll 1
  for (int c0 = h * kColsPerHalf; c0 < (h + 1) * kColsPerHalf; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
    tmem_ld_wait();
    const int n0 = nt * BN + c0;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n0 + j;
      float x = __uint_as_float(r[j]);
      if (bias != nullptr && n < N) x += bias[n];
      v[j] = apply_act(x, act);
    }
    if (pool > 1) {
      // max over groups of `pool` consecutive rows (pool divides 32); rows beyond M
      // never win because whole groups are either inside or outside M.
#pragma unroll
      for (int j = 0; j < 32; ++j)
        for (int o = 1; o < pool; o <<= 1) v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], o));
      if ((lane % pool) == 0 && grow < M) {
        float *o = out + (long long)(grow / pool) * ldo + n0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < N) o[j] = v[j];
      }
    } else if (grow < M) {
      float *o = out + (long long)grow * ldo + n0;
      const bool vst = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (n0 + 32 <= N);
      if (vst && !accumulate) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < N) o[j] = accumulate ? o[j] + v[j] : v[j];
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, BN);
}

}  // namespace msr3d

using namespace msr3d;

MSR3D_API long long msr3d_packed_weight_bytes(int N, int K, int BN, int BK) {
  if (N <= 0 || K <= 0 || BN <= 0 || BK <= 0) return 0;
  const long long Np = (long long)ceil_div(N, BN) * BN, Kp = (long long)ceil_div(K, BK) * BK;
  return Np * Kp * 4;
}

MSR3D_API int msr3d_pack_weight_f32(const float *W, int N, int K, long long stride_n,
                                    long long stride_k, const float *row_scale,
                                    const int32_t *col_perm, int BN, int BK, void *packed,
                                    msr3d_stream_t stream) {
  MSR3D_REQUIRE(W && packed, "pack_weight: null pointer");
  MSR3D_REQUIRE(N > 0 && K > 0, "pack_weight: bad sizes N=%d K=%d", N, K);
  MSR3D_REQUIRE(BN > 0 && BN % 8 == 0 && BK > 0 && BK % 16 == 0, "pack_weight: bad tile BN=%d BK=%d", BN, BK);
  const int Np = ceil_div(N, BN) * BN, Kp = ceil_div(K, BK) * BK;
  const long long total = (long long)Np * (Kp / 8);
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  pack_weight_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, N, K, stride_n, stride_k, row_scale,
                                                            col_perm, BN, BK, Np, Kp,
                                                            (uint8_t *)packed);
  return check_launch("pack_weight");
}

MSR3D_API int msr3d_linear_tc_f32(const float *A, long long stride_am, long long stride_ak, int M,
                                  int K, const void *packed_w, const float *bias, float *out,
                                  long long ldo, int N, int act, int pool, int split,
                                  int accumulate, msr3d_stream_t stream) {
  MSR3D_REQUIRE(A && packed_w && out, "linear_tc: null pointer");
  MSR3D_REQUIRE(M >= 0 && K > 0 && N > 0, "linear_tc: bad sizes M=%d K=%d N=%d", M, K, N);
  MSR3D_REQUIRE(act >= 0 && act <= 2, "linear_tc: bad act %d", act);
  MSR3D_REQUIRE(split == 1 || split == 3, "linear_tc: split must be 1 or 3");
  MSR3D_REQUIRE(pool <= 1 || (pool <= 32 && (pool & (pool - 1)) == 0 && M % pool == 0),
                "linear_tc: pool=%d must be a power of two <= 32 dividing M", pool);
  MSR3D_REQUIRE(!(accumulate && (pool > 1 || act != 0)), "linear_tc: accumulate needs act=0, pool<=1");
  if (M == 0) return MSR3D_OK;
  constexpr int BN = 128;
  const int Kp = ceil_div(K, kChunkK) * kChunkK;
  const size_t smem = 2 * (2 * (size_t)kTileM * kChunkK * 2 + (size_t)BN * kChunkK * 4) + 1024;
  static bool configured = false;
  if (!configured) {
    MSR3D_CUDA(cudaFuncSetAttribute(linear_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    configured = true;
  }
  dim3 grid(ceil_div(M, kTileM), ceil_div(N, BN));
  linear_tc_kernel<BN><<<grid, 256, smem, (cudaStream_t)stream>>>(
      A, stride_am, stride_ak, M, K, (const uint8_t *)packed_w, Kp, bias, out, ldo, N, act, pool,
      split, accumulate);
  return check_launch("linear_tc");
}
