// Dense layers of the frozen PointNet++ encoder (SA3 = GroupAll MLP 259 -> 256 -> 512 -> 768 with the max over
// the 16 points of a cloud, and the final fc 768 -> 768; pointnet2_modules.py:34-75, pytorch_utils.py:157-188,
// modules/layers/pointnet.py:53,63) with a TMA-fed activation operand:
//
//   out[M, N] = act( A[M, K] * W[N, K]^T + bias )       (optionally max over groups of `pool` consecutive rows)
//
// Arithmetic as in linear_tc.cu: bf16 hi/lo split, three tcgen05.mma kind::f16 per K = 16 slice (hh + hl + lh,
// ~16 mantissa bits, fp32 accumulation in TMEM); W is the pre-packed, BatchNorm-folded image (msr3d_pack_weight_f32)
// fetched with one 1-D bulk copy per chunk.  What changed is the A path.  Round 1 read the fp32 activations with
// strided global loads in 16 producer warps and converted them in registers: one chunk in flight per warp, so the
// kernel ran at a third of what its shared-memory traffic allows (3840 x 768 x 512: 37 us).  Here
//   warp 0      requests the K = 64 chunk of A as two [128 x 32 fp32] tensor-map TMA boxes (128 B swizzle) together
//               with the weight block, several chunks ahead;
//   warps 4-11  convert the landed fp32 tile IN PLACE into the swizzled K-major bf16 hi and lo tiles (the 32 KB of
//               fp32 are exactly the 16 + 16 KB of the two bf16 images), split the usual way;
//   warp 1      issues the MMAs (A: SWIZZLE_128B descriptors, B: the packed no-swizzle image), tcgen05.commit
//               frees the stage;
//   epilogue    bias + activation, then either a swizzled staging tile + TMA store (full lines, tails clipped in
//               hardware) or, with pooling, the warp-shuffle row-group max and a direct store of the few rows left.
// No split-K: the frozen encoder stays run-to-run bit-identical (tests/test_prompter_gpu.py).
#include "common.cuh"
#include "rng.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace msr3d {

using namespace umma;

constexpr int kLtM = 128;
constexpr int kLtChunk = 64;           // K elements per stage
constexpr int kLtConvWarps = 8;
constexpr int kLtThreads = 384;
constexpr int kLtFirstConv = 4;
constexpr uint32_t kLtBox = kLtM * tma::kRowBytes;     // 16 KB: one [128 x 32 fp32] box = one [128 x 64 bf16] tile

struct alignas(64) LinTmaParams {
  CUtensorMap map_a, map_out;
  CUtensorMap map_a_hi, map_a_lo;        // PS: the activation operand arrives as bf16 hi / lo tensors [M, K]
  uint16_t *out_hi, *out_lo;             // OS: the result leaves as bf16 hi / lo tensors [M, ldo] (next layer's operand)
  const uint8_t *PW;
  const float *bias;
  float *out;
  long long ldo;
  int M, N, K, nkc;
  int act, pool, tma_out;
};

template <int BN>
struct LinTmaCfg {
  static constexpr uint32_t kRaw = 2 * kLtBox;                       // fp32 chunk; afterwards [bf16 hi | bf16 lo]
  static constexpr uint32_t kBTile = BN * kLtChunk * 2;              // one packed bf16 image of the weight block
  static constexpr uint32_t kB = 2 * kBTile;
  static constexpr uint32_t kStage = kRaw + kB;
  static constexpr int kStages = (200 * 1024) / kStage > 4 ? 4 : (200 * 1024) / kStage;
  static constexpr uint32_t kOutStage = (BN / 32) * kLtBox;          // output staging, aliases the ring
  static_assert(kOutStage <= kStages * kStage, "output staging must fit in the ring");
  static constexpr size_t kSmem = (size_t)kStages * kStage + 1024;
  static constexpr uint32_t kTmemCols = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
};

__device__ __forceinline__ void lt_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float lt_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) {
    float cdf, pdf;
    gelu_cdf_pdf(v, cdf, pdf);
    return v * cdf;
  }
  return v;
}

// PS (pre-split): A is given as two bf16 tensors (hi, lo) written by the previous layer's epilogue — the TMA boxes
//     {64 bf16, 128 rows} land exactly as the swizzled operand images, no conversion pass (r2 timeline: the in-place
//     fp32 -> bf16 hi/lo conversion made SA3's 512 -> 768 layer 75 us on the encoder's SM partition, 3 us of it MMAs);
// OS (output split): the epilogue writes bf16 hi / lo instead of fp32 (for a PS consumer).
template <int BN, bool PS, bool OS>
__global__ void __launch_bounds__(kLtThreads, 1)
linear_tma_kernel(const __grid_constant__ LinTmaParams g) {
  using C = LinTmaCfg<BN>;
  constexpr int S = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[S], conv_done[S], mma_done[S];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * kLtM;
  const int nt = blockIdx.y;
  const int nkc = g.nkc;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv_done[s], kLtConvWarps);
      mbar_init(&mma_done[s], 1);
    }
    mbar_init(&acc_full, 1);
    fence_mbar_init();
    if (PS) {
      tma::prefetch_map(&g.map_a_hi);
      tma::prefetch_map(&g.map_a_lo);
    } else {
      tma::prefetch_map(&g.map_a);
    }
    if (g.tma_out) tma::prefetch_map(&g.map_out);
  }
  if (warp == 2) tmem_alloc(&s_tmem, C::kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = s_tmem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane)
    if (elect_one()) {
      auto load_w = [&](int i) {   // packed weight block (n tile, k chunk): hi image then lo image
        const int s = i % S;
        bulk_g2s(smem + (size_t)s * C::kStage + C::kRaw, g.PW + ((size_t)nt * nkc + i) * C::kB, C::kB, &full[s]);
      };
      auto load_a = [&](int i) {
        const int s = i % S;
        uint8_t *dst = smem + (size_t)s * C::kStage;
        if (PS) {
          tma::load_2d(dst, &g.map_a_hi, i * kLtChunk, m0, &full[s]);
          tma::load_2d(dst + kLtBox, &g.map_a_lo, i * kLtChunk, m0, &full[s]);
        } else {
          tma::load_2d(dst, &g.map_a, i * kLtChunk, m0, &full[s]);
          tma::load_2d(dst + kLtBox, &g.map_a, i * kLtChunk + 32, m0, &full[s]);
        }
      };
      const int pre = min(S, nkc);
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx(&full[i], C::kStage);
        load_w(i);                              // frozen weights: in flight while the previous kernel drains
      }
      pdl_wait();
      for (int i = 0; i < pre; ++i) load_a(i);
      pdl_launch_dependents();
      for (int i = pre; i < nkc; ++i) {
        const int s = i % S;
        mbar_wait(&mma_done[s], ((i / S) - 1) & 1);
        mbar_expect_tx(&full[s], C::kStage);
        load_w(i);
        load_a(i);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp waits, one lane issues)
    const uint32_t idesc = make_idesc_bf16(kLtM, BN);
#pragma unroll 1
    for (int i = 0; i < nkc; ++i) {
      const int s = i % S;
      if (PS) mbar_wait(&full[s], (i / S) & 1);
      else mbar_wait(&conv_done[s], (i / S) & 1);
      fence_after_sync();
      if (elect_one()) {
        const uint32_t base = smem_u32(smem + (size_t)s * C::kStage);
        const uint64_t a_hi = tma::desc_kmajor(base), a_lo = tma::desc_kmajor(base + kLtBox);
        const uint64_t b_hi = make_desc(base + C::kRaw, (uint32_t)BN * 16u, 128u);
        const uint64_t b_lo = make_desc(base + C::kRaw + C::kBTile, (uint32_t)BN * 16u, 128u);
        constexpr uint64_t kStepB = 2 * (uint64_t)BN;       // 16 B units per K = 16 slice of the packed image
#pragma unroll
        for (int ks = 0; ks < kLtChunk / 16; ++ks) {
          mma_bf16(tmem_d, a_hi + ks * 2, b_hi + ks * kStepB, idesc, (i | ks) ? 1u : 0u);
          mma_bf16(tmem_d, a_hi + ks * 2, b_lo + ks * kStepB, idesc, 1u);
          mma_bf16(tmem_d, a_lo + ks * 2, b_hi + ks * kStepB, idesc, 1u);
        }
        mma_commit(&mma_done[s]);
        if (i == nkc - 1) mma_commit(&acc_full);
      }
      __syncwarp();
    }
  } else if (warp >= kLtFirstConv) {
    // ------------------------------------------------------------------ converters, then epilogue
    pdl_wait();
    const int cw = warp - kLtFirstConv;
    const int ct = cw * 32 + lane;
    // this thread converts row `cr`, k half `cb` of every chunk: eight 16 B fp32 chunks in, four bf16 hi + four lo out
    const int cr = ct & 127, cb = ct >> 7;
#pragma unroll 1
    for (int i = 0; i < (PS ? 0 : nkc); ++i) {
      const int s = i % S;
      uint8_t *tile = smem + (size_t)s * C::kStage;
      mbar_wait(&full[s], (i / S) & 1);
      float4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4 *>(tile + cb * kLtBox + tma::sw128_offset(cr, c));
      tma::named_barrier(2, kLtConvWarps * 32);   // in place: every fp32 chunk is in registers before any bf16 chunk lands
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t hi[4], lo[4];
        split_pack2(v[2 * c].x, v[2 * c].y, hi[0], lo[0]);
        split_pack2(v[2 * c].z, v[2 * c].w, hi[1], lo[1]);
        split_pack2(v[2 * c + 1].x, v[2 * c + 1].y, hi[2], lo[2]);
        split_pack2(v[2 * c + 1].z, v[2 * c + 1].w, hi[3], lo[3]);
        const uint32_t o = tma::sw128_offset(cr, 4 * cb + c);
        *reinterpret_cast<uint4 *>(tile + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(tile + kLtBox + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&conv_done[s]);
    }
    // ---- epilogue: warp = (lane quarter q, column half)
    const int q = cw & 3, half = cw >> 2;
    const int r = q * 32 + lane;
    const int grow = m0 + r;
    const int n_base = nt * BN;
    mbar_wait(&acc_full, 0);
    fence_after_sync();
    constexpr int kColsPerHalf = BN / 2;
    if (OS) {
      // bf16 hi / lo rows of the next layer's operand: 16 columns = 32 B of each image per step
      const bool row_ok = grow < g.M;
#pragma unroll 1
      for (int c0 = half * kColsPerHalf; c0 < (half + 1) * kColsPerHalf; c0 += 16) {
        const int n0 = n_base + c0;
        if (n0 >= g.N) break;
        uint32_t acc[16];
        lt_tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
        tmem_ld_wait();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = lt_act(__uint_as_float(acc[2 * j]) + (g.bias != nullptr ? g.bias[n0 + 2 * j] : 0.f), g.act);
          const float b = lt_act(__uint_as_float(acc[2 * j + 1]) + (g.bias != nullptr ? g.bias[n0 + 2 * j + 1] : 0.f), g.act);
          split_pack2(a, b, hi[j], lo[j]);
        }
        if (row_ok) {
          uint4 *ph = reinterpret_cast<uint4 *>(g.out_hi + (size_t)grow * g.ldo + n0);
          uint4 *pl = reinterpret_cast<uint4 *>(g.out_lo + (size_t)grow * g.ldo + n0);
          ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
      }
    } else if (g.pool <= 1 && g.tma_out) {
#pragma unroll 1
      for (int c0 = half * kColsPerHalf; c0 < (half + 1) * kColsPerHalf; c0 += 16) {
        const int n0 = n_base + c0;
        float bv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) bv[j] = (g.bias != nullptr && n0 + j < g.N) ? g.bias[n0 + j] : 0.f;
        uint32_t acc[16];
        lt_tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
        tmem_ld_wait();
        uint8_t *box = smem + (size_t)(c0 >> 5) * kLtBox;
        const int ch0 = (c0 & 31) >> 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 t;
          t.x = lt_act(__uint_as_float(acc[4 * j]) + bv[4 * j], g.act);
          t.y = lt_act(__uint_as_float(acc[4 * j + 1]) + bv[4 * j + 1], g.act);
          t.z = lt_act(__uint_as_float(acc[4 * j + 2]) + bv[4 * j + 2], g.act);
          t.w = lt_act(__uint_as_float(acc[4 * j + 3]) + bv[4 * j + 3], g.act);
          *reinterpret_cast<float4 *>(box + tma::sw128_offset(r, ch0 + j)) = t;
        }
      }
      fence_async_smem();
      tma::named_barrier(1, kLtConvWarps * 32);
      if (cw == 0 && elect_one()) {
#pragma unroll
        for (int b = 0; b < BN / 32; ++b) {
          if (n_base + 32 * b >= g.N) break;
          tma::store_2d(&g.map_out, smem + (size_t)b * kLtBox, n_base + 32 * b, m0);
        }
        tma::store_commit();
        tma::store_wait_read();
      }
      __syncwarp();
    } else {
      // direct stores; pool > 1: max over groups of `pool` consecutive rows (a power of two <= 32 dividing M)
      const bool row_ok = grow < g.M;
#pragma unroll 1
      for (int c0 = half * kColsPerHalf; c0 < (half + 1) * kColsPerHalf; c0 += 16) {
        const int n0 = n_base + c0;
        if (n0 >= g.N) break;
        const int nv = min(16, g.N - n0);
        uint32_t acc[16];
        lt_tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = lt_act(__uint_as_float(acc[j]) + ((g.bias != nullptr && j < nv) ? g.bias[n0 + j] : 0.f), g.act);
        if (g.pool > 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            for (int o = 1; o < g.pool; o <<= 1) v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], o));
          if ((lane % g.pool) == 0 && row_ok) {
            float *o = g.out + (long long)(grow / g.pool) * g.ldo + n0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nv) o[j] = v[j];
          }
        } else if (row_ok) {
          float *o = g.out + (long long)grow * g.ldo + n0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nv) o[j] = v[j];
        }
      }
    }
  } else {
    pdl_sync();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_d, C::kTmemCols);
}

template <int BN, bool PS = false, bool OS = false>
static int launch_linear_tma(const LinTmaParams &g, cudaStream_t stream) {
  using C = LinTmaCfg<BN>;
  static bool configured = false;
  if (!configured) {
    MSR3D_CUDA(cudaFuncSetAttribute(linear_tma_kernel<BN, PS, OS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
    configured = true;
  }
  launch_pdl(linear_tma_kernel<BN, PS, OS>, dim3(ceil_div(g.M, kLtM), ceil_div(g.N, BN)), dim3(kLtThreads), C::kSmem, stream, g);
  return check_launch("linear_tma");
}

// 2-D bf16 row-major tensor [rows][cols], row pitch `ld` elements; box = 64 columns (128 B) x 128 rows, 128 B swizzle
static int make_map_bf16(CUtensorMap *map, const void *base, long long rows, long long cols, long long ld) {
  tma::EncodeTiledFn fn = tma::encode_fn();
  if (fn == nullptr) return set_error(MSR3D_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  MSR3D_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld & 7) == 0 && ld >= cols && rows > 0 && cols > 0,
                "bf16 tensor map: base must be 16 B aligned and the row pitch a multiple of 8 elements");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)kLtM};
  const cuuint32_t estride[2] = {1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(MSR3D_ERR_CUDA, "cuTensorMapEncodeTiled (bf16) failed (%d)", (int)r);
  return MSR3D_OK;
}

// Returns MSR3D_ERR_UNSUPPORTED (without touching the error string of a caller that falls back) when the
// arguments do not fit this kernel: A not K-contiguous / unaligned, unsupported tile, accumulate.
int linear_tma_try(const float *A, long long stride_am, long long stride_ak, int M, int K, const void *packed_w,
                   const float *bias, float *out, long long ldo, int N, int act, int pool, int split, int accumulate,
                   int tile_n, cudaStream_t stream) {
  if (g_tune[TUNE_LINEAR_TMA_OFF] != 0) return MSR3D_ERR_UNSUPPORTED;
  if (stride_ak != 1 || (stride_am & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || stride_am < K) return MSR3D_ERR_UNSUPPORTED;
  if (split != 3 || accumulate != 0) return MSR3D_ERR_UNSUPPORTED;
  const int bn = tile_n == 0 ? 128 : tile_n;
  if (bn != 64 && bn != 128 && bn != 192) return MSR3D_ERR_UNSUPPORTED;
  LinTmaParams g{};
  g.PW = static_cast<const uint8_t *>(packed_w);
  g.bias = bias; g.out = out; g.ldo = ldo;
  g.M = M; g.N = N; g.K = K; g.nkc = ceil_div(K, kLtChunk);
  g.act = act; g.pool = pool;
  int rc = tma::make_map_f32(&g.map_a, A, M, K, stride_am, kLtM, false);
  if (rc != MSR3D_OK) return rc;
  g.tma_out = (pool <= 1 && (ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && ldo >= N) ? 1 : 0;
  if (g.tma_out) {
    rc = tma::make_map_f32(&g.map_out, out, M, N, ldo, kLtM, false);
    if (rc != MSR3D_OK) return rc;
  }
  if (bn == 64) return launch_linear_tma<64>(g, stream);
  if (bn == 128) return launch_linear_tma<128>(g, stream);
  return launch_linear_tma<192>(g, stream);
}

}  // namespace msr3d

using namespace msr3d;

/* see include/msr3d_b200.h */
MSR3D_API int msr3d_linear_bf16x3_f32(const float *A, const void *A_hi, const void *A_lo, long long lda, int M, int K,
                                      const void *packed_w, const float *bias, float *out, void *out_hi, void *out_lo,
                                      long long ldo, int N, int act, int pool, msr3d_stream_t stream) {
  const bool ps = A_hi != nullptr, os = out_hi != nullptr;
  MSR3D_REQUIRE(packed_w && M >= 0 && K > 0 && N > 0, "linear_bf16x3: bad arguments");
  MSR3D_REQUIRE(ps ? (A_lo != nullptr) : (A != nullptr), "linear_bf16x3: activation operand missing");
  MSR3D_REQUIRE(os ? (out_lo != nullptr && pool <= 1 && (ldo & 7) == 0 && (N & 15) == 0 &&
                      ((reinterpret_cast<uintptr_t>(out_hi) | reinterpret_cast<uintptr_t>(out_lo)) & 15) == 0)
                   : (out != nullptr),
                "linear_bf16x3: output (hi / lo outputs need pool <= 1, N %% 16 == 0, ldo %% 8 == 0, 16 B alignment)");
  MSR3D_REQUIRE(pool <= 1 || (pool <= 32 && (pool & (pool - 1)) == 0 && M % pool == 0), "linear_bf16x3: pool=%d", pool);
  if (M == 0) return MSR3D_OK;
  if (!ps && !os)
    return linear_tma_try(A, lda, 1, M, K, packed_w, bias, out, ldo, N, act, pool, 3, 0, 128, (cudaStream_t)stream);
  LinTmaParams g{};
  g.PW = static_cast<const uint8_t *>(packed_w);
  g.bias = bias; g.out = out; g.ldo = ldo;
  g.out_hi = static_cast<uint16_t *>(out_hi); g.out_lo = static_cast<uint16_t *>(out_lo);
  g.M = M; g.N = N; g.K = K; g.nkc = ceil_div(K, kLtChunk);
  g.act = act; g.pool = pool; g.tma_out = 0;
  int rc;
  if (ps) {
    if ((rc = make_map_bf16(&g.map_a_hi, A_hi, M, K, lda)) != MSR3D_OK) return rc;
    if ((rc = make_map_bf16(&g.map_a_lo, A_lo, M, K, lda)) != MSR3D_OK) return rc;
  } else {
    MSR3D_REQUIRE((lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && lda >= K, "linear_bf16x3: A alignment");
    if ((rc = tma::make_map_f32(&g.map_a, A, M, K, lda, kLtM, false)) != MSR3D_OK) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (ps && os) return launch_linear_tma<128, true, true>(g, st);
  if (ps) return launch_linear_tma<128, true, false>(g, st);
  return launch_linear_tma<128, false, true>(g, st);
}
