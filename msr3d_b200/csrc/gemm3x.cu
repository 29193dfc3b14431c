// TMA-fed 3xTF32 GEMMs of the situated spatial-attention transformer (forward, data gradients and
// weight gradients): modules/layers/transformers.py:200-252,314-329 run these contractions through
// cuBLAS fp32; here they are tcgen05.mma kind::tf32 with both operands fetched by tensor-map TMA
// straight from the row-major fp32 tensors (activations, gradients, the flat parameter buffer and
// their transposes — see tma.cuh), fp32-grade through the big/small split done in shared memory.
//
//   out[M, N] (+)= epilogue( A[M, K] * B[N, K]^T + bias ) + residual
//
// One CTA = one 128 x BN output tile (x one K split), 384 threads, warp specialised:
//   warp 0      TMA producer: per 32-wide K chunk one box (K-major) or BN/32 resp. 4 boxes (MN-major) per
//               operand into a ring of stages; weight tiles ("static" operands that no kernel of the
//               current step writes) are requested BEFORE the programmatic-dependent-launch wait, i.e.
//               while the previous kernel of the chain is still running.
//   warps 4-11  converters: split the landed raw tile into big / small tf32 images in place (3 ALU ops
//               per element, no global traffic, no layout knowledge), then run the epilogue.
//   warp 1      MMA issuer: 4 K-steps x 3 products per chunk; tcgen05.commit frees the stage.
//   warp 2      TMEM allocation.
// Compared with linear_tc.cu (round 1: every CTA re-derived three bf16 images of both operands from
// global fp32 in registers, 5.5 instructions per element, 16 producer warps) the operand path costs no
// global-load instructions, no address arithmetic and a third of the ALU work.
#include "common.cuh"
#include "rng.cuh"
#include "tma.cuh"
#include "umma.cuh"

#include <mutex>

namespace msr3d {

using namespace umma;

namespace tma {
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_map_f32(CUtensorMap *map, const float *base, long long rows, long long cols, long long ld, int box_rows,
                 bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return set_error(MSR3D_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  MSR3D_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld & 3) == 0 && ld >= cols && rows > 0 && cols > 0,
                "tensor map: base must be 16 B aligned and the row pitch a multiple of 4 floats (rows=%lld cols=%lld ld=%lld)",
                rows, cols, ld);
  MSR3D_REQUIRE(box_rows >= 1 && box_rows <= 256, "tensor map: box rows %d", box_rows);
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(MSR3D_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return MSR3D_OK;
}
}  // namespace tma

constexpr int kTileM3 = 128;
constexpr int kChunk3 = 32;            // K elements per stage: one 128 B swizzle row
constexpr int kConvWarps = 8;          // converter / epilogue warps (warps 4..11)
constexpr int kThreads3 = 384;
constexpr int kFirstConvWarp = 4;

enum { EPI3_PLAIN = 0, EPI3_FFN1 = 2, EPI3_MUL_AUX = 3 };

struct alignas(64) Gemm3xParams {
  CUtensorMap map_a, map_b, map_out;
  int tma_out;                    // output goes through shared memory + TMA store / reduce-add (map_out valid)
  int a_mn, b_mn;                 // operand orientation (1 = MN-major, i.e. stored [k][mn])
  int b_static;                   // B is not written by any kernel of the current step (weights)
  int M, N, K;
  const float *bias;
  const float *res; long long ldr;
  const float *aux; long long ldaux;
  float *aux_out;
  float *out; long long ldo;
  int accumulate, epi;
  unsigned epi_site;
  Dropout drop;
};

__device__ __forceinline__ void tmem_ld16_3x(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void load16(const float *p, bool vec, int nv, float (&d)[16]) {
  if (vec) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 t = *reinterpret_cast<const float4 *>(p + j);
      d[j] = t.x; d[j + 1] = t.y; d[j + 2] = t.z; d[j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) d[j] = j < nv ? p[j] : 0.f;
  }
}
__device__ __forceinline__ void red_add4(float *p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Epilogue through shared memory: TMEM -> registers (bias, * aux, + residual) -> swizzled [128 x 32] boxes in the
// (now idle) operand ring -> ONE TMA store / reduce-add per box.  A thread owns a TMEM lane = an output ROW, so
// direct global stores touch 32 different 128 B lines per warp instruction (half a sector each): the round-1
// style epilogue below cost 1.2 us per 128 x 64 tile and 4.2 us per 128 x 256 tile of red.add
// (tools/trace_kernels.py); the bulk tensor copy writes full lines and clips the M / N tails in hardware.
template <int BN>
__device__ __forceinline__ void epilogue3x_tma(const Gemm3xParams &g, uint32_t tmem_d, uint8_t *stage, int m0, int n_base,
                                               int cw, int lane, uint64_t *acc_bar) {
  const int q = cw & 3, half = cw >> 2;
  const int r = q * 32 + lane;
  const int grow = m0 + r;
  const bool row_ok = grow < g.M;
  const bool first_split = blockIdx.z == 0;
  const bool use_bias = g.bias != nullptr && first_split;
  const bool use_res = g.res != nullptr && first_split && row_ok;
  const bool use_aux = g.epi == EPI3_MUL_AUX && row_ok;
  const bool res_vec = ((g.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.res) & 15) == 0);
  const bool aux_vec = ((g.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.aux) & 15) == 0);
  const bool bias_vec = (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0;
  constexpr int kSteps = BN / 32;          // 16-column steps per warp
  constexpr int kGroup = 2;                // steps whose global operands are in flight together
  bool waited = false;
#pragma unroll
  for (int st0 = 0; st0 < kSteps; st0 += kGroup) {
    // the global operands of a group of steps are requested together — those of the first group before the
    // wait for the accumulator
    float bv[kGroup][16], rv[kGroup][16], av[kGroup][16];
#pragma unroll
    for (int u = 0; u < kGroup; ++u) {
      const int n0 = n_base + half * (BN / 2) + (st0 + u) * 16;
      const int nv = min(16, g.N - n0);
      if (nv <= 0) continue;
      const bool full = nv == 16;
      if (use_bias) load16(g.bias + n0, full && bias_vec, nv, bv[u]);
      if (use_res) load16(g.res + (long long)grow * g.ldr + n0, full && res_vec, nv, rv[u]);
      if (use_aux) load16(g.aux + (long long)grow * g.ldaux + n0, full && aux_vec, nv, av[u]);
    }
    if (!waited) {
      mbar_wait(acc_bar, 0);
      fence_after_sync();
      waited = true;
      if (cw == 0 && lane == 0) trace_mark(6);
    }
#pragma unroll
    for (int u = 0; u < kGroup; ++u) {
      const int c0 = half * (BN / 2) + (st0 + u) * 16;
      uint32_t acc[16];
      tmem_ld16_3x(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] = __uint_as_float(acc[j]);
        if (use_bias) v[j] += bv[u][j];
        if (use_aux) v[j] *= av[u][j];
        if (use_res) v[j] += rv[u][j];
      }
      uint8_t *box = stage + (size_t)(c0 >> 5) * (kTileM3 * tma::kRowBytes);
      const int ch0 = (c0 & 31) >> 2;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4 *>(box + tma::sw128_offset(r, ch0 + j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  fence_async_smem();                                   // generic-proxy tile -> visible to the TMA engine
  tma::named_barrier(1, kConvWarps * 32);
  if (cw == 0 && tma::elect_one()) {
#pragma unroll
    for (int b = 0; b < BN / 32; ++b) {
      if (n_base + 32 * b >= g.N) break;
      const uint8_t *box = stage + (size_t)b * (kTileM3 * tma::kRowBytes);
      if (g.accumulate == 2) tma::reduce_add_2d(&g.map_out, box, n_base + 32 * b, m0);
      else tma::store_2d(&g.map_out, box, n_base + 32 * b, m0);
    }
    tma::store_commit();
    tma::store_wait_read();
  }
  __syncwarp();
}

// Direct epilogue (fallback: unaligned outputs, the dual-output FFN1 epilogue): 8 warps = 4 lane quarters x 2 column
// halves, 16 columns per step, row-per-thread global accesses.
template <int BN>
__device__ __noinline__ void epilogue3x(const Gemm3xParams &g, uint32_t tmem_d, int m0, int n_base, int cw, int lane,
                                        uint64_t *acc_bar) {
  const int q = cw & 3, half = cw >> 2;
  const int grow = m0 + q * 32 + lane;
  const bool row_ok = grow < g.M;
  const bool first_split = blockIdx.z == 0;
  const bool use_bias = g.bias != nullptr && first_split;
  const bool use_res = g.res != nullptr && first_split && row_ok;
  const bool use_aux = g.epi == EPI3_MUL_AUX && row_ok;
  const bool bias_vec = (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0;
  const bool res_vec = ((g.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.res) & 15) == 0);
  const bool aux_vec = ((g.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.aux) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(g.aux_out) & 15) == 0);
  const bool out_vec = ((g.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.out) & 15) == 0);
  Dropout drop = g.drop;
  if (g.epi == EPI3_FFN1) drop.resolve();
  bool waited = false;
#pragma unroll 1
  for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 16) {
    const int n0 = n_base + c0;
    if (n0 >= g.N) break;   // warp-uniform
    const int nv = min(16, g.N - n0);
    const bool full = nv == 16;
    float bv[16], rv[16], av[16];
    if (use_bias) load16(g.bias + n0, full && bias_vec, nv, bv);
    if (use_res) load16(g.res + (long long)grow * g.ldr + n0, full && res_vec, nv, rv);
    if (use_aux) load16(g.aux + (long long)grow * g.ldaux + n0, full && aux_vec, nv, av);
    if (!waited) {   // accumulator complete (the global operands of the first step are already in flight)
      mbar_wait(acc_bar, 0);
      fence_after_sync();
      waited = true;
      if (cw == 0 && lane == 0) trace_mark(6);
    }
    uint32_t r[16];
    tmem_ld16_3x(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
    tmem_ld_wait();
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + (use_bias ? bv[j] : 0.f);
    if (g.epi == EPI3_FFN1) {
      // out = dropout(gelu(x)), aux_out = gelu'(x) * mask   (requires N % 16 == 0, 16 B aligned rows)
      if (!row_ok) continue;
      const long long off = (long long)grow * g.ldaux + n0;
      float hh[16], gp[16];
#pragma unroll
      for (int j8 = 0; j8 < 16; j8 += 8) {
        float m[8];
        drop.mask8((uint64_t)(off + j8), g.epi_site, m);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x = v[j8 + j];
          float cdf, pdf;
          gelu_cdf_pdf(x, cdf, pdf);
          hh[j8 + j] = x * cdf * m[j];
          gp[j8 + j] = fmaf(x, pdf, cdf) * m[j];
        }
      }
      float *o = g.out + (long long)grow * g.ldo + n0;
      float *ao = g.aux_out + off;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        *reinterpret_cast<float4 *>(o + j) = make_float4(hh[j], hh[j + 1], hh[j + 2], hh[j + 3]);
        *reinterpret_cast<float4 *>(ao + j) = make_float4(gp[j], gp[j + 1], gp[j + 2], gp[j + 3]);
      }
      continue;
    }
    if (use_aux) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= av[j];
    }
    if (use_res) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += rv[j];
    }
    if (!row_ok) continue;
    float *o = g.out + (long long)grow * g.ldo + n0;
    if (g.accumulate == 0 && full && out_vec) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else if (g.accumulate == 2 && full && out_vec) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) red_add4(o + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else if (g.accumulate == 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nv) atomicAdd(o + j, v[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nv) o[j] = v[j];
    }
  }
}

template <int BN>
struct Gemm3xCfg {
  static constexpr uint32_t kABytes = kTileM3 * tma::kRowBytes;          // 16 KB: one image of the A chunk
  static constexpr uint32_t kBBytes = BN * tma::kRowBytes;
  static constexpr uint32_t kBigBytes = kABytes + kBBytes;               // TMA lands here; = small-image offset
  static constexpr uint32_t kStageBytes = 2 * kBigBytes;                 // [A big | B big | A small | B small]
  static constexpr int kStages = (200 * 1024) / kStageBytes >= 6 ? 6 : (200 * 1024) / kStageBytes;
  static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024;  // + alignment slack
  static constexpr uint32_t kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int kF4 = kBigBytes / 16;                             // float4s per stage to split
  static constexpr int kPer = kF4 / (kConvWarps * 32);
  static_assert(kF4 % (kConvWarps * 32) == 0, "stage must divide over the converter threads");
};

template <int BN>
__global__ void __launch_bounds__(kThreads3, 1)
gemm3x_kernel(const __grid_constant__ Gemm3xParams g) {
  using C = Gemm3xCfg<BN>;
  constexpr int S = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[S];       // TMA bytes landed
  __shared__ __align__(8) uint64_t conv_done[S];  // big / small images written (one arrival per converter warp)
  __shared__ __align__(8) uint64_t mma_done[S];   // tensor core finished reading the stage
  __shared__ __align__(8) uint64_t acc_full;      // accumulator complete
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * kTileM3;
  const int n0 = blockIdx.y * BN;
  const int nkc_total = (g.K + kChunk3 - 1) / kChunk3;
  const int per = (nkc_total + gridDim.z - 1) / gridDim.z;
  const int kc_begin = blockIdx.z * per;
  const int nkc = min(nkc_total, kc_begin + per) - kc_begin;
  if (nkc <= 0) { pdl_sync(); return; }   // uniform per CTA, nothing allocated yet

  if (tid == 0) {
    trace_mark(0);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv_done[s], kConvWarps);
      mbar_init(&mma_done[s], 1);
    }
    mbar_init(&acc_full, 1);
    fence_mbar_init();
    tma::prefetch_map(&g.map_a);
    tma::prefetch_map(&g.map_b);
    if (g.tma_out) tma::prefetch_map(&g.map_out);
  }
  if (warp == 2) tmem_alloc(&s_tmem, C::kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = s_tmem;
  if (tid == 0) trace_mark(1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane)
    if (tma::elect_one()) {
      auto load_a = [&](int i) {
        const int s = i % S;
        uint8_t *dst = smem + (size_t)s * C::kStageBytes;
        const int k0 = (kc_begin + i) * kChunk3;
        if (!g.a_mn) tma::load_2d(dst, &g.map_a, k0, m0, &full[s]);
        else {
#pragma unroll
          for (int b = 0; b < kTileM3 / 32; ++b) tma::load_2d(dst + b * tma::kMnBoxBytes, &g.map_a, m0 + 32 * b, k0, &full[s]);
        }
      };
      auto load_b = [&](int i) {
        const int s = i % S;
        uint8_t *dst = smem + (size_t)s * C::kStageBytes + C::kABytes;
        const int k0 = (kc_begin + i) * kChunk3;
        if (!g.b_mn) tma::load_2d(dst, &g.map_b, k0, n0, &full[s]);
        else {
#pragma unroll
          for (int b = 0; b < BN / 32; ++b) tma::load_2d(dst + b * tma::kMnBoxBytes, &g.map_b, n0 + 32 * b, k0, &full[s]);
        }
      };
      const int pre = min(S, nkc);
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx(&full[i], C::kBigBytes);
        if (g.b_static) load_b(i);            // weights: in flight while the previous kernel drains
      }
      pdl_wait();
      trace_mark(2);
      for (int i = 0; i < pre; ++i) {
        if (!g.b_static) load_b(i);
        load_a(i);
      }
      pdl_launch_dependents();
      for (int i = pre; i < nkc; ++i) {
        const int s = i % S;
        mbar_wait(&mma_done[s], ((i / S) - 1) & 1);
        mbar_expect_tx(&full[s], C::kBigBytes);
        load_b(i);
        load_a(i);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp waits, one lane issues)
    {
      const uint32_t idesc = tma::make_idesc_tf32(kTileM3, BN, g.a_mn != 0, g.b_mn != 0);
      const uint64_t step_a = tma::kstep_units(g.a_mn != 0), step_b = tma::kstep_units(g.b_mn != 0);
#pragma unroll 1
      for (int i = 0; i < nkc; ++i) {
        const int s = i % S;
        mbar_wait(&conv_done[s], (i / S) & 1);
        fence_after_sync();
        if (tma::elect_one()) {
          const uint32_t base = smem_u32(smem + (size_t)s * C::kStageBytes);
          const uint64_t a_big = g.a_mn ? tma::desc_mnmajor(base) : tma::desc_kmajor(base);
          const uint64_t b_big = g.b_mn ? tma::desc_mnmajor(base + C::kABytes) : tma::desc_kmajor(base + C::kABytes);
          constexpr uint64_t small = C::kBigBytes >> 4;
#pragma unroll
          for (int ks = 0; ks < kChunk3 / 8; ++ks)
            tma::mma_3x(tmem_d, a_big + ks * step_a, a_big + small + ks * step_a, b_big + ks * step_b,
                        b_big + small + ks * step_b, idesc, (i | ks) ? 1u : 0u);
          mma_commit(&mma_done[s]);
          if (i == nkc - 1) {
            trace_mark(5);
            mma_commit(&acc_full);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= kFirstConvWarp) {
    // ------------------------------------------------------------------ converters, then epilogue
    pdl_wait();
    const int cw = warp - kFirstConvWarp;
    const int ct = cw * 32 + lane;
#pragma unroll 1
    for (int i = 0; i < nkc; ++i) {
      const int s = i % S;
      uint4 *tile = reinterpret_cast<uint4 *>(smem + (size_t)s * C::kStageBytes);
      mbar_wait(&full[s], (i / S) & 1);
      if (i == 0 && ct == 0) trace_mark(3);
      float4 v[C::kPer];
#pragma unroll
      for (int j = 0; j < C::kPer; ++j) v[j] = *reinterpret_cast<const float4 *>(tile + ct + j * (kConvWarps * 32));
#pragma unroll
      for (int j = 0; j < C::kPer; ++j) {
        uint4 b, sm;
        tma::split_tf32(v[j].x, b.x, sm.x);
        tma::split_tf32(v[j].y, b.y, sm.y);
        tma::split_tf32(v[j].z, b.z, sm.z);
        tma::split_tf32(v[j].w, b.w, sm.w);
        tile[ct + j * (kConvWarps * 32)] = b;
        tile[ct + j * (kConvWarps * 32) + C::kBigBytes / 16] = sm;
      }
      fence_async_smem();     // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&conv_done[s]);
    }
    if (ct == 0) trace_mark(4);
    // the ring is idle once the accumulator is complete: stage 0 doubles as the output staging tile
    if (g.tma_out) epilogue3x_tma<BN>(g, tmem_d, smem, m0, n0, cw, lane, &acc_full);
    else epilogue3x<BN>(g, tmem_d, m0, n0, cw, lane, &acc_full);
    if (ct == 0) trace_mark(11);
  } else {
    pdl_sync();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_d, C::kTmemCols);
}

template <int BN>
static int launch3x(const Gemm3xParams &g, int ksplit, cudaStream_t stream) {
  using C = Gemm3xCfg<BN>;
  static bool configured = false;
  if (!configured) {
    MSR3D_CUDA(cudaFuncSetAttribute(gemm3x_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
    configured = true;
  }
  dim3 grid(ceil_div(g.M, kTileM3), ceil_div(g.N, BN), ksplit);
  launch_pdl(gemm3x_kernel<BN>, grid, dim3(kThreads3), C::kSmem, stream, g);
  return check_launch("gemm3x");
}

// why a descriptor cannot run on this kernel (nullptr = it can)
static const char *gemm3x_why_not(const msr3d_gemm_desc *d) {
  if (d == nullptr || !d->A || !d->B || !d->out) return "null pointer";
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return "empty problem";
  const bool a_k = d->stride_ak == 1, a_mn = !a_k && d->stride_am == 1;
  const bool b_k = d->stride_bk == 1, b_mn = !b_k && d->stride_bn == 1;
  if (!(a_k || a_mn) || !(b_k || b_mn)) return "an operand has no unit stride";
  const long long lda = a_k ? d->stride_am : d->stride_ak, ldb = b_k ? d->stride_bn : d->stride_bk;
  if ((lda & 3) || (ldb & 3)) return "operand row pitch is not a multiple of 4 floats";
  if ((reinterpret_cast<uintptr_t>(d->A) & 15) || (reinterpret_cast<uintptr_t>(d->B) & 15)) return "operand base not 16 B aligned";
  if (lda < (a_k ? d->K : d->M) || ldb < (b_k ? d->K : d->N)) return "row pitch smaller than the row";
  if (d->act != 0) return "activation (use epilogue 2)";
  if (d->epilogue != 0 && d->epilogue != 2 && d->epilogue != 3) return "epilogue";
  if (d->accumulate == 1) return "accumulate 1";
  if (d->row_map != nullptr || d->out_dtype != 0) return "row_map / typed output";
  if (d->ksplit > 1 && !(d->accumulate == 2 && d->epilogue == 0)) return "split-K needs accumulate 2 and a plain epilogue";
  if (d->epilogue == 2 && !(d->aux_out && d->N % 16 == 0 && d->ld_out % 4 == 0 && d->ld_aux % 4 == 0 && d->accumulate == 0 &&
                            !d->residual && (reinterpret_cast<uintptr_t>(d->out) % 16 == 0) &&
                            (reinterpret_cast<uintptr_t>(d->aux_out) % 16 == 0) && (reinterpret_cast<uintptr_t>(d->bias) % 16 == 0)))
    return "epilogue 2 alignment";
  if (d->epilogue == 3 && !d->aux) return "epilogue 3 needs aux";
  return nullptr;
}

}  // namespace msr3d

using namespace msr3d;

/* Development hook (tools/trace_kernels.py): device buffer of 16 int64 SM-clock slots per CTA that the gemm3x / ffn_pair
 * kernels fill at their phase boundaries; NULL (default) disables tracing. */
namespace msr3d { int set_trace_spatial(long long *buf); int set_trace_attn_block(long long *buf); }
MSR3D_API int msr3d_set_trace_buffer(long long *buf) {
  MSR3D_CUDA(cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)));
  const int rc = set_trace_attn_block(buf);
  if (rc != MSR3D_OK) return rc;
  return set_trace_spatial(buf);
}

/* 1 if msr3d_gemm_tf32x3_f32 accepts this descriptor, else 0 (msr3d_last_error() says why). */
MSR3D_API int msr3d_gemm_tf32x3_supported(const msr3d_gemm_desc *d) {
  const char *why = gemm3x_why_not(d);
  if (why != nullptr) { set_error(MSR3D_ERR_UNSUPPORTED, "gemm_tf32x3: %s", why); return 0; }
  return 1;
}

MSR3D_API int msr3d_gemm_tf32x3_f32(const msr3d_gemm_desc *d, int b_static, msr3d_stream_t stream) {
  const char *why = gemm3x_why_not(d);
  if (why != nullptr) return set_error(MSR3D_ERR_UNSUPPORTED, "gemm_tf32x3: %s", why);
  MSR3D_REQUIRE(d->dropout_p >= 0.f && d->dropout_p < 1.f, "gemm_tf32x3: bad dropout p");
  Gemm3xParams g{};
  g.a_mn = d->stride_ak == 1 ? 0 : 1;
  g.b_mn = d->stride_bk == 1 ? 0 : 1;
  g.b_static = b_static ? 1 : 0;
  g.M = d->M; g.N = d->N; g.K = d->K;
  int bn = d->tile_n;
  if (bn != 64 && bn != 128) {
    // few tokens (M ~ 240): narrow tiles while SMs are idle
    const int ks = d->ksplit < 1 ? 1 : d->ksplit;
    bn = (2 * ceil_div(d->M, kTileM3) * ceil_div(d->N, 128) * ks <= kNumSMs) ? 64 : 128;
  }
  int rc;
  if (!g.a_mn) rc = tma::make_map_f32(&g.map_a, d->A, d->M, d->K, d->stride_am, kTileM3, false);
  else rc = tma::make_map_f32(&g.map_a, d->A, d->K, d->M, d->stride_ak, tma::kBoxKRows, true);
  if (rc != MSR3D_OK) return rc;
  if (!g.b_mn) rc = tma::make_map_f32(&g.map_b, d->B, d->N, d->K, d->stride_bn, bn, false);
  else rc = tma::make_map_f32(&g.map_b, d->B, d->K, d->N, d->stride_bk, tma::kBoxKRows, true);
  if (rc != MSR3D_OK) return rc;
  g.bias = d->bias; g.res = d->residual; g.ldr = d->ld_residual;
  g.aux = d->aux; g.ldaux = d->ld_aux; g.aux_out = d->aux_out;
  g.out = d->out; g.ldo = d->ld_out;
  g.tma_out = (d->epilogue != EPI3_FFN1 && (d->ld_out & 3) == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 &&
               d->ld_out >= d->N && g_tune[TUNE_GEMM3X_DIRECT_EPI] == 0) ? 1 : 0;
  if (g.tma_out) {
    rc = tma::make_map_f32(&g.map_out, d->out, d->M, d->N, d->ld_out, kTileM3, false);
    if (rc != MSR3D_OK) return rc;
  }
  g.accumulate = d->accumulate; g.epi = d->epilogue; g.epi_site = d->epi_site;
  g.drop = make_dropout(d->seed, d->dropout_p, d->seed_ptr);
  const int nkc = ceil_div(d->K, kChunk3);
  int ksplit = d->ksplit < 1 ? 1 : d->ksplit;
  if (ksplit > nkc) ksplit = nkc;
  return bn == 64 ? launch3x<64>(g, ksplit, (cudaStream_t)stream) : launch3x<128>(g, ksplit, (cudaStream_t)stream);
}

namespace msr3d {
using namespace umma;

// =====================================================================================================
// Fused FFN pair:  out2[M, 256] += epi1( A[M, 256] * B1_slice^T ) * B2_slice^T      per (M tile, F slice)
//
//   forward  (transformers.py:326-327):  h = dropout(gelu(y2 W1^T + b1)),  f += h W2^T (+ b2)
//            A = y2, B1 = W1[slice, :] (K-major), B2 = W2[:, slice] (K-major); h and gelu'(u)*mask are
//            also stored for the backward pass.
//   backward:                            du = (df W2) * gelu'(u)*mask,      dy2 += du W1
//            A = df, B1 = W2[:, slice] as [k][n] (MN-major), B2 = W1[slice, :] as [k][n] (MN-major);
//            du is also stored (weight gradients of linear1 contract it over the tokens).
// The 128 x FS hidden tile never leaves the SM between the two GEMMs: epilogue 1 writes it as the big /
// small tf32 images of the second GEMM's A operand (swizzled K-major tiles) while TMA fetches the B2
// slice into the shared memory the first GEMM's ring just released.  ALL global traffic is bulk tensor
// copies: operands and the stored gelu'*mask tile arrive by TMA load; h / gelu'*mask / du leave as TMA
// stores of swizzled staging tiles; the per-slice partial sums of the second GEMM are reduced into the
// zeroed (forward) / pre-filled (backward: residual-path gradient) output with TMA reduce-add
// (cp.reduce.async.bulk.tensor, SASS UTMAREDG).  A thread owns a TMEM lane = a ROW of the tile, so direct
// global accesses from the epilogues touch 32 different lines per warp instruction: in the first version of
// this kernel the two epilogues were 11 of its 23 us (tools/trace_kernels.py).
// Replaces two GEMM launches and a 2 MB round trip of the hidden activation per layer and direction.
// =====================================================================================================
constexpr int kHid = 256;   // d_model of the spatial encoder (configs/msr3d.yaml:181)

struct alignas(64) FfnPairParams {
  CUtensorMap map_a, map_b1, map_b2;
  CUtensorMap map_a_s, map_b1_s, map_b2_s;   // pre-split variant: the small images (map_a / b1 / b2 = the big ones)
  CUtensorMap map_h;       // [M, F] forward: h (store); backward: du (store)
  CUtensorMap map_gp;      // [M, F] forward: gelu'(u) * mask (store); backward: the same tensor (load)
  CUtensorMap map_out2;    // [M, 256] reduce-add
  int M, F;
  const float *bias1;      // forward: linear1.bias [F]
  const float *bias2;      // forward: linear2.bias [256], added by the slice-0 CTAs
  unsigned site;
  Dropout drop;
};

template <int FS>
struct FfnPairCfg {
  static constexpr int kNch2 = FS / 32;                                        // K chunks of GEMM 2
  static constexpr uint32_t kBox = kTileM3 * tma::kRowBytes;                   // 16 KB: one 128 x 32 fp32 box
  static constexpr uint32_t kABytes = kBox;
  static constexpr uint32_t kB1Bytes = FS * tma::kRowBytes;
  static constexpr uint32_t kBig1 = kABytes + kB1Bytes;
  static constexpr uint32_t kStage1 = 2 * kBig1;
  static constexpr uint32_t kHTile = kBox;                                     // one 128 x 32 image of h
  static constexpr uint32_t kHBytes = 2 * kNch2 * kHTile;                      // [big chunks | small chunks]
  static constexpr uint32_t kB2Chunk = kHid * tma::kRowBytes;                  // 32 KB: 256 x 32
  static constexpr uint32_t kB2Big = kNch2 * kB2Chunk;
  static constexpr uint32_t kB2Bytes = 2 * kB2Big;
  static constexpr int kStages = 3;
  static constexpr uint32_t kRing = kStages * kStage1;
  static constexpr uint32_t kOut2 = (kHid / 32) * kBox;                        // 128 KB: staging of the second accumulator
  static constexpr uint32_t kU0 = kRing > kB2Bytes ? kRing : kB2Bytes;
  static constexpr uint32_t kUnion = kU0 > kOut2 ? kU0 : kOut2;                // GEMM-1 ring, then B2, then output staging
  static constexpr uint32_t kStageH = kB2Big;                                  // h / gelu' staging = the B2-small region
  static_assert(2 * kNch2 * kBox <= kB2Big, "hidden-tile staging must fit in the B2-small region");
  static constexpr size_t kSmem = (size_t)kHBytes + kUnion + 1024;
  static constexpr int kF4 = kBig1 / 16;
  static constexpr int kPer = kF4 / (kConvWarps * 32);
  static_assert(kF4 % (kConvWarps * 32) == 0, "stage must divide over the converter threads");
  static constexpr int kB2F4 = kB2Big / 16;
  static constexpr int kB2Per = kB2F4 / (kConvWarps * 32);
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kAcc2Col = 64;
};

// PS = operands arrive PRE-SPLIT (big / small tf32 images written once by their producers: the LayerNorm tails for the
// activations, msr3d_split_tf32_f32 for the weights): no conversion pass at all — round 2's trace showed the in-kernel
// split of the 128 x 256 activation tile, redone by every CTA of an M tile, as 5 of the kernel's 14 us.
template <int FS, bool BWD, bool PS>
__global__ void __launch_bounds__(kThreads3, 1)
ffn_pair_kernel(const __grid_constant__ FfnPairParams g) {
  using C = FfnPairCfg<FS>;
  constexpr int S = C::kStages;
  constexpr int kNkc1 = kHid / kChunk3;   // 8 chunks of GEMM 1
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *hbuf = smem;                    // A operand of GEMM 2: [big chunks | small chunks]
  uint8_t *ring = smem + C::kHBytes;       // stages of GEMM 1, later the B2 slice, later the output staging
  __shared__ __align__(8) uint64_t full[S], conv_done[S], mma_done[S];
  __shared__ __align__(8) uint64_t acc1_full, h_ready, b2_full, b2_conv, acc2_full, gp_full, b2s_full;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * kTileM3;
  const int f0 = blockIdx.y * FS;

  if (tid == 0) {
    trace_mark(0);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv_done[s], kConvWarps);
      mbar_init(&mma_done[s], 1);
    }
    mbar_init(&acc1_full, 1);
    mbar_init(&h_ready, kConvWarps);
    mbar_init(&b2_full, 1);
    mbar_init(&b2_conv, kConvWarps);
    mbar_init(&acc2_full, 1);
    mbar_init(&gp_full, 1);
    mbar_init(&b2s_full, 1);
    fence_mbar_init();
    tma::prefetch_map(&g.map_a);
    tma::prefetch_map(&g.map_b1);
    tma::prefetch_map(&g.map_b2);
    if (PS) {
      tma::prefetch_map(&g.map_a_s);
      tma::prefetch_map(&g.map_b1_s);
      tma::prefetch_map(&g.map_b2_s);
    }
    tma::prefetch_map(&g.map_h);
    tma::prefetch_map(&g.map_gp);
    tma::prefetch_map(&g.map_out2);
  }
  if (warp == 2) tmem_alloc(&s_tmem, C::kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem1 = s_tmem, tmem2 = s_tmem + C::kAcc2Col;
  if (tid == 0) trace_mark(1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane)
    if (tma::elect_one()) {
      auto load_b1 = [&](int i) {
        const int s = i % S;
        const int k0 = i * kChunk3;
#pragma unroll
        for (int img = 0; img < (PS ? 2 : 1); ++img) {
          uint8_t *dst = ring + (size_t)s * C::kStage1 + C::kABytes + (size_t)img * C::kBig1;
          const CUtensorMap *mp = img ? &g.map_b1_s : &g.map_b1;
          if (!BWD) tma::load_2d(dst, mp, k0, f0, &full[s]);                    // W1[f0 + r][k0 ..]
          else {
#pragma unroll
            for (int b = 0; b < FS / 32; ++b) tma::load_2d(dst + b * tma::kMnBoxBytes, mp, f0 + 32 * b, k0, &full[s]);   // W2[k0 + r][f ..]
          }
        }
      };
      auto load_a = [&](int i) {
        const int s = i % S;
        tma::load_2d(ring + (size_t)s * C::kStage1, &g.map_a, i * kChunk3, m0, &full[s]);
        if (PS) tma::load_2d(ring + (size_t)s * C::kStage1 + C::kBig1, &g.map_a_s, i * kChunk3, m0, &full[s]);
      };
      constexpr uint32_t kTx1 = PS ? C::kStage1 : C::kBig1;
      for (int i = 0; i < S; ++i) {
        mbar_expect_tx(&full[i], kTx1);
        load_b1(i);                              // weights: requested while the previous kernel drains
      }
      if (BWD) {
        // gelu'(u) * mask of this tile (written by the forward pass of this step, long complete): lands in the
        // "small" half of the hidden-tile buffer; every epilogue thread reads its own elements before it
        // overwrites them with the small image of du
        mbar_expect_tx(&gp_full, C::kNch2 * C::kBox);
#pragma unroll
        for (int c = 0; c < C::kNch2; ++c)
          tma::load_2d(hbuf + (size_t)(C::kNch2 + c) * C::kHTile, &g.map_gp, f0 + 32 * c, m0, &gp_full);
      }
      pdl_wait();
      trace_mark(2);
      for (int i = 0; i < S; ++i) load_a(i);
      pdl_launch_dependents();
      for (int i = S; i < kNkc1; ++i) {
        const int s = i % S;
        mbar_wait(&mma_done[s], ((i / S) - 1) & 1);
        mbar_expect_tx(&full[s], kTx1);
        load_b1(i);
        load_a(i);
      }
      // GEMM 1 has released the ring: fetch the B2 slice into it
      mbar_wait(&acc1_full, 0);
      mbar_expect_tx(&b2_full, C::kB2Big);
#pragma unroll
      for (int c = 0; c < C::kNch2; ++c) {
        uint8_t *dst = ring + (size_t)c * C::kB2Chunk;
        const int k0 = f0 + c * 32;
        if (!BWD) tma::load_2d(dst, &g.map_b2, k0, 0, &b2_full);                     // W2[n][k0 ..], 256 rows
        else {
#pragma unroll
          for (int b = 0; b < kHid / 32; ++b) tma::load_2d(dst + b * tma::kMnBoxBytes, &g.map_b2, 32 * b, k0, &b2_full);      // W1[k0 + r][n ..]
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp waits, one lane issues)
    {
      const uint32_t idesc1 = tma::make_idesc_tf32(kTileM3, FS, false, BWD);
      const uint64_t step_b1 = tma::kstep_units(BWD);
#pragma unroll 1
      for (int i = 0; i < kNkc1; ++i) {
        const int s = i % S;
        if (PS) mbar_wait(&full[s], (i / S) & 1);       // both images arrive by TMA
        else mbar_wait(&conv_done[s], (i / S) & 1);
        fence_after_sync();
        if (tma::elect_one()) {
          const uint32_t base = smem_u32(ring + (size_t)s * C::kStage1);
          const uint64_t a_big = tma::desc_kmajor(base);
          const uint64_t b_big = BWD ? tma::desc_mnmajor(base + C::kABytes) : tma::desc_kmajor(base + C::kABytes);
          constexpr uint64_t small = C::kBig1 >> 4;
#pragma unroll
          for (int ks = 0; ks < kChunk3 / 8; ++ks)
            tma::mma_3x(tmem1, a_big + ks * 2, a_big + small + ks * 2, b_big + ks * step_b1, b_big + small + ks * step_b1,
                        idesc1, (i | ks) ? 1u : 0u);
          mma_commit(&mma_done[s]);
          if (i == kNkc1 - 1) {
            trace_mark(5);
            mma_commit(&acc1_full);
          }
        }
        __syncwarp();
      }
      // GEMM 2: A = h images written by epilogue 1, B = the converted B2 slice
      mbar_wait(&h_ready, 0);
      if (PS) mbar_wait(&b2_full, 0);       // big image of the B2 slice (requested when GEMM 1 released the ring)
      else mbar_wait(&b2_conv, 0);
      fence_after_sync();
      const uint32_t idesc2 = tma::make_idesc_tf32(kTileM3, kHid, false, BWD);
      const uint64_t step_b2 = tma::kstep_units(BWD);
      constexpr uint64_t b_small_off = C::kB2Big >> 4;
      if (!PS) {
        if (tma::elect_one()) {
#pragma unroll
          for (int c = 0; c < C::kNch2; ++c) {
            const uint64_t a_big = tma::desc_kmajor(smem_u32(hbuf + (size_t)c * C::kHTile));
            const uint64_t a_small = tma::desc_kmajor(smem_u32(hbuf + (size_t)(C::kNch2 + c) * C::kHTile));
            const uint32_t bb = smem_u32(ring + (size_t)c * C::kB2Chunk);
            const uint64_t b_big = BWD ? tma::desc_mnmajor(bb) : tma::desc_kmajor(bb);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tma::mma_3x(tmem2, a_big + ks * 2, a_small + ks * 2, b_big + ks * step_b2, b_big + b_small_off + ks * step_b2,
                          idesc2, (c | ks) ? 1u : 0u);
          }
          mma_commit(&acc2_full);
        }
      } else {
        // the small image of B2 lands in the region the h / gelu' staging tiles occupy until their TMA stores have read
        // them: the two products that need only the big image go first, big x small follows when it has arrived
        if (tma::elect_one()) {
#pragma unroll
          for (int c = 0; c < C::kNch2; ++c) {
            const uint64_t a_big = tma::desc_kmajor(smem_u32(hbuf + (size_t)c * C::kHTile));
            const uint64_t a_small = tma::desc_kmajor(smem_u32(hbuf + (size_t)(C::kNch2 + c) * C::kHTile));
            const uint32_t bb = smem_u32(ring + (size_t)c * C::kB2Chunk);
            const uint64_t b_big = BWD ? tma::desc_mnmajor(bb) : tma::desc_kmajor(bb);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              tma::mma_tf32(tmem2, a_small + ks * 2, b_big + ks * step_b2, idesc2, (c | ks) ? 1u : 0u);
              tma::mma_tf32(tmem2, a_big + ks * 2, b_big + ks * step_b2, idesc2, 1u);
            }
          }
        }
        __syncwarp();
        mbar_wait(&b2s_full, 0);
        fence_after_sync();
        if (tma::elect_one()) {
#pragma unroll
          for (int c = 0; c < C::kNch2; ++c) {
            const uint64_t a_big = tma::desc_kmajor(smem_u32(hbuf + (size_t)c * C::kHTile));
            const uint32_t bb = smem_u32(ring + (size_t)c * C::kB2Chunk);
            const uint64_t b_big = BWD ? tma::desc_mnmajor(bb) : tma::desc_kmajor(bb);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tma::mma_tf32(tmem2, a_big + ks * 2, b_big + b_small_off + ks * step_b2, idesc2, 1u);
          }
          mma_commit(&acc2_full);
        }
      }
      __syncwarp();
    }
  } else if (warp >= kFirstConvWarp) {
    // ------------------------------------------------------------------ converters / epilogues
    pdl_wait();
    const int cw = warp - kFirstConvWarp;
    const int ct = cw * 32 + lane;
#pragma unroll 1
    for (int i = 0; i < (PS ? 0 : kNkc1); ++i) {
      const int s = i % S;
      uint4 *tile = reinterpret_cast<uint4 *>(ring + (size_t)s * C::kStage1);
      mbar_wait(&full[s], (i / S) & 1);
      if (i == 0 && ct == 0) trace_mark(3);
      float4 v[C::kPer];
#pragma unroll
      for (int j = 0; j < C::kPer; ++j) v[j] = *reinterpret_cast<const float4 *>(tile + ct + j * (kConvWarps * 32));
#pragma unroll
      for (int j = 0; j < C::kPer; ++j) {
        uint4 b, sm;
        tma::split_tf32(v[j].x, b.x, sm.x);
        tma::split_tf32(v[j].y, b.y, sm.y);
        tma::split_tf32(v[j].z, b.z, sm.z);
        tma::split_tf32(v[j].w, b.w, sm.w);
        tile[ct + j * (kConvWarps * 32)] = b;
        tile[ct + j * (kConvWarps * 32) + C::kBig1 / 16] = sm;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&conv_done[s]);
    }
    if (ct == 0) trace_mark(4);
    // ---- epilogue 1: hidden tile -> operand images of GEMM 2 + exact fp32 staging tiles for the TMA stores
    const int q = cw & 3, half = cw >> 2;
    const int r = q * 32 + lane;            // row of the tile = TMEM lane
    uint8_t *stage_h = ring + C::kStageH;                                   // kNch2 boxes: h (forward) / du (backward)
    uint8_t *stage_gp = stage_h + (size_t)C::kNch2 * C::kBox;               // kNch2 boxes: gelu'(u) * mask (forward)
    Dropout drop = g.drop;
    if (!BWD) drop.resolve();
    {
      constexpr int kSteps1 = FS / 32;      // 16-column steps of this warp
      float bvs[kSteps1][16];
      if (!BWD) {
#pragma unroll
        for (int st = 0; st < kSteps1; ++st) load16(g.bias1 + f0 + half * (FS / 2) + st * 16, true, 16, bvs[st]);
      } else {
        mbar_wait(&gp_full, 0);
#pragma unroll
        for (int st = 0; st < kSteps1; ++st) {
          const int c0 = half * (FS / 2) + st * 16;
          const uint8_t *src = hbuf + (size_t)(C::kNch2 + (c0 >> 5)) * C::kHTile;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 t = *reinterpret_cast<const float4 *>(src + tma::sw128_offset(r, ((c0 & 31) >> 2) + j));
            bvs[st][4 * j] = t.x; bvs[st][4 * j + 1] = t.y; bvs[st][4 * j + 2] = t.z; bvs[st][4 * j + 3] = t.w;
          }
        }
      }
      mbar_wait(&acc1_full, 0);
      fence_after_sync();
      if (ct == 0) trace_mark(6);
#pragma unroll
      for (int st = 0; st < kSteps1; ++st) {
        const int c0 = half * (FS / 2) + st * 16;
        const float (&bv)[16] = bvs[st];
        uint32_t acc[16];
        tmem_ld16_3x(tmem1 + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
        tmem_ld_wait();
        float hv[16], gp[16];
        if (!BWD) {
          const long long off = (long long)(m0 + r) * g.F + f0 + c0;        // element offset = dropout counter
#pragma unroll
          for (int j8 = 0; j8 < 16; j8 += 8) {
            float m[8];
            drop.mask8((uint64_t)(off + j8), g.site, m);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = __uint_as_float(acc[j8 + j]) + bv[j8 + j];
              float cdf, pdf;
              gelu_cdf_pdf(x, cdf, pdf);
              hv[j8 + j] = x * cdf * m[j];
              gp[j8 + j] = fmaf(x, pdf, cdf) * m[j];
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) hv[j] = __uint_as_float(acc[j]) * bv[j];
        }
        // columns c0 .. c0+15 of the slice = K chunk c0 / 32, 16 B chunks (c0 % 32) / 4 .. +3 of row r
        const int kc = c0 >> 5, ch0 = (c0 & 31) >> 2;
        uint8_t *big = hbuf + (size_t)kc * C::kHTile, *small = hbuf + (size_t)(C::kNch2 + kc) * C::kHTile;
        uint8_t *sh = stage_h + (size_t)kc * C::kBox, *sg = stage_gp + (size_t)kc * C::kBox;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 b, sm;
          tma::split_tf32(hv[4 * j + 0], b.x, sm.x);
          tma::split_tf32(hv[4 * j + 1], b.y, sm.y);
          tma::split_tf32(hv[4 * j + 2], b.z, sm.z);
          tma::split_tf32(hv[4 * j + 3], b.w, sm.w);
          const uint32_t o = tma::sw128_offset(r, ch0 + j);
          *reinterpret_cast<uint4 *>(big + o) = b;
          *reinterpret_cast<uint4 *>(small + o) = sm;
          *reinterpret_cast<float4 *>(sh + o) = make_float4(hv[4 * j], hv[4 * j + 1], hv[4 * j + 2], hv[4 * j + 3]);
          if (!BWD) *reinterpret_cast<float4 *>(sg + o) = make_float4(gp[4 * j], gp[4 * j + 1], gp[4 * j + 2], gp[4 * j + 3]);
        }
      }
    }
    fence_async_smem();
    fence_before_sync();    // order this warp's tcgen05.ld of acc1 before the issuer's next MMAs
    __syncwarp();
    if (lane == 0) mbar_arrive(&h_ready);
    if (ct == 0) trace_mark(7);
    // ---- hidden tile (and gelu'*mask) -> global: TMA stores of the staging boxes.  The staging area is the
    // B2-small region, so the stores must have READ it before the B2 split below writes there.
    tma::named_barrier(1, kConvWarps * 32);
    if (cw == 0 && tma::elect_one()) {
#pragma unroll
      for (int c = 0; c < C::kNch2; ++c) {
        tma::store_2d(&g.map_h, stage_h + (size_t)c * C::kBox, f0 + 32 * c, m0);
        if (!BWD) tma::store_2d(&g.map_gp, stage_gp + (size_t)c * C::kBox, f0 + 32 * c, m0);
      }
      tma::store_commit();
      tma::store_wait_read();
      if (PS) {
        // the staging tiles have been read: the small image of the B2 slice may land there now
        mbar_expect_tx(&b2s_full, C::kB2Big);
#pragma unroll
        for (int c = 0; c < C::kNch2; ++c) {
          uint8_t *dst = ring + C::kB2Big + (size_t)c * C::kB2Chunk;
          const int k0 = f0 + c * 32;
          if (!BWD) tma::load_2d(dst, &g.map_b2_s, k0, 0, &b2s_full);
          else {
#pragma unroll
            for (int b = 0; b < kHid / 32; ++b) tma::load_2d(dst + b * tma::kMnBoxBytes, &g.map_b2_s, 32 * b, k0, &b2s_full);
          }
        }
      }
    }
    tma::named_barrier(1, kConvWarps * 32);
    // ---- B2 slice: big / small split in place
    if (!PS) {
      uint4 *tile = reinterpret_cast<uint4 *>(ring);
      mbar_wait(&b2_full, 0);
      if (ct == 0) trace_mark(8);
#pragma unroll 4
      for (int j = 0; j < C::kB2Per; ++j) {
        const float4 v = *reinterpret_cast<const float4 *>(tile + ct + j * (kConvWarps * 32));
        uint4 b, sm;
        tma::split_tf32(v.x, b.x, sm.x);
        tma::split_tf32(v.y, b.y, sm.y);
        tma::split_tf32(v.z, b.z, sm.z);
        tma::split_tf32(v.w, b.w, sm.w);
        tile[ct + j * (kConvWarps * 32)] = b;
        tile[ct + j * (kConvWarps * 32) + C::kB2Big / 16] = sm;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&b2_conv);
      if (ct == 0) trace_mark(9);
    }
    // ---- epilogue 2: partial [128 x 256] -> staging boxes (the ring is idle once GEMM 2 is complete) -> TMA reduce-add
    {
      const bool add_bias = !BWD && g.bias2 != nullptr && blockIdx.y == 0;
      mbar_wait(&acc2_full, 0);
      fence_after_sync();
      if (ct == 0) trace_mark(10);
#pragma unroll 2
      for (int c0 = half * (kHid / 2); c0 < (half + 1) * (kHid / 2); c0 += 16) {
        float bv[16];
        if (add_bias) load16(g.bias2 + c0, true, 16, bv);
        uint32_t acc[16];
        tmem_ld16_3x(tmem2 + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
        tmem_ld_wait();
        uint8_t *box = ring + (size_t)(c0 >> 5) * C::kBox;
        const int ch0 = (c0 & 31) >> 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 t = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]), __uint_as_float(acc[4 * j + 2]),
                                 __uint_as_float(acc[4 * j + 3]));
          if (add_bias) { t.x += bv[4 * j]; t.y += bv[4 * j + 1]; t.z += bv[4 * j + 2]; t.w += bv[4 * j + 3]; }
          *reinterpret_cast<float4 *>(box + tma::sw128_offset(r, ch0 + j)) = t;
        }
      }
      fence_async_smem();
      tma::named_barrier(1, kConvWarps * 32);
      if (cw == 0 && tma::elect_one()) {
#pragma unroll
        for (int b = 0; b < kHid / 32; ++b) tma::reduce_add_2d(&g.map_out2, ring + (size_t)b * C::kBox, 32 * b, m0);
        tma::store_commit();
        tma::store_wait_read();
      }
      __syncwarp();
      if (ct == 0) trace_mark(11);
    }
  } else {
    pdl_sync();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(s_tmem, C::kTmemCols);
}

template <int FS, bool BWD, bool PS = false>
static int launch_ffn_pair(const FfnPairParams &g, cudaStream_t stream) {
  using C = FfnPairCfg<FS>;
  static bool configured = false;
  if (!configured) {
    MSR3D_CUDA(cudaFuncSetAttribute(ffn_pair_kernel<FS, BWD, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
    configured = true;
  }
  launch_pdl(ffn_pair_kernel<FS, BWD, PS>, dim3(ceil_div(g.M, kTileM3), g.F / FS), dim3(kThreads3), C::kSmem, stream, g);
  return check_launch("ffn_pair");
}

}  // namespace msr3d

// small images of the pre-split variant (same shapes and boxes as the big ones)
static int ffn_pair_small_maps(FfnPairParams &g, const float *A_s, const float *W1_s, const float *W2_s, int M, int F, bool bwd, int fs) {
  MSR3D_REQUIRE(A_s && W1_s && W2_s, "ffn_pair: null small image");
  int rc = tma::make_map_f32(&g.map_a_s, A_s, M, kHid, kHid, kTileM3, false);
  if (rc != MSR3D_OK) return rc;
  if (!bwd) {
    rc = tma::make_map_f32(&g.map_b1_s, W1_s, F, kHid, kHid, fs, false);
    if (rc != MSR3D_OK) return rc;
    return tma::make_map_f32(&g.map_b2_s, W2_s, kHid, F, F, kHid, false);
  }
  rc = tma::make_map_f32(&g.map_b1_s, W2_s, kHid, F, F, tma::kBoxKRows, true);
  if (rc != MSR3D_OK) return rc;
  return tma::make_map_f32(&g.map_b2_s, W1_s, F, kHid, kHid, tma::kBoxKRows, true);
}

static int ffn_pair_common(FfnPairParams &g, const float *A, const float *W1, const float *W2, int M, int F, bool bwd,
                           int fs, const float *h_or_du, const float *gp, const float *out2) {
  MSR3D_REQUIRE(A && W1 && W2 && h_or_du && gp && out2, "ffn_pair: null pointer");
  MSR3D_REQUIRE(M > 0 && F > 0 && (fs == 32 || fs == 64) && F % fs == 0, "ffn_pair: bad sizes M=%d F=%d slice=%d", M, F, fs);
  g.M = M; g.F = F;
  int rc = tma::make_map_f32(&g.map_a, A, M, kHid, kHid, kTileM3, false);
  if (rc != MSR3D_OK) return rc;
  if (!bwd) {
    // B1 = W1 [F][256] K-major rows of the slice; B2 = W2 [256][F] K-major, all 256 rows, k = slice columns
    rc = tma::make_map_f32(&g.map_b1, W1, F, kHid, kHid, fs, false);
    if (rc != MSR3D_OK) return rc;
    rc = tma::make_map_f32(&g.map_b2, W2, kHid, F, F, kHid, false);
  } else {
    // B1 = W2 [256][F] read as [k = h][n = f] (MN-major); B2 = W1 [F][256] read as [k = f][n = h] (MN-major)
    rc = tma::make_map_f32(&g.map_b1, W2, kHid, F, F, tma::kBoxKRows, true);
    if (rc != MSR3D_OK) return rc;
    rc = tma::make_map_f32(&g.map_b2, W1, F, kHid, kHid, tma::kBoxKRows, true);
  }
  if (rc != MSR3D_OK) return rc;
  rc = tma::make_map_f32(&g.map_h, h_or_du, M, F, F, kTileM3, false);
  if (rc != MSR3D_OK) return rc;
  rc = tma::make_map_f32(&g.map_gp, gp, M, F, F, kTileM3, false);
  if (rc != MSR3D_OK) return rc;
  return tma::make_map_f32(&g.map_out2, out2, M, kHid, kHid, kTileM3, false);
}

/* see include/msr3d_b200.h */
MSR3D_API int msr3d_ffn_pair_fwd_f32(const float *y2, const float *W1, const float *b1, const float *W2, const float *b2,
                                     int M, int F, unsigned long long seed, const unsigned long long *seed_ptr, float p,
                                     unsigned site, float *h, float *gp, float *f_accum, int slice, msr3d_stream_t stream) {
  MSR3D_REQUIRE(b1 && (reinterpret_cast<uintptr_t>(b1) & 15) == 0 && (!b2 || (reinterpret_cast<uintptr_t>(b2) & 15) == 0),
                "ffn_pair_fwd: biases must be 16 B aligned");
  MSR3D_REQUIRE(p >= 0.f && p < 1.f, "ffn_pair_fwd: bad dropout p");
  FfnPairParams g{};
  const int fs = slice == 32 ? 32 : 64;
  const int rc = ffn_pair_common(g, y2, W1, W2, M, F, false, fs, h, gp, f_accum);
  if (rc != MSR3D_OK) return rc;
  g.bias1 = b1; g.bias2 = b2; g.site = site;
  g.drop = make_dropout(seed, p, seed_ptr);
  return fs == 32 ? launch_ffn_pair<32, false>(g, (cudaStream_t)stream) : launch_ffn_pair<64, false>(g, (cudaStream_t)stream);
}

MSR3D_API int msr3d_ffn_pair_bwd_f32(const float *df, const float *W1, const float *W2, const float *gp, int M, int F,
                                     float *du, float *dy2_accum, int slice, msr3d_stream_t stream) {
  FfnPairParams g{};
  const int fs = slice == 32 ? 32 : 64;
  const int rc = ffn_pair_common(g, df, W1, W2, M, F, true, fs, du, gp, dy2_accum);
  if (rc != MSR3D_OK) return rc;
  g.drop = make_dropout(0, 0.f);
  return fs == 32 ? launch_ffn_pair<32, true>(g, (cudaStream_t)stream) : launch_ffn_pair<64, true>(g, (cudaStream_t)stream);
}

/* Pre-split variants: every operand is given as its big / small tf32 images (msr3d_split_tf32_f32 for the weights, the
 * LayerNorm-tail kernels for the activations); no conversion pass inside the kernel. */
MSR3D_API int msr3d_ffn_pair_fwd_ps_f32(const float *y2_big, const float *y2_small, const float *W1_big, const float *W1_small,
                                        const float *b1, const float *W2_big, const float *W2_small, const float *b2, int M,
                                        int F, unsigned long long seed, const unsigned long long *seed_ptr, float p,
                                        unsigned site, float *h, float *gp, float *f_accum, int slice, msr3d_stream_t stream) {
  MSR3D_REQUIRE(b1 && (reinterpret_cast<uintptr_t>(b1) & 15) == 0 && (!b2 || (reinterpret_cast<uintptr_t>(b2) & 15) == 0),
                "ffn_pair_fwd: biases must be 16 B aligned");
  MSR3D_REQUIRE(p >= 0.f && p < 1.f, "ffn_pair_fwd: bad dropout p");
  FfnPairParams g{};
  const int fs = slice == 32 ? 32 : 64;
  int rc = ffn_pair_common(g, y2_big, W1_big, W2_big, M, F, false, fs, h, gp, f_accum);
  if (rc != MSR3D_OK) return rc;
  rc = ffn_pair_small_maps(g, y2_small, W1_small, W2_small, M, F, false, fs);
  if (rc != MSR3D_OK) return rc;
  g.bias1 = b1; g.bias2 = b2; g.site = site;
  g.drop = make_dropout(seed, p, seed_ptr);
  return fs == 32 ? launch_ffn_pair<32, false, true>(g, (cudaStream_t)stream) : launch_ffn_pair<64, false, true>(g, (cudaStream_t)stream);
}

MSR3D_API int msr3d_ffn_pair_bwd_ps_f32(const float *df_big, const float *df_small, const float *W1_big, const float *W1_small,
                                        const float *W2_big, const float *W2_small, const float *gp, int M, int F, float *du,
                                        float *dy2_accum, int slice, msr3d_stream_t stream) {
  FfnPairParams g{};
  const int fs = slice == 32 ? 32 : 64;
  int rc = ffn_pair_common(g, df_big, W1_big, W2_big, M, F, true, fs, du, gp, dy2_accum);
  if (rc != MSR3D_OK) return rc;
  rc = ffn_pair_small_maps(g, df_small, W1_small, W2_small, M, F, true, fs);
  if (rc != MSR3D_OK) return rc;
  g.drop = make_dropout(0, 0.f);
  return fs == 32 ? launch_ffn_pair<32, true, true>(g, (cudaStream_t)stream) : launch_ffn_pair<64, true, true>(g, (cudaStream_t)stream);
}
