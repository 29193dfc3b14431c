// The attention block of one spatial-encoder layer as ONE kernel (forward):
//
//   qkvc = x Wqkvc^T + b           (w_qs | w_ks | w_vs | lang_cond_fc, transformers.py:205-222)
//   P    = softmax(log(clamp(sigmoid(cond . pairwise), 1e-6)) + q k^T / sqrt(dh)),  ctx = P v     (:224-248)
//   o    = ctx Wo^T + bo                                                                           (:249)
//   y2   = LN_1(x + drop_1(LN_a(drop_fc(o) + x)))                                                  (:250-251, 324-325)
//
// Round 1/2 ran this as four dependent launches (QKV+cond GEMM, attention core, out-proj GEMM, LayerNorm tail):
// 27 us of a 47 us layer, almost all of it launch hand-over, operand round trips through L2 and per-kernel
// set-up.  Scenes are independent through the whole block, and inside a scene the heads are independent up to
// the output projection, so:
//
//   one thread-block CLUSTER of 8 CTAs per scene, CTA = head h.
//   P1  q_h | k_h | v_h | cond_h = x Wqkvc[rows of head h]^T: tcgen05 kind::tf32 "3xTF32" (big/small split, fp32
//       grade), M = 128 (64 rows of the scene tile are real), N = 112, K = 256 streamed through a 3-stage ring:
//       x chunks by tensor-map TMA and split in place by the worker warps; the weight rows arrive PRE-SPLIT
//       (msr3d_split_tf32_f32 runs once per step on the flat parameter buffer), requested before the
//       programmatic-dependent-launch wait, i.e. while the previous kernel is still draining.
//   P2  epilogue drops q, k^T, v, cond (+bias) into shared memory (and qkvc into HBM for the backward pass);
//       attention on the CUDA cores, one warp per query row with the keys over the lanes — the arithmetic of
//       attn_fwd_kernel (spatial.cu) — writes P, LOC, ctx for the backward pass and ctx_h as the big / small
//       operand images of P3.
//   P3  partial out-projection o_h = ctx_h Wo[:, 32h .. 32h+32]^T (K = 32, N = 256; the Wo slice was fetched at
//       kernel start) -> fp32 partial tile in shared memory.
//   P4  cluster barrier; CTA r sums the eight partials of rows 8r .. 8r+7 through DISTRIBUTED SHARED MEMORY in a
//       fixed order (deterministic, no atomics, no zero-filled accumulator in HBM) and runs the two-LayerNorm
//       tail on them: y2 and the saved statistics go to HBM.
// 32 CTAs at config 2 (B = 4): the rest of the GPU stays free for the frozen encoder of the next batch and the
// weight-gradient GEMMs that run beside the chain.
#include "common.cuh"
#include "rng.cuh"
#include "rowops.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace msr3d {

using namespace umma;

// phase timeline for tools/trace_attn_block.py: compiled in only with -DMSR3D_TRACE (`make trace`); a mark is ~14
// instructions of one-shot code, and one-shot code costs ~5 cycles per instruction in instruction fetch
#if defined(MSR3D_TRACE)
#define AB_TRACE(cond, slot) do { if (cond) trace_mark(slot); } while (0)
#else
#define AB_TRACE(cond, slot) do { } while (0)
#endif

constexpr int kAbHeads = 8, kAbDh = 32, kAbCond = 6;
constexpr int kAbQkvc = 3 * kD + kAbHeads * kAbCond;   // 816
constexpr int kAbRT = 64;                              // rows of the scene tile kept in shared memory (L <= 64)
constexpr int kAbN1 = 112;                             // q 32 | k 32 | v 32 | cond 6 (+2 rows of the next head) | 8 unused
constexpr int kAbThreads = 384;
constexpr int kAbWorkers = 8;                          // warps 4..11: converters; then "row" warps or "pair" warps
constexpr int kAbRowThreads = 128;                     // warps 4, 5, 8, 9: own TMEM lanes 0..63 = the rows of the scene
constexpr int kAbPairThreads = 192;                    // warps 2, 3, 6, 7, 10, 11: (row, key) pair work, coalesced
constexpr int kAbStages = 3;
constexpr int kAbChunks = kD / 32;                     // 8 K chunks of 32
constexpr uint32_t kAbABytes = kAbRT * tma::kRowBytes;        // 8 KB: one image of a 64-row x 32-k operand tile
constexpr uint32_t kAbBBytes = kAbN1 * tma::kRowBytes;        // 14 KB: one image of a weight chunk
constexpr uint32_t kAbBLoaded = (96 + 8) * tma::kRowBytes;    // 13 KB of it are written by TMA
constexpr uint32_t kAbImg = kAbABytes + kAbBBytes;            // big image of a stage; = offset of the small image
constexpr uint32_t kAbStage = 2 * kAbImg;                     // [A big | B big | A small | B small] 44 KB
constexpr uint32_t kAbRing = kAbStages * kAbStage;            // 132 KB
constexpr uint32_t kAbWoImg = kD * tma::kRowBytes;            // 32 KB: Wo[:, 32h..32h+32] as 256 rows x 32 k
// after P1 the ring is idle and holds the operand tiles of the attention MMAs (every tile 1024 B aligned; an
// M = 128 MMA reads 16 KB from an A tile of which the first 8 KB = 64 rows are real: what follows only has to exist)
constexpr uint32_t kAbCtxBig = 0, kAbCtxSmall = 8 * 1024;     // A of P3
constexpr uint32_t kAbQBig = 16 * 1024, kAbQSmall = 24 * 1024;   // A of S = q k^T;    later P big   (chunks 0, 1)
constexpr uint32_t kAbKBig = 32 * 1024, kAbKSmall = 40 * 1024;   // B of S (N = 64);   later P small (chunks 0, 1)
constexpr uint32_t kAbPBig = kAbQBig, kAbPSmall = kAbKBig;
constexpr uint32_t kAbVtBig = 48 * 1024, kAbVtSmall = 56 * 1024; // B of P v: v^T [32 dims][64 keys], 2 chunks of 4 KB
constexpr uint32_t kAbCs = 64 * 1024;                         // cond + bias [64][8]
constexpr uint32_t kAbRed = 66 * 1024;                        // softmax partials: max [2][64], sum [2][64]
constexpr uint32_t kAbZs = 67 * 1024;                         // log(clamp(sigmoid(cond . pairwise))) [64][kAbZp]
constexpr int kAbZp = 65;
// Wo slice (fetched when the ring is idle)
constexpr uint32_t kAbWoOff = 84 * 1024;
static_assert(kAbZs + kAbRT * kAbZp * 4 <= kAbWoOff && kAbWoOff + 2 * kAbWoImg <= 148 * 1024, "Wo slice placement");
// pairwise features of the scene [L][L][5] (copied by warps 2, 3 while P1 runs; L <= 60, else read from global);
// once EVERY CTA of the cluster has consumed its copy (cluster barrier #1) the region is the exchange tile
// [8 heads][8 rows][256] the peers push their partial out-projections into
constexpr uint32_t kAbPw = 148 * 1024, kAbXchg = kAbPw;
static_assert(kAbPw >= kAbRing, "the pairwise features are copied while the ring is busy");
constexpr uint32_t kAbPwBytes = 60 * 60 * 5 * 4;              // 72 000
constexpr uint32_t kAbXchgBytes = kAbHeads * 8 * kD * 4;      // 64 KB
static_assert(kAbXchgBytes <= kAbPwBytes, "exchange tile = the pairwise-feature region");
constexpr size_t kAbSmem = (size_t)kAbPw + ((kAbPwBytes + 1023) / 1024) * 1024 + 1024;
constexpr uint32_t kAbTmemCols = 512;
constexpr uint32_t kAbAccSCol = 128, kAbAccPvCol = 192, kAbAcc3Col = 256;

struct alignas(64) AttnBlockFwdParams {
  CUtensorMap map_x;                     // [B*L, 256]  box {32, 64}
  CUtensorMap map_wq_b32, map_wq_s32;    // [816, 256]  box {32, 32}   big / small images of Wqkvc
  CUtensorMap map_wq_b8, map_wq_s8;      // [816, 256]  box {32, 8}
  CUtensorMap map_wo_b, map_wo_s;        // [256, 256]  box {32, 256}
  const float *x;                        // [B*L, 256] layer input (residual)
  const float *bqkvc, *bo;               // [816], [256]
  const float *pw;                       // [B, L, L, 5]
  const unsigned char *key_pad;          // [B, L] (nullable)
  LnParams lna, ln1;
  Dropout drop;
  unsigned site_fc, site_1;
  float *qkvc, *P, *LOC, *ctx;           // saved for the backward pass
  float *xhat_a, *rstd_a, *xhat_1, *rstd_1, *y2;
  float *y2_big, *y2_small;              // nullable: operand images of y2 for the pre-split FFN pair
  int B, L;
};

__device__ __forceinline__ void ab_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// cluster barrier, split: arrive (release) ... wait (acquire); every thread of every CTA executes both, in order
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// shared-memory address of the same variable in CTA `rank` of the cluster, and a 16 B store to it
__device__ __forceinline__ uint32_t cluster_map(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Shared memory through 32-bit shared-space addresses: the dynamic buffer is carved by address arithmetic, which
// makes the compiler fall back to GENERIC loads / stores (LD.E / ST.E with a descriptor register pair each) — twice
// the instructions in code that mostly runs once, where instruction fetch is what costs.
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts32u(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128u(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// one 16 B chunk (4 consecutive k) of row r of a swizzled K-major operand tile, both images
__device__ __forceinline__ void put4(uint32_t big, uint32_t small, int r, int chunk, float a, float b, float c, float d) {
  uint4 bg, sm;
  tma::split_tf32(a, bg.x, sm.x);
  tma::split_tf32(b, bg.y, sm.y);
  tma::split_tf32(c, bg.z, sm.z);
  tma::split_tf32(d, bg.w, sm.w);
  const uint32_t o = tma::sw128_offset(r, chunk);
  sts128u(big + o, bg);
  sts128u(small + o, sm);
}

__global__ void __launch_bounds__(kAbThreads, 1)
attn_block_fwd_kernel(const __grid_constant__ AttnBlockFwdParams g) {
  constexpr int S = kAbStages;
  extern __shared__ uint8_t smem_raw[];
  // the same offset in every CTA of the cluster (distributed-shared-memory addresses are mapped by offset)
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *ring = smem;
  uint8_t *wo = smem + kAbWoOff;
  const uint32_t sb = smem_u32(smem);          // shared-space address of the carved buffer
  __shared__ __align__(8) uint64_t full[S], conv_done[S], mma_done[S];
  __shared__ __align__(8) uint64_t acc1_full, wo_full, qk_ready, accs_full, p_ready, accpv_full, ctx_ready, acc3_full;
  __shared__ uint32_t s_tmem;
  // parameter vectors (static: only the optimiser of the previous step writes them), staged once so that no epilogue
  // waits for an L2 round trip: bias of this head's q | k | v | cond rows, out-projection bias, LayerNorm affine pairs
  __shared__ __align__(16) float s_bq[kAbN1], s_bo[kD], s_ga[kD], s_ba[kD], s_g1[kD], s_b1[kD];
  __shared__ float s_pad[kAbRT];      // 1 = padded key

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t h_;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(h_));
  const int h = (int)h_;                       // head = rank in the cluster
  const int b = blockIdx.y;
  const int L = g.L;
  const int row0 = b * L;                      // first token row of this scene

  if (tid == 0) {
    AB_TRACE(true, 0);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv_done[s], kAbWorkers);
      mbar_init(&mma_done[s], 1);
    }
    mbar_init(&acc1_full, 1);
    mbar_init(&wo_full, 1);
    mbar_init(&qk_ready, 4);
    mbar_init(&accs_full, 1);
    mbar_init(&p_ready, 4);
    mbar_init(&accpv_full, 1);
    mbar_init(&ctx_ready, 4);
    mbar_init(&acc3_full, 1);
    fence_mbar_init();
    tma::prefetch_map(&g.map_x);
    tma::prefetch_map(&g.map_wq_b32);
    tma::prefetch_map(&g.map_wq_s32);
    tma::prefetch_map(&g.map_wq_b8);
    tma::prefetch_map(&g.map_wq_s8);
    tma::prefetch_map(&g.map_wo_b);
    tma::prefetch_map(&g.map_wo_s);
  }
  if (warp == 2) tmem_alloc(&s_tmem, kAbTmemCols);
  if (tid >= 128) {
    const int i = tid - 128;     // 0..255
    s_bo[i] = g.bo[i];
    s_ga[i] = g.lna.gamma[i]; s_ba[i] = g.lna.beta[i];
    s_g1[i] = g.ln1.gamma[i]; s_b1[i] = g.ln1.beta[i];
    if (i < kAbRT) s_pad[i] = (g.key_pad != nullptr && i < g.L && g.key_pad[(size_t)blockIdx.y * g.L + i] != 0) ? 1.f : 0.f;
    if (i < 96) s_bq[i] = g.bqkvc[(i >> 5) * kD + h * kAbDh + (i & 31)];
    else if (i < 96 + kAbCond) s_bq[i] = g.bqkvc[3 * kD + h * kAbCond + (i - 96)];
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem1 = s_tmem, tmem_s = s_tmem + kAbAccSCol, tmem_pv = s_tmem + kAbAccPvCol, tmem3 = s_tmem + kAbAcc3Col;
  AB_TRACE(tid == 0, 1);

  const bool row_warp = warp >= 4 && (warp & 3) < 2;    // warps 4, 5, 8, 9
  const int qd = warp & 3, half = (warp - 4) >> 2;      // row warps: TMEM lane quarter (0, 1) and column half
  const int r = qd * 32 + lane;                         // row warps: this thread's row of the scene tile

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane)
    if (elect_one()) {
      // weights are static operands (written by the optimiser of the previous step only): before the PDL wait
      auto load_b = [&](int i) {
        const int s = i % S;
        const int k0 = i * 32;
#pragma unroll
        for (int img = 0; img < 2; ++img) {
          uint8_t *dst = ring + (size_t)s * kAbStage + (size_t)img * kAbImg + kAbABytes;
          const CUtensorMap *m32 = img ? &g.map_wq_s32 : &g.map_wq_b32;
          const CUtensorMap *m8 = img ? &g.map_wq_s8 : &g.map_wq_b8;
          tma::load_2d(dst, m32, k0, h * kAbDh, &full[s]);                        // w_qs rows of head h
          tma::load_2d(dst + 4096, m32, k0, kD + h * kAbDh, &full[s]);            // w_ks
          tma::load_2d(dst + 8192, m32, k0, 2 * kD + h * kAbDh, &full[s]);        // w_vs
          tma::load_2d(dst + 12288, m8, k0, 3 * kD + h * kAbCond, &full[s]);      // lang_cond_fc (6 rows + 2)
        }
      };
      auto load_a = [&](int i) {
        const int s = i % S;
        tma::load_2d(ring + (size_t)s * kAbStage, &g.map_x, i * 32, row0, &full[s]);
      };
      for (int i = 0; i < S; ++i) {
        mbar_expect_tx(&full[i], kAbABytes + 2 * kAbBLoaded);
        load_b(i);
      }
      pdl_wait();
      AB_TRACE(true, 2);
      for (int i = 0; i < S; ++i) load_a(i);
      pdl_launch_dependents();
      for (int i = S; i < kAbChunks; ++i) {
        const int s = i % S;
        mbar_wait(&mma_done[s], ((i / S) - 1) & 1);
        mbar_expect_tx(&full[s], kAbABytes + 2 * kAbBLoaded);
        load_b(i);
        load_a(i);
      }
      // the ring is idle once the accumulator is complete: the Wo slice lands behind the attention tiles
      mbar_wait(&acc1_full, 0);
      mbar_expect_tx(&wo_full, 2 * kAbWoImg);
      tma::load_2d(wo, &g.map_wo_b, h * kAbDh, 0, &wo_full);
      tma::load_2d(wo + kAbWoImg, &g.map_wo_s, h * kAbDh, 0, &wo_full);
    }
    __syncwarp();
    cluster_arrive_relaxed();      // #1
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc1 = tma::make_idesc_tf32(128, kAbN1, false, false);
#pragma unroll 1
    for (int i = 0; i < kAbChunks; ++i) {
      const int s = i % S;
      mbar_wait(&conv_done[s], (i / S) & 1);
      fence_after_sync();
      if (elect_one()) {
        const uint32_t base = smem_u32(ring + (size_t)s * kAbStage);
        const uint64_t a_big = tma::desc_kmajor(base), b_big = tma::desc_kmajor(base + kAbABytes);
        constexpr uint64_t small = kAbImg >> 4;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tma::mma_3x(tmem1, a_big + ks * 2, a_big + small + ks * 2, b_big + ks * 2, b_big + small + ks * 2, idesc1,
                      (i | ks) ? 1u : 0u);
        mma_commit(&mma_done[s]);
        if (i == kAbChunks - 1) mma_commit(&acc1_full);
      }
      __syncwarp();
    }
    cluster_arrive_relaxed();      // #1
    // S = q k^T  (scaled in the epilogue)
    mbar_wait(&qk_ready, 0);
    fence_after_sync();
    if (elect_one()) {
      const uint32_t idesc = tma::make_idesc_tf32(128, 64, false, false);
      const uint64_t a_big = tma::desc_kmajor(smem_u32(ring + kAbQBig)), a_small = tma::desc_kmajor(smem_u32(ring + kAbQSmall));
      const uint64_t b_big = tma::desc_kmajor(smem_u32(ring + kAbKBig)), b_small = tma::desc_kmajor(smem_u32(ring + kAbKSmall));
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        tma::mma_3x(tmem_s, a_big + ks * 2, a_small + ks * 2, b_big + ks * 2, b_small + ks * 2, idesc, ks ? 1u : 0u);
      mma_commit(&accs_full);
    }
    __syncwarp();
    // ctx_h = P v
    mbar_wait(&p_ready, 0);
    fence_after_sync();
    if (elect_one()) {
      const uint32_t idesc = tma::make_idesc_tf32(128, kAbDh, false, false);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint64_t a_big = tma::desc_kmajor(smem_u32(ring + kAbPBig + c * kAbABytes));
        const uint64_t a_small = tma::desc_kmajor(smem_u32(ring + kAbPSmall + c * kAbABytes));
        const uint64_t b_big = tma::desc_kmajor(smem_u32(ring + kAbVtBig + c * 4096));
        const uint64_t b_small = tma::desc_kmajor(smem_u32(ring + kAbVtSmall + c * 4096));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tma::mma_3x(tmem_pv, a_big + ks * 2, a_small + ks * 2, b_big + ks * 2, b_small + ks * 2, idesc, (c | ks) ? 1u : 0u);
      }
      mma_commit(&accpv_full);
    }
    __syncwarp();
    // o_h = ctx_h Wo_h^T
    mbar_wait(&ctx_ready, 0);
    mbar_wait(&wo_full, 0);
    fence_after_sync();
    if (elect_one()) {
      const uint32_t idesc3 = tma::make_idesc_tf32(128, kD, false, false);
      const uint64_t a_big = tma::desc_kmajor(smem_u32(ring + kAbCtxBig)), a_small = tma::desc_kmajor(smem_u32(ring + kAbCtxSmall));
      const uint64_t b_big = tma::desc_kmajor(smem_u32(wo)), b_small = tma::desc_kmajor(smem_u32(wo + kAbWoImg));
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        tma::mma_3x(tmem3, a_big + ks * 2, a_small + ks * 2, b_big + ks * 2, b_small + ks * 2, idesc3, ks ? 1u : 0u);
      mma_commit(&acc3_full);
    }
    __syncwarp();
  } else {
    pdl_wait();
    const int npw = L * L * 5;
    const bool pw_staged = (uint32_t)npw * 4u <= kAbPwBytes;
    if (warp < 4 && pw_staged) {
      // ---- warps 2, 3: pairwise features of the scene -> shared memory (asynchronous copies, no registers)
      const float *src = g.pw + (size_t)b * npw;
      const uint32_t dst = smem_u32(smem + kAbPw);
      const int i0 = (warp - 2) * 32 + lane;
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (npw & 3) == 0) {
#pragma unroll 4
        for (int i = i0; i < (npw >> 2); i += 64) cp_async16(dst + (uint32_t)i * 16u, src + (size_t)i * 4);
      } else {
#pragma unroll 4
        for (int i = i0; i < npw; i += 64) cp_async4(dst + (uint32_t)i * 4u, src + i);
      }
      cp_async_wait_all();
    }
    if (warp >= 4) {
      // ---- P1: split the landed x chunk into big / small images in place (the weights arrive split)
      const int ct = (warp - 4) * 32 + lane;
#pragma unroll 1
      for (int i = 0; i < kAbChunks; ++i) {
        const int s = i % S;
        const uint32_t tile = sb + (uint32_t)s * kAbStage + (uint32_t)ct * 16u;
        mbar_wait(&full[s], (i / S) & 1);
        AB_TRACE(i == 0 && ct == 0, 3);
        float4 v[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) v[j] = lds128(tile + j * 4096);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 bg, sm;
          tma::split_tf32(v[j].x, bg.x, sm.x);
          tma::split_tf32(v[j].y, bg.y, sm.y);
          tma::split_tf32(v[j].z, bg.z, sm.z);
          tma::split_tf32(v[j].w, bg.w, sm.w);
          sts128u(tile + j * 4096, bg);
          sts128u(tile + j * 4096 + kAbImg, sm);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_done[s]);
      }
      AB_TRACE(ct == 0, 4);
    }
    if (row_warp) {
      // ================================================================ row warps: thread = row r of the scene
      const uint32_t cs = sb + kAbCs, red = sb + kAbRed;
      const bool row_ok = r < L;
      {
        // ---- P1 epilogue: q, k -> operand tiles of S;  v -> v^T tiles;  cond -> smem;  qkvc -> HBM (backward)
        float *grow = g.qkvc + (size_t)(row0 + r) * kAbQkvc;
        mbar_wait(&acc1_full, 0);
        fence_after_sync();
        AB_TRACE(warp == 4 && lane == 0, 5);
        if (half == 1) {
          // cond first: the pair warps wait for it
          uint32_t acc[16];
          ab_tmem_ld16(tmem1 + ((uint32_t)(qd * 32) << 16) + 96u, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < kAbCond; ++j) {
            const float v = __uint_as_float(acc[j]) + s_bq[96 + j];
            sts32(cs + (uint32_t)(r * 8 + j) * 4u, v);
            if (row_ok) grow[3 * kD + h * kAbCond + j] = v;
          }
          named_arrive(3, 64 + kAbPairThreads);
        }
#pragma unroll 1
        for (int sec = half * 2; sec < 3 && sec < half * 2 + 2; ++sec) {      // 0 q, 1 k (half 0); 2 v (half 1)
          uint32_t acc[32];
          tmem_ld32(tmem1 + ((uint32_t)(qd * 32) << 16) + (uint32_t)(sec * 32), acc);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) + s_bq[sec * 32 + j];
          if (row_ok) {
            float *o = grow + sec * kD + h * kAbDh;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (sec < 2) {
            const uint32_t big = sb + (sec == 0 ? kAbQBig : kAbKBig), small = sb + (sec == 0 ? kAbQSmall : kAbKSmall);
#pragma unroll
            for (int j = 0; j < 8; ++j) put4(big, small, r, j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            // v^T: element (dim d, key r) of chunk r / 32
            const uint32_t big = sb + kAbVtBig + (r >> 5) * 4096, small = sb + kAbVtSmall + (r >> 5) * 4096;
            const int kk = r & 31;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              uint32_t bg, sm;
              tma::split_tf32(v[j], bg, sm);
              const uint32_t o = tma::sw128_offset(j, kk >> 2) + ((uint32_t)(kk & 3) << 2);
              sts32u(big + o, bg);
              sts32u(small + o, sm);
            }
          }
        }
        fence_before_sync();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&qk_ready);
        AB_TRACE(warp == 4 && lane == 0, 6);
        cluster_arrive_relaxed();   // #1 (completes when every pair warp of the cluster is done with the pairwise features)
      }
      {
        // ---- softmax: logits = log-loc (pair warps) + S * scale;  P -> HBM and operand tiles of P v
        const float scale = rsqrtf((float)kAbDh);
        const int nk = max(0, min(32, L - half * 32));      // keys of this half
        const uint32_t zrow = sb + kAbZs + (uint32_t)(r * kAbZp + half * 32) * 4u;
        tma::named_barrier(4, kAbRowThreads + kAbPairThreads);   // log-loc table complete
        AB_TRACE(warp == 4 && lane == 0, 14);
        mbar_wait(&accs_full, 0);
        fence_after_sync();
        AB_TRACE(warp == 4 && lane == 0, 15);
        uint32_t acc[32];
        tmem_ld32(tmem_s + ((uint32_t)(qd * 32) << 16) + (uint32_t)(half * 32), acc);
        tmem_ld_wait();
        float sj[32];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sj[j] = j < nk ? lds32(zrow + j * 4) + __uint_as_float(acc[j]) * scale : -INFINITY;
          mx = fmaxf(mx, sj[j]);
        }
        sts32(red + (uint32_t)(half * 64 + r) * 4u, mx);
        tma::named_barrier(2, kAbRowThreads);
        mx = fmaxf(lds32(red + (uint32_t)r * 4u), lds32(red + (uint32_t)(64 + r) * 4u));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sj[j] = __expf(sj[j] - mx);     // exp(-inf) = 0 for masked / out-of-range keys
          sum += sj[j];
        }
        sts32(red + (uint32_t)(128 + half * 64 + r) * 4u, sum);
        tma::named_barrier(2, kAbRowThreads);
        const float inv = 1.f / (lds32(red + (uint32_t)(128 + r) * 4u) + lds32(red + (uint32_t)(192 + r) * 4u));
#pragma unroll
        for (int j = 0; j < 32; ++j) sj[j] *= inv;
        if (row_ok && nk > 0) {
          float *prow = g.P + (((size_t)b * kAbHeads + h) * L + r) * L + half * 32;
          if ((L & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < nk) *reinterpret_cast<float4 *>(prow + j) = make_float4(sj[j], sj[j + 1], sj[j + 2], sj[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nk) prow[j] = sj[j];
          }
        }
        const uint32_t big = sb + kAbPBig + half * kAbABytes, small = sb + kAbPSmall + half * kAbABytes;
#pragma unroll
        for (int j = 0; j < 8; ++j) put4(big, small, r, j, sj[4 * j], sj[4 * j + 1], sj[4 * j + 2], sj[4 * j + 3]);
        fence_before_sync();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready);
        AB_TRACE(warp == 4 && lane == 0, 7);
      }
      {
        // ---- ctx_h: 16 dims per thread -> HBM (backward) and the operand tile of the out-projection
        mbar_wait(&accpv_full, 0);
        fence_after_sync();
        uint32_t acc[16];
        ab_tmem_ld16(tmem_pv + ((uint32_t)(qd * 32) << 16) + (uint32_t)(half * 16), acc);
        tmem_ld_wait();
        if (row_ok) {
          float *o = g.ctx + ((size_t)row0 + r) * kD + h * kAbDh + half * 16;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4 *>(o + j) = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                                             __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          put4(sb + kAbCtxBig, sb + kAbCtxSmall, r, half * 4 + j, __uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
               __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
        fence_before_sync();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctx_ready);
      }
      {
        // ---- partial out-projection of this head: rows are PUSHED into the exchange tile of the CTA that owns them
        const int rpc = (L + kAbHeads - 1) / kAbHeads;
        const int owner = min(r / rpc, kAbHeads - 1), lr = r - owner * rpc;
        const uint32_t dst = cluster_map(smem_u32(ring + kAbXchg) + (uint32_t)(((h * 8 + lr) * kD + half * 128) * 4), (uint32_t)owner);
        mbar_wait(&acc3_full, 0);
        fence_after_sync();
        AB_TRACE(warp == 4 && lane == 0, 8);
        cluster_wait();             // #1: the pairwise-feature region of every CTA is the exchange tile now
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 64) {
          uint32_t acc[32], acc2[32];
          tmem_ld32(tmem3 + ((uint32_t)(qd * 32) << 16) + (uint32_t)(half * 128 + c0), acc);
          tmem_ld32(tmem3 + ((uint32_t)(qd * 32) << 16) + (uint32_t)(half * 128 + c0 + 32), acc2);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              st_cluster_f4(dst + (uint32_t)((c0 + j) * 4), __uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                            __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              st_cluster_f4(dst + (uint32_t)((c0 + 32 + j) * 4), __uint_as_float(acc2[j]), __uint_as_float(acc2[j + 1]),
                            __uint_as_float(acc2[j + 2]), __uint_as_float(acc2[j + 3]));
          }
        }
        fence_before_sync();
        AB_TRACE(warp == 4 && lane == 0, 9);
      }
    } else {
      // ================================================================ pair warps (2, 3, 6, 7, 10, 11): the location
      // term of every (query row, key) pair, coalesced over the keys:  loc = sigmoid(cond . pairwise + c0),
      // LOC -> HBM (backward), log(clamp(loc, 1e-6)) (-inf for padded keys) -> shared memory for the softmax
      const int pi = (warp < 4 ? warp - 2 : (warp < 8 ? warp - 4 : warp - 6)) * 32 + lane;    // 0..191
      const uint32_t cs = sb + kAbCs, zs = sb + kAbZs;
      float *locs = g.LOC + ((size_t)b * kAbHeads + h) * L * L;
      const int npair = L * L;
      const int dl = kAbPairThreads / L, dt = kAbPairThreads - dl * L;
      int l = pi / L, t = pi - l * L;
      // cond (+bias) of every row and the pairwise features are in shared memory
      tma::named_barrier(3, 64 + kAbPairThreads);
      auto finish = [&](int p, float g0, float g1, float g2, float g3, float g4) {
        const float4 sw = lds128(cs + (uint32_t)l * 32u);
        const float sw4 = lds32(cs + (uint32_t)l * 32u + 16u), sw5 = lds32(cs + (uint32_t)l * 32u + 20u);
        float z = fmaf(sw.y, g0, sw.x);
        z = fmaf(sw.z, g1, z);
        z = fmaf(sw.w, g2, z);
        z = fmaf(sw4, g3, z);
        z = fmaf(sw5, g4, z);
        const bool pad = s_pad[t] != 0.f;
        const float lc = pad ? 0.f : __fdividef(1.f, 1.f + __expf(-z));
        const float lg = pad ? -INFINITY : __logf(fmaxf(lc, 1e-6f));
        locs[p] = lc;
        sts32(zs + (uint32_t)(l * kAbZp + t) * 4u, lg);
        t += dt; l += dl;
        if (t >= L) { t -= L; ++l; }
      };
      if (pw_staged) {
        const uint32_t pws = sb + kAbPw;
#pragma unroll 2
        for (int p = pi; p < npair; p += kAbPairThreads) {
          const uint32_t a = pws + (uint32_t)p * 20u;
          finish(p, lds32(a), lds32(a + 4), lds32(a + 8), lds32(a + 12), lds32(a + 16));
        }
      } else {
        const float *pws = g.pw + (size_t)b * npw;
#pragma unroll 1
        for (int p = pi; p < npair; p += kAbPairThreads) {
          const float *gp = pws + (size_t)p * 5;
          finish(p, gp[0], gp[1], gp[2], gp[3], gp[4]);
        }
      }
      AB_TRACE(warp == 2 && lane == 0, 13);
      named_arrive(4, kAbRowThreads + kAbPairThreads);
      cluster_arrive();              // #1: this warp has consumed its share of the pairwise features
    }
    if (!row_warp) cluster_wait();   // #1 (row warps waited before pushing)
  }
  if (warp < 2) cluster_wait();      // #1

  // ------------------------------------------------------------------ P4: sum the heads' partials, LayerNorm tail
  cluster_arrive();      // #2: this CTA's pushes are done ...
  const int cw = warp - 4;
  const int rpc = (L + kAbHeads - 1) / kAbHeads;      // rows per CTA (<= 8)
  const int l = h * rpc + cw;
  const bool tail_row = warp >= 4 && cw < rpc && l < L;
  Row8 xr, mfc, m1;
  if (tail_row) {        // residual row and dropout masks: independent of the pushes
    const size_t base = (size_t)(row0 + l) * kD;
    xr = row_load(g.x + base, lane);
    Dropout drop = g.drop;
    drop.resolve();
    mfc = drop_mask8(drop, base, lane, g.site_fc);
    m1 = drop_mask8(drop, base, lane, g.site_1);
  }
  cluster_wait();        //     ... and everybody else's have landed
  AB_TRACE(tid == 128, 10);
  if (warp >= 4) {
    if (tail_row) {
      const uint32_t xl = sb + kAbXchg + (uint32_t)(cw * kD + lane * 8) * 4u;
      Row8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = 0.f;
#pragma unroll
      for (int p = 0; p < kAbHeads; ++p) {      // fixed order: deterministic
        const float4 u = lds128(xl + (uint32_t)p * (8 * kD * 4)), w = lds128(xl + (uint32_t)p * (8 * kD * 4) + 16u);
        o.v[0] += u.x; o.v[1] += u.y; o.v[2] += u.z; o.v[3] += u.w;
        o.v[4] += w.x; o.v[5] += w.y; o.v[6] += w.z; o.v[7] += w.w;
      }
      const int row = row0 + l;
      const size_t base = (size_t)row * kD;
      const Row8 bo = row_load(s_bo, lane);
      Row8 v;
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] = (o.v[i] + bo.v[i]) * mfc.v[i] + xr.v[i];
      float ra, r1;
      Row8 y1 = ln_forward(v, s_ga, s_ba, lane, ra);
      row_store(g.xhat_a + base, lane, v);
      Row8 w;
#pragma unroll
      for (int i = 0; i < 8; ++i) w.v[i] = xr.v[i] + y1.v[i] * m1.v[i];
      Row8 out = ln_forward(w, s_g1, s_b1, lane, r1);
      row_store(g.xhat_1 + base, lane, w);
      row_store(g.y2 + base, lane, out);
      if (g.y2_big != nullptr) {
        Row8 bg, sm;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t ub, us;
          tma::split_tf32(out.v[i], ub, us);
          bg.v[i] = __uint_as_float(ub);
          sm.v[i] = __uint_as_float(us);
        }
        row_store(g.y2_big + base, lane, bg);
        row_store(g.y2_small + base, lane, sm);
      }
      if (lane == 0) { g.rstd_a[row] = ra; g.rstd_1[row] = r1; }
    }
  }
  AB_TRACE(tid == 128, 11);
  // nobody pushes into a CTA after barrier #2, so CTAs may exit independently
  fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    fence_after_sync();
    tmem_dealloc(s_tmem, kAbTmemCols);
  }
  AB_TRACE(tid == 128, 12);
}

// ---------------------------------------------------------------------------------------------------------------
// Backward of the attention core on the tensor cores (replaces attn_bwd_kernel for L <= 64; same inputs / outputs:
// transformers.py:205-248 differentiated).  One CTA per (head, scene):
//     dP = dctx v^T            dl = P o (dP - rowsum(dP o P))        (softmax backward, thread = query row)
//     dq = dl k / sqrt(dh)     dk = dl^T q / sqrt(dh)     dv = P^T dctx
//     d cond[l] = sum_t gz[l,t] [1, pairwise(l,t)],   gz = [loc >= 1e-6] dl sigmoid(-z)        (pair warps, keys over lanes)
// The CUDA-core kernel needed a cluster of four 16-warp CTAs per (head, scene) — 128 CTAs competing with the
// weight-gradient GEMMs for the chain's SMs, 20 us alone and 35 us beside them; the four contractions are 84 small MMAs.
// Operand tiles are written by the row warps straight from global memory (natural and transposed, big / small images).
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kBbGn = 0, kBbVn = 16 * 1024;                    // dctx [l][d] (A of dP), v [t][d] (B of dP); later dl [l][t]
constexpr uint32_t kBbDs = 0;                                       // dl [l][t]: big chunks 0,1 | small chunks 0,1 (A of dq)
constexpr uint32_t kBbDst = 32 * 1024;                              // dl^T [t][l] (A of dk)
constexpr uint32_t kBbPt = 64 * 1024;                               // P^T  [t][l] (A of dv)
constexpr uint32_t kBbQt = 96 * 1024, kBbKt = 112 * 1024, kBbGt = 128 * 1024;   // q^T, k^T, dctx^T [d][64]: B tiles, N = 32
constexpr uint32_t kBbDl = 144 * 1024;                              // dl as plain fp32 [64][kAbZp] for the pair warps
constexpr uint32_t kBbCs = kBbDl + 17 * 1024, kBbRed = kBbCs + 2 * 1024;
constexpr size_t kBbSmem = (size_t)kBbRed + 1024 + 1024;
constexpr uint32_t kBbTmemCols = 256;

struct alignas(16) AttnCoreBwdParams {
  const float *dctx;      // [B*L, ldc]
  const float *qkvc;      // [B*L, 816]
  const float *pw, *P, *LOC;
  float *dqkvc;           // [B*L, 816]
  int ldc, B, L;
};

// image pair of a 64-row operand tile whose rows are this thread's COLUMN index: element (row j0 + j, k = r), 32 rows
__device__ __forceinline__ void put_col32(uint32_t big, uint32_t small, int j0, int r, const float (&v)[32]) {
  const uint32_t cb = (uint32_t)(r >> 5) * 0;   // chunk offset is applied by the caller
  (void)cb;
  const int kk = r & 31;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    uint32_t bg, sm;
    tma::split_tf32(v[j], bg, sm);
    const uint32_t o = tma::sw128_offset(j0 + j, kk >> 2) + ((uint32_t)(kk & 3) << 2);
    sts32u(big + o, bg);
    sts32u(small + o, sm);
  }
}

__global__ void __launch_bounds__(kAbThreads, 1)
attn_core_bwd_tc_kernel(const __grid_constant__ AttnCoreBwdParams g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = smem_u32(smem);
  __shared__ __align__(8) uint64_t in_ready, accp_full, ds_ready, acco_full;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y, L = g.L;
  const int row0 = b * L;
  if (tid == 0) {
    mbar_init(&in_ready, 4);
    mbar_init(&accp_full, 1);
    mbar_init(&ds_ready, 4);
    mbar_init(&acco_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&s_tmem, kBbTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t t_dp = s_tmem, t_dq = s_tmem + 64, t_dk = s_tmem + 96, t_dv = s_tmem + 128;
  pdl_sync();
  const bool row_warp = warp >= 4 && (warp & 3) < 2;
  const int qd = warp & 3, half = (warp - 4) >> 2;
  const int r = qd * 32 + lane;
  const float scale = rsqrtf((float)kAbDh);

  if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    mbar_wait(&in_ready, 0);
    fence_after_sync();
    if (elect_one()) {
      const uint32_t idesc = tma::make_idesc_tf32(128, 64, false, false);
      const uint64_t ab = tma::desc_kmajor(sb + kBbGn), as = tma::desc_kmajor(sb + kBbGn + 8192);
      const uint64_t bb = tma::desc_kmajor(sb + kBbVn), bs = tma::desc_kmajor(sb + kBbVn + 8192);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) tma::mma_3x(t_dp, ab + ks * 2, as + ks * 2, bb + ks * 2, bs + ks * 2, idesc, ks ? 1u : 0u);
      mma_commit(&accp_full);
    }
    __syncwarp();
    mbar_wait(&ds_ready, 0);
    fence_after_sync();
    if (elect_one()) {
      const uint32_t idesc = tma::make_idesc_tf32(128, kAbDh, false, false);
      // A tiles: [big chunk 0 | big chunk 1 | small chunk 0 | small chunk 1], 8 KB each; B tiles: the same with 4 KB chunks
      const uint32_t a_of[3] = {kBbDs, kBbDst, kBbPt}, b_of[3] = {kBbKt, kBbQt, kBbGt};
      const uint32_t acc[3] = {t_dq, t_dk, t_dv};
#pragma unroll
      for (int m = 0; m < 3; ++m) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint64_t ab = tma::desc_kmajor(sb + a_of[m] + c * 8192), as = tma::desc_kmajor(sb + a_of[m] + 16384 + c * 8192);
          const uint64_t bb = tma::desc_kmajor(sb + b_of[m] + c * 4096), bs = tma::desc_kmajor(sb + b_of[m] + 8192 + c * 4096);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tma::mma_3x(acc[m], ab + ks * 2, as + ks * 2, bb + ks * 2, bs + ks * 2, idesc, (c | ks) ? 1u : 0u);
        }
      }
      mma_commit(&acco_full);
    }
    __syncwarp();
  } else if (row_warp) {
    // ------------------------------------------------------------------ row warps: thread = row r (query or key)
    const bool row_ok = r < L;
    const float *qrow = g.qkvc + (size_t)(row0 + r) * kAbQkvc;
    const int nk = max(0, min(32, L - half * 32));
    // P row of this thread's half (needed after dP): requested first
    float pr[32];
    {
      const float *prow = g.P + (((size_t)b * kAbHeads + h) * L + r) * L + half * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) pr[j] = (row_ok && j < nk) ? prow[j] : 0.f;
    }
    {
      // operand tiles of the four contractions: half 0 -> q^T, k^T;  half 1 -> v, dctx, dctx^T, cond
      float a[32], c[32];
      const float *p0 = half == 0 ? qrow + h * kAbDh : qrow + 2 * kD + h * kAbDh;                       // q | v
      const float *p1 = half == 0 ? qrow + kD + h * kAbDh : g.dctx + (size_t)(row0 + r) * g.ldc + h * kAbDh;   // k | dctx
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 u = row_ok ? *reinterpret_cast<const float4 *>(p0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 w = row_ok ? *reinterpret_cast<const float4 *>(p1 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        a[j] = u.x; a[j + 1] = u.y; a[j + 2] = u.z; a[j + 3] = u.w;
        c[j] = w.x; c[j + 1] = w.y; c[j + 2] = w.z; c[j + 3] = w.w;
      }
      const uint32_t ch = (uint32_t)(r >> 5) * 4096;
      if (half == 0) {
        put_col32(sb + kBbQt + ch, sb + kBbQt + 8192 + ch, 0, r, a);
        put_col32(sb + kBbKt + ch, sb + kBbKt + 8192 + ch, 0, r, c);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          put4(sb + kBbVn, sb + kBbVn + 8192, r, j, a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
          put4(sb + kBbGn, sb + kBbGn + 8192, r, j, c[4 * j], c[4 * j + 1], c[4 * j + 2], c[4 * j + 3]);
        }
        put_col32(sb + kBbGt + ch, sb + kBbGt + 8192 + ch, 0, r, c);
#pragma unroll
        for (int j = 0; j < kAbCond; ++j) sts32(sb + kBbCs + (uint32_t)(r * 8 + j) * 4u, row_ok ? qrow[3 * kD + h * kAbCond + j] : 0.f);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&in_ready);
    }
    {
      // ---- softmax backward on this thread's row: dl = P o (dP - sum_t dP P)
      mbar_wait(&accp_full, 0);
      fence_after_sync();
      uint32_t acc[32];
      tmem_ld32(t_dp + ((uint32_t)(qd * 32) << 16) + (uint32_t)(half * 32), acc);
      tmem_ld_wait();
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) part += __uint_as_float(acc[j]) * pr[j];
      sts32(sb + kBbRed + (uint32_t)(half * 64 + r) * 4u, part);
      tma::named_barrier(2, kAbRowThreads);
      const float rowdot = lds32(sb + kBbRed + (uint32_t)r * 4u) + lds32(sb + kBbRed + (uint32_t)(64 + r) * 4u);
      float dl[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) dl[j] = pr[j] * (__uint_as_float(acc[j]) - rowdot);    // 0 where P = 0 (padded / out of range)
      // dl [l][t]: A of dq (aliases the dP operands: every row warp is past accp_full, i.e. the dP MMAs are complete)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        put4(sb + kBbDs + half * 8192, sb + kBbDs + 16384 + half * 8192, r, j, dl[4 * j], dl[4 * j + 1], dl[4 * j + 2], dl[4 * j + 3]);
      // dl^T and P^T [t][l]: A of dk / dv; rows t = half * 32 + j, column l = r (chunk r / 32)
      const uint32_t chk = (uint32_t)(r >> 5) * 8192;
      put_col32(sb + kBbDst + chk, sb + kBbDst + 16384 + chk, half * 32, r, dl);
      put_col32(sb + kBbPt + chk, sb + kBbPt + 16384 + chk, half * 32, r, pr);
      // plain copy for the pair warps
#pragma unroll
      for (int j = 0; j < 32; ++j) sts32(sb + kBbDl + (uint32_t)(r * kAbZp + half * 32 + j) * 4u, dl[j]);
      fence_before_sync();
      fence_async_smem();
      tma::named_barrier(5, kAbRowThreads + 7 * 32);      // dl table complete: the pair warps may start
      if (lane == 0) mbar_arrive(&ds_ready);
    }
    {
      // ---- dq (row = query), dk, dv (row = key): 16 dims per thread
      mbar_wait(&acco_full, 0);
      fence_after_sync();
      uint32_t aq[16], ak[16], av[16];
      ab_tmem_ld16(t_dq + ((uint32_t)(qd * 32) << 16) + (uint32_t)(half * 16), aq);
      ab_tmem_ld16(t_dk + ((uint32_t)(qd * 32) << 16) + (uint32_t)(half * 16), ak);
      ab_tmem_ld16(t_dv + ((uint32_t)(qd * 32) << 16) + (uint32_t)(half * 16), av);
      tmem_ld_wait();
      if (row_ok) {
        float *o = g.dqkvc + (size_t)(row0 + r) * kAbQkvc + h * kAbDh + half * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          *reinterpret_cast<float4 *>(o + j) = make_float4(__uint_as_float(aq[j]) * scale, __uint_as_float(aq[j + 1]) * scale,
                                                           __uint_as_float(aq[j + 2]) * scale, __uint_as_float(aq[j + 3]) * scale);
          *reinterpret_cast<float4 *>(o + kD + j) = make_float4(__uint_as_float(ak[j]) * scale, __uint_as_float(ak[j + 1]) * scale,
                                                                __uint_as_float(ak[j + 2]) * scale, __uint_as_float(ak[j + 3]) * scale);
          *reinterpret_cast<float4 *>(o + 2 * kD + j) = make_float4(__uint_as_float(av[j]), __uint_as_float(av[j + 1]),
                                                                    __uint_as_float(av[j + 2]), __uint_as_float(av[j + 3]));
        }
      }
      fence_before_sync();
    }
  } else {
    // ------------------------------------------------------------------ pair warps (0, 2, 3, 6, 7, 10, 11): d cond
    const int pwi = warp == 0 ? 0 : (warp < 4 ? warp - 1 : (warp < 8 ? warp - 3 : warp - 5));     // 0..6
    const int t0 = lane, t1 = lane + 32;
    const bool ok0 = t0 < L, ok1 = t1 < L;
    const float *pws = g.pw + (size_t)b * L * L * 5;
    const float *locs = g.LOC + ((size_t)b * kAbHeads + h) * L * L;
    float ga[5], gb[5], la = 0.f, lb = 0.f;
    auto load_row = [&](int l) {
      const float *pr5 = pws + (size_t)l * L * 5;
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        ga[d] = ok0 ? pr5[t0 * 5 + d] : 0.f;
        gb[d] = ok1 ? pr5[t1 * 5 + d] : 0.f;
      }
      la = ok0 ? locs[(size_t)l * L + t0] : 0.f;
      lb = ok1 ? locs[(size_t)l * L + t1] : 0.f;
    };
    if (pwi < L) load_row(pwi);
    tma::named_barrier(5, kAbRowThreads + 7 * 32);
#pragma unroll 1
    for (int l = pwi; l < L; l += 7) {
      float fa[5], fb[5];
#pragma unroll
      for (int d = 0; d < 5; ++d) { fa[d] = ga[d]; fb[d] = gb[d]; }
      const float lca = la, lcb = lb;
      if (l + 7 < L) load_row(l + 7);
      const float4 sw = lds128(sb + kBbCs + (uint32_t)l * 32u);
      const float sw4 = lds32(sb + kBbCs + (uint32_t)l * 32u + 16u), sw5 = lds32(sb + kBbCs + (uint32_t)l * 32u + 20u);
      const float dla = lds32(sb + kBbDl + (uint32_t)(l * kAbZp + t0) * 4u), dlb = lds32(sb + kBbDl + (uint32_t)(l * kAbZp + t1) * 4u);
      float za = sw.x, zb = sw.x;
      za = fmaf(sw.y, fa[0], za); zb = fmaf(sw.y, fb[0], zb);
      za = fmaf(sw.z, fa[1], za); zb = fmaf(sw.z, fb[1], zb);
      za = fmaf(sw.w, fa[2], za); zb = fmaf(sw.w, fb[2], zb);
      za = fmaf(sw4, fa[3], za);  zb = fmaf(sw4, fb[3], zb);
      za = fmaf(sw5, fa[4], za);  zb = fmaf(sw5, fb[4], zb);
      // d/dz log(clamp(sigmoid(z), 1e-6)) = [loc >= 1e-6] sigmoid(-z)
      const float gza = (ok0 && lca >= 1e-6f) ? dla / (1.f + __expf(za)) : 0.f;
      const float gzb = (ok1 && lcb >= 1e-6f) ? dlb / (1.f + __expf(zb)) : 0.f;
      float acc6[6];
      acc6[0] = gza + gzb;
#pragma unroll
      for (int d = 0; d < 5; ++d) acc6[1 + d] = gza * fa[d] + gzb * fb[d];
#pragma unroll
      for (int d = 0; d < 6; ++d) acc6[d] = warp_sum(acc6[d]);
      if (lane < kAbCond) {
        float v = acc6[0];
#pragma unroll
        for (int d = 1; d < 6; ++d) v = lane == d ? acc6[d] : v;
        g.dqkvc[(size_t)(row0 + l) * kAbQkvc + 3 * kD + h * kAbCond + lane] = v;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    fence_after_sync();
    tmem_dealloc(s_tmem, kBbTmemCols);
  }
}

// big = x rounded to tf32, small = x - big: the operand images of the 3xTF32 products (tma.cuh), written once
// per step for the whole flat parameter buffer so that no GEMM CTA has to split a weight tile again.
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float4 *__restrict__ src, uint4 *__restrict__ big, uint4 *__restrict__ small, long long n4) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    uint4 bg, sm;
    tma::split_tf32(v.x, bg.x, sm.x);
    tma::split_tf32(v.y, bg.y, sm.y);
    tma::split_tf32(v.z, bg.z, sm.z);
    tma::split_tf32(v.w, bg.w, sm.w);
    big[i] = bg;
    small[i] = sm;
  }
}

}  // namespace msr3d

namespace msr3d {
int set_trace_attn_block(long long *buf) {
  MSR3D_CUDA(cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)));
  return MSR3D_OK;
}
}  // namespace msr3d

using namespace msr3d;

/* see include/msr3d_b200.h */
MSR3D_API int msr3d_split_tf32_f32(const float *src, float *big, float *small, long long n, msr3d_stream_t stream) {
  MSR3D_REQUIRE(src && big && small && n >= 0 && (n & 3) == 0, "split_tf32: null pointer or n %% 4 != 0");
  MSR3D_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(big) | reinterpret_cast<uintptr_t>(small)) & 15) == 0,
                "split_tf32: buffers must be 16 B aligned");
  if (n == 0) return MSR3D_OK;
  const long long n4 = n >> 2;
  const int blocks = (int)((n4 + 255) / 256 < (long long)kNumSMs * 4 ? (n4 + 255) / 256 : (long long)kNumSMs * 4);
  launch_pdl(split_tf32_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const float4 *>(src),
             reinterpret_cast<uint4 *>(big), reinterpret_cast<uint4 *>(small), n4);
  return check_launch("split_tf32");
}

MSR3D_API int msr3d_attn_block_supported(int L, int H, int hidden) {
  return (L >= 1 && L <= kAbRT && H == kAbHeads && hidden == kD) ? 1 : 0;
}

static int attn_block_fwd_impl(const float *x, const float *wqkvc_big, const float *wqkvc_small, const float *bqkvc,
                                       const float *wo_big, const float *wo_small, const float *bo, const float *pw,
                                       const unsigned char *key_pad, int B, int L, const float *gamma_a, const float *beta_a,
                                       const float *gamma_1, const float *beta_1, unsigned long long seed,
                                       const unsigned long long *seed_ptr, float p, unsigned site_fc, unsigned site_1,
                                       float *qkvc, float *P, float *LOC, float *ctx, float *xhat_a, float *rstd_a,
                                       float *xhat_1, float *rstd_1, float *y2, float *y2_big, float *y2_small,
                                       msr3d_stream_t stream) {
  MSR3D_REQUIRE(x && wqkvc_big && wqkvc_small && bqkvc && wo_big && wo_small && bo && pw && gamma_a && beta_a && gamma_1 &&
                    beta_1 && qkvc && P && LOC && ctx && xhat_a && rstd_a && xhat_1 && rstd_1 && y2,
                "attn_block_fwd: null pointer");
  MSR3D_REQUIRE(L >= 1 && L <= kAbRT, "attn_block_fwd: L=%d must be in [1,%d]", L, kAbRT);
  MSR3D_REQUIRE(p >= 0.f && p < 1.f, "attn_block_fwd: bad dropout p");
  MSR3D_REQUIRE(((reinterpret_cast<uintptr_t>(bqkvc) | reinterpret_cast<uintptr_t>(bo) | reinterpret_cast<uintptr_t>(qkvc)) & 15) == 0,
                "attn_block_fwd: biases / qkvc must be 16 B aligned");
  if (B <= 0) return MSR3D_OK;
  AttnBlockFwdParams g{};
  const long long M = (long long)B * L;
  int rc = tma::make_map_f32(&g.map_x, x, M, kD, kD, kAbRT, false);
  if (rc != MSR3D_OK) return rc;
  if ((rc = tma::make_map_f32(&g.map_wq_b32, wqkvc_big, kAbQkvc, kD, kD, 32, false)) != MSR3D_OK) return rc;
  if ((rc = tma::make_map_f32(&g.map_wq_s32, wqkvc_small, kAbQkvc, kD, kD, 32, false)) != MSR3D_OK) return rc;
  if ((rc = tma::make_map_f32(&g.map_wq_b8, wqkvc_big, kAbQkvc, kD, kD, 8, false)) != MSR3D_OK) return rc;
  if ((rc = tma::make_map_f32(&g.map_wq_s8, wqkvc_small, kAbQkvc, kD, kD, 8, false)) != MSR3D_OK) return rc;
  if ((rc = tma::make_map_f32(&g.map_wo_b, wo_big, kD, kD, kD, kD, false)) != MSR3D_OK) return rc;
  if ((rc = tma::make_map_f32(&g.map_wo_s, wo_small, kD, kD, kD, kD, false)) != MSR3D_OK) return rc;
  g.x = x; g.bqkvc = bqkvc; g.bo = bo; g.pw = pw; g.key_pad = key_pad;
  g.lna = LnParams{gamma_a, beta_a};
  g.ln1 = LnParams{gamma_1, beta_1};
  g.drop = make_dropout(seed, p, seed_ptr);
  g.site_fc = site_fc; g.site_1 = site_1;
  g.qkvc = qkvc; g.P = P; g.LOC = LOC; g.ctx = ctx;
  g.xhat_a = xhat_a; g.rstd_a = rstd_a; g.xhat_1 = xhat_1; g.rstd_1 = rstd_1; g.y2 = y2;
  g.y2_big = y2_big; g.y2_small = y2_small;
  g.B = B; g.L = L;
  static bool configured = false;
  if (!configured) {
    MSR3D_CUDA(cudaFuncSetAttribute(attn_block_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAbSmem));
    configured = true;
  }
  launch_pdl_cluster(attn_block_fwd_kernel, dim3(kAbHeads, B), dim3(kAbThreads), kAbSmem, (cudaStream_t)stream,
                     dim3(kAbHeads, 1, 1), g);
  return check_launch("attn_block_fwd");
}

MSR3D_API int msr3d_attn_block_fwd_f32(const float *x, const float *wqkvc_big, const float *wqkvc_small, const float *bqkvc,
                                       const float *wo_big, const float *wo_small, const float *bo, const float *pw,
                                       const unsigned char *key_pad, int B, int L, const float *gamma_a, const float *beta_a,
                                       const float *gamma_1, const float *beta_1, unsigned long long seed,
                                       const unsigned long long *seed_ptr, float p, unsigned site_fc, unsigned site_1,
                                       float *qkvc, float *P, float *LOC, float *ctx, float *xhat_a, float *rstd_a,
                                       float *xhat_1, float *rstd_1, float *y2, msr3d_stream_t stream) {
  return attn_block_fwd_impl(x, wqkvc_big, wqkvc_small, bqkvc, wo_big, wo_small, bo, pw, key_pad, B, L, gamma_a, beta_a, gamma_1,
                             beta_1, seed, seed_ptr, p, site_fc, site_1, qkvc, P, LOC, ctx, xhat_a, rstd_a, xhat_1, rstd_1, y2,
                             nullptr, nullptr, stream);
}

MSR3D_API int msr3d_attn_block_fwd_ps_f32(const float *x, const float *wqkvc_big, const float *wqkvc_small, const float *bqkvc,
                                          const float *wo_big, const float *wo_small, const float *bo, const float *pw,
                                          const unsigned char *key_pad, int B, int L, const float *gamma_a,
                                          const float *beta_a, const float *gamma_1, const float *beta_1,
                                          unsigned long long seed, const unsigned long long *seed_ptr, float p,
                                          unsigned site_fc, unsigned site_1, float *qkvc, float *P, float *LOC, float *ctx,
                                          float *xhat_a, float *rstd_a, float *xhat_1, float *rstd_1, float *y2,
                                          float *y2_big, float *y2_small, msr3d_stream_t stream) {
  MSR3D_REQUIRE(y2_big && y2_small, "attn_block_fwd_ps: null image pointer");
  return attn_block_fwd_impl(x, wqkvc_big, wqkvc_small, bqkvc, wo_big, wo_small, bo, pw, key_pad, B, L, gamma_a, beta_a, gamma_1,
                             beta_1, seed, seed_ptr, p, site_fc, site_1, qkvc, P, LOC, ctx, xhat_a, rstd_a, xhat_1, rstd_1, y2,
                             y2_big, y2_small, stream);
}

MSR3D_API int msr3d_attn_bwd_tc_f32(const float *dctx, int ldc, const float *qkvc, int ldq, const float *pw, const float *P,
                                    const float *LOC, int B, int L, int H, float *dqkvc, msr3d_stream_t stream) {
  MSR3D_REQUIRE(dctx && qkvc && pw && P && LOC && dqkvc, "attn_bwd_tc: null pointer");
  MSR3D_REQUIRE(L >= 1 && L <= kAbRT && H == kAbHeads && ldq == kAbQkvc && (ldc & 3) == 0, "attn_bwd_tc: needs L <= 64, 8 heads, ldq = 816");
  MSR3D_REQUIRE(((reinterpret_cast<uintptr_t>(dctx) | reinterpret_cast<uintptr_t>(qkvc) | reinterpret_cast<uintptr_t>(dqkvc)) & 15) == 0,
                "attn_bwd_tc: 16 B alignment");
  if (B <= 0) return MSR3D_OK;
  AttnCoreBwdParams g{};
  g.dctx = dctx; g.qkvc = qkvc; g.pw = pw; g.P = P; g.LOC = LOC; g.dqkvc = dqkvc;
  g.ldc = ldc; g.B = B; g.L = L;
  static bool configured = false;
  if (!configured) {
    MSR3D_CUDA(cudaFuncSetAttribute(attn_core_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBbSmem));
    configured = true;
  }
  launch_pdl(attn_core_bwd_tc_kernel, dim3(kAbHeads, B), dim3(kAbThreads), kBbSmem, (cudaStream_t)stream, g);
  return check_launch("attn_bwd_tc");
}
