// Hand-written sm_100a tensor-core plumbing: tcgen05.mma (kind::f16, bf16 in /
// f32 accumulate in TMEM), TMEM alloc/load, mbarriers and TMA bulk copies.
// Inline PTX only — no CUTLASS.  Field layouts follow the PTX ISA tables for
// the tcgen05 shared-memory matrix descriptor and instruction descriptor.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace msr3d {
namespace umma {

// ---------------------------------------------------------------- smem operand layout
// K-major, no swizzle ("interleaved") canonical layout, stored k-chunk-major:
// element (r, k) of an R-row bf16 operand tile lives at byte
//        (k / 8) * (R * 16)  +  r * 16  +  (k % 8) * 2.
// A core matrix (8 rows x 16 B) is 128 contiguous bytes; 8-row groups are
// 128 B apart (stride byte offset, SBO); K-adjacent core matrices are R*16 B
// apart (leading byte offset, LBO).  A warp writing one 16 B chunk per lane for
// 32 consecutive rows touches 512 contiguous bytes: conflict-free.
__host__ __device__ constexpr uint32_t tile_bytes(int rows, int k) { return (uint32_t)rows * k * 2; }
__host__ __device__ constexpr uint32_t chunk_offset(int rows, int r, int kchunk) {
  return (uint32_t)kchunk * rows * 16 + (uint32_t)r * 16;
}

// Tiles that are written from registers (fp32 operands split on the fly) use a PADDED K stride:
// K-adjacent core matrices are rows*16 + 16 bytes apart.  With the dense stride every 16 B chunk of
// one row group lands in the same four banks whatever its k index, so a warp whose lanes walk along
// k (the coalesced way to read K-contiguous global rows) would store with 8-way bank conflicts; one
// extra 16 B per k column spreads the eight k chunks of a 64-wide block over all 32 banks.  The
// tensor core does not care: LBO is a descriptor field.
__host__ __device__ constexpr uint32_t lbo_padded(int rows) { return (uint32_t)rows * 16u + 16u; }
__host__ __device__ constexpr uint32_t tile_bytes_padded(int rows, int k) { return lbo_padded(rows) * (uint32_t)(k / 8); }

// 64-bit shared-memory matrix descriptor (tcgen05): bits [0,14) start>>4,
// [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, [61,64) swizzle = 0 (none).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// descriptor of one K=16 slice (two 16 B chunks) of an R-row tile
__device__ __forceinline__ uint64_t kslice_desc(uint32_t tile_addr, int rows, int kstep) {
  return make_desc(tile_addr + (uint32_t)kstep * 2u * rows * 16u, (uint32_t)rows * 16u, 128u);
}

// 32-bit instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16
// (bits 7-9 = 1, 10-12 = 1), both K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a barrier that has not flipped after 2 s means a lost TMA/MMA
// completion (a bug); trap so the host sees a CUDA error instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 2000000000ull) __trap();
    }
  }
}

// one lane of a converged warp (CUTLASS's elect_one_sync): code under `if (elect_one())` compiles to straight
// uniform-datapath instructions (UTCHMMA, UTMALDG back to back); under `if (lane == 0)` the compiler wraps every
// such instruction in an ELECT / BRA.U.ANY serialisation loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- TMA bulk copy (UBLKCP)
// 1-D global -> shared bulk copy completing on an mbarrier (size multiple of 16 B).
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// make generic-proxy smem writes visible to the async proxy (tensor core / TMA)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, one thread issues.
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they complete
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 columns of f32: thread t of the warp receives lane (base_lane + t),
// columns [col, col+32).  `taddr` = (lane << 16) | column.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- split-bf16 arithmetic
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits survive.
// A*B ~= Ah*Bh + Ah*Bl + Al*Bh (the dropped Al*Bl term is ~2^-16 relative), all
// accumulated in f32 by the tensor core: three bf16 MMAs give near-fp32 results.
// cvt.rn.bf16x2.f32 packs two values per instruction; widening back is a shift / mask.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_k, float hi_k) {  // lower k in the low half
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_k), "f"(lo_k));
  return r;
}
__device__ __forceinline__ float bf16_lo_as_f32(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi_as_f32(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// packs two consecutive-k values into one 32-bit word per split level
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t &hi, uint32_t &lo) {
  hi = pack_bf16x2(a, b);
  lo = pack_bf16x2(a - bf16_lo_as_f32(hi), b - bf16_hi_as_f32(hi));
}
// three-way split x = h + m + l (24 mantissa bits): with the six products hh, hm, mh, hl, lh, mm the
// tensor-core result is fp32-grade even for badly conditioned sums (weight gradients).
__device__ __forceinline__ void split_pack2_3way(float a, float b, uint32_t &h, uint32_t &m, uint32_t &l) {
  h = pack_bf16x2(a, b);
  const float ar = a - bf16_lo_as_f32(h), br = b - bf16_hi_as_f32(h);
  m = pack_bf16x2(ar, br);
  l = pack_bf16x2(ar - bf16_lo_as_f32(m), br - bf16_hi_as_f32(m));
}

}  // namespace umma
}  // namespace msr3d
