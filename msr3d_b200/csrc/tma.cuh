// Tensor-map TMA (cp.async.bulk.tensor, SASS UTMALDG) plumbing and the tcgen05 kind::tf32 operand
// conventions used by the "3xTF32" GEMMs of the spatial transformer (gemm3x.cu).
//
// Operands are plain row-major fp32 tensors in global memory.  A TMA box of 32 floats (128 B) x R rows
// with CU_TENSOR_MAP_SWIZZLE_128B lands in shared memory as R rows of 128 B whose 16 B chunks are
// XOR-swizzled with (row % 8) — which is exactly the canonical SWIZZLE_128B layout tcgen05.mma reads,
// for BOTH orientations of an operand:
//   K-major   (tensor stored [mn][k], k contiguous):  box = {32 k, R mn rows};   8-row groups 1024 B apart
//             (SBO); a K = 8 slice of kind::tf32 is 32 B: the descriptor start address advances by 32 B.
//   MN-major  (tensor stored [k][mn], mn contiguous): box = {32 mn, 32 k rows}.  For 32-bit operands the
//             tensor core only reads the "128B swizzle with 32B atoms" layout in this orientation
//             (descriptor layout type 1, TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: the 32 B chunks of a
//             128 B row are XOR-swizzled with row % 4): groups of 4 k rows are 512 B apart (SBO), one K = 8
//             MMA consumes two of them, and 32-wide mn blocks are separate boxes LBO bytes apart.
// So transposed operands (W^T in data gradients, X^T / dY^T in weight gradients) need no transposition
// anywhere: only the descriptor bits and the box coordinates change.
//
// fp32-grade accuracy from tf32 tensor cores ("3xTF32"): x = big + small with big = rna_tf32(x),
// small = rna_tf32(x - big); A*B ~= As*Bb + Ab*Bs + Ab*Bb accumulated in fp32 in TMEM (the dropped
// As*Bs term and the representation error are ~2^-22 relative).  The split is done in shared memory,
// in place, on the raw tile TMA delivered: position independent, so it is the same code for every
// layout above.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace msr3d {
namespace tma {

// ---------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();   // resolved once through cudaGetDriverEntryPoint (no link-time libcuda dependency)

// 2-D fp32 row-major tensor [rows][cols] with row pitch `ld` elements; box = 32 columns x box_rows rows,
// 128 B swizzle (16 B atoms for K-major operands, 32 B atoms for MN-major ones), out-of-bounds elements
// read as zero.  Returns an msr3d status.
int make_map_f32(CUtensorMap *map, const float *base, long long rows, long long cols, long long ld, int box_rows,
                 bool mn_major);

#if defined(__CUDACC__)
// ---------------------------------------------------------------- device: TMA load
// box at (c_inner, c_outer) -> smem_dst (1024 B aligned), completes `bytes of the box` on `bar`
__device__ __forceinline__ void load_2d(void *smem_dst, const CUtensorMap *map, int c_inner, int c_outer, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          umma::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(umma::smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// shared -> global tile store / reduce-add (SASS UTMASTG / UTMAREDG): box at (c_inner, c_outer); rows and columns
// outside the tensor are clipped by the hardware.  Bulk-group completion: commit, then wait_read before the
// shared-memory source is reused or the CTA exits.
__device__ __forceinline__ void store_2d(const CUtensorMap *map, const void *smem_src, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(umma::smem_u32(smem_src)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void reduce_add_2d(const CUtensorMap *map, const void *smem_src, int c_inner, int c_outer) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(umma::smem_u32(smem_src)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// named barrier among `nthreads` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void prefetch_map(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---------------------------------------------------------------- device: SWIZZLE_128B descriptors
constexpr uint32_t kRowBytes = 128;              // one swizzle row: 32 floats
constexpr uint32_t kAtomBytes = 1024;            // 8 rows
constexpr uint32_t kBoxKRows = 32;               // k rows of one MN-major box
constexpr uint32_t kMnBoxBytes = kBoxKRows * kRowBytes;   // 4096: LBO of MN-major operands

// byte offset of 16 B chunk `c` (0..7) of row `r` inside a swizzled tile (for tiles written by threads)
__device__ __forceinline__ uint32_t sw128_offset(int r, int c) {
  return (uint32_t)r * kRowBytes + ((uint32_t)(c ^ (r & 7)) << 4);
}
// descriptor: start >> 4 at [0,14), LBO >> 4 at [16,30), SBO >> 4 at [32,46), version 1 at [46,48),
// layout type at [61,64): 2 = SWIZZLE_128B (16 B atoms), 1 = SWIZZLE_128B_BASE32B (32 B atoms)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// K-major tile (rows x 32 k): LBO is ignored by the hardware for swizzled K-major layouts
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr) { return desc_sw128(tile_addr, 16u, kAtomBytes, 2u); }
// MN-major tile (boxes of 32 k rows x 32 mn, kMnBoxBytes apart; groups of 4 k rows 512 B apart)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr) { return desc_sw128(tile_addr, kMnBoxBytes, 512u, 1u); }
// start-address increment (in the 16 B units of the descriptor field) per K = 8 step
__device__ __forceinline__ uint64_t kstep_units(bool mn_major) { return mn_major ? (kAtomBytes >> 4) : (32u >> 4); }

// instruction descriptor, kind::tf32: D = f32 (bits 4-5 = 1), A = B = tf32 (format 2 at bits 7-9 / 10-12),
// a_major / b_major at bits 15 / 16 (1 = MN-major), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the three products of one K = 8 slice, smallest first
__device__ __forceinline__ void mma_3x(uint32_t tmem_d, uint64_t a_big, uint64_t a_small, uint64_t b_big,
                                       uint64_t b_small, uint32_t idesc, uint32_t accumulate) {
  mma_tf32(tmem_d, a_small, b_big, idesc, accumulate);
  mma_tf32(tmem_d, a_big, b_small, idesc, 1u);
  mma_tf32(tmem_d, a_big, b_big, idesc, 1u);
}

using umma::elect_one;

// ---------------------------------------------------------------- device: big / small split
__device__ __forceinline__ uint32_t rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// big = x rounded to tf32 (nearest, ties away: add half an ulp of the 13 dropped bits, mask), small = x - big
// exactly; the tensor core ignores the low 13 bits of `small` (error 2^-21 |x|, the same order as the dropped
// small*small term).  Integer ALU + one FADD: cvt.rna.tf32.f32 runs at 16 lanes/clk/SM and made the split the
// bottleneck of the whole pipeline (0.58 us per 24 KB stage, tools/trace_kernels.py).
__device__ __forceinline__ void split_tf32(float x, uint32_t &big, uint32_t &small) {
  big = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  small = __float_as_uint(x - __uint_as_float(big));
}
// in place: tile[i] = big(x_i), tile[i + small_off] = small(x_i) for the float4s idx, idx + stride, ... < n_f4
__device__ __forceinline__ void split_region(uint8_t *tile, uint32_t small_off, int n_f4, int idx, int stride) {
  for (int i = idx; i < n_f4; i += stride) {
    uint4 *p = reinterpret_cast<uint4 *>(tile) + i;
    const float4 v = *reinterpret_cast<const float4 *>(p);
    uint4 b, s;
    split_tf32(v.x, b.x, s.x);
    split_tf32(v.y, b.y, s.y);
    split_tf32(v.z, b.z, s.z);
    split_tf32(v.w, b.w, s.w);
    *p = b;
    *reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(p) + small_off) = s;
  }
}
#endif  // __CUDACC__

}  // namespace tma
}  // namespace msr3d
