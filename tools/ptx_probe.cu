// Compile-only probe (not part of the library): the tcgen05 forms planned for round 2 assemble with this
// toolchain —  nvcc -gencode arch=compute_100a,code=sm_100a -c tools/ptx_probe.cu && cuobjdump -sass ptx_probe.o | grep UTC
// -> UTCATOMSWS.2CTA (cta_group::2 alloc), UTCHMMA (kind::tf32), UTCHMMA.2CTA, UTCBAR.2CTA.MULTICAST.
#include <stdint.h>
__global__ void k(uint32_t* out) {
  __shared__ uint32_t s_tmem;
  __shared__ uint64_t bar;
  uint32_t sa = (uint32_t)__cvta_generic_to_shared(&s_tmem);
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar);
  // cta_group::2 alloc / dealloc
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(sa), "r"(128u) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  uint32_t t = s_tmem;
  uint64_t da = 0, db = 0; uint32_t idesc = 0;
  // kind::tf32, cta_group::1 and ::2
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" :: "r"(t), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" :: "r"(t), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
  // multicast commit to both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" :: "r"(sb), "h"((uint16_t)3) : "memory");
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(t), "r"(128u) : "memory");
  out[0] = t;
}
