// Blackwell copy-engine primitives (sm_100a): mbarrier, bulk and tensor copies into shared memory.
// cp.async.bulk* shows up as UBLKCP / UTMALDG in SASS, the mbarrier operations as SYNCS.*.
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <stdint.h>

#include "gbp_math.cuh"

namespace gbp {

GBP_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

GBP_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
GBP_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
GBP_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// generic-proxy accesses of this thread to shared memory are ordered before later async-proxy (TMA) writes
GBP_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// rows [row0, row0 + box rows) x 512 bytes of a quad-SoA array, for warp-tile wt -> dst
GBP_DEV void tma_load_rows(void* dst, const CUtensorMap* map, uint32_t wt, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"((int)(wt * 128u)), "r"(0), "r"(smem_u32(bar))
      : "memory");
}
GBP_DEV void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct SweepMaps {  // TMA descriptors of the two big quad-SoA arrays (box = [rows x 128 floats])
  CUtensorMap fac;   // [14][E_pad * 4] floats, box 14 x 128
  CUtensorMap mcam;  // [7][E_pad * 4] floats, box 7 x 128
};

}  // namespace gbp
