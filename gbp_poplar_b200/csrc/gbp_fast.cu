// OPT-IN contracted arithmetic (gbp_opts.fast_math): the sweep kernel of gbp_sweep_tma.cuh compiled a second time with
// the fp32 primitives of gbp_math.cuh written as plain operators, so that nvcc may contract a * b + c into one FMA
// (the default build spells every operation as __fmul_rn / __fadd_rn to stay bit-identical to the reference's
// codelets).  Same data layout, same staging, same order of operations -- only the rounding of each multiply-add
// differs (one rounding instead of two).  GBP amplifies such differences over many sweeps, so this
// build is NOT bit-comparable with the reference: it is validated at the north-star tolerance instead (one
// teacher-forced sweep <= 1e-4 per block, plateau reprojection error within 1 %: tests/test_fast_math_gpu.py) and
// reported beside the bit-faithful default, never instead of it.
//
// The whole header tree is re-instantiated under another namespace (gbp_fast) so the two builds cannot collide; the
// host hands the DeviceGraph over as bytes (identical layout, checked at compile time against its size).
#define gbp gbp_fast
#define GBP_FAST_MATH 1
#include "gbp_kernels.cuh"
#undef gbp

#include "gbp_fast.h"

namespace {
template <bool PREP, bool UPPER>
int launch(const gbp_fast::DeviceGraph& g, const gbp_fast::SweepMaps& m, unsigned grid, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(gbp_fast::k_sweep_tma<PREP, true, UPPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, GBP_T_SMEM) != cudaSuccess)
      return (int)cudaGetLastError();
    attr = true;
  }
  gbp_fast::k_sweep_tma<PREP, true, UPPER><<<grid, GBP_TW * 32, GBP_T_SMEM, s>>>(g, m);
  return (int)cudaGetLastError();
}
}  // namespace

size_t gbp_fast_graph_bytes() { return sizeof(gbp_fast::DeviceGraph); }

int gbp_fast_launch_sweep(const void* graph_bytes, const void* maps_bytes, int prep, int upper, unsigned grid, cudaStream_t s) {
  gbp_fast::DeviceGraph g;
  gbp_fast::SweepMaps m;
  std::memcpy(&g, graph_bytes, sizeof(g));
  std::memcpy(&m, maps_bytes, sizeof(m));
  if (prep) return upper ? launch<true, true>(g, m, grid, s) : launch<true, false>(g, m, grid, s);
  return upper ? launch<false, true>(g, m, grid, s) : launch<false, false>(g, m, grid, s);
}
