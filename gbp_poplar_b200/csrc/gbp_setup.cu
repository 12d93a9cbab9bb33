// Problem setup ON THE DEVICE: the index maps the reference derives on the host with O(V*E) / O(E^2)
// loops -- variable degrees (ba/ba.cpp:514-521), message slots (ba/ba.cpp:267-279) -- and the per-edge records of the
// device layout (gbp_layout.h) are built by kernels from the caller's raw arrays:
//
//   stage A  (before the arena can be sized)   degrees of every camera / landmark by atomics, index range check,
//            warp-tiles per camera, their total and the two maximum degrees -> 16 bytes back to the host
//   stage B  exclusive scans (first edge slot / first warp-tile of every camera, first message of every landmark);
//            two STABLE radix sorts of the edge ids, by camera and by landmark: the position of edge e in the
//            landmark-sorted order IS the place of its landmark-bound message (landmark order, slot = number of
//            earlier edges of that landmark = the reference's message slot made dense), its position in the
//            camera-sorted order minus the camera's first gives its camera slot; then the edge-state records, the
//            warp-tile table, the packed landmark priors, the belief-update blocks and the first observing camera of
//            every landmark (SLAM)
//
// Nothing here is O(E) on the host; the host-side copies of the maps that get_tensor / set_tensor need are rebuilt
// lazily from the device arrays the first time such a call is made (gbp_cuda_api.cu: ensure_host_maps).
#include "gbp_setup.h"

#include <cub/cub.cuh>

#include "gbp_layout.h"

namespace gbp {

namespace {

#define SETUP_TRY(expr)                       \
  do {                                        \
    const cudaError_t e__ = (expr);           \
    if (e__ != cudaSuccess) return (int)e__;  \
  } while (0)

__global__ void k_setup_hist(const uint32_t* __restrict__ cam_ids, const uint32_t* __restrict__ lmk_ids, const uint32_t E, const uint32_t C,
                             const uint32_t L, uint32_t* __restrict__ deg_c, uint32_t* __restrict__ deg_l, uint32_t* __restrict__ info) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const uint32_t c = cam_ids[e], l = lmk_ids[e];
  if (c >= C || l >= L) {
    info[3] = 1u;  // edge index out of range
    return;
  }
  atomicAdd(deg_c + c, 1u);
  atomicAdd(deg_l + l, 1u);
}

// tiles[c] = warp-tiles of camera c; info = {total warp-tiles, max camera degree, max landmark degree, range error}
__global__ void k_setup_degrees(const uint32_t* __restrict__ deg_c, const uint32_t* __restrict__ deg_l, const uint32_t C, const uint32_t L,
                                uint32_t* __restrict__ tiles, uint32_t* __restrict__ info) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t t = 0, mc = 0, ml = 0;
  if (i < C) {
    mc = deg_c[i];
    t = (mc + 31u) / 32u;
    tiles[i] = t;
  }
  if (i < L) ml = deg_l[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t += __shfl_down_sync(0xffffffffu, t, o);
    mc = max(mc, __shfl_down_sync(0xffffffffu, mc, o));
    ml = max(ml, __shfl_down_sync(0xffffffffu, ml, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (t) atomicAdd(info + 0, t);
    if (mc) atomicMax(info + 1, mc);
    if (ml) atomicMax(info + 2, ml);
  }
}

__global__ void k_setup_iota(uint32_t* __restrict__ v, const uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// padding defaults of every edge slot (the edges overwrite theirs in k_setup_records)
__global__ void k_setup_fill(float4* __restrict__ recA, float* __restrict__ var, uint32_t* __restrict__ edge_orig, const uint32_t EP,
                             uint32_t* __restrict__ lmk_first_cam, const uint32_t L) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < EP) {
    recA[s] = make_float4(0.f, __int_as_float(0), __uint_as_float(GBP_FLAG_PAD), 0.f);
    var[s] = 1.f;
    edge_orig[s] = 0xffffffffu;
  }
  if (s < L) lmk_first_cam[s] = 0xffffffffu;
}

// i = position in the camera-sorted / landmark-sorted order; by_cam[i], by_lmk[i] = the edge at that position
__global__ void k_setup_slots(const uint32_t* __restrict__ cam_ids, const uint32_t* __restrict__ by_cam, const uint32_t* __restrict__ by_lmk,
                              const uint32_t* __restrict__ cam_first, const uint32_t* __restrict__ cam_wt_begin, const uint32_t E,
                              uint32_t* __restrict__ pos_of_orig, uint32_t* __restrict__ lpos) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E) return;
  const uint32_t ec = by_cam[i];
  const uint32_t c = cam_ids[ec];
  pos_of_orig[ec] = cam_wt_begin[c] * 32u + (i - cam_first[c]);  // slot = number of earlier edges of that camera
  lpos[by_lmk[i]] = i;                                            // landmark order: message position == sorted position
}

__global__ void k_setup_records(const SetupInputs in, const uint32_t* __restrict__ pos_of_orig, const uint32_t* __restrict__ lpos,
                                float4* __restrict__ recA, float4* __restrict__ recB, float* __restrict__ var,
                                uint32_t* __restrict__ edge_orig, uint32_t* __restrict__ lmk_first_cam) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= in.E) return;
  const uint32_t s = pos_of_orig[e];
  const uint32_t c = in.cam_ids[e], l = in.lmk_ids[e];
  const uint32_t act = in.active_flag ? (in.active_flag[e] == 1u ? 1u : 0u) : 1u;
  recA[s] = make_float4(in.damping ? in.damping[e] : 0.f, __int_as_float(in.damping_count ? in.damping_count[e] : -15),
                        __uint_as_float(act ? GBP_FLAG_ACTIVE : 0u), 0.f);
  const float2 z = reinterpret_cast<const float2*>(in.measurements)[e];
  recB[s] = make_float4(z.x, z.y, __uint_as_float(l), __uint_as_float(lpos[e]));
  var[s] = in.meas_variances[e];
  edge_orig[s] = in.edge_global ? in.edge_global[e] : e;  // quirk Q7 compares GLOBAL edge ids
  atomicMin(lmk_first_cam + l, c);
}

// one thread per camera: its warp-tiles {camera, factors in the tile | warp-tiles of the camera << 8}
__global__ void k_setup_wt_info(const uint32_t* __restrict__ deg_c, const uint32_t* __restrict__ cam_wt_begin, const uint32_t C,
                                uint2* __restrict__ wt_info) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const uint32_t t0 = cam_wt_begin[c], t1 = cam_wt_begin[c + 1], d = deg_c[c];
  for (uint32_t t = t0; t < t1; ++t) wt_info[t] = make_uint2(c, min(32u, d - (t - t0) * 32u) | ((t1 - t0) << 8));
}

// one thread per landmark: packed prior [eta 3 | Lambda 9]; one thread per belief-update block: its record
__global__ void k_setup_lmks(const float* __restrict__ prior_eta, const float* __restrict__ prior_lam, const uint32_t* __restrict__ lmk_ptr,
                             const uint32_t L, const uint32_t per_block, float4* __restrict__ lmk_prior, uint4* __restrict__ lmk_blk) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const float* e = prior_eta + (size_t)l * 3;
  const float* m = prior_lam + (size_t)l * 9;
  float4* o = lmk_prior + (size_t)l * 3;
  o[0] = make_float4(e[0], e[1], e[2], m[0]);
  o[1] = make_float4(m[1], m[2], m[3], m[4]);
  o[2] = make_float4(m[5], m[6], m[7], m[8]);
  if (l % per_block == 0) {
    const uint32_t l1 = min(L, l + per_block);
    lmk_blk[l / per_block] = make_uint4(l, 0u, lmk_ptr[l], lmk_ptr[l1]);   // .y: mask of boundary landmarks, set by k_boundary_records on a shard
  }
}

inline uint32_t blocks(uint64_t n, uint32_t t = 256) { return (uint32_t)((n + t - 1) / t); }
inline int bits_for(uint32_t n) {  // radix-sort key bits for ids in [0, n)
  int b = 1;
  while (b < 32 && (1ull << b) < n) ++b;
  return b;
}

}  // namespace

size_t setup_temp_bytes(uint32_t E, uint32_t C, uint32_t L) {
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)E);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(std::max(C, L) + 1));
  const size_t a = 256;
  auto up = [&](size_t b) { return (b + a - 1) / a * a; };
  // deg_c | tiles | cam_first | deg_l | iota | keys_out | by_cam | by_lmk | lpos | info | cub temp
  return up((size_t)(C + 1) * 4) * 3 + up((size_t)(L + 1) * 4) + up((size_t)E * 4) * 5 + up(64) + up(std::max(sort_bytes, scan_bytes)) + 4096;
}

int setup_stage_a(cudaStream_t s, const SetupInputs& in, char* temp, SetupTemp* t) {
  const uint32_t E = in.E, C = in.C, L = in.L;
  const size_t a = 256;
  auto take = [&](size_t bytes) {
    char* p = temp;
    temp += (bytes + a - 1) / a * a;
    return p;
  };
  t->deg_c = (uint32_t*)take((size_t)(C + 1) * 4);
  t->tiles = (uint32_t*)take((size_t)(C + 1) * 4);
  t->cam_first = (uint32_t*)take((size_t)(C + 1) * 4);
  t->deg_l = (uint32_t*)take((size_t)(L + 1) * 4);
  t->iota = (uint32_t*)take((size_t)E * 4);
  t->keys_out = (uint32_t*)take((size_t)E * 4);
  t->by_cam = (uint32_t*)take((size_t)E * 4);
  t->by_lmk = (uint32_t*)take((size_t)E * 4);
  t->lpos = (uint32_t*)take((size_t)E * 4);
  t->info = (uint32_t*)take(64);
  t->cub_temp = temp;
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)E);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(std::max(C, L) + 1));
  t->cub_bytes = std::max(sort_bytes, scan_bytes);
  // the degree arrays (and their extra last element, which the scans turn into the totals) start from zero
  SETUP_TRY(cudaMemsetAsync(t->deg_c, 0, (size_t)((char*)(t->deg_l + L + 1) - (char*)t->deg_c), s));  // deg_c | tiles | cam_first | deg_l are contiguous
  SETUP_TRY(cudaMemsetAsync(t->info, 0, 64, s));
  if (E) k_setup_hist<<<blocks(E), 256, 0, s>>>(in.cam_ids, in.lmk_ids, E, C, L, t->deg_c, t->deg_l, t->info);
  if (std::max(C, L)) k_setup_degrees<<<blocks(std::max(C, L)), 256, 0, s>>>(t->deg_c, t->deg_l, C, L, t->tiles, t->info);
  SETUP_TRY(cudaGetLastError());
  return 0;
}

int setup_stage_b(cudaStream_t s, const SetupInputs& in, const SetupTemp& t, const SetupOutputs& out) {
  const uint32_t E = in.E, C = in.C, L = in.L;
  size_t tb = t.cub_bytes;
  // first edge / first warp-tile of every camera, first message of every landmark (C + 1 / L + 1 entries: the last is the total)
  SETUP_TRY(cub::DeviceScan::ExclusiveSum(t.cub_temp, tb, t.deg_c, t.cam_first, (int)(C + 1), s));
  tb = t.cub_bytes;
  SETUP_TRY(cub::DeviceScan::ExclusiveSum(t.cub_temp, tb, t.tiles, out.cam_wt_begin, (int)(C + 1), s));
  tb = t.cub_bytes;
  SETUP_TRY(cub::DeviceScan::ExclusiveSum(t.cub_temp, tb, t.deg_l, out.lmk_ptr, (int)(L + 1), s));
  k_setup_fill<<<blocks(std::max(out.E_pad, L)), 256, 0, s>>>(out.recA, out.var, out.edge_orig, out.E_pad, out.lmk_first_cam, L);
  if (E) {
    k_setup_iota<<<blocks(E), 256, 0, s>>>(t.iota, E);
    // stable sorts of the edge ids: by camera, by landmark
    tb = t.cub_bytes;
    SETUP_TRY(cub::DeviceRadixSort::SortPairs(t.cub_temp, tb, in.cam_ids, t.keys_out, t.iota, t.by_cam, (int)E, 0, bits_for(C), s));
    tb = t.cub_bytes;
    SETUP_TRY(cub::DeviceRadixSort::SortPairs(t.cub_temp, tb, in.lmk_ids, t.keys_out, t.iota, t.by_lmk, (int)E, 0, bits_for(L), s));
    k_setup_slots<<<blocks(E), 256, 0, s>>>(in.cam_ids, t.by_cam, t.by_lmk, t.cam_first, out.cam_wt_begin, E, out.pos_of_orig, t.lpos);
    k_setup_records<<<blocks(E), 256, 0, s>>>(in, out.pos_of_orig, t.lpos, out.recA, out.recB, out.var, out.edge_orig, out.lmk_first_cam);
  }
  if (C) k_setup_wt_info<<<blocks(C, 128), 128, 0, s>>>(t.deg_c, out.cam_wt_begin, C, out.wt_info);
  if (L) k_setup_lmks<<<blocks(L), 256, 0, s>>>(in.lmk_priors_eta, in.lmk_priors_lambda, out.lmk_ptr, L, out.lmk_per_block, out.lmk_prior, out.lmk_blk);
  SETUP_TRY(cudaGetLastError());
  return 0;
}

}  // namespace gbp
