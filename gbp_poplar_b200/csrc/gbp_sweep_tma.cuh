// k_sweep_tma: the GBP sweep of every factor with the per-factor records staged by the Blackwell
// copy engines (included by gbp_kernels.cuh; same arithmetic, bit for bit, as k_sweep).
//
// What changed against the cp.async kernel (k_sweep, kept as the reference implementation of the staging):
//   * The 22 contiguous 512-byte rows a warp-tile needs -- factor potential 14, previous camera-bound message 7,
//     edge state 1 -- plus its camera record arrive through TMA: ONE elected lane issues two
//     cp.async.bulk.tensor.2d copies (a [14 x 512 B] and a [7 x 512 B] box of the quad-SoA arrays, tensor maps
//     built at init) and two cp.async.bulk copies, all completing on an mbarrier the warp then waits on.
//     The 26 per-lane LDGSTS with their 64-bit address arithmetic, the wait_group and the register-held
//     landmark records of k_sweep are gone from the instruction stream.
//   * Only the per-factor GATHERS stay per-lane cp.async: the previous landmark-bound message (3 quads at the
//     factor's position in landmark order) and the landmark's belief record (3 quads + the quad of squared
//     mean differences, below).  They live in a single-buffered 7-row region that is consumed at the TOP of a
//     tile (3x3 inverse and the belief-minus-message vectors of the camera-bound message: 15 live registers)
//     and refilled for the next tile right away, a whole tile of arithmetic before it is needed.
//   * dmu (gbp_codelets.cpp:268-277) no longer needs the two variables' previous means per factor: the belief
//     kernel leaves  acc6 = sum over the camera's 6 dofs of (old - new)^2  (sequential from +0, in the camera
//     record) and  sq[i] = (old_i - new_i)^2  of the landmark (lmk_sq), so  dmu = sqrt(((acc6 + sq0) + sq1) + sq2)
//     -- the very same fp32 operations in the same order, nine of them hoisted per variable.
//   * Warp-tiles are handed out by a device-side queue (one atomicAdd per tile, fetched two tiles ahead) after
//     the first two static rounds, so the makespan of a shard does not jump by a whole round when it holds a
//     few tiles more than a multiple of the resident warps.
#pragma once

#include "gbp_tma.cuh"

namespace gbp {

#ifndef GBP_TW
#define GBP_TW 10  // warps per block of k_sweep_tma (one block per SM); measured on config 4: 8 warps (two buffers) 129.6 us, 10 warps (one buffer) 123.9 us, 12 warps 131.5 us
#endif
#define GBP_T_ROWS 22        // TMA rows per buffer: potential 0..13 | camera-bound message 14..20 | edge state 21
#define GBP_T_FAC 0
#define GBP_T_MCAM 14
#define GBP_T_RECA 21
#define GBP_G_ROWS 7         // gathered rows: landmark-bound message 0..2 | landmark belief 3..5 | squared mean differences 6
#define GBP_G_MLMK 0
#define GBP_G_LB 3
#define GBP_G_SQ 6
#define GBP_T_SCAM_QUADS 14  // camera record: belief eta 0..5 | lambda 6..41 | mean 42..47 | previous mean 48..53 | acc6 54 | pad
#define GBP_T_TX_BYTES ((GBP_T_ROWS * 32 + GBP_T_SCAM_QUADS) * 16)
// GBP_TW = 8: two buffers per warp (the reduction runs through the consumed one);
// GBP_TW = 12: one buffer per warp + a reduction scratch of its own (27 rows of GBP_RED_STRIDE floats).
#define GBP_T_NBUF (GBP_TW <= 8 ? 2 : 1)
#define GBP_T_RED_QUADS (GBP_T_NBUF == 1 ? (27 * GBP_RED_STRIDE / 4) : 0)
// per warp: [NBUF buffers x 22 rows x 32 quads | 7 gathered rows x 32 quads | NBUF camera records | mbarriers | scratch]
// (rounded up to whole 128-byte lines: the destination of a tensor copy must be 128-byte aligned)
#define GBP_T_WARP_QUADS ((GBP_T_NBUF * GBP_T_ROWS * 32 + GBP_G_ROWS * 32 + GBP_T_NBUF * GBP_T_SCAM_QUADS + 1 + GBP_T_RED_QUADS + 7) / 8 * 8)
#define GBP_T_SMEM (GBP_TW * GBP_T_WARP_QUADS * 16)

template <int Q0, int N>
GBP_DEV void rows_read(const float4* rows, uint32_t lane, float (&out)[N * 4]) {
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const float4 v = rows[(Q0 + q) * 32 + lane];
    out[q * 4] = v.x; out[q * 4 + 1] = v.y; out[q * 4 + 2] = v.z; out[q * 4 + 3] = v.w;
  }
}

// one elected lane: everything of warp-tile wt that is contiguous -> buffer `tb` / camera record `sc`
GBP_DEV void tma_issue_tile(const DeviceGraph& g, const SweepMaps& maps, float4* tb, float4* sc, uint64_t* bar, const uint32_t wt,
                            const uint32_t cam) {
  mbar_expect_tx(bar, GBP_T_TX_BYTES);
  tma_load_rows(tb + GBP_T_FAC * 32, &maps.fac, wt, bar);
  tma_load_rows(tb + GBP_T_MCAM * 32, &maps.mcam, wt, bar);
  bulk_load(tb + GBP_T_RECA * 32, g.recA + (size_t)wt * 32, 512u, bar);
  bulk_load(sc, g.cam_rec + (size_t)cam * 16, GBP_T_SCAM_QUADS * 16u, bar);
}

// every lane: the gathered records of ITS factor of the next warp-tile (landmark l, message position lpos)
GBP_DEV void gather_issue(const DeviceGraph& g, float4* gat, const uint32_t l, const uint32_t lpos, const uint32_t lane) {
#pragma unroll
  for (int q = 0; q < GBP_MLMK_QUADS; ++q) cp_async16(gat + (GBP_G_MLMK + q) * 32 + lane, g.mlmk + (size_t)lpos * GBP_MLMK_QUADS + q);
#pragma unroll
  for (int q = 0; q < 3; ++q) cp_async16(gat + (GBP_G_LB + q) * 32 + lane, g.lmk_b + (size_t)l * GBP_LMKB_QUADS + q);
  cp_async16(gat + GBP_G_SQ * 32 + lane, g.lmk_sq + l);
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}

// staged factor record: rows 0..8 = eta 0..8 | ll(lower) 9..14 | cl 15..32 | cc(lower) 0..2 at 33..35, rows 9..13 = cc(lower) 3..20 | pad
#define GBP_LLF(i, j) head[GBP_FAC_LL + gbp_sym(i, j)]
#define GBP_CCF(i, j) ((gbp_sym(i, j) < 3) ? head[GBP_FAC_CC + gbp_sym(i, j)] : tail[gbp_sym(i, j) - 3])

// Front half of the camera-bound message (gbp_codelets.cpp:446-462): everything that reads the landmark belief and
// the previous landmark-bound message -- the 3x3 inverse of (Lambda_ll + belief - message) and eta_l + belief - message.
// Afterwards the gathered rows are dead.  ple = previous f->lmk eta (needed again by the landmark-bound damping).
GBP_DEV void cam_msg_front(const float4* tb, const float4* gat, const uint32_t lane, float (&Li)[9], float (&ed)[3], float (&ple)[3]) {
  float pl[12], lb[12];
  rows_read<GBP_G_MLMK, 3>(gat, lane, pl);
  rows_read<GBP_G_LB, 3>(gat, lane, lb);
  float head[16];  // eta 0..8 | ll(lower) 9..14
  rows_read<GBP_T_FAC, 4>(tb, lane, head);
  float Ld[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Ld[i * 3 + j] = fs(fa(GBP_LLF(i, j), lb[3 + i * 3 + j]), pl[3 + i * 3 + j]);
  inv3(Ld, Li);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    ed[i] = fs(fa(head[GBP_FAC_ETA + 6 + i], lb[i]), pl[i]);
    ple[i] = pl[i];
  }
}

// Back half (gbp_codelets.cpp:446-462, 619-627): Schur complement over the landmark block with the inverse from
// cam_msg_front.  nc: eta 0..5 | lower lambda 6..26 | pad; ncu: the strict upper triangle (row-major, i<j).
template <bool UPPER>
GBP_DEV void cam_msg_back(const float4* tb, const uint32_t lane, const float (&Li)[9], const float (&ed)[3], const float damping,
                          float (&nc)[28], float (&ncu)[16]) {
  const float omd = fs(1.0f, damping);
  float head[36], tail[20];
  rows_read<GBP_T_FAC, 9>(tb, lane, head);
  rows_read<GBP_T_FAC + 9, 5>(tb, lane, tail);
  const float* eta = head + GBP_FAC_ETA;
  const float* cl = head + GBP_FAC_CL;
  float pc[8];  // previous f->cam eta 0..5
  rows_read<GBP_T_MCAM, 2>(tb, lane, pc);
  float P[18];  // Lambda_cl * inv (6x3)
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      P[i * 3 + j] = fa(fa(fm(cl[i * 3], Li[j]), fm(cl[i * 3 + 1], Li[3 + j])), fm(cl[i * 3 + 2], Li[6 + j]));
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float acc = fa(fa(fm(P[i * 3], ed[0]), fm(P[i * 3 + 1], ed[1])), fm(P[i * 3 + 2], ed[2]));
    const float h = fs(eta[i], acc);
    nc[i] = fa(fm(h, omd), fm(pc[i], damping));
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      if (i >= j || UPPER) {
        const float acc = fa(fa(fm(P[i * 3], cl[j * 3]), fm(P[i * 3 + 1], cl[j * 3 + 1])), fm(P[i * 3 + 2], cl[j * 3 + 2]));
        const float v = fs(GBP_CCF(i, j), acc);
        if (i >= j) nc[GBP_MCAM_LOWER + lt(i, j)] = v;
        else ncu[gbp_upper(i, j)] = v;
      } else {
        ncu[gbp_upper(i, j)] = 0.f;
      }
    }
  nc[27] = 0.f;
  ncu[15] = 0.f;
}

// Landmark-bound message (gbp_codelets.cpp:536-552, 691-699): Schur complement over the camera block, one inv6x6.
// sc: camera belief eta 0..5 | lambda 6..41.  ple: previous f->lmk eta.
GBP_DEV void lmk_msg(const float4* tb, const float* sc, const uint32_t lane, const float damping, const float (&ple)[3],
                     float (&nl)[12]) {
  const float omd = fs(1.0f, damping);
  float head[36], tail[20];
  rows_read<GBP_T_FAC, 9>(tb, lane, head);
  rows_read<GBP_T_FAC + 9, 5>(tb, lane, tail);
  const float* eta = head + GBP_FAC_ETA;
  const float* cl = head + GBP_FAC_CL;
  float pc[28];  // previous f->cam message: eta 0..5, lower lambda 6..26
  rows_read<GBP_T_MCAM, 7>(tb, lane, pc);
  float Ai[36];
  {
    float Ld[21];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) Ld[lt(i, j)] = fs(fa(GBP_CCF(i, j), sc[6 + i * 6 + j]), pc[GBP_MCAM_LOWER + lt(i, j)]);
    inv6(Ld, Ai);
  }
  float P[18];  // Lambda_lc * inv  (3x6), Lambda_lc(i,k) = Lambda_cl(k,i)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      float acc = fm(cl[i], Ai[j]);
#pragma unroll
      for (int k = 1; k < 6; ++k) acc = fa(acc, fm(cl[k * 3 + i], Ai[k * 6 + j]));
      P[i * 6 + j] = acc;
    }
  float ed[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) ed[i] = fs(fa(eta[i], sc[i]), pc[i]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = fm(P[i * 6], ed[0]);
#pragma unroll
    for (int k = 1; k < 6; ++k) acc = fa(acc, fm(P[i * 6 + k], ed[k]));
    const float h = fs(eta[6 + i], acc);
    nl[i] = fa(fm(h, omd), fm(ple[i], damping));
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float acc = fm(P[i * 6], cl[j]);
#pragma unroll
      for (int k = 1; k < 6; ++k) acc = fa(acc, fm(P[i * 6 + k], cl[k * 3 + j]));
      nl[3 + i * 3 + j] = fs(GBP_LLF(i, j), acc);
    }
}
#undef GBP_LLF
#undef GBP_CCF

// PrepMessageVertex (gbp_codelets.cpp:241-378) from the hoisted per-variable sums.  Out of line on the paths that
// need more than them: an edge whose oldmu is still the streamed one (first sweep after init / set_tensor) and the
// accumulating relinearisation itself.
__device__ __noinline__ float dmu_from_edge_oldmu(const DeviceGraph g, const size_t e, const uint32_t l, const float c0, const float c1,
                                                  const float c2, const float c3, const float c4, const float c5) {
  const float x[9] = {c0, c1, c2, c3, c4, c5, 0.f, 0.f, 0.f};
  const float4 m = g.lmk_b[(size_t)l * GBP_LMKB_QUADS + 3];
  const float xl[3] = {m.x, m.y, m.z};
  float acc = 0.f;  // gbp_codelets.cpp:268-277
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float old = g.oldmu_edge ? g.oldmu_edge[(size_t)i * g.E_pad + e] : 0.f;
    const float d = fs(old, i < 6 ? x[i] : xl[i - 6]);
    acc = fa(acc, fm(d, d));
  }
  return acc;
}

GBP_DEV void prep_factor_tma(const DeviceGraph& g, float4* tb, const float* sc, const float4 sq, const uint32_t cam, const uint32_t l,
                             const size_t e, const uint32_t lane, float& damping, int& dcount, uint32_t& flags, float& dmu) {
  if (dcount == 0) damping = g.hp.maxeta_damping;  // gbp_codelets.cpp:245-248
  dcount += 1;
  float acc;
  if (flags & GBP_FLAG_MUVALID) {
    acc = fa(fa(fa(sc[54], sq.x), sq.y), sq.z);
  } else {
    acc = dmu_from_edge_oldmu(g, e, l, sc[42], sc[43], sc[44], sc[45], sc[46], sc[47]);
  }
  dmu = __fsqrt_rn(acc);
  flags |= GBP_FLAG_MUVALID;
  if (dmu < g.hp.dmu_threshold && dcount > g.hp.min_linear_iters - g.hp.num_undamped_iters) {
    damping = 0.0f;  // gbp_codelets.cpp:280-283
    dcount = -g.hp.num_undamped_iters;
    // the staged record is the current potential: accumulate onto it (quirk Q1), write it back
    const float4 m = g.lmk_b[(size_t)l * GBP_LMKB_QUADS + 3];
    const float4 rb = g.recB[e];
    const uint32_t robust = relinearise_record(tb + lane, 32, g.fac + e, g.E_pad, tb + lane, g.cam_lin + (size_t)cam * 5,
                                               make_float4(g.K[0], g.K[1], g.K[2], g.K[3]), g.hp.Nstds, rb.x, rb.y, g.var[e],
                                               sc[42], sc[43], sc[44], sc[45], sc[46], sc[47], m.x, m.y, m.z);
    flags = (flags & ~GBP_FLAG_ROBUST) | (robust ? GBP_FLAG_ROBUST : 0u);
    const uint32_t m_act = __activemask();
    if (lane == (uint32_t)__ffs(m_act) - 1u) atomicAdd(g.relin_ring + (g.relin_ring[GBP_RELIN_RING] % GBP_RELIN_RING), __popc(m_act));
  }
}

// Lane-ordered sums of the 32 camera-bound messages through `red` (rows of GBP_RED_STRIDE floats, one row per message
// entry): one pass over [eta 6 | lower 21]; with UPPER a second pass over the 15 strictly upper entries.  Every entry is
// summed over the lanes in lane order by ONE lane (row_sum_lane_order), so the result does not depend on the split.
template <bool UPPER>
GBP_DEV void reduce_cam_messages(float* red, const uint32_t lane, const float (&nc)[28], const float (&ncu)[16], float* __restrict__ out42) {
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 27; ++k) red[k * GBP_RED_STRIDE + lane] = nc[k];
  __syncwarp();
  warp_cam_reduce<false>(red, lane, out42);
  if (UPPER) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 15; ++k) red[k * GBP_RED_STRIDE + lane] = ncu[k];
    __syncwarp();
    if (lane < 15) {
      const float acc = row_sum_lane_order(red + lane * GBP_RED_STRIDE);
      // entry `lane` of the row-major strict upper triangle (i < j) inside [eta 6 | Lambda 36 row-major]
      const uint32_t i = (lane >= 14) ? 4u : (lane >= 12) ? 3u : (lane >= 9) ? 2u : (lane >= 5) ? 1u : 0u;
      const uint32_t j = lane - (5u * i - i * (i - 1u) / 2u) + i + 1u;
      out42[6 + i * 6 + j] = acc;
    }
  }
}

// One warp-tile from a landed buffer.  lid = {landmark, message position} of this lane's factor; lid_n the same for
// the warp's next tile (gather issued here once the gathered rows are consumed).  refill(): called once every lane has
// read the last staged value (single-buffered configuration: the copy engine may overwrite the buffer from here on).
template <bool PREP, bool MSG, bool UPPER, class Refill>
GBP_DEV void sweep_tile_tma(const DeviceGraph& g, float4* tb, const float* sc, float4* gat, float* red, const uint32_t wt, const uint2 ti,
                            const uint2 lid, const bool has_next, const uint2 lid_n, const uint32_t lane, Refill&& refill) {
  const size_t e = (size_t)wt * 32 + lane;
  const bool valid = lane < (ti.y & 0xffu);  // padding slots hold no factor (upper bits: warp-tiles of the camera)
  const float4 ra = tb[GBP_T_RECA * 32 + lane];
  float damping = ra.x;
  int dcount = __float_as_int(ra.y);
  uint32_t flags = __float_as_uint(ra.z);
  float dmu = ra.w;
  const bool active = valid && (flags & GBP_FLAG_ACTIVE) != 0;
  const size_t lpos = lid.y;

  if (PREP && active) prep_factor_tma(g, tb, sc, gat[GBP_G_SQ * 32 + lane], ti.x, lid.x, e, lane, damping, dcount, flags, dmu);

  float Li[9], ed3[3], ple[3];
  if (MSG) cam_msg_front(tb, gat, lane, Li, ed3, ple);
  // the gathered rows of this tile are consumed (every lane reads and refills only its own column)
  if (has_next) gather_issue(g, gat, lid_n.x, lid_n.y, lane);

  float nc[28];   // new f->cam message record: eta 0..5 | lower lambda 6..26 | pad
  float ncu[16];  // its strict upper triangle (row-major, i<j): only summed into the camera partial
  float nl[12];   // new f->lmk message
  if (MSG) {
    if (active) {
      lmk_msg(tb, sc, lane, damping, ple, nl);
      cam_msg_back<UPPER>(tb, lane, Li, ed3, damping, nc, ncu);
    } else {
      // inactive (or padding) slot: its messages are zero (gbp_codelets.cpp:464-468 etc.)
#pragma unroll
      for (int k = 0; k < 28; ++k) nc[k] = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) ncu[k] = 0.f;
#pragma unroll
      for (int k = 0; k < 12; ++k) nl[k] = 0.f;
    }
  }
  // generic accesses to the staged rows end here (a relinearising lane also WROTE its potential into them)
  fence_proxy_async();
  __syncwarp();
  refill();
  if (MSG) {
    if (active || (valid && (flags & GBP_FLAG_HASMSG))) {
      float4* p = g.mlmk + lpos * GBP_MLMK_QUADS;
#pragma unroll
      for (int q = 0; q < 3; ++q) p[q] = make_float4(nl[q * 4], nl[q * 4 + 1], nl[q * 4 + 2], nl[q * 4 + 3]);
      store_cam_message(g, e, nc, ncu);
      flags = active ? (flags | GBP_FLAG_HASMSG) : (flags & ~GBP_FLAG_HASMSG);
    }
  }
  // the state record only changes in this kernel when prep ran here or the has-message flag toggled
  if (valid && (PREP ? (active || MSG) : flags != __float_as_uint(ra.z)))
    g.recA[e] = make_float4(damping, __int_as_float(dcount), __uint_as_float(flags), dmu);
  if (MSG) reduce_cam_messages<UPPER>(red, lane, nc, ncu, g.cam_partial + (size_t)wt * GBP_CAMPART_STRIDE);
}

// NBUF = 2: the next tile's rows are requested at the top of a tile into the other buffer (a whole tile of arithmetic
// ahead); the lane-ordered reduction runs through the consumed buffer.  8 warps per SM.
// NBUF = 1: ONE buffer per warp, refilled as soon as the message functions have read it -- while the warp stores its
// results and reduces them through a scratch of its own; the latency of the refill is covered by the other warps of
// the scheduler, of which there are three instead of two: 12 warps per SM at <= 168 registers.
template <bool PREP, bool MSG, bool UPPER>
__global__ void __launch_bounds__(GBP_TW * 32, 1) k_sweep_tma(const DeviceGraph g, const __grid_constant__ SweepMaps maps) {
  extern __shared__ __align__(1024) float4 smem4[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* wbase = smem4 + (size_t)warp * GBP_T_WARP_QUADS;
  float4* tbuf = wbase;                                           // [NBUF][22 rows][32]
  float4* gat = wbase + GBP_T_NBUF * GBP_T_ROWS * 32;             // [7 rows][32]
  float4* scam = gat + GBP_G_ROWS * 32;                           // [NBUF][14]
  uint64_t* bars = reinterpret_cast<uint64_t*>(scam + GBP_T_NBUF * GBP_T_SCAM_QUADS);  // [NBUF]
  float* red_own = reinterpret_cast<float*>(scam + GBP_T_NBUF * GBP_T_SCAM_QUADS + 1);  // [27 rows][36] (NBUF == 1 only)
  const uint32_t n_wt = g.E_pad / 32;
  const uint32_t n_static = gridDim.x * GBP_TW;           // tiles of one static round
  // consecutive warp-tiles go to different SMs, so small graphs spread over the whole chip
  uint32_t wt = warp * gridDim.x + blockIdx.x;
  if (wt >= n_wt) return;
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < GBP_T_NBUF; ++b) mbar_init(bars + b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  fence_proxy_async();
  __syncwarp();

  // prologue: everything of the first warp-tile, ids of the second, a ticket for the third
  uint2 ti = __ldg(g.wt_info + wt);
  // {landmark id, message position} of a lane's factor = the second half of its recB record
  const uint2* lrec = reinterpret_cast<const uint2*>(g.recB) + 1;
  uint2 lid = __ldg(lrec + 2 * ((size_t)wt * 32 + lane));
  if (lane == 0) tma_issue_tile(g, maps, tbuf, scam, bars, wt, ti.x);
  gather_issue(g, gat, lid.x, lid.y, lane);
  uint32_t wt_n = wt + n_static;  // the second round is static too
  if (wt_n >= n_wt) wt_n = 0xffffffffu;
  uint2 ti_n = make_uint2(0u, 0u), lid_n = make_uint2(0u, 0u);
  uint32_t ticket = 0;
  bool drawn = false;
  if (wt_n != 0xffffffffu) {
    ti_n = __ldg(g.wt_info + wt_n);
    lid_n = __ldg(lrec + 2 * ((size_t)wt_n * 32 + lane));
    if (g.tile_queue) {
      ticket = queue_draw(g, lane);
      drawn = true;
    }
  }
  uint32_t buf = 0, phase = 0;  // phase: bit b = parity the next wait on buffer b expects
  for (;;) {
    const bool has_next = wt_n != 0xffffffffu;
    uint32_t wt_nn = 0xffffffffu;
    uint2 ti_nn = make_uint2(0u, 0u), lid_nn = make_uint2(0u, 0u);
    if (has_next) {
      if (GBP_T_NBUF == 2 && lane == 0)  // the other buffer was released at the end of the previous tile
        tma_issue_tile(g, maps, tbuf + (buf ^ 1) * GBP_T_ROWS * 32, scam + (buf ^ 1) * GBP_T_SCAM_QUADS, bars + (buf ^ 1), wt_n, ti_n.x);
      if (g.tile_queue) {
        if (drawn) wt_nn = queue_resolve(ticket, n_wt, n_static);   // drawn one tile ago
        drawn = wt_nn != 0xffffffffu;
        if (drawn) ticket = queue_draw(g, lane);
      } else {  // static round-robin (GBP_TILE_QUEUE=0)
        wt_nn = wt_n + n_static;
        if (wt_nn >= n_wt) wt_nn = 0xffffffffu;
      }
      if (wt_nn != 0xffffffffu) {
        ti_nn = __ldg(g.wt_info + wt_nn);
        lid_nn = __ldg(lrec + 2 * ((size_t)wt_nn * 32 + lane));
      }
    }
    // this tile: the bulk copies complete on the buffer's mbarrier, the gathers on the lane's cp.async group
    mbar_wait(bars + buf, (phase >> buf) & 1u);
    phase ^= 1u << buf;
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    float4* tb = tbuf + buf * GBP_T_ROWS * 32;
    float4* sc4 = scam + buf * GBP_T_SCAM_QUADS;
    sweep_tile_tma<PREP, MSG, UPPER>(g, tb, reinterpret_cast<const float*>(sc4), gat,
                                     GBP_T_NBUF == 1 ? red_own : reinterpret_cast<float*>(tb), wt, ti, lid, has_next, lid_n, lane, [&]() {
                                       if (GBP_T_NBUF == 1 && has_next && lane == 0) tma_issue_tile(g, maps, tb, sc4, bars + buf, wt_n, ti_n.x);
                                     });
    if (GBP_T_NBUF == 2) {
      // the reduction wrote through this buffer; the copy engine rewrites it two tiles from now
      fence_proxy_async();
      __syncwarp();
    }
    if (!has_next) break;
    wt = wt_n; wt_n = wt_nn;
    ti = ti_n; ti_n = ti_nn;
    lid = lid_n; lid_n = lid_nn;
    if (GBP_T_NBUF == 2) buf ^= 1;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  if (g.tile_queue) tile_queue_done(g, lane, n_wt, n_static);
}

}  // namespace gbp
