// CUDA kernels of the GBP sweep (sm_100a, fp32 SIMT; tensor cores are
// deliberately unused: the work is ~2.5 kflop of tiny 3x3/6x6 solves per 0.7 KB
// of streamed state, i.e. HBM-bound -- see DESIGN.md).
//
//   k_sweep<PREP,MSG>   one thread per factor, persistent software-pipelined warps.
//                       PREP = PrepMessageVertex        gbp_codelets.cpp:241-378
//                       MSG  = the four message vertices gbp_codelets.cpp:411-709
//                              + on-chip reduction of the camera-bound messages
//   k_update_cams/_lmks belief update (prog_ub, ba/ba.cpp:104-139) fused with the
//                       per-variable mean (inf2mean hoisted out of PrepMessageVertex,
//                       which recomputes it once per adjacent edge, :264-265)
//   k_relinearise_all   RelineariseFactorVertex          gbp_codelets.cpp:38-171
//   k_weaken            WeakenPriorVertex                gbp_codelets.cpp:184-196
//   k_metric/_finish    eval_reprojection_error          ba/util.cpp:74-144
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gbp_layout.h"
#include "gbp_math.cuh"
#include "gbp_tma.cuh"

#define GBP_RELIN_RING 32  // sweeps of relinearisation history kept on the device

namespace gbp {

struct DeviceGraph {
  // sizes
  uint32_t C, L, E, E_pad;  // E_pad: edge slots (multiple of GBP_TILE)
  // per-edge-slot records (quad-SoA, see gbp_layout.h)
  float4* fac;        // [18][E_pad]
  float4* mcam;       // [7][E_pad]
  float4* mcam_up;    // [4][E_pad] strict upper triangle of the camera message Lambda, or nullptr (gbp_opts.store_full_messages)
  float4* mlmk;       // [E][3] factor->landmark messages in LANDMARK order (see gbp_layout.h)
  float4* recA;       // [E_pad] {damping, damping_count, flags, dmu}
  float4* recB;       // [E_pad] {z.x, z.y, landmark id, position of the landmark message}
  float* var;         // [E_pad] meas_variances (only read when a factor relinearises)
  float* oldmu_edge;  // [9][E_pad] or nullptr (= zeros); read only while !MUVALID
  uint32_t* edge_orig;  // [E_pad] original edge id, 0xffffffff for padding
  // tiles
  uint2* wt_info;            // [E_pad/32] per warp-tile {camera, number of real factors (0..32)}
  uint32_t* cam_wt_begin;    // [C+1] first warp-tile of every camera
  float* cam_partial;        // [E_pad/32][42] per-warp-tile sums of the camera-bound messages
  // cameras
  float* cam_b_eta;      // [C][6]
  float* cam_b_lam;      // [C][36]
  float* cam_mean;       // [C][6]
  float* cam_mean_prev;  // [C][6]
  float4* cam_lin;       // [C][5] camera-only linearisation constants of the mean: R 9 | num 9 | den | pad (cam_lin_consts)
  float4* cam_rec;       // [C][16] packed {belief eta 6 | belief lambda 36 | mean 6 | previous mean 6 | pad}: what k_sweep stages
  float* cam_prior_eta;  // [C][6]
  float* cam_prior_lam;  // [C][36]
  float* cam_scaling;    // [C]
  uint32_t* cam_wflag;   // [C]
  // landmarks
  float4* lmk_b;          // [L][4] {eta3, lam9, mean3, pad}
  float4* lmk_mean_prev;  // [L]
  float4* lmk_sq;         // [L] {(old - new)^2 of the three mean components, 0}: the landmark's terms of dmu (k_sweep_tma)
  float4* lmk_prior;      // [L][3] {eta3, lam9}
  float* lmk_scaling;     // [L]
  uint32_t* lmk_wflag;    // [L]
  uint32_t* lmk_ptr;      // [L+1] first message of every landmark in mlmk (messages in original edge order)
  uint4* lmk_blk;         // [n_lmk_blocks] {first landmark, mask of its boundary landmarks (shards), first message, one past the last} of every belief-update block (32 landmarks; staged when <= GBP_LMK_CAP messages)
  uint32_t n_lmk_blocks;
  // multi-GPU shard (all null / 0 on a single-GPU handle): boundary landmarks = landmarks
  // that other ranks observe too; their beliefs are formed from all-gathered partials
  uint32_t* bnd_local;    // [n_bnd_local] local landmark id
  uint32_t* bnd_slot;     // [n_bnd_local] position in the global boundary list
  uint32_t* bnd_span;     // [n_bnd_local] bit r: rank r observes the landmark -- the ranks its partials go to / come from
  uint4* bnd_rec;         // [n_bnd_local] {local landmark, its first message, one past its last, position in the global boundary list}
  float4* bnd_send;       // [n_bnd_global][3]  this rank's partial sums (zero where it has no factor)
  const float4* bnd_recv; // [world][n_bnd_global][3]  all ranks' partial sums
  uint32_t n_bnd_local, n_bnd_global, world, rank;
  // peer-to-peer exchange over NVLink (null = the NCCL all-gather path): every rank maps the
  // receive buffers and arrival flags of all ranks (CUDA IPC)
  uint4** peer_recv;      // [world] -> that rank's receive buffer [2 parities][world][n_bnd_global][3 quads][2] of tagged pairs {a, step, b, step}
  const uint4* p2p_recv;  // this rank's own receive buffer
  uint32_t* p2p_error;    // set when a wait for a peer timed out
  uint32_t* p2p_step;     // [2] {completed exchange steps, blocks of the current k_update_vars that have read it}
  long long p2p_timeout;  // clock64 ticks a block waits for its peers before it gives up (p2p_error)
  // metric exchange over the same peer mappings (no collective call in a sweep that asks for the metric)
  double** peer_metric;       // [world] -> that rank's metric receive buffer [2 parities][world][8]
  uint32_t** peer_mflag;      // [world] -> that rank's metric arrival flags [world]
  const double* metric_recv;  // this rank's own receive buffer
  uint32_t* metric_flag;      // this rank's own arrival flags
  uint32_t* metric_step;      // [1] completed metric exchanges
  uint32_t* relin_list;   // [E] edge slots that relinearise this sweep (compacted by k_prep_pass)
  uint32_t* relin_count;  // [1]
  uint32_t* relin_ring;   // [GBP_RELIN_RING + 1] relinearisations of the last sweeps; [GBP_RELIN_RING] = sweep counter
  uint32_t* tile_queue;   // [2] {tickets handed out, warps done}: warp-tile queue of the sweep kernels
  unsigned long long* dbg_ts;  // [32][8] per-exchange timestamps (diagnostic builds with -DGBP_DEBUG_TS, GBP_DEBUG_TS=1), else null
  float K[4];             // fx fy cx cy
  Hyper hp;
};

GBP_DEV float4 ldg4(const float4* p) { return __ldg(p); }

// ---- L2 residency control --------------------------------------------------------
// Per sweep the factor kernel streams ~0.6 KB per factor that is touched exactly once
// (factor potentials, camera-bound messages) and ~0.1 KB per factor that is touched
// again within microseconds (landmark-bound messages: written here, gathered by the
// belief kernel, read back by the next sweep; the small per-edge state records).  On
// B200 the second class fits the 126 MB L2 for graphs of a few million factors, so it
// is tagged evict_last while the streams are tagged evict_first.
#ifndef GBP_L2_HINTS
#define GBP_L2_HINTS 0  // (measured: no gain on B200, kept for experiments) bitmask: 1 keep landmark messages, 2 keep edge-state records, 4 stream loads, 8 stream stores
#endif
GBP_DEV uint64_t l2_policy_stream() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
GBP_DEV uint64_t l2_policy_keep() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
template <int EN>
GBP_DEV float4 ld4_hint(const float4* ptr, uint64_t pol) {
  if (!EN) return *ptr;
#if GBP_L2_HINTS
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol) : "memory");
  return v;
#else
  return *ptr;
#endif
}
template <int EN>
GBP_DEV void st4_hint(float4* ptr, const float4 v, uint64_t pol) {
  if (!EN) {
    *ptr = v;
    return;
  }
#if GBP_L2_HINTS
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;\n"
               ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
#else
  *ptr = v;
#endif
}

template <int N>
GBP_DEV void load_quads(const float4* base, size_t stride, size_t e, float (&out)[N * 4]) {
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const float4 v = base[(size_t)q * stride + e];
    out[q * 4 + 0] = v.x;
    out[q * 4 + 1] = v.y;
    out[q * 4 + 2] = v.z;
    out[q * 4 + 3] = v.w;
  }
}

template <int N>
GBP_DEV void store_quads(float4* base, size_t stride, size_t e, const float (&in)[N * 4]) {
#pragma unroll
  for (int q = 0; q < N; ++q)
    base[(size_t)q * stride + e] = make_float4(in[q * 4], in[q * 4 + 1], in[q * 4 + 2], in[q * 4 + 3]);
}

// ---- factor record (un)packing ---------------------------------------------------
// Registers hold a factor potential un-packed as f[72] = [eta 9 | ll 9 | cl 18 | cc 36]
// (what linearise_accumulate works on); memory holds the 14-quad packed record of
// gbp_layout.h (lower triangles of the symmetric ll / cc blocks).
#define GBP_F_ETA 0
#define GBP_F_LL 9
#define GBP_F_CL 18
#define GBP_F_CC 36
GBP_DEV void fac_unpack(const float (&r)[GBP_FAC_QUADS * 4], float (&f)[72]) {
#pragma unroll
  for (int i = 0; i < 9; ++i) f[GBP_F_ETA + i] = r[GBP_FAC_ETA + i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) f[GBP_F_LL + i * 3 + j] = r[GBP_FAC_LL + gbp_sym(i, j)];
#pragma unroll
  for (int i = 0; i < 18; ++i) f[GBP_F_CL + i] = r[GBP_FAC_CL + i];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) f[GBP_F_CC + i * 6 + j] = r[GBP_FAC_CC + gbp_sym(i, j)];
}
GBP_DEV void fac_pack(const float (&f)[72], float (&r)[GBP_FAC_QUADS * 4]) {
#pragma unroll
  for (int i = 0; i < 9; ++i) r[GBP_FAC_ETA + i] = f[GBP_F_ETA + i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) r[GBP_FAC_LL + gbp_lt(i, j)] = f[GBP_F_LL + i * 3 + j];
#pragma unroll
  for (int i = 0; i < 18; ++i) r[GBP_FAC_CL + i] = f[GBP_F_CL + i];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) r[GBP_FAC_CC + gbp_lt(i, j)] = f[GBP_F_CC + i * 6 + j];
  r[54] = 0.f;
  r[55] = 0.f;
}

// (Re)linearisation of one factor (gbp_codelets.cpp:285-373 / :53-168): rare and
// register hungry, so it is kept out of line of the streaming path.  `src` is where the
// current packed record is read from (quad q at src[q * src_stride]; nullptr = start from
// zero, RelineariseFactorVertex), the new record goes to global memory and, when `stage`
// is given, into the caller's shared-memory stage slot as well.  Arguments are passed by
// value so the kernel parameter block never has its address taken.
__device__ __noinline__ uint32_t relinearise_record(const float4* src, size_t src_stride, float4* dst, size_t dst_stride,
                                                    float4* stage, const float4* cam_lin, float4 K4, float Nstds, float z0,
                                                    float z1, float var,
                                                    float c0, float c1, float c2, float c3, float c4, float c5, float l0,
                                                    float l1, float l2) {
  float f[72];
  if (src) {  // quirk Q1: accumulate onto the old blocks
    float r[GBP_FAC_QUADS * 4];
#pragma unroll
    for (int q = 0; q < GBP_FAC_QUADS; ++q) {
      const float4 v = src[(size_t)q * src_stride];
      r[q * 4] = v.x; r[q * 4 + 1] = v.y; r[q * 4 + 2] = v.z; r[q * 4 + 3] = v.w;
    }
    fac_unpack(r, f);
  } else {
#pragma unroll
    for (int i = 0; i < 72; ++i) f[i] = 0.f;
  }
  const float K[4] = {K4.x, K4.y, K4.z, K4.w};
  const float x_kf[6] = {c0, c1, c2, c3, c4, c5};
  const float x_l[3] = {l0, l1, l2};
  float(&eta)[9] = *reinterpret_cast<float(*)[9]>(f + GBP_F_ETA);
  float(&ll)[9] = *reinterpret_cast<float(*)[9]>(f + GBP_F_LL);
  float(&cl)[18] = *reinterpret_cast<float(*)[18]>(f + GBP_F_CL);
  float(&cc)[36] = *reinterpret_cast<float(*)[36]>(f + GBP_F_CC);
  float Rn[20];  // R 9 | num 9 | den | pad of this factor's camera
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const float4 v = __ldg(cam_lin + q);
    Rn[q * 4] = v.x; Rn[q * 4 + 1] = v.y; Rn[q * 4 + 2] = v.z; Rn[q * 4 + 3] = v.w;
  }
  const float(&R)[9] = *reinterpret_cast<const float(*)[9]>(Rn);
  const float(&num)[9] = *reinterpret_cast<const float(*)[9]>(Rn + 9);
  const uint32_t robust = linearise_accumulate(z0, z1, var, K, x_kf, x_l, R, num, Rn[18], Nstds, eta, ll, cl, cc);
  float r[GBP_FAC_QUADS * 4];
  fac_pack(f, r);
#pragma unroll
  for (int q = 0; q < GBP_FAC_QUADS; ++q) {
    const float4 v = make_float4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
    dst[(size_t)q * dst_stride] = v;
    if (stage) stage[q * 32] = v;
  }
  return robust;
}

// ---- k_sweep ---------------------------------------------------------------------
// Persistent, software-pipelined: one block of GBP_SW_WARPS warps per SM; every WARP is
// autonomous (no block-wide barrier) and walks over warp-tiles (32 consecutive edge slots
// = 32 factors of one camera, one thread per factor) with a double-buffered
// shared-memory stage:
//   top of iteration t   issue the cp.async (LDGSTS, L1-bypassing) copies of warp-tile
//                        t+1 -- 26 quads per factor: potential 14, previous camera
//                        message 7, previous landmark message 3, the two edge-state
//                        records -- plus its camera record; start the register loads of
//                        the landmark beliefs of t+1 (their landmark ids were fetched
//                        during t-1); fetch tile info / landmark ids of t+2;
//                        wait for the copies of t (issued one iteration ago);
//   body                 PrepMessageVertex on the hoisted per-variable means, both
//                        messages from the staged data, stores straight from registers;
//   end                  lane-ordered sum of the 32 camera-bound messages through the
//                        consumed stage (serial fp32 adds: the result is defined
//                        independently of the hardware), one 42-float partial per warp.
// So every DRAM byte a warp needs is requested a whole tile of arithmetic (~4 us) before
// it is used and the SM never sits in a load phase.
#ifndef GBP_SW_WARPS
#define GBP_SW_WARPS 8   // warps per block (one block per SM)
#endif
#define GBP_NBUF 2       // double-buffered stage
#ifndef GBP_L2_PREFETCH
#define GBP_L2_PREFETCH 0  // 1: bulk, 2: per-line L2 prefetch of warp-tile t+2 (measured: 138 -> 145 / 143 us, kept for experiments)
#endif
#define GBP_WARPS (GBP_TILE / 32)  // warp-tiles per 128-slot tile
#define GBP_SQ 26  // quads per factor in a stage
#define GBP_SQ_FAC 0
#define GBP_SQ_MCAM GBP_FAC_QUADS                      // 14
#define GBP_SQ_MLMK (GBP_FAC_QUADS + GBP_MCAM_QUADS)   // 21
#define GBP_SQ_RECA (GBP_SQ_MLMK + GBP_MLMK_QUADS)     // 24
#define GBP_SQ_RECB (GBP_SQ_RECA + 1)                  // 25
#define GBP_STAGE_QUADS (GBP_SQ * 32)
#define GBP_SCAM 56  // per-warp copy of: belief eta 6 | belief lambda 36 | mean 6 | previous mean 6
#define GBP_RED_STRIDE 36  // floats per row of the reduction scratch: 32 lanes + 4 pad, rows 16-byte aligned for LDS.128
#define GBP_SWEEP_SMEM (GBP_SW_WARPS * GBP_NBUF * (GBP_STAGE_QUADS * 16 + GBP_SCAM * 4))

GBP_DEV void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}

template <int Q0, int N>
GBP_DEV void stage_read(const float4* stage, uint32_t lane, float (&out)[N * 4]) {
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const float4 v = stage[(Q0 + q) * 32 + lane];
    out[q * 4] = v.x; out[q * 4 + 1] = v.y; out[q * 4 + 2] = v.z; out[q * 4 + 3] = v.w;
  }
}

// all per-factor records of warp-tile wt (+ the camera record) -> shared memory, asynchronously
GBP_DEV void issue_stage(const DeviceGraph& g, float4* stage, float* s_cam, const uint32_t wt, const uint32_t cam,
                         const uint32_t lpos, const uint32_t lane) {
  const size_t e = (size_t)wt * 32 + lane;
  cp_async16(stage + GBP_SQ_RECA * 32 + lane, g.recA + e);
  cp_async16(stage + GBP_SQ_RECB * 32 + lane, g.recB + e);
  if (lane < GBP_SCAM / 4) cp_async16(s_cam + lane * 4, g.cam_rec + (size_t)cam * 16 + lane);
#pragma unroll
  for (int q = 0; q < GBP_MLMK_QUADS; ++q) cp_async16(stage + (GBP_SQ_MLMK + q) * 32 + lane, g.mlmk + (size_t)lpos * GBP_MLMK_QUADS + q);
#pragma unroll
  for (int q = 0; q < GBP_MCAM_QUADS; ++q) cp_async16(stage + (GBP_SQ_MCAM + q) * 32 + lane, g.mcam + (size_t)q * g.E_pad + e);
#pragma unroll
  for (int q = 0; q < GBP_FAC_QUADS; ++q) cp_async16(stage + (GBP_SQ_FAC + q) * 32 + lane, g.fac + (size_t)q * g.E_pad + e);
}

// landmark record of one factor: belief eta 0..2 | lambda 3..11 | mean 12..14 | - | previous mean 16..18 | -
GBP_DEV void load_lmk_belief(const DeviceGraph& g, const uint32_t l, float (&lb)[20]) {
  const float4* p = g.lmk_b + (size_t)l * GBP_LMKB_QUADS;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = p[q];
    lb[q * 4] = v.x; lb[q * 4 + 1] = v.y; lb[q * 4 + 2] = v.z; lb[q * 4 + 3] = v.w;
  }
  const float4 mp = g.lmk_mean_prev[l];
  lb[16] = mp.x; lb[17] = mp.y; lb[18] = mp.z; lb[19] = mp.w;
}

// lane-ordered sum of the 42 camera-message values of the warp's 32 factors.
// UPPER == false: only eta and the lower triangle of Lambda (27 values, rows 0..26 of `red` in the order
// [eta 6 | lower 21]) are summed -- one pass over the lanes instead of two; the strict upper triangle of the
// partial is left untouched (see k_sweep).
// one row of the scratch (the 32 lanes' values of one message entry), summed in lane order; eight 16-byte reads
GBP_DEV float row_sum_lane_order(const float* row) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
  float4 v = r4[0];
  float acc = fa(fa(fa(v.x, v.y), v.z), v.w);
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    v = r4[k];
    acc = fa(fa(fa(fa(acc, v.x), v.y), v.z), v.w);
  }
  return acc;
}

template <bool UPPER>
GBP_DEV void warp_cam_reduce(float* red, uint32_t lane, float* __restrict__ out42) {
  if (UPPER) {
    out42[lane] = row_sum_lane_order(red + lane * GBP_RED_STRIDE);
    if (lane < GBP_CAMPART - 32) out42[32 + lane] = row_sum_lane_order(red + (32 + lane) * GBP_RED_STRIDE);
  } else if (lane < 27) {
    const float acc = row_sum_lane_order(red + lane * GBP_RED_STRIDE);
    // position of value `lane` of [eta 6 | lower 21] inside [eta 6 | Lambda 36 row-major]
    uint32_t pos = lane;
    if (lane >= 6) {
      const uint32_t k = lane - 6;
      const uint32_t i = (k >= 15) ? 5u : (k >= 10) ? 4u : (k >= 6) ? 3u : (k >= 3) ? 2u : (k >= 1) ? 1u : 0u;
      pos = 6 + i * 6 + (k - i * (i + 1) / 2);
    }
    out42[pos] = acc;
  }
}

GBP_DEV void store_cam_message(const DeviceGraph& g, const size_t e, const float (&nc)[28], const float (&ncu)[16]) {
#pragma unroll
  for (int q = 0; q < GBP_MCAM_QUADS; ++q)
    g.mcam[(size_t)q * g.E_pad + e] = make_float4(nc[q * 4], nc[q * 4 + 1], nc[q * 4 + 2], nc[q * 4 + 3]);
  if (g.mcam_up) {  // gbp_opts.store_full_messages
#pragma unroll
    for (int q = 0; q < 4; ++q)
      g.mcam_up[(size_t)q * g.E_pad + e] = make_float4(ncu[q * 4], ncu[q * 4 + 1], ncu[q * 4 + 2], ncu[q * 4 + 3]);
  }
}

// ---- the two factor->variable messages, from a landed stage ------------------------
// Staged factor record: quads 0..8 = eta 0..8 | ll(lower) 9..14 | cl 15..32 | cc(lower) 0..2 at 33..35,
// quads 9..13 = cc(lower) 3..20 | pad.
#define GBP_LLF(i, j) head[GBP_FAC_LL + gbp_sym(i, j)]
#define GBP_CCF(i, j) ((gbp_sym(i, j) < 3) ? head[GBP_FAC_CC + gbp_sym(i, j)] : tail[gbp_sym(i, j) - 3])

// Message to the landmark (gbp_codelets.cpp:536-552, 691-699): Schur complement over the camera block,
// one inv6x6; eta damped with the previous message.  s_cam: camera belief eta 0..5 | lambda 6..41.
GBP_DEV void msg_to_landmark(const float4* stage, const float* s_cam, const uint32_t lane, const float damping,
                             float (&nl)[12]) {
  const float omd = fs(1.0f, damping);
  float head[36], tail[20];
  stage_read<GBP_SQ_FAC, 9>(stage, lane, head);
  stage_read<GBP_SQ_FAC + 9, 5>(stage, lane, tail);
  const float* eta = head + GBP_FAC_ETA;
  const float* cl = head + GBP_FAC_CL;
  float pc[28];  // previous f->cam message: eta 0..5, lower lambda 6..26
  stage_read<GBP_SQ_MCAM, 7>(stage, lane, pc);
  const float4 ple = stage[GBP_SQ_MLMK * 32 + lane];  // previous f->lmk eta (x, y, z)
  const float pl[3] = {ple.x, ple.y, ple.z};
  float Ai[36];
  {
    float Ld[21];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j)
        Ld[lt(i, j)] = fs(fa(GBP_CCF(i, j), s_cam[6 + i * 6 + j]), pc[GBP_MCAM_LOWER + lt(i, j)]);
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
  for (int i = 0; i < 6; ++i) ed[i] = fs(fa(eta[i], s_cam[i]), pc[i]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = fm(P[i * 6], ed[0]);
#pragma unroll
    for (int k = 1; k < 6; ++k) acc = fa(acc, fm(P[i * 6 + k], ed[k]));
    const float h = fs(eta[6 + i], acc);
    nl[i] = fa(fm(h, omd), fm(pl[i], damping));
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

// Message to the camera (gbp_codelets.cpp:446-462, 619-627): Schur complement over the landmark block,
// one inv3x3.  nc: eta 0..5 | lower lambda 6..26 | pad; ncu: the strict upper triangle (row-major, i<j).
template <bool UPPER>
GBP_DEV void msg_to_camera(const float4* stage, const uint32_t lane, const float (&lb)[20], const float damping,
                           float (&nc)[28], float (&ncu)[16]) {
  const float omd = fs(1.0f, damping);
  float head[36], tail[20];
  stage_read<GBP_SQ_FAC, 9>(stage, lane, head);
  stage_read<GBP_SQ_FAC + 9, 5>(stage, lane, tail);
  const float* eta = head + GBP_FAC_ETA;
  const float* cl = head + GBP_FAC_CL;
  float pl[12];  // previous f->lmk message: eta 0..2, lambda 3..11
  stage_read<GBP_SQ_MLMK, 3>(stage, lane, pl);
  float pc[8];  // previous f->cam eta 0..5
  stage_read<GBP_SQ_MCAM, 2>(stage, lane, pc);
  float Ld[9], Li[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Ld[i * 3 + j] = fs(fa(GBP_LLF(i, j), lb[3 + i * 3 + j]), pl[3 + i * 3 + j]);
  inv3(Ld, Li);
  float P[18];  // Lambda_cl * inv (6x3)
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      P[i * 3 + j] = fa(fa(fm(cl[i * 3], Li[j]), fm(cl[i * 3 + 1], Li[3 + j])), fm(cl[i * 3 + 2], Li[6 + j]));
  float ed[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) ed[i] = fs(fa(eta[6 + i], lb[i]), pl[i]);
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
#undef GBP_LLF
#undef GBP_CCF

// PrepMessageVertex on the hoisted per-variable means (gbp_codelets.cpp:241-378): damping state
// machine, dmu, conditional accumulating relinearisation (quirk Q1) with Huber (quirk Q2).  The
// staged factor record of a relinearising lane is rewritten in place (and in global memory).
GBP_DEV void prep_factor(const DeviceGraph& g, float4* stage, const float* s_cam, const uint32_t cam, const size_t e,
                         const uint32_t lane, const float (&lb)[20], const float4 rb, float& damping, int& dcount, uint32_t& flags, float& dmu) {
  if (dcount == 0) damping = g.hp.maxeta_damping;  // gbp_codelets.cpp:245-248
  dcount += 1;
  float x_kf[6], x_l[3], old[9];
#pragma unroll
  for (int i = 0; i < 6; ++i) x_kf[i] = s_cam[42 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) x_l[i] = lb[12 + i];
  if (flags & GBP_FLAG_MUVALID) {
#pragma unroll
    for (int i = 0; i < 6; ++i) old[i] = s_cam[48 + i];
    old[6] = lb[16]; old[7] = lb[17]; old[8] = lb[18];
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) old[i] = g.oldmu_edge ? g.oldmu_edge[(size_t)i * g.E_pad + e] : 0.f;
  }
  float acc = 0.f;  // gbp_codelets.cpp:268-277
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float d = fs(old[i], x_kf[i]);
    acc = fa(acc, fm(d, d));
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float d = fs(old[6 + i], x_l[i]);
    acc = fa(acc, fm(d, d));
  }
  dmu = __fsqrt_rn(acc);
  flags |= GBP_FLAG_MUVALID;
  if (dmu < g.hp.dmu_threshold && dcount > g.hp.min_linear_iters - g.hp.num_undamped_iters) {
    damping = 0.0f;  // gbp_codelets.cpp:280-283
    dcount = -g.hp.num_undamped_iters;
    // the staged record is the current potential: accumulate onto it (quirk Q1), write it back
    const uint32_t robust = relinearise_record(stage + lane, 32, g.fac + e, g.E_pad, stage + lane, g.cam_lin + (size_t)cam * 5,
                                               make_float4(g.K[0], g.K[1], g.K[2], g.K[3]), g.hp.Nstds, rb.x, rb.y, g.var[e],
                                               x_kf[0], x_kf[1], x_kf[2], x_kf[3], x_kf[4], x_kf[5], x_l[0], x_l[1], x_l[2]);
    flags = (flags & ~GBP_FLAG_ROBUST) | (robust ? GBP_FLAG_ROBUST : 0u);
    // relinearisations of this sweep, one atomic per warp that has any (drives the choice between the
    // fused and the two-pass sweep); costs nothing on the common path
    const uint32_t m = __activemask();
    if (lane == (uint32_t)__ffs(m) - 1u) atomicAdd(g.relin_ring + (g.relin_ring[GBP_RELIN_RING] % GBP_RELIN_RING), __popc(m));
  }
}

// One warp-tile: prep + messages of 32 factors from the landed stage.
// lb: landmark record of this lane's factor (load_lmk_belief).
template <bool PREP, bool MSG, bool UPPER>
GBP_DEV void sweep_tile(const DeviceGraph& g, float4* stage, const float* s_cam, const uint32_t wt, const uint2 ti,
                        const float (&lb)[20], const uint32_t lane) {
  const size_t e = (size_t)wt * 32 + lane;
  const bool valid = lane < (ti.y & 0xffu);  // padding slots hold no factor (upper bits of ti.y: warp-tiles of the camera)
  const float4 ra = stage[GBP_SQ_RECA * 32 + lane];
  const float4 rb = stage[GBP_SQ_RECB * 32 + lane];
  float damping = ra.x;
  int dcount = __float_as_int(ra.y);
  uint32_t flags = __float_as_uint(ra.z);
  float dmu = ra.w;
  const bool active = valid && (flags & GBP_FLAG_ACTIVE) != 0;
  const size_t lpos = __float_as_uint(rb.w);  // where this factor's landmark-bound message lives

  if (PREP && active) prep_factor(g, stage, s_cam, ti.x, e, lane, lb, rb, damping, dcount, flags, dmu);

  float nc[28];   // new f->cam message record: eta 0..5 | lower lambda 6..26 | pad
  float ncu[16];  // its strict upper triangle (row-major, i<j): only summed into the camera partial
  if (MSG) {
    if (active) {
      float nl[12];
      msg_to_landmark(stage, s_cam, lane, damping, nl);
      float4* p = g.mlmk + lpos * GBP_MLMK_QUADS;
#pragma unroll
      for (int q = 0; q < 3; ++q) p[q] = make_float4(nl[q * 4], nl[q * 4 + 1], nl[q * 4 + 2], nl[q * 4 + 3]);
      msg_to_camera<UPPER>(stage, lane, lb, damping, nc, ncu);
      store_cam_message(g, e, nc, ncu);
      flags |= GBP_FLAG_HASMSG;
    } else {
      // inactive (or padding) slot: its messages are zero (gbp_codelets.cpp:464-468 etc.)
#pragma unroll
      for (int k = 0; k < 28; ++k) nc[k] = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) ncu[k] = 0.f;
      if (valid && (flags & GBP_FLAG_HASMSG)) {
        store_cam_message(g, e, nc, ncu);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 3; ++q) g.mlmk[lpos * GBP_MLMK_QUADS + q] = z4;
        flags &= ~GBP_FLAG_HASMSG;
      }
    }
  }
  // the state record only changes in this kernel when prep ran here or the has-message flag toggled
  if (valid && (PREP ? (active || MSG) : flags != __float_as_uint(ra.z)))
    g.recA[e] = make_float4(damping, __int_as_float(dcount), __uint_as_float(flags), dmu);

  if (MSG) {
    // lane-ordered reduction through the warp's own (now consumed) stage
    float* red = reinterpret_cast<float*>(stage);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 6; ++i) red[i * GBP_RED_STRIDE + lane] = nc[i];
    if (UPPER) {
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j)
          red[(6 + i * 6 + j) * GBP_RED_STRIDE + lane] =
              (i >= j) ? nc[GBP_MCAM_LOWER + lt(i, j)] : ncu[gbp_upper(i, j)];
    } else {
#pragma unroll
      for (int k = 0; k < 21; ++k) red[(6 + k) * GBP_RED_STRIDE + lane] = nc[GBP_MCAM_LOWER + k];
    }
    __syncwarp();
    warp_cam_reduce<UPPER>(red, lane, g.cam_partial + (size_t)wt * GBP_CAMPART_STRIDE);
    __syncwarp();
  }
}

// UPPER == false: the strict upper triangle of the factor->camera message Lambda is neither computed nor
// summed.  Nothing in the algorithm ever reads it back (inv6x6 and the staged camera belief use the lower
// triangle only); it only reaches the caller through cam_beliefs_lambda.  gbp_cuda_iterate therefore runs
// every sweep of a call without per-sweep metrics but the LAST in this mode (k_update_vars mirrors the lower triangle meanwhile): the
// belief of a camera is the sum of the messages of the latest sweep alone, so after the call every tensor is
// what the all-UPPER sequence would have produced, bit for bit.
template <bool PREP, bool MSG, bool UPPER>
__global__ void __launch_bounds__(GBP_SW_WARPS * 32, 1) k_sweep(const DeviceGraph g) {
  extern __shared__ float4 smem4[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* stage_base = smem4 + warp * (GBP_NBUF * GBP_STAGE_QUADS);
  float* scam_base = reinterpret_cast<float*>(smem4 + GBP_SW_WARPS * GBP_NBUF * GBP_STAGE_QUADS) + warp * (GBP_NBUF * GBP_SCAM);
  const uint32_t n_wt = g.E_pad / 32;
  const uint32_t stride = gridDim.x * GBP_SW_WARPS;
  // consecutive warp-tiles go to different SMs, so small graphs spread over the whole chip
  uint32_t wt = warp * gridDim.x + blockIdx.x;
  if (wt >= n_wt) return;

  // prologue: everything of the first warp-tile, ids of the second
  uint2 ti = __ldg(g.wt_info + wt);
  uint32_t buf = 0;
  // {landmark id, message position} of a lane's factor = the second half of its recB record
  const uint2* lrec = reinterpret_cast<const uint2*>(g.recB) + 1;
  uint2 lid = __ldg(lrec + 2 * ((size_t)wt * 32 + lane));
  issue_stage(g, stage_base, scam_base, wt, ti.x, lid.y, lane);
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  float lb[20];
  load_lmk_belief(g, lid.x, lb);
  uint32_t wt_n = wt + stride;
  uint2 ti_n = make_uint2(0u, 0u);
  uint2 lid_n = make_uint2(0u, 0u);
  if (wt_n < n_wt) {
    ti_n = __ldg(g.wt_info + wt_n);
    lid_n = __ldg(lrec + 2 * ((size_t)wt_n * 32 + lane));
  }
  for (;;) {
    const bool has_next = wt_n < n_wt;
    float lb_n[20];
    uint2 ti_nn = make_uint2(0u, 0u);
    uint2 lid_nn = make_uint2(0u, 0u);
    const uint32_t wt_nn = wt_n + stride;
    if (has_next) {
      issue_stage(g, stage_base + (buf ^ 1) * GBP_STAGE_QUADS, scam_base + (buf ^ 1) * GBP_SCAM, wt_n, ti_n.x, lid_n.y, lane);
      load_lmk_belief(g, lid_n.x, lb_n);
      if (wt_nn < n_wt) {
        ti_nn = __ldg(g.wt_info + wt_nn);
        lid_nn = __ldg(lrec + 2 * ((size_t)wt_nn * 32 + lane));
#if GBP_L2_PREFETCH == 1
        // warp-tile t+2: pull its 23 contiguous 512-byte rows (potential 14, camera message 7, the two
        // edge-state records) into L2 now, one bulk prefetch per lane, so that the cp.async copies issued
        // at the top of the next iteration are served from L2 instead of queueing behind DRAM
        const float4* row = nullptr;
        if (lane < GBP_FAC_QUADS) row = g.fac + (size_t)lane * g.E_pad;
        else if (lane < GBP_FAC_QUADS + GBP_MCAM_QUADS) row = g.mcam + (size_t)(lane - GBP_FAC_QUADS) * g.E_pad;
        else if (lane == GBP_FAC_QUADS + GBP_MCAM_QUADS) row = g.recA;
        else if (lane == GBP_FAC_QUADS + GBP_MCAM_QUADS + 1) row = g.recB;
        if (row) asm volatile("cp.async.bulk.prefetch.L2.global [%0], 512;\n" ::"l"(row + (size_t)wt_nn * 32) : "memory");
#elif GBP_L2_PREFETCH == 2
        // the same 23 rows x 4 lines of 128 B, one prefetch.global.L2 per line, three per lane
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const uint32_t i = lane + 32u * k, r = i >> 2;
          if (r < GBP_FAC_QUADS + GBP_MCAM_QUADS + 2) {
            const float4* row = (r < GBP_FAC_QUADS) ? g.fac + (size_t)r * g.E_pad
                                : (r < GBP_FAC_QUADS + GBP_MCAM_QUADS) ? g.mcam + (size_t)(r - GBP_FAC_QUADS) * g.E_pad
                                : (r == GBP_FAC_QUADS + GBP_MCAM_QUADS) ? g.recA : g.recB;
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(row + (size_t)wt_nn * 32 + (i & 3u) * 8) : "memory");
          }
        }
#endif
      }
    }
    // the copies of THIS warp-tile were committed one iteration ago: leave the newest group in flight
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 1;\n" ::: "memory");
    __syncwarp();
    sweep_tile<PREP, MSG, UPPER>(g, stage_base + buf * GBP_STAGE_QUADS, scam_base + buf * GBP_SCAM, wt, ti, lb, lane);
    if (!has_next) break;
    wt = wt_n; wt_n = wt_nn;
    ti = ti_n; ti_n = ti_nn;
    lid_n = lid_nn;
#pragma unroll
    for (int i = 0; i < 20; ++i) lb[i] = lb_n[i];
    buf ^= 1;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// Warp-tiles beyond the first two (static) rounds come from a device-side queue.  tile_queue = {tickets handed out,
// warps that have finished}; the last warp of the launch to finish rewinds both, so every launch starts from an empty
// queue without a memset node.  A ticket is DRAWN one tile before it is needed (the atomic's round trip is hidden
// behind a tile of arithmetic) and a warp only draws again while its last ticket was valid, so no valid ticket is
// ever dropped.  The makespan of a launch then no longer jumps by a whole round when the graph (or a shard) holds a
// few warp-tiles more than a multiple of the resident warps.
GBP_DEV uint32_t queue_draw(const DeviceGraph& g, const uint32_t lane) {
  uint32_t t = 0;
  if (lane == 0) t = atomicAdd(g.tile_queue, 1u);
  return t;
}
GBP_DEV uint32_t queue_resolve(const uint32_t ticket, const uint32_t n_wt, const uint32_t n_static) {
  const uint32_t t = __shfl_sync(0xffffffffu, ticket, 0) + 2u * n_static;
  return t < n_wt ? t : 0xffffffffu;
}
GBP_DEV void tile_queue_done(const DeviceGraph& g, const uint32_t lane, const uint32_t n_wt, const uint32_t n_static) {
  if (lane != 0) return;
  const uint32_t n_warps = n_static < n_wt ? n_static : n_wt;  // warps of this launch that had a first tile
  __threadfence();
  if (atomicAdd(g.tile_queue + 1, 1u) == n_warps - 1u) {
    g.tile_queue[0] = 0u;
    g.tile_queue[1] = 0u;
  }
}

// ---- PrepMessageVertex as its own pass, relinearisation by compaction ---------------------------
// k_prep_pass runs the damping state machine of every factor (one thread per edge slot; it only
// needs the 16-byte state record and the means of the two variables) and appends the slots whose
// factor relinearises this sweep to a list; k_relin_list relinearises exactly those, one thread
// each.  Whether 3 % of the factors relinearise per sweep (a de-synchronised graph) or all of them
// in one sweep out of eleven, the heavy path runs in fully populated warps and stays out of the
// instruction stream of the message kernel.
__global__ void __launch_bounds__(GBP_TILE) k_prep_pass(const DeviceGraph g) {
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  const size_t e = (size_t)blockIdx.x * GBP_TILE + tid;
  const float4 ra = g.recA[e];
  float damping = ra.x;
  int dcount = __float_as_int(ra.y);
  uint32_t flags = __float_as_uint(ra.z);
  float dmu = ra.w;
  const bool active = !(flags & GBP_FLAG_PAD) && (flags & GBP_FLAG_ACTIVE);
  bool need = false;
  if (active) {
    const uint32_t cam = g.wt_info[e >> 5].x;
    const uint32_t l = __float_as_uint(g.recB[e].z);
    const float* crec = reinterpret_cast<const float*>(g.cam_rec + (size_t)cam * 16);  // ... | mean 42..47 | previous mean 48..53
    const float4 lm = g.lmk_b[(size_t)l * GBP_LMKB_QUADS + 3];
    if (dcount == 0) damping = g.hp.maxeta_damping;  // gbp_codelets.cpp:245-248
    dcount += 1;
    float old[9];
    if (flags & GBP_FLAG_MUVALID) {
      const float4 lp = g.lmk_mean_prev[l];
#pragma unroll
      for (int i = 0; i < 6; ++i) old[i] = crec[48 + i];
      old[6] = lp.x; old[7] = lp.y; old[8] = lp.z;
    } else {
#pragma unroll
      for (int i = 0; i < 9; ++i) old[i] = g.oldmu_edge ? g.oldmu_edge[(size_t)i * g.E_pad + e] : 0.f;
    }
    const float x_l[3] = {lm.x, lm.y, lm.z};
    float acc = 0.f;  // gbp_codelets.cpp:268-277
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float d = fs(old[i], crec[42 + i]);
      acc = fa(acc, fm(d, d));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float d = fs(old[6 + i], x_l[i]);
      acc = fa(acc, fm(d, d));
    }
    dmu = __fsqrt_rn(acc);
    flags |= GBP_FLAG_MUVALID;
    if (dmu < g.hp.dmu_threshold && dcount > g.hp.min_linear_iters - g.hp.num_undamped_iters) {
      damping = 0.0f;  // gbp_codelets.cpp:280-283
      dcount = -g.hp.num_undamped_iters;
      need = true;
    }
    g.recA[e] = make_float4(damping, __int_as_float(dcount), __uint_as_float(flags), dmu);
  }
  // warp-aggregated append
  const uint32_t mask = __ballot_sync(0xffffffffu, need);
  if (mask) {
    uint32_t base = 0;
    if (lane == 0) {
      base = atomicAdd(g.relin_count, __popc(mask));
      atomicAdd(g.relin_ring + (g.relin_ring[GBP_RELIN_RING] % GBP_RELIN_RING), __popc(mask));
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (need) g.relin_list[base + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)e;
  }
}

__global__ void __launch_bounds__(GBP_TILE) k_relin_list(const DeviceGraph g) {
  const uint32_t i = blockIdx.x * GBP_TILE + threadIdx.x;
  if (i >= *g.relin_count) return;
  const size_t e = g.relin_list[i];
  const float4 rb = g.recB[e];
  const uint32_t cam = g.wt_info[e >> 5].x;
  const float* crec = reinterpret_cast<const float*>(g.cam_rec + (size_t)cam * 16);
  const float4 lm = g.lmk_b[(size_t)__float_as_uint(rb.z) * GBP_LMKB_QUADS + 3];
  // accumulate onto the current potential (quirk Q1), Huber re-evaluated (quirk Q2)
  const uint32_t robust = relinearise_record(g.fac + e, g.E_pad, g.fac + e, g.E_pad, nullptr, g.cam_lin + (size_t)cam * 5,
                                             make_float4(g.K[0], g.K[1], g.K[2], g.K[3]), g.hp.Nstds, rb.x, rb.y, g.var[e], crec[42],
                                             crec[43], crec[44], crec[45], crec[46], crec[47], lm.x, lm.y, lm.z);
  uint32_t* fl = reinterpret_cast<uint32_t*>(g.recA + e) + 2;
  *fl = (*fl & ~GBP_FLAG_ROBUST) | (robust ? GBP_FLAG_ROBUST : 0u);
}

// Recompute the per-warp camera partial sums from the stored messages (used
// after set_tensor on the camera message tensors; the strict upper triangle is
// the mirrored lower one, see gbp_layout.h).
__global__ void __launch_bounds__(GBP_TILE) k_cam_partials(const DeviceGraph g) {
  __shared__ __align__(16) float s_red[GBP_TILE / 32][GBP_CAMPART * GBP_RED_STRIDE];
  const uint32_t tile = blockIdx.x, tid = threadIdx.x;
  const uint32_t warp = tid >> 5, lane = tid & 31;
  const size_t e = (size_t)tile * GBP_TILE + tid;
  float m[GBP_MCAM_QUADS * 4];
  load_quads<GBP_MCAM_QUADS>(g.mcam, g.E_pad, e, m);
  float* red = s_red[warp];
#pragma unroll
  for (int i = 0; i < 6; ++i) red[i * GBP_RED_STRIDE + lane] = m[i];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) red[(6 + i * 6 + j) * GBP_RED_STRIDE + lane] = m[GBP_MCAM_LOWER + gbp_sym(i, j)];
  if (g.mcam_up) {
    float u[16];
    load_quads<4>(g.mcam_up, g.E_pad, e, u);
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = i + 1; j < 6; ++j) red[(6 + i * 6 + j) * GBP_RED_STRIDE + lane] = u[gbp_upper(i, j)];
  }
  __syncwarp();
  warp_cam_reduce<true>(red, lane, g.cam_partial + ((size_t)tile * (GBP_TILE / 32) + warp) * GBP_CAMPART_STRIDE);
}

// b += the factor->landmark messages of landmark l, strictly in slot order (= original
// edge order, the reference's message slots 1..deg); they are contiguous in mlmk.
GBP_DEV void lmk_accumulate(const DeviceGraph& g, const uint32_t l, float (&b)[12]) {
  const uint32_t k0 = g.lmk_ptr[l], k1 = g.lmk_ptr[l + 1];
  for (uint32_t k = k0; k < k1; ++k) {
    const float4* p = g.mlmk + (size_t)k * GBP_MLMK_QUADS;
    const float4 v0 = p[0], v1 = p[1], v2 = p[2];
    b[0] = fa(b[0], v0.x); b[1] = fa(b[1], v0.y); b[2] = fa(b[2], v0.z); b[3] = fa(b[3], v0.w);
    b[4] = fa(b[4], v1.x); b[5] = fa(b[5], v1.y); b[6] = fa(b[6], v1.z); b[7] = fa(b[7], v1.w);
    b[8] = fa(b[8], v2.x); b[9] = fa(b[9], v2.y); b[10] = fa(b[10], v2.z); b[11] = fa(b[11], v2.w);
  }
}

// the landmark's three terms of dmu (gbp_codelets.cpp:268-277): (old - new)^2 per mean component
GBP_DEV float4 lmk_sq_terms(const float4 prev, const float (&mean)[3]) {
  const float d0 = fs(prev.x, mean[0]), d1 = fs(prev.y, mean[1]), d2 = fs(prev.z, mean[2]);
  return make_float4(fm(d0, d0), fm(d1, d1), fm(d2, d2), 0.f);
}

// belief record of one landmark: [eta 3 | Lambda 9 | mean 3 | pad]; shift != 0 keeps the
// mean the last PrepMessageVertex pass used as the "old mu" (Copy(mu, oldmu), ba/ba.cpp:898)
GBP_DEV void lmk_store_belief(const DeviceGraph& g, const uint32_t l, const float (&b)[12], const int shift) {
  const float eta[3] = {b[0], b[1], b[2]};
  float lam[9], mean[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) lam[i] = b[3 + i];
  inf2mean3(eta, lam, mean);
  float4* o = g.lmk_b + (size_t)l * GBP_LMKB_QUADS;
  float4 prev;
  if (shift) {
    const float4 oldq = o[3];
    prev = make_float4(oldq.x, oldq.y, oldq.z, 0.f);
    g.lmk_mean_prev[l] = prev;
  } else {
    prev = g.lmk_mean_prev[l];
  }
  g.lmk_sq[l] = lmk_sq_terms(prev, mean);
  o[0] = make_float4(b[0], b[1], b[2], b[3]);
  o[1] = make_float4(b[4], b[5], b[6], b[7]);
  o[2] = make_float4(b[8], b[9], b[10], b[11]);
  o[3] = make_float4(mean[0], mean[1], mean[2], 0.f);
}

GBP_DEV void lmk_load_prior(const DeviceGraph& g, const uint32_t l, float (&b)[12]) {
  // the sum starts from +0 like a zero-initialised accumulator: 0 + prior (turns a -0 prior into +0)
  const float4* p = g.lmk_prior + (size_t)l * 3;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float4 v = p[q];
    b[q * 4] = fa(0.0f, v.x); b[q * 4 + 1] = fa(0.0f, v.y); b[q * 4 + 2] = fa(0.0f, v.z); b[q * 4 + 3] = fa(0.0f, v.w);
  }
}

#define GBP_LMK_PER_BLOCK (GBP_TILE / 4)  // four lanes per landmark
// Belief update + per-variable mean (inf2mean hoisted out of PrepMessageVertex).  shift != 0: the mean that the
// last PrepMessageVertex pass used becomes the "old mu" (Copy(mu, oldmu),
// ba/ba.cpp:898) before the new mean is stored.
// Belief update of the cameras (prog_ub, ba/ba.cpp:104-139, camera half) + per-camera mean and rotation.
// One WARP per camera, four cameras per block.  The per-warp-tile partial sums k_sweep left for the camera are one
// contiguous run of cam_partial: lane 0 stages it with ONE bulk copy (completion on the warp's mbarrier) while the
// lanes fetch the prior entries and the old mean, so a camera is two memory round trips deep (its tile range, then
// everything else) before the arithmetic starts: the lanes add the partials to the prior in warp-tile order -- entry
// `lane` and entry `lane + 32` of [eta 6 | Lambda 36] -- and lane 0 inverts the 6x6 belief (a serial LDL^T) and
// forms the linearisation constants.
#define GBP_CAM_PER_BLOCK 4
#define GBP_CAM_STAGE_TILES 30  // warp-tiles of one camera a warp stages (30 x 176 B); longer runs finish through direct loads
GBP_DEV void update_cameras(const DeviceGraph& g, float* s_buf, uint64_t* s_bars, const int shift, const uint32_t block, const int lower_only) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t c = block * GBP_CAM_PER_BLOCK + warp;
  if (warp >= GBP_CAM_PER_BLOCK || c >= g.C) return;  // whole warp
  float* s_part = s_buf + warp * (GBP_CAM_STAGE_TILES * GBP_CAMPART_STRIDE);
  uint64_t* bar = s_bars + warp;
  const uint32_t t0 = g.cam_wt_begin[c], t1 = g.cam_wt_begin[c + 1];
  const uint32_t n_st = (t1 - t0 < GBP_CAM_STAGE_TILES) ? t1 - t0 : GBP_CAM_STAGE_TILES;
  if (lane == 0 && n_st) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
    const uint32_t bytes = n_st * (uint32_t)(GBP_CAMPART_STRIDE * sizeof(float));
    mbar_expect_tx(bar, bytes);
    bulk_load(s_part, g.cam_partial + (size_t)t0 * GBP_CAMPART_STRIDE, bytes, bar);
  }
  // entries `lane` and `lane + 32` of [eta 6 | Lambda 36 row-major]; after a sweep that skipped the strict upper triangle of
  // the camera messages (k_sweep<.., UPPER = false>) the upper entries of the belief mirror the lower ones
  float acc[2] = {0.f, 0.f};
  bool skip[2] = {true, true};
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t ent = lane + 32u * half;
    if (ent < GBP_CAMPART) {
      const uint32_t bi = ent >= 6 ? (ent - 6) / 6 : 0u, bj = ent >= 6 ? (ent - 6) % 6 : 0u;
      skip[half] = lower_only && ent >= 6 && bi < bj;
      // the sum starts from +0 like a zero-initialised accumulator: 0 + prior (turns a -0 prior into +0)
      if (!skip[half]) acc[half] = fa(0.0f, (ent < 6) ? g.cam_prior_eta[c * 6 + ent] : g.cam_prior_lam[c * 36 + (ent - 6)]);
    }
  }
  float prev6 = 0.f;  // lanes 0..5: the mean the last PrepMessageVertex pass used (Copy(mu, oldmu), ba/ba.cpp:898)
  if (lane < 6) prev6 = shift ? g.cam_mean[c * 6 + lane] : g.cam_mean_prev[c * 6 + lane];
  __syncwarp();
  if (n_st) mbar_wait(bar, 0);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t ent = lane + 32u * half;
    if (ent < GBP_CAMPART && !skip[half]) {
      float a = acc[half];
      for (uint32_t t = 0; t < n_st; ++t) a = fa(a, s_part[t * GBP_CAMPART_STRIDE + ent]);  // warp-tile order
      for (uint32_t t = t0 + n_st; t < t1; t += 8) {  // the rest of a long run: eight direct loads in flight
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (t + u < t1) ? g.cam_partial[(size_t)(t + u) * GBP_CAMPART_STRIDE + ent] : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (t + u < t1) a = fa(a, v[u]);
      }
      acc[half] = a;
    }
  }
  // the 42 sums go through the (consumed) staging buffer: lane 0 needs all of them for the inverse, the mirrored entries
  // come from their transposes
  __syncwarp();
  float* s_b = s_part;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t ent = lane + 32u * half;
    if (ent < GBP_CAMPART && !skip[half]) s_b[ent] = acc[half];
  }
  __syncwarp();
  float* rec = reinterpret_cast<float*>(g.cam_rec + (size_t)c * 16);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t ent = lane + 32u * half;
    if (ent < GBP_CAMPART) {
      const uint32_t bi = ent >= 6 ? (ent - 6) / 6 : 0u, bj = ent >= 6 ? (ent - 6) % 6 : 0u;
      const float v = skip[half] ? s_b[6 + bj * 6 + bi] : s_b[ent];
      if (ent < 6) g.cam_b_eta[c * 6 + ent] = v;
      else g.cam_b_lam[c * 36 + (ent - 6)] = v;
      rec[ent] = v;
    }
  }
  float prev[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) prev[i] = __shfl_sync(0xffffffffu, prev6, i);
  if (lane == 0) {
    float eta[6], lamL[21], mean[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) eta[i] = s_b[i];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) lamL[lt(i, j)] = s_b[6 + i * 6 + j];
    inf2mean6(eta, lamL, mean);
    const float w[3] = {mean[3], mean[4], mean[5]};
    float R[9], num[9], den;
    cam_lin_consts(w, R, num, den);
    float acc6 = 0.f;  // the camera's six terms of dmu (gbp_codelets.cpp:268-277), hoisted out of the factors
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      g.cam_mean_prev[c * 6 + i] = prev[i];
      g.cam_mean[c * 6 + i] = mean[i];
      rec[42 + i] = mean[i];
      rec[48 + i] = prev[i];
      const float d = fs(prev[i], mean[i]);
      acc6 = fa(acc6, fm(d, d));
    }
    rec[54] = acc6;
    float* lin = reinterpret_cast<float*>(g.cam_lin + (size_t)c * 5);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      lin[i] = R[i];
      lin[9 + i] = num[i];
    }
    lin[18] = den;
  }
}

// Belief update of the landmarks (prog_ub, landmark half) + per-landmark mean:
// GBP_LMK_PER_BLOCK landmarks per block.
// Four lanes per landmark: lane q < 3 owns quad q of [eta 3 | Lambda 9].
// acc += quad q of the landmark's factor->landmark messages, strictly in slot order (= original
// edge order, the reference's message slots 1..deg; contiguous in mlmk), four independent
// 16-byte loads in flight per lane.
#ifndef GBP_LMK_BATCH
#define GBP_LMK_BATCH 4
#endif
GBP_DEV float4 lmk_sum_quad(const DeviceGraph& g, const uint32_t l, const uint32_t q, float4 acc) {
  const uint32_t k0 = g.lmk_ptr[l], k1 = g.lmk_ptr[l + 1];
  for (uint32_t k = k0; k < k1; k += GBP_LMK_BATCH) {
    float4 v[GBP_LMK_BATCH];
#pragma unroll
    for (int u = 0; u < GBP_LMK_BATCH; ++u)
      if (k + u < k1) v[u] = g.mlmk[(size_t)(k + u) * GBP_MLMK_QUADS + q];
#pragma unroll
    for (int u = 0; u < GBP_LMK_BATCH; ++u)
      if (k + u < k1) {
        acc.x = fa(acc.x, v[u].x); acc.y = fa(acc.y, v[u].y); acc.z = fa(acc.z, v[u].z); acc.w = fa(acc.w, v[u].w);
      }
  }
  return acc;
}
GBP_DEV float4 lmk_prior_quad(const DeviceGraph& g, const uint32_t l, const uint32_t q) {
  // the sum starts from +0 like a zero-initialised accumulator: 0 + prior (turns a -0 prior into +0)
  const float4 pr = g.lmk_prior[(size_t)l * 3 + q];
  return make_float4(fa(0.0f, pr.x), fa(0.0f, pr.y), fa(0.0f, pr.z), fa(0.0f, pr.w));
}
// The landmark's first lane collects the 12 sums, forms the mean and stores the belief record
// [eta 3 | Lambda 9 | mean 3 | pad].  Must be reached by all 32 lanes of the warp.
// prev_in (have_prev): the landmark's previous mean fetched by the caller ahead of time -- the old mean (shift) or
// lmk_mean_prev -- so that the load does not sit at the end of the block's dependency chain
GBP_DEV void lmk_finish_quads(const DeviceGraph& g, const uint32_t l, const uint32_t q, const float4 acc, const bool mine,
                              const int shift, const bool have_prev = false, const float4 prev_in = make_float4(0.f, 0.f, 0.f, 0.f)) {
  float b[12];
  const uint32_t base = (threadIdx.x & 31) & ~3u;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    b[j * 4 + 0] = __shfl_sync(0xffffffffu, acc.x, base + j);
    b[j * 4 + 1] = __shfl_sync(0xffffffffu, acc.y, base + j);
    b[j * 4 + 2] = __shfl_sync(0xffffffffu, acc.z, base + j);
    b[j * 4 + 3] = __shfl_sync(0xffffffffu, acc.w, base + j);
  }
  if (!mine) return;
  float4* o = g.lmk_b + (size_t)l * GBP_LMKB_QUADS;
  if (q == 0) {
    const float eta[3] = {b[0], b[1], b[2]};
    float lam[9], mean[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) lam[i] = b[3 + i];
    inf2mean3(eta, lam, mean);
    float4 prev;
    if (shift) {  // Copy(mu, oldmu), ba/ba.cpp:898
      const float4 oldq = have_prev ? prev_in : o[3];
      prev = make_float4(oldq.x, oldq.y, oldq.z, 0.f);
      g.lmk_mean_prev[l] = prev;
    } else {
      prev = have_prev ? prev_in : g.lmk_mean_prev[l];
    }
    g.lmk_sq[l] = lmk_sq_terms(prev, mean);
    o[3] = make_float4(mean[0], mean[1], mean[2], 0.f);
  }
  if (q < 3) o[q] = acc;
}

// A block updates the landmarks [l0, l1) of its record lmk_blk[b] = {l0, l1, k0, k1} -- at most 32, four lanes each.
// Their messages are ONE contiguous run [k0, k1) of mlmk (landmark order), so a single cp.async.bulk brings the whole
// run into shared memory (one elected thread, completion on an mbarrier) instead of every lane chasing its landmark's
// messages through dependent 16-byte loads; the sums then read shared memory, strictly in slot order as before.
// Everything a lane needs besides (its message range, its prior quad, the landmark's previous mean) is requested in
// the same breath, so a block is two memory round trips deep: the record, then all of the rest.  A landmark with
// more than GBP_LMK_CAP messages has a block of its own and takes the direct path.
#define GBP_LMK_CAP 448  // messages a block stages (21 KB)
GBP_DEV void update_landmarks(const DeviceGraph& g, float4* s_msg, uint64_t* s_bars, const int shift, const uint32_t block) {
  uint64_t& s_bar = s_bars[0];
  const uint4 rec = __ldg(g.lmk_blk + block);
  const uint32_t l0 = rec.x, l1 = min(l0 + GBP_LMK_PER_BLOCK, g.L), k0 = rec.z, k1 = rec.w;
  const bool staged = k1 - k0 <= GBP_LMK_CAP && k1 > k0;
  if (staged && threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
    const uint32_t bytes = (k1 - k0) * (uint32_t)(GBP_MLMK_QUADS * sizeof(float4));
    mbar_expect_tx(&s_bar, bytes);
    bulk_load(s_msg, g.mlmk + (size_t)k0 * GBP_MLMK_QUADS, bytes, &s_bar);
  }
  const uint32_t l = l0 + (threadIdx.x >> 2), q = threadIdx.x & 3;
  const bool here = l < l1;
  // boundary landmarks of a multi-GPU shard are finished by the exchange blocks / kernels: one bit per landmark in the
  // block's record (a per-landmark word fetched beside the other inputs cost the launch 1.4 us, measured)
  static_assert(GBP_LMK_PER_BLOCK == 32 && GBP_TILE == 4 * GBP_LMK_PER_BLOCK, "one mask bit per landmark of a block");
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), prev = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t a0 = 0, a1 = 0;
  if (here && q < 3) {
    acc = lmk_prior_quad(g, l, q);
    a0 = g.lmk_ptr[l];
    a1 = g.lmk_ptr[l + 1];
  }
  if (here && q == 0) prev = shift ? g.lmk_b[(size_t)l * GBP_LMKB_QUADS + 3] : g.lmk_mean_prev[l];
  const bool mine = here && !((rec.y >> (threadIdx.x >> 2)) & 1u);
  if (staged) {
    __syncthreads();  // the barrier is initialised before anyone waits on it
    mbar_wait(&s_bar, 0);
    if (mine && q < 3) {
      const float4* m = s_msg + (size_t)(a0 - k0) * GBP_MLMK_QUADS + q;
      for (uint32_t k = a0; k < a1; ++k, m += GBP_MLMK_QUADS) {
        const float4 v = *m;
        acc.x = fa(acc.x, v.x); acc.y = fa(acc.y, v.y); acc.z = fa(acc.z, v.z); acc.w = fa(acc.w, v.w);
      }
    }
  } else if (mine && q < 3) {
    acc = lmk_sum_quad(g, l, q, acc);
  }
  lmk_finish_quads(g, l, q, acc, mine, shift, true, prev);
}

// ---- multi-GPU: partial sums fused with their exchange over peer memory ----------------
// Landmarks that other ranks observe too ("boundary") get
//     belief = (0 + prior) + partial[rank 0] + partial[rank 1] + ...
// where partial[r] is rank r's sum of its own messages (slot order, from +0): the same
// operations on every rank, so all replicas stay bit-identical.
// boundary_push: the FIRST blocks of k_update_vars form this rank's partial sums and store them, every float tagged
// with the exchange step, into this rank's exchange buffer (default) -- or straight into the buffer of every rank
// that observes the landmark (GBP_XCHG_PUSH=1) -- through NVLink-mapped peer pointers (CUDA IPC between processes,
// direct pointers inside one process).
// boundary_finish: later blocks of the same grid poll the tagged words of every observing rank -- over NVLink in that
// rank's buffer (default), or in their own buffer (push) -- until they carry the current step, add them in rank
// order and finish the boundary landmarks.  No collective call, no communication stream, no host synchronisation,
// no flag and no fence: the transfer overlaps the cameras and the interior landmarks of the same launch.
// Buffers are double-buffered by step parity: a peer may already be at step s+1 while this rank still reads step s,
// never at s+2 (it needs this rank's step-s+1 partials first).
#ifdef GBP_DEBUG_TS
GBP_DEV unsigned long long dbg_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define GBP_TS_MAX(g, step, i) do { if ((g).dbg_ts && threadIdx.x == 0) atomicMax((g).dbg_ts + ((step) % 32u) * 8u + (i), dbg_now()); } while (0)
#define GBP_TS_MIN(g, step, i) do { if ((g).dbg_ts && threadIdx.x == 0) atomicMin((g).dbg_ts + ((step) % 32u) * 8u + (i), dbg_now()); } while (0)
#else
#define GBP_TS_MAX(g, step, i) do { } while (0)
#define GBP_TS_MIN(g, step, i) do { } while (0)
#endif

GBP_DEV void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
GBP_DEV uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Flag-in-data exchange (the scheme of NCCL's LL protocol): every float of a partial sum travels as an 8-byte word
// {value bits, exchange step}; an aligned 8-byte store is indivisible on the way to a peer, so the receiver needs no
// fence, no flag and no counter -- it polls the very words it is going to add until their tag is the step it is in.
// The receive buffers start out zeroed and steps count from 1, so a stale word can never pass for a fresh one; two
// parities because a peer may already push step s+1 while this rank still reads step s (never s+2: it needs this
// rank's step-s+1 partials first).
GBP_DEV void st_pair_sys(uint4* p, const float a, const float b, const uint32_t tag) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "r"(__float_as_uint(a)), "r"(tag), "r"(__float_as_uint(b)), "r"(tag) : "memory");
}
GBP_DEV uint4 ld_pair_sys(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

GBP_DEV void boundary_push(const DeviceGraph& g, const uint32_t step, const uint32_t block, const bool pull) {
  const uint32_t k = block * GBP_LMK_PER_BLOCK + (threadIdx.x >> 2), q = threadIdx.x & 3;
  if (k < g.n_bnd_local && q < 3) {
    // one record per boundary landmark (k_boundary_records): the partial sum is two memory round trips deep --
    // the record, then up to eight of its messages at a time, added strictly in slot order from +0
    const uint4 rec = __ldg(g.bnd_rec + k);
    const uint32_t span = __ldg(g.bnd_span + k);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t m = rec.y; m < rec.z; m += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (m + u < rec.z) v[u] = g.mlmk[(size_t)(m + u) * GBP_MLMK_QUADS + q];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (m + u < rec.z) {
          acc.x = fa(acc.x, v[u].x); acc.y = fa(acc.y, v[u].y); acc.z = fa(acc.z, v[u].z); acc.w = fa(acc.w, v[u].w);
        }
    }
    // quad q of the landmark's slot in THIS rank's lane of the receive buffer = two 16-byte stores of tagged pairs,
    // to the ranks that observe this landmark (no other rank ever reads it)
    const size_t off = (((size_t)((step & 1u) * g.world + g.rank) * g.n_bnd_global + rec.w) * 3 + q) * 2;
    // pull (default): into this rank's own buffer only -- the peers' finish blocks read it from there over NVLink.  A
    // kernel that has STORED into peer memory pays ~3.8 us at its end (measured, N=2: the system-scope flush of the
    // remote writes), one that has only LOADED from a peer does not.  push (GBP_XCHG_PUSH=1, kept for A/B): into the
    // buffer of every rank that observes the landmark.
    for (uint32_t m = pull ? (span & (1u << g.rank)) : span; m; m &= m - 1) {
      uint4* dst = g.peer_recv[__ffs(m) - 1] + off;
      st_pair_sys(dst, acc.x, acc.y, step);
      st_pair_sys(dst + 1, acc.z, acc.w, step);
    }
  }
  GBP_TS_MAX(g, step, 1);  // last push block has stored its partials
}

GBP_DEV void boundary_finish(const DeviceGraph& g, const int shift, const uint32_t step, const uint32_t block, const bool pull) {
  // everything that does not depend on the peers is fetched before the first poll: the record, the prior, the previous mean
  const uint32_t k = block * GBP_LMK_PER_BLOCK + (threadIdx.x >> 2), q = threadIdx.x & 3;
  const bool mine = k < g.n_bnd_local;
  uint4 rec = make_uint4(0u, 0u, 0u, 0u);
  uint32_t span = 0;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), prev = make_float4(0.f, 0.f, 0.f, 0.f);
  if (mine) {
    rec = __ldg(g.bnd_rec + k);
    span = __ldg(g.bnd_span + k);
    if (q < 3) acc = lmk_prior_quad(g, rec.x, q);
    if (q == 0) prev = shift ? g.lmk_b[(size_t)rec.x * GBP_LMKB_QUADS + 3] : g.lmk_mean_prev[rec.x];
  }
  bool ok = true;
  if (mine && q < 3) {
    // rank order over the ranks that observe the landmark.  The others contribute +0 to the sum over ALL ranks that
    // defines the belief, and acc + (+0) == acc bit for bit (acc starts as 0 + prior, so it is never -0): skipped.
    const size_t base = ((size_t)(step & 1u) * g.world * g.n_bnd_global + rec.w) * 6 + (size_t)q * 2;
    const long long t0 = clock64();
    for (uint32_t m = span; m && ok; m &= m - 1) {  // ascending rank order
      const uint32_t r = __ffs(m) - 1;
      // rank r's lane: in that rank's own buffer (pull; this rank's own lane is local) or in this rank's (push)
      const uint4* src = (pull && r != g.rank ? (const uint4*)g.peer_recv[r] : g.p2p_recv) + base + (size_t)r * g.n_bnd_global * 6;
      uint4 a = ld_pair_sys(src), b = ld_pair_sys(src + 1);
      while (a.y != step || a.w != step || b.y != step || b.w != step) {  // that rank's partial has not landed yet
        if (clock64() - t0 > g.p2p_timeout) {  // a peer died or never made the matching call: do not hang the GPU
          *g.p2p_error = 1u;                   // (host-mapped, sticky: every later call on the handle fails)
          ok = false;
          break;
        }
        __nanosleep(20);
        a = ld_pair_sys(src);
        b = ld_pair_sys(src + 1);
      }
      acc.x = fa(acc.x, __uint_as_float(a.x)); acc.y = fa(acc.y, __uint_as_float(a.z));
      acc.z = fa(acc.z, __uint_as_float(b.x)); acc.w = fa(acc.w, __uint_as_float(b.z));
    }
  }
  GBP_TS_MIN(g, step, 2);  // first / last finish block that has all its partials
  GBP_TS_MAX(g, step, 3);
  // nothing is stored from a receive buffer that is not complete (the four lanes of a landmark decide together)
  ok = __all_sync(0xffffffffu, ok);
  lmk_finish_quads(g, rec.x, q, acc, mine && ok, shift, true, prev);
  GBP_TS_MAX(g, step, 4);  // last finish block done
}

// {local landmark, first message, one past the last message, position in the global boundary list} of every boundary
// landmark this rank touches (built once at init from the device-side lmk_ptr), and its bit in the landmark-block record
__global__ void k_boundary_records(const DeviceGraph g) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= g.n_bnd_local) return;
  const uint32_t l = g.bnd_local[k];
  g.bnd_rec[k] = make_uint4(l, g.lmk_ptr[l], g.lmk_ptr[l + 1], g.bnd_slot[k]);
  atomicOr(&g.lmk_blk[l / GBP_LMK_PER_BLOCK].y, 1u << (l % GBP_LMK_PER_BLOCK));   // the landmark blocks skip it
}

// prog_ub in one launch.  Block roles, in dispatch order:
//   [multi-GPU, peer-to-peer] boundary_push blocks  -- first, so the partials are there when the peers look for them
//   camera blocks     -- few, long-running (a serial 6x6 inverse + Rodrigues per camera, one warp each), nearly idle:
//                        they overlap with the bandwidth-bound landmark blocks that fill the chip
//   [multi-GPU, peer-to-peer] boundary_finish blocks -- by default BEFORE the landmark blocks (finish_after = 0): they
//                        poll for the peers' partials while the landmark blocks stream, instead of forming a serial
//                        tail after them (measured at N=2: behind 80-100 % of the landmark blocks +0.5..2 us per sweep)
//   landmark blocks   -- GBP_LMK_PER_BLOCK landmarks each, their messages staged by one bulk copy
// The register budget (GBP_UV_BLOCKS blocks per SM) fits all paths.  n_push == 0: no fused exchange.
#ifndef GBP_FINISH_AT_DEFAULT
#define GBP_FINISH_AT_DEFAULT 0   // percent of the landmark blocks dispatched before the finish blocks of the exchange
#endif
#ifndef GBP_UV_BLOCKS
#define GBP_UV_BLOCKS 9
#endif
__global__ void __launch_bounds__(GBP_TILE, GBP_UV_BLOCKS) k_update_vars(const DeviceGraph g, const int shift, const uint32_t n_push,
                                                                         const int lower_only, const uint32_t finish_after) {
  // one staging buffer per block, used by whichever role the block plays (landmark message run / camera partial runs)
  __shared__ __align__(128) float4 s_stage[GBP_LMK_CAP * GBP_MLMK_QUADS];
  __shared__ __align__(8) uint64_t s_bars[GBP_CAM_PER_BLOCK];
  static_assert(GBP_CAM_PER_BLOCK * GBP_CAM_STAGE_TILES * GBP_CAMPART_STRIDE * 4 <= GBP_LMK_CAP * GBP_MLMK_QUADS * 16, "camera staging fits");
  const uint32_t nb_lmk = g.n_lmk_blocks;
  const uint32_t nb_cam = (g.C + GBP_CAM_PER_BLOCK - 1) / GBP_CAM_PER_BLOCK;
  if (shift && blockIdx.x == 0 && threadIdx.x == 0) {  // a sweep ended: open the next slot of the relinearisation ring
    const uint32_t next = g.relin_ring[GBP_RELIN_RING] + 1;
    g.relin_ring[next % GBP_RELIN_RING] = 0;
    g.relin_ring[GBP_RELIN_RING] = next;
  }
  // Block roles in dispatch order: [push][cameras][landmarks 0 .. finish_after)[finish][landmarks finish_after ..).
  // The finish blocks are the only ones that wait for a peer; placed behind the first `finish_after` landmark blocks
  // they start a few microseconds into the launch, so a peer that arrives that much later costs this rank nothing,
  // and the rest of the landmark blocks run beside them (no tail).
  const uint32_t fin0 = n_push + nb_cam + min(finish_after, nb_lmk);   // first finish block
  // The exchange step lives on the device (the same sequence on every rank), so the launch has no per-sweep argument
  // and can be replayed from a CUDA graph.  Only the exchange blocks (push / finish) need it: thread 0 of each reads it
  // and THEN takes a ticket; the block that draws the last of the 2 n_push tickets knows every other one has read the
  // step and advances it for the next launch.  The ticket's round trip overlaps the block's work; camera and landmark
  // blocks do not take part at all.
  __shared__ uint32_t s_step;
  uint32_t ticket = 0;
  const bool xchg_block = n_push && (blockIdx.x < n_push || (blockIdx.x >= fin0 && blockIdx.x < fin0 + n_push));
  if (xchg_block) {
    if (threadIdx.x == 0) {
      s_step = *(volatile uint32_t*)g.p2p_step + 1u;
      ticket = atomicAdd(g.p2p_step + 1, 1u);
    }
    __syncthreads();
  }
#ifdef GBP_DEBUG_TS   // the timeline build stamps every block: all of them read the step (diagnostics only)
  const uint32_t step = xchg_block ? s_step : (n_push ? *(volatile uint32_t*)g.p2p_step + 1u : 0u);
#else
  const uint32_t step = xchg_block ? s_step : 0u;
#endif
  uint32_t b = blockIdx.x;
  GBP_TS_MIN(g, step, 0);  // first block of the launch
  if (b < n_push) {
    if (!(lower_only & 16)) boundary_push(g, step, b, lower_only & 64);   // bit 6: pull mode   // bits 4, 5: timing diagnostics (GBP_XCHG_DEBUG)
  } else if ((b -= n_push) < nb_cam) {
    if (!(lower_only & 2)) update_cameras(g, reinterpret_cast<float*>(s_stage), s_bars, shift, b, lower_only & 1);   // bits 1, 2: timing diagnostics (GBP_UV_DEBUG)
    GBP_TS_MAX(g, step, 6);  // last camera block done
  } else if (xchg_block) {
    if (!(lower_only & 32)) boundary_finish(g, shift, step, blockIdx.x - fin0, lower_only & 64);
  } else {
    b -= nb_cam;
    if (blockIdx.x >= fin0) b -= n_push;   // landmark blocks behind the finish blocks
    if (b < nb_lmk) {
      if (!(lower_only & 4)) update_landmarks(g, s_stage, s_bars, shift, b);
      GBP_TS_MAX(g, step, 5);  // last landmark block done
    }
  }
  if (xchg_block && threadIdx.x == 0 && ticket == 2 * n_push - 1) {
    g.p2p_step[1] = 0u;
    g.p2p_step[0] = step;
  }
}

// ---- multi-GPU boundary landmarks, NCCL all-gather path (fallback when CUDA IPC is unavailable) ----
// k_boundary_partial: this rank's partial sum (local factor->landmark messages in slot
// order, starting from +0) of every boundary landmark it touches, written to its slot of
// the exchange buffer.  k_boundary_finish, after the all-gather: belief = (0 + prior) +
// partial[rank 0] + partial[rank 1] + ... -- the same operations on every rank, so all
// replicas of a boundary landmark stay bit-identical.
__global__ void __launch_bounds__(GBP_TILE) k_boundary_partial(const DeviceGraph g) {
  const uint32_t k = blockIdx.x * GBP_TILE + threadIdx.x;
  if (k >= g.n_bnd_local) return;
  const uint32_t l = g.bnd_local[k];
  float b[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) b[i] = 0.f;
  lmk_accumulate(g, l, b);
  float4* o = g.bnd_send + (size_t)g.bnd_slot[k] * 3;
  o[0] = make_float4(b[0], b[1], b[2], b[3]);
  o[1] = make_float4(b[4], b[5], b[6], b[7]);
  o[2] = make_float4(b[8], b[9], b[10], b[11]);
}

__global__ void __launch_bounds__(GBP_TILE) k_boundary_finish(const DeviceGraph g, const int shift) {
  const uint32_t k = blockIdx.x * GBP_TILE + threadIdx.x;
  if (k >= g.n_bnd_local) return;
  const uint32_t l = g.bnd_local[k];
  const size_t slot = g.bnd_slot[k];
  float b[12];
  lmk_load_prior(g, l, b);
  for (uint32_t r = 0; r < g.world; ++r) {
    const float4* p = g.bnd_recv + ((size_t)r * g.n_bnd_global + slot) * 3;
    const float4 v0 = p[0], v1 = p[1], v2 = p[2];
    b[0] = fa(b[0], v0.x); b[1] = fa(b[1], v0.y); b[2] = fa(b[2], v0.z); b[3] = fa(b[3], v0.w);
    b[4] = fa(b[4], v1.x); b[5] = fa(b[5], v1.y); b[6] = fa(b[6], v1.z); b[7] = fa(b[7], v1.w);
    b[8] = fa(b[8], v2.x); b[9] = fa(b[9], v2.y); b[10] = fa(b[10], v2.z); b[11] = fa(b[11], v2.w);
  }
  lmk_store_belief(g, l, b, shift);
}

// RelineariseFactorVertex on every factor, active or not (ba/ba.cpp:68-97).
__global__ void __launch_bounds__(GBP_TILE) k_relinearise_all(const DeviceGraph g) {
  const uint32_t tile = blockIdx.x, tid = threadIdx.x;
  const size_t e = (size_t)tile * GBP_TILE + tid;
  float4 ra = g.recA[e];
  uint32_t flags = __float_as_uint(ra.z);
  if (flags & GBP_FLAG_PAD) return;
  const uint32_t c = g.wt_info[e >> 5].x;
  const float4 rb = g.recB[e];
  const uint32_t l = __float_as_uint(rb.z);
  float x_kf[6], x_l[3];
#pragma unroll
  for (int i = 0; i < 6; ++i) x_kf[i] = g.cam_mean[c * 6 + i];
  const float4 m = g.lmk_b[(size_t)l * GBP_LMKB_QUADS + 3];
  x_l[0] = m.x; x_l[1] = m.y; x_l[2] = m.z;
  const uint32_t robust = relinearise_record(nullptr, 0, g.fac + e, g.E_pad, nullptr, g.cam_lin + (size_t)c * 5, make_float4(g.K[0], g.K[1], g.K[2], g.K[3]),
                                             g.hp.Nstds, rb.x, rb.y, g.var[e], x_kf[0], x_kf[1], x_kf[2], x_kf[3], x_kf[4],
                                             x_kf[5], x_l[0], x_l[1], x_l[2]);
  flags = (flags & ~GBP_FLAG_ROBUST) | (robust ? GBP_FLAG_ROBUST : 0u);
  ra.z = __uint_as_float(flags);
  g.recA[e] = ra;
}

// WeakenPriorVertex on every variable (ba/ba.cpp:165-182).
__global__ void k_weaken(const DeviceGraph g) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.C) {
    const uint32_t fl = g.cam_wflag[i];
    if (fl >= 1 && fl <= 5) {  // quirk Q5
      g.cam_wflag[i] = fl - 1;
      const float s = g.cam_scaling[i];
      for (int k = 0; k < 6; ++k) g.cam_prior_eta[i * 6 + k] = fm(g.cam_prior_eta[i * 6 + k], s);
      for (int k = 0; k < 36; ++k) g.cam_prior_lam[i * 36 + k] = fm(g.cam_prior_lam[i * 36 + k], s);
    }
  } else if (i < g.C + g.L) {
    const uint32_t l = i - g.C;
    const uint32_t fl = g.lmk_wflag[l];
    if (fl >= 1 && fl <= 5) {
      g.lmk_wflag[l] = fl - 1;
      const float s = g.lmk_scaling[l];
      float4* p = g.lmk_prior + (size_t)l * 3;
      for (int q = 0; q < 3; ++q) {
        float4 v = p[q];
        v.x = fm(v.x, s); v.y = fm(v.y, s); v.z = fm(v.z, s); v.w = fm(v.w, s);
        p[q] = v;
      }
    }
  }
}

// ---- SLAM keyframe insertion on the device (ba/slam.cpp:1020-1046) ---------------------
// The reference reads the camera beliefs and all priors back (READ_PROG / READ_PRIORS), runs
// update_flags (ba/dataio.cpp:477-508) and initialise_new_kf (ba/util.cpp:183-223) on the host and
// streams priors, flags and damping counts in again (NEW_KEYFRAME).  Here the same arithmetic runs
// where the data lives.  k_kf_pose: one thread -- mean of the previous keyframe (double-precision
// LU with partial pivoting, the host restatement's solve_small), prior eta of the new keyframe,
// and the world point 1 m in front of the previous keyframe (prior mean of the new landmarks).
// kf_scratch: {pw.x, pw.y, pw.z, status bits (1 = singular belief)}.
__global__ void k_kf_pose(const DeviceGraph g, const uint32_t new_cam, float* __restrict__ kf_scratch) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const uint32_t prev = new_cam - 1;
  double A[36], b[6], x[6];
  for (int i = 0; i < 36; ++i) A[i] = (double)g.cam_b_lam[(size_t)prev * 36 + i];
  for (int i = 0; i < 6; ++i) b[i] = (double)g.cam_b_eta[(size_t)prev * 6 + i];
  bool singular = false;
  for (int k = 0; k < 6 && !singular; ++k) {
    int p = k;
    for (int i = k + 1; i < 6; ++i)
      if (fabs(A[i * 6 + k]) > fabs(A[p * 6 + k])) p = i;
    if (A[p * 6 + k] == 0.0) {
      singular = true;
      break;
    }
    if (p != k) {
      for (int j = 0; j < 6; ++j) {
        const double t = A[k * 6 + j];
        A[k * 6 + j] = A[p * 6 + j];
        A[p * 6 + j] = t;
      }
      const double t = b[k];
      b[k] = b[p];
      b[p] = t;
    }
    for (int i = k + 1; i < 6; ++i) {
      const double f = __ddiv_rn(A[i * 6 + k], A[k * 6 + k]);
      for (int j = k; j < 6; ++j) A[i * 6 + j] = __dsub_rn(A[i * 6 + j], __dmul_rn(f, A[k * 6 + j]));
      b[i] = __dsub_rn(b[i], __dmul_rn(f, b[k]));
    }
  }
  if (singular) {
    kf_scratch[3] = __uint_as_float(1u);
    return;
  }
  for (int i = 5; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < 6; ++j) s = __dsub_rn(s, __dmul_rn(A[i * 6 + j], x[j]));
    x[i] = __ddiv_rn(s, A[i * 6 + i]);
  }
  float mu[6];
  for (int i = 0; i < 6; ++i) mu[i] = (float)x[i];
  // eta of the new keyframe's prior = its (current) prior Lambda times the previous keyframe's mean
  const float* lam_new = g.cam_prior_lam + (size_t)new_cam * 36;
  for (int i = 0; i < 6; ++i) {
    float v = 0.f;
    for (int j = 0; j < 6; ++j) v = fa(v, fm(lam_new[i * 6 + j], mu[j]));
    g.cam_prior_eta[(size_t)new_cam * 6 + i] = v;
  }
  // point 1 m in front of the previous keyframe: R(w)^T ([0 0 1] - t), R in the term order of
  // ba/util.cpp:20-32 (sin/cos in double, rounded once)
  float R[9];
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.f : 0.f;
  const float w[3] = {mu[3], mu[4], mu[5]};
  const float theta = __fsqrt_rn(fa(fa(fm(w[0], w[0]), fm(w[1], w[1])), fm(w[2], w[2])));
  if (!((double)theta < 1e-6)) {
    const float H[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
    const float sn = (float)sin((double)theta), cs = (float)cos((double)theta);
    const float a = fd(sn, theta), bb = fd(fs(1.f, cs), fm(theta, theta));
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        float h2 = 0.f;
        for (int k = 0; k < 3; ++k) h2 = fa(h2, fm(H[i * 3 + k], H[k * 3 + j]));
        R[i * 3 + j] = fa(R[i * 3 + j], fa(fm(a, H[i * 3 + j]), fm(bb, h2)));
      }
  }
  const float d[3] = {fs(0.f, mu[0]), fs(0.f, mu[1]), fs(1.f, mu[2])};
  for (int i = 0; i < 3; ++i) kf_scratch[i] = fa(fa(fm(R[i], d[0]), fm(R[3 + i], d[1])), fm(R[6 + i], d[2]));
  kf_scratch[3] = __uint_as_float(0u);
}

// k_kf_apply: thread i handles edge slot i (activate the factors of the new keyframe, reset every
// damping count: ba/slam.cpp:1039-1041, quirk Q10), landmark i (weaken flag = steps for the
// landmarks this keyframe observes first, 0 otherwise; their prior eta = prior Lambda x the point in
// front of the previous keyframe -- quirk Q6, intended semantics: the test is `flag == 5` whatever
// `steps` is, ba/util.cpp:214) and camera i (weaken flag).
__global__ void __launch_bounds__(GBP_TILE) k_kf_apply(const DeviceGraph g, const uint32_t new_cam, const uint32_t steps,
                                                       const int dcount_reset, const uint32_t* __restrict__ lmk_first_cam,
                                                       const float* __restrict__ kf_scratch) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (__float_as_uint(kf_scratch[3]) != 0u) return;  // singular belief: leave the state untouched
  if (i < g.E_pad) {
    float4 ra = g.recA[i];
    uint32_t flags = __float_as_uint(ra.z);
    if (!(flags & GBP_FLAG_PAD)) {
      if (g.wt_info[i >> 5].x == new_cam) flags |= GBP_FLAG_ACTIVE;
      ra.y = __int_as_float(dcount_reset);
      ra.z = __uint_as_float(flags);
      g.recA[i] = ra;
    }
  }
  if (i < g.L) {
    const uint32_t fl = (lmk_first_cam[i] == new_cam) ? steps : 0u;
    g.lmk_wflag[i] = fl;
    if (fl == 5u) {
      const float pw[3] = {kf_scratch[0], kf_scratch[1], kf_scratch[2]};
      float4* p = g.lmk_prior + (size_t)i * 3;
      float4 q0 = p[0];
      const float4 q1 = p[1], q2 = p[2];
      const float lam[9] = {q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
      q0.x = fa(fa(fm(lam[0], pw[0]), fm(lam[1], pw[1])), fm(lam[2], pw[2]));
      q0.y = fa(fa(fm(lam[3], pw[0]), fm(lam[4], pw[1])), fm(lam[5], pw[2]));
      q0.z = fa(fa(fm(lam[6], pw[0]), fm(lam[7], pw[1])), fm(lam[8], pw[2]));
      p[0] = q0;
    }
  }
  if (i < g.C) g.cam_wflag[i] = (i == new_cam) ? steps : 0u;
}

// ---- metric (ba/util.cpp:74-144) ------------------------------------------------
// The reference inverts every belief on the host (Eigen, fp32 LU) for the error it prints.  Here
// the means used ONLY for the metric are solved in double precision (a register-resident
// LU of the full matrix), so the reported error does not carry the fp32 inversion noise of weakly
// constrained variables (a few 1e-3 relative); the sweep itself keeps using the codelets' fp32
// LDL^T means.  k_metric_prep: one thread per variable; k_metric: one thread per factor.
struct MetricPartial {
  double sum_norm, sum_sq;
  uint32_t n_relins, n_robust, n_active, pad;
};

// x = A^-1 b in double for a belief precision A (positive definite up to the fp32 rounding of its
// entries; the FULL matrix is used, like the reference's Eigen inverse, not just one triangle):
// LU without pivoting, compile-time indices (registers only), forward and backward substitution.
template <int N>
GBP_DEV void solve_double(const double (&A)[N * N], const double (&b)[N], double (&x)[N]) {
  double M[N * N], y[N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) M[i] = A[i];
  static_for<0, N>([&](auto kc) {  // Doolittle: rows below k are reduced, multipliers stored in place
    constexpr int k = decltype(kc)::value;
    const double inv = 1.0 / M[k * N + k];
    static_for<k + 1, N>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      const double f = M[r * N + k] * inv;
      M[r * N + k] = f;
      static_for<k + 1, N>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        M[r * N + c] -= f * M[k * N + c];
      });
    });
  });
  static_for<0, N>([&](auto ic) {  // L y = b
    constexpr int i = decltype(ic)::value;
    double v = b[i];
    static_for<0, i>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      v -= M[i * N + k] * y[k];
    });
    y[i] = v;
  });
  static_for<0, N>([&](auto rc) {  // U x = y
    constexpr int i = N - 1 - decltype(rc)::value;
    double v = y[i];
    static_for<i + 1, N>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      v -= M[i * N + k] * x[k];
    });
    x[i] = v / M[i * N + i];
  });
}

// met_cam: [C][16] doubles = mean 6 | R 9 | pad;  met_lmk: [L][4] doubles = mean 3 | pad
__global__ void k_metric_prep(const DeviceGraph g, double* __restrict__ met_cam, double* __restrict__ met_lmk) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.C) {
    double A[36], b[6], x[6];
    for (int k = 0; k < 36; ++k) A[k] = (double)g.cam_b_lam[i * 36 + k];
    for (int k = 0; k < 6; ++k) b[k] = (double)g.cam_b_eta[i * 6 + k];
    solve_double<6>(A, b, x);
    double* o = met_cam + (size_t)i * 16;
    for (int k = 0; k < 6; ++k) o[k] = x[k];
    // Rodrigues in double (bafuncs.cpp:32-55)
    const double w0 = x[3], w1 = x[4], w2 = x[5];
    const double th = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (th > 1e-6) {
      const double a = sin(th) / th, c = (1.0 - cos(th)) / (th * th);
      const double H[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
      for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) {
          double h2 = 0;
          for (int k = 0; k < 3; ++k) h2 += H[r * 3 + k] * H[k * 3 + q];
          R[r * 3 + q] += a * H[r * 3 + q] + c * h2;
        }
    }
    for (int k = 0; k < 9; ++k) o[6 + k] = R[k];
  } else if (i < g.C + g.L) {
    const uint32_t l = i - g.C;
    const float4* p = g.lmk_b + (size_t)l * GBP_LMKB_QUADS;
    const float4 q0 = p[0], q1 = p[1], q2 = p[2];
    double A[9] = {q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
    double b[3] = {q0.x, q0.y, q0.z}, x[3];
    solve_double<3>(A, b, x);
    double* o = met_lmk + (size_t)l * 4;
    o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = 0.0;
  }
}

struct DeviceStats {  // == gbp_iter_stats
  float reproj_mean, cost;
  uint32_t n_relins, n_robust, n_active, reserved;
};

// Fixed-order reduction of the per-block partials in double by ONE block of NT threads, then the result.
// raw != nullptr (multi-GPU): the five sums are written as doubles for k_metric_combine instead.
// cursor != nullptr (CUDA-graph replay): the result goes to out[*cursor] and the cursor advances.
struct MetricPeers {  // the peer mappings the metric exchange needs (all null on a single-GPU / NCCL handle)
  double** peer_metric;
  uint32_t** peer_mflag;
  const uint32_t* metric_step;
  uint32_t world, rank;
};
GBP_DEV MetricPeers metric_peers(const DeviceGraph& g) {
  MetricPeers px;
  px.peer_metric = g.peer_metric; px.peer_mflag = g.peer_mflag; px.metric_step = g.metric_step;
  px.world = g.world; px.rank = g.rank;
  return px;
}

template <int NT>
GBP_DEV void metric_finish(const MetricPartial* parts, const uint32_t n_parts, DeviceStats* __restrict__ out,
                           double* __restrict__ raw, uint32_t* __restrict__ cursor, const MetricPeers px) {
  __shared__ double s_d[2][NT];
  __shared__ uint32_t s_u[3][NT];
  const uint32_t tid = threadIdx.x;
  double a = 0.0, b = 0.0;
  uint32_t r = 0, ro = 0, ac = 0;
  for (uint32_t t = tid; t < n_parts; t += NT) {
    // written by other blocks of the same launch: read past L1
    const double2 d = __ldcg(reinterpret_cast<const double2*>(parts + t));
    const uint4 u = __ldcg(reinterpret_cast<const uint4*>(parts + t) + 1);
    a += d.x; b += d.y;
    r += u.x; ro += u.y; ac += u.z;
  }
  s_d[0][tid] = a; s_d[1][tid] = b; s_u[0][tid] = r; s_u[1][tid] = ro; s_u[2][tid] = ac;
  __syncthreads();
  for (int o = NT / 2; o > 0; o >>= 1) {
    if (tid < o) {
      s_d[0][tid] += s_d[0][tid + o]; s_d[1][tid] += s_d[1][tid + o];
      s_u[0][tid] += s_u[0][tid + o]; s_u[1][tid] += s_u[1][tid + o]; s_u[2][tid] += s_u[2][tid + o];
    }
    __syncthreads();
  }
  if (tid == 0 && raw) {
    raw[0] = s_d[0][0]; raw[1] = s_d[1][0];
    raw[2] = (double)s_u[0][0]; raw[3] = (double)s_u[1][0]; raw[4] = (double)s_u[2][0];
    raw[5] = raw[6] = raw[7] = 0.0;
    if (px.peer_metric) {
      // the five sums of this rank go straight into every rank's receive buffer (parity of the exchange step),
      // then the step is published in every rank's arrival flag -- the protocol of boundary_push
      const uint32_t mstep = *px.metric_step + 1u;
      const size_t off = ((size_t)(mstep & 1u) * px.world + px.rank) * 8;
      for (uint32_t r = 0; r < px.world; ++r)
        for (int i = 0; i < 8; ++i) px.peer_metric[r][off + i] = raw[i];
      __threadfence_system();
      for (uint32_t r = 0; r < px.world; ++r) st_release_sys(px.peer_mflag[r] + px.rank, mstep);
    }
  } else if (tid == 0) {
    DeviceStats s;
    s.n_active = s_u[2][0];
    s.reproj_mean = (float)(s_d[0][0] / (double)s.n_active);
    s.cost = (float)s_d[1][0];
    s.n_relins = s_u[0][0];
    s.n_robust = s_u[1][0];
    s.reserved = 0;
    if (cursor) {
      out[*cursor] = s;
      *cursor += 1;
    } else {
      *out = s;
    }
  }
}

// Persistent blocks: block b walks the 128-slot tiles b, b + gridDim.x, ... in order and every thread keeps
// its own double sums, so the number of partials the finishing block has to add is the (small, fixed) grid size
// and the summation order is fixed.  The last block to finish (a ticket counter) adds them up and writes the
// result: no separate finishing launch.
__global__ void __launch_bounds__(GBP_TILE) k_metric(const DeviceGraph g, const uint32_t n_active_total, const uint32_t n_tiles,
                                                    const double* __restrict__ met_cam, const double* __restrict__ met_lmk,
                                                    MetricPartial* out, uint32_t* __restrict__ ticket,
                                                    DeviceStats* __restrict__ stats_out, double* __restrict__ raw,
                                                    uint32_t* __restrict__ cursor) {
  __shared__ double s_f[2][GBP_TILE / 32];
  __shared__ uint32_t s_u[3][GBP_TILE / 32];
  const uint32_t tid = threadIdx.x;
  double nrm = 0.0, sq = 0.0;
  uint32_t relin = 0, robust = 0, act = 0;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const size_t e = (size_t)tile * GBP_TILE + tid;
    // the three coalesced per-slot loads are independent: issue them together
    const float4 ra = g.recA[e];
    const uint32_t eo = g.edge_orig[e];
    const float4 rb = g.recB[e];
    const uint32_t cam = g.wt_info[e >> 5].x;
    const uint32_t flags = __float_as_uint(ra.z);
    if (flags & GBP_FLAG_PAD) continue;
    robust += (flags & GBP_FLAG_ROBUST) ? 1u : 0u;
    relin += (__float_as_int(ra.y) == -g.hp.num_undamped_iters) ? 1u : 0u;
    act += (flags & GBP_FLAG_ACTIVE) ? 1u : 0u;
    // quirk Q7: the reference evaluates edges [0, n_active) of the ORIGINAL order
    if (eo < n_active_total) {
      const double* cm = met_cam + (size_t)cam * 16;  // mean 6 | R 9 (same address across the warp)
      const double* m = met_lmk + (size_t)__float_as_uint(rb.z) * 4;
      const double m0 = m[0], m1 = m[1], m2 = m[2];
      double y[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) y[i] = cm[6 + i * 3] * m0 + cm[6 + i * 3 + 1] * m1 + cm[6 + i * 3 + 2] * m2 + cm[i];
      const double iz = 1.0 / y[2];
      const double u = ((double)g.K[0] * y[0] + (double)g.K[2] * y[2]) * iz;
      const double v = ((double)g.K[1] * y[1] + (double)g.K[3] * y[2]) * iz;
      const double r0 = (double)rb.x - u, r1 = (double)rb.y - v;
      const double s2 = r0 * r0 + r1 * r1;
      nrm += sqrt(s2);
      sq += 0.5 * s2;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nrm += __shfl_down_sync(0xffffffffu, nrm, o);
    sq += __shfl_down_sync(0xffffffffu, sq, o);
    relin += __shfl_down_sync(0xffffffffu, relin, o);
    robust += __shfl_down_sync(0xffffffffu, robust, o);
    act += __shfl_down_sync(0xffffffffu, act, o);
  }
  const int w = tid >> 5;
  if ((tid & 31) == 0) {
    s_f[0][w] = nrm; s_f[1][w] = sq;
    s_u[0][w] = relin; s_u[1][w] = robust; s_u[2][w] = act;
  }
  __syncthreads();
  if (tid == 0) {
    MetricPartial p = {0.0, 0.0, 0u, 0u, 0u, 0u};
    for (int i = 0; i < GBP_TILE / 32; ++i) {
      p.sum_norm += s_f[0][i]; p.sum_sq += s_f[1][i];
      p.n_relins += s_u[0][i]; p.n_robust += s_u[1][i]; p.n_active += s_u[2][i];
    }
    out[blockIdx.x] = p;
  }
  __shared__ uint32_t s_last;
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  metric_finish<GBP_TILE>(out, gridDim.x, stats_out, raw, cursor, metric_peers(g));
  if (tid == 0) *ticket = 0u;
}


// The empty graph (no k_metric launch): zero sums through the same finishing code.
__global__ void __launch_bounds__(GBP_TILE) k_metric_finish(const DeviceGraph g, const MetricPartial* __restrict__ parts,
                                                           const uint32_t n_parts, DeviceStats* __restrict__ out,
                                                           double* __restrict__ raw, uint32_t* __restrict__ cursor) {
  metric_finish<GBP_TILE>(parts, n_parts, out, raw, cursor, metric_peers(g));
}

// multi-GPU: sum the per-rank metric sums [world][8] in rank order.  all != nullptr: the NCCL all-gathered copy;
// else the peer-to-peer receive buffer of the current exchange step, once every rank's flag has arrived.
// cursor != nullptr (CUDA-graph replay): the result goes to out[*cursor] and the cursor advances.
__global__ void k_metric_combine(const DeviceGraph g, const double* __restrict__ all, DeviceStats* __restrict__ out,
                                 uint32_t* __restrict__ cursor) {
  if (blockIdx.x != 0) return;
  __shared__ uint32_t s_timed_out;
  if (threadIdx.x == 0) s_timed_out = 0u;
  __syncthreads();
  const double* src = all;
  uint32_t mstep = 0;
  if (!all) {
    mstep = *g.metric_step + 1u;
    if (threadIdx.x < g.world) {
      const long long t0 = clock64();
      while ((int32_t)(ld_acquire_sys(g.metric_flag + threadIdx.x) - mstep) < 0) {
        if (clock64() - t0 > g.p2p_timeout) {
          *g.p2p_error = 1u;
          s_timed_out = 1u;
          break;
        }
        __nanosleep(100);
      }
    }
    src = g.metric_recv + (size_t)(mstep & 1u) * g.world * 8;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  if (!all) *g.metric_step = mstep;
  double a[5] = {0, 0, 0, 0, 0};
  if (!s_timed_out)
    for (uint32_t r = 0; r < g.world; ++r)
      for (int i = 0; i < 5; ++i) a[i] += __ldcg(src + r * 8 + i);
  DeviceStats s;
  s.n_active = (uint32_t)a[4];
  s.reproj_mean = (float)(a[0] / a[4]);
  s.cost = (float)a[1];
  s.n_relins = (uint32_t)a[2];
  s.n_robust = (uint32_t)a[3];
  s.reserved = 0;
  if (cursor) {
    out[*cursor] = s;
    *cursor += 1;
  } else {
    *out = s;
  }
}

// READ_PROG helpers: gather per-edge scalars back into the reference's edge order
// and unpack the landmark belief records.
__global__ void k_export_edges(const DeviceGraph g, const uint32_t* __restrict__ pos_of_orig,
                               float* __restrict__ damping, int32_t* __restrict__ dcount,
                               uint32_t* __restrict__ robust) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.E) return;
  const float4 ra = g.recA[pos_of_orig[i]];
  damping[i] = ra.x;
  dcount[i] = __float_as_int(ra.y);
  robust[i] = (__float_as_uint(ra.z) & GBP_FLAG_ROBUST) ? 1u : 0u;
}

__global__ void k_export_lmk_beliefs(const DeviceGraph g, float* __restrict__ eta, float* __restrict__ lam) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= g.L) return;
  const float4* p = g.lmk_b + (size_t)l * GBP_LMKB_QUADS;
  const float4 a = p[0], b = p[1], c = p[2];
  eta[l * 3] = a.x; eta[l * 3 + 1] = a.y; eta[l * 3 + 2] = a.z;
  float* o = lam + (size_t)l * 9;
  o[0] = a.w; o[1] = b.x; o[2] = b.y; o[3] = b.z; o[4] = b.w; o[5] = c.x; o[6] = c.y; o[7] = c.z; o[8] = c.w;
}

// NEW_KEYFRAME / set_tensor helper: scatter per-edge scalars given in the
// reference's edge order into the edge-slot records (null pointer = keep).
__global__ void k_import_edges(const DeviceGraph g, const uint32_t* __restrict__ pos_of_orig,
                               const float* __restrict__ damping, const int32_t* __restrict__ dcount,
                               const uint32_t* __restrict__ active, const uint32_t* __restrict__ robust,
                               const float* __restrict__ dmu, const int clear_muvalid) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.E) return;
  const uint32_t p = pos_of_orig[i];
  float4 ra = g.recA[p];
  uint32_t flags = __float_as_uint(ra.z);
  if (damping) ra.x = damping[i];
  if (dcount) ra.y = __int_as_float(dcount[i]);
  if (active) flags = (flags & ~GBP_FLAG_ACTIVE) | ((active[i] == 1u) ? GBP_FLAG_ACTIVE : 0u);
  if (robust) flags = (flags & ~GBP_FLAG_ROBUST) | (robust[i] ? GBP_FLAG_ROBUST : 0u);
  if (clear_muvalid) flags &= ~GBP_FLAG_MUVALID;
  if (dmu) ra.w = dmu[i];
  ra.z = __uint_as_float(flags);
  g.recA[p] = ra;
}

// set_tensor on a belief tensor: recompute the per-variable means (and camera
// rotations) from the stored beliefs without re-summing the messages.
__global__ void k_means_from_beliefs(const DeviceGraph g) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.C) {
    float eta[6], lamL[21], mean[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) eta[k] = g.cam_b_eta[i * 6 + k];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int j = 0; j <= r; ++j) lamL[lt(r, j)] = g.cam_b_lam[i * 36 + r * 6 + j];
    inf2mean6(eta, lamL, mean);
    const float w[3] = {mean[3], mean[4], mean[5]};
    float R[9], num[9], den;
    cam_lin_consts(w, R, num, den);
    float* rec = reinterpret_cast<float*>(g.cam_rec + (size_t)i * 16);
    float acc6 = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      g.cam_mean[i * 6 + k] = mean[k];
      rec[k] = eta[k];
      rec[42 + k] = mean[k];
      const float prev = g.cam_mean_prev[i * 6 + k];
      rec[48 + k] = prev;
      const float d = fs(prev, mean[k]);
      acc6 = fa(acc6, fm(d, d));
    }
    rec[54] = acc6;
#pragma unroll
    for (int k = 0; k < 36; ++k) rec[6 + k] = g.cam_b_lam[i * 36 + k];
    float* lin = reinterpret_cast<float*>(g.cam_lin + (size_t)i * 5);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      lin[k] = R[k];
      lin[9 + k] = num[k];
    }
    lin[18] = den;
  } else if (i < g.C + g.L) {
    const uint32_t l = i - g.C;
    float4* o = g.lmk_b + (size_t)l * GBP_LMKB_QUADS;
    const float4 a = o[0], b = o[1], c = o[2];
    const float eta[3] = {a.x, a.y, a.z};
    const float lam[9] = {a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    float mean[3];
    inf2mean3(eta, lam, mean);
    o[3] = make_float4(mean[0], mean[1], mean[2], 0.f);
    g.lmk_sq[l] = lmk_sq_terms(g.lmk_mean_prev[l], mean);
  }
}

}  // namespace gbp

#include "gbp_sweep_tma.cuh"
