// Device data layout of the GBP hot path, shared by kernels and host-side
// (de)serialisation.  Reference tensors (ba/ba.cpp:665-687,759-775) are AoS and
// padded to the maximum degree; here every per-factor tensor is stored in
// "quad-SoA" form: record field f of edge slot e lives in float4 number
// (f/4)*E_pad + e, lane f%4, so that a warp of 32 consecutive factors moves
// 512 contiguous bytes per 128-bit load/store.
//
// Edge slots: the factors of one camera are contiguous (stable in input order)
// and padded to a multiple of GBP_TILE; a thread block owns one tile, hence one
// camera, and reduces its camera-bound messages on chip.
#pragma once
#include <stdint.h>

#define GBP_TILE 128  // factors per thread block (one thread per factor)

// ---- factor potential record: 72 floats = 18 quads -------------------------
// [eta 9 | Lambda_ll 9 | Lambda_cl 18 (6x3) | Lambda_cc 36].  Lambda_lc is not
// stored: the reference always sets it to Lambda_cl^T (gbp_codelets.cpp:158-162,363-367).
#define GBP_FAC_QUADS 18
#define GBP_FAC_ETA 0
#define GBP_FAC_LL 9
#define GBP_FAC_CL 18
#define GBP_FAC_CC 36

// ---- factor->camera message record: 44 floats = 11 quads --------------------
// [eta 6 | lower triangle of Lambda (21, row-major i>=j) | pad | strict upper (15) | pad]
// The first 7 quads are everything a sweep READS back (inv6x6 only touches the
// lower triangle, matlib.cpp:193-206); the upper part is kept for the belief sum.
#define GBP_MCAM_QUADS 11
#define GBP_MCAM_READ_QUADS 7
#define GBP_MCAM_LOWER 6
#define GBP_MCAM_UPPER 28

// ---- factor->landmark message record: 12 floats = 3 quads, AoS per edge slot
// [eta 3 | Lambda 9] so the landmark-side gather reads 48 contiguous bytes.
#define GBP_MLMK_QUADS 3

// ---- per-landmark belief record: 16 floats = 4 quads (64 B, sector aligned) --
// [eta 3 | Lambda 9 | mean 3 | pad]
#define GBP_LMKB_QUADS 4

// ---- per-edge state quads -----------------------------------------------------
// recA (read+write): {damping, damping_count (int bits), flags (uint bits), dmu}
// recB (read only) : {z.x, z.y, meas_variance, landmark id (uint bits)}
#define GBP_FLAG_ACTIVE 1u    // active_flag                         ba/ba.cpp:765
#define GBP_FLAG_ROBUST 2u    // robust_flag                         ba/ba.cpp:766
#define GBP_FLAG_MUVALID 4u   // oldmu of this edge == previous mean of its variables
#define GBP_FLAG_HASMSG 8u    // the message records of this edge are non-zero
#define GBP_FLAG_PAD 16u      // padding slot (no factor)

// ---- per-tile camera partial: 42 floats [eta 6 | Lambda 36 row-major] -----------
#define GBP_CAMPART 42

#ifdef __CUDACC__
#define GBP_HD __host__ __device__
#else
#define GBP_HD
#endif

GBP_HD inline int gbp_lt(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j
// field index of Lambda(i,j) inside the camera message record
GBP_HD inline int gbp_mcam_lam_field(int i, int j) {
  if (i >= j) return GBP_MCAM_LOWER + gbp_lt(i, j);
  return GBP_MCAM_UPPER + (5 * i - i * (i - 1) / 2) + (j - i - 1);
}
