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

// ---- factor potential record: 56 floats = 14 quads -------------------------
// [eta 9 | Lambda_ll lower 6 | Lambda_cl 18 (6x3) | Lambda_cc lower 21 | pad 2].
// Lambda_lc is not stored: the reference always sets it to Lambda_cl^T
// (gbp_codelets.cpp:158-162,363-367).  Lambda_cc and Lambda_ll are stored as
// packed lower triangles: the reference builds them as J^T J (+ the previous
// block, quirk Q1) divided by a scalar, entry (i,j) and (j,i) by the very same
// fp32 operations (matlib.cpp:60-68, gbp_codelets.cpp:109-125,313-329), so the
// two triangles are bit-identical and storing one loses nothing.
#define GBP_FAC_QUADS 14
#define GBP_FAC_ETA 0
#define GBP_FAC_LL 9    // 6 floats, lower triangle row-major (i>=j)
#define GBP_FAC_CL 15   // 18 floats, 6x3 row-major
#define GBP_FAC_CC 33   // 21 floats, lower triangle row-major (i>=j)

// ---- factor->camera message record: 28 floats = 7 quads ----------------------
// [eta 6 | lower triangle of Lambda (21, row-major i>=j) | pad].
// This is everything the algorithm ever READS of a camera-bound message: inv6x6
// only touches the lower triangle of (Lambda_cc + belief - message)
// (matlib.cpp:193-206), and the belief sum over the FULL 6x6 message is formed
// on chip by the kernel that computes it.  The strict upper triangle (equal to
// the lower one up to fp32 rounding of the Schur complement) is therefore not
// stored; get_tensor mirrors the lower triangle.
#define GBP_MCAM_QUADS 7
#define GBP_MCAM_LOWER 6

// ---- factor->landmark message record: 12 floats = 3 quads [eta 3 | Lambda 9] --------
// Stored in LANDMARK order, not edge-slot order: the message of edge e lives at
// lmk_ptr[landmark(e)] + (number of earlier edges of that landmark) -- the reference's
// message slot (ba/ba.cpp:267-279) made dense.  The belief update then streams each
// landmark's messages as one contiguous run (no index indirection, no gather), and the
// scattered 48-byte accesses move into the factor kernel, whose software pipeline hides
// their latency.
#define GBP_MLMK_QUADS 3

// ---- per-landmark belief record: 16 floats = 4 quads (64 B, sector aligned) --
// [eta 3 | Lambda 9 | mean 3 | pad]
#define GBP_LMKB_QUADS 4

// ---- per-edge state quads -----------------------------------------------------
// recA (read+write): {damping, damping_count (int bits), flags (uint bits), dmu}
// recB (read only) : {z.x, z.y, landmark id, position of the landmark-bound message (uint bits)}
// var  (read only) : meas_variance, separate because only a relinearising factor reads it
#define GBP_FLAG_ACTIVE 1u    // active_flag                         ba/ba.cpp:765
#define GBP_FLAG_ROBUST 2u    // robust_flag                         ba/ba.cpp:766
#define GBP_FLAG_MUVALID 4u   // oldmu of this edge == previous mean of its variables
#define GBP_FLAG_HASMSG 8u    // the message records of this edge are non-zero
#define GBP_FLAG_PAD 16u      // padding slot (no factor)

// ---- per-tile camera partial: 42 floats [eta 6 | Lambda 36 row-major] -----------
#define GBP_CAMPART 42
#define GBP_CAMPART_STRIDE 44  // floats between the partials of consecutive warp-tiles: 176 bytes, so that a camera's run of
                               // partials is 16-byte aligned and sized (one bulk copy stages it in the belief update)

#ifdef __CUDACC__
#define GBP_HD __host__ __device__
#else
#define GBP_HD
#endif

GBP_HD inline int gbp_lt(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j
GBP_HD inline int gbp_sym(int i, int j) { return i >= j ? gbp_lt(i, j) : gbp_lt(j, i); }
// index of (i,j), i<j, in a row-major packed strict upper triangle of a 6x6
GBP_HD inline int gbp_upper(int i, int j) { return (5 * i - i * (i - 1) / 2) + (j - i - 1); }
// field index of Lambda(i,j) inside the camera message record (upper triangle mirrored)
GBP_HD inline int gbp_mcam_lam_field(int i, int j) { return GBP_MCAM_LOWER + gbp_sym(i, j); }
