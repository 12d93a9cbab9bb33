/*
 * gbp_host.h -- C ABI of the host-side problem setup that surrounds the GBP
 * hot path (pure host code, no CUDA): the BAL-format loader, the prior /
 * scaling / flag preparation of the reference's ba.cpp and slam.cpp mains,
 * and the SLAM keyframe bookkeeping.  It produces the `gbp_problem` that
 * gbp_cuda_init consumes.  Reference paths are relative to the reference tree.
 */
#ifndef GBP_HOST_H_
#define GBP_HOST_H_

#include <stdint.h>

#include "gbp_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Parsed input file = BALProblem (include/dataio.h:13-75, ba/dataio.cpp:17-57). */
typedef struct gbp_bal gbp_bal;

/* The flags of ./ba and ./slam (ba/ba.cpp:394-476, ba/slam.cpp:394-476). */
typedef struct gbp_cli_options {
  int n_iters;                   /* --n_iters 1500 (ba only)            */
  int iters_between_kfs;         /* --iters_between_kfs 700 (slam only) */
  int n_ipus;                    /* --ipus 1  -> number of GPUs         */
  int cams_per_tile;             /* --camspertile 1 (accepted, unused)  */
  int profile;                   /* --profile false                     */
  float transnoise;              /* --tn 0                              */
  float rotnoise;                /* --rn 0                              */
  float lmktrans_noise;          /* --ltn 0                             */
  int av_depth_on;               /* --avdepth_on false                  */
  float av_depth;                /* --avdepth 1                         */
  float reproj_meas_var;         /* --reproj_meas_var 4                 */
  float prior_std_weaker_factor; /* --prior_std_weaker_factor 100       */
  float first_cam_prior_std;     /* --first_cam_prior_std 0.01          */
  float steps;                   /* --steps 5                           */
  int iters_before_damping;      /* --undamped_start 15                 */
  int verbose;                   /* --v false                           */
  uint32_t noise_seed;           /* 0 = clock-seeded like the reference (ba/dataio.cpp:334) */
} gbp_cli_options;

void gbp_cli_options_default(gbp_cli_options* o);

/* BALProblem::LoadFile.  GBP_ERR_IO if the file cannot be opened; a short /
 * malformed file prints "Invalid UW data file." once per bad token and
 * carries on, like the reference (ba/dataio.cpp:59-65). */
int gbp_bal_load(const char* path, gbp_bal** out);
/* Same from caller memory (used by the synthetic generator and tests):
 * cameras[6C] and points[3L] as doubles, observations[2E]. */
int gbp_bal_from_arrays(uint32_t C, uint32_t L, uint32_t E, const double intrinsics[4],
                        const uint32_t* cam_idx, const uint32_t* lmk_idx, const double* observations,
                        const double* cameras, const double* points, gbp_bal** out);
int gbp_bal_save(const gbp_bal* b, const char* path);
/* A copy of `b` whose camera / point parameters are the means of the given beliefs (what READ_PROG
 * returns): the optimised problem in the input format -- the purpose of the reference's never-called
 * save_cam_means / save_lmk_means (ba/dataio.cpp:205-255). */
int gbp_bal_with_means(const gbp_bal* b, const float* cam_beliefs_eta, const float* cam_beliefs_lambda,
                       const float* lmk_beliefs_eta, const float* lmk_beliefs_lambda, gbp_bal** out);
void gbp_bal_free(gbp_bal* b);
int gbp_bal_dims(const gbp_bal* b, uint32_t* C, uint32_t* L, uint32_t* E);
const uint32_t* gbp_bal_camera_index(const gbp_bal* b);
const uint32_t* gbp_bal_point_index(const gbp_bal* b);
const double* gbp_bal_observations(const gbp_bal* b);
const double* gbp_bal_parameters(const gbp_bal* b);     /* 6C camera then 3L point doubles */
const double* gbp_bal_intrinsics(const gbp_bal* b);     /* fx fy cx cy */

/* Everything ba.cpp:489-590 / slam.cpp:489-597 prepares on the host before
 * WRITE_PROG: measurements, prior means (+ optional noise / average-depth
 * init), set_prior_lambda (ba/dataio.cpp:67-117, O(E) here), weakening
 * scale factors (ba/ba.cpp:560-572), damping state and the BA or SLAM flag
 * schedule (ba/ba.cpp:588-590, ba/dataio.cpp:455-475). */
typedef struct gbp_setup gbp_setup;
#define GBP_MODE_BA 0
#define GBP_MODE_SLAM 1
int gbp_setup_create(const gbp_bal* b, const gbp_cli_options* o, int mode, gbp_setup** out);
void gbp_setup_free(gbp_setup* s);
const gbp_problem* gbp_setup_problem(const gbp_setup* s);

/* SLAM keyframe insertion, host half (ba/slam.cpp:1020-1041):
 * update_flags (ba/dataio.cpp:477-508) + initialise_new_kf (ba/util.cpp:183-223)
 * + damping_count reset to -15.  The caller passes what READ_PROG / READ_PRIORS
 * returned; the four prior arrays are updated in place and, together with the
 * setup's flag arrays (gbp_setup_problem), are what NEW_KEYFRAME streams in.
 * Returns the number of newly observed landmarks in *n_new_lmks. */
int gbp_setup_next_keyframe(gbp_setup* s, const float* cam_beliefs_eta,
                            const float* cam_beliefs_lambda, float* cam_priors_eta,
                            float* cam_priors_lambda, float* lmk_priors_eta,
                            float* lmk_priors_lambda, int32_t* damping_count, int* n_new_lmks);
/* data_counter of slam.cpp (number of inserted keyframes so far). */
int gbp_setup_data_counter(const gbp_setup* s);

/* Synthetic BAL-format problem: cameras on a smooth
 * trajectory, landmarks in the viewing frusta, each seen by ~obs_per_point
 * nearby cameras, edges sorted by camera.  Deterministic in `seed`. */
int gbp_synth_generate(uint32_t n_cameras, uint32_t n_points, double obs_per_point, uint32_t seed,
                       gbp_bal** out);

#ifdef __cplusplus
}
#endif
#endif /* GBP_HOST_H_ */
