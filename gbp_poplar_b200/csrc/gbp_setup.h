// Device-side problem setup (gbp_setup.cu): interface between the C ABI layer and the setup kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace gbp {

struct SetupInputs {  // DEVICE copies of the caller's raw arrays (gbp_problem); optional ones may be null
  uint32_t E, C, L;
  const uint32_t* cam_ids;        // [E]
  const uint32_t* lmk_ids;        // [E]
  const float* measurements;      // [2E]
  const float* meas_variances;    // [E]
  const uint32_t* active_flag;    // [E] or null (all active)
  const float* damping;           // [E] or null (0)
  const int32_t* damping_count;   // [E] or null (-15)
  const uint32_t* edge_global;    // [E] or null (edge e is global edge e)
  const float* lmk_priors_eta;    // [3L]
  const float* lmk_priors_lambda; // [9L]
};

struct SetupTemp {  // scratch of the two stages, carved out of one device block of setup_temp_bytes()
  uint32_t *deg_c, *tiles, *cam_first, *deg_l, *iota, *keys_out, *by_cam, *by_lmk, *lpos;
  uint32_t* info;  // {total warp-tiles, max camera degree, max landmark degree, index-range error}
  void* cub_temp;
  size_t cub_bytes;
};

struct SetupOutputs {  // arrays of the handle's arena that stage B fills
  uint32_t E_pad, lmk_per_block;
  float4* recA;
  float4* recB;
  float* var;
  uint32_t* edge_orig;
  uint2* wt_info;
  uint32_t* cam_wt_begin;   // [C+1]
  uint32_t* lmk_ptr;        // [L+1]
  uint32_t* pos_of_orig;    // [E]
  uint32_t* lmk_first_cam;  // [L]
  float4* lmk_prior;        // [L][3]
  uint4* lmk_blk;           // [ceil(L / lmk_per_block)]
};

size_t setup_temp_bytes(uint32_t E, uint32_t C, uint32_t L);
// both return 0 or a cudaError_t
int setup_stage_a(cudaStream_t s, const SetupInputs& in, char* temp, SetupTemp* t);
int setup_stage_b(cudaStream_t s, const SetupInputs& in, const SetupTemp& t, const SetupOutputs& out);

}  // namespace gbp
