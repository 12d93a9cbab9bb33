// TEST INFRASTRUCTURE ONLY.
//
// Adapter that drives the reference's own vertex classes, compiled UNMODIFIED
// from /root/reference/ba/gbp_codelets.cpp behind oracle/shim, through the
// same function signatures as oracle/gbp_restated.hpp.  Only built into
// oracle/_ref/libgbp_ref.so (oracle/Makefile target `ref`), and only where
// /root/reference exists.  No reference source is copied into this repo.
#pragma once
#include <cstdint>

#include "gbp_restated.hpp"  // for Hyper only

// The reference translation unit (defines the vertex classes and the five
// hyper-parameter globals; must be included exactly once).
#include "gbp_codelets.cpp"

namespace gbp_ref_backend {

using gbp_restated::Hyper;

inline void set_hyper(const Hyper& hp) {
  ::maxeta_damping = hp.maxeta_damping;
  ::num_undamped_iters = hp.num_undamped_iters;
  ::dmu_threshold = hp.dmu_threshold;
  ::min_linear_iters = hp.min_linear_iters;
  ::Nstds = hp.Nstds;
}

static unsigned kSix = 6, kThree = 3;

template <class V>
inline void bind_factor_blocks(V& v, float* f_eta, float* f_lam) {
  v.factor_eta_.bind(f_eta, 9);
  v.factor_lambda_cc_.bind(f_lam, 36);        // ba/ba.cpp:93
  v.factor_lambda_cl_.bind(f_lam + 36, 18);   // ba/ba.cpp:95
  v.factor_lambda_lc_.bind(f_lam + 54, 18);   // ba/ba.cpp:96
  v.factor_lambda_ll_.bind(f_lam + 72, 9);    // ba/ba.cpp:94
}

inline void relinearise_factor(const float* z, float var, const float* K, const float* kf_eta,
                               const float* kf_lam, const float* lmk_eta, const float* lmk_lam,
                               const Hyper&, float* f_eta, float* f_lam, uint32_t* robust) {
  RelineariseFactorVertex v;
  v.measurement.bind(const_cast<float*>(z), 2);
  v.meas_variance.bind(&var);
  v.K_.bind(const_cast<float*>(K), 9);
  v.kf_belief_eta_.bind(const_cast<float*>(kf_eta), 6);
  v.kf_belief_lambda_.bind(const_cast<float*>(kf_lam), 36);
  v.lmk_belief_eta_.bind(const_cast<float*>(lmk_eta), 3);
  v.lmk_belief_lambda_.bind(const_cast<float*>(lmk_lam), 9);
  bind_factor_blocks(v, f_eta, f_lam);
  v.robust_flag.bind(robust);
  v.compute();
}

inline void prep_message(uint32_t active, float* damping, int32_t* damping_count, uint32_t* robust,
                         const float* z, float var, const float* K, const float* kf_eta,
                         const float* kf_lam, const float* lmk_eta, const float* lmk_lam,
                         const float* oldmu, float* mu, float* dmu, const Hyper&, float* f_eta,
                         float* f_lam) {
  PrepMessageVertex v;
  v.damping.bind(damping);
  v.damping_count.bind(damping_count);
  v.active_flag.bind(&active);
  v.robust_flag.bind(robust);
  v.measurement.bind(const_cast<float*>(z), 2);
  v.K_.bind(const_cast<float*>(K), 9);
  v.meas_variance.bind(&var);
  v.kf_belief_eta_.bind(const_cast<float*>(kf_eta), 6);
  v.kf_belief_lambda_.bind(const_cast<float*>(kf_lam), 36);
  v.lmk_belief_eta_.bind(const_cast<float*>(lmk_eta), 3);
  v.lmk_belief_lambda_.bind(const_cast<float*>(lmk_lam), 9);
  v.oldmu.bind(const_cast<float*>(oldmu), 9);
  v.mu.bind(mu, 9);
  v.dmu.bind(dmu);
  bind_factor_blocks(v, f_eta, f_lam);
  v.compute();
}

inline void cam_message_eta(uint32_t active, float damping, const float* f_eta, const float* f_lam,
                            const float* lmk_b_eta, const float* lmk_b_lam, const float* p_lmk_eta,
                            const float* p_lmk_lam, const float* p_cam_eta, float* out) {
  ComputeCamMessageEtaVertex v;     // wiring: ba/ba.cpp:303-320
  v.damping.bind(&damping);
  v.active_flag.bind(&active);
  v.outedge_dofs.bind(&kSix);
  v.nonoutedge_dofs.bind(&kThree);
  v.f_outedge_eta_.bind(const_cast<float*>(f_eta), 6);
  v.f_nonoutedge_eta_.bind(const_cast<float*>(f_eta) + 6, 3);
  v.f_noe_noe_lambda_.bind(const_cast<float*>(f_lam) + 72, 9);
  v.f_oe_noe_lambda_.bind(const_cast<float*>(f_lam) + 36, 18);
  v.belief_nonoutedge_eta_.bind(const_cast<float*>(lmk_b_eta), 3);
  v.belief_nonoutedge_lambda_.bind(const_cast<float*>(lmk_b_lam), 9);
  v.pmess_nonoutedge_eta_.bind(const_cast<float*>(p_lmk_eta), 3);
  v.pmess_nonoutedge_lambda_.bind(const_cast<float*>(p_lmk_lam), 9);
  v.pmess_outedge_eta_.bind(const_cast<float*>(p_cam_eta), 6);
  v.mess_outedge_eta_.bind(out, 6);
  v.compute();
}

inline void lmk_message_eta(uint32_t active, float damping, const float* f_eta, const float* f_lam,
                            const float* cam_b_eta, const float* cam_b_lam, const float* p_cam_eta,
                            const float* p_cam_lam, const float* p_lmk_eta, float* out) {
  ComputeLmkMessageEtaVertex v;     // wiring: ba/ba.cpp:336-353
  v.damping.bind(&damping);
  v.active_flag.bind(&active);
  v.outedge_dofs.bind(&kThree);
  v.nonoutedge_dofs.bind(&kSix);
  v.f_outedge_eta_.bind(const_cast<float*>(f_eta) + 6, 3);
  v.f_nonoutedge_eta_.bind(const_cast<float*>(f_eta), 6);
  v.f_noe_noe_lambda_.bind(const_cast<float*>(f_lam), 36);
  v.f_oe_noe_lambda_.bind(const_cast<float*>(f_lam) + 54, 18);
  v.belief_nonoutedge_eta_.bind(const_cast<float*>(cam_b_eta), 6);
  v.belief_nonoutedge_lambda_.bind(const_cast<float*>(cam_b_lam), 36);
  v.pmess_nonoutedge_eta_.bind(const_cast<float*>(p_cam_eta), 6);
  v.pmess_nonoutedge_lambda_.bind(const_cast<float*>(p_cam_lam), 36);
  v.pmess_outedge_eta_.bind(const_cast<float*>(p_lmk_eta), 3);
  v.mess_outedge_eta_.bind(out, 3);
  v.compute();
}

inline void cam_message_lambda(uint32_t active, const float* f_lam, const float* lmk_b_lam,
                               const float* p_lmk_lam, float* out) {
  ComputeCamMessageLambdaVertex v;  // wiring: ba/ba.cpp:322-333
  v.outedge_dofs.bind(&kSix);
  v.nonoutedge_dofs.bind(&kThree);
  v.active_flag.bind(&active);
  v.f_oe_oe_lambda_.bind(const_cast<float*>(f_lam), 36);
  v.f_noe_noe_lambda_.bind(const_cast<float*>(f_lam) + 72, 9);
  v.f_oe_noe_lambda_.bind(const_cast<float*>(f_lam) + 36, 18);
  v.f_noe_oe_lambda_.bind(const_cast<float*>(f_lam) + 54, 18);
  v.belief_nonoutedge_lambda_.bind(const_cast<float*>(lmk_b_lam), 9);
  v.pmess_nonoutedge_lambda_.bind(const_cast<float*>(p_lmk_lam), 9);
  v.mess_outedge_lambda_.bind(out, 36);
  v.compute();
}

inline void lmk_message_lambda(uint32_t active, const float* f_lam, const float* cam_b_lam,
                               const float* p_cam_lam, float* out) {
  ComputeLmkMessageLambdaVertex v;  // wiring: ba/ba.cpp:355-366
  v.outedge_dofs.bind(&kThree);
  v.nonoutedge_dofs.bind(&kSix);
  v.active_flag.bind(&active);
  v.f_oe_oe_lambda_.bind(const_cast<float*>(f_lam) + 72, 9);
  v.f_noe_noe_lambda_.bind(const_cast<float*>(f_lam), 36);
  v.f_oe_noe_lambda_.bind(const_cast<float*>(f_lam) + 54, 18);
  v.f_noe_oe_lambda_.bind(const_cast<float*>(f_lam) + 36, 18);
  v.belief_nonoutedge_lambda_.bind(const_cast<float*>(cam_b_lam), 36);
  v.pmess_nonoutedge_lambda_.bind(const_cast<float*>(p_cam_lam), 36);
  v.mess_outedge_lambda_.bind(out, 9);
  v.compute();
}

inline void weaken_prior(float scaling, uint32_t* flag, float* eta, int n_eta, float* lam,
                         int n_lam) {
  WeakenPriorVertex v;              // wiring: ba/ba.cpp:165-182
  v.scaling.bind(&scaling);
  v.weaken_flag.bind(flag);
  v.prior_eta.bind(eta, n_eta);
  v.prior_lambda.bind(lam, n_lam);
  v.compute();
}

}  // namespace gbp_ref_backend
