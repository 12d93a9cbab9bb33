// TEST INFRASTRUCTURE ONLY -- the CPU oracle; never linked into the product.
//
// Plain C++ restatement of the arithmetic of the reference's GBP codelets
// (/root/reference/ba/gbp_codelets.cpp, matlib.cpp, bafuncs.cpp).  Every
// function keeps the reference's fp32 operation ORDER (including the
// accumulate-into-output behaviour of matMul, matlib.cpp:54) so that, built
// with `-O2 -ffp-contract=off`, it is bit-identical to the reference sources
// compiled behind oracle/shim (oracle/_ref/libgbp_ref.so).  That identity is
// what tests/test_oracle_pin.py checks, together with the committed golden
// vectors under tests/golden/ (generated from oracle/_ref by
// tests/golden/make_golden.py).
//
// Parity status: PINNED against the reference's own codelet sources compiled
// in this repo's authoring container (the reference ships no tests or golden
// vectors of its own -- SURVEY.md section 4).
#pragma once
#include <cmath>
#include <cstdint>

namespace gbp_restated {

struct Hyper {               // gbp_codelets.cpp:11-16
  float maxeta_damping = 0.4f;
  int num_undamped_iters = 8;
  float dmu_threshold = 3e-3f;
  int min_linear_iters = 10;
  float Nstds = 2.5f;
};

// ---- tiny dense helpers: all ACCUMULATE into C like matlib.cpp:47-89 -------
// C[MxN] += A[MxK] * B[KxN]                      (matlib.cpp:50-58)
template <int M, int K, int N>
inline void mm_acc(const float* A, const float* B, float* C) {
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < K; ++k) C[i * N + j] += A[i * K + k] * B[k * N + j];
}
// C[MxN] += A^T * B with A stored [KxM], B [KxN]  (matlib.cpp:60-68)
template <int K, int M, int N>
inline void mm_tn_acc(const float* A, const float* B, float* C) {
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < K; ++k) C[i * N + j] += A[k * M + i] * B[k * N + j];
}
// C[MxN] += A * B^T with A [MxK], B stored [NxK]  (matlib.cpp:70-78)
template <int M, int K, int N>
inline void mm_nt_acc(const float* A, const float* B, float* C) {
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < K; ++k) C[i * N + j] += A[i * K + k] * B[j * K + k];
}

// Closed-form 3x3 inverse, nine divisions by the determinant (matlib.cpp:143-161).
inline void inv3(const float* M, float* R) {
  const float det = M[0] * (M[4] * M[8] - M[7] * M[5]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
                    M[2] * (M[3] * M[7] - M[4] * M[6]);
  R[0] = (M[4] * M[8] - M[7] * M[5]) / det;
  R[1] = (M[2] * M[7] - M[1] * M[8]) / det;
  R[2] = (M[1] * M[5] - M[2] * M[4]) / det;
  R[3] = (M[5] * M[6] - M[3] * M[8]) / det;
  R[4] = (M[0] * M[8] - M[2] * M[6]) / det;
  R[5] = (M[3] * M[2] - M[0] * M[5]) / det;
  R[6] = (M[3] * M[7] - M[6] * M[4]) / det;
  R[7] = (M[6] * M[1] - M[0] * M[7]) / det;
  R[8] = (M[0] * M[4] - M[3] * M[1]) / det;
}

// 6x6 inverse by un-pivoted LDL^T, explicit inverse of the unit upper factor,
// then two dense products (matlib.cpp:163-222).  Reads only the LOWER
// triangle of A.  Ainv must come in zeroed (it is accumulated into).
inline void inv6(const float* A, float* Ainv) {
  float D[36] = {}, LT[36] = {}, LTinv[36] = {}, T[36] = {};
  for (int j = 0; j < 6; ++j) {                                  // matlib.cpp:193-206
    LT[j * 6 + j] = 1.0f;
    D[j * 6 + j] = A[j * 6 + j];
    for (int k = 0; k < j; ++k) D[j * 6 + j] -= LT[k * 6 + j] * LT[k * 6 + j] * D[k * 6 + k];
    for (int i = j + 1; i < 6; ++i) {
      LT[j * 6 + i] = (1 / D[j * 6 + j]) * A[i * 6 + j];
      for (int k = 0; k < j; ++k)
        LT[j * 6 + i] -= (1 / D[j * 6 + j]) * LT[k * 6 + i] * LT[k * 6 + j] * D[k * 6 + k];
    }
  }
  for (int j = 0; j < 6; ++j) D[j * 6 + j] = 1 / D[j * 6 + j];   // matlib.cpp:209-212
  for (int j = 0; j < 6; ++j) {                                  // matlib.cpp:163-178
    LTinv[j * 6 + j] = 1 / LT[j * 6 + j];
    for (int i = 0; i < j; ++i)
      for (int k = 0; k < j; ++k) LTinv[i * 6 + j] += LTinv[i * 6 + k] * LT[k * 6 + j];
    for (int m = 0; m < j; ++m) LTinv[m * 6 + j] /= -LT[j * 6 + j];
  }
  mm_acc<6, 6, 6>(LTinv, D, T);                                  // matlib.cpp:220
  mm_nt_acc<6, 6, 6>(T, LTinv, Ainv);                            // matlib.cpp:221
}

// mean += inv(lambda) * eta                                     (bafuncs.cpp:3-15)
inline void inf2mean6(const float* eta, const float* lam, float* mean) {
  float sigma[36] = {};
  inv6(lam, sigma);
  mm_acc<6, 6, 1>(sigma, eta, mean);
}
inline void inf2mean3(const float* eta, const float* lam, float* mean) {
  float sigma[9] = {};
  inv3(lam, sigma);
  mm_acc<3, 3, 1>(sigma, eta, mean);
}

inline void hat3(const float* v, float* H) {                     // bafuncs.cpp:20-28
  H[1] = -v[2];
  H[2] = v[1];
  H[3] = v[2];
  H[5] = -v[0];
  H[6] = -v[1];
  H[7] = v[0];
}

// Rodrigues formula; R comes in zeroed                          (bafuncs.cpp:32-55)
inline void so3exp(const float* v, float* R) {
  R[0] = 1.f;
  R[4] = 1.f;
  R[8] = 1.f;
  const float theta = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (theta > 1e-6f) {
    const float s = std::sin(theta), c = std::cos(theta);
    float H[9] = {}, H2[9] = {};
    hat3(v, H);
    mm_acc<3, 3, 3>(H, H, H2);
    for (int i = 0; i < 9; ++i) {
      R[i] += (s / theta) * H[i];
      R[i] += ((1 - c) / (theta * theta)) * H2[i];
    }
  }
}

// 4x4 world->camera transform, last row all zero                (bafuncs.cpp:59-80)
inline void transf_w2c(const float* x, float* T) {
  T[15] = 0.f;
  T[3] = x[0];
  T[7] = x[1];
  T[11] = x[2];
  float R[9] = {};
  so3exp(x + 3, R);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j];
}

// Pinhole projection; K is the row-major 3x3                    (bafuncs.cpp:83-103)
inline void hfunc(const float* cam, const float* lmk, const float* K, float* hx) {
  float T[16] = {}, yh[4] = {lmk[0], lmk[1], lmk[2], 1.0f}, y[4] = {};
  transf_w2c(cam, T);
  mm_acc<4, 4, 1>(T, yh, y);
  hx[0] = K[0] * (y[0] / y[2]) + K[2];
  hx[1] = K[4] * (y[1] / y[2]) + K[5];
}

// Analytic Jacobians; Jkf[2x6], Jlmk[2x3] come in zeroed        (bafuncs.cpp:107-213)
inline void jac(const float* cam, const float* lmk, const float* K, float* Jkf, float* Jlmk) {
  float T[16] = {}, R[9];
  transf_w2c(cam, T);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = T[i * 4 + j];
  float yh[4] = {lmk[0], lmk[1], lmk[2], 1.0f}, y[4] = {};
  mm_acc<4, 4, 1>(T, yh, y);

  float Jp[6] = {};
  Jp[0] = K[0] / y[2];
  Jp[2] = -(K[0] * y[0]) / (y[2] * y[2]);
  Jp[4] = K[4] / y[2];
  Jp[5] = -(K[4] * y[1]) / (y[2] * y[2]);

  mm_acc<2, 3, 3>(Jp, R, Jlmk);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) Jkf[i * 6 + j] = Jp[i * 3 + j];

  const float* v = cam + 3;
  float vh[9] = {}, ph[9] = {}, vv[9], RtI[9] = {}, Rph[9] = {}, num[9] = {}, dR[9] = {}, Jr[6] = {};
  hat3(v, vh);
  hat3(lmk, ph);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) vv[i * 3 + j] = v[i] * v[j];
  for (int i = 0; i < 3; ++i) {
    RtI[i * 3 + i] = -1.f;
    for (int j = 0; j < 3; ++j) RtI[i * 3 + j] += R[j * 3 + i];
  }
  mm_acc<3, 3, 3>(R, ph, Rph);
  mm_acc<3, 3, 3>(RtI, vh, num);
  for (int i = 0; i < 9; ++i) num[i] += vv[i];
  float den = 0;
  for (int i = 0; i < 3; ++i) den += v[i] * v[i];
  mm_acc<3, 3, 3>(Rph, num, dR);
  for (int i = 0; i < 9; ++i) dR[i] = -dR[i] / den;
  mm_acc<2, 3, 3>(Jp, dR, Jr);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) Jkf[i * 6 + j + 3] = Jr[i * 3 + j];
}

// Shared body of RelineariseFactorVertex (gbp_codelets.cpp:90-168) and of the
// in-loop relinearisation of PrepMessageVertex (:285-373): ACCUMULATES J^T J
// and J^T(...) into the factor blocks, then divides by the Huber-modified
// variance.  The caller zeroes (or not -- quirk Q1) the blocks beforehand.
inline void linearise_accumulate(const float* z, float var, const float* K, const float* x_kf,
                                 const float* x_lmk, const Hyper& hp, float* f_eta, float* cc,
                                 float* cl, float* lc, float* ll, uint32_t* robust) {
  float Jkf[12] = {}, Jl[6] = {};
  jac(x_kf, x_lmk, K, Jkf, Jl);
  mm_tn_acc<2, 6, 6>(Jkf, Jkf, cc);
  mm_tn_acc<2, 3, 3>(Jl, Jl, ll);
  mm_tn_acc<2, 6, 3>(Jkf, Jl, cl);

  float hx[2] = {};
  hfunc(x_kf, x_lmk, K, hx);

  float eb[2] = {}, x0[9], J[18] = {};
  for (int i = 0; i < 6; ++i) x0[i] = x_kf[i];
  for (int i = 0; i < 3; ++i) x0[i + 6] = x_lmk[i];
  for (int i = 0; i < 2; ++i) {
    for (int j = 0; j < 6; ++j) J[i * 9 + j] = Jkf[i * 6 + j];
    for (int j = 0; j < 3; ++j) J[i * 9 + j + 6] = Jl[i * 3 + j];
  }
  mm_acc<2, 9, 1>(J, x0, eb);
  for (int i = 0; i < 2; ++i) eb[i] = eb[i] + z[i];
  for (int i = 0; i < 2; ++i) eb[i] = eb[i] - hx[i];
  mm_tn_acc<2, 9, 1>(J, eb, f_eta);

  // Huber; note the double-precision sub-expression (0.5 is a double literal).
  const float err = std::sqrt((hx[0] - z[0]) * (hx[0] - z[0]) + (hx[1] - z[1]) * (hx[1] - z[1]));
  float mvar = var;
  if (err > hp.Nstds * std::sqrt(var)) {
    *robust = 1;
    mvar = var * err * err / (2 * (hp.Nstds * std::sqrt(var) * err - 0.5 * hp.Nstds * hp.Nstds * var));
  } else {
    *robust = 0;
  }
  for (int i = 0; i < 36; ++i) cc[i] /= mvar;
  for (int i = 0; i < 9; ++i) ll[i] /= mvar;
  for (int i = 0; i < 18; ++i) cl[i] /= mvar;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 6; ++j) lc[i * 6 + j] = cl[j * 3 + i];
  for (int i = 0; i < 9; ++i) f_eta[i] /= mvar;
}

// RelineariseFactorVertex::compute                         (gbp_codelets.cpp:38-171)
// f_lam is the block-packed [cc 36 | cl 18 | lc 18 | ll 9] of ba/ba.cpp:93-96.
inline void relinearise_factor(const float* z, float var, const float* K, const float* kf_eta,
                               const float* kf_lam, const float* lmk_eta, const float* lmk_lam,
                               const Hyper& hp, float* f_eta, float* f_lam, uint32_t* robust) {
  for (int i = 0; i < 9; ++i) f_eta[i] = 0;
  for (int i = 0; i < 81; ++i) f_lam[i] = 0;
  float x_kf[6] = {}, x_l[3] = {};
  inf2mean6(kf_eta, kf_lam, x_kf);
  inf2mean3(lmk_eta, lmk_lam, x_l);
  linearise_accumulate(z, var, K, x_kf, x_l, hp, f_eta, f_lam, f_lam + 36, f_lam + 54, f_lam + 72,
                       robust);
}

// PrepMessageVertex::compute                               (gbp_codelets.cpp:241-378)
inline void prep_message(uint32_t active, float* damping, int32_t* damping_count, uint32_t* robust,
                         const float* z, float var, const float* K, const float* kf_eta,
                         const float* kf_lam, const float* lmk_eta, const float* lmk_lam,
                         const float* oldmu, float* mu, float* dmu, const Hyper& hp, float* f_eta,
                         float* f_lam) {
  if (active != 1) return;
  if (0 == *damping_count) *damping = hp.maxeta_damping;
  *damping_count += 1;
  float x_kf[6] = {}, x_l[3] = {};
  inf2mean6(kf_eta, kf_lam, x_kf);
  inf2mean3(lmk_eta, lmk_lam, x_l);
  *dmu = 0.0f;
  for (int i = 0; i < 6; ++i) {
    *dmu += (oldmu[i] - x_kf[i]) * (oldmu[i] - x_kf[i]);
    mu[i] = x_kf[i];
  }
  for (int i = 0; i < 3; ++i) {
    *dmu += (oldmu[i + 6] - x_l[i]) * (oldmu[i + 6] - x_l[i]);
    mu[i + 6] = x_l[i];
  }
  *dmu = std::sqrt(*dmu);
  if ((*dmu < hp.dmu_threshold) && (*damping_count > hp.min_linear_iters - hp.num_undamped_iters)) {
    *damping = 0.0f;
    *damping_count = -hp.num_undamped_iters;
    // Q1: no zeroing here -> the new linearisation is ADDED to the old blocks.
    linearise_accumulate(z, var, K, x_kf, x_l, hp, f_eta, f_lam, f_lam + 36, f_lam + 54,
                         f_lam + 72, robust);
  }
}

// ComputeCamMessageEtaVertex::compute                      (gbp_codelets.cpp:411-471)
inline void cam_message_eta(uint32_t active, float damping, const float* f_eta, const float* f_lam,
                            const float* lmk_b_eta, const float* lmk_b_lam, const float* p_lmk_eta,
                            const float* p_lmk_lam, const float* p_cam_eta, float* out) {
  if (active != 1) {
    for (int i = 0; i < 6; ++i) out[i] = 0.0f;
    return;
  }
  const float *eta_c = f_eta, *eta_l = f_eta + 6, *cl = f_lam + 36, *ll = f_lam + 72;
  float Ld[9], Li[9] = {}, P[18] = {}, ed[3], es[6] = {};
  for (int i = 0; i < 9; ++i) Ld[i] = ll[i] + lmk_b_lam[i];
  for (int i = 0; i < 9; ++i) Ld[i] = Ld[i] - p_lmk_lam[i];
  inv3(Ld, Li);
  mm_acc<6, 3, 3>(cl, Li, P);
  for (int i = 0; i < 3; ++i) ed[i] = eta_l[i] + lmk_b_eta[i];
  for (int i = 0; i < 3; ++i) ed[i] = ed[i] - p_lmk_eta[i];
  mm_acc<6, 3, 1>(P, ed, es);
  for (int i = 0; i < 6; ++i) {
    const float h = eta_c[i] - es[i];
    out[i] = h * (1 - damping) + p_cam_eta[i] * damping;
  }
}

// ComputeLmkMessageEtaVertex::compute                      (gbp_codelets.cpp:503-562)
inline void lmk_message_eta(uint32_t active, float damping, const float* f_eta, const float* f_lam,
                            const float* cam_b_eta, const float* cam_b_lam, const float* p_cam_eta,
                            const float* p_cam_lam, const float* p_lmk_eta, float* out) {
  if (active != 1) {
    for (int i = 0; i < 3; ++i) out[i] = 0.0f;
    return;
  }
  const float *eta_c = f_eta, *eta_l = f_eta + 6, *cc = f_lam, *lc = f_lam + 54;
  float Ld[36], Li[36] = {}, P[18] = {}, ed[6], es[3] = {};
  for (int i = 0; i < 36; ++i) Ld[i] = cc[i] + cam_b_lam[i];
  for (int i = 0; i < 36; ++i) Ld[i] = Ld[i] - p_cam_lam[i];
  inv6(Ld, Li);
  mm_acc<3, 6, 6>(lc, Li, P);
  for (int i = 0; i < 6; ++i) ed[i] = eta_c[i] + cam_b_eta[i];
  for (int i = 0; i < 6; ++i) ed[i] = ed[i] - p_cam_eta[i];
  mm_acc<3, 6, 1>(P, ed, es);
  for (int i = 0; i < 3; ++i) {
    const float h = eta_l[i] - es[i];
    out[i] = h * (1 - damping) + p_lmk_eta[i] * damping;
  }
}

// ComputeCamMessageLambdaVertex::compute                   (gbp_codelets.cpp:592-637)
inline void cam_message_lambda(uint32_t active, const float* f_lam, const float* lmk_b_lam,
                               const float* p_lmk_lam, float* out) {
  if (active != 1) {
    for (int i = 0; i < 36; ++i) out[i] = 0.0f;
    return;
  }
  const float *cc = f_lam, *cl = f_lam + 36, *lc = f_lam + 54, *ll = f_lam + 72;
  float Ld[9], Li[9] = {}, P[18] = {}, S[36] = {};
  for (int i = 0; i < 9; ++i) Ld[i] = ll[i] + lmk_b_lam[i];
  for (int i = 0; i < 9; ++i) Ld[i] = Ld[i] - p_lmk_lam[i];
  inv3(Ld, Li);
  mm_acc<6, 3, 3>(cl, Li, P);
  mm_acc<6, 3, 6>(P, lc, S);
  for (int i = 0; i < 36; ++i) out[i] = cc[i] - S[i];
}

// ComputeLmkMessageLambdaVertex::compute                   (gbp_codelets.cpp:664-709)
inline void lmk_message_lambda(uint32_t active, const float* f_lam, const float* cam_b_lam,
                               const float* p_cam_lam, float* out) {
  if (active != 1) {
    for (int i = 0; i < 9; ++i) out[i] = 0.0f;
    return;
  }
  const float *cc = f_lam, *cl = f_lam + 36, *lc = f_lam + 54, *ll = f_lam + 72;
  float Ld[36], Li[36] = {}, P[18] = {}, S[9] = {};
  for (int i = 0; i < 36; ++i) Ld[i] = cc[i] + cam_b_lam[i];
  for (int i = 0; i < 36; ++i) Ld[i] = Ld[i] - p_cam_lam[i];
  inv6(Ld, Li);
  mm_acc<3, 6, 6>(lc, Li, P);
  mm_acc<3, 6, 3>(P, cl, S);
  for (int i = 0; i < 9; ++i) out[i] = ll[i] - S[i];
}

// WeakenPriorVertex::compute                               (gbp_codelets.cpp:184-196)
inline void weaken_prior(float scaling, uint32_t* flag, float* eta, int n_eta, float* lam,
                         int n_lam) {
  if (*flag >= 1 && *flag <= 5) {   // quirk Q5: only flags 1..5 act
    *flag -= 1;
    for (int i = 0; i < n_eta; ++i) eta[i] *= scaling;
    for (int i = 0; i < n_lam; ++i) lam[i] *= scaling;
  }
}

}  // namespace gbp_restated
