// Device-side fp32 helpers of the GBP hot path (sm_100a).
//
// These are the __device__ counterparts of the reference's header-style math
// (ba/matlib.cpp, ba/bafuncs.cpp).  Two rules shape them:
//  (1) ORDER-FAITHFUL: every result is produced by the same sequence of IEEE
//      fp32 operations as the reference (whose matMul accumulates term by
//      term, matlib.cpp:54), written with __fmul_rn/__fadd_rn/... so nvcc can
//      never contract them into FMAs.  GBP is chaotic at rounding level
//      (SURVEY.md fact 4); keeping the op order makes a sweep bit-comparable
//      with the CPU oracle instead of merely "close".
//  (2) ZERO-SKIPPING: terms the reference multiplies by structural zeros
//      (D is diagonal, L^-T is unit upper triangular, hat matrices, ...) are
//      dropped.  x + (+-0) == x, so results are unchanged while the 6x6
//      inverse shrinks from ~1.2 kflop to ~0.4 kflop.
// Everything is fully unrolled on compile-time indices so the small matrices
// live in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace gbp {

#define GBP_DEV __device__ __forceinline__

#ifndef GBP_FAST_MATH
GBP_DEV float fm(float a, float b) { return __fmul_rn(a, b); }
GBP_DEV float fa(float a, float b) { return __fadd_rn(a, b); }
GBP_DEV float fs(float a, float b) { return __fsub_rn(a, b); }
GBP_DEV float fd(float a, float b) { return __fdiv_rn(a, b); }
#else
// gbp_fast.cu (opt-in, gbp_opts.fast_math): plain operators, which nvcc may contract into FMAs; the division stays IEEE
GBP_DEV float fm(float a, float b) { return a * b; }
GBP_DEV float fa(float a, float b) { return a + b; }
GBP_DEV float fs(float a, float b) { return a - b; }
GBP_DEV float fd(float a, float b) { return __fdiv_rn(a, b); }
#endif

// index of (i,j), i>=j, in a row-major packed lower triangle
__host__ __device__ constexpr int lt(int i, int j) { return i * (i + 1) / 2 + j; }

struct Hyper {  // ba/gbp_codelets.cpp:11-16
  float maxeta_damping;
  int num_undamped_iters;
  float dmu_threshold;
  int min_linear_iters;
  float Nstds;
};

// inv3x3 (matlib.cpp:143-161): closed form, nine divisions by det.
GBP_DEV void inv3(const float (&M)[9], float (&R)[9]) {
  const float c0 = fs(fm(M[4], M[8]), fm(M[7], M[5]));
  const float c1 = fs(fm(M[3], M[8]), fm(M[5], M[6]));
  const float c2 = fs(fm(M[3], M[7]), fm(M[4], M[6]));
  const float det = fa(fs(fm(M[0], c0), fm(M[1], c1)), fm(M[2], c2));
  R[0] = fd(c0, det);
  R[1] = fd(fs(fm(M[2], M[7]), fm(M[1], M[8])), det);
  R[2] = fd(fs(fm(M[1], M[5]), fm(M[2], M[4])), det);
  R[3] = fd(fs(fm(M[5], M[6]), fm(M[3], M[8])), det);
  R[4] = fd(fs(fm(M[0], M[8]), fm(M[2], M[6])), det);
  R[5] = fd(fs(fm(M[3], M[2]), fm(M[0], M[5])), det);
  R[6] = fd(fs(fm(M[3], M[7]), fm(M[6], M[4])), det);
  R[7] = fd(fs(fm(M[6], M[1]), fm(M[0], M[7])), det);
  R[8] = fd(fs(fm(M[0], M[4]), fm(M[3], M[1])), det);
}

// Compile-time loop: f(integral_constant<int, I>) for I in [B, E).  Used where every index
// must be a constant expression so the small matrices are scalarised into registers (nvcc
// gives up on `#pragma unroll` for the deepest triangular nests and falls back to a
// local-memory array, which costs long-scoreboard stalls in the hot loop).
template <int B, int E, class F>
GBP_DEV void static_for(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}

// inv6x6 (matlib.cpp:163-222): un-pivoted LDL^T of the LOWER triangle,
// explicit inverse of the unit upper factor, then L^-T D^-1 L^-1.
// A: packed lower triangle (21).  Ai: full 6x6 row-major (not bitwise symmetric).
GBP_DEV void inv6(const float (&A)[21], float (&Ai)[36]) {
  float D[6], rD[6];
  float U[36];  // U[j*6+i], j<i : the reference's LT(j,i)
  static_for<0, 6>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    float d = A[lt(j, j)];
    static_for<0, j>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      d = fs(d, fm(fm(U[k * 6 + j], U[k * 6 + j]), D[k]));
    });
    D[j] = d;
    const float r = fd(1.0f, d);
    rD[j] = r;
    static_for<j + 1, 6>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      float v = fm(r, A[lt(i, j)]);
      static_for<0, j>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        v = fs(v, fm(fm(fm(r, U[k * 6 + i]), U[k * 6 + j]), D[k]));
      });
      U[j * 6 + i] = v;
    });
  });
  // W = U^-1 (unit upper triangular), inv_uppertriang (matlib.cpp:163-178)
  float W[36];
  static_for<1, 6>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    static_for<0, j>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      float acc = U[i * 6 + j];  // 0 + 1*U(i,j)
      static_for<i + 1, j>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        acc = fa(acc, fm(W[i * 6 + k], U[k * 6 + j]));
      });
      W[i * 6 + j] = -acc;  // /= -1
    });
  });
  // T = W * D^-1 (upper triangular), Ai = T * W^T
  float T[36];
  static_for<0, 6>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    T[i * 6 + i] = rD[i];
    static_for<i + 1, 6>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      T[i * 6 + k] = fm(W[i * 6 + k], rD[k]);
    });
  });
  static_for<0, 6>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    static_for<0, 6>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      constexpr int k0 = (i > j) ? i : j;
      // first non-zero term: T(i,k0) * W(j,k0), with W(j,j) == 1
      float acc = (k0 == j) ? T[i * 6 + j] : fm(T[i * 6 + k0], W[j * 6 + k0]);
      static_for<k0 + 1, 6>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        acc = fa(acc, fm(T[i * 6 + k], W[j * 6 + k]));
      });
      Ai[i * 6 + j] = acc;
    });
  });
}

// mean = inv(lambda) * eta  (bafuncs.cpp:3-15), lambda given as packed lower triangle.
GBP_DEV void inf2mean6(const float (&eta)[6], const float (&lamL)[21], float (&mean)[6]) {
  float S[36];
  inv6(lamL, S);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float acc = fm(S[i * 6], eta[0]);
#pragma unroll
    for (int k = 1; k < 6; ++k) acc = fa(acc, fm(S[i * 6 + k], eta[k]));
    mean[i] = acc;
  }
}

GBP_DEV void inf2mean3(const float (&eta)[3], const float (&lam)[9], float (&mean)[3]) {
  float S[9];
  inv3(lam, S);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    mean[i] = fa(fa(fm(S[i * 3], eta[0]), fm(S[i * 3 + 1], eta[1])), fm(S[i * 3 + 2], eta[2]));
}

// Rodrigues (bafuncs.cpp:32-55).  sin/cos are evaluated in double and rounded
// once, which reproduces a correctly-rounded host sinf/cosf.
GBP_DEV void so3exp(const float (&v)[3], float (&R)[9]) {
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0f : 0.0f;
  const float theta = __fsqrt_rn(fa(fa(fm(v[0], v[0]), fm(v[1], v[1])), fm(v[2], v[2])));
  if (theta > 1e-6f) {
    const float s = (float)sin((double)theta), c = (float)cos((double)theta);
    const float H[9] = {0.f, -v[2], v[1], v[2], 0.f, -v[0], -v[1], v[0], 0.f};
    float H2[9];
    // H*H, skipping the structurally zero factors (term order preserved)
    H2[0] = fa(fm(H[1], H[3]), fm(H[2], H[6]));
    H2[1] = fm(H[2], H[7]);
    H2[2] = fm(H[1], H[5]);
    H2[3] = fm(H[5], H[6]);
    H2[4] = fa(fm(H[3], H[1]), fm(H[5], H[7]));
    H2[5] = fm(H[3], H[2]);
    H2[6] = fm(H[7], H[3]);
    H2[7] = fm(H[6], H[1]);
    H2[8] = fa(fm(H[6], H[2]), fm(H[7], H[5]));
    const float a = fd(s, theta), b = fd(fs(1.0f, c), fm(theta, theta));
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      R[i] = fa(R[i], fm(a, H[i]));
      R[i] = fa(R[i], fm(b, H2[i]));
    }
  }
}

// A(3x3) * hat(h): two non-zero terms per entry, reference term order.
GBP_DEV void mul_hat(const float (&A)[9], const float (&h)[3], float (&O)[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    O[i * 3 + 0] = fa(fm(A[i * 3 + 1], h[2]), fm(A[i * 3 + 2], -h[1]));
    O[i * 3 + 1] = fa(fm(A[i * 3 + 0], -h[2]), fm(A[i * 3 + 2], h[0]));
    O[i * 3 + 2] = fa(fm(A[i * 3 + 0], h[1]), fm(A[i * 3 + 1], -h[0]));
  }
}

// One (re)linearisation of a reprojection factor around (x_kf, x_l):
// Jac (bafuncs.cpp:107-213), hfunc (:83-103), J^T J, J^T(J x0 + z - h) and the
// Huber reweighting (gbp_codelets.cpp:95-168 == :299-373).  The J^T J / J^T r
// terms are ADDED onto the incoming blocks (zero them first for
// RelineariseFactorVertex; leave them for the in-loop relinearisation, quirk Q1).
// Layout of the blocks: eta[9], ll[9], cl[18] (6x3), cc[36]; lc is cl^T.
// Camera-only part of Jac (bafuncs.cpp:107-213): R = so3exp(w), num = (R^T - I)[w]x + w w^T and
// den = |w|^2 of dRp/dw = -R [p]x num / den.  The reference recomputes them per factor; here they
// are formed once per camera when its mean changes (k_update_vars) -- same operations, same values.
GBP_DEV void cam_lin_consts(const float (&w)[3], float (&R)[9], float (&num)[9], float& den) {
  so3exp(w, R);
  float RtI[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) RtI[i * 3 + j] = (i == j) ? fa(-1.0f, R[i * 3 + i]) : R[j * 3 + i];
  mul_hat(RtI, w, num);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) num[i * 3 + j] = fa(num[i * 3 + j], fm(w[i], w[j]));
  den = fa(fa(fm(w[0], w[0]), fm(w[1], w[1])), fm(w[2], w[2]));
}

// Projection of the landmark into the camera and its Jacobians at (x_kf, x_l): hfunc (bafuncs.cpp:83-103) and Jac
// (bafuncs.cpp:107-213) given the camera-only constants of cam_lin_consts.  Jk: 2x6 row-major (d h / d pose),
// Jl: 2x3 (d h / d landmark), (h0, h1) = h(x0).
GBP_DEV void project_jac(const float (&K)[4] /* fx fy cx cy */, const float (&x_kf)[6], const float (&x_l)[3], const float (&R)[9],
                         const float (&num)[9], const float den, float (&Jk)[12], float (&Jl)[6], float& h0, float& h1) {
  const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
  float y[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    y[i] = fa(fa(fa(fm(R[i * 3], x_l[0]), fm(R[i * 3 + 1], x_l[1])), fm(R[i * 3 + 2], x_l[2])), x_kf[i]);
  // J_proj = [a 0 b; 0 c d]
  const float ja = fd(fx, y[2]);
  const float jb = fd(-fm(fx, y[0]), fm(y[2], y[2]));
  const float jc = fd(fy, y[2]);
  const float jd = fd(-fm(fy, y[1]), fm(y[2], y[2]));
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    Jl[j] = fa(fm(ja, R[j]), fm(jb, R[6 + j]));
    Jl[3 + j] = fa(fm(jc, R[3 + j]), fm(jd, R[6 + j]));
  }
  Jk[0] = ja; Jk[1] = 0.f; Jk[2] = jb;
  Jk[6] = 0.f; Jk[7] = jc; Jk[8] = jd;
  // rotation part: dRp/dw = -R [p]x (w w^T + (R^T - I)[w]x) / |w|^2
  float Rph[9], dR[9];
  mul_hat(R, x_l, Rph);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float acc = fa(fa(fm(Rph[i * 3], num[j]), fm(Rph[i * 3 + 1], num[3 + j])), fm(Rph[i * 3 + 2], num[6 + j]));
      dR[i * 3 + j] = fd(-acc, den);
    }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    Jk[3 + j] = fa(fm(ja, dR[j]), fm(jb, dR[6 + j]));
    Jk[9 + j] = fa(fm(jc, dR[3 + j]), fm(jd, dR[6 + j]));
  }
  h0 = fa(fm(fx, fd(y[0], y[2])), cx);
  h1 = fa(fm(fy, fd(y[1], y[2])), cy);
}

GBP_DEV uint32_t linearise_accumulate(const float z0, const float z1, const float var,
                                      const float (&K)[4] /* fx fy cx cy */,
                                      const float (&x_kf)[6], const float (&x_l)[3],
                                      const float (&R)[9], const float (&num)[9], const float den,
                                      const float Nstds, float (&eta)[9], float (&ll)[9],
                                      float (&cl)[18], float (&cc)[36]) {
  float Jk[12], Jl[6], h0, h1;
  project_jac(K, x_kf, x_l, R, num, den, Jk, Jl, h0, h1);
  // J^T J accumulated onto the blocks (matMul(.., true, false), matlib.cpp:60-68)
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) cc[i * 6 + j] = fa(fa(cc[i * 6 + j], fm(Jk[i], Jk[j])), fm(Jk[6 + i], Jk[6 + j]));
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) ll[i * 3 + j] = fa(fa(ll[i * 3 + j], fm(Jl[i], Jl[j])), fm(Jl[3 + i], Jl[3 + j]));
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) cl[i * 3 + j] = fa(fa(cl[i * 3 + j], fm(Jk[i], Jl[j])), fm(Jk[6 + i], Jl[3 + j]));
  // eb = J x0 + z - h ; eta += J^T eb
  float eb[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float acc = fm(Jk[r * 6], x_kf[0]);
#pragma unroll
    for (int j = 1; j < 6; ++j) acc = fa(acc, fm(Jk[r * 6 + j], x_kf[j]));
#pragma unroll
    for (int j = 0; j < 3; ++j) acc = fa(acc, fm(Jl[r * 3 + j], x_l[j]));
    eb[r] = acc;
  }
  eb[0] = fs(fa(eb[0], z0), h0);
  eb[1] = fs(fa(eb[1], z1), h1);
#pragma unroll
  for (int i = 0; i < 6; ++i) eta[i] = fa(fa(eta[i], fm(Jk[i], eb[0])), fm(Jk[6 + i], eb[1]));
#pragma unroll
  for (int i = 0; i < 3; ++i) eta[6 + i] = fa(fa(eta[6 + i], fm(Jl[i], eb[0])), fm(Jl[3 + i], eb[1]));
  // Huber (gbp_codelets.cpp:135-141); the reference evaluates the denominator
  // in double because of the 0.5 literal.
  const float e0 = fs(h0, z0), e1 = fs(h1, z1);
  const float err = __fsqrt_rn(fa(fm(e0, e0), fm(e1, e1)));
  const float sd = __fsqrt_rn(var);
  float mvar = var;
  uint32_t robust = 0;
  if (err > fm(Nstds, sd)) {
    robust = 1;
    const float numer = fm(fm(var, err), err);
    const double den2 = __dmul_rn(2.0, __dsub_rn((double)fm(fm(Nstds, sd), err),
                                                 __dmul_rn(__dmul_rn(__dmul_rn(0.5, (double)Nstds), (double)Nstds), (double)var)));
    mvar = (float)__ddiv_rn((double)numer, den2);
  }
  // Division of the 54 distinct entries by the (Huber-adjusted) variance.  When the variance is a power of
  // two -- the default --reproj_meas_var 4 of every non-robust factor -- x / var and x * (1 / var) are
  // roundings of the same real number, hence the same bits for every x, and the multiplication costs one
  // instruction instead of ten.  Taken when all lanes of the warp that are here qualify.
  const uint32_t mb = __float_as_uint(mvar);
  const bool pow2 = (mb & 0x007fffffu) == 0u && mb >= 0x01000000u && mb <= 0x7e000000u;
  if (__all_sync(__activemask(), pow2)) {
    const float r = __uint_as_float(0x7f000000u - mb);  // exactly 1 / mvar
#pragma unroll
    for (int i = 0; i < 36; ++i) cc[i] = fm(cc[i], r);
#pragma unroll
    for (int i = 0; i < 9; ++i) ll[i] = fm(ll[i], r);
#pragma unroll
    for (int i = 0; i < 18; ++i) cl[i] = fm(cl[i], r);
#pragma unroll
    for (int i = 0; i < 9; ++i) eta[i] = fm(eta[i], r);
  } else {
#pragma unroll
    for (int i = 0; i < 36; ++i) cc[i] = fd(cc[i], mvar);
#pragma unroll
    for (int i = 0; i < 9; ++i) ll[i] = fd(ll[i], mvar);
#pragma unroll
    for (int i = 0; i < 18; ++i) cl[i] = fd(cl[i], mvar);
#pragma unroll
    for (int i = 0; i < 9; ++i) eta[i] = fd(eta[i], mvar);
  }
  return robust;
}

}  // namespace gbp
