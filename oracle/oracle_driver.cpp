// TEST INFRASTRUCTURE ONLY -- the CPU oracle for the GBP hot path.
//
// Restates the reference's Poplar programs (WRITE/LINEARISE/GBP/WEAKEN/READ/
// READ_PRIORS/NEW_KEYFRAME, ba/ba.cpp:860-916, ba/slam.cpp:913-928) as plain
// loops over the reference's own padded tensor layouts, and exposes them
// through a C ABI that mirrors include/gbp_cuda.h (prefix gbp_oracle_).
//
// Built twice (oracle/Makefile):
//   oracle/libgbp_oracle.so    codelet arithmetic = oracle/gbp_restated.hpp   ("port")
//   oracle/_ref/libgbp_ref.so  codelet arithmetic = the reference's own
//                              gbp_codelets.cpp compiled behind oracle/shim    ("reference")
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load either library.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/gbp_cuda.h"

// Transcendentals: both oracle builds bind sinf/cosf (the only libm calls of
// the hot path besides sqrt, bafuncs.cpp:39-40) to the definitions below --
// the double-precision result rounded once, i.e. a correctly rounded libm --
// instead of glibc's <1-ulp-but-not-correctly-rounded sinf/cosf.  The CUDA
// path evaluates them the same way, which makes CPU and GPU trajectories
// comparable bit for bit.  (Linked with -Wl,-Bsymbolic, oracle/Makefile.)
extern "C" {
__attribute__((noinline)) float sinf(float x) {
  volatile double d = std::sin((double)x);
  return (float)d;
}
__attribute__((noinline)) float cosf(float x) {
  volatile double d = std::cos((double)x);
  return (float)d;
}
// gcc fuses sinf(x)+cosf(x) into one sincosf call
__attribute__((noinline)) void sincosf(float x, float* s, float* c) {
  volatile double ds = std::sin((double)x), dc = std::cos((double)x);
  *s = (float)ds;
  *c = (float)dc;
}
}

#ifdef GBP_ORACLE_USE_REFERENCE
#include "ref_backend.hpp"
namespace be = gbp_ref_backend;
#define GBP_ORACLE_KIND "reference"
#else
#include "gbp_restated.hpp"
namespace be = gbp_restated;
#define GBP_ORACLE_KIND "port"
#endif

using gbp_restated::Hyper;

namespace {

struct Oracle {
  uint32_t C = 0, L = 0, E = 0, SK = 0, SL = 0;  // SK/SL = max degree + 1 (slot 0 = prior)
  std::vector<uint32_t> cam_ids, lmk_ids, slot_c, slot_l;  // slot = edges_c / edges_l of ba.cpp:267-279
  std::vector<float> z, var;
  float K[9];
  Hyper hp;
  int nthreads = 1;
  int reduce_order = 0;  // 0 = serial slot order; 1 = the CUDA path's tile order; 2 = its multi-GPU order (see update_beliefs)
  uint32_t shard_world = 1;
  std::vector<uint16_t> slot_rank;     // [L*SL] rank owning the camera of each landmark message slot (0xffff = none)
  std::vector<uint8_t> lmk_boundary;   // [L] observed from more than one rank
  // variable-side tensors (ba/ba.cpp:665-687)
  std::vector<float> cam_b_eta, cam_b_lam, lmk_b_eta, lmk_b_lam;
  std::vector<float> cam_scaling, lmk_scaling;
  std::vector<uint32_t> cam_wflag, lmk_wflag;
  std::vector<float> cam_m_eta, cam_m_lam, lmk_m_eta, lmk_m_lam;      // messages
  std::vector<float> pcam_m_eta, pcam_m_lam, plmk_m_eta, plmk_m_lam;  // previous messages
  // factor-side tensors (ba/ba.cpp:759-775)
  std::vector<float> damping, mu, oldmu, dmu, f_eta, f_lam;
  std::vector<int32_t> damping_count;
  std::vector<uint32_t> active, robust;
  double last_ms = 0;
};

thread_local std::string g_err;

inline size_t cme(const Oracle& o, uint32_t c, uint32_t slot) { return ((size_t)o.SK * c + slot) * 6; }
inline size_t cml(const Oracle& o, uint32_t c, uint32_t slot) { return ((size_t)o.SK * c + slot) * 36; }
inline size_t lme(const Oracle& o, uint32_t l, uint32_t slot) { return ((size_t)o.SL * l + slot) * 3; }
inline size_t lml(const Oracle& o, uint32_t l, uint32_t slot) { return ((size_t)o.SL * l + slot) * 9; }

// prog_ub: belief = sum over slots, fp32 (ba/ba.cpp:104-139).  popops::reduce does
// not specify its summation order; two realisations are provided:
//   reduce_order 0: serial slot order  ((prior + m1) + m2) + ...
//   reduce_order 1: the order of the CUDA path -- camera messages are summed serially
//     inside chunks of 32 consecutive slots (one warp's factors), and the chunk sums
//     are added to the prior in order (chunks are counted in whole 128-slot tiles).
//     Landmarks are serial in both modes.
//   reduce_order 2: order 1 for cameras; a landmark observed from more than one rank of a
//     camera-range partition (gbp_oracle_set_shard_bounds) is summed the way the multi-GPU
//     path does: per-rank partial sums (serial, starting from +0), then
//     (0 + prior) + partial[0] + partial[1] + ...; all other landmarks stay serial.
float cam_chunk_sum(const Oracle& o, const std::vector<float>& m, uint32_t c, uint32_t d, uint32_t dim,
                    uint32_t chunk) {
  float acc = 0.f;
  for (uint32_t i = 0; i < 32; ++i) {
    const uint32_t slot = chunk * 32 + i + 1;  // slot 0 is the prior
    const float v = (slot < o.SK) ? m[((size_t)o.SK * c + slot) * dim + d] : 0.f;
    acc = (i == 0) ? v : acc + v;
  }
  return acc;
}

void update_beliefs(Oracle& o) {
  const uint32_t n_chunks = ((o.SK - 1 + 127) / 128) * 4;
#pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for (int64_t c = 0; c < (int64_t)o.C; ++c) {
    // slots beyond the camera's degree are zero, so summing max-degree many chunks is exact
    for (int d = 0; d < 6; ++d) {
      float s = 0.f;
      if (o.reduce_order == 0) {
        for (uint32_t k = 0; k < o.SK; ++k) s += o.cam_m_eta[cme(o, c, k) + d];
      } else {
        s += o.cam_m_eta[cme(o, c, 0) + d];
        for (uint32_t t = 0; t < n_chunks; ++t) s += cam_chunk_sum(o, o.cam_m_eta, c, d, 6, t);
      }
      o.cam_b_eta[c * 6 + d] = s;
    }
    for (int d = 0; d < 36; ++d) {
      float s = 0.f;
      if (o.reduce_order == 0) {
        for (uint32_t k = 0; k < o.SK; ++k) s += o.cam_m_lam[cml(o, c, k) + d];
      } else {
        s += o.cam_m_lam[cml(o, c, 0) + d];
        for (uint32_t t = 0; t < n_chunks; ++t) s += cam_chunk_sum(o, o.cam_m_lam, c, d, 36, t);
      }
      o.cam_b_lam[c * 36 + d] = s;
    }
  }
#pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for (int64_t l = 0; l < (int64_t)o.L; ++l) {
    if (o.reduce_order == 2 && o.lmk_boundary[l]) {
      const uint16_t* sr = &o.slot_rank[(size_t)l * o.SL];
      for (int d = 0; d < 12; ++d) {
        const bool is_eta = d < 3;
        const std::vector<float>& m = is_eta ? o.lmk_m_eta : o.lmk_m_lam;
        const size_t base = is_eta ? lme(o, l, 0) + d : lml(o, l, 0) + (d - 3);
        const size_t dim = is_eta ? 3 : 9;
        float s = 0.f;
        s += m[base];
        for (uint32_t r = 0; r < o.shard_world; ++r) {
          float part = 0.f;
          for (uint32_t k = 1; k < o.SL; ++k)
            if (sr[k] == r) part += m[base + (size_t)k * dim];
          s += part;
        }
        if (is_eta) o.lmk_b_eta[l * 3 + d] = s; else o.lmk_b_lam[l * 9 + (d - 3)] = s;
      }
      continue;
    }
    for (int d = 0; d < 3; ++d) {
      float s = 0.f;
      for (uint32_t k = 0; k < o.SL; ++k) s += o.lmk_m_eta[lme(o, l, k) + d];
      o.lmk_b_eta[l * 3 + d] = s;
    }
    for (int d = 0; d < 9; ++d) {
      float s = 0.f;
      for (uint32_t k = 0; k < o.SL; ++k) s += o.lmk_m_lam[lml(o, l, k) + d];
      o.lmk_b_lam[l * 9 + d] = s;
    }
  }
}

// cs_relinearise: every edge, no active check (ba/ba.cpp:68-97).
void relinearise_all(Oracle& o) {
#pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for (int64_t e = 0; e < (int64_t)o.E; ++e) {
    const uint32_t c = o.cam_ids[e], l = o.lmk_ids[e];
    be::relinearise_factor(&o.z[2 * e], o.var[e], o.K, &o.cam_b_eta[c * 6], &o.cam_b_lam[c * 36],
                           &o.lmk_b_eta[l * 3], &o.lmk_b_lam[l * 9], o.hp, &o.f_eta[9 * e],
                           &o.f_lam[81 * e], &o.robust[e]);
  }
}

// cs_compmess_prep (ba/ba.cpp:249,282-300).
void prep_all(Oracle& o) {
#pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for (int64_t e = 0; e < (int64_t)o.E; ++e) {
    const uint32_t c = o.cam_ids[e], l = o.lmk_ids[e];
    be::prep_message(o.active[e], &o.damping[e], &o.damping_count[e], &o.robust[e], &o.z[2 * e],
                     o.var[e], o.K, &o.cam_b_eta[c * 6], &o.cam_b_lam[c * 36], &o.lmk_b_eta[l * 3],
                     &o.lmk_b_lam[l * 9], &o.oldmu[9 * e], &o.mu[9 * e], &o.dmu[e], o.hp,
                     &o.f_eta[9 * e], &o.f_lam[81 * e]);
  }
}

// cs_computemessages: the four message vertices of every edge (ba/ba.cpp:252-366).
void messages_all(Oracle& o) {
#pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for (int64_t e = 0; e < (int64_t)o.E; ++e) {
    const uint32_t c = o.cam_ids[e], l = o.lmk_ids[e];
    const uint32_t sc = o.slot_c[e] + 1, sl = o.slot_l[e] + 1;
    const float* fe = &o.f_eta[9 * e];
    const float* fl = &o.f_lam[81 * e];
    be::cam_message_eta(o.active[e], o.damping[e], fe, fl, &o.lmk_b_eta[l * 3], &o.lmk_b_lam[l * 9],
                        &o.plmk_m_eta[lme(o, l, sl)], &o.plmk_m_lam[lml(o, l, sl)],
                        &o.pcam_m_eta[cme(o, c, sc)], &o.cam_m_eta[cme(o, c, sc)]);
    be::cam_message_lambda(o.active[e], fl, &o.lmk_b_lam[l * 9], &o.plmk_m_lam[lml(o, l, sl)],
                           &o.cam_m_lam[cml(o, c, sc)]);
    be::lmk_message_eta(o.active[e], o.damping[e], fe, fl, &o.cam_b_eta[c * 6], &o.cam_b_lam[c * 36],
                        &o.pcam_m_eta[cme(o, c, sc)], &o.pcam_m_lam[cml(o, c, sc)],
                        &o.plmk_m_eta[lme(o, l, sl)], &o.lmk_m_eta[lme(o, l, sl)]);
    be::lmk_message_lambda(o.active[e], fl, &o.cam_b_lam[c * 36], &o.pcam_m_lam[cml(o, c, sc)],
                           &o.lmk_m_lam[lml(o, l, sl)]);
  }
}

// cs_weaken_prior (ba/ba.cpp:165-182): acts on slot 0 of the message tensors.
void weaken_all(Oracle& o) {
  for (uint32_t c = 0; c < o.C; ++c)
    be::weaken_prior(o.cam_scaling[c], &o.cam_wflag[c], &o.cam_m_eta[cme(o, c, 0)], 6,
                     &o.cam_m_lam[cml(o, c, 0)], 36);
  for (uint32_t l = 0; l < o.L; ++l)
    be::weaken_prior(o.lmk_scaling[l], &o.lmk_wflag[l], &o.lmk_m_eta[lme(o, l, 0)], 3,
                     &o.lmk_m_lam[lml(o, l, 0)], 9);
}

void commit_messages(Oracle& o) {  // ba/ba.cpp:902-905
  o.pcam_m_eta = o.cam_m_eta;
  o.pcam_m_lam = o.cam_m_lam;
  o.plmk_m_eta = o.lmk_m_eta;
  o.plmk_m_lam = o.lmk_m_lam;
}

void gbp_sweep(Oracle& o) {  // gbp_iter_prog, ba/ba.cpp:895-905
  prep_all(o);
  o.oldmu = o.mu;
  messages_all(o);
  update_beliefs(o);
  commit_messages(o);
}

// ---- metric: eval_reprojection_error (ba/util.cpp:74-144) -----------------
// The reference inverts the full fp32 belief matrices with Eigen (absent
// here); this restatement solves the same systems in double by Gaussian
// elimination with partial pivoting and sums in double.
bool solve_n(int n, const double* A_in, const double* b_in, double* x) {
  double A[36], b[6];
  for (int i = 0; i < n * n; ++i) A[i] = A_in[i];
  for (int i = 0; i < n; ++i) b[i] = b_in[i];
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(A[i * n + k]) > std::fabs(A[p * n + k])) p = i;
    if (A[p * n + k] == 0.0) return false;
    if (p != k) {
      for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[p * n + j]);
      std::swap(b[k], b[p]);
    }
    for (int i = k + 1; i < n; ++i) {
      const double f = A[i * n + k] / A[k * n + k];
      for (int j = k; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
      b[i] -= f * b[k];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * x[j];
    x[i] = s / A[i * n + i];
  }
  return true;
}

void so3exp_d(const double* w, double* R) {  // ba/util.cpp:20-32
  const double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if (th < 1e-6) return;
  const double H[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  const double a = std::sin(th) / th, b = (1 - std::cos(th)) / (th * th);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double h2 = 0;
      for (int k = 0; k < 3; ++k) h2 += H[i * 3 + k] * H[k * 3 + j];
      R[i * 3 + j] += a * H[i * 3 + j] + b * h2;
    }
}

void eval_metrics(const Oracle& o, gbp_iter_stats* st) {
  std::vector<double> cmu((size_t)o.C * 6), cR((size_t)o.C * 9), lmu((size_t)o.L * 3);
#pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for (int64_t c = 0; c < (int64_t)o.C; ++c) {
    double A[36], b[6];
    for (int i = 0; i < 36; ++i) A[i] = o.cam_b_lam[c * 36 + i];
    for (int i = 0; i < 6; ++i) b[i] = o.cam_b_eta[c * 6 + i];
    solve_n(6, A, b, &cmu[c * 6]);
    so3exp_d(&cmu[c * 6 + 3], &cR[c * 9]);
  }
#pragma omp parallel for schedule(static) num_threads(o.nthreads)
  for (int64_t l = 0; l < (int64_t)o.L; ++l) {
    double A[9], b[3];
    for (int i = 0; i < 9; ++i) A[i] = o.lmk_b_lam[l * 9 + i];
    for (int i = 0; i < 3; ++i) b[i] = o.lmk_b_eta[l * 3 + i];
    solve_n(3, A, b, &lmu[l * 3]);
  }
  uint32_t n_active = 0, n_robust = 0, n_relins = 0;
  for (uint32_t e = 0; e < o.E; ++e) {
    n_active += o.active[e];
    n_robust += o.robust[e];
    if (o.damping_count[e] == -o.hp.num_undamped_iters) n_relins += 1;  // ba.cpp:1016-1020
  }
  double sum_norm = 0, sum_sq = 0;
  // Q7: the reference walks e in [0, n_active) -- active edges are a prefix
  // because the shipped files are camera-sorted (ba/util.cpp:95-99).
#pragma omp parallel for schedule(static) reduction(+ : sum_norm, sum_sq) num_threads(o.nthreads)
  for (int64_t e = 0; e < (int64_t)n_active; ++e) {
    const uint32_t c = o.cam_ids[e], l = o.lmk_ids[e];
    const double *R = &cR[c * 9], *t = &cmu[c * 6], *p = &lmu[l * 3];
    double y[3];
    for (int i = 0; i < 3; ++i) y[i] = R[i * 3] * p[0] + R[i * 3 + 1] * p[1] + R[i * 3 + 2] * p[2] + t[i];
    const double u = (o.K[0] * y[0] + o.K[1] * y[1] + o.K[2] * y[2]) / y[2];
    const double v = (o.K[3] * y[0] + o.K[4] * y[1] + o.K[5] * y[2]) / y[2];
    const double r0 = o.z[2 * e] - u, r1 = o.z[2 * e + 1] - v;
    const double sq = r0 * r0 + r1 * r1;
    sum_norm += std::sqrt(sq);
    sum_sq += 0.5 * sq;
  }
  st->reproj_mean = (float)(sum_norm / (double)n_active);
  st->cost = (float)sum_sq;
  st->n_relins = n_relins;
  st->n_robust = n_robust;
  st->n_active = n_active;
  st->reserved = 0;
}

struct TensorRef {
  void* ptr;
  size_t nbytes;
};

bool find_tensor(Oracle& o, const std::string& n, TensorRef* t) {
#define T_(name, vec)                                   \
  if (n == name) {                                      \
    t->ptr = (void*)(vec).data();                       \
    t->nbytes = (vec).size() * sizeof((vec)[0]);        \
    return true;                                        \
  }
  T_("cam_beliefs_eta", o.cam_b_eta)
  T_("cam_beliefs_lambda", o.cam_b_lam)
  T_("lmk_beliefs_eta", o.lmk_b_eta)
  T_("lmk_beliefs_lambda", o.lmk_b_lam)
  T_("cam_messages_eta", o.cam_m_eta)
  T_("cam_messages_lambda", o.cam_m_lam)
  T_("lmk_messages_eta", o.lmk_m_eta)
  T_("lmk_messages_lambda", o.lmk_m_lam)
  T_("pcam_messages_eta", o.pcam_m_eta)
  T_("pcam_messages_lambda", o.pcam_m_lam)
  T_("plmk_messages_eta", o.plmk_m_eta)
  T_("plmk_messages_lambda", o.plmk_m_lam)
  T_("factor_potentials_eta", o.f_eta)
  T_("factor_potentials_lambda", o.f_lam)
  T_("damping", o.damping)
  T_("damping_count", o.damping_count)
  T_("mu", o.mu)
  T_("oldmu", o.oldmu)
  T_("dmu", o.dmu)
  T_("active_flag", o.active)
  T_("robust_flag", o.robust)
  T_("measurements", o.z)
  T_("meas_variances", o.var)
  T_("cam_scaling", o.cam_scaling)
  T_("lmk_scaling", o.lmk_scaling)
  T_("cam_weaken_flag", o.cam_wflag)
  T_("lmk_weaken_flag", o.lmk_wflag)
#undef T_
  return false;
}

}  // namespace

extern "C" {

const char* gbp_oracle_kind(void) { return GBP_ORACLE_KIND; }
const char* gbp_oracle_last_error(void) { return g_err.c_str(); }

int gbp_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int gbp_oracle_set_reduce_order(void* h, int mode) {
  if (!h || mode < 0 || mode > 1) return GBP_ERR_ARG;
  ((Oracle*)h)->reduce_order = mode;
  return GBP_OK;
}

// Camera-range partition of the multi-GPU path: bounds[world+1]; selects reduce_order 2.
int gbp_oracle_set_shard_bounds(void* h, const uint32_t* bounds, uint32_t world) {
  if (!h || !bounds || world == 0) return GBP_ERR_ARG;
  Oracle& o = *(Oracle*)h;
  o.shard_world = world;
  o.slot_rank.assign((size_t)o.L * o.SL, 0xffff);
  o.lmk_boundary.assign(o.L, 0);
  std::vector<uint16_t> first(o.L, 0xffff);
  for (uint32_t e = 0; e < o.E; ++e) {
    uint32_t r = 0;
    while (r + 1 < world && o.cam_ids[e] >= bounds[r + 1]) ++r;
    const uint32_t l = o.lmk_ids[e];
    o.slot_rank[(size_t)l * o.SL + o.slot_l[e] + 1] = (uint16_t)r;
    if (first[l] == 0xffff) first[l] = (uint16_t)r;
    else if (first[l] != r) o.lmk_boundary[l] = 1;
  }
  o.reduce_order = 2;
  return GBP_OK;
}

int gbp_oracle_set_threads(void* h, int n) {
  if (!h || n < 1) return GBP_ERR_ARG;
#ifndef _OPENMP
  n = 1;
#endif
  ((Oracle*)h)->nthreads = n;
  return GBP_OK;
}

// Graph build + WRITE_PROG + LINEARISE_PROG (ba/ba.cpp:659-986).
int gbp_oracle_init(const gbp_problem* p, const gbp_opts* opts, void** out) {
  if (!p || !out || !p->cam_ids || !p->lmk_ids || !p->measurements || !p->meas_variances) {
    g_err = "null problem field";
    return GBP_ERR_ARG;
  }
  Oracle* o = new Oracle();
  o->C = p->n_keyframes;
  o->L = p->n_points;
  o->E = p->n_edges;
  if (opts) {
    o->hp.maxeta_damping = opts->maxeta_damping;
    o->hp.num_undamped_iters = opts->num_undamped_iters;
    o->hp.dmu_threshold = opts->dmu_threshold;
    o->hp.min_linear_iters = opts->min_linear_iters;
    o->hp.Nstds = opts->Nstds;
  }
#ifdef GBP_ORACLE_USE_REFERENCE
  be::set_hyper(o->hp);
#endif
  const uint32_t C = o->C, L = o->L, E = o->E;
  o->cam_ids.assign(p->cam_ids, p->cam_ids + E);
  o->lmk_ids.assign(p->lmk_ids, p->lmk_ids + E);
  o->z.assign(p->measurements, p->measurements + 2 * (size_t)E);
  o->var.assign(p->meas_variances, p->meas_variances + E);
  std::memcpy(o->K, p->K, sizeof(o->K));
  // slot of each edge at its two variables = number of earlier edges touching
  // the same variable (ba/ba.cpp:267-279), computed in O(E).
  std::vector<uint32_t> dc(C, 0), dl(L, 0);
  o->slot_c.resize(E);
  o->slot_l.resize(E);
  for (uint32_t e = 0; e < E; ++e) {
    if (o->cam_ids[e] >= C || o->lmk_ids[e] >= L) {
      g_err = "edge index out of range";
      delete o;
      return GBP_ERR_ARG;
    }
    o->slot_c[e] = dc[o->cam_ids[e]]++;
    o->slot_l[e] = dl[o->lmk_ids[e]]++;
  }
  o->SK = (C ? *std::max_element(dc.begin(), dc.end()) : 0) + 1;  // ba/ba.cpp:594
  o->SL = (L ? *std::max_element(dl.begin(), dl.end()) : 0) + 1;  // ba/ba.cpp:595
  o->cam_b_eta.assign((size_t)C * 6, 0.f);
  o->cam_b_lam.assign((size_t)C * 36, 0.f);
  o->lmk_b_eta.assign((size_t)L * 3, 0.f);
  o->lmk_b_lam.assign((size_t)L * 9, 0.f);
  o->cam_scaling.assign(p->cam_scaling, p->cam_scaling + C);
  o->lmk_scaling.assign(p->lmk_scaling, p->lmk_scaling + L);
  o->cam_wflag.assign(p->cam_weaken_flag, p->cam_weaken_flag + C);
  o->lmk_wflag.assign(p->lmk_weaken_flag, p->lmk_weaken_flag + L);
  // Quirk Q4: all non-prior slots and every p-message start at zero.
  o->cam_m_eta.assign((size_t)C * o->SK * 6, 0.f);
  o->cam_m_lam.assign((size_t)C * o->SK * 36, 0.f);
  o->lmk_m_eta.assign((size_t)L * o->SL * 3, 0.f);
  o->lmk_m_lam.assign((size_t)L * o->SL * 9, 0.f);
  o->pcam_m_eta = o->cam_m_eta;
  o->pcam_m_lam = o->cam_m_lam;
  o->plmk_m_eta = o->lmk_m_eta;
  o->plmk_m_lam = o->lmk_m_lam;
  for (uint32_t c = 0; c < C; ++c) {  // ba/ba.cpp:880-881
    std::memcpy(&o->cam_m_eta[cme(*o, c, 0)], p->cam_priors_eta + 6 * (size_t)c, 6 * sizeof(float));
    std::memcpy(&o->cam_m_lam[cml(*o, c, 0)], p->cam_priors_lambda + 36 * (size_t)c, 36 * sizeof(float));
  }
  for (uint32_t l = 0; l < L; ++l) {  // ba/ba.cpp:882-883
    std::memcpy(&o->lmk_m_eta[lme(*o, l, 0)], p->lmk_priors_eta + 3 * (size_t)l, 3 * sizeof(float));
    std::memcpy(&o->lmk_m_lam[lml(*o, l, 0)], p->lmk_priors_lambda + 9 * (size_t)l, 9 * sizeof(float));
  }
  o->damping.assign(E, 0.f);
  if (p->damping) o->damping.assign(p->damping, p->damping + E);
  o->damping_count.assign(E, -15);
  if (p->damping_count) o->damping_count.assign(p->damping_count, p->damping_count + E);
  o->mu.assign((size_t)E * 9, 0.f);
  if (p->mu) o->mu.assign(p->mu, p->mu + (size_t)E * 9);
  o->oldmu.assign((size_t)E * 9, 0.f);
  if (p->oldmu) o->oldmu.assign(p->oldmu, p->oldmu + (size_t)E * 9);
  o->dmu.assign(E, 0.f);
  o->active.assign(E, 1u);
  if (p->active_flag) o->active.assign(p->active_flag, p->active_flag + E);
  o->robust.assign(E, 0u);
  o->f_eta.assign((size_t)E * 9, 0.f);
  o->f_lam.assign((size_t)E * 81, 0.f);
  // LINEARISE_PROG (ba/ba.cpp:890-893)
  update_beliefs(*o);
  relinearise_all(*o);
  *out = o;
  return GBP_OK;
}

int gbp_oracle_free(void* h) {
  delete (Oracle*)h;
  return GBP_OK;
}

int gbp_oracle_weaken_priors(void* h) {  // ba/ba.cpp:863-865
  Oracle& o = *(Oracle*)h;
  weaken_all(o);
  update_beliefs(o);
  return GBP_OK;
}

int gbp_oracle_eval(void* h, gbp_iter_stats* st) {
  eval_metrics(*(Oracle*)h, st);
  return GBP_OK;
}

int gbp_oracle_iterate(void* h, int n_sweeps, gbp_iter_stats* stats) {
  Oracle& o = *(Oracle*)h;
  double ms = 0;
  for (int i = 0; i < n_sweeps; ++i) {
    const auto t0 = std::chrono::steady_clock::now();
    gbp_sweep(o);
    const auto t1 = std::chrono::steady_clock::now();
    ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (stats) eval_metrics(o, &stats[i]);
  }
  o.last_ms = ms;  // sweeps only; metric evaluation excluded
  return GBP_OK;
}

int gbp_oracle_last_timing(void* h, double* ms_total) {
  *ms_total = ((Oracle*)h)->last_ms;
  return GBP_OK;
}

int gbp_oracle_get_beliefs(void* h, float* cam_eta, float* cam_lambda, float* lmk_eta,
                           float* lmk_lambda, float* damping, int32_t* damping_count,
                           uint32_t* robust_flag) {  // READ_PROG, ba/ba.cpp:908-916
  Oracle& o = *(Oracle*)h;
  if (cam_eta) std::memcpy(cam_eta, o.cam_b_eta.data(), o.cam_b_eta.size() * 4);
  if (cam_lambda) std::memcpy(cam_lambda, o.cam_b_lam.data(), o.cam_b_lam.size() * 4);
  if (lmk_eta) std::memcpy(lmk_eta, o.lmk_b_eta.data(), o.lmk_b_eta.size() * 4);
  if (lmk_lambda) std::memcpy(lmk_lambda, o.lmk_b_lam.data(), o.lmk_b_lam.size() * 4);
  if (damping) std::memcpy(damping, o.damping.data(), o.damping.size() * 4);
  if (damping_count) std::memcpy(damping_count, o.damping_count.data(), o.damping_count.size() * 4);
  if (robust_flag) std::memcpy(robust_flag, o.robust.data(), o.robust.size() * 4);
  return GBP_OK;
}

int gbp_oracle_get_priors(void* h, float* cam_eta, float* cam_lambda, float* lmk_eta,
                          float* lmk_lambda) {  // READ_PRIORS, ba/slam.cpp:913-917
  Oracle& o = *(Oracle*)h;
  for (uint32_t c = 0; c < o.C; ++c) {
    if (cam_eta) std::memcpy(cam_eta + 6 * (size_t)c, &o.cam_m_eta[cme(o, c, 0)], 24);
    if (cam_lambda) std::memcpy(cam_lambda + 36 * (size_t)c, &o.cam_m_lam[cml(o, c, 0)], 144);
  }
  for (uint32_t l = 0; l < o.L; ++l) {
    if (lmk_eta) std::memcpy(lmk_eta + 3 * (size_t)l, &o.lmk_m_eta[lme(o, l, 0)], 12);
    if (lmk_lambda) std::memcpy(lmk_lambda + 9 * (size_t)l, &o.lmk_m_lam[lml(o, l, 0)], 36);
  }
  return GBP_OK;
}

int gbp_oracle_add_keyframe(void* h, const int32_t* damping_count, const float* cam_prior_eta,
                            const float* cam_prior_lambda, const float* lmk_prior_eta,
                            const float* lmk_prior_lambda, const uint32_t* active_flag,
                            const uint32_t* cam_weaken_flag,
                            const uint32_t* lmk_weaken_flag) {  // NEW_KEYFRAME, ba/slam.cpp:919-928
  Oracle& o = *(Oracle*)h;
  if (damping_count) o.damping_count.assign(damping_count, damping_count + o.E);
  for (uint32_t c = 0; c < o.C; ++c) {
    if (cam_prior_eta) std::memcpy(&o.cam_m_eta[cme(o, c, 0)], cam_prior_eta + 6 * (size_t)c, 24);
    if (cam_prior_lambda) std::memcpy(&o.cam_m_lam[cml(o, c, 0)], cam_prior_lambda + 36 * (size_t)c, 144);
  }
  for (uint32_t l = 0; l < o.L; ++l) {
    if (lmk_prior_eta) std::memcpy(&o.lmk_m_eta[lme(o, l, 0)], lmk_prior_eta + 3 * (size_t)l, 12);
    if (lmk_prior_lambda) std::memcpy(&o.lmk_m_lam[lml(o, l, 0)], lmk_prior_lambda + 9 * (size_t)l, 36);
  }
  if (active_flag) o.active.assign(active_flag, active_flag + o.E);
  if (cam_weaken_flag) o.cam_wflag.assign(cam_weaken_flag, cam_weaken_flag + o.C);
  if (lmk_weaken_flag) o.lmk_wflag.assign(lmk_weaken_flag, lmk_weaken_flag + o.L);
  update_beliefs(o);
  return GBP_OK;
}

// codelet-level entry points
int gbp_oracle_relinearise_factors(void* h) {
  relinearise_all(*(Oracle*)h);
  return GBP_OK;
}
int gbp_oracle_prep_messages(void* h) {
  prep_all(*(Oracle*)h);
  return GBP_OK;
}
int gbp_oracle_compute_messages(void* h) {  // Copy(mu,oldmu) + cs_computemessages (ba/ba.cpp:898-899)
  Oracle& o = *(Oracle*)h;
  o.oldmu = o.mu;
  messages_all(o);
  return GBP_OK;
}
int gbp_oracle_update_beliefs(void* h) {
  update_beliefs(*(Oracle*)h);
  return GBP_OK;
}
int gbp_oracle_commit_messages(void* h) {  // the four Copy(messages, pmessages), ba/ba.cpp:902-905
  commit_messages(*(Oracle*)h);
  return GBP_OK;
}
int gbp_oracle_weaken_prior_vertices(void* h) {
  weaken_all(*(Oracle*)h);
  return GBP_OK;
}

int gbp_oracle_dims(void* h, uint32_t* C, uint32_t* L, uint32_t* E, uint32_t* max_nkfedges,
                    uint32_t* max_nlmkedges) {
  Oracle& o = *(Oracle*)h;
  if (C) *C = o.C;
  if (L) *L = o.L;
  if (E) *E = o.E;
  if (max_nkfedges) *max_nkfedges = o.SK - 1;
  if (max_nlmkedges) *max_nlmkedges = o.SL - 1;
  return GBP_OK;
}

int gbp_oracle_tensor_nbytes(void* h, const char* name, size_t* nbytes) {
  TensorRef t;
  if (!find_tensor(*(Oracle*)h, name, &t)) return GBP_ERR_NAME;
  *nbytes = t.nbytes;
  return GBP_OK;
}
int gbp_oracle_get_tensor(void* h, const char* name, void* dst, size_t nbytes) {
  TensorRef t;
  if (!find_tensor(*(Oracle*)h, name, &t)) return GBP_ERR_NAME;
  if (t.nbytes != nbytes) return GBP_ERR_SIZE;
  std::memcpy(dst, t.ptr, nbytes);
  return GBP_OK;
}
int gbp_oracle_set_tensor(void* h, const char* name, const void* src, size_t nbytes) {
  TensorRef t;
  if (!find_tensor(*(Oracle*)h, name, &t)) return GBP_ERR_NAME;
  if (t.nbytes != nbytes) return GBP_ERR_SIZE;
  std::memcpy(t.ptr, src, nbytes);
  return GBP_OK;
}

// ---- single-vertex probes (unit parity of the math helpers) ---------------
// out must hold 36 / 9 floats.  Exposed so the tests can compare the device
// helpers and the two oracle builds on identical inputs.
void gbp_oracle_inv6x6(const float* A, float* out) {
  for (int i = 0; i < 36; ++i) out[i] = 0.f;
#ifdef GBP_ORACLE_USE_REFERENCE
  Mat<float> a(const_cast<float*>(A), 6, 6), r(out, 6, 6);
  inv6x6(a, r);
#else
  gbp_restated::inv6(A, out);
#endif
}
void gbp_oracle_inv3x3(const float* A, float* out) {
#ifdef GBP_ORACLE_USE_REFERENCE
  Mat<float> a(const_cast<float*>(A), 3, 3), r(out, 3, 3);
  inv3x3(a, r);
#else
  gbp_restated::inv3(A, out);
#endif
}
// hx[2], Jkf[12], Jlmk[6] for camera x[6], landmark p[3], K[9].
void gbp_oracle_project(const float* x, const float* p, const float* K, float* hx, float* Jkf,
                        float* Jlmk) {
  for (int i = 0; i < 2; ++i) hx[i] = 0.f;
  for (int i = 0; i < 12; ++i) Jkf[i] = 0.f;
  for (int i = 0; i < 6; ++i) Jlmk[i] = 0.f;
#ifdef GBP_ORACLE_USE_REFERENCE
  Mat<float> mx(const_cast<float*>(x), 6, 1), mp(const_cast<float*>(p), 3, 1),
      mK(const_cast<float*>(K), 3, 3), mh(hx, 2, 1), mJk(Jkf, 2, 6), mJl(Jlmk, 2, 3);
  hfunc(mx, mp, mK, mh);
  Jac(mx, mp, mK, mJk, mJl);
#else
  gbp_restated::hfunc(x, p, K, hx);
  gbp_restated::jac(x, p, K, Jkf, Jlmk);
#endif
}

}  // extern "C"
