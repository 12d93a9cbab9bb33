// Host-side problem setup around the GBP hot path (pure host code, no CUDA).
//
// Re-hosts, without Eigen / TBB / Boost (absent in this image), what the
// reference's mains do before WRITE_PROG and at SLAM keyframe insertion:
//   BALProblem::LoadFile            ba/dataio.cpp:17-57
//   set_prior_lambda                ba/dataio.cpp:67-117  (O(E) here, O(V*E) there)
//   reprojectionJacFn               ba/util.cpp:48-72
//   weakening scale factors         ba/ba.cpp:560-572
//   noise / average-depth init      ba/dataio.cpp:330-453
//   create_flags / update_flags     ba/dataio.cpp:455-508
//   initialise_new_kf               ba/util.cpp:183-223
// plus the synthetic BAL-format generator for the large configs.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "../../include/gbp_host.h"

void gbp_set_error(const std::string& s);  // gbp_cuda_api.cu / host_error.cpp

struct gbp_bal {
  uint32_t C = 0, L = 0, E = 0;
  double intr[4] = {0, 0, 0, 0};  // fx fy cx cy
  std::vector<uint32_t> cam_idx, lmk_idx;
  std::vector<double> obs;     // 2E
  std::vector<double> params;  // 6C + 3L
};

struct gbp_setup {
  gbp_problem prob;
  gbp_cli_options opt;
  int mode = GBP_MODE_BA;
  uint32_t steps = 5;
  uint32_t data_counter = 0;
  std::vector<uint32_t> cam_ids, lmk_ids;
  std::vector<float> z, var;
  std::vector<float> cam_mean, lmk_mean;
  std::vector<float> cam_p_eta, cam_p_lam, lmk_p_eta, lmk_p_lam;
  std::vector<float> cam_scaling, lmk_scaling;
  std::vector<uint32_t> cam_wflag, lmk_wflag, lmk_active, active;
  std::vector<float> damping, mu, oldmu;
  std::vector<int32_t> damping_count;
};

namespace {

// ---- small fp32 helpers (Eigen-free restatements of ba/util.cpp:11-46) -----
// trig_in_double: sin/cos evaluated in double and rounded once (what a correctly rounded sinf/cosf
// returns; glibc's differ from that by 1 ulp for <0.1 % of the arguments).  Used where the device
// evaluates the same expression (k_kf_pose), so that host and device keyframe insertion agree bit for bit.
void so3exp_f(const float* w, float* R, bool trig_in_double = false) {  // ba/util.cpp:20-32
  const float theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.f : 0.f;
  if (theta < 1e-6) return;
  const float H[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
  const float sn = trig_in_double ? (float)std::sin((double)theta) : std::sin(theta);
  const float cs = trig_in_double ? (float)std::cos((double)theta) : std::cos(theta);
  const float a = sn / theta, b = (1 - cs) / (theta * theta);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float h2 = 0.f;
      for (int k = 0; k < 3; ++k) h2 += H[i * 3 + k] * H[k * 3 + j];
      R[i * 3 + j] += a * H[i * 3 + j] + b * h2;
    }
}

void so3log_f(const float* R, float* w) {  // ba/util.cpp:34-46
  const float d = 0.5f * (R[0] + R[4] + R[8] - 1);
  const float s = std::acos(d) / (2 * std::sqrt(1 - d * d));
  w[0] = s * (R[7] - R[5]);
  w[1] = s * (R[2] - R[6]);
  w[2] = s * (R[3] - R[1]);
}

// max |J| of the 2x9 reprojection Jacobian (ba/util.cpp:48-72, ba/dataio.cpp:83-85).
float max_abs_reproj_jac(const float* cam, const float* lmk, const float* K) {
  float R[9];
  so3exp_f(cam + 3, R);
  float Rl[3], pc[3], p[3];
  for (int i = 0; i < 3; ++i) {
    Rl[i] = R[i * 3] * lmk[0] + R[i * 3 + 1] * lmk[1] + R[i * 3 + 2] * lmk[2];
    pc[i] = Rl[i] + cam[i];
  }
  for (int i = 0; i < 3; ++i) p[i] = K[i * 3] * pc[0] + K[i * 3 + 1] * pc[1] + K[i * 3 + 2] * pc[2];
  const float p2sq = (float)std::pow((double)p[2], 2);
  const float jp[6] = {1 / p[2], 0.f, -p[0] / p2sq, 0.f, 1 / p[2], -p[1] / p2sq};
  float jpK[6];  // j_proj * K
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j)
      jpK[i * 3 + j] = jp[i * 3] * K[j] + jp[i * 3 + 1] * K[3 + j] + jp[i * 3 + 2] * K[6 + j];
  const float dR[9] = {0.f, Rl[2], -Rl[1], -Rl[2], 0.f, Rl[0], Rl[1], -Rl[0], 0.f};  // -hat(R*lmk)
  float m = 0.f;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) {
      const float a = jpK[i * 3 + j];
      const float b = jpK[i * 3] * dR[j] + jpK[i * 3 + 1] * dR[3 + j] + jpK[i * 3 + 2] * dR[6 + j];
      const float c = jpK[i * 3] * R[j] + jpK[i * 3 + 1] * R[3 + j] + jpK[i * 3 + 2] * R[6 + j];
      m = std::max(m, std::max(std::fabs(a), std::max(std::fabs(b), std::fabs(c))));
    }
  return m;
}

// x = A^-1 b for a small dense system (partial pivoting, double).
bool solve_small(int n, const float* A_in, const float* b_in, float* x_out) {
  double A[36], b[6], x[6];
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
  for (int i = 0; i < n; ++i) x_out[i] = (float)x[i];
  return true;
}

// World position of the point 1 m in front of a camera with pose x[6]
// (Tw2c.inverse() * (0,0,1,1), ba/dataio.cpp:427-440, ba/util.cpp:199-211).
void point_in_front(const float* x, float* p_world, bool trig_in_double = false) {
  float R[9];
  so3exp_f(x + 3, R, trig_in_double);
  const float d[3] = {0.f - x[0], 0.f - x[1], 1.f - x[2]};
  for (int i = 0; i < 3; ++i) p_world[i] = R[i] * d[0] + R[3 + i] * d[1] + R[6 + i] * d[2];  // R^T d
}

unsigned noise_seed(const gbp_cli_options& o) {
  if (o.noise_seed) return o.noise_seed;
  return (unsigned)std::chrono::system_clock::now().time_since_epoch().count();
}

// ---- tokenizer for the BAL-like text format -------------------------------
struct Tokens {
  const char* p;
  const char* end;
  bool warned = false;
  void skip() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  }
  void bad() { std::cout << "Invalid UW data file."; }  // ba/dataio.cpp:59-65
  bool next_u32(uint32_t* v) {
    skip();
    if (p >= end) {
      bad();
      return false;
    }
    char* q;
    const long long x = std::strtoll(p, &q, 10);
    if (q == p) {
      bad();
      return false;
    }
    *v = (uint32_t)x;
    p = q;
    return true;
  }
  bool next_f64(double* v) {
    skip();
    if (p >= end) {
      bad();
      return false;
    }
    char* q;
    const double x = std::strtod(p, &q);
    if (q == p) {
      bad();
      return false;
    }
    *v = x;
    p = q;
    return true;
  }
};

// deterministic RNG for the synthetic generator
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  double normal() {
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
  int poisson(double mean) {
    const double l = std::exp(-mean);
    int k = 0;
    double p = 1.0;
    do {
      ++k;
      p *= uni();
    } while (p > l);
    return k - 1;
  }
};

void so3exp_d(const double* w, double* R) {
  const double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if (th < 1e-12) return;
  const double H[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  const double a = std::sin(th) / th, b = (1 - std::cos(th)) / (th * th);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double h2 = 0;
      for (int k = 0; k < 3; ++k) h2 += H[i * 3 + k] * H[k * 3 + j];
      R[i * 3 + j] += a * H[i * 3 + j] + b * h2;
    }
}

}  // namespace

extern "C" {

void gbp_cli_options_default(gbp_cli_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->n_iters = 1500;
  o->iters_between_kfs = 700;
  o->n_ipus = 1;
  o->cams_per_tile = 1;
  o->av_depth = 1.f;
  o->reproj_meas_var = 4.f;
  o->prior_std_weaker_factor = 100.f;
  o->first_cam_prior_std = 0.01f;
  o->steps = 5.f;
  o->iters_before_damping = 15;
}

int gbp_bal_load(const char* path, gbp_bal** out) {
  if (!path || !out) return GBP_ERR_ARG;
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    gbp_set_error(std::string("unable to open file ") + path);
    return GBP_ERR_IO;
  }
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<char> buf((size_t)n + 1);
  const size_t got = std::fread(buf.data(), 1, (size_t)n, f);
  std::fclose(f);
  buf[got] = 0;
  Tokens t{buf.data(), buf.data() + got};
  gbp_bal* b = new gbp_bal();
  t.next_u32(&b->C);
  t.next_u32(&b->L);
  t.next_u32(&b->E);
  for (int i = 0; i < 4; ++i) t.next_f64(&b->intr[i]);
  b->cam_idx.assign(b->E, 0);
  b->lmk_idx.assign(b->E, 0);
  b->obs.assign((size_t)2 * b->E, 0.0);
  b->params.assign((size_t)6 * b->C + (size_t)3 * b->L, 0.0);
  bool ok = true;
  for (uint32_t i = 0; i < b->E && ok; ++i) {
    ok = t.next_u32(&b->cam_idx[i]) && t.next_u32(&b->lmk_idx[i]) && t.next_f64(&b->obs[2 * (size_t)i]) &&
         t.next_f64(&b->obs[2 * (size_t)i + 1]);
  }
  for (size_t i = 0; i < b->params.size() && ok; ++i) ok = t.next_f64(&b->params[i]);
  *out = b;
  return GBP_OK;
}

int gbp_bal_from_arrays(uint32_t C, uint32_t L, uint32_t E, const double intrinsics[4],
                        const uint32_t* cam_idx, const uint32_t* lmk_idx, const double* observations,
                        const double* cameras, const double* points, gbp_bal** out) {
  if (!out || !intrinsics || !cam_idx || !lmk_idx || !observations || !cameras || !points) return GBP_ERR_ARG;
  gbp_bal* b = new gbp_bal();
  b->C = C;
  b->L = L;
  b->E = E;
  for (int i = 0; i < 4; ++i) b->intr[i] = intrinsics[i];
  b->cam_idx.assign(cam_idx, cam_idx + E);
  b->lmk_idx.assign(lmk_idx, lmk_idx + E);
  b->obs.assign(observations, observations + 2 * (size_t)E);
  b->params.assign(cameras, cameras + 6 * (size_t)C);
  b->params.insert(b->params.end(), points, points + 3 * (size_t)L);
  *out = b;
  return GBP_OK;
}

int gbp_bal_save(const gbp_bal* b, const char* path) {  // format: sequences/README.md:5-16
  if (!b || !path) return GBP_ERR_ARG;
  FILE* f = std::fopen(path, "w");
  if (!f) {
    gbp_set_error(std::string("unable to open file for writing ") + path);
    return GBP_ERR_IO;
  }
  std::fprintf(f, "%u %u %u\n", b->C, b->L, b->E);
  std::fprintf(f, "%.9g %.9g %.9g %.9g\n", b->intr[0], b->intr[1], b->intr[2], b->intr[3]);
  for (uint32_t e = 0; e < b->E; ++e)
    std::fprintf(f, "%u %u     %.6e %.6e\n", b->cam_idx[e], b->lmk_idx[e], b->obs[2 * (size_t)e],
                 b->obs[2 * (size_t)e + 1]);
  for (double v : b->params) std::fprintf(f, "%.16e\n", v);
  std::fclose(f);
  return GBP_OK;
}

// A copy of `b` whose camera / point parameters are the means of the given beliefs (mu = Lambda^-1 eta,
// solved in double with partial pivoting).  What the reference's never-called save_cam_means /
// save_lmk_means (ba/dataio.cpp:205-255) were meant to give: the optimised problem, here in the INPUT format.
int gbp_bal_with_means(const gbp_bal* b, const float* cam_eta, const float* cam_lambda, const float* lmk_eta,
                       const float* lmk_lambda, gbp_bal** out) {
  if (!b || !cam_eta || !cam_lambda || !lmk_eta || !lmk_lambda || !out) return GBP_ERR_ARG;
  auto solve = [](int n, const float* lam, const float* eta, double* x) {
    double A[36], r[6];
    for (int i = 0; i < n * n; ++i) A[i] = lam[i];
    for (int i = 0; i < n; ++i) r[i] = eta[i];
    for (int k = 0; k < n; ++k) {
      int piv = k;
      for (int i = k + 1; i < n; ++i)
        if (std::fabs(A[i * n + k]) > std::fabs(A[piv * n + k])) piv = i;
      if (piv != k) {
        for (int c = 0; c < n; ++c) std::swap(A[k * n + c], A[piv * n + c]);
        std::swap(r[k], r[piv]);
      }
      for (int i = k + 1; i < n; ++i) {
        const double f = A[i * n + k] / A[k * n + k];
        for (int c = k; c < n; ++c) A[i * n + c] -= f * A[k * n + c];
        r[i] -= f * r[k];
      }
    }
    for (int k = n - 1; k >= 0; --k) {
      double acc = r[k];
      for (int c = k + 1; c < n; ++c) acc -= A[k * n + c] * x[c];
      x[k] = acc / A[k * n + k];
    }
  };
  gbp_bal* o = new gbp_bal(*b);
  for (uint32_t c = 0; c < b->C; ++c) solve(6, cam_lambda + (size_t)36 * c, cam_eta + (size_t)6 * c, &o->params[(size_t)6 * c]);
  for (uint32_t l = 0; l < b->L; ++l)
    solve(3, lmk_lambda + (size_t)9 * l, lmk_eta + (size_t)3 * l, &o->params[(size_t)6 * b->C + (size_t)3 * l]);
  *out = o;
  return GBP_OK;
}

void gbp_bal_free(gbp_bal* b) { delete b; }

int gbp_bal_dims(const gbp_bal* b, uint32_t* C, uint32_t* L, uint32_t* E) {
  if (!b) return GBP_ERR_ARG;
  if (C) *C = b->C;
  if (L) *L = b->L;
  if (E) *E = b->E;
  return GBP_OK;
}
const uint32_t* gbp_bal_camera_index(const gbp_bal* b) { return b->cam_idx.data(); }
const uint32_t* gbp_bal_point_index(const gbp_bal* b) { return b->lmk_idx.data(); }
const double* gbp_bal_observations(const gbp_bal* b) { return b->obs.data(); }
const double* gbp_bal_parameters(const gbp_bal* b) { return b->params.data(); }
const double* gbp_bal_intrinsics(const gbp_bal* b) { return b->intr; }

int gbp_setup_create(const gbp_bal* b, const gbp_cli_options* opt_in, int mode, gbp_setup** out) {
  if (!b || !out) return GBP_ERR_ARG;
  gbp_cli_options o;
  if (opt_in) o = *opt_in; else gbp_cli_options_default(&o);
  const uint32_t C = b->C, L = b->L, E = b->E;
  for (uint32_t e = 0; e < E; ++e)
    if (b->cam_idx[e] >= C || b->lmk_idx[e] >= L) {
      gbp_set_error("edge refers to a camera / landmark outside the declared counts");
      return GBP_ERR_ARG;
    }
  if (mode == GBP_MODE_SLAM && C < 2) {
    gbp_set_error("SLAM mode needs at least two keyframes");
    return GBP_ERR_ARG;
  }
  gbp_setup* s = new gbp_setup();
  s->opt = o;
  s->mode = mode;
  s->steps = (uint32_t)o.steps;
  std::memset(&s->prob, 0, sizeof(s->prob));
  // K, measurements (ba/ba.cpp:494-512)
  const float K[9] = {(float)b->intr[0], 0.f, (float)b->intr[2], 0.f, (float)b->intr[1],
                      (float)b->intr[3], 0.f, 0.f, 1.f};
  std::memcpy(s->prob.K, K, sizeof(K));
  s->cam_ids = b->cam_idx;
  s->lmk_ids = b->lmk_idx;
  s->z.resize((size_t)2 * E);
  for (size_t i = 0; i < s->z.size(); ++i) s->z[i] = (float)b->obs[i];
  s->var.assign(E, o.reproj_meas_var);
  // prior means (ba/ba.cpp:523-534)
  s->cam_mean.resize((size_t)6 * C);
  for (size_t i = 0; i < s->cam_mean.size(); ++i) s->cam_mean[i] = (float)b->params[i];
  s->lmk_mean.resize((size_t)3 * L);
  for (size_t i = 0; i < s->lmk_mean.size(); ++i) s->lmk_mean[i] = (float)b->params[(size_t)6 * C + i];

  // optional noise (ba/ba.cpp:536-548, ba/dataio.cpp:330-453)
  if (o.transnoise != 0.f) {
    std::cout << "\nAdding Gaussian noise with std: " << o.transnoise
              << "m to the keyframe translaton intialisations\n";
    std::default_random_engine gen(noise_seed(o));
    std::normal_distribution<float> nd(0.f, o.transnoise);
    for (uint32_t c = 2; c < C; ++c)
      for (int i = 0; i < 3; ++i) s->cam_mean[c * 6 + i] += nd(gen);
  }
  if (o.rotnoise != 0.f) {
    std::cout << "Adding Gaussian noise with std: " << o.rotnoise << " to the keyframe rotation intialisations\n";
    std::default_random_engine gen(noise_seed(o));
    std::normal_distribution<float> nd(0.f, o.rotnoise);
    for (uint32_t c = 2; c < C; ++c) {
      const float ang = nd(gen) * (float)M_PI / 180;
      const int axis = std::rand() % 3;
      const float cs = std::cos(ang), sn = std::sin(ang);
      float Rn[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      if (axis == 0) { Rn[4] = cs; Rn[5] = -sn; Rn[7] = sn; Rn[8] = cs; }
      else if (axis == 1) { Rn[0] = cs; Rn[2] = sn; Rn[6] = -sn; Rn[8] = cs; }
      else { Rn[0] = cs; Rn[1] = -sn; Rn[3] = sn; Rn[4] = cs; }
      float* x = &s->cam_mean[c * 6];
      float R[9], Rnew[9], tmp[3], tnew[3];
      so3exp_f(x + 3, R);
      // Tc2w = [R^T, -R^T t];  R_c2w' = Rn R^T  =>  R' = R Rn^T,  t' = R' R^T t
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          float v = 0.f;
          for (int k = 0; k < 3; ++k) v += R[i * 3 + k] * Rn[j * 3 + k];
          Rnew[i * 3 + j] = v;
        }
      for (int i = 0; i < 3; ++i) tmp[i] = R[i] * x[0] + R[3 + i] * x[1] + R[6 + i] * x[2];
      for (int i = 0; i < 3; ++i) tnew[i] = Rnew[i * 3] * tmp[0] + Rnew[i * 3 + 1] * tmp[1] + Rnew[i * 3 + 2] * tmp[2];
      for (int i = 0; i < 3; ++i) x[i] = tnew[i];
      so3log_f(Rnew, x + 3);
    }
  }
  if (o.lmktrans_noise != 0.f && !o.av_depth_on) {
    std::cout << "Adding Gaussian noise with std: " << o.lmktrans_noise << "m to the landmark intialisations\n";
    std::default_random_engine gen(noise_seed(o));
    std::normal_distribution<float> nd(0.f, o.lmktrans_noise);
    for (size_t i = 0; i < s->lmk_mean.size(); ++i) s->lmk_mean[i] += nd(gen);
  } else if (o.av_depth_on) {
    std::cout << "Initialising all landmarks at an average depth of: " << o.av_depth << "\n";
    // every landmark takes the point in front of the lowest-numbered camera observing it
    std::vector<uint32_t> first_cam(L, UINT32_MAX);
    for (uint32_t e = 0; e < E; ++e) first_cam[s->lmk_ids[e]] = std::min(first_cam[s->lmk_ids[e]], s->cam_ids[e]);
    std::vector<float> front((size_t)3 * C);
    for (uint32_t c = 0; c < C; ++c) point_in_front(&s->cam_mean[c * 6], &front[c * 3]);
    for (uint32_t l = 0; l < L; ++l)
      if (first_cam[l] != UINT32_MAX)
        for (int i = 0; i < 3; ++i) s->lmk_mean[l * 3 + i] = front[first_cam[l] * 3 + i];
  }

  // set_prior_lambda (ba/dataio.cpp:67-117): lambda = max|J|^2 / var on the diagonal.
  // The Jacobian is evaluated at the FILE values (not the noised means), once per edge.
  std::vector<float> cam_maxj(C, 0.f), lmk_maxj(L, 0.f);
  for (uint32_t e = 0; e < E; ++e) {
    const uint32_t c = s->cam_ids[e], l = s->lmk_ids[e];
    float cam[6], lmk[3];
    for (int i = 0; i < 6; ++i) cam[i] = (float)b->params[(size_t)6 * c + i];
    for (int i = 0; i < 3; ++i) lmk[i] = (float)b->params[(size_t)6 * C + (size_t)3 * l + i];
    const float m = max_abs_reproj_jac(cam, lmk, K);
    cam_maxj[c] = std::max(cam_maxj[c], m);
    lmk_maxj[l] = std::max(lmk_maxj[l], m);
  }
  s->cam_p_eta.assign((size_t)6 * C, 0.f);
  s->cam_p_lam.assign((size_t)36 * C, 0.f);
  s->lmk_p_eta.assign((size_t)3 * L, 0.f);
  s->lmk_p_lam.assign((size_t)9 * L, 0.f);
  for (uint32_t c = 0; c < C; ++c) {
    const float lam = (float)(std::pow((double)cam_maxj[c], 2) / o.reproj_meas_var);
    for (int i = 0; i < 6; ++i) {
      s->cam_p_eta[c * 6 + i] = s->cam_mean[c * 6 + i] * lam;
      s->cam_p_lam[(size_t)c * 36 + i * 6 + i] = lam;
    }
  }
  for (uint32_t l = 0; l < L; ++l) {
    const float lam = (float)(std::pow((double)lmk_maxj[l], 2) / o.reproj_meas_var);
    for (int i = 0; i < 3; ++i) {
      s->lmk_p_eta[l * 3 + i] = s->lmk_mean[l * 3 + i] * lam;
      s->lmk_p_lam[(size_t)l * 9 + i * 3 + i] = lam;
    }
  }
  // weakening scale factors (ba/ba.cpp:560-572)
  s->cam_scaling.resize(C);
  for (uint32_t c = 0; c < C; ++c) {
    if (c == 0 || c == 1)
      s->cam_scaling[c] = (float)std::exp(-1 / o.steps * std::log(s->cam_p_lam[(size_t)c * 36] *
                                                                   std::pow((double)o.first_cam_prior_std, 2)));
    else
      s->cam_scaling[c] = std::exp(-2 / o.steps * std::log(o.prior_std_weaker_factor));
  }
  s->lmk_scaling.assign(L, std::exp(-2 / o.steps * std::log(o.prior_std_weaker_factor)));
  // damping state (ba/ba.cpp:580-584)
  s->damping.assign(E, 0.f);
  s->damping_count.assign(E, -o.iters_before_damping);
  if (mode == GBP_MODE_BA) {  // ba/ba.cpp:588-590
    s->active.assign(E, 1u);
    s->cam_wflag.assign(C, (uint32_t)o.steps);
    s->lmk_wflag.assign(L, (uint32_t)o.steps);
    s->lmk_active.assign(L, (uint32_t)o.steps);
  } else {  // create_flags, ba/dataio.cpp:455-475
    s->active.assign(E, 0u);
    s->cam_wflag.assign(C, 0u);
    s->lmk_wflag.assign(L, 0u);
    s->cam_wflag[0] = s->steps;
    s->cam_wflag[1] = s->steps;
    for (uint32_t e = 0; e < E; ++e)
      if (s->cam_ids[e] == 0 || s->cam_ids[e] == 1) {
        s->active[e] = 1;
        s->lmk_wflag[s->lmk_ids[e]] = s->steps;
      }
    s->lmk_active = s->lmk_wflag;
  }
  gbp_problem& p = s->prob;
  p.n_keyframes = C;
  p.n_points = L;
  p.n_edges = E;
  p.cam_ids = s->cam_ids.data();
  p.lmk_ids = s->lmk_ids.data();
  p.measurements = s->z.data();
  p.meas_variances = s->var.data();
  p.cam_priors_eta = s->cam_p_eta.data();
  p.cam_priors_lambda = s->cam_p_lam.data();
  p.lmk_priors_eta = s->lmk_p_eta.data();
  p.lmk_priors_lambda = s->lmk_p_lam.data();
  p.cam_scaling = s->cam_scaling.data();
  p.lmk_scaling = s->lmk_scaling.data();
  p.cam_weaken_flag = s->cam_wflag.data();
  p.lmk_weaken_flag = s->lmk_wflag.data();
  // optional arrays that hold nothing but the defaults of include/gbp_cuda.h are passed as NULL: the engine then
  // neither copies nor uploads them (12 bytes per factor at init)
  p.active_flag = (mode == GBP_MODE_BA) ? nullptr : s->active.data();         // BA: every factor active (ba/ba.cpp:588)
  p.damping = nullptr;                                                          // all zero (ba/ba.cpp:580)
  p.damping_count = (o.iters_before_damping == 15) ? nullptr : s->damping_count.data();  // -15 unless --iters_before_damping says otherwise
  // mu / oldmu: the reference streams zeros (ba/ba.cpp:582-583); NULL means exactly that (include/gbp_cuda.h) and
  // spares the engine two 9E-float arrays it would only scan for a non-zero
  p.mu = nullptr;
  p.oldmu = nullptr;
  *out = s;
  return GBP_OK;
}

void gbp_setup_free(gbp_setup* s) { delete s; }
const gbp_problem* gbp_setup_problem(const gbp_setup* s) { return s ? &s->prob : nullptr; }
int gbp_setup_data_counter(const gbp_setup* s) { return s ? (int)s->data_counter : -1; }

int gbp_setup_next_keyframe(gbp_setup* s, const float* cam_b_eta, const float* cam_b_lam,
                            float* cam_p_eta, float* cam_p_lam, float* lmk_p_eta, float* lmk_p_lam,
                            int32_t* damping_count, int* n_new_lmks) {
  if (!s || !cam_b_eta || !cam_b_lam || !cam_p_eta || !cam_p_lam || !lmk_p_eta || !lmk_p_lam) return GBP_ERR_ARG;
  const uint32_t C = s->prob.n_keyframes, L = s->prob.n_points, E = s->prob.n_edges;
  if (s->data_counter + 2 >= C) {
    gbp_set_error("no keyframe left to add");
    return GBP_ERR_ARG;
  }
  s->data_counter += 1;
  const uint32_t dc = s->data_counter;
  // update_flags (ba/dataio.cpp:477-508)
  for (uint32_t e = 0; e < E; ++e) {
    if (s->cam_ids[e] == dc + 1) s->active[e] = 1;
    if (s->cam_ids[e] <= dc + 1) s->lmk_wflag[s->lmk_ids[e]] = s->steps;
  }
  std::fill(s->cam_wflag.begin(), s->cam_wflag.end(), 0u);
  s->cam_wflag[dc + 1] = s->steps;
  uint32_t sum = 0;
  for (uint32_t l = 0; l < L; ++l) {
    s->lmk_wflag[l] -= s->lmk_active[l];
    s->lmk_active[l] += s->lmk_wflag[l];
    sum += s->lmk_wflag[l];
  }
  if (n_new_lmks) *n_new_lmks = s->steps ? (int)sum / (int)s->steps : 0;
  // initialise_new_kf (ba/util.cpp:183-223)
  float prev_mu[6];
  if (!solve_small(6, cam_b_lam + (size_t)dc * 36, cam_b_eta + (size_t)dc * 6, prev_mu)) {
    gbp_set_error("singular belief of the previous keyframe");
    return GBP_ERR_ARG;
  }
  const float* lam_new = cam_p_lam + (size_t)(dc + 1) * 36;
  for (int i = 0; i < 6; ++i) {
    float v = 0.f;
    for (int j = 0; j < 6; ++j) v += lam_new[i * 6 + j] * prev_mu[j];
    cam_p_eta[(size_t)(dc + 1) * 6 + i] = v;
  }
  float pw[3];
  point_in_front(prev_mu, pw, true);
  for (uint32_t l = 0; l < L; ++l) {
    // Quirk Q6: the reference indexes lmk_weaken_flag_[data_counter*n_points + i]
    // (out of bounds); the intended test is "newly observed landmark".
    if (s->lmk_wflag[l] == 5) {
      const float* Ll = lmk_p_lam + (size_t)l * 9;
      for (int i = 0; i < 3; ++i) lmk_p_eta[(size_t)l * 3 + i] = Ll[i * 3] * pw[0] + Ll[i * 3 + 1] * pw[1] + Ll[i * 3 + 2] * pw[2];
    }
  }
  // ba/slam.cpp:1039-1041 (quirk Q10: hard-coded -15)
  std::fill(s->damping_count.begin(), s->damping_count.end(), -15);
  if (damping_count) std::fill(damping_count, damping_count + E, -15);
  return GBP_OK;
}

int gbp_synth_generate(uint32_t C, uint32_t L, double obs_per_point, uint32_t seed, gbp_bal** out) {
  if (!out || C < 2 || L < 1 || obs_per_point < 2.0) return GBP_ERR_ARG;
  Rng rng(seed);
  const double fx = 517.306408, fy = 516.469215, cx = 318.64304, cy = 255.313989;  // fr1xyz intrinsics
  // ground-truth cameras: smooth trajectory along +x, looking down +z
  std::vector<double> cam_gt((size_t)6 * C), Rgt((size_t)9 * C);
  for (uint32_t i = 0; i < C; ++i) {
    const double ctr[3] = {0.05 * i, 0.1 * std::sin(0.01 * i), 0.05 * std::cos(0.013 * i)};
    double w[3] = {0.02 * std::sin(0.013 * i) + 0.004, 0.03 * std::cos(0.007 * i) + 0.003,
                   0.01 * std::sin(0.02 * i) + 0.002};
    double* R = &Rgt[(size_t)9 * i];
    so3exp_d(w, R);
    double* x = &cam_gt[(size_t)6 * i];
    for (int r = 0; r < 3; ++r) x[r] = -(R[r * 3] * ctr[0] + R[r * 3 + 1] * ctr[1] + R[r * 3 + 2] * ctr[2]);
    for (int r = 0; r < 3; ++r) x[3 + r] = w[r];
  }
  std::vector<double> pts_gt((size_t)3 * L);
  struct Obs { uint32_t c, l; double u, v; };
  std::vector<Obs> obs;
  obs.reserve((size_t)(L * (obs_per_point + 1)));
  std::vector<uint32_t> deg(C, 0);
  const int W = 40;
  std::vector<int> cand(2 * W + 1);
  std::vector<Obs> mine;
  for (uint32_t l = 0; l < L; ++l) {
    const uint32_t a = (uint32_t)(((uint64_t)l * C) / L);  // anchor camera, spread evenly
    double* p = &pts_gt[(size_t)3 * l];
    // re-draw the landmark until at least two cameras see it inside the image
    for (int attempt = 0; attempt < 200; ++attempt) {
      const double u = 20 + 600 * rng.uni(), v = 20 + 440 * rng.uni(), d = 0.5 + 4.5 * rng.uni();
      const double yc[3] = {d * (u - cx) / fx, d * (v - cy) / fy, d};
      const double* R = &Rgt[(size_t)9 * a];
      const double* x = &cam_gt[(size_t)6 * a];
      const double dd[3] = {yc[0] - x[0], yc[1] - x[1], yc[2] - x[2]};
      for (int i = 0; i < 3; ++i) p[i] = R[i] * dd[0] + R[3 + i] * dd[1] + R[6 + i] * dd[2];
      int k = rng.poisson(obs_per_point);
      k = std::max(2, std::min(k, (int)(2 * obs_per_point)));
      for (int i = 0; i <= 2 * W; ++i) cand[i] = i - W;
      for (int i = 2 * W; i > 0; --i) std::swap(cand[i], cand[rng.next() % (uint64_t)(i + 1)]);
      mine.clear();
      // the anchor is tried first; then random nearby cameras
      for (int ci = -1; ci <= 2 * W && (int)mine.size() < k; ++ci) {
        const int off = (ci < 0) ? 0 : cand[ci];
        if (ci >= 0 && off == 0) continue;
        const long cc = (long)a + off;
        if (cc < 0 || cc >= (long)C) continue;
        const double* Rc = &Rgt[(size_t)9 * cc];
        const double* xc = &cam_gt[(size_t)6 * cc];
        double y[3];
        for (int i = 0; i < 3; ++i) y[i] = Rc[i * 3] * p[0] + Rc[i * 3 + 1] * p[1] + Rc[i * 3 + 2] * p[2] + xc[i];
        if (y[2] < 0.45) continue;
        const double pu = fx * y[0] / y[2] + cx, pv = fy * y[1] / y[2] + cy;
        if (pu < 22 || pu > 618 || pv < 22 || pv > 458) continue;
        const double nu = std::max(-2.5, std::min(2.5, rng.normal())), nv = std::max(-2.5, std::min(2.5, rng.normal()));
        mine.push_back({(uint32_t)cc, l, pu + nu, pv + nv});
      }
      if (mine.size() >= 2) break;
    }
    for (const Obs& ob : mine) {
      obs.push_back(ob);
      deg[ob.c]++;
    }
  }
  // counting sort by camera (stable in landmark order)
  const size_t E = obs.size();
  std::vector<size_t> start(C + 1, 0);
  for (uint32_t c = 0; c < C; ++c) start[c + 1] = start[c] + deg[c];
  std::vector<uint32_t> ci(E), li(E);
  std::vector<double> zz(2 * E);
  std::vector<size_t> fill(start.begin(), start.end() - 1);
  for (const Obs& ob : obs) {
    const size_t k = fill[ob.c]++;
    ci[k] = ob.c;
    li[k] = ob.l;
    zz[2 * k] = ob.u;
    zz[2 * k + 1] = ob.v;
  }
  // initial values = ground truth + noise (first two cameras anchor the gauge).  The
  // pose noise is applied to the camera CENTRE and orientation (t = -R c), so a
  // rotation perturbation does not swing the camera around the world origin.
  std::vector<double> cam0 = cam_gt, pts0 = pts_gt;
  for (uint32_t c = 2; c < C; ++c) {
    const double* Rg = &Rgt[(size_t)9 * c];
    const double* xg = &cam_gt[(size_t)6 * c];
    double ctr[3], w[3], Rn[9];
    for (int i = 0; i < 3; ++i) ctr[i] = -(Rg[i] * xg[0] + Rg[3 + i] * xg[1] + Rg[6 + i] * xg[2]) + 0.01 * rng.normal();
    for (int i = 0; i < 3; ++i) w[i] = xg[3 + i] + 0.5 * M_PI / 180.0 * rng.normal();
    so3exp_d(w, Rn);
    for (int i = 0; i < 3; ++i) cam0[(size_t)6 * c + i] = -(Rn[i * 3] * ctr[0] + Rn[i * 3 + 1] * ctr[1] + Rn[i * 3 + 2] * ctr[2]);
    for (int i = 0; i < 3; ++i) cam0[(size_t)6 * c + 3 + i] = w[i];
  }
  for (size_t i = 0; i < pts0.size(); ++i) pts0[i] += 0.02 * rng.normal();
  const double intr[4] = {fx, fy, cx, cy};
  return gbp_bal_from_arrays(C, L, (uint32_t)E, intr, ci.data(), li.data(), zz.data(), cam0.data(), pts0.data(), out);
}

}  // extern "C"
