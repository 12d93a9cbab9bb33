// Host-side graph partitioning for the multi-GPU path (pure host code, no CUDA):
// contiguous camera ranges balanced by edge count, the rank-local sub-problem and
// the boundary-landmark lists of include/gbp_cuda.h.
//
// Replaces the reference's static placement of cameras / landmarks / factors on the
// tiles of 2^k IPUs (ba/ba.cpp:617-631,717-753,795-834); the exchange Poplar compiles
// from graph.connect + reduce becomes one explicit all-gather of boundary partials.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gbp_cuda.h"

void gbp_set_error(const std::string& s);  // host_error.cpp

struct gbp_shard {
  gbp_shard_plan plan;
  gbp_problem prob;
  std::vector<uint32_t> cam_bounds;  // [world+1]
  std::vector<uint32_t> cam_ids, lmk_ids, active, cam_wflag, lmk_wflag;
  std::vector<float> z, var, cam_pe, cam_pl, lmk_pe, lmk_pl, cam_sc, lmk_sc, damping, mu, oldmu;
  std::vector<int32_t> dcount;
  std::vector<uint32_t> lmk_global, edge_global, bnd_local, bnd_slot;
  uint32_t n_active_global = 0;
};

namespace {

// rank r owns the cameras whose cumulative edge count first reaches r/world of the total
int camera_bounds(const gbp_problem* p, uint32_t world, std::vector<uint32_t>& bounds) {
  const uint32_t C = p->n_keyframes, L = p->n_points, E = p->n_edges;
  std::vector<uint64_t> deg(C, 0);
  for (uint32_t e = 0; e < E; ++e) {
    if (p->cam_ids[e] >= C || p->lmk_ids[e] >= L) {
      gbp_set_error("edge index out of range");
      return GBP_ERR_ARG;
    }
    deg[p->cam_ids[e]]++;
  }
  bounds.assign(world + 1, C);
  bounds[0] = 0;
  uint64_t cum = 0;
  uint32_t r = 1;
  for (uint32_t c = 0; c < C && r < world; ++c) {
    cum += deg[c];
    while (r < world && cum * world >= (uint64_t)r * E && cum > 0) bounds[r++] = c + 1;
  }
  for (uint32_t i = 1; i <= world; ++i) bounds[i] = std::max(bounds[i], bounds[i - 1]);
  bounds[world] = C;
  return GBP_OK;
}

inline uint32_t rank_of_camera(const std::vector<uint32_t>& bounds, uint32_t c) {
  return (uint32_t)(std::upper_bound(bounds.begin() + 1, bounds.end(), c) - bounds.begin() - 1);
}

// first / last rank observing each landmark (0xffffffff = unobserved)
void landmark_rank_span(const gbp_problem* p, const std::vector<uint32_t>& bounds, std::vector<uint32_t>& lo,
                        std::vector<uint32_t>& hi) {
  lo.assign(p->n_points, 0xffffffffu);
  hi.assign(p->n_points, 0u);
  for (uint32_t e = 0; e < p->n_edges; ++e) {
    const uint32_t r = rank_of_camera(bounds, p->cam_ids[e]);
    const uint32_t l = p->lmk_ids[e];
    lo[l] = std::min(lo[l], r);
    hi[l] = std::max(hi[l], r);
  }
}

template <class T>
void slice(std::vector<T>& dst, const T* src, const std::vector<uint32_t>& idx, size_t width) {
  dst.clear();
  if (!src) return;
  dst.resize(idx.size() * width);
  for (size_t i = 0; i < idx.size(); ++i) std::memcpy(&dst[i * width], src + (size_t)idx[i] * width, width * sizeof(T));
}

}  // namespace

extern "C" {

int gbp_cuda_plan_shard(const gbp_problem* p, uint32_t world, uint32_t rank, gbp_shard_plan* out) {
  if (!p || !out || world == 0 || rank >= world) {
    gbp_set_error("bad shard arguments");
    return GBP_ERR_ARG;
  }
  std::vector<uint32_t> bounds, lo, hi;
  int rc = camera_bounds(p, world, bounds);
  if (rc) return rc;
  landmark_rank_span(p, bounds, lo, hi);
  uint32_t n_local_edges = 0, n_local_points = 0, n_boundary = 0;
  std::vector<uint8_t> touched(p->n_points, 0);
  for (uint32_t e = 0; e < p->n_edges; ++e)
    if (rank_of_camera(bounds, p->cam_ids[e]) == rank) {
      n_local_edges++;
      touched[p->lmk_ids[e]] = 1;
    }
  for (uint32_t l = 0; l < p->n_points; ++l) {
    n_local_points += touched[l];
    if (lo[l] != 0xffffffffu && lo[l] != hi[l]) n_boundary++;
  }
  out->world = world;
  out->rank = rank;
  out->cam_begin = bounds[rank];
  out->cam_end = bounds[rank + 1];
  out->n_local_edges = n_local_edges;
  out->n_local_points = n_local_points;
  out->n_boundary_points = n_boundary;
  out->reserved = 0;
  return GBP_OK;
}

int gbp_shard_build(const gbp_problem* p, uint32_t world, uint32_t rank, gbp_shard** out) {
  if (!p || !out || world == 0 || rank >= world) {
    gbp_set_error("bad shard arguments");
    return GBP_ERR_ARG;
  }
  if ((p->n_edges && (!p->cam_ids || !p->lmk_ids || !p->measurements || !p->meas_variances)) ||
      (p->n_keyframes && (!p->cam_priors_eta || !p->cam_priors_lambda || !p->cam_scaling || !p->cam_weaken_flag)) ||
      (p->n_points && (!p->lmk_priors_eta || !p->lmk_priors_lambda || !p->lmk_scaling || !p->lmk_weaken_flag))) {
    gbp_set_error("gbp_problem has a null required array");
    return GBP_ERR_ARG;
  }
  gbp_shard* s = new gbp_shard();
  int rc = camera_bounds(p, world, s->cam_bounds);
  if (rc) {
    delete s;
    return rc;
  }
  const uint32_t L = p->n_points, E = p->n_edges;
  const uint32_t c0 = s->cam_bounds[rank], c1 = s->cam_bounds[rank + 1];
  std::vector<uint32_t> lo, hi;
  landmark_rank_span(p, s->cam_bounds, lo, hi);
  // local edges (global order), local landmarks (ascending global id)
  std::vector<uint32_t> lmk_local(L, 0xffffffffu);
  for (uint32_t e = 0; e < E; ++e) {
    const uint32_t c = p->cam_ids[e];
    if (c >= c0 && c < c1) {
      s->edge_global.push_back(e);
      lmk_local[p->lmk_ids[e]] = 0;
    }
    if (!p->active_flag || p->active_flag[e] == 1u) s->n_active_global++;
  }
  uint32_t n_boundary = 0;
  for (uint32_t l = 0; l < L; ++l) {
    const bool boundary = lo[l] != 0xffffffffu && lo[l] != hi[l];
    if (lmk_local[l] == 0) {
      lmk_local[l] = (uint32_t)s->lmk_global.size();
      s->lmk_global.push_back(l);
      if (boundary) {
        s->bnd_local.push_back(lmk_local[l]);
        s->bnd_slot.push_back(n_boundary);
      }
    }
    if (boundary) n_boundary++;
  }
  const uint32_t nE = (uint32_t)s->edge_global.size(), nL = (uint32_t)s->lmk_global.size(), nC = c1 - c0;
  s->cam_ids.resize(nE);
  s->lmk_ids.resize(nE);
  for (uint32_t i = 0; i < nE; ++i) {
    const uint32_t e = s->edge_global[i];
    s->cam_ids[i] = p->cam_ids[e] - c0;
    s->lmk_ids[i] = lmk_local[p->lmk_ids[e]];
  }
  std::vector<uint32_t> cams(nC);
  for (uint32_t i = 0; i < nC; ++i) cams[i] = c0 + i;
  slice(s->z, p->measurements, s->edge_global, 2);
  slice(s->var, p->meas_variances, s->edge_global, 1);
  slice(s->active, p->active_flag, s->edge_global, 1);
  slice(s->damping, p->damping, s->edge_global, 1);
  slice(s->dcount, p->damping_count, s->edge_global, 1);
  slice(s->mu, p->mu, s->edge_global, 9);
  slice(s->oldmu, p->oldmu, s->edge_global, 9);
  slice(s->cam_pe, p->cam_priors_eta, cams, 6);
  slice(s->cam_pl, p->cam_priors_lambda, cams, 36);
  slice(s->cam_sc, p->cam_scaling, cams, 1);
  slice(s->cam_wflag, p->cam_weaken_flag, cams, 1);
  slice(s->lmk_pe, p->lmk_priors_eta, s->lmk_global, 3);
  slice(s->lmk_pl, p->lmk_priors_lambda, s->lmk_global, 9);
  slice(s->lmk_sc, p->lmk_scaling, s->lmk_global, 1);
  slice(s->lmk_wflag, p->lmk_weaken_flag, s->lmk_global, 1);
  gbp_problem& q = s->prob;
  std::memset(&q, 0, sizeof(q));
  q.n_keyframes = nC;
  q.n_points = nL;
  q.n_edges = nE;
  std::memcpy(q.K, p->K, sizeof(q.K));
  auto ptr = [](auto& v) { return v.empty() ? nullptr : v.data(); };
  q.cam_ids = s->cam_ids.data();
  q.lmk_ids = s->lmk_ids.data();
  q.measurements = s->z.data();
  q.meas_variances = s->var.data();
  q.cam_priors_eta = s->cam_pe.data();
  q.cam_priors_lambda = s->cam_pl.data();
  q.lmk_priors_eta = s->lmk_pe.data();
  q.lmk_priors_lambda = s->lmk_pl.data();
  q.cam_scaling = s->cam_sc.data();
  q.lmk_scaling = s->lmk_sc.data();
  q.cam_weaken_flag = s->cam_wflag.data();
  q.lmk_weaken_flag = s->lmk_wflag.data();
  q.active_flag = ptr(s->active);
  q.damping = ptr(s->damping);
  q.damping_count = ptr(s->dcount);
  q.mu = ptr(s->mu);
  q.oldmu = ptr(s->oldmu);
  s->plan.world = world;
  s->plan.rank = rank;
  s->plan.cam_begin = c0;
  s->plan.cam_end = c1;
  s->plan.n_local_edges = nE;
  s->plan.n_local_points = nL;
  s->plan.n_boundary_points = n_boundary;
  s->plan.reserved = 0;
  *out = s;
  return GBP_OK;
}

void gbp_shard_free(gbp_shard* s) { delete s; }
const gbp_problem* gbp_shard_problem(const gbp_shard* s) { return s ? &s->prob : nullptr; }
const gbp_shard_plan* gbp_shard_get_plan(const gbp_shard* s) { return s ? &s->plan : nullptr; }
const uint32_t* gbp_shard_lmk_global(const gbp_shard* s) { return s ? s->lmk_global.data() : nullptr; }
const uint32_t* gbp_shard_edge_global(const gbp_shard* s) { return s ? s->edge_global.data() : nullptr; }
uint32_t gbp_shard_n_boundary_local(const gbp_shard* s) { return s ? (uint32_t)s->bnd_local.size() : 0; }
const uint32_t* gbp_shard_boundary_local(const gbp_shard* s) { return s ? s->bnd_local.data() : nullptr; }
const uint32_t* gbp_shard_boundary_slot(const gbp_shard* s) { return s ? s->bnd_slot.data() : nullptr; }
uint32_t gbp_shard_n_active_global(const gbp_shard* s) { return s ? s->n_active_global : 0; }
const uint32_t* gbp_shard_cam_bounds(const gbp_shard* s) { return s ? s->cam_bounds.data() : nullptr; }

}  // extern "C"
