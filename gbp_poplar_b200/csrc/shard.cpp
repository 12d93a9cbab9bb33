// Host-side graph partitioning for the multi-GPU path (pure host code, no CUDA):
// contiguous camera ranges balanced by warp-tile count, the rank-local sub-problem and
// the boundary-landmark lists of include/gbp_cuda.h.
//
// Replaces the reference's static placement of cameras / landmarks / factors on the
// tiles of 2^k IPUs (ba/ba.cpp:617-631,717-753,795-834); the exchange Poplar compiles
// from graph.connect + reduce becomes one explicit all-gather of boundary partials.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
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
  std::vector<uint32_t> lmk_global, edge_global, bnd_local, bnd_slot, bnd_span;
  uint32_t n_active_global = 0;
};

namespace {

// rank r owns the cameras whose cumulative warp-tile count first reaches r/world of the total
int camera_bounds(const gbp_problem* p, uint32_t world, std::vector<uint32_t>& bounds) {
  const uint32_t C = p->n_keyframes, L = p->n_points, E = p->n_edges;
  std::vector<uint64_t> deg(C, 0);
  for (uint32_t e = 0; e < E;) {  // by runs of one camera (the shipped files are camera-sorted)
    const uint32_t c = p->cam_ids[e];
    if (c >= C) {
      gbp_set_error("edge index out of range");
      return GBP_ERR_ARG;
    }
    const uint32_t e0 = e;
    uint32_t bad = 0;
    while (e < E && p->cam_ids[e] == c) bad |= (p->lmk_ids[e++] >= L) ? 1u : 0u;
    if (bad) {
      gbp_set_error("edge index out of range");
      return GBP_ERR_ARG;
    }
    deg[c] += e - e0;
  }
  // Balanced by WARP-TILES (a camera's factors are padded to whole 32-slot tiles and a tile is the unit of work of the
  // sweep kernel), not by edges: the ranks advance in lock step, so the job runs at the pace of the rank with the
  // most tiles.
  bounds.assign(world + 1, C);
  bounds[0] = 0;
  uint64_t total = 0;
  for (uint32_t c = 0; c < C; ++c) total += (deg[c] + 31) / 32;
  uint64_t cum = 0;
  uint32_t r = 1;
  for (uint32_t c = 0; c < C && r < world; ++c) {
    cum += (deg[c] + 31) / 32;
    while (r < world && cum * world >= (uint64_t)r * total && cum > 0) bounds[r++] = c + 1;
  }
  for (uint32_t i = 1; i <= world; ++i) bounds[i] = std::max(bounds[i], bounds[i - 1]);
  bounds[world] = C;
  return GBP_OK;
}

// owner rank of every camera (a table: one lookup per edge instead of a binary search)
std::vector<uint32_t> camera_ranks(const std::vector<uint32_t>& bounds, uint32_t C) {
  std::vector<uint32_t> rank(C, 0u);
  for (uint32_t r = 0; r + 1 < bounds.size(); ++r)
    for (uint32_t c = bounds[r]; c < bounds[r + 1] && c < C; ++c) rank[c] = r;
  return rank;
}

// first / last rank observing each landmark (0xffffffff = unobserved)
void landmark_rank_span(const gbp_problem* p, const std::vector<uint32_t>& cam_rank, std::vector<uint32_t>& lo,
                        std::vector<uint32_t>& hi) {
  lo.assign(p->n_points, 0xffffffffu);
  hi.assign(p->n_points, 0u);
  for (uint32_t e = 0; e < p->n_edges; ++e) {
    const uint32_t r = cam_rank[p->cam_ids[e]];
    const uint32_t l = p->lmk_ids[e];
    lo[l] = std::min(lo[l], r);
    hi[l] = std::max(hi[l], r);
  }
}

// sel = src[idx] (rows of `width` elements), returned as a pointer: into `dst` (a copy) or -- when `view` is set
// and idx is one contiguous ascending run (the edges of a camera range in a camera-sorted file, a range of
// cameras) -- straight into src, without touching memory.  skip_if_zero: nullptr when every selected element is
// zero (the streamed mu / oldmu of a fresh problem: the library treats a null array as zeros).
template <class T>
const T* slice(std::vector<T>& dst, const T* src, const std::vector<uint32_t>& idx, size_t width, bool view,
               bool skip_if_zero = false) {
  dst.clear();
  if (!src || idx.empty()) return nullptr;
  const bool contiguous = (size_t)idx.back() - (size_t)idx.front() + 1 == idx.size();
  if (skip_if_zero) {
    bool any = false;
    if (contiguous) {
      // block-wise, branch-free inside a block so the compiler vectorises it (-0.0f counts as non-zero: harmless)
      const unsigned char* b = reinterpret_cast<const unsigned char*>(src + (size_t)idx.front() * width);
      const size_t n = idx.size() * width * sizeof(T);
      for (size_t i = 0; i < n && !any; i += 4096) {
        unsigned char acc = 0;
        for (size_t k = i, k1 = std::min(n, i + 4096); k < k1; ++k) acc |= b[k];
        any = acc != 0;
      }
    } else {
      for (size_t i = 0; i < idx.size() && !any; ++i)
        for (size_t k = 0; k < width && !any; ++k) any = src[(size_t)idx[i] * width + k] != T(0);
    }
    if (!any) return nullptr;
  }
  if (contiguous) {
    const T* b = src + (size_t)idx.front() * width;
    if (view) return b;
    dst.assign(b, b + idx.size() * width);
    return dst.data();
  }
  dst.resize(idx.size() * width);
  for (size_t i = 0; i < idx.size(); ++i) std::memcpy(&dst[i * width], src + (size_t)idx[i] * width, width * sizeof(T));
  return dst.data();
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
  const std::vector<uint32_t> cam_rank = camera_ranks(bounds, p->n_keyframes);
  landmark_rank_span(p, cam_rank, lo, hi);
  uint32_t n_local_edges = 0, n_local_points = 0, n_boundary = 0;
  std::vector<uint8_t> touched(p->n_points, 0);
  for (uint32_t e = 0; e < p->n_edges; ++e)
    if (cam_rank[p->cam_ids[e]] == rank) {
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

static int shard_build_impl(const gbp_problem* p, uint32_t world, uint32_t rank, bool view, gbp_shard** out) {
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
  if (world > 0xffffu) {
    gbp_set_error("too many ranks");
    return GBP_ERR_ARG;
  }
  gbp_shard* s = new gbp_shard();
  const bool timing = std::getenv("GBP_INIT_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[gbp shard] %-30s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  int rc = camera_bounds(p, world, s->cam_bounds);
  lap("camera bounds");
  if (rc) {
    delete s;
    return rc;
  }
  const uint32_t L = p->n_points, E = p->n_edges;
  const uint32_t c0 = s->cam_bounds[rank], c1 = s->cam_bounds[rank + 1];
  // one pass over the global edge list: first / last rank of every landmark, this rank's edges (global
  // order) and landmarks, the active-edge count of the whole graph
  const std::vector<uint32_t> cam_rank = camera_ranks(s->cam_bounds, p->n_keyframes);
  std::vector<uint32_t> lo(L, 0xffffffffu), hi(L, 0u);
  std::vector<uint32_t> rank_mask(world <= 32 ? L : 0, 0u);  // bit r: rank r observes the landmark (worlds of up to 32 ranks)
  std::vector<uint32_t> lmk_local(L, 0xffffffffu);
  s->edge_global.reserve((size_t)E / world + (size_t)E / (8 * world) + 64);
  for (uint32_t e = 0; e < E; ++e) {
    const uint32_t r = cam_rank[p->cam_ids[e]];
    const uint32_t l = p->lmk_ids[e];
    lo[l] = std::min(lo[l], r);
    hi[l] = std::max(hi[l], r);
    if (!rank_mask.empty()) rank_mask[l] |= 1u << r;
    if (r == rank) {
      s->edge_global.push_back(e);
      lmk_local[l] = 0;
    }
    if (!p->active_flag || p->active_flag[e] == 1u) s->n_active_global++;
  }
  lap("edge pass");
  uint32_t n_boundary = 0;
  for (uint32_t l = 0; l < L; ++l) {
    const bool boundary = lo[l] != 0xffffffffu && lo[l] != hi[l];
    if (lmk_local[l] == 0) {
      lmk_local[l] = (uint32_t)s->lmk_global.size();
      s->lmk_global.push_back(l);
      if (boundary) {
        s->bnd_local.push_back(lmk_local[l]);
        s->bnd_slot.push_back(n_boundary);
        s->bnd_span.push_back(rank_mask.empty() ? 0u : rank_mask[l]);  // the ranks that contribute a partial sum to it
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
  lap("landmark pass + local ids");
  std::vector<uint32_t> cams(nC);
  for (uint32_t i = 0; i < nC; ++i) cams[i] = c0 + i;
  gbp_problem& q = s->prob;
  std::memset(&q, 0, sizeof(q));
  q.n_keyframes = nC;
  q.n_points = nL;
  q.n_edges = nE;
  std::memcpy(q.K, p->K, sizeof(q.K));
  q.cam_ids = s->cam_ids.data();
  q.lmk_ids = s->lmk_ids.data();
  q.measurements = slice(s->z, p->measurements, s->edge_global, 2, view);
  q.meas_variances = slice(s->var, p->meas_variances, s->edge_global, 1, view);
  q.active_flag = slice(s->active, p->active_flag, s->edge_global, 1, view);
  q.damping = slice(s->damping, p->damping, s->edge_global, 1, view);
  q.damping_count = slice(s->dcount, p->damping_count, s->edge_global, 1, view);
  lap("slices: z var active damping dcount");
  q.mu = slice(s->mu, p->mu, s->edge_global, 9, view, true);
  q.oldmu = slice(s->oldmu, p->oldmu, s->edge_global, 9, view, true);
  lap("slices: mu oldmu");
  q.cam_priors_eta = slice(s->cam_pe, p->cam_priors_eta, cams, 6, view);
  q.cam_priors_lambda = slice(s->cam_pl, p->cam_priors_lambda, cams, 36, view);
  q.cam_scaling = slice(s->cam_sc, p->cam_scaling, cams, 1, view);
  q.cam_weaken_flag = slice(s->cam_wflag, p->cam_weaken_flag, cams, 1, view);
  q.lmk_priors_eta = slice(s->lmk_pe, p->lmk_priors_eta, s->lmk_global, 3, view);
  q.lmk_priors_lambda = slice(s->lmk_pl, p->lmk_priors_lambda, s->lmk_global, 9, view);
  q.lmk_scaling = slice(s->lmk_sc, p->lmk_scaling, s->lmk_global, 1, view);
  q.lmk_weaken_flag = slice(s->lmk_wflag, p->lmk_weaken_flag, s->lmk_global, 1, view);
  lap("slices: variables");
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

int gbp_shard_build(const gbp_problem* p, uint32_t world, uint32_t rank, gbp_shard** out) {
  return shard_build_impl(p, world, rank, false, out);
}
int gbp_shard_build_view(const gbp_problem* p, uint32_t world, uint32_t rank, gbp_shard** out) {
  return shard_build_impl(p, world, rank, true, out);
}

// After gbp_shard_build_view: drop every member of the local problem that points INTO the caller's arrays (the
// per-edge / per-camera runs that were not copied).  What stays is owned by the shard: the index maps, the plan,
// the boundary lists and whatever had to be gathered.  gbp_cuda_init_shard / gbp_cuda_init_group call this once the
// device copy exists, so the shard they retain (gbp_cuda_shard_info) never dangles.
void gbp_shard_detach_views(gbp_shard* s) {
  if (!s) return;
  gbp_problem& q = s->prob;
  auto own = [](const void* ptr, const void* vec_data) { return ptr != nullptr && ptr == vec_data; };
#define GBP_DETACH(member, vec) if (!own(q.member, s->vec.data())) q.member = nullptr
  GBP_DETACH(measurements, z);
  GBP_DETACH(meas_variances, var);
  GBP_DETACH(active_flag, active);
  GBP_DETACH(damping, damping);
  GBP_DETACH(damping_count, dcount);
  GBP_DETACH(mu, mu);
  GBP_DETACH(oldmu, oldmu);
  GBP_DETACH(cam_priors_eta, cam_pe);
  GBP_DETACH(cam_priors_lambda, cam_pl);
  GBP_DETACH(cam_scaling, cam_sc);
  GBP_DETACH(cam_weaken_flag, cam_wflag);
  GBP_DETACH(lmk_priors_eta, lmk_pe);
  GBP_DETACH(lmk_priors_lambda, lmk_pl);
  GBP_DETACH(lmk_scaling, lmk_sc);
  GBP_DETACH(lmk_weaken_flag, lmk_wflag);
#undef GBP_DETACH
}

void gbp_shard_free(gbp_shard* s) { delete s; }
const gbp_problem* gbp_shard_problem(const gbp_shard* s) { return s ? &s->prob : nullptr; }
const gbp_shard_plan* gbp_shard_get_plan(const gbp_shard* s) { return s ? &s->plan : nullptr; }
const uint32_t* gbp_shard_lmk_global(const gbp_shard* s) { return s ? s->lmk_global.data() : nullptr; }
const uint32_t* gbp_shard_edge_global(const gbp_shard* s) { return s ? s->edge_global.data() : nullptr; }
uint32_t gbp_shard_n_boundary_local(const gbp_shard* s) { return s ? (uint32_t)s->bnd_local.size() : 0; }
const uint32_t* gbp_shard_boundary_local(const gbp_shard* s) { return s ? s->bnd_local.data() : nullptr; }
const uint32_t* gbp_shard_boundary_slot(const gbp_shard* s) { return s ? s->bnd_slot.data() : nullptr; }
const uint32_t* gbp_shard_boundary_ranks(const gbp_shard* s) { return s ? s->bnd_span.data() : nullptr; }
uint32_t gbp_shard_n_active_global(const gbp_shard* s) { return s ? s->n_active_global : 0; }
const uint32_t* gbp_shard_cam_bounds(const gbp_shard* s) { return s ? s->cam_bounds.data() : nullptr; }

}  // extern "C"
