// C ABI of the GBP hot path (include/gbp_cuda.h): handle, device layout
// construction, program entry points and tensor (de)serialisation.
//
// What used to be Poplar graph construction -- tensors mapped onto IPU tiles and
// vertices wired to tensor slices (ba/ba.cpp:45-371, 658-937) -- becomes: build
// the edge-slot layout (gbp_layout.h), upload, and launch the kernels of
// gbp_kernels.cuh on one stream.  There is no CPU path in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gbp_cuda.h"
#include "gbp_kernels.cuh"
#include "gbp_setup.h"
#include "gbp_fast.h"
#include "nccl_dyn.h"


void gbp_set_error(const std::string& s);  // host_error.cpp
extern "C" void gbp_shard_detach_views(gbp_shard* s);  // shard.cpp

namespace {

using gbp::DeviceGraph;

#define GBP_CUDA_TRY(expr)                                                              \
  do {                                                                                  \
    cudaError_t err__ = (expr);                                                         \
    if (err__ != cudaSuccess) {                                                         \
      gbp_set_error(std::string(#expr) + ": " + cudaGetErrorString(err__));             \
      return GBP_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

template <class T>
int dev_alloc(T** p, size_t n, bool zero = true) {
  *p = nullptr;
  if (n == 0) n = 1;
  GBP_CUDA_TRY(cudaMalloc((void**)p, n * sizeof(T)));
  if (zero) GBP_CUDA_TRY(cudaMemset(*p, 0, n * sizeof(T)));
  return GBP_OK;
}

inline uint32_t f2u(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}
inline int32_t f2i(float f) {
  int32_t i;
  std::memcpy(&i, &f, 4);
  return i;
}

}  // namespace

struct gbp_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  DeviceGraph g;
  uint32_t C = 0, L = 0, E = 0, E_pad = 0, n_tiles = 0, SK = 1, SL = 1;
  // host-side index maps (reference edge order <-> edge slots)
  std::vector<uint32_t> cam_ids, lmk_ids, slot_c, slot_l, pos_of_orig, active_host, lmk_ptr;
  uint32_t n_active = 0;
  std::vector<float> mu_init;     // host copy of the streamed `mu` (empty = zeros)
  std::vector<float> oldmu_init;  // host copy of the streamed `oldmu` (empty = zeros)
  bool pending_shift = false;     // a PrepMessageVertex pass ran since the last belief update
  // Slot 0 of the reference's p*_messages tensors holds the priors as they were at the
  // last Copy(messages, pmessages) (ba.cpp:902-905).  Nothing ever reads it, so it is
  // tracked lazily: snapshotted only when the priors are about to change.
  bool p_in_sync = false;         // a sweep ran since the priors last changed
  float* d_pprior_cam_eta = nullptr;
  float* d_pprior_cam_lam = nullptr;
  float4* d_pprior_lmk = nullptr;
  int use_graph = 0;
  cudaAccessPolicyWindow l2_window{};  // num_bytes == 0: none
  // CUDA-graph replay of one steady-state sweep (with / without the metric): one launch per sweep
  // instead of 2-5, which is what bounds the small graphs of the reference sequences
  cudaGraphExec_t sweep_graph[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [skip_upper * 4 + two_pass * 2 + with_stats]
  int skip_upper = 1;               // all sweeps of a gbp_cuda_iterate call but the last skip the upper triangle of the camera messages (GBP_SKIP_UPPER=0 disables)
  DeviceGraph graph_g;               // the kernel arguments the graphs were captured with
  gbp::DeviceStats* graph_stats = nullptr;
  uint32_t graph_n_active = 0;      // k_metric's argument at capture time
  bool capturing = false;
  uint32_t* d_stat_cursor = nullptr;
  // sweep flavour: 0 = one fused kernel (prep + messages), 1 = prep pass + compacted relinearisation +
  // message-only kernel.  Bit-identical; chosen from the relinearisation pattern of the last sweeps.
  int relin_mode = 0;         // gbp_opts.relin_mode: 0 auto, 1 fused, 2 two-pass
  int two_pass = 0;
  uint32_t sweeps_since_choice = 0;
  int num_sms = 148;
  // staging for READ_PROG
  uint32_t* d_pos_of_orig = nullptr;
  uint32_t* d_cam_ids = nullptr;   // [E] | [E] the caller's camera / landmark ids, kept for ensure_host_maps
  uint32_t* d_lmk_ids = nullptr;
  bool host_maps_ready = false;
  // device-side keyframe insertion (gbp_cuda_add_keyframe_device)
  uint32_t* d_lmk_first_cam = nullptr;      // [L] lowest camera index observing the landmark (0xffffffff: none)
  float* d_kf_scratch = nullptr;            // [4] see k_kf_pose
  std::vector<uint32_t> new_lmks_at_cam;    // [C] landmarks first observed by each camera
  std::vector<uint32_t> cam_edge_ptr;       // [C+1] CSR over original edge ids: the factors of every camera
  std::vector<uint32_t> cam_edge_ids;       // [E]
  float* d_exp_lmk_eta = nullptr;
  float* d_exp_lmk_lam = nullptr;
  float* d_exp_damping = nullptr;
  int32_t* d_exp_dcount = nullptr;
  uint32_t* d_exp_robust = nullptr;
  // metric
  gbp::MetricPartial* d_metric_parts = nullptr;
  uint32_t* d_metric_ticket = nullptr;  // blocks of k_metric that are done (the last one finishes)
  double* d_met_cam = nullptr;      // [C][16] double-precision camera means + rotations (metric only)
  double* d_met_lmk = nullptr;      // [L][4]  double-precision landmark means (metric only)
  gbp::DeviceStats* d_stats = nullptr;
  size_t d_stats_cap = 0;
  // pinned staging for the small per-call read-backs (per-sweep metrics, relinearisation ring): one
  // asynchronous copy each and ONE stream synchronisation per gbp_cuda_iterate call
  void* pin_block = nullptr;   // page-locked staging block (pinned_get): error word | ring | stats
  size_t pin_block_bytes = 0;
  gbp_iter_stats* pin_stats = nullptr;
  size_t pin_stats_cap = 0;
  uint32_t* pin_ring = nullptr;
  bool d_stats_in_arena = false;
  // timing
  int profile = 0;
  std::vector<cudaEvent_t> prof_events;
  float last_ms_factor = 0.f, last_ms_variable = 0.f;
  std::vector<float> sweep_ms_factor, sweep_ms_variable;  // per sweep of the last profiled gbp_cuda_iterate
  float last_ms = 0.f;
  uint64_t kernels_launched = 0;
  uint64_t last_kernels = 0;
  std::vector<void*> allocs;
  std::vector<void*> pool_allocs;  // from the stream-ordered pool (arena_alloc)
  // multi-GPU shard (null / 0 on a single-GPU handle)
  gbp_shard* shard = nullptr;
  uint32_t world = 1, rank = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_send = nullptr, ev_recv = nullptr;
  // peer-to-peer exchange (CUDA IPC): one exported block per rank [flags | counters | receive buffer]
  int p2p = 0;                      // 1 = boundary partials are pushed straight into the peers' buffers
  void* p2p_block = nullptr;        // this rank's exported block
  std::vector<void*> p2p_peers;     // the other ranks' blocks, mapped (nullptr for this rank)
  double* d_metric_raw = nullptr;   // [8]         this rank's metric sums
  double* d_metric_all = nullptr;   // [world][8]  all-gathered
  uint64_t exchanges = 0;
  uint32_t* p2p_err_host = nullptr; // host-mapped word a kernel sets when a wait for a peer timed out (sticky)
  struct GroupShared* group = nullptr;  // single-process group (gbp_cuda_init_group): owner of the exchange blocks
  // state of the gbp_cuda_iterate call in flight (iterate_begin .. iterate_finish)
  cudaGraphExec_t it_exec = nullptr, it_exec_lower = nullptr;
  bool it_graph = false, it_prof = false;
  uint64_t it_k0 = 0;
  bool l2_limit_changed = false;
  // TMA-staged sweep kernel (k_sweep_tma): descriptors of the two big quad-SoA arrays
  gbp::SweepMaps maps;
  int use_tma = 0;
  int fast_math = 0;  // gbp_opts.fast_math: the contracted-FMA build of the sweep kernel (gbp_fast.cu)
  // GBP_XCHG_PUSH=1 (read when the handle is built): boundary partial sums are stored into the observers' buffers
  // instead of read from the owners' (see boundary_push); every rank of a job must use the same setting
  int xchg_push = std::getenv("GBP_XCHG_PUSH") ? (std::atoi(std::getenv("GBP_XCHG_PUSH")) != 0) : 0;
};

// A single-process group of shard handles: the exchange blocks of all ranks live (and die) together, so that
// freeing one handle never pulls memory from under a peer that is still pushing into it.
struct GroupShared {
  std::vector<void*> blocks;
  std::vector<int> devices;
  int refs = 0;
};

namespace {

// The per-handle arena comes from a stream-ordered memory pool PRIVATE to this library (one per device, release
// threshold lifted), so the block of a freed handle is handed to the next gbp_cuda_init of the process without a
// trip to the driver (a fresh cudaMalloc of 0.65 GB costs anything between 1 and 200 ms depending on the box) and
// the host application's own default pool is left alone.  gbp_cuda_release_cached_memory() returns it.
struct DevicePool {
  int device;
  cudaMemPool_t pool;
};
std::vector<DevicePool>& pools() {
  static std::vector<DevicePool> v;
  return v;
}
std::mutex& pools_mutex() {
  static std::mutex m;
  return m;
}
cudaMemPool_t private_pool(int device) {
  std::lock_guard<std::mutex> lk(pools_mutex());
  for (const DevicePool& d : pools())
    if (d.device == device) return d.pool;
  int supported = 0;
  cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, device);
  if (const char* env = std::getenv("GBP_ARENA_POOL")) supported = supported && std::atoi(env) != 0;
  cudaMemPool_t pool = nullptr;
  if (supported) {
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
      pool = nullptr;
      cudaGetLastError();
    }
  }
  pools().push_back({device, pool});
  return pool;
}

int arena_alloc(gbp_handle* h, char** p, size_t bytes) {
  *p = nullptr;
  if (cudaMemPool_t pool = private_pool(h->device)) {
    if (cudaMallocFromPoolAsync((void**)p, std::max<size_t>(bytes, 16), pool, h->stream) == cudaSuccess) {
      h->pool_allocs.push_back((void*)*p);
      return GBP_OK;
    }
    cudaGetLastError();
  }
  GBP_CUDA_TRY(cudaMalloc((void**)p, std::max<size_t>(bytes, 16)));
  h->allocs.push_back((void*)*p);
  return GBP_OK;
}

// Persisting-L2 set-aside: a context-wide limit.  The first handle that raises it remembers the previous value;
// the last handle with a window to go restores it (and only then resets the persisting lines).
struct L2State {
  int device;
  int users;
  size_t prev_limit;
};
std::vector<L2State>& l2_states() {
  static std::vector<L2State> v;
  return v;
}
L2State& l2_state(int device) {
  for (L2State& s : l2_states())
    if (s.device == device) return s;
  l2_states().push_back({device, 0, 0});
  return l2_states().back();
}

// Page-locked host memory is expensive to get and to give back on some boxes (cudaHostAlloc / cudaFreeHost of a few
// KB: 10-150 ms measured), so the small staging blocks of the handles are recycled process-wide: a freed handle's
// block is handed to the next gbp_cuda_init.  All blocks are mapped + portable, so any of them can serve as the
// device-visible error word of a sharded handle.  gbp_cuda_release_cached_memory() frees them.
struct PinnedBlock {
  void* p;
  size_t bytes;
};
std::vector<PinnedBlock>& pinned_free_list() {
  static std::vector<PinnedBlock> v;
  return v;
}
void* pinned_get(size_t bytes, size_t* got) {
  {
    std::lock_guard<std::mutex> lk(pools_mutex());
    std::vector<PinnedBlock>& fl = pinned_free_list();
    for (size_t i = 0; i < fl.size(); ++i)
      if (fl[i].bytes >= bytes) {
        PinnedBlock b = fl[i];
        fl.erase(fl.begin() + (long)i);
        *got = b.bytes;
        return b.p;
      }
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  *got = bytes;
  return p;
}
void pinned_put(void* p, size_t bytes) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(pools_mutex());
  pinned_free_list().push_back({p, bytes});
}
// layout of a handle's staging block: [error word + padding 64 B | relinearisation ring 192 B | per-sweep stats ...]
constexpr size_t PIN_OFF_RING = 64, PIN_OFF_STATS = 256;
constexpr size_t STATS_PREALLOC = 4096;  // sweeps of per-sweep metrics a handle can return without growing its buffers

template <class T>
int h_alloc(gbp_handle* h, T** p, size_t n, bool zero = true) {
  int rc = dev_alloc(p, n, zero);
  if (rc == GBP_OK) h->allocs.push_back((void*)*p);
  return rc;
}

template <class T>
int upload(T* dst, const T* src, size_t n, cudaStream_t s) {
  if (n == 0) return GBP_OK;
  GBP_CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
  return GBP_OK;
}
template <class T>
int download(T* dst, const T* src, size_t n, cudaStream_t s) {
  if (n == 0) return GBP_OK;
  GBP_CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, s));
  return GBP_OK;
}

inline uint32_t metric_grid(const gbp_handle* h) { return std::min<uint32_t>(h->n_tiles, 8u * (uint32_t)h->num_sms); }
inline uint32_t lmks_grid(const gbp_handle* h) { return h->g.n_lmk_blocks; }
inline uint32_t cams_grid(const gbp_handle* h) { return (h->C + GBP_CAM_PER_BLOCK - 1) / GBP_CAM_PER_BLOCK; }

int priors_about_to_change(gbp_handle* h) {
  if (!h->p_in_sync) return GBP_OK;
  GBP_CUDA_TRY(cudaMemcpyAsync(h->d_pprior_cam_eta, h->g.cam_prior_eta, 6 * (size_t)h->C * 4, cudaMemcpyDeviceToDevice, h->stream));
  GBP_CUDA_TRY(cudaMemcpyAsync(h->d_pprior_cam_lam, h->g.cam_prior_lam, 36 * (size_t)h->C * 4, cudaMemcpyDeviceToDevice, h->stream));
  GBP_CUDA_TRY(cudaMemcpyAsync(h->d_pprior_lmk, h->g.lmk_prior, 3 * (size_t)h->L * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
  h->p_in_sync = false;
  return GBP_OK;
}

// NCCL communicators are cached per unique id for the life of the process: building one
// costs ~1 s, and a host program that solves several problems in a row (or the SLAM-style
// re-initialisation of an engine) should pay it once.  Handles only borrow them.
struct CommEntry {
  unsigned char id[128];
  uint32_t world, rank;
  int device;
  ncclComm_t comm;
};
std::vector<CommEntry>& comm_cache() {
  static std::vector<CommEntry> cache;
  return cache;
}

#define GBP_NCCL_TRY(expr)                                                                         \
  do {                                                                                             \
    ncclResult_t r__ = (expr);                                                                     \
    if (r__ != ncclSuccess) {                                                                      \
      gbp_set_error(std::string(#expr) + ": " + gbp::nccl_api().GetErrorString(r__));              \
      return GBP_ERR_COMM;                                                                         \
    }                                                                                              \
  } while (0)

// prog_ub (ba/ba.cpp:104-139).  On a shard the boundary landmarks go first: their partial
// sums are all-gathered on the communication stream while the main stream updates the
// interior landmarks and the cameras; k_boundary_finish then waits for the gather.
int launch_update_vars(gbp_handle* h, bool lower_only_in = false, bool skip_cams = false) {
  static const int uv_debug = std::getenv("GBP_UV_DEBUG") ? std::atoi(std::getenv("GBP_UV_DEBUG")) : 0;  // timing diagnostics: 1 skips the cameras, 2 the landmarks (results are then WRONG)
  // bit 0: mirror the lower triangle of the camera beliefs; bit 1: the cameras are already done (fused into the sweep)
  // where the finish blocks of the boundary exchange sit among the landmark blocks, in percent of the latter (see
  // k_update_vars): 0 = before all of them, 100 = behind all of them
  static const int finish_at = std::getenv("GBP_FINISH_AT") ? std::min(100, std::max(0, std::atoi(std::getenv("GBP_FINISH_AT")))) : GBP_FINISH_AT_DEFAULT;
  // timing diagnostics of the exchange (results are then WRONG): 1 = push blocks idle, 2 = finish blocks idle, 3 = both
  static const int xchg_debug = std::getenv("GBP_XCHG_DEBUG") ? std::atoi(std::getenv("GBP_XCHG_DEBUG")) : 0;
  const int lower_only = (lower_only_in ? 1 : 0) | (skip_cams ? 2 : 0) | (uv_debug << 1) | ((xchg_debug & 3) << 4) | (h->xchg_push ? 0 : 64);
  const int shift = h->pending_shift ? 1 : 0;
  const uint32_t grid = lmks_grid(h);
  const bool exchange = h->shard && h->g.n_bnd_global > 0;
  const uint32_t bgrid = (h->g.n_bnd_local + GBP_TILE - 1) / GBP_TILE;
  if (exchange && h->p2p) {
    // one launch: the first blocks form the partial sums (tagged with the exchange step), later blocks finish the
    // boundary landmarks once the tagged partials of every observing rank are there (see boundary_push / boundary_finish)
    const uint32_t n_x = std::max((h->g.n_bnd_local + GBP_LMK_PER_BLOCK - 1) / GBP_LMK_PER_BLOCK, 1u);
    const uint32_t finish_after = (uint32_t)((uint64_t)grid * finish_at / 100);
    gbp::k_update_vars<<<n_x + cams_grid(h) + grid + n_x, GBP_TILE, 0, h->stream>>>(h->g, shift, n_x, lower_only, finish_after);
    h->kernels_launched++;
    h->exchanges++;
  } else {
    if (exchange) {
      if (bgrid) {
        gbp::k_boundary_partial<<<bgrid, GBP_TILE, 0, h->stream>>>(h->g);
        h->kernels_launched++;
      }
      GBP_CUDA_TRY(cudaEventRecord(h->ev_send, h->stream));
      GBP_CUDA_TRY(cudaStreamWaitEvent(h->comm_stream, h->ev_send, 0));
      GBP_NCCL_TRY(gbp::nccl_api().AllGather(h->g.bnd_send, (void*)h->g.bnd_recv, (size_t)h->g.n_bnd_global * 12, ncclFloat,
                                             h->comm, h->comm_stream));
      GBP_CUDA_TRY(cudaEventRecord(h->ev_recv, h->comm_stream));
      h->exchanges++;
    }
    if (grid + cams_grid(h)) {
      gbp::k_update_vars<<<grid + cams_grid(h), GBP_TILE, 0, h->stream>>>(h->g, shift, 0u, lower_only, 0u);
      h->kernels_launched++;
    }
    if (exchange) {
      GBP_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_recv, 0));
      if (bgrid) {
        gbp::k_boundary_finish<<<bgrid, GBP_TILE, 0, h->stream>>>(h->g, shift);
        h->kernels_launched++;
      }
    }
  }
  h->pending_shift = false;
  GBP_CUDA_TRY(cudaGetLastError());
  return GBP_OK;
}

// upper == false: the strict upper triangle of the camera messages is skipped (see k_sweep); only valid when the
// following belief update runs with lower_only and another, complete sweep follows before anything is read
template <bool PREP, bool MSG>
int launch_sweep(gbp_handle* h, bool upper = true) {
  if (h->n_tiles) {
    // persistent: one block per SM (fewer when the graph has fewer warp-tiles than that)
    const uint32_t n_wt = h->E_pad / 32;
    const uint32_t grid = std::min<uint32_t>((uint32_t)h->num_sms, n_wt);
    const bool full = upper || !MSG || h->g.mcam_up;
    if constexpr (MSG) {
      if (h->use_tma && h->fast_math) {
        if (gbp_fast_launch_sweep(&h->g, &h->maps, PREP ? 1 : 0, full ? 1 : 0, grid, h->stream) != 0) {
          gbp_set_error(std::string("fast-math sweep launch: ") + cudaGetErrorString(cudaGetLastError()));
          return GBP_ERR_CUDA;
        }
      } else if (h->use_tma) {
        if (full) gbp::k_sweep_tma<PREP, true, true><<<grid, GBP_TW * 32, GBP_T_SMEM, h->stream>>>(h->g, h->maps);
        else gbp::k_sweep_tma<PREP, true, false><<<grid, GBP_TW * 32, GBP_T_SMEM, h->stream>>>(h->g, h->maps);
      } else if (full) {  // GBP_SWEEP=cpasync: the round-1 kernel (per-lane cp.async staging, static tile order), kept as the reference
        gbp::k_sweep<PREP, true, true><<<grid, GBP_SW_WARPS * 32, GBP_SWEEP_SMEM, h->stream>>>(h->g);
      } else {
        gbp::k_sweep<PREP, true, false><<<grid, GBP_SW_WARPS * 32, GBP_SWEEP_SMEM, h->stream>>>(h->g);
      }
    } else {
      gbp::k_sweep<PREP, MSG, true><<<grid, GBP_SW_WARPS * 32, GBP_SWEEP_SMEM, h->stream>>>(h->g);
    }
    h->kernels_launched++;
  }
  if (PREP) h->pending_shift = true;
  if (MSG) h->p_in_sync = true;
  GBP_CUDA_TRY(cudaGetLastError());
  return GBP_OK;
}

// PrepMessageVertex of every factor: state-machine pass, then the listed factors relinearise
int launch_prep(gbp_handle* h) {
  if (h->n_tiles) {
    GBP_CUDA_TRY(cudaMemsetAsync(h->g.relin_count, 0, sizeof(uint32_t), h->stream));
    gbp::k_prep_pass<<<h->n_tiles, GBP_TILE, 0, h->stream>>>(h->g);
    gbp::k_relin_list<<<(h->E + GBP_TILE - 1) / GBP_TILE, GBP_TILE, 0, h->stream>>>(h->g);
    h->kernels_launched += 2;
  }
  h->pending_shift = true;
  GBP_CUDA_TRY(cudaGetLastError());
  return GBP_OK;
}

// one full sweep of the factors (prep + messages)
int launch_full_sweep(gbp_handle* h, bool upper = true) {
  if (!h->two_pass) return launch_sweep<true, true>(h, upper);
  int rc = launch_prep(h);
  if (!rc) rc = launch_sweep<false, true>(h, upper);
  return rc;
}
// may the sweeps of a call other than the last skip the upper triangle of the camera messages?
inline bool can_skip_upper(const gbp_handle* h) { return h->skip_upper && !h->g.mcam_up; }

int launch_metric(gbp_handle* h, gbp::DeviceStats* d_out) {
  uint32_t* cursor = h->capturing ? h->d_stat_cursor : nullptr;
  if (h->n_tiles) {
    const uint32_t nv = h->C + h->L;
    gbp::k_metric_prep<<<(nv + 127) / 128, 128, 0, h->stream>>>(h->g, h->d_met_cam, h->d_met_lmk);
    gbp::k_metric<<<metric_grid(h), GBP_TILE, 0, h->stream>>>(h->g, h->n_active, h->n_tiles, h->d_met_cam, h->d_met_lmk,
                                                              h->d_metric_parts, h->d_metric_ticket, d_out,
                                                              h->shard ? h->d_metric_raw : nullptr, h->shard ? nullptr : cursor);
    h->kernels_launched += 2;
  } else {
    gbp::k_metric_finish<<<1, GBP_TILE, 0, h->stream>>>(h->g, h->d_metric_parts, 0u, d_out, h->shard ? h->d_metric_raw : nullptr,
                                                        h->shard ? nullptr : cursor);
    h->kernels_launched++;
  }
  if (h->shard) {  // every rank reports the metric of the WHOLE graph
    if (h->p2p) {
      // the finishing block of k_metric pushed this rank's sums into every rank's receive buffer; wait for all
      gbp::k_metric_combine<<<1, 32, 0, h->stream>>>(h->g, nullptr, d_out, cursor);
    } else {
      GBP_CUDA_TRY(cudaEventRecord(h->ev_send, h->stream));
      GBP_CUDA_TRY(cudaStreamWaitEvent(h->comm_stream, h->ev_send, 0));
      GBP_NCCL_TRY(gbp::nccl_api().AllGather(h->d_metric_raw, h->d_metric_all, 8, ncclDouble, h->comm, h->comm_stream));
      GBP_CUDA_TRY(cudaEventRecord(h->ev_recv, h->comm_stream));
      GBP_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_recv, 0));
      gbp::k_metric_combine<<<1, 32, 0, h->stream>>>(h->g, h->d_metric_all, d_out, cursor);
    }
    h->kernels_launched++;
  }
  GBP_CUDA_TRY(cudaGetLastError());
  return GBP_OK;
}

// a kernel of this handle gave up waiting for a peer (bounded spin in boundary_finish / k_metric_combine)
int check_peer_error(gbp_handle* h) {
  if (h->p2p_err_host && *(volatile uint32_t*)h->p2p_err_host) {
    gbp_set_error("multi-GPU exchange: timed out waiting for a peer's boundary partials; the handle's state is invalid");
    return GBP_ERR_COMM;
  }
  return GBP_OK;
}

// (re)binds the handle's page-locked staging block with room for `n` per-sweep stats
int pin_block_bind(gbp_handle* h, size_t n) {
  size_t got = 0;
  void* p = pinned_get(PIN_OFF_STATS + n * sizeof(gbp_iter_stats), &got);
  if (!p) {
    gbp_set_error("cudaHostAlloc of the staging block failed");
    return GBP_ERR_CUDA;
  }
  if (h->pin_block) {
    std::memcpy(p, h->pin_block, PIN_OFF_STATS);  // error word + ring keep their values
    pinned_put(h->pin_block, h->pin_block_bytes);
  } else {
    std::memset(p, 0, PIN_OFF_STATS);
  }
  h->pin_block = p;
  h->pin_block_bytes = got;
  h->pin_ring = (uint32_t*)((char*)p + PIN_OFF_RING);
  h->pin_stats = (gbp_iter_stats*)((char*)p + PIN_OFF_STATS);
  h->pin_stats_cap = (got - PIN_OFF_STATS) / sizeof(gbp_iter_stats);
  return GBP_OK;
}

int ensure_stats(gbp_handle* h, size_t n) {
  if (n > h->pin_stats_cap) {
    if (h->p2p_err_host) {  // the error word of a sharded handle is baked into its kernels' arguments: it cannot move
      gbp_set_error("per-sweep metrics of more than 4096 sweeps per call are not supported on a sharded handle");
      return GBP_ERR_ARG;
    }
    int rc = pin_block_bind(h, std::max<size_t>(n, STATS_PREALLOC));
    if (rc) return rc;
  }
  if (n <= h->d_stats_cap) return GBP_OK;
  if (h->d_stats && !h->d_stats_in_arena) cudaFree(h->d_stats);
  h->d_stats = nullptr;
  h->d_stats_in_arena = false;
  GBP_CUDA_TRY(cudaMalloc((void**)&h->d_stats, n * sizeof(gbp::DeviceStats)));
  h->d_stats_cap = n;
  return GBP_OK;
}

int set_device(const gbp_handle* h) {
  GBP_CUDA_TRY(cudaSetDevice(h->device));
  return GBP_OK;
}

// ---- host <-> device record conversion -----------------------------------------
struct HostRecs {  // full host mirrors of the per-edge records, fetched on demand
  std::vector<float4> fac, mcam, mlmk, recA, recB;
};

int fetch(gbp_handle* h, std::vector<float4>& v, const float4* d, size_t n) {
  v.resize(n);
  int rc = download(v.data(), d, n, h->stream);
  if (rc) return rc;
  GBP_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return GBP_OK;
}
int push(gbp_handle* h, const std::vector<float4>& v, float4* d) {
  int rc = upload(d, v.data(), v.size(), h->stream);
  if (rc) return rc;
  GBP_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return GBP_OK;
}

inline float& quad_field(std::vector<float4>& v, size_t stride, size_t e, int field) {
  float4& q = v[(size_t)(field / 4) * stride + e];
  return (&q.x)[field % 4];
}
inline float& aos_field(std::vector<float4>& v, int quads, size_t e, int field) {
  float4& q = v[e * quads + field / 4];
  return (&q.x)[field % 4];
}

enum TensorId {
  T_CAM_B_ETA, T_CAM_B_LAM, T_LMK_B_ETA, T_LMK_B_LAM, T_CAM_M_ETA, T_CAM_M_LAM, T_LMK_M_ETA, T_LMK_M_LAM,
  T_F_ETA, T_F_LAM, T_DAMPING, T_DCOUNT, T_MU, T_OLDMU, T_DMU, T_ACTIVE, T_ROBUST, T_Z, T_VAR,
  T_CAM_SCALING, T_LMK_SCALING, T_CAM_WFLAG, T_LMK_WFLAG, T_NONE
};

TensorId tensor_id(const std::string& n) {
  if (n == "cam_beliefs_eta") return T_CAM_B_ETA;
  if (n == "cam_beliefs_lambda") return T_CAM_B_LAM;
  if (n == "lmk_beliefs_eta") return T_LMK_B_ETA;
  if (n == "lmk_beliefs_lambda") return T_LMK_B_LAM;
  // messages and previous messages share one buffer: a factor only ever reads
  // its OWN previous messages, so they are updated in place (no Copy, ba.cpp:902-905)
  if (n == "cam_messages_eta" || n == "pcam_messages_eta") return T_CAM_M_ETA;
  if (n == "cam_messages_lambda" || n == "pcam_messages_lambda") return T_CAM_M_LAM;
  if (n == "lmk_messages_eta" || n == "plmk_messages_eta") return T_LMK_M_ETA;
  if (n == "lmk_messages_lambda" || n == "plmk_messages_lambda") return T_LMK_M_LAM;
  if (n == "factor_potentials_eta") return T_F_ETA;
  if (n == "factor_potentials_lambda") return T_F_LAM;
  if (n == "damping") return T_DAMPING;
  if (n == "damping_count") return T_DCOUNT;
  if (n == "mu") return T_MU;
  if (n == "oldmu") return T_OLDMU;
  if (n == "dmu") return T_DMU;
  if (n == "active_flag") return T_ACTIVE;
  if (n == "robust_flag") return T_ROBUST;
  if (n == "measurements") return T_Z;
  if (n == "meas_variances") return T_VAR;
  if (n == "cam_scaling") return T_CAM_SCALING;
  if (n == "lmk_scaling") return T_LMK_SCALING;
  if (n == "cam_weaken_flag") return T_CAM_WFLAG;
  if (n == "lmk_weaken_flag") return T_LMK_WFLAG;
  return T_NONE;
}

size_t tensor_elems(const gbp_handle* h, TensorId id) {
  const size_t C = h->C, L = h->L, E = h->E, SK = h->SK, SL = h->SL;
  switch (id) {
    case T_CAM_B_ETA: return 6 * C;
    case T_CAM_B_LAM: return 36 * C;
    case T_LMK_B_ETA: return 3 * L;
    case T_LMK_B_LAM: return 9 * L;
    case T_CAM_M_ETA: return C * SK * 6;
    case T_CAM_M_LAM: return C * SK * 36;
    case T_LMK_M_ETA: return L * SL * 3;
    case T_LMK_M_LAM: return L * SL * 9;
    case T_F_ETA: return 9 * E;
    case T_F_LAM: return 81 * E;
    case T_MU: case T_OLDMU: return 9 * E;
    case T_Z: return 2 * E;
    case T_DAMPING: case T_DCOUNT: case T_DMU: case T_ACTIVE: case T_ROBUST: case T_VAR: return E;
    case T_CAM_SCALING: case T_CAM_WFLAG: return C;
    case T_LMK_SCALING: case T_LMK_WFLAG: return L;
    default: return 0;
  }
}

int recompute_means(gbp_handle* h) {
  const uint32_t n = h->C + h->L;
  if (n) {
    gbp::k_means_from_beliefs<<<(n + 127) / 128, 128, 0, h->stream>>>(h->g);
    h->kernels_launched++;
  }
  GBP_CUDA_TRY(cudaGetLastError());
  return GBP_OK;
}

int import_edges(gbp_handle* h, const float* damping, const int32_t* dcount, const uint32_t* active,
                 const uint32_t* robust, const float* dmu, int clear_muvalid) {
  const uint32_t E = h->E;
  if (!E) return GBP_OK;
  float* d_damp = nullptr;
  int32_t* d_dc = nullptr;
  uint32_t *d_act = nullptr, *d_rob = nullptr;
  float* d_dmu = nullptr;
  int rc = GBP_OK;
  auto stage = [&](auto** d, const auto* src) -> int {
    if (!src) return GBP_OK;
    int r = dev_alloc(d, (size_t)E, false);
    if (r) return r;
    return upload(*d, src, (size_t)E, h->stream);
  };
  if (!rc) rc = stage(&d_damp, damping);
  if (!rc) rc = stage(&d_dc, dcount);
  if (!rc) rc = stage(&d_act, active);
  if (!rc) rc = stage(&d_rob, robust);
  if (!rc) rc = stage(&d_dmu, dmu);
  if (!rc) {
    gbp::k_import_edges<<<(E + 255) / 256, 256, 0, h->stream>>>(h->g, h->d_pos_of_orig, d_damp, d_dc, d_act, d_rob,
                                                                  d_dmu, clear_muvalid);
    h->kernels_launched++;
    if (cudaGetLastError() != cudaSuccess) rc = GBP_ERR_CUDA;
  }
  cudaStreamSynchronize(h->stream);
  cudaFree(d_damp); cudaFree(d_dc); cudaFree(d_act); cudaFree(d_rob); cudaFree(d_dmu);
  if (active) {
    h->active_host.assign(active, active + E);
    if (!h->shard) {  // a shard keeps the active-edge count of the whole graph
      h->n_active = 0;
      for (uint32_t e = 0; e < E; ++e) h->n_active += (active[e] == 1u) ? 1u : 0u;
    }
  }
  return rc;
}

int upload_lmk_priors(gbp_handle* h, const float* eta, const float* lam) {
  // partial update allowed: fetch current, patch, push
  std::vector<float4> pr((size_t)h->L * 3);
  int rc = download(pr.data(), h->g.lmk_prior, pr.size(), h->stream);
  if (rc) return rc;
  GBP_CUDA_TRY(cudaStreamSynchronize(h->stream));
  for (size_t l = 0; l < h->L; ++l) {
    float* r = &pr[l * 3].x;
    if (eta) for (int i = 0; i < 3; ++i) r[i] = eta[l * 3 + i];
    if (lam) for (int i = 0; i < 9; ++i) r[3 + i] = lam[l * 9 + i];
  }
  rc = upload(h->g.lmk_prior, pr.data(), pr.size(), h->stream);
  if (rc) return rc;
  GBP_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return GBP_OK;
}

// ---- peer-to-peer exchange set-up ------------------------------------------------------------
// Every rank owns one block
//   [- | metric arrival flags W | - | metric receive buffer 2 x W x 8 doubles |
//    boundary receive buffer 2 parities x W x n_boundary x 12 tagged words {value, exchange step}]
// that every other rank maps: through CUDA IPC when the ranks are processes (the handles travel over the NCCL
// communicator once), directly (same device, or cudaDeviceEnablePeerAccess) when they are the handles of a
// single-process group.
constexpr size_t P2P_OFF_MFLAG = 1024, P2P_OFF_METRIC = 4096;
constexpr size_t P2P_OFF_RECV = 4096 + 2 * 256 * 8 * sizeof(double);  // metric buffer sized for <= 256 ranks
constexpr uint32_t P2P_MAX_WORLD = 256;

size_t p2p_block_bytes(uint32_t W, uint32_t n_bnd_global) {
  const size_t recv_bytes = (size_t)2 * W * n_bnd_global * 6 * sizeof(uint4);  // every float travels as {value, step}
  return ((P2P_OFF_RECV + recv_bytes + (2u << 20) - 1) >> 21) << 21;  // whole 2 MB pages
}

// bases[r] = rank r's block as seen from this rank's device (bases[rank] = its own)
int p2p_wire(gbp_handle* h, const std::vector<void*>& bases) {
  DeviceGraph& g = h->g;
  const uint32_t W = h->world;
  std::vector<uint4*> recv(W);
  std::vector<uint32_t*> mflag(W);
  std::vector<double*> metric(W);
  for (uint32_t r = 0; r < W; ++r) {
    char* base = (char*)bases[r];
    mflag[r] = (uint32_t*)(base + P2P_OFF_MFLAG);
    metric[r] = (double*)(base + P2P_OFF_METRIC);
    recv[r] = (uint4*)(base + P2P_OFF_RECV);
  }
  int rc = h_alloc(h, &g.peer_recv, W);
  if (!rc) rc = h_alloc(h, &g.peer_mflag, W);
  if (!rc) rc = h_alloc(h, &g.peer_metric, W);
  if (rc) return rc;
  GBP_CUDA_TRY(cudaMemcpy(g.peer_recv, recv.data(), sizeof(uint4*) * W, cudaMemcpyHostToDevice));
  GBP_CUDA_TRY(cudaMemcpy(g.peer_mflag, mflag.data(), sizeof(uint32_t*) * W, cudaMemcpyHostToDevice));
  GBP_CUDA_TRY(cudaMemcpy(g.peer_metric, metric.data(), sizeof(double*) * W, cudaMemcpyHostToDevice));
  char* own = (char*)bases[h->rank];
  g.metric_flag = (uint32_t*)(own + P2P_OFF_MFLAG);
  g.metric_recv = (const double*)(own + P2P_OFF_METRIC);
  g.p2p_recv = (const uint4*)(own + P2P_OFF_RECV);
  // the error word lives in mapped host memory: a timed-out kernel sets it, every synchronising entry point reads it
  h->p2p_err_host = (uint32_t*)h->pin_block;
  *h->p2p_err_host = 0u;
  GBP_CUDA_TRY(cudaHostGetDevicePointer((void**)&g.p2p_error, h->p2p_err_host, 0));
  // how long a block waits for its peers: generous (a peer may be instantiating a graph or sit in a debugger),
  // GBP_P2P_TIMEOUT_S overrides
  double seconds = 60.0;
  if (const char* env = std::getenv("GBP_P2P_TIMEOUT_S")) seconds = std::max(0.001, std::atof(env));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
  g.p2p_timeout = (long long)(seconds * 1e3 * (khz > 0 ? khz : 1965000));
  h->p2p = 1;
  return GBP_OK;
}

// Processes (one per GPU): returns GBP_OK with h->p2p == 0 when IPC is not available (the NCCL all-gather path
// is used instead).
int setup_p2p(gbp_handle* h, int mode) {
  DeviceGraph& g = h->g;
  const uint32_t W = h->world;
  h->p2p = 0;
  if (mode == 1 || g.n_bnd_global == 0) return GBP_OK;
  if (W > 32) {  // the per-landmark rank masks of the peer-to-peer exchange are 32 bits wide
    if (mode == 2) {
      gbp_set_error("peer-to-peer exchange supports up to 32 ranks");
      return GBP_ERR_ARG;
    }
    return GBP_OK;
  }
  const size_t total = p2p_block_bytes(W, g.n_bnd_global);
  int ok = 1;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (W > P2P_MAX_WORLD) ok = 0;
  if (ok && cudaMalloc(&h->p2p_block, total) != cudaSuccess) ok = 0;
  if (ok && cudaMemset(h->p2p_block, 0, total) != cudaSuccess) ok = 0;
  if (ok && cudaDeviceSynchronize() != cudaSuccess) ok = 0;  // zeroed BEFORE any peer can learn the handle and push into it
  if (ok && cudaIpcGetMemHandle(&mine, h->p2p_block) != cudaSuccess) ok = 0;
  cudaGetLastError();
  // all-gather {ok, handle} over the communicator (also tells every rank whether ALL ranks can do it)
  struct Item { int ok; int pad[3]; cudaIpcMemHandle_t hd; };
  static_assert(sizeof(Item) == 80, "item layout");
  Item it;
  it.ok = ok; it.pad[0] = it.pad[1] = it.pad[2] = 0; it.hd = mine;
  Item* d_items = nullptr;
  GBP_CUDA_TRY(cudaMalloc((void**)&d_items, sizeof(Item) * (W + 1)));
  GBP_CUDA_TRY(cudaMemcpyAsync(d_items + W, &it, sizeof(Item), cudaMemcpyHostToDevice, h->comm_stream));
  GBP_NCCL_TRY(gbp::nccl_api().AllGather(d_items + W, d_items, sizeof(Item), ncclChar, h->comm, h->comm_stream));
  std::vector<Item> all(W);
  GBP_CUDA_TRY(cudaMemcpyAsync(all.data(), d_items, sizeof(Item) * W, cudaMemcpyDeviceToHost, h->comm_stream));
  GBP_CUDA_TRY(cudaStreamSynchronize(h->comm_stream));
  cudaFree(d_items);
  for (uint32_t r = 0; r < W; ++r) ok = ok && all[r].ok;
  h->p2p_peers.assign(W, nullptr);
  for (uint32_t r = 0; r < W && ok; ++r) {
    if (r == h->rank) continue;
    if (cudaIpcOpenMemHandle(&h->p2p_peers[r], all[r].hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      h->p2p_peers[r] = nullptr;
      ok = 0;
    }
  }
  cudaGetLastError();
  // second round: fall back everywhere unless every rank mapped every peer
  int* d_ok = nullptr;
  GBP_CUDA_TRY(cudaMalloc((void**)&d_ok, sizeof(int) * (W + 1)));
  GBP_CUDA_TRY(cudaMemcpyAsync(d_ok + W, &ok, sizeof(int), cudaMemcpyHostToDevice, h->comm_stream));
  GBP_NCCL_TRY(gbp::nccl_api().AllGather(d_ok + W, d_ok, sizeof(int), ncclChar, h->comm, h->comm_stream));
  std::vector<int> oks(W);
  GBP_CUDA_TRY(cudaMemcpyAsync(oks.data(), d_ok, sizeof(int) * W, cudaMemcpyDeviceToHost, h->comm_stream));
  GBP_CUDA_TRY(cudaStreamSynchronize(h->comm_stream));
  cudaFree(d_ok);
  for (uint32_t r = 0; r < W; ++r) ok = ok && oks[r];
  if (!ok) {
    for (void*& q : h->p2p_peers)
      if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
    if (h->p2p_block) cudaFree(h->p2p_block);
    h->p2p_block = nullptr;
    cudaGetLastError();
    if (mode == 2) {
      gbp_set_error("peer-to-peer exchange requested but CUDA IPC mapping of the peers failed");
      return GBP_ERR_COMM;
    }
    return GBP_OK;
  }
  std::vector<void*> bases(W);
  for (uint32_t r = 0; r < W; ++r) bases[r] = (r == h->rank) ? h->p2p_block : h->p2p_peers[r];
  return p2p_wire(h, bases);
}

// TMA descriptors for k_sweep_tma: the factor potentials and the camera-bound messages seen as 2-D fp32 tensors
// [rows][E_pad * 4] whose box is one warp-tile: [rows] x [128 floats = the 32 lanes' quads of one row].
// GBP_SWEEP=cpasync selects the round-1 kernel (k_sweep: per-lane cp.async staging) instead.
int setup_tma(gbp_handle* h) {
  h->use_tma = 1;
  if (const char* env = std::getenv("GBP_SWEEP")) h->use_tma = std::strcmp(env, "cpasync") != 0;
  if (!h->use_tma || !h->n_tiles) return GBP_OK;
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      gbp_set_error("cuTensorMapEncodeTiled is not available from this driver");
      return GBP_ERR_CUDA;
    }
    encode = (EncodeFn)fn;
  }
  int promo = 2;  // CU_TENSOR_MAP_L2_PROMOTION_L2_128B
  if (const char* env = std::getenv("GBP_TMA_L2PROMO")) promo = std::max(0, std::min(3, std::atoi(env)));
  auto make = [&](CUtensorMap* m, void* base, uint32_t rows) -> int {
    const cuuint64_t dims[2] = {(cuuint64_t)h->E_pad * 4, rows};
    const cuuint64_t strides[1] = {(cuuint64_t)h->E_pad * 16};
    const cuuint32_t box[2] = {128, rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      gbp_set_error("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
      return GBP_ERR_CUDA;
    }
    return GBP_OK;
  };
  if (h->fast_math && gbp_fast_graph_bytes() != sizeof(DeviceGraph)) {
    gbp_set_error("fast-math build out of sync with the default build (DeviceGraph layout)");
    return GBP_ERR_CUDA;
  }
  int rc = make(&h->maps.fac, h->g.fac, GBP_FAC_QUADS);
  if (!rc) rc = make(&h->maps.mcam, h->g.mcam, GBP_MCAM_QUADS);
  if (rc) return rc;
  GBP_CUDA_TRY(cudaFuncSetAttribute(gbp::k_sweep_tma<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GBP_T_SMEM));
  GBP_CUDA_TRY(cudaFuncSetAttribute(gbp::k_sweep_tma<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GBP_T_SMEM));
  GBP_CUDA_TRY(cudaFuncSetAttribute(gbp::k_sweep_tma<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GBP_T_SMEM));
  GBP_CUDA_TRY(cudaFuncSetAttribute(gbp::k_sweep_tma<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GBP_T_SMEM));
  return GBP_OK;
}

// CUDA loads kernels lazily, and loading one can need a context-wide synchronisation.  In a single-process group
// rank A's belief update spins on its peers while rank B launches a kernel for the first time: if that launch had
// to load the kernel it would wait for A, which waits for B.  Every kernel of the library is therefore loaded
// up front (cudaFuncGetAttributes forces the load).
template <class K>
void preload(K kernel) {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, kernel);
}
void preload_kernels() {
  preload(gbp::k_sweep<true, true, true>); preload(gbp::k_sweep<true, true, false>);
  preload(gbp::k_sweep<false, true, true>); preload(gbp::k_sweep<false, true, false>);
  preload(gbp::k_sweep_tma<true, true, true>); preload(gbp::k_sweep_tma<true, true, false>);
  preload(gbp::k_sweep_tma<false, true, true>); preload(gbp::k_sweep_tma<false, true, false>);
  preload(gbp::k_prep_pass); preload(gbp::k_relin_list); preload(gbp::k_cam_partials); preload(gbp::k_update_vars);
  preload(gbp::k_boundary_partial); preload(gbp::k_boundary_records); preload(gbp::k_boundary_finish); preload(gbp::k_relinearise_all); preload(gbp::k_weaken);
  preload(gbp::k_kf_pose); preload(gbp::k_kf_apply); preload(gbp::k_metric_prep); preload(gbp::k_metric);
  preload(gbp::k_metric_finish); preload(gbp::k_metric_combine); preload(gbp::k_export_edges);
  preload(gbp::k_export_lmk_beliefs); preload(gbp::k_import_edges); preload(gbp::k_means_from_beliefs);
  cudaGetLastError();
}

struct PhaseTimer {  // GBP_INIT_TIMING=1: wall time of the phases of gbp_cuda_init on stderr
  bool on = std::getenv("GBP_INIT_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[gbp init] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

// Host-side copies of the index maps (reference edge order <-> edge slots, message slots): only get_tensor / set_tensor
// and the SLAM bookkeeping need them, so they are rebuilt from the device arrays the first time such a call is made --
// gbp_cuda_init itself does no O(E) work on the host (the maps are built by the kernels of gbp_setup.cu).
int ensure_host_maps(gbp_handle* h) {
  if (h->host_maps_ready) return GBP_OK;
  const uint32_t C = h->C, L = h->L, E = h->E;
  h->cam_ids.resize(E);
  h->lmk_ids.resize(E);
  h->pos_of_orig.resize(E);
  h->lmk_ptr.assign(L + 1, 0);
  std::vector<uint32_t> first_cam(L, 0xffffffffu);
  int rc = download(h->cam_ids.data(), h->d_cam_ids, E, h->stream);
  if (!rc) rc = download(h->lmk_ids.data(), h->d_lmk_ids, E, h->stream);
  if (!rc) rc = download(h->pos_of_orig.data(), h->d_pos_of_orig, E, h->stream);
  if (!rc) rc = download(h->lmk_ptr.data(), h->g.lmk_ptr, (size_t)L + 1, h->stream);
  if (!rc) rc = download(first_cam.data(), h->d_lmk_first_cam, L, h->stream);
  if (rc) return rc;
  GBP_CUDA_TRY(cudaStreamSynchronize(h->stream));
  // message slots = number of earlier edges of the same variable (ba/ba.cpp:267-279), in O(E)
  std::vector<uint32_t> deg_c(C, 0), deg_l(L, 0);
  h->slot_c.resize(E);
  h->slot_l.resize(E);
  for (uint32_t e = 0; e < E; ++e) {
    h->slot_c[e] = deg_c[h->cam_ids[e]]++;
    h->slot_l[e] = deg_l[h->lmk_ids[e]]++;
  }
  if (h->active_host.empty()) h->active_host.assign(E, 1u);
  // SLAM bookkeeping: landmarks first observed by each camera, the factors of every camera
  h->new_lmks_at_cam.assign(C, 0u);
  for (uint32_t l = 0; l < L; ++l)
    if (first_cam[l] != 0xffffffffu) h->new_lmks_at_cam[first_cam[l]]++;
  h->cam_edge_ptr.assign(C + 1, 0u);
  for (uint32_t c = 0; c < C; ++c) h->cam_edge_ptr[c + 1] = h->cam_edge_ptr[c] + deg_c[c];
  h->cam_edge_ids.resize(E);
  for (uint32_t e = 0; e < E; ++e) h->cam_edge_ids[h->cam_edge_ptr[h->cam_ids[e]] + h->slot_c[e]] = e;
  h->host_maps_ready = true;
  return GBP_OK;
}

// a block from the library's stream-ordered pool that is given back as soon as the stream has passed this point
struct ScopedPoolBlock {
  gbp_handle* h;
  char* p = nullptr;
  bool from_pool = false;
  explicit ScopedPoolBlock(gbp_handle* hh) : h(hh) {}
  int alloc(size_t bytes) {
    if (cudaMemPool_t pool = private_pool(h->device)) {
      if (cudaMallocFromPoolAsync((void**)&p, std::max<size_t>(bytes, 16), pool, h->stream) == cudaSuccess) {
        from_pool = true;
        return GBP_OK;
      }
      cudaGetLastError();
    }
    GBP_CUDA_TRY(cudaMalloc((void**)&p, std::max<size_t>(bytes, 16)));
    return GBP_OK;
  }
  ~ScopedPoolBlock() {
    if (!p) return;
    if (from_pool) {
      cudaFreeAsync(p, h->stream);
    } else {
      cudaStreamSynchronize(h->stream);
      cudaFree(p);
    }
  }
};

int build(gbp_handle* h, const gbp_problem* p, const gbp_opts* o, const uint32_t* edge_global = nullptr) {
  PhaseTimer pt;
  const uint32_t C = p->n_keyframes, L = p->n_points, E = p->n_edges;
  h->C = C; h->L = L; h->E = E;
  cudaStream_t s = h->stream;
  int rc = GBP_OK;
  // ---- stage A: raw arrays to the device, degrees + warp-tile count back (gbp_setup.cu)
  // the two id arrays stay resident (the lazily built host maps come from them); everything else of this block is scratch
  GBP_CUDA_TRY(cudaMalloc((void**)&h->d_cam_ids, std::max<size_t>((size_t)E * 8, 16)));
  h->allocs.push_back((void*)h->d_cam_ids);
  h->d_lmk_ids = h->d_cam_ids + E;
  auto up256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t raw_bytes = up256((size_t)E * 8) + up256((size_t)E * 4) * 5 + up256((size_t)L * 12) + up256((size_t)L * 36);
  ScopedPoolBlock tmp(h);
  rc = tmp.alloc(raw_bytes + gbp::setup_temp_bytes(E, C, L));
  if (rc) return rc;
  char* cur = tmp.p;
  auto take = [&](size_t bytes) {
    char* q = cur;
    cur += up256(bytes);
    return q;
  };
  float* d_z = (float*)take((size_t)E * 8);
  float* d_var = (float*)take((size_t)E * 4);
  uint32_t* d_active = (uint32_t*)take((size_t)E * 4);
  float* d_damping = (float*)take((size_t)E * 4);
  int32_t* d_dcount = (int32_t*)take((size_t)E * 4);
  uint32_t* d_eglobal = (uint32_t*)take((size_t)E * 4);
  float* d_lpe = (float*)take((size_t)L * 12);
  float* d_lpl = (float*)take((size_t)L * 36);
  char* d_setup_tmp = cur;
#define U_(dst, src, n) if (!rc) rc = upload(dst, src, (size_t)(n), s)
  U_(h->d_cam_ids, p->cam_ids, E);
  U_(h->d_lmk_ids, p->lmk_ids, E);
  gbp::SetupInputs in{};
  in.E = E; in.C = C; in.L = L;
  in.cam_ids = h->d_cam_ids;
  in.lmk_ids = h->d_lmk_ids;
  gbp::SetupTemp st{};
  if (!rc && gbp::setup_stage_a(s, in, d_setup_tmp, &st) != 0) {
    gbp_set_error(std::string("device setup (stage A): ") + cudaGetErrorString(cudaGetLastError()));
    rc = GBP_ERR_CUDA;
  }
  uint32_t info[4] = {0, 0, 0, 0};
  if (!rc) rc = download(info, st.info, 4, s);
  // the rest of the raw arrays travels while the host waits for those 16 bytes
  U_(d_z, p->measurements, 2 * (size_t)E);
  U_(d_var, p->meas_variances, E);
  if (p->active_flag) U_(d_active, p->active_flag, E);
  if (p->damping) U_(d_damping, p->damping, E);
  if (p->damping_count) U_(d_dcount, p->damping_count, E);
  if (edge_global) U_(d_eglobal, edge_global, E);
  U_(d_lpe, p->lmk_priors_eta, 3 * (size_t)L);
  U_(d_lpl, p->lmk_priors_lambda, 9 * (size_t)L);
  if (rc) return rc;
  // host work that overlaps the transfers: active-edge count, the streamed mu / oldmu (normally null or zero)
  h->n_active = E;
  if (p->active_flag) {
    h->active_host.assign(p->active_flag, p->active_flag + E);
    h->n_active = 0;
    for (uint32_t e = 0; e < E; ++e) h->n_active += (p->active_flag[e] == 1u) ? 1u : 0u;
  }
  if (h->shard) h->n_active = gbp_shard_n_active_global(h->shard);
  auto any_nonzero = [&](const float* a) {
    if (!a) return false;
    const size_t n = (size_t)9 * E;
    for (size_t i = 0; i < n; ++i)
      if (a[i] != 0.f) return true;
    return false;
  };
  if (any_nonzero(p->mu)) h->mu_init.assign(p->mu, p->mu + (size_t)9 * E);
  if (any_nonzero(p->oldmu)) h->oldmu_init.assign(p->oldmu, p->oldmu + (size_t)9 * E);
  GBP_CUDA_TRY(cudaStreamSynchronize(s));
  if (info[3]) {
    gbp_set_error("edge index out of range");
    return GBP_ERR_ARG;
  }
  h->SK = info[1] + 1;
  h->SL = info[2] + 1;
  // warp-tiles: each camera's factors are padded to a multiple of 32 edge slots (one thread per
  // factor, one camera per warp); the slot count is rounded up to whole GBP_TILE blocks for the
  // helper kernels (trailing warp-tiles hold no factor)
  const uint64_t n_wt64 = ((uint64_t)info[0] + GBP_WARPS - 1) / GBP_WARPS * GBP_WARPS;
  const uint64_t epad64 = n_wt64 * 32;
  if (epad64 >= 0xffffffffull) {
    gbp_set_error("problem too large for 32-bit edge slots");
    return GBP_ERR_ARG;
  }
  h->E_pad = (uint32_t)epad64;
  h->n_tiles = h->E_pad / GBP_TILE;
  const uint32_t n_wt = h->E_pad / 32;
  const size_t EP = h->E_pad;
  const uint32_t n_lmk_blocks = (L + GBP_LMK_PER_BLOCK - 1) / GBP_LMK_PER_BLOCK;
  pt.lap("raw uploads + degrees (device)");
  // ---- device allocation (everything the reference relies on being zero IS zeroed, quirk Q4)
  DeviceGraph& g = h->g;
  std::memset(&g, 0, sizeof(g));
  g.C = C; g.L = L; g.E = E; g.E_pad = h->E_pad;
  g.n_lmk_blocks = n_lmk_blocks;
  g.K[0] = p->K[0]; g.K[1] = p->K[4]; g.K[2] = p->K[2]; g.K[3] = p->K[5];
  g.hp.maxeta_damping = o->maxeta_damping;
  g.hp.num_undamped_iters = o->num_undamped_iters;
  g.hp.dmu_threshold = o->dmu_threshold;
  g.hp.min_linear_iters = o->min_linear_iters;
  g.hp.Nstds = o->Nstds;
  // one device arena for all per-handle buffers: a single cudaMalloc + cudaMemset instead of ~40
  // (everything the reference relies on being zero IS zeroed, quirk Q4)
  std::vector<std::pair<void**, size_t>> arena;
  size_t arena_bytes = 0;
  auto reserve = [&](void** pp, size_t bytes) {
    arena.emplace_back(pp, arena_bytes);
    arena_bytes += (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
  };
#define A_(ptr, n) reserve((void**)&(ptr), (size_t)(n) * sizeof(*(ptr)))
  A_(g.fac, GBP_FAC_QUADS * EP);
  A_(g.mcam, GBP_MCAM_QUADS * EP);
  if (o->store_full_messages) A_(g.mcam_up, 4 * EP);
  // buffers that are touched again within a sweep or by the next one (landmark-bound messages: written and
  // read back by k_sweep, streamed by k_update_vars; edge state; camera partials; landmark records) are kept
  // together so that ONE access-policy window can pin them in L2 (l2_persist below)
  const size_t reuse_begin = arena_bytes;
  A_(g.mlmk, GBP_MLMK_QUADS * (size_t)E);
  const size_t reuse_mlmk_end = arena_bytes;
  A_(g.recA, EP);
  A_(g.cam_partial, (size_t)n_wt * GBP_CAMPART_STRIDE);
  A_(g.lmk_b, GBP_LMKB_QUADS * (size_t)L);
  A_(g.lmk_mean_prev, L);
  A_(g.lmk_sq, L);
  const size_t reuse_end = arena_bytes;
  A_(g.var, EP);
  A_(g.recB, EP);
  A_(g.edge_orig, EP);
  A_(g.wt_info, n_wt);
  A_(g.cam_rec, 16 * (size_t)C);
  A_(g.cam_wt_begin, C + 1);
  A_(g.cam_b_eta, 6 * (size_t)C);
  A_(g.cam_b_lam, 36 * (size_t)C);
  A_(g.cam_mean, 6 * (size_t)C);
  A_(g.cam_mean_prev, 6 * (size_t)C);
  A_(g.cam_lin, 5 * (size_t)C);
  A_(g.cam_prior_eta, 6 * (size_t)C);
  A_(g.cam_prior_lam, 36 * (size_t)C);
  A_(g.cam_scaling, C);
  A_(g.cam_wflag, C);
  A_(g.lmk_prior, 3 * (size_t)L);
  A_(g.lmk_scaling, L);
  A_(g.lmk_wflag, L);
  A_(g.lmk_ptr, L + 1);
  A_(g.lmk_blk, n_lmk_blocks);
  A_(h->d_pos_of_orig, E);
  A_(h->d_lmk_first_cam, L);
  A_(h->d_kf_scratch, 4);
  A_(h->d_metric_parts, h->n_tiles);
  A_(h->d_metric_ticket, 1);
  A_(h->d_stat_cursor, 1);
  A_(h->d_stats, STATS_PREALLOC);
  A_(g.relin_list, E);
  A_(g.relin_count, 1);
  A_(g.relin_ring, GBP_RELIN_RING + 1);
  A_(g.tile_queue, 2);
  A_(h->d_exp_lmk_eta, 3 * (size_t)L);   // READ_PROG staging (unpacked landmark beliefs, per-edge scalars in edge order)
  A_(h->d_exp_lmk_lam, 9 * (size_t)L);
  A_(h->d_exp_damping, E);
  A_(h->d_exp_dcount, E);
  A_(h->d_exp_robust, E);
  A_(h->d_met_cam, 16 * (size_t)C);
  A_(h->d_met_lmk, 4 * (size_t)L);
  A_(h->d_pprior_cam_eta, 6 * (size_t)C);
  A_(h->d_pprior_cam_lam, 36 * (size_t)C);
  A_(h->d_pprior_lmk, 3 * (size_t)L);
  if (h->shard) {
    const uint32_t nbl = gbp_shard_n_boundary_local(h->shard);
    g.n_bnd_local = nbl;
    g.n_bnd_global = gbp_shard_get_plan(h->shard)->n_boundary_points;
    g.world = h->world;
    g.rank = h->rank;
    A_(g.bnd_local, nbl);
    A_(g.bnd_slot, nbl);
    A_(g.bnd_span, nbl);
    A_(g.bnd_rec, nbl);
    A_(g.bnd_send, 3 * (size_t)g.n_bnd_global);
    reserve((void**)&g.bnd_recv, 3 * (size_t)g.n_bnd_global * h->world * sizeof(float4));
    A_(g.p2p_step, 2);
    A_(g.metric_step, 1);
    A_(h->d_metric_raw, 8);
    A_(h->d_metric_all, 8 * (size_t)h->world);
  }
#undef A_
  {
    char* base = nullptr;
    rc = arena_alloc(h, &base, arena_bytes);
    if (rc) return rc;
    GBP_CUDA_TRY(cudaMemsetAsync(base, 0, arena_bytes, h->stream));  // ordered before the uploads below
    for (auto& r : arena) *r.first = base + r.second;
    h->d_stats_cap = STATS_PREALLOC;
    h->d_stats_in_arena = true;
    rc = pin_block_bind(h, STATS_PREALLOC);
    if (rc) return rc;
    // L2 residency of the re-used buffers (B200: 126 MB L2, at most 79 MB of it can be set aside).
    // GBP_L2_PERSIST: 0 off, 1 (default) the landmark-bound messages, 2 all re-used buffers.  Measured on
    // config 4 (profiles/): 161.9 us per sweep without, 155.1 us with the 48 MB of landmark-bound messages
    // pinned (k_update_vars 28 -> 23.5 us, k_sweep 142 -> 140 us), 167 us with all 77 MB (too little L2 is
    // left for the streams); windows with a hit ratio below 1 only lose.
    int mode = 1;
    if (const char* env = std::getenv("GBP_L2_PERSIST")) mode = std::atoi(env);
    h->l2_window = cudaAccessPolicyWindow{};
    h->l2_window.num_bytes = 0;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
    size_t want = (mode == 1 ? reuse_mlmk_end : reuse_end) - reuse_begin;
    // never more than 48 MB: a window that needs a hit ratio below 1, or a set-aside that leaves the streams less
    // than ~70 MB of L2, loses (measured); on larger graphs the window covers the first 48 MB of the messages
    size_t window_cap_mb = 48;
    if (const char* env = std::getenv("GBP_L2_WINDOW_MB")) window_cap_mb = (size_t)std::atoi(env);
    want = std::min(want, window_cap_mb << 20);
    size_t cap_mb = 1u << 20;
    if (const char* env = std::getenv("GBP_L2_SETASIDE_MB")) cap_mb = (size_t)std::atoi(env);
    if (mode > 0 && max_persist > 0 && max_window > 0 && want >= ((size_t)8 << 20)) {  // small graphs live in L2 anyway
      const size_t win = std::min(want, (size_t)max_window);
      size_t cur = 0;
      cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
      const size_t set_aside = std::min(std::min(std::max(cur, win), (size_t)max_persist), cap_mb << 20);
      {
        std::lock_guard<std::mutex> lk(pools_mutex());
        L2State& ls = l2_state(h->device);
        if (ls.users == 0) ls.prev_limit = cur;  // restored by the last handle with a window (gbp_cuda_free)
        ls.users++;
        h->l2_limit_changed = true;
      }
      if (set_aside != cur) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside);
      h->l2_window.base_ptr = base + reuse_begin;
      h->l2_window.num_bytes = win;
      h->l2_window.hitRatio = std::min(1.0f, (float)((double)set_aside / (double)win));
      h->l2_window.hitProp = cudaAccessPropertyPersisting;
      h->l2_window.missProp = cudaAccessPropertyStreaming;
      cudaStreamAttrValue av;
      av.accessPolicyWindow = h->l2_window;
      if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) {
        cudaGetLastError();
        h->l2_window.num_bytes = 0;
      }
      if (std::getenv("GBP_INIT_TIMING"))
        std::fprintf(stderr, "[gbp init] L2 window %.1f MB, set-aside %.1f MB (max %.1f MB), hit ratio %.2f\n", win / 1048576.0,
                     set_aside / 1048576.0, max_persist / 1048576.0, h->l2_window.hitRatio);
    }
  }
  pt.lap("cudaMalloc + memset");
  // ---- stage B: scans, stable sorts, edge records, warp-tile table, packed priors -- all into the arena
  in.measurements = d_z;
  in.meas_variances = d_var;
  in.active_flag = p->active_flag ? d_active : nullptr;
  in.damping = p->damping ? d_damping : nullptr;
  in.damping_count = p->damping_count ? d_dcount : nullptr;
  in.edge_global = edge_global ? d_eglobal : nullptr;
  in.lmk_priors_eta = d_lpe;
  in.lmk_priors_lambda = d_lpl;
  gbp::SetupOutputs so{};
  so.E_pad = h->E_pad; so.lmk_per_block = GBP_LMK_PER_BLOCK;
  so.recA = g.recA; so.recB = g.recB; so.var = g.var; so.edge_orig = g.edge_orig; so.wt_info = g.wt_info;
  so.cam_wt_begin = g.cam_wt_begin; so.lmk_ptr = g.lmk_ptr; so.pos_of_orig = h->d_pos_of_orig;
  so.lmk_first_cam = h->d_lmk_first_cam; so.lmk_prior = g.lmk_prior; so.lmk_blk = g.lmk_blk;
  if (gbp::setup_stage_b(s, in, st, so) != 0) {
    gbp_set_error(std::string("device setup (stage B): ") + cudaGetErrorString(cudaGetLastError()));
    return GBP_ERR_CUDA;
  }
  U_(g.cam_prior_eta, p->cam_priors_eta, 6 * (size_t)C);
  U_(g.cam_prior_lam, p->cam_priors_lambda, 36 * (size_t)C);
  U_(g.cam_scaling, p->cam_scaling, C);
  U_(g.cam_wflag, p->cam_weaken_flag, C);
  U_(g.lmk_scaling, p->lmk_scaling, L);
  U_(g.lmk_wflag, p->lmk_weaken_flag, L);
  if (h->shard) {
    U_(g.bnd_local, gbp_shard_boundary_local(h->shard), g.n_bnd_local);
    U_(g.bnd_slot, gbp_shard_boundary_slot(h->shard), g.n_bnd_local);
    U_(g.bnd_span, gbp_shard_boundary_ranks(h->shard), g.n_bnd_local);
    if (!rc && g.n_bnd_local) {
      gbp::k_boundary_records<<<(g.n_bnd_local + 255) / 256, 256, 0, s>>>(g);  // after stage B: it reads lmk_ptr
      h->kernels_launched++;
    }
  }
  std::vector<float> oldmu_t;
  if (!rc && !h->oldmu_init.empty()) {  // a streamed per-edge oldmu (rare): transposed into edge-slot order
    rc = ensure_host_maps(h);
    if (!rc) rc = h_alloc(h, &g.oldmu_edge, 9 * EP);
    oldmu_t.assign(9 * EP, 0.f);
    for (uint32_t e = 0; e < E && !rc; ++e)
      for (int i = 0; i < 9; ++i) oldmu_t[(size_t)i * EP + h->pos_of_orig[e]] = h->oldmu_init[(size_t)9 * e + i];
    U_(g.oldmu_edge, oldmu_t.data(), oldmu_t.size());
  }
#undef U_
  if (rc) return rc;
  GBP_CUDA_TRY(cudaStreamSynchronize(s));
#define GBP_SMEM_ATTR(k) GBP_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, GBP_SWEEP_SMEM))
  GBP_SMEM_ATTR((gbp::k_sweep<true, true, true>)); GBP_SMEM_ATTR((gbp::k_sweep<true, true, false>));
  GBP_SMEM_ATTR((gbp::k_sweep<false, true, true>)); GBP_SMEM_ATTR((gbp::k_sweep<false, true, false>));
  GBP_SMEM_ATTR((gbp::k_sweep<true, false, true>));
#undef GBP_SMEM_ATTR
  GBP_CUDA_TRY(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
  rc = setup_tma(h);
  if (rc) return rc;
  if (std::getenv("GBP_DEBUG_TS")) {  // diagnostic builds (-DGBP_DEBUG_TS): per-exchange timestamps of k_update_vars
    rc = h_alloc(h, &g.dbg_ts, 32 * 8);
    if (rc) return rc;
    std::vector<unsigned long long> init(32 * 8, 0ull);
    for (int i = 0; i < 32; ++i) init[i * 8 + 0] = init[i * 8 + 2] = ~0ull;  // the two minima
    GBP_CUDA_TRY(cudaMemcpy(g.dbg_ts, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
  }
  if (const char* env = std::getenv("GBP_TILE_QUEUE"))
    if (std::atoi(env) == 0) g.tile_queue = nullptr;  // static round-robin over the warp-tiles (diagnostics)
  pt.lap("uploads");
  return GBP_OK;
}

// LINEARISE_PROG (ba/ba.cpp:890-893): beliefs <- priors, then linearise every factor.  Enqueued only: on a
// multi-GPU handle the belief update contains the first boundary exchange, so every rank must have enqueued it
// before any rank waits (linearise_wait).
int linearise_prog(gbp_handle* h) {
  h->pending_shift = false;
  int rc = launch_update_vars(h);
  if (rc) return rc;
  if (h->n_tiles) {
    gbp::k_relinearise_all<<<h->n_tiles, GBP_TILE, 0, h->stream>>>(h->g);
    h->kernels_launched++;
  }
  GBP_CUDA_TRY(cudaGetLastError());
  return GBP_OK;
}
int linearise_wait(gbp_handle* h) {
  GBP_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return check_peer_error(h);
}

// The fused sweep kernel is fastest while relinearisations come in lock step (all factors in one sweep out of
// ~11, none in between: what the reference's uniform --undamped_start produces on a fresh graph); once they are
// spread over the sweeps most warps contain a relinearising lane and the compacting two-pass sweep wins (config 4:
// 171 us either way against 161 / 262 us for the fused kernel).  Decided from the last <= 32 sweeps.
int choose_sweep_flavour(gbp_handle* h, const uint32_t* ring_in = nullptr) {
  if (h->relin_mode == 1) { h->two_pass = 0; return GBP_OK; }
  if (h->relin_mode == 2) { h->two_pass = 1; return GBP_OK; }
  uint32_t local[GBP_RELIN_RING + 1];
  const uint32_t* ring = ring_in;
  if (!ring) {
    GBP_CUDA_TRY(cudaMemcpyAsync(local, h->g.relin_ring, sizeof(local), cudaMemcpyDeviceToHost, h->stream));
    GBP_CUDA_TRY(cudaStreamSynchronize(h->stream));
    ring = local;
  }
  const uint32_t n = std::min<uint32_t>(ring[GBP_RELIN_RING], GBP_RELIN_RING);  // completed sweeps on record
  if (n < 8) return GBP_OK;
  uint32_t busy = 0;
  for (uint32_t k = 1; k <= n; ++k)  // slot of the sweep counter itself is the (empty) next one
    if (ring[(ring[GBP_RELIN_RING] - k) % GBP_RELIN_RING] > h->E / 500) busy++;
  h->two_pass = (busy * 2 > n) ? 1 : 0;
  return GBP_OK;
}

void drop_graphs(gbp_handle* h) {
  for (cudaGraphExec_t& e : h->sweep_graph) {
    if (e) cudaGraphExecDestroy(e);
    e = nullptr;
  }
}

// The instantiated graph of ONE steady-state sweep: k_sweep, k_update_vars (shift = 1) and, with
// stats, the three metric kernels writing to d_stats[cursor++].  Captured lazily, re-captured when
// the kernel arguments (DeviceGraph, stats buffer) changed.
int sweep_graph(gbp_handle* h, bool with_stats, bool upper, cudaGraphExec_t* out) {
  if (std::memcmp(&h->graph_g, &h->g, sizeof(DeviceGraph)) != 0 || h->graph_stats != h->d_stats ||
      h->graph_n_active != h->n_active)
    drop_graphs(h);
  cudaGraphExec_t& exec = h->sweep_graph[(upper ? 0 : 4) + (h->two_pass ? 2 : 0) + (with_stats ? 1 : 0)];
  if (!exec) {
    const uint64_t k0 = h->kernels_launched;
    cudaGraph_t graph = nullptr;
    GBP_CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    h->capturing = true;
    int rc = launch_full_sweep(h, upper);
    if (!rc) rc = launch_update_vars(h, !upper);
    if (!rc && with_stats) rc = launch_metric(h, h->d_stats);
    h->capturing = false;
    const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    h->kernels_launched = k0;
    if (rc || ce != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      gbp_set_error("CUDA graph capture of the sweep failed");
      return rc ? rc : GBP_ERR_CUDA;
    }
    if (h->l2_window.num_bytes) {  // captured kernel nodes do not inherit the stream's access-policy window
      size_t n_nodes = 0;
      cudaGraphGetNodes(graph, nullptr, &n_nodes);
      std::vector<cudaGraphNode_t> nodes(n_nodes);
      if (n_nodes) cudaGraphGetNodes(graph, nodes.data(), &n_nodes);
      for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeAttrValue av;
        av.accessPolicyWindow = h->l2_window;
        cudaGraphKernelNodeSetAttribute(nd, cudaKernelNodeAttributeAccessPolicyWindow, &av);
      }
      cudaGetLastError();
    }
    PhaseTimer pti;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    pti.lap("  cudaGraphInstantiate");
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      exec = nullptr;
      gbp_set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
      return GBP_ERR_CUDA;
    }
    h->graph_g = h->g;
    h->graph_stats = h->d_stats;
    h->graph_n_active = h->n_active;
  }
  *out = exec;
  return GBP_OK;
}

}  // namespace

extern "C" {

const char* gbp_cuda_version(void) { return "gbp-b200 0.1 (sm_100a)"; }

void gbp_opts_default(gbp_opts* o) {
  std::memset(o, 0, sizeof(*o));
  o->device = 0;
  o->maxeta_damping = 0.4f;
  o->num_undamped_iters = 8;
  o->dmu_threshold = 3e-3f;
  o->min_linear_iters = 10;
  o->Nstds = 2.5f;
  o->use_cuda_graph = 1;
  const char* env = std::getenv("GBP_CUDA_GRAPH");  // "0" disables the graph replay (diagnostics)
  if (env && env[0] == '0') o->use_cuda_graph = 0;
}

int gbp_cuda_init(const gbp_problem* p, const gbp_opts* o_in, gbp_handle** out) {
  if (!p || !out) {
    gbp_set_error("null argument");
    return GBP_ERR_ARG;
  }
  if ((p->n_edges && (!p->cam_ids || !p->lmk_ids || !p->measurements || !p->meas_variances)) ||
      (p->n_keyframes && (!p->cam_priors_eta || !p->cam_priors_lambda || !p->cam_scaling || !p->cam_weaken_flag)) ||
      (p->n_points && (!p->lmk_priors_eta || !p->lmk_priors_lambda || !p->lmk_scaling || !p->lmk_weaken_flag))) {
    gbp_set_error("gbp_problem has a null required array");
    return GBP_ERR_ARG;
  }
  gbp_opts o;
  if (o_in) o = *o_in; else gbp_opts_default(&o);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    gbp_set_error("no CUDA device available (the GBP hot path has no CPU fallback)");
    return GBP_ERR_CUDA;
  }
  if (o.device < 0 || o.device >= ndev) {
    gbp_set_error("device ordinal out of range");
    return GBP_ERR_ARG;
  }
  gbp_handle* h = new gbp_handle();
  h->device = o.device;
  h->use_graph = o.use_cuda_graph;
  h->fast_math = o.fast_math ? 1 : 0;
  if (const char* env = std::getenv("GBP_SKIP_UPPER")) h->skip_upper = std::atoi(env) != 0;
  h->relin_mode = o.relin_mode;
  h->two_pass = (o.relin_mode == 2) ? 1 : 0;
  int rc = set_device(h);
  if (!rc && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) rc = GBP_ERR_CUDA;
  if (!rc && cudaEventCreate(&h->ev0) != cudaSuccess) rc = GBP_ERR_CUDA;
  if (!rc && cudaEventCreate(&h->ev1) != cudaSuccess) rc = GBP_ERR_CUDA;

  if (!rc) rc = build(h, p, &o);
  if (!rc) rc = linearise_prog(h);
  if (!rc) rc = linearise_wait(h);
  if (rc) {
    if (rc == GBP_ERR_CUDA && !*gbp_cuda_last_error()) gbp_set_error("CUDA initialisation failed");
    gbp_cuda_free(h);
    return rc;
  }
  *out = h;
  return GBP_OK;
}

int gbp_cuda_release_cached_memory(void) {
  std::lock_guard<std::mutex> lk(pools_mutex());
  for (const DevicePool& d : pools())
    if (d.pool) cudaMemPoolTrimTo(d.pool, 0);
  for (const PinnedBlock& b : pinned_free_list()) cudaFreeHost(b.p);
  pinned_free_list().clear();
  cudaGetLastError();
  return GBP_OK;
}

int gbp_cuda_free(gbp_handle* h) {
  if (!h) return GBP_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->l2_limit_changed) {
    // the last handle with a persisting window on this device un-pins the lines and restores the set-aside the
    // application had; while other handles are alive their windows are left alone
    std::lock_guard<std::mutex> lk(pools_mutex());
    L2State& ls = l2_state(h->device);
    if (--ls.users == 0) {
      cudaCtxResetPersistingL2Cache();
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ls.prev_limit);
      cudaGetLastError();
    }
  }
  for (void* p : h->allocs) cudaFree(p);
  for (void* p : h->pool_allocs) cudaFreeAsync(p, h->stream);  // stays in the pool for the next handle
  if (h->stream && !h->pool_allocs.empty()) cudaStreamSynchronize(h->stream);
  if (h->d_stats && !h->d_stats_in_arena) cudaFree(h->d_stats);
  pinned_put(h->pin_block, h->pin_block_bytes);  // recycled by the next handle
  for (cudaEvent_t ev : h->prof_events) cudaEventDestroy(ev);
  drop_graphs(h);
  if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);  // the communicator itself stays cached
  if (h->p2p && !h->group) {
    // peers may still be pushing into this rank's block: freeing a sharded handle is collective
    int* d_b = nullptr;
    if (cudaMalloc((void**)&d_b, sizeof(int) * (h->world + 1)) == cudaSuccess) {
      cudaMemsetAsync(d_b, 0, sizeof(int) * (h->world + 1), h->comm_stream);
      gbp::nccl_api().AllGather(d_b + h->world, d_b, sizeof(int), ncclChar, h->comm, h->comm_stream);
      cudaStreamSynchronize(h->comm_stream);
      cudaFree(d_b);
    }
    for (void* q : h->p2p_peers)
      if (q) cudaIpcCloseMemHandle(q);
    cudaFree(h->p2p_block);
  }
  if (h->group && --h->group->refs == 0) {  // the last handle of a single-process group frees every rank's block
    for (size_t r = 0; r < h->group->blocks.size(); ++r) {
      cudaSetDevice(h->group->devices[r]);
      cudaFree(h->group->blocks[r]);
    }
    cudaSetDevice(h->device);
    delete h->group;
  }
  if (h->ev_send) cudaEventDestroy(h->ev_send);
  if (h->ev_recv) cudaEventDestroy(h->ev_recv);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->shard) gbp_shard_free(h->shard);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return GBP_OK;
}

int gbp_cuda_dims(gbp_handle* h, uint32_t* C, uint32_t* L, uint32_t* E, uint32_t* mk, uint32_t* ml) {
  if (!h) return GBP_ERR_ARG;
  if (C) *C = h->C;
  if (L) *L = h->L;
  if (E) *E = h->E;
  if (mk) *mk = h->SK - 1;
  if (ml) *ml = h->SL - 1;
  return GBP_OK;
}

void* gbp_cuda_stream(gbp_handle* h) { return h ? (void*)h->stream : nullptr; }

int gbp_cuda_synchronize(gbp_handle* h) {
  if (!h) return GBP_ERR_ARG;
  int rc = set_device(h);
  if (rc) return rc;
  GBP_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return check_peer_error(h);
}

int gbp_cuda_weaken_priors(gbp_handle* h) {
  if (!h) return GBP_ERR_ARG;
  int rc = gbp_cuda_weaken_prior_vertices(h);
  if (rc) return rc;
  return launch_update_vars(h);
}

int gbp_cuda_weaken_prior_vertices(gbp_handle* h) {
  if (!h) return GBP_ERR_ARG;
  int rc = set_device(h);
  if (!rc) rc = priors_about_to_change(h);
  if (rc) return rc;
  const uint32_t n = h->C + h->L;
  if (n) {
    gbp::k_weaken<<<(n + 255) / 256, 256, 0, h->stream>>>(h->g);
    h->kernels_launched++;
  }
  GBP_CUDA_TRY(cudaGetLastError());
  return GBP_OK;
}

int gbp_cuda_iterate_async(gbp_handle* h, int n_sweeps) {
  if (!h || n_sweeps < 0) return GBP_ERR_ARG;
  int rc = set_device(h);
  for (int i = 0; i < n_sweeps && !rc; ++i) {
    const bool upper = i == n_sweeps - 1 || !can_skip_upper(h);
    rc = launch_full_sweep(h, upper);
    if (!rc) rc = launch_update_vars(h, !upper);
  }
  return rc;
}

}  // extern "C"

namespace {

// gbp_cuda_iterate in four phases, so that a single-process group can interleave the ranks (every rank's sweep i
// must be enqueued before any rank waits: the belief update of a sweep spins on its peers' partials):
//   iterate_begin    buffers, events, the CUDA graphs to replay
//   iterate_enqueue  sweep i of n
//   iterate_end      closing event + the asynchronous read-backs
//   iterate_finish   ONE stream synchronisation, results, the sweep-flavour choice
int iterate_begin(gbp_handle* h, int n_sweeps, bool stats) {
  int rc = set_device(h);
  if (!rc) rc = check_peer_error(h);
  if (rc) return rc;
  PhaseTimer pt;
  if (stats) {
    rc = ensure_stats(h, (size_t)n_sweeps);
    if (rc) return rc;
    pt.lap("iterate: stats buffers");
  }
  h->it_k0 = h->kernels_launched;
  h->it_prof = h->profile != 0;
  if (h->it_prof) {
    while (h->prof_events.size() < (size_t)3 * n_sweeps) {
      cudaEvent_t ev;
      GBP_CUDA_TRY(cudaEventCreate(&ev));
      h->prof_events.push_back(ev);
    }
  }
  // steady state: replay the captured sweep (profiling goes through the plain launches, as does a sharded handle
  // whose exchange is the NCCL all-gather on the communication stream)
  const bool graph_ok = !h->shard || h->p2p || h->g.n_bnd_global == 0;
  h->it_graph = h->use_graph && !h->it_prof && graph_ok && !(h->shard && stats && !h->p2p) && n_sweeps > 0 && h->n_tiles;
  h->it_exec = h->it_exec_lower = nullptr;
  GBP_CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  if (h->it_graph) {
    if (stats) GBP_CUDA_TRY(cudaMemsetAsync(h->d_stat_cursor, 0, sizeof(uint32_t), h->stream));
    // every sweep but the last skips the strict upper triangle of the camera messages (see k_sweep)
    rc = sweep_graph(h, stats, true, &h->it_exec);
    // (not when per-sweep metrics are requested: the metric inverts the FULL camera belief, like the reference's)
    if (!rc && n_sweeps > 1 && !stats && can_skip_upper(h)) rc = sweep_graph(h, false, false, &h->it_exec_lower);
    if (rc) return rc;
    pt.lap("iterate: sweep graphs");
  }
  return GBP_OK;
}

int iterate_enqueue(gbp_handle* h, int i, int n_sweeps, bool stats) {
  int rc = set_device(h);
  if (rc) return rc;
  if (h->it_graph) {
    GBP_CUDA_TRY(cudaGraphLaunch((h->it_exec_lower && i + 1 < n_sweeps) ? h->it_exec_lower : h->it_exec, h->stream));
    h->kernels_launched += (uint64_t)((stats ? (h->shard ? 5 : 4) : 2) + (h->two_pass ? 2 : 0));
    h->pending_shift = false;
    h->p_in_sync = true;
    return GBP_OK;
  }
  const bool prof = h->it_prof;
  const bool upper = i == n_sweeps - 1 || stats || !can_skip_upper(h);
  if (prof) GBP_CUDA_TRY(cudaEventRecord(h->prof_events[3 * i], h->stream));
  rc = launch_full_sweep(h, upper);
  if (prof) GBP_CUDA_TRY(cudaEventRecord(h->prof_events[3 * i + 1], h->stream));
  if (!rc) rc = launch_update_vars(h, !upper);
  if (prof) GBP_CUDA_TRY(cudaEventRecord(h->prof_events[3 * i + 2], h->stream));
  if (!rc && stats) rc = launch_metric(h, h->d_stats + i);
  return rc;
}

int iterate_end(gbp_handle* h, int n_sweeps, bool stats) {
  int rc = set_device(h);
  if (rc) return rc;
  GBP_CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  if (stats && n_sweeps)
    GBP_CUDA_TRY(cudaMemcpyAsync(h->pin_stats, h->d_stats, (size_t)n_sweeps * sizeof(gbp_iter_stats), cudaMemcpyDeviceToHost,
                                 h->stream));
  // every 16 sweeps the relinearisation ring (132 bytes) rides along with the same synchronisation
  h->sweeps_since_choice += (uint32_t)n_sweeps;
  if (h->sweeps_since_choice >= 16 && h->E && h->relin_mode == 0) {
    GBP_CUDA_TRY(cudaMemcpyAsync(h->pin_ring, h->g.relin_ring, (GBP_RELIN_RING + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                 h->stream));
  }
  return GBP_OK;
}

int iterate_finish(gbp_handle* h, int n_sweeps, gbp_iter_stats* stats) {
  int rc = set_device(h);
  if (rc) return rc;
  GBP_CUDA_TRY(cudaStreamSynchronize(h->stream));
  rc = check_peer_error(h);
  if (rc) return rc;
  if (stats && n_sweeps) std::memcpy(stats, h->pin_stats, (size_t)n_sweeps * sizeof(gbp_iter_stats));
  GBP_CUDA_TRY(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  h->last_kernels = h->kernels_launched - h->it_k0;
  if (h->sweeps_since_choice >= 16 && h->E) {
    h->sweeps_since_choice = 0;
    rc = choose_sweep_flavour(h, h->relin_mode == 0 ? h->pin_ring : nullptr);
    if (rc) return rc;
  }
  h->last_ms_factor = h->last_ms_variable = 0.f;
  if (h->it_prof) {
    double a = 0, b = 0;
    h->sweep_ms_factor.assign((size_t)n_sweeps, 0.f);
    h->sweep_ms_variable.assign((size_t)n_sweeps, 0.f);
    for (int i = 0; i < n_sweeps; ++i) {
      float t;
      GBP_CUDA_TRY(cudaEventElapsedTime(&t, h->prof_events[3 * i], h->prof_events[3 * i + 1]));
      a += t;
      h->sweep_ms_factor[(size_t)i] = t;
      GBP_CUDA_TRY(cudaEventElapsedTime(&t, h->prof_events[3 * i + 1], h->prof_events[3 * i + 2]));
      b += t;
      h->sweep_ms_variable[(size_t)i] = t;
    }
    h->last_ms_factor = (float)a;
    h->last_ms_variable = (float)b;
  }
  return GBP_OK;
}

}  // namespace

extern "C" {

int gbp_cuda_iterate(gbp_handle* h, int n_sweeps, gbp_iter_stats* stats) {
  if (!h || n_sweeps < 0) return GBP_ERR_ARG;
  int rc = iterate_begin(h, n_sweeps, stats != nullptr);
  for (int i = 0; i < n_sweeps && !rc; ++i) rc = iterate_enqueue(h, i, n_sweeps, stats != nullptr);
  if (!rc) rc = iterate_end(h, n_sweeps, stats != nullptr);
  if (!rc) rc = iterate_finish(h, n_sweeps, stats);
  return rc;
}

int gbp_cuda_iterate_until(gbp_handle* h, int max_sweeps, int check_every, float rel_tol, float diverge_factor,
                           gbp_iter_stats* stats, int* n_done, int* stop_reason) {
  if (!h || max_sweeps < 0 || check_every <= 0 || !n_done || !stop_reason) {
    gbp_set_error("bad gbp_cuda_iterate_until arguments");
    return GBP_ERR_ARG;
  }
  std::vector<gbp_iter_stats> block((size_t)check_every);
  float run_min = 0.f, ms_total = 0.f;
  float last_block_end = -1.f;
  uint64_t kernels = 0;
  bool have_min = false;
  int done = 0, reason = GBP_STOP_MAX_SWEEPS;
  while (done < max_sweeps) {
    const int m = std::min(check_every, max_sweeps - done);
    int rc = gbp_cuda_iterate(h, m, block.data());
    if (rc) return rc;
    ms_total += h->last_ms;
    kernels += h->last_kernels;
    if (stats) std::memcpy(stats + done, block.data(), (size_t)m * sizeof(gbp_iter_stats));
    done += m;
    bool diverged = false;
    for (int i = 0; i < m; ++i) {
      const float err = block[(size_t)i].reproj_mean;
      if (!(err == err)) diverged = true;  // NaN
      if (!have_min || err < run_min) { run_min = err; have_min = true; }
      if (diverge_factor > 0.f && err > diverge_factor * run_min) diverged = true;
    }
    if (diverged) {
      reason = GBP_STOP_DIVERGED;
      break;
    }
    const float end = block[(size_t)m - 1].reproj_mean;
    // converged: the error still went down (or stood still) over the block, by less than rel_tol.  A block over
    // which it ROSE is not convergence: the run goes on until the rise trips the divergence rule or max_sweeps
    const float gain = last_block_end - end;
    if (m == check_every && last_block_end >= 0.f && gain >= 0.f && gain < rel_tol * last_block_end) {
      reason = GBP_STOP_CONVERGED;
      break;
    }
    last_block_end = end;
  }
  h->last_ms = ms_total;
  h->last_kernels = kernels;
  *n_done = done;
  *stop_reason = reason;
  return GBP_OK;
}

int gbp_cuda_set_profile(gbp_handle* h, int enabled) {
  if (!h) return GBP_ERR_ARG;
  h->profile = enabled ? 1 : 0;
  return GBP_OK;
}

int gbp_cuda_last_kernel_times(gbp_handle* h, float* ms_factor, float* ms_variable) {
  if (!h) return GBP_ERR_ARG;
  if (ms_factor) *ms_factor = h->last_ms_factor;
  if (ms_variable) *ms_variable = h->last_ms_variable;
  return GBP_OK;
}

int gbp_cuda_last_sweep_times(gbp_handle* h, float* ms_factor_kernel, float* ms_variable_kernel, int capacity, int* n_sweeps) {
  if (!h || capacity < 0) return GBP_ERR_ARG;
  const int n = (int)std::min<size_t>(h->sweep_ms_factor.size(), (size_t)capacity);
  for (int i = 0; i < n; ++i) {
    if (ms_factor_kernel) ms_factor_kernel[i] = h->sweep_ms_factor[(size_t)i];
    if (ms_variable_kernel) ms_variable_kernel[i] = h->sweep_ms_variable[(size_t)i];
  }
  if (n_sweeps) *n_sweeps = (int)h->sweep_ms_factor.size();
  return GBP_OK;
}

int gbp_cuda_debug_timestamps(gbp_handle* h, uint64_t* out, int capacity) {
  if (!h || !out || capacity < 0) return GBP_ERR_ARG;
  if (!h->g.dbg_ts) return 0;
  const int n = std::min(capacity, 32 * 8);
  if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess ||
      cudaMemcpy(out, h->g.dbg_ts, (size_t)n * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
    return GBP_ERR_CUDA;
  // re-arm the minima for the next window
  std::vector<unsigned long long> init(32 * 8, 0ull);
  for (int i = 0; i < 32; ++i) init[i * 8 + 0] = init[i * 8 + 2] = ~0ull;
  cudaMemcpy(h->g.dbg_ts, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
  return n;
}

int gbp_cuda_last_timing(gbp_handle* h, float* ms_total, uint64_t* kernels) {
  if (!h) return GBP_ERR_ARG;
  if (ms_total) *ms_total = h->last_ms;
  if (kernels) *kernels = h->last_kernels;
  return GBP_OK;
}

int gbp_cuda_eval(gbp_handle* h, gbp_iter_stats* out) {
  if (!h || !out) return GBP_ERR_ARG;
  int rc = set_device(h);
  if (!rc) rc = ensure_stats(h, 1);
  if (!rc) rc = launch_metric(h, h->d_stats);
  if (rc) return rc;
  GBP_CUDA_TRY(cudaMemcpyAsync(out, h->d_stats, sizeof(gbp_iter_stats), cudaMemcpyDeviceToHost, h->stream));
  GBP_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return check_peer_error(h);
}

int gbp_cuda_get_beliefs(gbp_handle* h, float* cam_eta, float* cam_lambda, float* lmk_eta, float* lmk_lambda,
                         float* damping, int32_t* damping_count, uint32_t* robust_flag) {
  if (!h) return GBP_ERR_ARG;
  int rc = set_device(h);
  if (rc) return rc;
  cudaStream_t s = h->stream;
  if (cam_eta) rc = download(cam_eta, h->g.cam_b_eta, 6 * (size_t)h->C, s);
  if (!rc && cam_lambda) rc = download(cam_lambda, h->g.cam_b_lam, 36 * (size_t)h->C, s);
  if (!rc && (lmk_eta || lmk_lambda) && h->L) {
    if (!rc) {
      gbp::k_export_lmk_beliefs<<<(h->L + 255) / 256, 256, 0, s>>>(h->g, h->d_exp_lmk_eta, h->d_exp_lmk_lam);
      h->kernels_launched++;
      if (lmk_eta) rc = download(lmk_eta, h->d_exp_lmk_eta, 3 * (size_t)h->L, s);
      if (!rc && lmk_lambda) rc = download(lmk_lambda, h->d_exp_lmk_lam, 9 * (size_t)h->L, s);
    }
  }
  if (!rc && (damping || damping_count || robust_flag) && h->E) {
    if (!rc) {
      gbp::k_export_edges<<<(h->E + 255) / 256, 256, 0, s>>>(h->g, h->d_pos_of_orig, h->d_exp_damping,
                                                               h->d_exp_dcount, h->d_exp_robust);
      h->kernels_launched++;
      if (damping) rc = download(damping, h->d_exp_damping, (size_t)h->E, s);
      if (!rc && damping_count) rc = download(damping_count, h->d_exp_dcount, (size_t)h->E, s);
      if (!rc && robust_flag) rc = download(robust_flag, h->d_exp_robust, (size_t)h->E, s);
    }
  }
  if (rc) return rc;
  GBP_CUDA_TRY(cudaGetLastError());
  GBP_CUDA_TRY(cudaStreamSynchronize(s));
  return check_peer_error(h);
}

int gbp_cuda_get_priors(gbp_handle* h, float* cam_eta, float* cam_lambda, float* lmk_eta, float* lmk_lambda) {
  if (!h) return GBP_ERR_ARG;
  int rc = set_device(h);
  if (rc) return rc;
  cudaStream_t s = h->stream;
  if (cam_eta) rc = download(cam_eta, h->g.cam_prior_eta, 6 * (size_t)h->C, s);
  if (!rc && cam_lambda) rc = download(cam_lambda, h->g.cam_prior_lam, 36 * (size_t)h->C, s);
  std::vector<float4> pr;
  if (!rc && (lmk_eta || lmk_lambda)) {
    pr.resize((size_t)h->L * 3);
    rc = download(pr.data(), h->g.lmk_prior, pr.size(), s);
  }
  if (rc) return rc;
  GBP_CUDA_TRY(cudaStreamSynchronize(s));
  for (size_t l = 0; l < h->L && !pr.empty(); ++l) {
    const float* r = &pr[l * 3].x;
    if (lmk_eta) for (int i = 0; i < 3; ++i) lmk_eta[l * 3 + i] = r[i];
    if (lmk_lambda) for (int i = 0; i < 9; ++i) lmk_lambda[l * 9 + i] = r[3 + i];
  }
  return GBP_OK;
}

int gbp_cuda_add_keyframe(gbp_handle* h, const int32_t* damping_count, const float* cam_prior_eta,
                          const float* cam_prior_lambda, const float* lmk_prior_eta, const float* lmk_prior_lambda,
                          const uint32_t* active_flag, const uint32_t* cam_weaken_flag,
                          const uint32_t* lmk_weaken_flag) {
  if (!h) return GBP_ERR_ARG;
  if (h->shard) {
    gbp_set_error("add_keyframe (incremental SLAM) is single-GPU only");
    return GBP_ERR_ARG;
  }
  int rc = set_device(h);
  if (rc) return rc;
  cudaStream_t s = h->stream;
  rc = priors_about_to_change(h);
  if (!rc && cam_prior_eta) rc = upload(h->g.cam_prior_eta, cam_prior_eta, 6 * (size_t)h->C, s);
  if (!rc && cam_prior_lambda) rc = upload(h->g.cam_prior_lam, cam_prior_lambda, 36 * (size_t)h->C, s);
  if (!rc && (lmk_prior_eta || lmk_prior_lambda)) rc = upload_lmk_priors(h, lmk_prior_eta, lmk_prior_lambda);
  if (!rc && cam_weaken_flag) rc = upload(h->g.cam_wflag, cam_weaken_flag, (size_t)h->C, s);
  if (!rc && lmk_weaken_flag) rc = upload(h->g.lmk_wflag, lmk_weaken_flag, (size_t)h->L, s);
  if (!rc && (damping_count || active_flag)) rc = import_edges(h, nullptr, damping_count, active_flag, nullptr, nullptr, 0);
  if (!rc) rc = launch_update_vars(h);  // prog_ub at the end of NEW_KEYFRAME (slam.cpp:928)
  if (rc) return rc;
  GBP_CUDA_TRY(cudaStreamSynchronize(s));
  return GBP_OK;
}

int gbp_cuda_add_keyframe_device(gbp_handle* h, uint32_t new_cam, uint32_t steps, int* n_new_lmks) {
  if (!h) return GBP_ERR_ARG;
  if (h->shard) {
    gbp_set_error("add_keyframe (incremental SLAM) is single-GPU only");
    return GBP_ERR_ARG;
  }
  if (new_cam < 1 || new_cam >= h->C) {
    gbp_set_error("add_keyframe_device: keyframe index out of range");
    return GBP_ERR_ARG;
  }
  int rc = set_device(h);
  if (rc) return rc;
  cudaStream_t s = h->stream;
  rc = priors_about_to_change(h);
  if (rc) return rc;
  gbp::k_kf_pose<<<1, 32, 0, s>>>(h->g, new_cam, h->d_kf_scratch);
  const uint32_t n = std::max(std::max(h->E_pad, h->L), h->C);
  gbp::k_kf_apply<<<(n + GBP_TILE - 1) / GBP_TILE, GBP_TILE, 0, s>>>(h->g, new_cam, steps, -15, h->d_lmk_first_cam,
                                                                    h->d_kf_scratch);
  h->kernels_launched += 2;
  GBP_CUDA_TRY(cudaGetLastError());
  rc = launch_update_vars(h);  // prog_ub at the end of NEW_KEYFRAME (slam.cpp:928)
  if (rc) return rc;
  uint32_t status = 0;
  GBP_CUDA_TRY(cudaMemcpyAsync(&status, h->d_kf_scratch + 3, sizeof(status), cudaMemcpyDeviceToHost, s));
  GBP_CUDA_TRY(cudaStreamSynchronize(s));
  if (status) {
    gbp_set_error("singular belief of the previous keyframe");
    return GBP_ERR_ARG;
  }
  // host mirrors of the flags (get_tensor("active_flag"), the metric's active-edge count)
  rc = ensure_host_maps(h);
  if (rc) return rc;
  uint32_t newly = 0;
  for (uint32_t k = h->cam_edge_ptr[new_cam]; k < h->cam_edge_ptr[new_cam + 1]; ++k) {
    const uint32_t e = h->cam_edge_ids[k];
    if (h->active_host[e] != 1u) {
      h->active_host[e] = 1u;
      ++newly;
    }
  }
  h->n_active += newly;
  if (n_new_lmks) *n_new_lmks = (int)h->new_lmks_at_cam[new_cam];
  return GBP_OK;
}

int gbp_cuda_relinearise_factors(gbp_handle* h) {
  if (!h) return GBP_ERR_ARG;
  int rc = set_device(h);
  if (rc) return rc;
  if (h->n_tiles) {
    gbp::k_relinearise_all<<<h->n_tiles, GBP_TILE, 0, h->stream>>>(h->g);
    h->kernels_launched++;
  }
  GBP_CUDA_TRY(cudaGetLastError());
  return GBP_OK;
}

int gbp_cuda_prep_messages(gbp_handle* h) {
  if (!h) return GBP_ERR_ARG;
  int rc = set_device(h);
  if (rc) return rc;
  return launch_prep(h);
}

int gbp_cuda_compute_messages(gbp_handle* h) {
  if (!h) return GBP_ERR_ARG;
  int rc = set_device(h);
  if (rc) return rc;
  return launch_sweep<false, true>(h);
}

int gbp_cuda_update_beliefs(gbp_handle* h) {
  if (!h) return GBP_ERR_ARG;
  int rc = set_device(h);
  if (rc) return rc;
  return launch_update_vars(h);
}

int gbp_cuda_tensor_nbytes(gbp_handle* h, const char* name, size_t* nbytes) {
  if (!h || !name || !nbytes) return GBP_ERR_ARG;
  const TensorId id = tensor_id(name);
  if (id == T_NONE) {
    gbp_set_error(std::string("unknown tensor ") + name);
    return GBP_ERR_NAME;
  }
  *nbytes = tensor_elems(h, id) * 4;
  return GBP_OK;
}

int gbp_cuda_get_tensor(gbp_handle* h, const char* name, void* dst, size_t nbytes) {
  if (!h || !name || !dst) return GBP_ERR_ARG;
  const TensorId id = tensor_id(name);
  if (id == T_NONE) {
    gbp_set_error(std::string("unknown tensor ") + name);
    return GBP_ERR_NAME;
  }
  if (nbytes != tensor_elems(h, id) * 4) {
    gbp_set_error(std::string("size mismatch for tensor ") + name);
    return GBP_ERR_SIZE;
  }
  int rc = set_device(h);
  if (!rc) rc = ensure_host_maps(h);
  if (rc) return rc;
  cudaStream_t s = h->stream;
  float* out = (float*)dst;
  uint32_t* outu = (uint32_t*)dst;
  int32_t* outi = (int32_t*)dst;
  const size_t C = h->C, L = h->L, E = h->E, EP = h->E_pad, SK = h->SK, SL = h->SL;
  const bool stale_p = (name[0] == 'p') && !h->p_in_sync;  // slot 0 of a p*-tensor, see gbp_handle
  std::vector<float4> v;
  switch (id) {
    case T_CAM_B_ETA: rc = download(out, h->g.cam_b_eta, 6 * C, s); break;
    case T_CAM_B_LAM: rc = download(out, h->g.cam_b_lam, 36 * C, s); break;
    case T_CAM_SCALING: rc = download(out, h->g.cam_scaling, C, s); break;
    case T_LMK_SCALING: rc = download(out, h->g.lmk_scaling, L, s); break;
    case T_CAM_WFLAG: rc = download(outu, h->g.cam_wflag, C, s); break;
    case T_LMK_WFLAG: rc = download(outu, h->g.lmk_wflag, L, s); break;
    case T_LMK_B_ETA:
    case T_LMK_B_LAM: {
      rc = fetch(h, v, h->g.lmk_b, L * GBP_LMKB_QUADS);
      if (rc) break;
      for (size_t l = 0; l < L; ++l) {
        const float* r = &v[l * GBP_LMKB_QUADS].x;
        if (id == T_LMK_B_ETA) for (int i = 0; i < 3; ++i) out[l * 3 + i] = r[i];
        else for (int i = 0; i < 9; ++i) out[l * 9 + i] = r[3 + i];
      }
      break;
    }
    case T_CAM_M_ETA:
    case T_CAM_M_LAM: {
      const int d = (id == T_CAM_M_ETA) ? 6 : 36;
      std::memset(out, 0, nbytes);
      std::vector<float> pr(d * C);
      if (stale_p) rc = download(pr.data(), id == T_CAM_M_ETA ? h->d_pprior_cam_eta : h->d_pprior_cam_lam, pr.size(), s);
      else rc = download(pr.data(), id == T_CAM_M_ETA ? h->g.cam_prior_eta : h->g.cam_prior_lam, pr.size(), s);
      if (!rc) rc = fetch(h, v, h->g.mcam, GBP_MCAM_QUADS * EP);
      std::vector<float4> vu;
      if (!rc && h->g.mcam_up && id == T_CAM_M_LAM) rc = fetch(h, vu, h->g.mcam_up, 4 * EP);
      if (rc) break;
      for (size_t c = 0; c < C; ++c) std::memcpy(out + c * SK * d, pr.data() + c * d, d * 4);
      for (size_t e = 0; e < E; ++e) {
        float* o = out + ((size_t)h->cam_ids[e] * SK + h->slot_c[e] + 1) * d;
        const size_t pos = h->pos_of_orig[e];
        if (id == T_CAM_M_ETA) {
          for (int i = 0; i < 6; ++i) o[i] = quad_field(v, EP, pos, i);
        } else {
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j)
              o[i * 6 + j] = (j > i && !vu.empty()) ? quad_field(vu, EP, pos, gbp_upper(i, j))
                                                    : quad_field(v, EP, pos, gbp_mcam_lam_field(i, j));
        }
      }
      break;
    }
    case T_LMK_M_ETA:
    case T_LMK_M_LAM: {
      const int d = (id == T_LMK_M_ETA) ? 3 : 9;
      const int off = (id == T_LMK_M_ETA) ? 0 : 3;
      std::memset(out, 0, nbytes);
      std::vector<float4> pr;
      rc = fetch(h, pr, stale_p ? h->d_pprior_lmk : h->g.lmk_prior, L * 3);
      if (!rc) rc = fetch(h, v, h->g.mlmk, GBP_MLMK_QUADS * E);
      if (rc) break;
      for (size_t l = 0; l < L; ++l)
        for (int i = 0; i < d; ++i) out[l * SL * d + i] = aos_field(pr, 3, l, off + i);
      for (size_t e = 0; e < E; ++e) {
        float* o = out + ((size_t)h->lmk_ids[e] * SL + h->slot_l[e] + 1) * d;
        const size_t k = (size_t)h->lmk_ptr[h->lmk_ids[e]] + h->slot_l[e];
        for (int i = 0; i < d; ++i) o[i] = aos_field(v, GBP_MLMK_QUADS, k, off + i);
      }
      break;
    }
    case T_F_ETA:
    case T_F_LAM: {
      rc = fetch(h, v, h->g.fac, GBP_FAC_QUADS * EP);
      if (rc) break;
      for (size_t e = 0; e < E; ++e) {
        const size_t pos = h->pos_of_orig[e];
        if (id == T_F_ETA) {
          for (int i = 0; i < 9; ++i) out[e * 9 + i] = quad_field(v, EP, pos, GBP_FAC_ETA + i);
        } else {  // block-packed [cc 36 | cl 18 | lc 18 | ll 9], ba/ba.cpp:93-96
          float* o = out + e * 81;
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) o[i * 6 + j] = quad_field(v, EP, pos, GBP_FAC_CC + gbp_sym(i, j));
          for (int i = 0; i < 18; ++i) o[36 + i] = quad_field(v, EP, pos, GBP_FAC_CL + i);
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 6; ++j) o[54 + i * 6 + j] = quad_field(v, EP, pos, GBP_FAC_CL + j * 3 + i);
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) o[72 + i * 3 + j] = quad_field(v, EP, pos, GBP_FAC_LL + gbp_sym(i, j));
        }
      }
      break;
    }
    case T_DAMPING: case T_DCOUNT: case T_DMU: case T_ACTIVE: case T_ROBUST: {
      rc = fetch(h, v, h->g.recA, EP);
      if (rc) break;
      for (size_t e = 0; e < E; ++e) {
        const float4 r = v[h->pos_of_orig[e]];
        if (id == T_DAMPING) out[e] = r.x;
        else if (id == T_DCOUNT) outi[e] = f2i(r.y);
        else if (id == T_DMU) out[e] = r.w;
        else if (id == T_ACTIVE) outu[e] = h->active_host[e];
        else outu[e] = (f2u(r.z) & GBP_FLAG_ROBUST) ? 1u : 0u;
      }
      break;
    }
    case T_Z: {
      rc = fetch(h, v, h->g.recB, EP);
      if (rc) break;
      for (size_t e = 0; e < E; ++e) {
        const float4 r = v[h->pos_of_orig[e]];
        out[2 * e] = r.x; out[2 * e + 1] = r.y;
      }
      break;
    }
    case T_VAR: {
      std::vector<float> vv(EP);
      rc = download(vv.data(), h->g.var, EP, s);
      if (!rc) GBP_CUDA_TRY(cudaStreamSynchronize(s));
      for (size_t e = 0; e < E && !rc; ++e) out[e] = vv[h->pos_of_orig[e]];
      break;
    }
    case T_MU: case T_OLDMU: {
      // mu == oldmu after a full sweep (Copy(mu,oldmu), ba.cpp:898).  For an edge
      // whose prep has run, both equal the variable means that prep used.
      std::vector<float> cm(6 * C);
      std::vector<float4> lm, lb;
      const bool pend = h->pending_shift;
      rc = download(cm.data(), pend ? h->g.cam_mean : h->g.cam_mean_prev, cm.size(), s);
      if (!rc) rc = fetch(h, v, h->g.recA, EP);
      if (!rc && pend) rc = fetch(h, lb, h->g.lmk_b, L * GBP_LMKB_QUADS);
      if (!rc && !pend) rc = fetch(h, lm, h->g.lmk_mean_prev, L);
      if (rc) break;
      const std::vector<float>& init = (id == T_MU) ? h->mu_init : h->oldmu_init;
      for (size_t e = 0; e < E; ++e) {
        float* o = out + e * 9;
        if (f2u(v[h->pos_of_orig[e]].z) & GBP_FLAG_MUVALID) {
          const size_t c = h->cam_ids[e], l = h->lmk_ids[e];
          for (int i = 0; i < 6; ++i) o[i] = cm[c * 6 + i];
          const float* m = pend ? &lb[l * GBP_LMKB_QUADS + 3].x : &lm[l].x;
          for (int i = 0; i < 3; ++i) o[6 + i] = m[i];
        } else {
          for (int i = 0; i < 9; ++i) o[i] = init.empty() ? 0.f : init[e * 9 + i];
        }
      }
      break;
    }
    default: return GBP_ERR_NAME;
  }
  if (rc) return rc;
  GBP_CUDA_TRY(cudaStreamSynchronize(s));
  return check_peer_error(h);
}

int gbp_cuda_set_tensor(gbp_handle* h, const char* name, const void* src, size_t nbytes) {
  if (!h || !name || !src) return GBP_ERR_ARG;
  const TensorId id = tensor_id(name);
  if (id == T_NONE) {
    gbp_set_error(std::string("unknown tensor ") + name);
    return GBP_ERR_NAME;
  }
  if (nbytes != tensor_elems(h, id) * 4) {
    gbp_set_error(std::string("size mismatch for tensor ") + name);
    return GBP_ERR_SIZE;
  }
  int rc = set_device(h);
  if (!rc) rc = ensure_host_maps(h);
  if (rc) return rc;
  cudaStream_t s = h->stream;
  const float* in = (const float*)src;
  const uint32_t* inu = (const uint32_t*)src;
  const int32_t* ini = (const int32_t*)src;
  const size_t C = h->C, L = h->L, E = h->E, EP = h->E_pad, SK = h->SK, SL = h->SL;
  const bool is_p = (name[0] == 'p');
  if (id == T_CAM_M_ETA || id == T_CAM_M_LAM || id == T_LMK_M_ETA || id == T_LMK_M_LAM) {
    // slot 0: a p*-tensor writes the bookkeeping copy of the priors, a message tensor the priors
    rc = priors_about_to_change(h);
    if (rc) return rc;
    if (is_p) h->p_in_sync = false;
  }
  std::vector<float4> v;
  switch (id) {
    case T_CAM_B_ETA: rc = upload(h->g.cam_b_eta, in, 6 * C, s); if (!rc) rc = recompute_means(h); break;
    case T_CAM_B_LAM: rc = upload(h->g.cam_b_lam, in, 36 * C, s); if (!rc) rc = recompute_means(h); break;
    case T_CAM_SCALING: rc = upload(h->g.cam_scaling, in, C, s); break;
    case T_LMK_SCALING: rc = upload(h->g.lmk_scaling, in, L, s); break;
    case T_CAM_WFLAG: rc = upload(h->g.cam_wflag, inu, C, s); break;
    case T_LMK_WFLAG: rc = upload(h->g.lmk_wflag, inu, L, s); break;
    case T_LMK_B_ETA:
    case T_LMK_B_LAM: {
      rc = fetch(h, v, h->g.lmk_b, L * GBP_LMKB_QUADS);
      if (rc) break;
      for (size_t l = 0; l < L; ++l) {
        float* r = &v[l * GBP_LMKB_QUADS].x;
        if (id == T_LMK_B_ETA) for (int i = 0; i < 3; ++i) r[i] = in[l * 3 + i];
        else for (int i = 0; i < 9; ++i) r[3 + i] = in[l * 9 + i];
      }
      rc = push(h, v, h->g.lmk_b);
      if (!rc) rc = recompute_means(h);
      break;
    }
    case T_CAM_M_ETA:
    case T_CAM_M_LAM: {
      const int d = (id == T_CAM_M_ETA) ? 6 : 36;
      std::vector<float> pr(d * C);
      for (size_t c = 0; c < C; ++c) std::memcpy(pr.data() + c * d, in + c * SK * d, d * 4);
      if (is_p) rc = upload(id == T_CAM_M_ETA ? h->d_pprior_cam_eta : h->d_pprior_cam_lam, pr.data(), pr.size(), s);
      else rc = upload(id == T_CAM_M_ETA ? h->g.cam_prior_eta : h->g.cam_prior_lam, pr.data(), pr.size(), s);
      if (!rc) GBP_CUDA_TRY(cudaStreamSynchronize(s));
      if (!rc) rc = fetch(h, v, h->g.mcam, GBP_MCAM_QUADS * EP);
      std::vector<float4> vu;
      if (!rc && h->g.mcam_up && id == T_CAM_M_LAM) rc = fetch(h, vu, h->g.mcam_up, 4 * EP);
      if (rc) break;
      for (size_t e = 0; e < E; ++e) {
        const float* o = in + ((size_t)h->cam_ids[e] * SK + h->slot_c[e] + 1) * d;
        const size_t pos = h->pos_of_orig[e];
        if (id == T_CAM_M_ETA) {
          for (int i = 0; i < 6; ++i) quad_field(v, EP, pos, i) = o[i];
        } else {
          for (int i = 0; i < 6; ++i)  // the lower triangle is what the algorithm reads (gbp_layout.h)
            for (int j = 0; j <= i; ++j) quad_field(v, EP, pos, gbp_mcam_lam_field(i, j)) = o[i * 6 + j];
          if (!vu.empty())
            for (int i = 0; i < 6; ++i)
              for (int j = i + 1; j < 6; ++j) quad_field(vu, EP, pos, gbp_upper(i, j)) = o[i * 6 + j];
        }
      }
      rc = push(h, v, h->g.mcam);
      if (!rc && !vu.empty()) rc = push(h, vu, h->g.mcam_up);
      if (!rc && h->n_tiles) {
        gbp::k_cam_partials<<<h->n_tiles, GBP_TILE, 0, s>>>(h->g);
        h->kernels_launched++;
      }
      break;
    }
    case T_LMK_M_ETA:
    case T_LMK_M_LAM: {
      const int d = (id == T_LMK_M_ETA) ? 3 : 9;
      const int off = (id == T_LMK_M_ETA) ? 0 : 3;
      std::vector<float4> pr;
      float4* d_pr = is_p ? h->d_pprior_lmk : h->g.lmk_prior;
      rc = fetch(h, pr, d_pr, L * 3);
      if (!rc) rc = fetch(h, v, h->g.mlmk, GBP_MLMK_QUADS * E);
      if (rc) break;
      for (size_t l = 0; l < L; ++l)
        for (int i = 0; i < d; ++i) aos_field(pr, 3, l, off + i) = in[l * SL * d + i];
      for (size_t e = 0; e < E; ++e) {
        const float* o = in + ((size_t)h->lmk_ids[e] * SL + h->slot_l[e] + 1) * d;
        const size_t k = (size_t)h->lmk_ptr[h->lmk_ids[e]] + h->slot_l[e];
        for (int i = 0; i < d; ++i) aos_field(v, GBP_MLMK_QUADS, k, off + i) = o[i];
      }
      rc = push(h, pr, d_pr);
      if (!rc) rc = push(h, v, h->g.mlmk);
      break;
    }
    case T_F_ETA:
    case T_F_LAM: {
      rc = fetch(h, v, h->g.fac, GBP_FAC_QUADS * EP);
      if (rc) break;
      for (size_t e = 0; e < E; ++e) {
        const size_t pos = h->pos_of_orig[e];
        if (id == T_F_ETA) {
          for (int i = 0; i < 9; ++i) quad_field(v, EP, pos, GBP_FAC_ETA + i) = in[e * 9 + i];
        } else {  // Lambda_lc (o[54..71]) is implied by Lambda_cl; Lambda_cc / Lambda_ll are symmetric
          const float* o = in + e * 81;     // (gbp_layout.h): their lower triangles are stored
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j <= i; ++j) quad_field(v, EP, pos, GBP_FAC_CC + gbp_lt(i, j)) = o[i * 6 + j];
          for (int i = 0; i < 18; ++i) quad_field(v, EP, pos, GBP_FAC_CL + i) = o[36 + i];
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j <= i; ++j) quad_field(v, EP, pos, GBP_FAC_LL + gbp_lt(i, j)) = o[72 + i * 3 + j];
        }
      }
      rc = push(h, v, h->g.fac);
      break;
    }
    case T_DAMPING: rc = import_edges(h, in, nullptr, nullptr, nullptr, nullptr, 0); break;
    case T_DCOUNT: rc = import_edges(h, nullptr, ini, nullptr, nullptr, nullptr, 0); break;
    case T_ACTIVE: rc = import_edges(h, nullptr, nullptr, inu, nullptr, nullptr, 0); break;
    case T_ROBUST: rc = import_edges(h, nullptr, nullptr, nullptr, inu, nullptr, 0); break;
    case T_DMU: rc = import_edges(h, nullptr, nullptr, nullptr, nullptr, in, 0); break;
    case T_Z: {
      rc = fetch(h, v, h->g.recB, EP);
      if (rc) break;
      for (size_t e = 0; e < E; ++e) {
        float4& r = v[h->pos_of_orig[e]];
        r.x = in[2 * e]; r.y = in[2 * e + 1];
      }
      rc = push(h, v, h->g.recB);
      break;
    }
    case T_VAR: {
      std::vector<float> vv(EP, 1.f);
      for (size_t e = 0; e < E; ++e) vv[h->pos_of_orig[e]] = in[e];
      rc = upload(h->g.var, vv.data(), EP, s);
      if (!rc) GBP_CUDA_TRY(cudaStreamSynchronize(s));
      break;
    }
    case T_MU:
      // `mu` is pure output of PrepMessageVertex (overwritten before it is read);
      // remember it only so that get_tensor echoes it for edges that never ran.
      h->mu_init.assign(in, in + 9 * E);
      break;
    case T_OLDMU: {
      // per-edge oldmu: stored verbatim and used by the next prep of every edge
      h->oldmu_init.assign(in, in + 9 * E);
      if (!h->g.oldmu_edge) rc = h_alloc(h, &h->g.oldmu_edge, 9 * EP);
      if (rc) break;
      std::vector<float> t(9 * EP, 0.f);
      for (size_t e = 0; e < E; ++e)
        for (int i = 0; i < 9; ++i) t[(size_t)i * EP + h->pos_of_orig[e]] = in[9 * e + i];
      rc = upload(h->g.oldmu_edge, t.data(), t.size(), s);
      if (!rc) GBP_CUDA_TRY(cudaStreamSynchronize(s));
      if (!rc) rc = import_edges(h, nullptr, nullptr, nullptr, nullptr, nullptr, 1);
      h->pending_shift = false;
      break;
    }
    default: return GBP_ERR_NAME;
  }
  if (rc) return rc;
  GBP_CUDA_TRY(cudaGetLastError());
  GBP_CUDA_TRY(cudaStreamSynchronize(s));
  return GBP_OK;
}

int gbp_cuda_nccl_unique_id(void* id128) {
  if (!id128) return GBP_ERR_ARG;
  gbp::NcclApi& nc = gbp::nccl_api();
  if (!nc.load()) {
    gbp_set_error(nc.error);
    return GBP_ERR_COMM;
  }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  GBP_NCCL_TRY(nc.GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  return GBP_OK;
}

int gbp_cuda_init_shard(const gbp_problem* p, const gbp_opts* o_in, uint32_t world, uint32_t rank,
                        const void* nccl_unique_id, gbp_handle** out) {
  if (world == 1 && rank == 0) return gbp_cuda_init(p, o_in, out);
  if (!p || !out || world == 0 || rank >= world || !nccl_unique_id) {
    gbp_set_error("bad gbp_cuda_init_shard arguments");
    return GBP_ERR_ARG;
  }
  gbp_opts o;
  if (o_in) o = *o_in; else gbp_opts_default(&o);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    gbp_set_error("no CUDA device available (the GBP hot path has no CPU fallback)");
    return GBP_ERR_CUDA;
  }
  if (o.device < 0 || o.device >= ndev) {
    gbp_set_error("device ordinal out of range");
    return GBP_ERR_ARG;
  }
  gbp::NcclApi& nc = gbp::nccl_api();
  if (!nc.load()) {
    gbp_set_error(nc.error);
    return GBP_ERR_COMM;
  }
  gbp_shard* sh = nullptr;
  int rc = gbp_shard_build_view(p, world, rank, &sh);  // p's arrays are only read during this call
  if (rc) return rc;
  gbp_handle* h = new gbp_handle();
  h->device = o.device;
  h->use_graph = o.use_cuda_graph;
  h->fast_math = o.fast_math ? 1 : 0;
  if (const char* env = std::getenv("GBP_SKIP_UPPER")) h->skip_upper = std::atoi(env) != 0;
  h->relin_mode = o.relin_mode;
  h->two_pass = (o.relin_mode == 2) ? 1 : 0;
  h->shard = sh;
  h->world = world;
  h->rank = rank;
  rc = set_device(h);
  if (!rc && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) rc = GBP_ERR_CUDA;
  if (!rc && cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking) != cudaSuccess) rc = GBP_ERR_CUDA;
  if (!rc && cudaEventCreate(&h->ev0) != cudaSuccess) rc = GBP_ERR_CUDA;
  if (!rc && cudaEventCreate(&h->ev1) != cudaSuccess) rc = GBP_ERR_CUDA;

  if (!rc && cudaEventCreateWithFlags(&h->ev_send, cudaEventDisableTiming) != cudaSuccess) rc = GBP_ERR_CUDA;
  if (!rc && cudaEventCreateWithFlags(&h->ev_recv, cudaEventDisableTiming) != cudaSuccess) rc = GBP_ERR_CUDA;
  if (!rc) {
    for (const CommEntry& ce : comm_cache())
      if (ce.world == world && ce.rank == rank && ce.device == o.device && std::memcmp(ce.id, nccl_unique_id, 128) == 0)
        h->comm = ce.comm;
  }
  if (!rc && !h->comm) {
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, sizeof(id));
    ncclResult_t r = nc.CommInitRank(&h->comm, (int)world, id, (int)rank);
    if (r != ncclSuccess) {
      gbp_set_error(std::string("ncclCommInitRank: ") + nc.GetErrorString(r));
      h->comm = nullptr;
      rc = GBP_ERR_COMM;
    } else {
      CommEntry ce;
      std::memcpy(ce.id, nccl_unique_id, 128);
      ce.world = world; ce.rank = rank; ce.device = o.device; ce.comm = h->comm;
      comm_cache().push_back(ce);
    }
  }
  if (!rc) rc = build(h, gbp_shard_problem(sh), &o, gbp_shard_edge_global(sh));
  gbp_shard_detach_views(sh);  // the caller may free p's arrays after this call: the retained shard keeps no pointer into them
  if (!rc) rc = setup_p2p(h, o.exchange);
  if (!rc) rc = linearise_prog(h);
  if (!rc) rc = linearise_wait(h);
  if (rc) {
    if (rc == GBP_ERR_CUDA && !*gbp_cuda_last_error()) gbp_set_error("CUDA initialisation failed");
    gbp_cuda_free(h);
    return rc;
  }
  *out = h;
  return GBP_OK;
}

// ---- single-process group: `world` shard handles of one problem driven by one host thread -------------------
int gbp_cuda_init_group(const gbp_problem* p, const gbp_opts* o_in, uint32_t world, const int* devices, gbp_handle** out) {
  if (!p || !out || world == 0 || world > 32) {  // (the per-landmark rank masks of the exchange are 32 bits wide)
    gbp_set_error("bad gbp_cuda_init_group arguments");
    return GBP_ERR_ARG;
  }
  gbp_opts o;
  if (o_in) o = *o_in; else gbp_opts_default(&o);
  for (uint32_t r = 0; r < world; ++r) out[r] = nullptr;
  if (world == 1) {
    if (devices) o.device = devices[0];
    return gbp_cuda_init(p, &o, out);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    gbp_set_error("no CUDA device available (the GBP hot path has no CPU fallback)");
    return GBP_ERR_CUDA;
  }
  std::vector<int> dev(world);
  for (uint32_t r = 0; r < world; ++r) {
    dev[r] = devices ? devices[r] : (int)(r % (uint32_t)ndev);
    if (dev[r] < 0 || dev[r] >= ndev) {
      gbp_set_error("device ordinal out of range");
      return GBP_ERR_ARG;
    }
  }
  auto fail = [&](int rc) {
    if (rc == GBP_ERR_CUDA && !*gbp_cuda_last_error()) gbp_set_error("CUDA initialisation failed");
    for (uint32_t r = 0; r < world; ++r)
      if (out[r]) {
        cudaSetDevice(out[r]->device);
        cudaStreamSynchronize(out[r]->stream);
      }
    for (uint32_t r = 0; r < world; ++r)
      if (out[r]) {
        gbp_cuda_free(out[r]);
        out[r] = nullptr;
      }
    return rc;
  };
  int rc = GBP_OK;
  // every device of the group must be able to address every other one's memory
  for (uint32_t a = 0; a < world && !rc; ++a)
    for (uint32_t b = 0; b < world && !rc; ++b) {
      if (dev[a] == dev[b]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, dev[a], dev[b]);
      if (!can) {
        gbp_set_error("gbp_cuda_init_group: the devices of the group cannot access each other's memory");
        rc = GBP_ERR_COMM;
        break;
      }
      cudaSetDevice(dev[a]);
      const cudaError_t e = cudaDeviceEnablePeerAccess(dev[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        gbp_set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        rc = GBP_ERR_COMM;
      }
      cudaGetLastError();
    }
  if (rc) return rc;
  for (uint32_t r = 0; r < world && !rc; ++r) {
    gbp_shard* sh = nullptr;
    rc = gbp_shard_build_view(p, world, r, &sh);
    if (rc) break;
    gbp_handle* h = new gbp_handle();
    out[r] = h;
    h->device = dev[r];
    h->use_graph = o.use_cuda_graph;
    h->fast_math = o.fast_math ? 1 : 0;
    if (const char* env = std::getenv("GBP_SKIP_UPPER")) h->skip_upper = std::atoi(env) != 0;
    h->relin_mode = o.relin_mode;
    h->two_pass = (o.relin_mode == 2) ? 1 : 0;
    h->shard = sh;
    h->world = world;
    h->rank = r;
    rc = set_device(h);
    if (!rc && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) rc = GBP_ERR_CUDA;
    if (!rc && cudaEventCreate(&h->ev0) != cudaSuccess) rc = GBP_ERR_CUDA;
    if (!rc && cudaEventCreate(&h->ev1) != cudaSuccess) rc = GBP_ERR_CUDA;
    if (!rc) preload_kernels();
    gbp_opts or_ = o;
    or_.device = dev[r];
    if (!rc) rc = build(h, gbp_shard_problem(sh), &or_, gbp_shard_edge_global(sh));
    gbp_shard_detach_views(sh);
  }
  if (rc) return fail(rc);
  // Shards that SHARE a device (a test configuration: the whole multi-GPU protocol on a one-GPU box) can only make
  // progress while the blocks spinning for a peer leave room for that peer's kernels: refuse graphs whose
  // boundary would fill the device with waiting blocks.
  for (uint32_t a = 0; a < world; ++a)
    for (uint32_t b = a + 1; b < world; ++b)
      if (dev[a] == dev[b]) {
        const uint32_t n_x = (std::max(out[a]->g.n_bnd_local, out[b]->g.n_bnd_local) + GBP_LMK_PER_BLOCK - 1) / GBP_LMK_PER_BLOCK;
        // (a waiting block holds ~1/10 of an SM's registers, a sweep block of the peer needs ~9/10: two waiting
        // blocks on every SM would starve it; at most world - 1 ranks wait at a time)
        if (n_x * (world - 1) > (uint32_t)out[a]->num_sms) {
          gbp_set_error("gbp_cuda_init_group: too many boundary landmarks for shards that share one device");
          return fail(GBP_ERR_ARG);
        }
      }
  // one exchange block per rank on its device, addressed directly by every peer
  GroupShared* gs = new GroupShared();
  gs->blocks.assign(world, nullptr);
  gs->devices = dev;
  const size_t total = p2p_block_bytes(world, out[0]->g.n_bnd_global);
  for (uint32_t r = 0; r < world && !rc; ++r) {
    cudaSetDevice(dev[r]);
    if (cudaMalloc(&gs->blocks[r], total) != cudaSuccess || cudaMemset(gs->blocks[r], 0, total) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess)
      rc = GBP_ERR_CUDA;
  }
  if (rc) {
    for (uint32_t r = 0; r < world; ++r)
      if (gs->blocks[r]) {
        cudaSetDevice(dev[r]);
        cudaFree(gs->blocks[r]);
      }
    delete gs;
    cudaGetLastError();
    return fail(rc);
  }
  for (uint32_t r = 0; r < world; ++r) {
    out[r]->group = gs;
    gs->refs++;
  }
  for (uint32_t r = 0; r < world && !rc; ++r) {
    rc = set_device(out[r]);
    if (!rc) rc = p2p_wire(out[r], gs->blocks);
  }
  for (uint32_t r = 0; r < world && !rc; ++r) {
    rc = set_device(out[r]);
    if (!rc) rc = linearise_prog(out[r]);
  }
  for (uint32_t r = 0; r < world && !rc; ++r) {
    rc = set_device(out[r]);
    if (!rc) rc = linearise_wait(out[r]);
  }
  if (rc) return fail(rc);
  return GBP_OK;
}

int gbp_cuda_group_iterate(gbp_handle** hs, uint32_t world, int n_sweeps, gbp_iter_stats* stats) {
  if (!hs || world == 0 || n_sweeps < 0) return GBP_ERR_ARG;
  for (uint32_t r = 0; r < world; ++r)
    if (!hs[r]) return GBP_ERR_ARG;
  int rc = GBP_OK;
  for (uint32_t r = 0; r < world && !rc; ++r) rc = iterate_begin(hs[r], n_sweeps, stats != nullptr);
  // sweep by sweep over the ranks: no rank's launch queue can fill up with work that waits for a peer that has
  // not been enqueued yet
  for (int i = 0; i < n_sweeps && !rc; ++i)
    for (uint32_t r = 0; r < world && !rc; ++r) rc = iterate_enqueue(hs[r], i, n_sweeps, stats != nullptr);
  for (uint32_t r = 0; r < world && !rc; ++r) rc = iterate_end(hs[r], n_sweeps, stats != nullptr);
  // every rank reports the metric of the whole graph: rank 0's copy is returned
  std::vector<gbp_iter_stats> scratch(stats ? (size_t)n_sweeps : 0);
  for (uint32_t r = 0; r < world && !rc; ++r) rc = iterate_finish(hs[r], n_sweeps, stats ? (r == 0 ? stats : scratch.data()) : nullptr);
  return rc;
}

int gbp_cuda_group_weaken_priors(gbp_handle** hs, uint32_t world) {
  if (!hs || world == 0) return GBP_ERR_ARG;
  int rc = GBP_OK;
  for (uint32_t r = 0; r < world && !rc; ++r) rc = hs[r] ? gbp_cuda_weaken_priors(hs[r]) : GBP_ERR_ARG;  // enqueues only
  return rc;
}

int gbp_cuda_group_eval(gbp_handle** hs, uint32_t world, gbp_iter_stats* out) {
  if (!hs || world == 0 || !out) return GBP_ERR_ARG;
  int rc = GBP_OK;
  for (uint32_t r = 0; r < world && !rc; ++r) {
    gbp_handle* h = hs[r];
    if (!h) return GBP_ERR_ARG;
    rc = set_device(h);
    if (!rc) rc = ensure_stats(h, 1);
    if (!rc) rc = launch_metric(h, h->d_stats);
  }
  for (uint32_t r = 0; r < world && !rc; ++r) {
    gbp_handle* h = hs[r];
    rc = set_device(h);
    if (rc) break;
    gbp_iter_stats tmp;
    GBP_CUDA_TRY(cudaMemcpyAsync(r == 0 ? out : &tmp, h->d_stats, sizeof(gbp_iter_stats), cudaMemcpyDeviceToHost, h->stream));
    GBP_CUDA_TRY(cudaStreamSynchronize(h->stream));
    rc = check_peer_error(h);
  }
  return rc;
}

int gbp_cuda_group_free(gbp_handle** hs, uint32_t world) {
  if (!hs) return GBP_OK;
  for (uint32_t r = 0; r < world; ++r)
    if (hs[r]) {
      cudaSetDevice(hs[r]->device);
      if (hs[r]->stream) cudaStreamSynchronize(hs[r]->stream);
    }
  for (uint32_t r = 0; r < world; ++r)
    if (hs[r]) {
      gbp_cuda_free(hs[r]);
      hs[r] = nullptr;
    }
  return GBP_OK;
}

const gbp_shard* gbp_cuda_shard_info(gbp_handle* h) { return h ? h->shard : nullptr; }

int gbp_cuda_exchange_mode(gbp_handle* h) {
  if (!h || !h->shard || h->g.n_bnd_global == 0) return 0;
  return h->p2p ? 2 : 1;
}

}  // extern "C"
