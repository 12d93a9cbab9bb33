/*
 * gbp_cuda.h -- C ABI of the B200-native GBP bundle-adjustment hot path.
 *
 * This is the drop-in boundary for the Poplar graph/engine layer of
 * joeaortiz/gbp-poplar (reference paths below are relative to the reference
 * tree).  The reference host loop talks to its device through
 *   - numbered programs run with engine.run(id)        ba/ba.cpp:925-934, ba/slam.cpp:937-948
 *   - named FIFOs bound to caller-owned host buffers   ba/ba.cpp:940-976
 * Every entry point below replaces one of those programs / stream groups.
 * All pointers are caller-owned HOST memory in the reference's own (AoS,
 * padded) layouts; the library owns every device buffer.  No torch / C++
 * types cross this boundary.
 *
 * Return value: 0 = GBP_OK, negative = error (text via gbp_cuda_last_error()).
 * A handle is not thread-safe; distinct handles are independent.
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with GBP_ERR_CUDA.
 */
#ifndef GBP_CUDA_H_
#define GBP_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GBP_OK 0
#define GBP_ERR_ARG (-1)      /* bad argument / inconsistent problem        */
#define GBP_ERR_CUDA (-2)     /* CUDA runtime error or no device            */
#define GBP_ERR_NAME (-3)     /* unknown tensor name                        */
#define GBP_ERR_SIZE (-4)     /* nbytes does not match the tensor           */
#define GBP_ERR_IO (-5)       /* file could not be opened / parsed          */
#define GBP_ERR_COMM (-6)     /* multi-GPU exchange not initialised/failed  */

/* Fixed sizes of the reference's factor graph (ba/ba.cpp:576-578). */
#define GBP_CAM_DOFS 6
#define GBP_LMK_DOFS 3
#define GBP_FACTOR_DOFS 9

/*
 * One bundle-adjustment problem exactly as the reference's host code hands it
 * to WRITE_PROG (ba/ba.cpp:868-886, stream bindings ba/ba.cpp:940-961).
 * Arrays marked "may be NULL" default to what ba.cpp streams (zeros).
 */
typedef struct gbp_problem {
  uint32_t n_keyframes;               /* C  */
  uint32_t n_points;                  /* L  */
  uint32_t n_edges;                   /* E  */
  const uint32_t* cam_ids;            /* [E]   measurements_camIDs  ba/ba.cpp:505 */
  const uint32_t* lmk_ids;            /* [E]   measurements_lIDs    ba/ba.cpp:506 */
  const float* measurements;          /* [2E]  ba/ba.cpp:504 */
  const float* meas_variances;        /* [E]   ba/ba.cpp:503 */
  float K[9];                         /* shared 3x3 intrinsics, row-major; the reference
                                         replicates it per edge (ba/ba.cpp:494-501) */
  const float* cam_priors_eta;        /* [6C]  slot 0 of cam_messages_eta   ba/ba.cpp:880 */
  const float* cam_priors_lambda;     /* [36C] slot 0 of cam_messages_lambda ba/ba.cpp:881 */
  const float* lmk_priors_eta;        /* [3L]  ba/ba.cpp:882 */
  const float* lmk_priors_lambda;     /* [9L]  ba/ba.cpp:883 */
  const float* cam_scaling;           /* [C]   ba/ba.cpp:561-568 */
  const float* lmk_scaling;           /* [L]   ba/ba.cpp:569-572 */
  const uint32_t* cam_weaken_flag;    /* [C]   ba/ba.cpp:589 */
  const uint32_t* lmk_weaken_flag;    /* [L]   ba/ba.cpp:590 */
  const uint32_t* active_flag;        /* [E]   ba/ba.cpp:588; may be NULL = all 1 */
  const float* damping;               /* [E]   ba/ba.cpp:580; may be NULL = 0 */
  const int32_t* damping_count;       /* [E]   ba/ba.cpp:581; may be NULL = -15 */
  const float* mu;                    /* [9E]  ba/ba.cpp:582; may be NULL = 0 */
  const float* oldmu;                 /* [9E]  ba/ba.cpp:583; may be NULL = 0 */
} gbp_problem;

/* Hyper-parameters = the file-scope globals of ba/gbp_codelets.cpp:11-16. */
typedef struct gbp_opts {
  int device;                 /* CUDA device ordinal (default 0)                       */
  float maxeta_damping;       /* 0.4   gbp_codelets.cpp:11 */
  int num_undamped_iters;     /* 8     gbp_codelets.cpp:12 */
  float dmu_threshold;        /* 3e-3  gbp_codelets.cpp:13 */
  int min_linear_iters;       /* 10    gbp_codelets.cpp:14 */
  float Nstds;                /* 2.5   gbp_codelets.cpp:16 */
  int use_cuda_graph;         /* 1 (default): gbp_cuda_iterate replays a captured CUDA graph of one sweep (one launch
                                 per sweep instead of 2-5: what bounds the small graphs of the reference sequences) */
  int store_full_messages;    /* 0 (default): a factor->camera message keeps eta + the LOWER triangle of
                                 Lambda -- all the algorithm ever reads back (inv6x6 works on the lower
                                 triangle, matlib.cpp:193-206; the belief sum over the full 6x6 message is
                                 formed on chip) -- and get_tensor("cam_messages_lambda") mirrors it.
                                 1: the strict upper triangle is stored too (+64 B per factor and sweep),
                                 so that tensor is bit-identical to the reference's as well.  Beliefs,
                                 every other tensor and the trajectory are identical in both modes.     */
  int exchange;               /* multi-GPU boundary exchange: 0 = auto (peer-to-peer over NVLink -- the
                                 belief-update kernel reads the peers' step-tagged partial sums through
                                 CUDA-IPC-mapped memory -- when every rank can map every peer, else NCCL all-gather),
                                 1 = NCCL all-gather, 2 = peer-to-peer or fail                          */
  int relin_mode;             /* how a sweep handles the in-loop relinearisation (results are bit-identical):
                                 1 = one fused kernel; 2 = state-machine pass + compacted relinearisation +
                                 message-only kernel; 0 (default) = chosen from the pattern of the last sweeps
                                 (fused while relinearisations come in lock step, two-pass once they are spread) */
  int fast_math;              /* 0 (default): every fp32 operation of the sweep is the reference's, in its order, never
                                 contracted -- results are bit-identical to the reference's codelets.
                                 1: the sweep kernel built with contracted multiply-adds (one rounding instead of two
                                 per a*b+c; ~25 % fewer instructions).  NOT bit-comparable: GBP amplifies the rounding
                                 difference over the sweeps; validated at the north-star tolerance (one sweep <= 1e-4
                                 per block, plateau reprojection error within 1 %).  Single-GPU and sharded handles. */
  int reserved[3];
} gbp_opts;

/* Per-sweep metrics = what the reference's host computes after READ_PROG
 * (ba/ba.cpp:1011-1028, ba/util.cpp:74-144), evaluated on the device. */
typedef struct gbp_iter_stats {
  float reproj_mean;          /* mean |z - h(mu)| over active edges   util.cpp:126,143 */
  float cost;                 /* sum 0.5*|r|^2, un-robustified         util.cpp:127     */
  uint32_t n_relins;          /* #edges with damping_count == -num_undamped_iters  ba.cpp:1016-1020 */
  uint32_t n_robust;          /* sum of robust_flag over all edges     ba.cpp:1013-1015 */
  uint32_t n_active;          /* #active edges                          util.cpp:95-97   */
  uint32_t reserved;
} gbp_iter_stats;

typedef struct gbp_handle gbp_handle;

/* Sharding description for the multi-GPU path (one process per GPU).
 * rank r owns cameras [cam_begin, cam_end). */
typedef struct gbp_shard_plan {
  uint32_t world;             /* number of ranks                                   */
  uint32_t rank;              /* this rank                                         */
  uint32_t cam_begin;         /* first camera owned by this rank                   */
  uint32_t cam_end;           /* one past the last owned camera                    */
  uint32_t n_local_edges;     /* edges whose camera is owned here                  */
  uint32_t n_local_points;    /* landmarks touched by local edges                  */
  uint32_t n_boundary_points; /* landmarks observed from >1 rank (global count)    */
  uint32_t reserved;
} gbp_shard_plan;

const char* gbp_cuda_last_error(void);
const char* gbp_cuda_version(void);
void gbp_opts_default(gbp_opts* o);

/* Graph build + WRITE_PROG + LINEARISE_PROG (ba/ba.cpp:659-986): uploads the
 * problem, zero-fills all non-prior message slots (the reference relies on
 * zeroed device memory, ba/ba.cpp:868-886), sums the priors into the beliefs
 * and linearises every factor (RelineariseFactorVertex, no active check). */
int gbp_cuda_init(const gbp_problem* p, const gbp_opts* o, gbp_handle** out);

/* Engine teardown. */
int gbp_cuda_free(gbp_handle* h);

/* WEAKEN_PRIORS (ba/ba.cpp:863-865): WeakenPriorVertex on every variable,
 * then the belief update. */
int gbp_cuda_weaken_priors(gbp_handle* h);

/* GBP_PROG x n_sweeps (ba/ba.cpp:895-905).  If stats != NULL it must hold
 * n_sweeps entries; entry i is evaluated on the beliefs after sweep i.
 * When the call returns every tensor equals what n_sweeps x GBP_PROG leave behind.
 * With stats == NULL the sweeps before the last one do not form the strict upper
 * triangle of the factor->camera message Lambda (the algorithm never reads it back;
 * it only reaches cam_beliefs_lambda, which the last sweep of the call rebuilds). */
int gbp_cuda_iterate(gbp_handle* h, int n_sweeps, gbp_iter_stats* stats);

/* Convergence control on top of gbp_cuda_iterate (the reference has none: it runs a
 * fixed --n_iters and, on fr1desk, past the point where GBP diverges).  Runs at most max_sweeps
 * sweeps in blocks of `check_every`, evaluating the metric on the device after every sweep, and stops
 *   - GBP_STOP_CONVERGED: the mean reprojection error improved by less than rel_tol (relative) over the
 *     last `check_every` sweeps, or
 *   - GBP_STOP_DIVERGED:  the error exceeded `diverge_factor` x its running minimum (0 disables), or
 *   - GBP_STOP_MAX_SWEEPS.
 * Prior weakening is the caller's schedule (call this after the weakening phase).  stats (may be NULL)
 * receives one entry per executed sweep (capacity max_sweeps); *n_done = executed sweeps. */
#define GBP_STOP_MAX_SWEEPS 0
#define GBP_STOP_CONVERGED 1
#define GBP_STOP_DIVERGED 2
int gbp_cuda_iterate_until(gbp_handle* h, int max_sweeps, int check_every, float rel_tol, float diverge_factor,
                           gbp_iter_stats* stats, int* n_done, int* stop_reason);

/* Metrics of the current beliefs (what the host prints as "Initial
 * Reprojection error", ba/ba.cpp:992-996). */
int gbp_cuda_eval(gbp_handle* h, gbp_iter_stats* out);

/* READ_PROG (ba/ba.cpp:908-916). Any pointer may be NULL to skip it. */
int gbp_cuda_get_beliefs(gbp_handle* h, float* cam_eta, float* cam_lambda, float* lmk_eta,
                         float* lmk_lambda, float* damping, int32_t* damping_count,
                         uint32_t* robust_flag);

/* READ_PRIORS (ba/slam.cpp:913-917): slot 0 of the four message tensors. */
int gbp_cuda_get_priors(gbp_handle* h, float* cam_eta, float* cam_lambda, float* lmk_eta,
                        float* lmk_lambda);

/* NEW_KEYFRAME (ba/slam.cpp:919-928) including the trailing belief update.
 * Streams damping_count, the four prior tensors, active_flag and both weaken
 * flag tensors; `damping` itself is not re-streamed (as in the reference). */
int gbp_cuda_add_keyframe(gbp_handle* h, const int32_t* damping_count, const float* cam_prior_eta,
                          const float* cam_prior_lambda, const float* lmk_prior_eta,
                          const float* lmk_prior_lambda, const uint32_t* active_flag,
                          const uint32_t* cam_weaken_flag, const uint32_t* lmk_weaken_flag);

/* The whole keyframe insertion of slam.cpp:1020-1046 on the device, without the
 * READ_PROG / READ_PRIORS / NEW_KEYFRAME host round trip: update_flags
 * (ba/dataio.cpp:477-508: factors of camera `new_cam` become active, `new_cam` and the
 * landmarks it is the first camera to observe get weaken flag `steps`, every other
 * weaken flag 0), initialise_new_kf (ba/util.cpp:183-223: prior eta of `new_cam` =
 * its prior Lambda x the mean of keyframe new_cam-1; prior eta of the new landmarks =
 * their prior Lambda x the point 1 m in front of keyframe new_cam-1), damping_count
 * reset to -15 (slam.cpp:1039-1041) and NEW_KEYFRAME's trailing belief update.
 * new_cam = data_counter + 1 of slam.cpp.  Leaves the handle in exactly (bit for bit)
 * the state that gbp_cuda_get_beliefs + gbp_cuda_get_priors +
 * gbp_setup_next_keyframe + gbp_cuda_add_keyframe produce.  Four bytes cross the bus.
 * *n_new_lmks (may be NULL) = number of newly observed landmarks. */
int gbp_cuda_add_keyframe_device(gbp_handle* h, uint32_t new_cam, uint32_t steps, int* n_new_lmks);

/* Codelet-level entry points = Execute(cs_*) of one compute set on the
 * handle's state (used by the parity tests).
 *   relinearise_factors      cs_relinearise         ba/ba.cpp:68-97   gbp_codelets.cpp:20-172
 *   prep_messages            cs_compmess_prep       ba/ba.cpp:234,249 gbp_codelets.cpp:215-379
 *   compute_messages         Copy(mu,oldmu) + cs_computemessages + the four
 *                            Copy(messages,pmessages) that always follow it in
 *                            GBP_PROG (ba/ba.cpp:898-905); the copies are pure
 *                            double-buffer bookkeeping and commute with update_beliefs
 *   update_beliefs           prog_ub                ba/ba.cpp:104-139
 *   weaken_prior_vertices    cs_weaken_prior        ba/ba.cpp:144-185 (no belief update) */
int gbp_cuda_relinearise_factors(gbp_handle* h);
int gbp_cuda_prep_messages(gbp_handle* h);
int gbp_cuda_compute_messages(gbp_handle* h);
int gbp_cuda_update_beliefs(gbp_handle* h);
int gbp_cuda_weaken_prior_vertices(gbp_handle* h);

/* Snapshot / restore of any device tensor by its reference name, in the
 * reference's padded layout (ba/ba.cpp:665-687,759-775):
 *   cam_beliefs_eta[6C] cam_beliefs_lambda[36C] lmk_beliefs_eta[3L] lmk_beliefs_lambda[9L]
 *   cam_messages_eta[C*SK*6] cam_messages_lambda[C*SK*36] lmk_messages_eta[L*SL*3]
 *   lmk_messages_lambda[L*SL*9] and the p* previous-message twins
 *   factor_potentials_eta[9E] factor_potentials_lambda[81E]
 *   damping[E] damping_count[E] mu[9E] oldmu[9E] dmu[E] active_flag[E] robust_flag[E]
 *   measurements[2E] meas_variances[E] cam_scaling[C] lmk_scaling[L]
 *   cam_weaken_flag[C] lmk_weaken_flag[L]
 * with SK = max_nkfedges+1, SL = max_nlmkedges+1 (slot 0 = prior). */
int gbp_cuda_tensor_nbytes(gbp_handle* h, const char* name, size_t* nbytes);
int gbp_cuda_get_tensor(gbp_handle* h, const char* name, void* dst, size_t nbytes);
int gbp_cuda_set_tensor(gbp_handle* h, const char* name, const void* src, size_t nbytes);

/* Graph shape as the reference derives it (ba/ba.cpp:514-521,594-595). */
int gbp_cuda_dims(gbp_handle* h, uint32_t* n_keyframes, uint32_t* n_points, uint32_t* n_edges,
                  uint32_t* max_nkfedges, uint32_t* max_nlmkedges);

/* Device timing of the last gbp_cuda_iterate call (CUDA events on the
 * handle's stream): total milliseconds and number of kernels launched. */
int gbp_cuda_last_timing(gbp_handle* h, float* ms_total, uint64_t* kernels_launched);

/* Per-kernel device timing: when enabled, gbp_cuda_iterate brackets every kernel
 * with CUDA events on the handle's stream; gbp_cuda_last_kernel_times returns the
 * summed durations of the factor kernel (k_sweep) and of the variable kernel
 * (k_update_vars) over the last gbp_cuda_iterate call. */
int gbp_cuda_set_profile(gbp_handle* h, int enabled);
int gbp_cuda_last_kernel_times(gbp_handle* h, float* ms_factor_kernel, float* ms_variable_kernel);
/* The same per sweep of the last profiled call (arrays of `capacity` entries, either may be NULL); *n_sweeps = sweeps
 * on record.  A sweep in which the factors relinearise is a different kernel workload than one in which they do not:
 * the roofline is reported per class (bench.py). */
int gbp_cuda_last_sweep_times(gbp_handle* h, float* ms_factor_kernel, float* ms_variable_kernel, int capacity, int* n_sweeps);

/* ---- device-side unit checks of the math helpers (test support) -------- */
/* The __device__ helpers the sweep kernels inline, run on caller-supplied inputs (one thread per item): inv6x6 /
 * inv3x3 (ba/matlib.cpp:143-222; A, out: [n][36] / [n][9] row-major) and so3exp + hfunc + Jac
 * (ba/bafuncs.cpp:32-213; X [n][6] poses, P [n][3] points, K9 the 3x3 intrinsics; hx [n][2], Jkf [n][12] = 2x6,
 * Jlmk [n][6] = 2x3).  tests/test_device_helpers.py pins them to the reference's own functions. */
int gbp_cuda_test_inv6x6(const float* A, float* out, int n);
int gbp_cuda_test_inv3x3(const float* A, float* out, int n);
int gbp_cuda_test_project(const float* X, const float* P, const float* K9, float* hx, float* Jkf, float* Jlmk, int n);

/* Diagnostic builds only (-DGBP_DEBUG_TS and GBP_DEBUG_TS=1 in the environment): %globaltimer stamps of the phases of
 * the last 32 belief updates, [32][8] = {first block, last push block, first / last finish block past its wait,
 * last finish block done, last landmark block done, last camera block done, -}.  Returns the number of values
 * copied (0 in a normal build). */
int gbp_cuda_debug_timestamps(gbp_handle* h, uint64_t* out, int capacity);

/* ---- asynchronous / resident use (bench, multi-GPU) ------------------- */
/* Enqueue n sweeps on the handle's stream without synchronising. */
int gbp_cuda_iterate_async(gbp_handle* h, int n_sweeps);
int gbp_cuda_synchronize(gbp_handle* h);
/* The cudaStream_t the handle launches on (as void*), for event timing. */
void* gbp_cuda_stream(gbp_handle* h);

/* ---- multi-GPU: camera-range sharding with boundary-landmark exchange --- */
/* The reference scales by spreading variables and factors over the tiles of
 * 2^k IPUs (--ipus, ba/ba.cpp:617-631,717-753,795-834) and lets Poplar compile
 * the exchange.  Here: one process per GPU; rank r owns a contiguous camera
 * range (balanced by edge count), every factor of those cameras and a replica
 * of every landmark they observe.  Landmarks observed from more than one rank
 * are "boundary" landmarks; per sweep each rank contributes the partial sum of
 * its own factor->landmark messages (12 floats) for them -- the kernel that forms
 * the partial sums stores them straight into every rank's receive buffer over
 * NVLink (CUDA IPC peer mappings; NCCL all-gather as the fallback) -- and every
 * rank forms
 *     belief = prior + partial[rank 0] + partial[rank 1] + ...
 * in rank order, so all replicas hold identical bits. */

/* Pure host: the camera range and counts of rank `rank`. */
int gbp_cuda_plan_shard(const gbp_problem* p, uint32_t world, uint32_t rank, gbp_shard_plan* out);

/* Pure host: the rank-local sub-problem of a GLOBAL problem.  Local cameras are
 * [cam_begin, cam_end) renumbered from 0; local landmarks are the landmarks the
 * local edges touch, in ascending global id; local edges keep their global order. */
typedef struct gbp_shard gbp_shard;
int gbp_shard_build(const gbp_problem* p, uint32_t world, uint32_t rank, gbp_shard** out);
/* The same shard without copying what is one contiguous run of the global arrays (the per-edge arrays of a
 * camera range in a camera-sorted problem, the per-camera arrays): those members of gbp_shard_problem() then
 * point INTO p's arrays, which must outlive every use of them.  gbp_cuda_init_shard builds its shard this way
 * (it only reads the arrays during the call). */
int gbp_shard_build_view(const gbp_problem* p, uint32_t world, uint32_t rank, gbp_shard** out);
void gbp_shard_free(gbp_shard* s);
const gbp_problem* gbp_shard_problem(const gbp_shard* s);    /* local ids, arrays owned by s (see _view) */
const gbp_shard_plan* gbp_shard_get_plan(const gbp_shard* s);
const uint32_t* gbp_shard_lmk_global(const gbp_shard* s);    /* [n_local_points] global landmark id */
const uint32_t* gbp_shard_edge_global(const gbp_shard* s);   /* [n_local_edges]  global edge id     */
/* The boundary landmarks this rank touches: local landmark id and position in
 * the global boundary list (ascending global id, identical on every rank). */
uint32_t gbp_shard_n_boundary_local(const gbp_shard* s);
const uint32_t* gbp_shard_boundary_local(const gbp_shard* s);
const uint32_t* gbp_shard_boundary_slot(const gbp_shard* s);
/* [n_boundary_local] bit r set = rank r observes the landmark, i.e. contributes a partial sum to its belief (worlds of
 * up to 32 ranks; all zero above that, where the engine uses the NCCL all-gather).  The peer-to-peer exchange sends a
 * partial to exactly those ranks and waits for exactly those ranks' partials. */
const uint32_t* gbp_shard_boundary_ranks(const gbp_shard* s);
/* Number of active edges of the GLOBAL problem (quirk Q7: the metric covers global edges [0, n_active)). */
uint32_t gbp_shard_n_active_global(const gbp_shard* s);
/* camera-range bounds of all ranks: [world+1] */
const uint32_t* gbp_shard_cam_bounds(const gbp_shard* s);

/* 128-byte NCCL unique id (rank 0 creates it, the caller's plumbing -- e.g.
 * torch.distributed -- broadcasts it). */
int gbp_cuda_nccl_unique_id(void* id128);
/* Build the handle for this rank's shard of the GLOBAL problem `p` (its arrays are
 * only read during the call) and join the NCCL communicator.  Afterwards every entry point works on the LOCAL shard
 * (sizes via gbp_cuda_dims, index maps via gbp_cuda_shard_info); iterate /
 * weaken_priors / update_beliefs include the boundary exchange and are
 * collective: every rank must make the same sequence of calls.  add_keyframe
 * is single-GPU only. */
int gbp_cuda_init_shard(const gbp_problem* p, const gbp_opts* o, uint32_t world, uint32_t rank,
                        const void* nccl_unique_id, gbp_handle** out);
/* ---- multi-GPU in ONE process (the reference's own multi-chip mode is one host process driving 2^k IPUs,
 * --ipus, ba/ba.cpp:617-631) ------------------------------------------------------------------------------
 * Builds all `world` shard handles of the GLOBAL problem `p` in this process: out[r] is rank r's handle on
 * CUDA device devices[r] (NULL: device r modulo the device count).  The boundary exchange is the same fused
 * peer-to-peer protocol as between processes, over directly addressed peer memory
 * (cudaDeviceEnablePeerAccess); no NCCL, no IPC.  Device ordinals may repeat: shards that share a GPU are the
 * way the whole multi-rank protocol is exercised on a one-GPU box (refused when their boundary is large
 * enough for the waiting blocks to starve the peer's kernels).
 * Programs that contain an exchange must be driven through the group calls below -- they enqueue on every
 * rank before they wait on any; per-handle read-backs (get_beliefs, get_tensor, ...) work on each out[r]. */
int gbp_cuda_init_group(const gbp_problem* p, const gbp_opts* o, uint32_t world, const int* devices, gbp_handle** out);
/* GBP_PROG x n_sweeps on every rank (gbp_cuda_iterate semantics); stats (may be NULL) = the metric of the
 * WHOLE graph after every sweep, identical on all ranks. */
int gbp_cuda_group_iterate(gbp_handle** hs, uint32_t world, int n_sweeps, gbp_iter_stats* stats);
int gbp_cuda_group_weaken_priors(gbp_handle** hs, uint32_t world);
int gbp_cuda_group_eval(gbp_handle** hs, uint32_t world, gbp_iter_stats* out);
/* Synchronises every rank, then frees all handles (hs[r] = NULL afterwards). */
int gbp_cuda_group_free(gbp_handle** hs, uint32_t world);

/* Device memory of freed handles is kept in a pool private to this library (one per device) for the next
 * gbp_cuda_init of the process; this returns it to the driver. */
int gbp_cuda_release_cached_memory(void);

/* The shard a handle was built from (NULL for a single-GPU handle).  Its index maps, plan and boundary lists
 * are owned by the handle; the array members of gbp_shard_problem() that were views of the caller's problem
 * are NULL (the caller may free its arrays after init). */
const gbp_shard* gbp_cuda_shard_info(gbp_handle* h);
/* How this handle exchanges boundary partials: 0 = no exchange (single GPU or no boundary
 * landmarks), 1 = NCCL all-gather, 2 = peer-to-peer over NVLink (inside the belief-update kernel). */
int gbp_cuda_exchange_mode(gbp_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* GBP_CUDA_H_ */
