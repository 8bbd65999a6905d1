/*
 * navmodel_b200.h -- C-ABI of the B200-native hot path of elimas9/navigation_model.
 *
 * One shared library (navigation_model_b200/libnavmodel_b200.so, built by __graft_entry__.build()
 * with nvcc for sm_100a) exports exactly the entry points below.  Plain pointers and sizes only: no
 * torch / pybind11 / Eigen types.  The Python host layer binds them with ctypes
 * (navigation_model_b200/_native.py); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns int: 0 (or a non-negative count/flag) on success, < 0 on error
 *     (NM_ERR_*); the message is available from nm_last_error() (thread-local).
 *   - pointers prefixed d_ are DEVICE pointers owned by the caller (torch tensors used as buffers);
 *     h_ are HOST pointers.  The library never frees caller memory.  Handles (nm_maze, nm_streams,
 *     nm_mouse) own what they allocate and are released with nm_*_free.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are
 *     asynchronous with respect to the host unless stated; re-entrant per (device, stream).
 *   - there is NO CPU fallback: compute entry points fail with NM_ERR_CUDA when no device is usable.
 *
 * Reference interfaces replaced (paths relative to the reference repo root):
 *   nm_mouse_*        src/_navigation_model/bindings.cpp:11-62, mouse.cpp:5-92, mouse.hpp:12-45
 *                     (the reference's only native component; consumed at simulation/mouse.py:7-12)
 *   nm_maze_*         simulation/maze.py:107-122 (dense tables of the list-of-lists layout), :380-395,
 *                     :488-501
 *   nm_stream_build   simulation/rl.py:85-102 (ConditioningExperiment.run transition extraction),
 *                     :177-182,189 (PositiveConditioning candidate next states), :419-424 (replays)
 *   nm_vtable_sweep   simulation/rl.py:72-117 ConditioningExperiment.run + RLAlgorithm.update
 *                     (:176-194, :219-237, :256-262) + ShuffledReplays.offline_replays (:419-424),
 *                     batched over a parameter grid
 *   nm_loglik_sweep   the same recurrence + softmax log-probability of the observed choice built
 *                     from simulation/exploration_model.py:84-93 (extension: the reference has no
 *                     likelihood; see DESIGN.md "parity unpinned")
 *   nm_argmin_nll     new (grid-sweep epilogue; per-shard best fit fed to the NCCL all-gather)
 *   nm_simulate       simulation/experiment.py:23-51 StandardExperiment.run, with
 *                     exploration_model.py:68-109, behavioural_components.py:315-544, maze.py:563-627,
 *                     mouse.cpp:64-87 and analysis/metrics.py:192-248,425-442,697-731 fused
 */
#ifndef NAVMODEL_B200_H
#define NAVMODEL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NM_ABI_VERSION 2

/* error codes */
#define NM_OK 0
#define NM_ERR_INVALID (-1)  /* bad argument (shape, range, NULL)        -> Python ValueError   */
#define NM_ERR_RUNTIME (-2)  /* reference-style RuntimeError condition   -> Python RuntimeError */
#define NM_ERR_CUDA (-3)     /* CUDA runtime / launch failure, no device -> Python RuntimeError */
#define NM_ERR_NOMEM (-4)    /* host or device allocation failed         -> Python MemoryError  */

/* value-learning algorithms (simulation/rl.py:154-262) */
#define NM_ALG_TD0 0      /* rpe = r + g*V[s1] - V[s]                  rl.py:257-259 */
#define NM_ALG_POSITIVE 1 /* rpe = r + g*max_k V[cand_k] - V[s]        rl.py:184-187 */
#define NM_ALG_NEGATIVE 2 /* rpe = r + g*min_k V[cand_k] - V[s]        rl.py:227-230 */

#define NM_MAX_CAND 8    /* max candidate next states per step (= max number of relative actions) */
#define NM_OBS_NONE 0xFF /* nm_step_rec.obs when the arrival tile is not among the candidates */

int nm_version(void);
const char *nm_last_error(void);
/* number of usable CUDA devices (0 when there is none); never fails */
int nm_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * Mouse: pose + action endpoints (host object; replaces the pybind11 module `_navigation_model`)
 * ---------------------------------------------------------------------------------------------- */
typedef struct nm_mouse nm_mouse;

/* actions: row-major [n_actions, n_cols]; n_cols must be 2 when n_actions > 0 (mouse.cpp:5-9 raises
 * ValueError otherwise -> NULL is returned and nm_last_error() is set). */
nm_mouse *nm_mouse_create(const double *init_pos, double init_ori, const double *actions, int n_actions,
                          int n_cols, int save_history);
void nm_mouse_free(nm_mouse *m);
/* out[6] = {pos.x, pos.y, ori, init_pos.x, init_pos.y, init_ori} */
int nm_mouse_state(const nm_mouse *m, double *out6);
int nm_mouse_num_actions(const nm_mouse *m);
/* out: row-major [n_actions, 2]; R(ori) * a_i + pos  (mouse.cpp:64-73) */
int nm_mouse_endpoints(const nm_mouse *m, double *out);
/* returns 1 if the mouse moved, 0 if new_pos isApprox(pos)  (mouse.cpp:75-87) */
int nm_mouse_move_to(nm_mouse *m, double x, double y);
/* n move_to calls in a row, points: [n, 2]; returns what the last one returned.  The replay passes of
 * PositiveConditioning / NegativeConditioning move the learner's mouse to every replayed arrival point
 * (rl.py:189, 232): the device launch computes the tables, this call leaves the mouse where the reference does. */
int nm_mouse_move_through(nm_mouse *m, const double *points, int64_t n);
int nm_mouse_enable_history(nm_mouse *m, int enable);
int nm_mouse_reset_history(nm_mouse *m);
/* init_pos / init_ori may be NULL (= keep)  (mouse.cpp:43-53) */
int nm_mouse_reset(nm_mouse *m, const double *init_pos, const double *init_ori);
int64_t nm_mouse_history_len(const nm_mouse *m);
/* pos_out: [len, 2], ori_out: [len] */
int nm_mouse_history(const nm_mouse *m, double *pos_out, double *ori_out);
/* Adopt a trajectory computed elsewhere (the simulation kernel): the n poses are appended to the history
 * (when enabled) exactly as n move_to calls would have, and the last one becomes the current pose. */
int nm_mouse_adopt_trajectory(nm_mouse *m, const double *pos, const double *ori, int64_t n);

/* ------------------------------------------------------------------------------------------------
 * Maze: dense tile tables resident on the device
 * ---------------------------------------------------------------------------------------------- */
typedef struct nm_maze nm_maze;

/* h_tiles: u8[X*Y], x-major (index = x*Y + y, x = text row; maze.py:126-143, 266-276); a tile is
 * visitable iff its code != 0 (VOID; maze.py:56-57, 99-100).  h_subareas: u8[X*Y] or NULL.
 * device < 0 builds a host-only handle (usable by nm_stream_build and the nm_maze_* queries). */
nm_maze *nm_maze_upload(const uint8_t *h_tiles, const uint8_t *h_subareas, int X, int Y, double size_tile,
                        int device, void *stream);
void nm_maze_free(nm_maze *mz);
/* number of visitable tiles = number of compact states */
int nm_maze_num_states(const nm_maze *mz);
/* int(x // size_tile), int(y // size_tile) with Python float floor-division (maze.py:393-394) */
int nm_maze_cont2disc(const nm_maze *mz, double x, double y, int *dx, int *dy);
/* ray / AABB segment test (maze.py:563-611): 1 possible, 0 not */
int nm_maze_movement_possible_cont(const nm_maze *mz, double x1, double y1, double x2, double y2);
/* closest visitable tile (maze.py:469-484) */
int nm_maze_closest_visitable_cont(const nm_maze *mz, double x, double y, int *dx, int *dy);

/* ------------------------------------------------------------------------------------------------
 * Step streams: the parameter-independent part of a conditioning session
 * ---------------------------------------------------------------------------------------------- */
/* One transition (rl.py:12-24 `Transition`, flattened).  32 bytes, 16-byte aligned.  State ids are
 * COMPACT ids of visitable tiles in x-major order (0 .. num_states-1). */
typedef struct nm_step_rec {
  double r;                    /* reward[i]                                      rl.py:95       */
  uint16_t s;                  /* start state                                    rl.py:87,102   */
  uint16_t s1;                 /* arrival state                                  rl.py:94       */
  uint8_t ncand;               /* K = len(next_states), 0..NM_MAX_CAND           rl.py:182      */
  uint8_t obs;                 /* first k with cand[k] == s1, else NM_OBS_NONE   (extension)    */
  uint16_t reserved;
  uint16_t cand[NM_MAX_CAND];  /* next_states in action order (unused slots = 0) rl.py:177-182  */
} nm_step_rec;

/* Build the T-1 online records of one session on the HOST (native C++, no GPU needed).
 *   trajectory [T,2], reward [T] (reward[0] unused, rl.py:95).
 *   mouse: the algorithm's own Mouse (rl.py:173); it is driven exactly like the reference drives it
 *   (move_to(start), endpoints, move_to(arrival) per transition, rl.py:178-189 -- history included) and
 *   is left at the final pose.  NULL skips the endpoint geometry (TD0 without likelihood): ncand = 0,
 *   obs = NM_OBS_NONE.
 * Returns the number of records written (T-1), or < 0.  A step with zero candidates is an error when
 * a mouse is given (the reference raises from np.max([]), rl.py:184). */
int64_t nm_stream_build(const nm_maze *mz, const double *trajectory, const double *reward, int64_t T,
                        nm_mouse *mouse, nm_step_rec *out);

/* Records of n INDEPENDENT transitions starts[i] -> arrivals[i] ([n,2] continuous points) between the given tiles
 * ([n,2] int32 (x, y), both visitable): what RLAlgorithm.update is handed by ShuffledReplays' imaginary replays
 * (simulation/rl.py:388-418), one fresh Transition per update.  The algorithm's mouse is driven per transition as in
 * rl.py:178-189 (move_to(start), endpoints -> candidates, move_to(arrival)); NULL: no candidates (TD0).
 * Returns n or < 0. */
int64_t nm_stream_build_pairs(const nm_maze *mz, const double *starts, const double *arrivals,
                              const int32_t *start_tiles, const int32_t *arrival_tiles, const double *reward, int64_t n,
                              nm_mouse *mouse, nm_step_rec *out);

typedef struct nm_streams nm_streams;

/* Upload S concatenated session streams.  Session j owns records [h_offsets[j], h_offsets[j+1]);
 * its first h_online[j] records are the observed (online) transitions, the remaining ones are the
 * shuffled-replay passes in execution order (rl.py:419-424) and update V only.
 * Copies are enqueued on `stream` from the (ideally pinned) host buffers, which must stay valid until
 * the stream has been synchronised. */
nm_streams *nm_streams_upload(const nm_step_rec *h_records, const int64_t *h_offsets, const int64_t *h_online,
                              int S, int num_states, int device, void *stream);
/* Re-upload the records of streams with the same layout (same offsets / online counts) into the existing
 * device allocation: one H2D copy enqueued on `stream`, no allocation. */
int nm_streams_update(nm_streams *st, const nm_step_rec *h_records, void *stream);
void nm_streams_free(nm_streams *st);

/* ------------------------------------------------------------------------------------------------
 * Session ingestion on the device (many sessions at once)
 * ---------------------------------------------------------------------------------------------- */
/* Session.__init__(new_sampling_time=...) subsampling (analysis/data.py:39-74) of S raw recordings stored back to
 * back: session j owns samples [d_offsets[j], d_offsets[j+1]) of d_time [n], d_traj [n,2] and (optional) d_reward [n].
 * For every window of at least `sampling_time` seconds the first finite point and the first reward that is neither 0
 * nor NaN (else 0.0) are kept; windows without a finite point are skipped; NaN timestamps are skipped.  Session j's
 * d_counts[j] subsampled points are written from index d_offsets[j] of d_traj_out / d_reward_out (same layout as
 * the input: no compaction).  All pointers are DEVICE pointers. */
int nm_sessions_subsample_device(const double *d_time, const double *d_traj, const double *d_reward,
                                 const int64_t *d_offsets, int S, double sampling_time, double *d_traj_out,
                                 double *d_reward_out, int64_t *d_counts, void *stream);

/* compute_orientations (analysis/metrics.py:653-694) of S sessions laid out like above (session j: d_counts[j] points
 * from index d_offsets[j]).  d_ori_out [.] gets one heading per point at the same indices; d_first_moved [S] the index
 * of the first step that moved, -1 when a session never moves (its headings are NaN; the reference raises IndexError).
 * n_actions > 0 (with a device maze): a step only counts as a move when its end tile lies under one of the endpoints
 * of the actions rotated by the previous heading (":668-671").  All d_ pointers are DEVICE pointers. */
int nm_sessions_orientations_device(const nm_maze *mz, const double *d_traj, const int64_t *d_offsets,
                                    const int64_t *d_counts, int S, const double *h_actions, int n_actions,
                                    double *d_ori_out, int64_t *d_first_moved, void *stream);

/* adjust_positions_to_maze (analysis/data.py:372-388) on n DEVICE points [n,2]: a point outside the visitable tiles
 * moves to the centre of the nearest visitable tile; points with a NaN coordinate stay.  d_out may alias d_points. */
int nm_positions_adjust_device(const nm_maze *mz, const double *d_points, int64_t n, double *d_out, void *stream);

/* nm_stream_build for S sessions at once, on the device: session j's h_counts[j] >= 1 points start at index
 * h_offsets[j] of d_traj [.,2] / d_reward [.] (e.g. the output of nm_sessions_subsample_device) and yield
 * h_counts[j] - 1 online records.  n_actions = 0: no candidates (TD0 value learning); otherwise every session gets its
 * own algorithm mouse, Mouse(trajectory[0], 0.0, actions) or h_mouse_init[j] = (x, y, ori).  The returned handle is
 * resident on the maze's device and is used like one from nm_streams_upload.  Synchronises `stream`.  NULL on error
 * (a step without visitable candidate is an error, rl.py:184). */
nm_streams *nm_streams_build_device(const nm_maze *mz, const double *d_traj, const double *d_reward,
                                    const int64_t *h_offsets, const int64_t *h_counts, int S, const double *h_actions,
                                    int n_actions, const double *h_mouse_init, void *stream);

/* New streams = the online records of every session of `src` (built on the device) followed by replay passes:
 * h_order holds, session after session, h_counts[j] indices into session j's online records -- the order in which
 * ShuffledReplays.offline_replays presents them again (n_times cumulative in-place Generator.shuffle passes,
 * rl.py:419-424; the permutation is parameter-independent, so it is drawn once on the host).  The records are gathered
 * on the device; `src` is left untouched.  NULL on error. */
nm_streams *nm_streams_with_replays(const nm_streams *src, const int64_t *h_order, const int64_t *h_counts,
                                    void *stream);
/* Copy the records of a stream handle back to the host ([total] nm_step_rec); synchronises `stream`. */
int nm_streams_download(const nm_streams *st, nm_step_rec *h_records, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Parameter-grid sweeps (one thread = one parameter set; V-table resident in shared memory)
 * ---------------------------------------------------------------------------------------------- */
/* V-table learning over P parameter sets x S sessions (each session starts from v_init).
 *   d_alpha, d_gamma: [P]
 *   d_v_init: NULL (zeros, rl.py:66) | [X*Y] shared (v_init_per_unit = 0) | [P, X*Y] (= 1)  rl.py:63-64
 *   d_v_out:  [S, P, X*Y] dense tables (non-visitable tiles keep their v_init value)
 */
int nm_vtable_sweep(int alg, const nm_maze *mz, const nm_streams *st, const double *d_alpha,
                    const double *d_gamma, int64_t P, const double *d_v_init, int v_init_per_unit,
                    double *d_v_out, void *stream);

/* ONE parameter set over ONE session of the streams, with what RLAlgorithm.update RETURNS per step: the reference puts
 * Update(transition, dv, |rpe|) of every online step into the replay buffer (rl.py:97-98, 190-194, 233-237, 260-262).
 * Same recurrence and operations as nm_vtable_sweep (bit-identical table).  This is the launch behind
 * ConditioningExperiment.run / do_replays for the built-in learners.
 *   d_v_init: NULL (zeros) | [X*Y];  d_v_out: [X*Y] dense;  d_dv_out, d_rpe_out: [records of the session] or NULL
 *   (records past the session's online count are replay passes: their entries are filled too). */
int nm_vtable_trace(int alg, const nm_maze *mz, const nm_streams *st, int session, double alpha, double gamma,
                    const double *d_v_init, double *d_v_out, double *d_dv_out, double *d_rpe_out, void *stream);
/* largest number of visitable tiles nm_vtable_trace holds (its table lives in shared memory) */
int nm_vtable_trace_max_states(void);

/* Largest number of visitable tiles (compact states) the per-parameter-set value-table kernels can hold in SHARED
 * memory (863 with 227 KB: 32 table columns of 8 bytes per state next to the stream rings).  Larger mazes run the same
 * sweeps with their tables in device memory (T[state][parameter set], coalesced, L2 / HBM resident; also on request
 * with NM_FIT_TABLES=global): same results, no size limit. */
int nm_vtable_max_states(void);

/* The sweeps that keep per-set tables in device memory (nm_loglik_sweep / nm_vtable_sweep mode 2, nm_qtable_sweep mode
 * 2) hold ONE grow-only scratch block per (device, stream) (states x parameter sets x 8 bytes of the largest sweep so
 * far on that stream) for the life of the process.  This call waits for the devices and gives the blocks back; later sweeps allocate again. */
int nm_scratch_release(void);

/* Same recurrence + log-likelihood of the observed choices under the softmax policy
 * p_k = softmax_k(beta * V[cand_k]) (exploration_model.py:84-93), evaluated BEFORE the update of the
 * step, accumulated in FP64 in step order over the online records whose obs != NM_OBS_NONE.
 *   d_ll_out: [S, P];  d_v_out: [S, P, X*Y] or NULL.
 */
int nm_loglik_sweep(int alg, const nm_maze *mz, const nm_streams *st, const double *d_alpha,
                    const double *d_beta, const double *d_gamma, int64_t P, const double *d_v_init,
                    int v_init_per_unit, double *d_ll_out, double *d_v_out, void *stream);

/* Which kernel the last nm_loglik_sweep on `stream` (current device) used:
 *   1 = one value table per WARP in shared memory, shared by an aligned run of 32 consecutive parameter sets with equal
 *       (alpha, gamma) -- the device checks the arrays itself, so a Cartesian grid whose beta axis varies fastest (a
 *       multiple of 32 points) takes this path without any hint (NM_FIT_GROUP=0 disables it);
 *   2 = one table per parameter set in DEVICE memory (T[state][set], coalesced, L2 / HBM resident): arbitrary parameter
 *       sets, any maze size (the default for them; NM_FIT_TABLES=shared selects 0 instead);
 *   0 = one table per parameter set in shared memory (mazes of at most nm_vtable_max_states() states).
 * Results are bit-identical in all three.  Synchronises the stream: diagnostics and tests only. */
int nm_loglik_sweep_mode(void *stream, int *mode_out);

/* Per-shard best fit: nll[p] = -(sum_s ll[s, p]) (sessions added in index order), then the minimum and
 * the LOWEST index attaining it (numpy argmin).  NaN entries are skipped; if all are NaN idx = -1.
 *   d_nll_out: [P] or NULL;  d_best_out: {double best_nll; int64 best_idx} (16 bytes, device).
 */
int nm_argmin_nll(const double *d_ll, int S, int64_t P, double *d_nll_out, void *d_best_out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Forward simulation of many mice (one thread = one mouse)
 * ---------------------------------------------------------------------------------------------- */
/* behavioural components, in the order the BehaviouralModel sums them (exploration_model.py:84-86) */
#define NM_COMP_ANXIETY 1      /* behavioural_components.py:315-376 (StandardTiles and GridTiles variants) */
#define NM_COMP_BCOST 2        /* :379-416  BiomechanicalCostComponent                                  */
#define NM_COMP_EXPERIENCE 3   /* :419-482  ExperienceComponent                                         */
#define NM_COMP_BPERSISTENCE 4 /* :485-516  BiomechPersistenceComponent                                 */
#define NM_COMP_VTABLE 5       /* :519-544  ValueTableComponent                                         */
#define NM_MAX_COMPONENTS 8

/* All d_ arrays are device pointers; [N] = one entry per mouse.  Unused components: NULL. */
#define NM_TILE_ZONE_IN 1u    /* tile belongs to the shock zone                                       */
#define NM_TILE_ZONE_OUT 2u   /* standing here ends a visit of the shock zone                         */
#define NM_TILE_CORRIDOR_A 4u /* arms(): subarea 1                                                    */
#define NM_TILE_CORRIDOR_B 8u /* arms(): subarea 7                                                    */
#define NM_MAX_SUBAREAS 16
#define NM_MAX_ORI_BINS 16

typedef struct nm_sim_args {
  int64_t n_mice;
  int32_t n_steps;                         /* iterations of StandardExperiment.run (experiment.py:33)      */
  int32_t n_actions;                       /* A <= NM_MAX_CAND                                             */
  const double *h_actions;                 /* HOST [A,2] relative displacements of the Mouse               */
  const uint8_t *h_zero_action;            /* HOST [A] 1 where the component's discrete action is (0,0)    */
  int32_t n_components;                    /* <= NM_MAX_COMPONENTS                                         */
  int32_t component_kind[NM_MAX_COMPONENTS];   /* NM_COMP_*, in model order                                */
  int32_t component_active[NM_MAX_COMPONENTS]; /* Component.active (behavioural_components.py:237,283-285) */
  const double *d_beta;                    /* [N]   BehaviouralModel beta (exploration_model.py:50)        */
  const double *d_init_pose;               /* [N,3] Mouse(init_pos, init_ori)                              */
  const double *d_init_prev;               /* [N,2] init_pos_discrete as doubles (exploration_model.py:47) */
  const double *d_anxiety;                 /* [N,5] W_anx, p1, p2, p3, p4 (p4: GridTiles decision points)  */
  const double *d_bcost_weight;            /* [N]   W_bcost                                                */
  const double *d_bcost_table;             /* [N,A] vonmises(k).pdf(angle) * gamma-mask, host-precomputed
                                              (behavioural_components.py:395-406)                          */
  const double *d_experience;              /* [N,6] W_exp, xi, kappa_home, kappa_explo, theta_explo,
                                              theta_home                                                   */
  const double *d_bpersistence;            /* [N,4] W_bper, bp, Wm, Wnm                                    */
  const double *d_vtable_weight;           /* [N]   W_values                                               */
  const double *d_vtable;                  /* normalised table(s) (behavioural_components.py:530-535):
                                              [X*Y] shared or [N,X*Y]                                      */
  int32_t vtable_per_mouse;
  int32_t model_kind;                      /* 0 BehaviouralModel (softmax over components);
                                              1 RandomModel (uniform rng.choice, exploration_model.py:131-144;
                                              components ignored)                                          */
  const double *d_draws;                   /* [T,N] replayed uniforms of Generator.choice, or NULL         */
  const uint64_t *d_pcg_state;             /* [N,4] PCG64 {state_hi, state_lo, inc_hi, inc_lo} or NULL     */
  double visit_percentage;                 /* it_perc_visited_maze(percentage), metrics.py:698             */
  /* outputs (NULL = not wanted) */
  double *d_final_pose;                    /* [N,3]                                                        */
  uint64_t *d_pcg_state_out;               /* [N,4] generator state after the run                          */
  double *d_hist_pos;                      /* [T+1,N,2] mouse.history_position (time-major)                */
  double *d_hist_ori;                      /* [T+1,N]   mouse.history_orientation                          */
  uint8_t *d_hist_choice;                  /* [T,N]     index (into the A actions) of the chosen endpoint  */
  uint32_t *d_occupancy;                   /* [N,X*Y]   occupancy_map over the T+1 history positions       */
  int32_t *d_it_visit;                     /* [N]       it_perc_visited_maze                               */
  int32_t *d_count_moving;                 /* [N]       motion_analysis count_moving                       */
  int32_t *d_tile_counts;                  /* [N,8]     tile_analysis counts by tile code                  */
  int32_t *d_status;                       /* [N]       0 ok; t+1 if step t had no reachable endpoint
                                              (the reference raises from np.max([]))                        */
  unsigned long long *d_occupancy_total;   /* [X*Y]     sum of the occupancy maps over all mice (atomics)  */
  double *d_model_state;                   /* [N,5]     BehaviouralModel._prev_pos (x, y), _moving (0/1),
                                              ExperienceComponent._familiarity, _ment_state (0 EXPLO/1 HOME) */
  const uint32_t *d_pcg_u32;               /* [N,2] numpy's buffered 32-bit half {has_uint32, uinteger} or NULL */
  uint32_t *d_pcg_u32_out;                 /* [N,2] ... after the run                                       */
  uint64_t device_seed;                    /* use_device_seed != 0: every mouse seeds its own PCG64 stream ON THE
                                              DEVICE from (device_seed, mouse index) (splitmix64 -> pcg64_srandom);
                                              fast path for millions of mice, NOT numpy's default_rng seeding   */
  int32_t use_device_seed;
  int32_t reserved1;
  /* ---- more fused trajectory metrics (analysis/metrics.py; all optional, NULL / 0 = not wanted) ---- */
  const uint8_t *d_tile_flags;             /* [X*Y] per tile: NM_TILE_ZONE_IN | NM_TILE_ZONE_OUT | NM_TILE_CORRIDOR_A |
                                              NM_TILE_CORRIDOR_B (host-built from subareas / tile lines)       */
  int32_t *d_zone_metrics;                 /* [N,4] latency_enter_shock_zone (metrics.py:488-516; n_steps if never),
                                              positions inside the zone (occupancy_shock_zone :519-534),
                                              shock_zone_entries (:537-586), corridor switches (arms :735-769) */
  int32_t *d_static_intervals;             /* [N] not-moving runs of moving_bool_list (hist_time_moving :445-467
                                              -> compute_static_intervals :621-635)                            */
  int32_t *d_subarea_counts;               /* [N,NM_MAX_SUBAREAS] positions by subarea index (subareas :470-485);
                                              the maze must have been uploaded with subareas                   */
  int32_t n_ori_bins;                      /* > 0: histogram of the turns ori[t+1]-ori[t] over the steps that
                                              change tile (hist_orientations :306-359), <= NM_MAX_ORI_BINS      */
  int32_t ori_distance;                    /* hist_orientations(distance=True): turns in bin 2 that jump more
                                              than 1.5 tiles go to the extra last entry                        */
  uint32_t ori_tile_filter;                /* != 0: hist_orientations_filtered (:383-422), bit c = tile code c of
                                              the step's START tile is accepted                                */
  int32_t reserved2;
  const double *h_ori_edges;               /* HOST [n_ori_bins+1] the edges numpy.histogram uses               */
  double ori_min_lim, ori_max_lim;         /* turns outside are shifted by one full turn (:331-334)            */
  int32_t *d_ori_hist;                     /* [N, n_ori_bins+1]; last entry = the distance=True extra count    */
  /* ---- continuing an earlier run: StandardExperiment.run called again on the same model (experiment.py:33-51
   *      keeps no per-run state: the components and the model carry on where they were); NULL = fresh objects ---- */
  const uint32_t *d_init_visits;           /* [N,X*Y] ExperienceComponent._maze_map_exp before the run
                                              (behavioural_components.py:437-441).  The visit map of the kernel
                                              starts from it, so every output derived from that map (d_occupancy,
                                              d_occupancy_total, d_tile_counts, d_it_visit, zone / subarea
                                              occupancy) counts the seeded visits too                          */
  const double *d_init_model_state;        /* [N,3] BehaviouralModel._moving (0/1), ExperienceComponent._familiarity,
                                              _ment_state (0 EXPLO / 1 HOME) before the run                     */
  uint32_t init_visits_max;                /* upper bound of d_init_visits (picks the counter width)           */
  int32_t reserved3;
  /* ---- a shard of a larger population (mice split over the GPUs of a box in contiguous blocks) ---- */
  int64_t mouse_index_base;                /* use_device_seed: mouse n seeds its stream from (device_seed,
                                              mouse_index_base + n), so a population gives the same mice whatever the
                                              number of shards                                                  */
  /* ---- mazes whose visit map does not fit shared memory (nm_simulate_fits == 2) ---- */
  uint32_t *d_visit_scratch;               /* [X*Y, visit_scratch_columns] 32-bit words of device memory for the visit
                                              maps (count[tile][launched thread]); contents on entry are ignored.
                                              NULL when the shared-memory map is used                             */
  int64_t visit_scratch_columns;           /* >= nm_simulate_scratch_columns(n_mice)                             */
} nm_sim_args;

/* StandardExperiment.run(n_steps) for n_mice independent mice on the maze's device. */
int nm_simulate(const nm_maze *mz, const nm_sim_args *args, void *stream);

/* sizeof(nm_sim_args) of the built library: a binding checks its own mirror of the struct against it at load time. */
int nm_sim_args_size(void);

/* Where the visit maps of an X x Y maze live for runs of n_steps: 1 = shared memory (the counters are 16 bits wide
 * below 65535 steps; mazes up to about 50 x 50 tiles); 2 = device memory (the caller passes nm_sim_args.d_visit_scratch;
 * same step body and results; up to about 66 000 tiles, whose tile / grid-line tables must still fit shared memory);
 * 0 = the maze is too large for nm_simulate. */
int nm_simulate_fits(int X, int Y, int n_steps);
/* columns of nm_sim_args.d_visit_scratch nm_simulate needs for n_mice (the launched threads: whole CTAs) */
int64_t nm_simulate_scratch_columns(int64_t n_mice);

/* Vectorised Maze.is_movement_possible_cont on the device: d_segments [n,4] = (x1,y1,x2,y2),
 * d_out u8[n] (1 possible).  maze.py:563-611 */
int nm_maze_movements_possible_device(const nm_maze *mz, const double *d_segments, int64_t n, uint8_t *d_out,
                                      void *stream);

/* ------------------------------------------------------------------------------------------------
 * Model-based agent (EXTENSION, parity unpinned: the reference has no model-based agent)
 * ---------------------------------------------------------------------------------------------- */
/* One warp = one parameter set.  Learned reward model Rhat[s1] += alpha * (r - Rhat[s1]), n_sweeps synchronous
 * value-iteration sweeps V'[u] = max_a (Rhat[next(u,a)] + gamma * V[next(u,a)]) after every observed step, and
 * the softmax log-likelihood (exploration_model.py:84-93) of the observed tile step under
 * Q(s,a) = Rhat[next(s,a)] + gamma * V[next(s,a)], evaluated before the step's updates.
 *   d_next: int16[num_states, n_actions] compact successor states, -1 where the centre-to-centre movement is
 *           not possible (Maze.is_movement_possible, maze.py:512-527); the streams only need s, s1, r.
 *   d_v_init: [X*Y] shared or NULL;  d_ll_out: [S, P];  d_v_out, d_r_out: [S, P, X*Y] or NULL. */
int nm_modelbased_sweep(const nm_maze *mz, const nm_streams *st, const int16_t *d_next, int n_actions, int n_sweeps,
                        const double *d_alpha, const double *d_beta, const double *d_gamma, int64_t P,
                        const double *d_v_init, double *d_ll_out, double *d_v_out, double *d_r_out, void *stream);

/* Which mapping the last nm_modelbased_sweep on `stream` (current device) used: 0 = a warp per parameter set,
 * 1 = a warp per aligned run of 32 consecutive parameter sets with equal (alpha, gamma) -- V, Rhat and the
 * value-iteration sweeps are then computed once per run (checked on the device, bit-identical results;
 * NM_MB_GROUP=0 forces 0).  Synchronises the stream: diagnostics and tests only. */
int nm_modelbased_sweep_mode(void *stream, int *mode_out);

/* ------------------------------------------------------------------------------------------------
 * Tabular Q-learning agent with an explicit Q[s,a] table (EXTENSION, parity unpinned: the reference's learners keep
 * state values only, simulation/rl.py:154-262)
 * ---------------------------------------------------------------------------------------------- */
/* Eight lanes = one parameter set, lane k = action k.  Per step (s -> s1, r) whose observed action
 * a = first a with next[s,a] == s1 exists (other steps are skipped): the softmax log-likelihood
 * (exploration_model.py:84-93) of a under beta * Q[s, available actions], evaluated before the update, then
 * Q[s,a] += alpha * (r + gamma * max_{a' available at s1} Q[s1,a'] - Q[s,a])  (update form of rl.py:184-187;
 * bootstrap 0 when s1 has no available action).  Replay records update Q only.
 *   d_next: int16[num_states, n_actions] as for nm_modelbased_sweep; the streams only need s, s1, r.
 *   d_q_init: [X*Y, n_actions] shared or NULL (zeros);  d_ll_out: [S, P];  d_q_out: [S, P, X*Y, n_actions] or NULL. */
int nm_qtable_sweep(const nm_maze *mz, const nm_streams *st, const int16_t *d_next, int n_actions,
                    const double *d_alpha, const double *d_beta, const double *d_gamma, int64_t P,
                    const double *d_q_init, double *d_ll_out, double *d_q_out, void *stream);

/* Which kernel the last nm_qtable_sweep on `stream` (current device) used: 1 = one table per warp in shared memory,
 * shared by an aligned run of 32 consecutive parameter sets with equal (alpha, gamma) (checked on the device;
 * NM_Q_GROUP=0 disables it); 2 = one table per parameter set in device memory (T[state * 8 + action][set], coalesced;
 * arbitrary sets, any maze size -- the default for them); 0 = one table per parameter set in shared memory, eight lanes
 * each (NM_Q_TABLES=shared).  Bit-identical results.  Synchronises the stream: diagnostics and tests only. */
int nm_qtable_sweep_mode(void *stream, int *mode_out);

/* ------------------------------------------------------------------------------------------------
 * Batched distances between per-mouse metric results and data metrics (analysis/statistics.py:10-103)
 * ---------------------------------------------------------------------------------------------- */
#define NM_DIST_NORMALIZED 0 /* normalized_distance  statistics.py:67-92  */
#define NM_DIST_MANHATTAN 1  /* manhattan_distance   statistics.py:95-103 */
#define NM_DIST_JS 2         /* js_distance on the count vectors, statistics.py:10-54 */
/* d_x: [N, L] one result vector per mouse; d_ref: [L] the data metric; d_out: [N]. */
int nm_distance_rows(int kind, const double *d_x, const double *d_ref, int64_t N, int L, double *d_out, void *stream);

/* FP64 issue-rate micro-benchmark used for the roofline denominator: every thread runs `iters`
 * iterations of 8 independent DFMA chains; returns the instruction count via *h_fp64_instr.
 * Launches on `stream`; the caller times it with CUDA events. */
int nm_fp64_peak_probe(int blocks, int threads, int iters, double *d_sink, int64_t *h_fp64_instr, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NAVMODEL_B200_H */
