// Forward simulation of many independent mice: StandardExperiment.run on the device.
//
// Mapping (DESIGN.md "simulation kernel"):
//   * one THREAD owns one mouse: pose, softmax policy state, experience map and generator live in
//     registers / shared memory for the whole run; nothing is written to HBM per step unless a
//     history is requested;
//   * the visit-count map is kept in shared memory as count[tile][thread] (u16 while n_steps < 65535, else u32; in device
//     memory for mazes beyond ~50 x 50 tiles): one map serves both the ExperienceComponent (visits so far,
//     behavioural_components.py:456-479) and the fused occupancy metric (analysis/metrics.py:204-216);
//   * the per-mouse booleans travel in one register word, the reachability of the eight endpoints is cached per mouse and
//     the ray walks of the mice that moved are shared by the 32 lanes of the warp; the reference's standard
//     five-component model has its own instantiation with straight-line component bodies (kStdWord);
//   * the maze tile table is staged once per CTA into shared memory; the ray / AABB walk reads it there;
//   * random draws are either replayed (`draws[T, N]`, coalesced) or generated on the device with a
//     numpy-identical PCG64 (XSL-RR 128/64) from host-computed (state, inc).
//
// Reference loop restated here, operation for operation (no FMA contraction, -fmad=false):
//   StandardExperiment.run                     simulation/experiment.py:33-51
//   Mouse.get_endpoints / move_to              src/_navigation_model/mouse.cpp:64-87
//   Maze.are_movements_possible_cont           simulation/maze.py:563-627
//   BehaviouralModel.decision_making           simulation/exploration_model.py:68-109
//   components                                 simulation/behavioural_components.py:315-544
//   Generator.choice(a, p=p)                   numpy: cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, 'right')
//   fused metrics                              analysis/metrics.py:192-248, 425-442, 697-731
#include <cuda_runtime.h>

#include <cstdlib>
#include <type_traits>

#include "nm_common.h"
#include "nm_libm.h"
#include "nm_pose.h"
#include "nm_simgeom.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
#include "nm_exp.cuh"

// A mouse's biomechanical-cost row comes from L2 every step (64 bytes per mouse: far more per SM than the 28 KB of L1 left
// beside the shared memory), and read right before its use, behind the `moving` test, that latency was exposed.  The row
// is therefore read through a plain (non-volatile, side-effect-free) asm load: the compiler may -- and does -- issue the
// eight loads early in the step, above the branches and the shared-memory asm statements that an ordinary load does not
// cross (measured: 96.1 -> 94.2 ms per 1M x 1000; an unconditional __ldg in C++ is NOT moved and costs 97.4 ms).  Because
// the load can run for every mouse of every model, the pointer must always be valid: nm_zero_row stands in when the model
// has no biomechanical-cost component.
__device__ const double nm_zero_row[NM_MAX_CAND] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
__device__ __forceinline__ double nm_ldg_early(const double *p) {
  double v;
  asm("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

constexpr int kA = NM_MAX_CAND;
constexpr int kSimThreads = 128;
#ifndef NM_SIM_BCW_GLOBAL
#define NM_SIM_BCW_GLOBAL 1  // the biomechanical-cost tables stay in device memory: 64 B of shared memory less per mouse
                             // buy a 7th resident CTA per SM (measured: 144.7 ms against 154.3 ms per 1M x 1000 steps)
#endif
constexpr int kBcwRows = NM_SIM_BCW_GLOBAL ? 0 : kA;
constexpr size_t kSimFixedBytes = ExpSim::kRegionBytes + 16 * kA;  // exponential table region + relative actions
#ifndef NM_SIM_MIN_CTAS
#define NM_SIM_MIN_CTAS 4  // resident CTAs per SM the register allocation is capped for (32-bit visit counters)
#endif
#ifndef NM_SIM_MIN_CTAS_U16
#define NM_SIM_MIN_CTAS_U16 7  // same, for the variant with 16-bit visit counters (72 registers; more warps beat fewer spills)
#endif

struct SimDev {
  nm_sim_args a;  // device-visible copy (host pointers inside are not dereferenced on the device)
  const uint8_t *tiles;
  int X, Y, n_visitable;
  double st;
  double act[kA][2];
  uint8_t zero_act[kA];
  uint32_t zero_mask;    // bit i = zero_act[i]
  uint32_t comp_word;    // 4 bits per ACTIVE component, in model order: its NM_COMP_* kind (0 terminates the list)
  uint8_t null_act[kA];  // exactly (0, 0): the endpoint is the position itself
  uint32_t null_mask;    // bit i = null_act[i]
  uint8_t walk_ord[kA];  // ordinal of action i among the actions that need a ray walk
  uint32_t walk_list;    // 4 bits per entry: walk_list >> 4j & 7 = j-th action that needs a ray walk
  int n_walk;            // number of those
  const uint8_t *subareas;                 // [X*Y] or NULL
  double ori_edges[NM_MAX_ORI_BINS + 1];   // numpy.histogram edges of the turn histogram
  int64_t scratch_columns;                 // GMAP: columns (= launched threads) of the device-memory visit map
  int n_visit_thr;                         // smallest number of visited tiles that reaches a.visit_percentage (metrics.py:711-721)
};

constexpr uint32_t kStdWord = NM_COMP_ANXIETY | (NM_COMP_BCOST << 4) | (NM_COMP_EXPERIENCE << 8) |
                              (NM_COMP_BPERSISTENCE << 12) | (NM_COMP_VTABLE << 16);

// cos / sin / atan2 of the pose chain: nm_host_sincos / nm_host_atan2 (nm_libm.h) give the HOST library's results, so
// that a simulated trajectory equals the reference's to the last bit (mouse.cpp:66, 78 call std::cos / sin / atan2).
__device__ __forceinline__ uint64_t rotr64(uint64_t v, unsigned r) { return (v >> r) | (v << ((64 - r) & 63)); }

// numpy PCG64: 128-bit LCG step, then XSL-RR output of the NEW state; random() = (u64 >> 11) * 2^-53
struct Pcg64 {
  unsigned __int128 state, inc;
  uint32_t has32, buf32;  // numpy's buffered second half of a 64-bit output (bitgen_t has_uint32 / uinteger)
  __device__ __forceinline__ uint64_t next64() {
    const unsigned __int128 mult = ((unsigned __int128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
    state = state * mult + inc;
    const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
    return rotr64(hi ^ lo, (unsigned)(hi >> 58));
  }
  __device__ __forceinline__ double next_double() {
    return (double)(next64() >> 11) * (1.0 / 9007199254740992.0);
  }
  __device__ __forceinline__ uint32_t next32() {  // low half first, the high half is buffered
    if (has32) {
      has32 = 0u;
      return buf32;
    }
    const uint64_t v = next64();
    has32 = 1u;
    buf32 = (uint32_t)(v >> 32);
    return (uint32_t)v;
  }
  // Generator.integers(0, n) for n <= 2^32: buffered 32-bit Lemire rejection; n == 1 consumes nothing.
  // (Generator.choice(a) without p is a[integers(0, len(a))].)
  __device__ __forceinline__ uint32_t bounded(uint32_t n) {
    if (n <= 1u) return 0u;
    uint64_t m = (uint64_t)next32() * n;
    uint32_t left = (uint32_t)m;
    if (left < n) {
      const uint32_t thr = (0xffffffffu - (n - 1u)) % n;
      while (left < thr) {
        m = (uint64_t)next32() * n;
        left = (uint32_t)m;
      }
    }
    return (uint32_t)(m >> 32);
  }
};

// the per-mouse booleans of the step loop, one bit each of one register word
struct Bits {
  uint32_t w;
  template <int B>
  __device__ __forceinline__ bool get() const {
    return (w & (1u << B)) != 0u;
  }
  template <int B>
  __device__ __forceinline__ void set(bool v) {
    w = v ? (w | (1u << B)) : (w & ~(1u << B));
  }
};
enum { F_ACTIVE, F_DIRTY, F_MOVING, F_HOME, F_INSIDE, F_INZONE, F_PREVMOVED, F_LIVE };

// NT = threads per CTA (compile-time: every per-mouse table is indexed [row][thread], so row strides and the
// fixed-size tables' offsets become immediates).
// CntT = type of a visit counter: u16 when n_steps + 1 < 65536 (half the shared memory: one more resident CTA), else u32.
// SEEDED = the run continues an earlier one (d_init_visits / d_init_model_state); a separate instantiation, so that
// the fresh-start kernel keeps its register allocation.
// GMAP = the visit map lives in DEVICE memory (a.d_visit_scratch, count[tile][launched thread] u32: the 32 mice of a warp
// touch 32 consecutive words of a row) and the shared normalised value table is read from device memory too: mazes whose
// map does not fit shared memory (beyond ~50 x 50 tiles) run the same step body -- same operations, same results -- with
// only the tile / grid-line tables in shared memory (up to ~66 000 tiles).
// CW = the model's component word when it is known at compile time (kStdWord: the reference's standard model --
// anxiety, biomechanical cost, experience, biomechanical persistence, value table, in that order): the five component
// bodies follow one another in a straight line, no per-component dispatch.  0 = any component list (run-time loop).
template <int NT, typename CntT, bool SEEDED, bool GMAP, uint32_t CW = 0u>
__global__ void __launch_bounds__(NT, (((sizeof(CntT) == 2 || GMAP) ? NM_SIM_MIN_CTAS_U16 : NM_SIM_MIN_CTAS) * kSimThreads) / NT)
    nm_sim_kernel(const SimDev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nT = d.X * d.Y;
  constexpr int nt = NT;
  const int tid = threadIdx.x;
  // 2^(j/64) (1 KB region) | actions | rank / mailbox scratch | anxiety value*weight by tile code [5][nt] | biomechanical cost*weight by action [kA][nt] |
  // visit map [nT][nt] u32 | shared normalised value table [nT] | geometry tables
  double *act_sm = reinterpret_cast<double *>(smem_raw + ExpSim::kRegionBytes);  // [kA][2] relative actions
  uint8_t *rank_sm = reinterpret_cast<uint8_t *>(act_sm + 2 * kA);      // [warps][32] rank -> lane of the dirty mice
  uint16_t *res_sm = reinterpret_cast<uint16_t *>(rank_sm + nt);        // [thread][kA] walk results: tile | ok << 15
  double *anx_sm = act_sm + 2 * kA + (nt / 32) * 4 + 2 * nt;
  double *bcw_sm = anx_sm + 5 * nt;
  // visit map: count[tile * cs + column]; shared memory (cs = nt, column = tid) or device memory (GMAP)
  CntT *count = GMAP ? reinterpret_cast<CntT *>(d.a.d_visit_scratch) : reinterpret_cast<CntT *>(bcw_sm + kBcwRows * nt);
  const size_t cs = GMAP ? (size_t)d.scratch_columns : (size_t)nt;
  double *vt_sm = GMAP ? bcw_sm + kBcwRows * nt : reinterpret_cast<double *>(count + (size_t)nT * nt);
  ExpSim::fill(smem_raw, tid);
  if (tid < 2 * kA) act_sm[tid] = d.act[tid >> 1][tid & 1];
  NmSimGeom geo;  // tile table, grid-line tables (nm_simgeom.cuh)
  geo.build(reinterpret_cast<unsigned char *>(GMAP ? vt_sm : vt_sm + nT), d.tiles, d.a.d_tile_flags, d.X, d.Y, d.st);
  if (!GMAP && d.a.d_vtable && !d.a.vtable_per_mouse)
    for (int i = tid; i < nT; i += nt) vt_sm[i] = d.a.d_vtable[i];
  const nm_sim_args &a = d.a;
  const int64_t N = a.n_mice;
  const int64_t n_raw = (int64_t)blockIdx.x * nt + tid;
  CntT *cnt = count + (GMAP ? (size_t)n_raw : (size_t)tid);  // this thread's column
  // shared-memory map: one pinned shared-window address, rows nt * sizeof(CntT) bytes apart (compile-time stride)
  const uint32_t a_cnt = GMAP ? 0u : nm_pin_u32((uint32_t)__cvta_generic_to_shared(cnt));
  auto cnt_ld = [&](int tile) -> uint32_t {
    if constexpr (GMAP) {
      return (uint32_t)cnt[(size_t)tile * cs];
    } else {
      uint32_t v;
      if constexpr (sizeof(CntT) == 2)
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a_cnt + (uint32_t)tile * (uint32_t)(nt * sizeof(CntT))));
      else
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a_cnt + (uint32_t)tile * (uint32_t)(nt * sizeof(CntT))));
      return v;
    }
  };
  auto cnt_st = [&](int tile, uint32_t v) {
    if constexpr (GMAP) {
      cnt[(size_t)tile * cs] = (CntT)v;
    } else {
      if constexpr (sizeof(CntT) == 2)
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_cnt + (uint32_t)tile * (uint32_t)(nt * sizeof(CntT))), "r"(v));
      else
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(a_cnt + (uint32_t)tile * (uint32_t)(nt * sizeof(CntT))), "r"(v));
    }
  };
  if (SEEDED && a.d_init_visits) {  // continuing an earlier run: ExperienceComponent._maze_map_exp as it was left
    const uint32_t *iv = a.d_init_visits + (n_raw < N ? n_raw : N - 1) * nT;
    for (int i = 0; i < nT; ++i) cnt_st(i, iv[i]);
  } else {
    for (int i = 0; i < nT; ++i) cnt_st(i, 0u);
  }
  __syncthreads();

  // The warp shares the ray walks of its mice (below), so all 32 lanes run the step loop together: the lanes past the
  // last mouse (final CTA only) shadow mouse N - 1 and store nothing.
  const bool live = n_raw < N;
  const int64_t n = live ? n_raw : N - 1;
  const int A = a.n_actions;
  const int T = a.n_steps;
  const uint32_t zero_mask = nm_pin_u32(d.zero_mask), comp_word = nm_pin_u32(d.comp_word);
  const uint32_t null_mask = d.null_mask;  // bit i = action i is exactly (0, 0)

  int status = 0;
  ExpSim er;
  er.load(smem_raw);
  {
    // ---- per-mouse state -------------------------------------------------------------------------
    double px = a.d_init_pose[3 * n], py = a.d_init_pose[3 * n + 1], ori = a.d_init_pose[3 * n + 2];
    double prevx = a.d_init_prev[2 * n], prevy = a.d_init_prev[2 * n + 1];
    // the per-mouse booleans travel in ONE register word (as separate bools they were spilled to local memory byte by
    // byte and every test waited for a reload)
    Bits fl;
    fl.w = 0u;
    fl.set<F_LIVE>(live);
    fl.set<F_MOVING>(false);  // BehaviouralModel._moving (True when the last choice was to STAY)
    const double beta = a.d_beta[n];
    // value * weight is the same product at every step (Component.weighted_values, behavioural_components.py:277-285):
    // it is formed once here and looked up afterwards.
    // this thread's value * weight tables: anxiety by tile code at a_thr + code * nt * 8, biomechanical cost by action
    // after them -- one pinned shared-window address, every offset a compile-time constant
    uint32_t a_thr = nm_pin_u32((uint32_t)__cvta_generic_to_shared(anx_sm + tid));
    constexpr uint32_t kRow = (uint32_t)nt * 8u;
    if (a.d_anxiety) {
      const double *q = a.d_anxiety + 5 * n;
      const double w = q[0];
      nm_sts_f64(a_thr + 0u * kRow, -1.0 * w);                  // VOID
      nm_sts_f64(a_thr + 2u * kRow, q[1] * w);                  // CORNER: p1                    behavioural_components.py:364
      nm_sts_f64(a_thr + 1u * kRow, (q[2] * q[1]) * w);         // WALL:   p2 * p1
      nm_sts_f64(a_thr + 3u * kRow, (q[3] * q[2] * q[1]) * w);  // CENTRE: p3 * p2 * p1
      nm_sts_f64(a_thr + 4u * kRow, q[4] * w);                  // DECISION_POINT: p4            :332-333
    }
    double bc_zero = 0.0;  // 0.0 * W_bcost: the component's contribution when the mouse has just moved (:410-413)
    double bc_w = 0.0;
    const double *bc_row = a.d_bcost_table ? a.d_bcost_table + n * A : nm_zero_row;
    if (a.d_bcost_table) {
      const double w = a.d_bcost_weight[n];
      bc_zero = 0.0 * w;
      bc_w = w;
#if !NM_SIM_BCW_GLOBAL
#pragma unroll
      for (int i = 0; i < kA; ++i) nm_sts_f64(a_thr + (5u + (uint32_t)i) * kRow, i < A ? a.d_bcost_table[n * A + i] * w : 0.0);
#endif
    }
    double ex_w = 0, xi = 0, k_home = 0, k_explo = 0, th_explo = 0, th_home = 0, fam = 0.0;
    // F_HOME clear: ExperienceComponent.State.EXPLO at start (:440)
    if (a.d_experience) {
      const double *q = a.d_experience + 6 * n;
      ex_w = q[0]; xi = q[1]; k_home = q[2]; k_explo = q[3]; th_explo = q[4]; th_home = q[5];
    }
    if (SEEDED && a.d_init_model_state) {  // the model and the component carry on where an earlier run left them
      const double *q = a.d_init_model_state + 3 * n;
      fl.set<F_MOVING>(q[0] != 0.0);
      fam = q[1];
      fl.set<F_HOME>(q[2] != 0.0);
    }
    a_thr = nm_pin_u32(a_thr);  // re-pinned after the stores above: the step loop's (plain) loads cannot move above them
    double bpw_plain = 0, bpw_m = 0, bpw_nm = 0;  // bp * W, (Wm*bp) * W, (Wnm*bp) * W   (:502-513)
    if (a.d_bpersistence) {
      const double *q = a.d_bpersistence + 4 * n;
      bpw_plain = q[1] * q[0];
      bpw_m = (q[2] * q[1]) * q[0];
      bpw_nm = (q[3] * q[1]) * q[0];
    }
    double vt_w = 0;
    const double *vt = nullptr;
    if (a.d_vtable) {
      vt_w = a.d_vtable_weight[n];
      vt = a.vtable_per_mouse ? a.d_vtable + n * nT : (GMAP ? a.d_vtable : vt_sm);
    }
    const uint32_t a_vt = nm_pin_u32((uint32_t)__cvta_generic_to_shared(vt_sm));  // the shared normalised value table
    Pcg64 rng;
    rng.state = 0;
    rng.inc = 0;
    rng.has32 = 0u;
    rng.buf32 = 0u;
    if (a.d_pcg_state) {
      const uint64_t *q = a.d_pcg_state + 4 * n;
      rng.state = ((unsigned __int128)q[0] << 64) | q[1];
      rng.inc = ((unsigned __int128)q[2] << 64) | q[3];
      if (a.d_pcg_u32) {
        rng.has32 = a.d_pcg_u32[2 * n];
        rng.buf32 = a.d_pcg_u32[2 * n + 1];
      }
    } else if (a.use_device_seed) {
      // pcg64_srandom(initstate, initseq) with both 128-bit values drawn from splitmix64(device_seed, mouse index)
      uint64_t z = a.device_seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(4 * (n + a.mouse_index_base) + 1);
      uint64_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        z += 0x9E3779B97F4A7C15ULL;
        uint64_t y = z;
        y = (y ^ (y >> 30)) * 0xBF58476D1CE4E5B9ULL;
        y = (y ^ (y >> 27)) * 0x94D049BB133111EBULL;
        w[i] = y ^ (y >> 31);
      }
      const unsigned __int128 initstate = ((unsigned __int128)w[0] << 64) | w[1];
      const unsigned __int128 initseq = ((unsigned __int128)w[2] << 64) | w[3];
      rng.state = 0;
      rng.inc = (initseq << 1) | 1u;
      rng.next64();
      rng.state += initstate;
      rng.next64();
    }

    // ---- fused metrics over the history positions (history[0] = initial position, mouse.cpp:58-61) ----
    int n_visited = 0, it_visit = T + 1, count_moving = 0;
    // shock zone (latency_enter_shock_zone / shock_zone_entries, metrics.py:488-586), corridors (arms, :735-769),
    // not-moving runs (hist_time_moving -> compute_static_intervals, :445-467, 621-635): moving_bool_list[0] is False,
    // so the history opens with a not-moving run
    int zone_latency = T, zone_entries = 0, corridor_switches = 0, corridor_side = 0, static_runs = 1;

    int ohist[NM_MAX_ORI_BINS + 1];
#pragma unroll
    for (int i = 0; i <= NM_MAX_ORI_BINS; ++i) ohist[i] = 0;
    int cur_tx = geo.tile(px), cur_ty = geo.tile(py);
    auto record_position = [&](int t_idx, int tx, int ty, bool inside) {
      // occupancy_map / tile_analysis / it_perc_visited_maze on history entry t_idx
      if (!inside) return;
      const int tile = tx * d.Y + ty;
      if (a.d_tile_flags) {
        // staged in shared memory with the tile table (mazes of up to 4096 tiles), else read from device memory
        const unsigned f = NmSimGeom::flags_staged(d.X, d.Y) ? geo.tile_flags(tile) : a.d_tile_flags[tile];
        if (f & NM_TILE_ZONE_IN) {
          if (zone_latency == T && t_idx < T) zone_latency = t_idx;  // "never" and "at the last position" are both T
          if (!fl.get<F_INZONE>()) {
            zone_entries += 1;
            fl.set<F_INZONE>(true);
          }
        } else if ((f & NM_TILE_ZONE_OUT) && fl.get<F_INZONE>()) {
          fl.set<F_INZONE>(false);
        }
        const int side = (int)((f >> 2) & 3u);
        if (side) {
          if (corridor_side && corridor_side != side) corridor_switches += 1;
          corridor_side = side;
        }
      }
      const uint32_t c = cnt_ld(tile);
      cnt_st(tile, c + 1u);
      if (c == 0u) {
        n_visited += 1;
        // perc = n_visited * 100 / number_visitable_tiles >= percentage      metrics.py:711-721
        // (the quotient is monotone in n_visited: the host evaluates it once per count and passes the first count that
        //  reaches the percentage, nm_simulate)
        if (it_visit == T + 1 && n_visited >= d.n_visit_thr) it_visit = t_idx;
      }
    };
    fl.set<F_INSIDE>(cur_tx >= 0 && cur_tx < d.X && cur_ty >= 0 && cur_ty < d.Y);
    // NOTE: the ExperienceComponent increments count[x, y] of the CURRENT tile at every decision
    // (:456) -- that is exactly "occupancy over history[0..t]", so history entry t is recorded at the
    // top of step t and the final entry after the loop.
    if (live && a.d_hist_pos) {
      a.d_hist_pos[2 * n] = px;
      a.d_hist_pos[2 * n + 1] = py;
    }
    if (live && a.d_hist_ori) a.d_hist_ori[n] = ori;

    // Reachability of the A endpoints (experiment.py:36-38) depends on the pose only, and most mice keep their pose
    // in a given step (the null action wins): every mouse caches its reachability mask `vis` and the tiles under its
    // endpoints (16 bits each), and the walks of the mice whose pose did change ("dirty") are spread over all 32
    // lanes of the warp -- item = (dirty mouse, action), 32 items per round, inputs and results travel by shuffles.
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = tid & 31;
    uint8_t *rank_w = rank_sm + (tid >> 5) * 32;
    uint16_t *res_w = res_sm + (tid >> 5) * (32 * kA);  // the warp's mailbox: [lane][action]
    const int n_walk = d.n_walk;
    const unsigned div_walk = n_walk > 0 ? (65536u + (unsigned)n_walk - 1u) / (unsigned)n_walk : 0u;  // item / n_walk
    fl.set<F_ACTIVE>(true);
    fl.set<F_DIRTY>(true);
    unsigned vis = 0u;
    uint4 et = make_uint4(0u, 0u, 0u, 0u);  // eight 16-bit entries: tile under endpoint i | reachable << 15
    double c = 1.0, s = 0.0;
    auto et_get = [&](int i) -> int {  // i is a compile-time constant after unrolling
      const uint32_t w = i < 2 ? et.x : i < 4 ? et.y : i < 6 ? et.z : et.w;
      return (int)((i & 1) ? ((w >> 16) & 0x7fffu) : (w & 0x7fffu));
    };
#pragma unroll
    for (int i = 0; i < kA; ++i) res_w[lane * kA + i] = 0;

    for (int t = 0; t < T; ++t) {
      __syncwarp();
      const bool active = fl.get<F_ACTIVE>();
      if (__ballot_sync(kFull, active) == 0u) break;
      const bool cur_inside = fl.get<F_INSIDE>();
      if (active) record_position(t, cur_tx, cur_ty, cur_inside);
      const bool need = active && fl.get<F_DIRTY>();
      const unsigned dmask = __ballot_sync(kFull, need);
      if (dmask != 0u) {
        const bool start_ok = cur_inside && geo.tile_code(cur_tx * d.Y + cur_ty) != 0;  // tile under the mouse (maze.py:578)
        const int cur_tile = cur_inside ? cur_tx * d.Y + cur_ty : 0;
        if (need) {
          nm_host_sincos(ori, &s, &c);
          // null displacement: the endpoint IS the position, R(ori)*0 + p == p bit for bit, and no grid line is
          // crossed -- reachable iff the tile under the mouse is visitable
#pragma unroll 1
          for (uint32_t m = null_mask; m != 0u; m &= m - 1u)
            res_w[lane * kA + (__ffs((int)m) - 1)] = (uint16_t)(cur_tile | (start_ok ? 0x8000 : 0));
        }
        const int nd = __popc(dmask);
        if (n_walk > 0) {
          const int my_rank = __popc(dmask & ((1u << lane) - 1u));
          if (need) rank_w[my_rank] = (uint8_t)lane;
          __syncwarp();
          const int info = (start_ok ? 1 : 0) | ((cur_inside ? cur_tx : 0) << 1) | ((cur_inside ? cur_ty : 0) << 16);
          const int items = nd * n_walk;
#pragma unroll 1
          for (int base = 0; base < items; base += 32) {
            const int it = base + lane;
            const bool valid = it < items;
            const int r = valid ? (int)(((unsigned)it * div_walk) >> 16) : 0;  // exact for it < 256
            const int j = valid ? it - r * n_walk : 0;
            const int i = (int)((d.walk_list >> (4 * j)) & 7u);
            const int src = rank_w[r];
            const double spx = __shfl_sync(kFull, px, src), spy = __shfl_sync(kFull, py, src);
            const double sc = __shfl_sync(kFull, c, src), ss = __shfl_sync(kFull, s, src);
            const int sinfo = __shfl_sync(kFull, info, src);
            double ex, ey;
            nm_pose_endpoint(sc, ss, spx, spy, act_sm[2 * i], act_sm[2 * i + 1], &ex, &ey);
            int tile_e;
            const bool ok = geo.move_ok((sinfo & 1) != 0, (sinfo >> 1) & 0x7fff, sinfo >> 16, spx, spy, ex, ey, &tile_e);
            if (valid) res_w[src * kA + i] = (uint16_t)(tile_e | (ok ? 0x8000 : 0));  // into the owner's mailbox
          }
        }
        __syncwarp();
        if (need) {
          et = *reinterpret_cast<const uint4 *>(res_w + lane * kA);
          vis = ((et.x >> 15) & 1u) | ((et.x >> 30) & 2u) | ((et.y >> 13) & 4u) | ((et.y >> 28) & 8u) |
                ((et.z >> 11) & 16u) | ((et.z >> 26) & 32u) | ((et.w >> 9) & 64u) | ((et.w >> 24) & 128u);
        }
        fl.set<F_DIRTY>(false);
      }
      if (!active) continue;
      int etile[kA];
#pragma unroll
      for (int i = 0; i < kA; ++i) etile[i] = et_get(i);
      const int K = __popc(vis);
      if (K == 0) {  // np.max([]) raises in the reference (exploration_model.py:92)
        status = t + 1;
        fl.set<F_ACTIVE>(false);
        continue;
      }
      int act_idx = -1;
      if (CW != 0u || a.model_kind == 0) {  // ---------------- BehaviouralModel -----------------------------------------
      // -- component updates                                                 exploration_model.py:79-81
      if (a.d_experience) {  // :456-466 (count of the current tile was just incremented above)
        const double nvis = cur_inside ? (double)cnt_ld(cur_tx * d.Y + cur_ty) : 0.0;
        fam = (1.0 - xi) * fam + xi * nvis;
        if (fl.get<F_HOME>() && fam >= th_explo) fl.set<F_HOME>(false);
        else if (!fl.get<F_HOME>() && fam < th_home) fl.set<F_HOME>(true);
      }
      // -- val_fin = beta * sum_c weighted_values_c                           :84-87
      // (candidates stay in action order: the reference's list is the action list filtered by
      //  reachability, so "k-th candidate" == "k-th reachable action" and nothing needs compacting)
      // Components outermost (one uniform dispatch per component and step), the eight endpoints innermost and
      // unrolled; every endpoint is evaluated (an unreachable one reads the tile it would land on, or tile 0; slots
      // beyond the A actions read tile 0 and zero tables -- no per-slot `i < A` test) and the reachable ones are
      // selected afterwards, so the warp stays converged.  Per endpoint the sum still runs over the
      // components in model order starting from 0.0 (:84-86).
      const bool moving = fl.get<F_MOVING>(), home = fl.get<F_HOME>();
      double val[kA];
#pragma unroll
      for (int i = 0; i < kA; ++i) val[i] = 0.0;
      // one component's weighted values added to the running sums; KIND is a compile-time constant
      auto component = [&](auto kind_tag) {
        constexpr int KIND = decltype(kind_tag)::value;
        if constexpr (KIND == NM_COMP_ANXIETY) {
#pragma unroll
          for (int i = 0; i < kA; ++i) {
            const int code = (int)geo.tile_code(etile[i]);
            val[i] = val[i] + nm_lds_f64(a_thr + (uint32_t)(code > 4 ? 0 : code) * kRow);
          }
        } else if constexpr (KIND == NM_COMP_BCOST) {  // cached[is_visitable] if moving else zeros   :410-413
#pragma unroll
          for (int i = 0; i < kA; ++i)
#if NM_SIM_BCW_GLOBAL
            val[i] = val[i] + (moving ? nm_ldg_early(bc_row + (i < A ? i : A - 1)) * bc_w : bc_zero);
#else
            val[i] = val[i] + (moving ? nm_lds_f64(a_thr + (5u + (uint32_t)i) * kRow) : bc_zero);
#endif
        } else if constexpr (KIND == NM_COMP_EXPERIENCE) {  // :473-479
          // home: (1 - e) * W, else e * W -- fma(sg, e, one) is fl(1 - e) resp. e itself (0 + e), one instruction
          const double kk = home ? -k_home : -k_explo, sg = home ? -1.0 : 1.0, one = home ? 1.0 : 0.0;
#pragma unroll
          for (int i = 0; i < kA; ++i) {
            // (double)count by the 2^52 trick: exact for a u32, and no I2F on the quarter-rate pipe
            const double cn = __hiloint2double(0x43300000, (int)cnt_ld(etile[i])) - 4503599627370496.0;
            const double ee = er.exp_nonpos(kk * cn);
            val[i] = val[i] + fma(sg, ee, one) * ex_w;
          }
        } else if constexpr (KIND == NM_COMP_BPERSISTENCE) {  // :508-513
          const double vz = moving ? bpw_m : bpw_plain, vn = moving ? bpw_plain : bpw_nm;
#pragma unroll
          for (int i = 0; i < kA; ++i) val[i] = val[i] + ((zero_mask >> i) & 1u ? vz : vn);
        } else if constexpr (KIND == NM_COMP_VTABLE) {  // :537-541
          if (!GMAP && !a.vtable_per_mouse) {  // the shared table: pinned shared-window address
#pragma unroll
            for (int i = 0; i < kA; ++i) val[i] = val[i] + nm_lds_f64(a_vt + 8u * (uint32_t)etile[i]) * vt_w;
          } else {
#pragma unroll
            for (int i = 0; i < kA; ++i) val[i] = val[i] + vt[etile[i]] * vt_w;
          }
        }
      };
      if constexpr (CW != 0u) {  // the component list is known: straight line
        if constexpr (((CW >> 0) & 15u) != 0u) component(std::integral_constant<int, (int)((CW >> 0) & 15u)>{});
        if constexpr (((CW >> 4) & 15u) != 0u) component(std::integral_constant<int, (int)((CW >> 4) & 15u)>{});
        if constexpr (((CW >> 8) & 15u) != 0u) component(std::integral_constant<int, (int)((CW >> 8) & 15u)>{});
        if constexpr (((CW >> 12) & 15u) != 0u) component(std::integral_constant<int, (int)((CW >> 12) & 15u)>{});
        if constexpr (((CW >> 16) & 15u) != 0u) component(std::integral_constant<int, (int)((CW >> 16) & 15u)>{});
      } else {
#pragma unroll 1
        // the active components' kinds travel in ONE register word (an inactive component contributes integer 0,
        // :283-285, and is left out on the host): no per-step loads from the launch-parameter arrays
        for (uint32_t cw = comp_word; cw != 0u; cw >>= 4) {
          switch (cw & 15u) {
            case NM_COMP_ANXIETY: component(std::integral_constant<int, NM_COMP_ANXIETY>{}); break;
            case NM_COMP_BCOST: component(std::integral_constant<int, NM_COMP_BCOST>{}); break;
            case NM_COMP_EXPERIENCE: component(std::integral_constant<int, NM_COMP_EXPERIENCE>{}); break;
            case NM_COMP_BPERSISTENCE: component(std::integral_constant<int, NM_COMP_BPERSISTENCE>{}); break;
            case NM_COMP_VTABLE: component(std::integral_constant<int, NM_COMP_VTABLE>{}); break;
            default: break;
          }
        }
      }
      // -- val_fin = beta * val; the unreachable slots are parked at -inf, so that the maximum over the candidates is a
      // plain tree of compare-selects over all eight (NaN is out of contract; which of two equal values wins changes
      // nothing below)
      double vmax;
      {
#pragma unroll
        for (int i = 0; i < kA; ++i) {
          const double b = beta * val[i];
          val[i] = (vis & (1u << i)) != 0u ? b : -INFINITY;
        }
        const double m01 = val[0] > val[1] ? val[0] : val[1], m23 = val[2] > val[3] ? val[2] : val[3];
        const double m45 = val[4] > val[5] ? val[4] : val[5], m67 = val[6] > val[7] ? val[6] : val[7];
        const double m03 = m01 > m23 ? m01 : m23, m47 = m45 > m67 ? m45 : m67;
        vmax = m03 > m47 ? m03 : m47;
      }
      // -- softmax + numpy Generator.choice                                   :92-98
      // every slot is evaluated (an unreachable one on -inf: whatever comes out is discarded) and the unreachable ones
      // are set to 0 afterwards: adding +0.0 changes no partial sum, so the sums below run over all eight slots without
      // a test and give the bits of the sums over the candidates
      double e[kA];
#pragma unroll
      for (int i = 0; i < kA; ++i) {
        const double ex = er.exp_nonpos(val[i] - vmax);
        e[i] = (vis & (1u << i)) != 0u ? ex : 0.0;
      }
      // numpy's sum: pairwise for n == 8, sequential from the first candidate for n < 8
      const double sum_pw = ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
      double sum_sq = e[0];
#pragma unroll
      for (int i = 1; i < kA; ++i) sum_sq = sum_sq + e[i];
      const double sum = K == kA ? sum_pw : sum_sq;
      // p = e / sum, cdf = cumsum(p), choice = first k with cdf_k / cdf_last > u.  The 2 K divisions are replaced by
      // one reciprocal and K + 1 multiplications (p = e * (1/sum); cdf_k > u * cdf_last): like the device exp, this
      // moves p by <= 1 ulp against numpy, which can only change a choice when u falls within ~1e-16 of a cdf step.
      // The cdf is non-decreasing and flat across unreachable slots, so the first candidate whose cdf exceeds the
      // threshold is the NUMBER of slots at or below it (searchsorted, side='right').
      const double inv_sum = 1.0 / sum;
      double cdf[kA];
      double run = e[0] * inv_sum;
      cdf[0] = run;
#pragma unroll
      for (int i = 1; i < kA; ++i) {
        run = run + e[i] * inv_sum;  // cumsum
        cdf[i] = run;
      }
      const double u = a.d_draws ? a.d_draws[(int64_t)t * N + n] : rng.next_double();
      const double thr = u * run;  // u * cdf[-1]
      int below = 0;
#pragma unroll
      for (int i = 0; i < kA; ++i) below += (cdf[i] <= thr) ? 1 : 0;
      const int last_vis = 31 - __clz((int)vis);
      act_idx = below < last_vis ? below : last_vis;  // beyond the last candidate: cannot happen for u < 1
      } else {  // ---------------- RandomModel: rng.choice(next_positions_cont)   exploration_model.py:141 ----
        uint32_t pick;
        if (a.d_draws) {  // replayed uniforms: floor(u * K) (convenience mode, not numpy's draw)
          const double u = a.d_draws[(int64_t)t * N + n];
          pick = (uint32_t)(u * (double)K);
          if (pick >= (uint32_t)K) pick = (uint32_t)K - 1u;
        } else {
          pick = rng.bounded((uint32_t)K);
        }
        uint32_t seen = 0u;
#pragma unroll
        for (int i = 0; i < kA; ++i)
          if (vis & (1u << i)) {
            if (seen == pick) act_idx = i;
            ++seen;
          }
      }
      double nx, ny;  // the chosen endpoint, recomputed (same operations, same bits)
      nm_pose_endpoint(c, s, px, py, d.act[act_idx][0], d.act[act_idx][1], &nx, &ny);
      // -- moving flag                                                        :101-107
      fl.set<F_MOVING>((prevx == nx) && (prevy == ny));
      prevx = nx;
      prevy = ny;
      // -- Mouse.move_to                                                      mouse.cpp:75-87
      const double ori_before = ori;
      if (!nm_pos_is_approx(nx, ny, px, py)) {
        ori = nm_host_atan2(ny - py, nx - px);
        px = nx;
        py = ny;
        fl.set<F_DIRTY>(true);  // new pose: the cached reachability is stale
      }
      const int ntx = geo.tile(px), nty = geo.tile(py);
      const bool tile_changed = ntx != cur_tx || nty != cur_ty;
      if (tile_changed) count_moving += 1;  // motion_analysis, metrics.py:437-440
      if (fl.get<F_PREVMOVED>() && !tile_changed) static_runs += 1;
      fl.set<F_PREVMOVED>(tile_changed);
      if (a.n_ori_bins > 0 && tile_changed &&
          (a.ori_tile_filter == 0u || (cur_inside && ((a.ori_tile_filter >> geo.tile_code(cur_tx * d.Y + cur_ty)) & 1u)))) {
        // hist_orientations (:329-356): the turn of this step, brought back by one full turn when it leaves
        // [min_lim, max_lim]; numpy.histogram's bins (left-closed, the last one closed on both sides)
        double rel = ori - ori_before;
        if (rel > a.ori_max_lim) rel = rel - 6.283185307179586;
        else if (rel < a.ori_min_lim) rel = rel - (-6.283185307179586);
        const int nb = a.n_ori_bins;
        if (rel >= d.ori_edges[0] && rel <= d.ori_edges[nb]) {
          int b = 0;
          for (int k = 1; k < nb; ++k) b += (rel >= d.ori_edges[k]) ? 1 : 0;
          const int ddx = ntx - cur_tx, ddy = nty - cur_ty;
          if (a.ori_distance && b == 2 && ddx * ddx + ddy * ddy > 2) b = nb;  // math.dist(...) > 1.5   (:339-343)
          ohist[b] += 1;
        }
      }
      cur_tx = ntx;
      cur_ty = nty;
      fl.set<F_INSIDE>(cur_tx >= 0 && cur_tx < d.X && cur_ty >= 0 && cur_ty < d.Y);
      if (fl.get<F_LIVE>() && a.d_hist_pos) {
        a.d_hist_pos[((int64_t)(t + 1) * N + n) * 2] = px;
        a.d_hist_pos[((int64_t)(t + 1) * N + n) * 2 + 1] = py;
      }
      if (fl.get<F_LIVE>() && a.d_hist_ori) a.d_hist_ori[(int64_t)(t + 1) * N + n] = ori;
      if (fl.get<F_LIVE>() && a.d_hist_choice) a.d_hist_choice[(int64_t)t * N + n] = (uint8_t)act_idx;
    }
    if (status == 0) record_position(T, cur_tx, cur_ty, fl.get<F_INSIDE>());

    if (live) {
    if (a.d_final_pose) {
      a.d_final_pose[3 * n] = px;
      a.d_final_pose[3 * n + 1] = py;
      a.d_final_pose[3 * n + 2] = ori;
    }
    if (a.d_pcg_state_out) {
      uint64_t *q = a.d_pcg_state_out + 4 * n;
      q[0] = (uint64_t)(rng.state >> 64); q[1] = (uint64_t)rng.state;
      q[2] = (uint64_t)(rng.inc >> 64);   q[3] = (uint64_t)rng.inc;
    }
    if (a.d_pcg_u32_out) {
      a.d_pcg_u32_out[2 * n] = rng.has32;
      a.d_pcg_u32_out[2 * n + 1] = rng.buf32;
    }
    if (a.d_it_visit) a.d_it_visit[n] = it_visit;
    if (a.d_count_moving) a.d_count_moving[n] = count_moving;
    if (a.d_tile_counts) {  // tile_analysis counts = the visit map summed by tile code (metrics.py:219-248)
#pragma unroll 1
      for (int code = 0; code < 8; ++code) {
        int acc = 0;
        for (int tile = 0; tile < nT; ++tile)
          if ((int)(geo.tile_code(tile) & 7u) == code) acc += (int)cnt_ld(tile);
        a.d_tile_counts[8 * n + code] = acc;
      }
    }
    if (a.d_status) a.d_status[n] = status;
    if (a.d_zone_metrics) {
      int zone_occ = 0;  // positions inside the zone = the visit map summed over the zone's tiles
      if (a.d_tile_flags)
        for (int tile = 0; tile < nT; ++tile)
          if (a.d_tile_flags[tile] & NM_TILE_ZONE_IN) zone_occ += (int)cnt_ld(tile);
      int *q = a.d_zone_metrics + 4 * n;
      q[0] = zone_latency; q[1] = zone_occ; q[2] = zone_entries; q[3] = corridor_switches;
    }
    if (a.d_static_intervals) a.d_static_intervals[n] = static_runs;
    if (a.d_subarea_counts) {
#pragma unroll 1
      for (int sub = 0; sub < NM_MAX_SUBAREAS; ++sub) {
        int acc = 0;
        for (int tile = 0; tile < nT; ++tile)
          if (d.subareas[tile] == sub) acc += (int)cnt_ld(tile);
        a.d_subarea_counts[(int64_t)NM_MAX_SUBAREAS * n + sub] = acc;
      }
    }
    if (a.d_ori_hist)
      for (int b = 0; b <= a.n_ori_bins; ++b) a.d_ori_hist[(int64_t)(a.n_ori_bins + 1) * n + b] = ohist[b];
    if (a.d_model_state) {
      double *q = a.d_model_state + 5 * n;
      q[0] = prevx; q[1] = prevy; q[2] = fl.get<F_MOVING>() ? 1.0 : 0.0; q[3] = fam; q[4] = fl.get<F_HOME>() ? 1.0 : 0.0;
    }
    }  // live
  }

  // ---- occupancy maps: coalesced write through a transposed read of count[tile][thread] -------------
  if (a.d_occupancy || a.d_occupancy_total) {
    __syncthreads();
    const int64_t n0 = (int64_t)blockIdx.x * nt;
    const int64_t np = (N - n0 < nt) ? (N - n0) : nt;
    if (a.d_occupancy) {
      uint32_t *dst = a.d_occupancy + n0 * nT;
      for (int64_t i = tid; i < np * nT; i += nt) {
        const int u = (int)(i / nT), tile = (int)(i - (int64_t)u * nT);
        dst[i] = count[(size_t)tile * cs + (GMAP ? (size_t)n0 : (size_t)0) + u];
      }
    }
    if (a.d_occupancy_total) {
      for (int tile = tid; tile < nT; tile += nt) {
        unsigned long long acc = 0ull;
        for (int u = 0; u < (int)np; ++u) acc += count[(size_t)tile * cs + (GMAP ? (size_t)n0 : (size_t)0) + u];
        if (acc) atomicAdd(a.d_occupancy_total + tile, acc);
      }
    }
  }
}

__global__ void __launch_bounds__(256) nm_segments_kernel(const uint8_t *__restrict__ g_tiles, int X, int Y, double st,
                                                          const double *__restrict__ seg, int64_t n,
                                                          uint8_t *__restrict__ out) {
  // the simulation kernel's geometry (tabulated grid lines, branch-free walk) on caller-supplied segments
  extern __shared__ __align__(16) unsigned char sm_geo[];
  NmSimGeom geo;
  geo.build(sm_geo, g_tiles, nullptr, X, Y, st);
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x1 = seg[4 * i], y1 = seg[4 * i + 1], x2 = seg[4 * i + 2], y2 = seg[4 * i + 3];
    const int tx = geo.tile(x1), ty = geo.tile(y1);
    const bool s_in = (unsigned)tx < (unsigned)X && (unsigned)ty < (unsigned)Y;
    const bool start_ok = s_in && geo.tile_code(s_in ? tx * Y + ty : 0) != 0;
    int et;
    out[i] = geo.move_ok(start_ok, tx, ty, x1, y1, x2, y2, &et) ? 1 : 0;
  }
}

}  // namespace

extern "C" int nm_simulate(const nm_maze *mz, const nm_sim_args *args, void *stream) {
  NM_CHECK_ARG(mz && args, "nm_simulate: NULL argument");
  NM_CHECK_ARG(mz->device >= 0, "nm_simulate: the maze has not been uploaded to a CUDA device");
  const nm_sim_args &a = *args;
  NM_CHECK_ARG(a.n_mice >= 0 && a.n_steps >= 0 && a.mouse_index_base >= 0, "nm_simulate: negative size");
  NM_CHECK_ARG(a.n_actions >= 1 && a.n_actions <= NM_MAX_CAND && a.h_actions,
               "nm_simulate: need 1..%d relative actions", NM_MAX_CAND);
  NM_CHECK_ARG(a.n_components >= 0 && a.n_components <= NM_MAX_COMPONENTS, "nm_simulate: too many components");
  NM_CHECK_ARG(a.d_beta && a.d_init_pose && a.d_init_prev, "nm_simulate: NULL beta / init_pose / init_prev");
  NM_CHECK_ARG((a.d_draws != nullptr) + (a.d_pcg_state != nullptr) + (a.use_device_seed != 0) == 1,
               "nm_simulate: give exactly one of d_draws (replayed uniforms), d_pcg_state and use_device_seed");
  for (int i = 0; i < a.n_components; ++i) {
    const int k = a.component_kind[i];
    const bool ok = (k == NM_COMP_ANXIETY && a.d_anxiety) || (k == NM_COMP_BCOST && a.d_bcost_table && a.d_bcost_weight) ||
                    (k == NM_COMP_EXPERIENCE && a.d_experience) || (k == NM_COMP_BPERSISTENCE && a.d_bpersistence) ||
                    (k == NM_COMP_VTABLE && a.d_vtable && a.d_vtable_weight);
    NM_CHECK_ARG(ok, "nm_simulate: component %d (kind %d) has no parameter array", i, k);
    if ((k == NM_COMP_BCOST || k == NM_COMP_BPERSISTENCE))
      NM_CHECK_ARG(a.h_zero_action, "nm_simulate: h_zero_action is required by the biomechanical components");
  }
  NM_CHECK_ARG(a.n_ori_bins >= 0 && a.n_ori_bins <= NM_MAX_ORI_BINS, "nm_simulate: at most %d orientation bins",
               NM_MAX_ORI_BINS);
  NM_CHECK_ARG(a.n_ori_bins == 0 || (a.h_ori_edges && a.d_ori_hist), "nm_simulate: the orientation histogram needs "
               "h_ori_edges and d_ori_hist");
  NM_CHECK_ARG(!a.ori_distance || a.n_ori_bins == 6, "nm_simulate: hist_orientations(distance=True) is defined for "
               "six-bin histograms");
  NM_CHECK_ARG(!a.d_zone_metrics || a.d_tile_flags, "nm_simulate: d_zone_metrics needs d_tile_flags");
  NM_CHECK_ARG(!a.d_subarea_counts || mz->d_subareas, "nm_simulate: d_subarea_counts needs a maze with subareas");
  if (a.n_mice == 0) return NM_OK;
  if (cudaSetDevice(mz->device) != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaSetDevice(%d) failed", mz->device);
  SimDev d;
  d.a = a;
  d.subareas = mz->d_subareas;
  for (int i = 0; i <= NM_MAX_ORI_BINS; ++i) d.ori_edges[i] = (a.n_ori_bins > 0 && i <= a.n_ori_bins) ? a.h_ori_edges[i] : 0.0;
  if (a.n_ori_bins == 0) d.a.d_ori_hist = nullptr;
  d.tiles = mz->d_tiles;
  d.X = mz->X;
  d.Y = mz->Y;
  d.n_visitable = mz->num_states;
  d.st = mz->size_tile;
  d.zero_mask = 0u;
  d.comp_word = 0u;
  {
    int slot = 0;
    for (int i = 0; i < a.n_components; ++i)
      if (a.component_active[i]) d.comp_word |= (uint32_t)(a.component_kind[i] & 15) << (4 * slot++);
  }
  for (int i = 0; i < kA; ++i) {
    d.act[i][0] = i < a.n_actions ? a.h_actions[2 * i] : 0.0;
    d.act[i][1] = i < a.n_actions ? a.h_actions[2 * i + 1] : 0.0;
    d.zero_act[i] = (i < a.n_actions && a.h_zero_action) ? a.h_zero_action[i] : 0;
    if (d.zero_act[i]) d.zero_mask |= 1u << i;
  }
  // it_perc_visited_maze (metrics.py:711-721): first count of visited tiles whose percentage reaches the bar -- the
  // same two conversions and division the kernel would do at every first visit, once per possible count here
  d.n_visit_thr = 2147483647;
  for (int nv = 1; nv <= mz->X * mz->Y + 1; ++nv)
    if ((double)(nv * 100) / (double)d.n_visitable >= a.visit_percentage) {
      d.n_visit_thr = nv;
      break;
    }
  d.n_walk = 0;
  d.walk_list = 0u;
  d.null_mask = 0u;
  for (int i = 0; i < kA; ++i) {
    d.null_act[i] = (i < a.n_actions && d.act[i][0] == 0.0 && d.act[i][1] == 0.0) ? 1 : 0;
    if (d.null_act[i]) d.null_mask |= 1u << i;
    d.walk_ord[i] = (uint8_t)d.n_walk;
    if (i < a.n_actions && !d.null_act[i]) {
      d.walk_list |= (uint32_t)i << (4 * d.n_walk);
      d.n_walk += 1;
    }
  }
  const int nT = mz->X * mz->Y;
  int threads = kSimThreads;
  const size_t geo_bytes = NmSimGeom::smem_bytes(mz->X, mz->Y);
  // 2^(j/64) region | actions | rank -> lane scratch | per-mouse anxiety and biomechanical-cost tables | visit map | shared
  // value table | geometry
  // a tile is visited at most n_steps + 1 times; NM_SIM_COUNTERS=32 forces the wide variant (tests)
  const char *env_cnt = getenv("NM_SIM_COUNTERS");
  const int64_t count_max = (int64_t)a.n_steps + (a.d_init_visits ? (int64_t)a.init_visits_max : 0);
  NM_CHECK_ARG(count_max < 4294967295LL, "nm_simulate: seeded visits + n_steps overflow the 32-bit visit counters");
  bool small_counts = count_max < 65535 && !(env_cnt && env_cnt[0] == '3');
  size_t cnt_bytes = small_counts ? sizeof(uint16_t) : sizeof(uint32_t);
  const size_t per_thread = (size_t)(1 + 2 * kA) + (size_t)(5 + kBcwRows) * sizeof(double);
  auto smem_for = [&](int th) {
    return kSimFixedBytes + per_thread * th + (size_t)nT * th * cnt_bytes + (size_t)nT * sizeof(double) + geo_bytes;
  };
  // the visit map in device memory: mazes whose map does not fit shared memory even for one warp, or NM_SIM_MAP=global
  const char *env_map = getenv("NM_SIM_MAP");
  const bool gmap = smem_for(32) > 227 * 1024 || (env_map && env_map[0] == 'g');
  size_t smem;
  if (gmap) {
    small_counts = false;
    cnt_bytes = sizeof(uint32_t);
    smem = kSimFixedBytes + per_thread * threads + geo_bytes;
    NM_CHECK_ARG(smem <= 227 * 1024, "nm_simulate: the tile and grid-line tables of a maze of %d tiles do not fit shared "
                 "memory (about 66 000 tiles at most)", nT);
    const int64_t columns = ((a.n_mice + threads - 1) / threads) * threads;
    NM_CHECK_ARG(a.d_visit_scratch && a.visit_scratch_columns >= columns,
                 "nm_simulate: a maze of %d tiles keeps its visit map in device memory: pass d_visit_scratch with "
                 "X*Y x nm_simulate_scratch_columns(n_mice) = %lld 32-bit words", nT, (long long)columns);
    d.scratch_columns = a.visit_scratch_columns;
  } else {
    while (threads > 32 && smem_for(threads) > 200 * 1024) threads /= 2;
    smem = smem_for(threads);
    d.scratch_columns = 0;
  }
  void (*kern)(const SimDev);
  const bool seeded = a.d_init_visits || a.d_init_model_state;
#define NM_SIM_PICK(CNT, SEED) \
  (threads == 128 ? nm_sim_kernel<128, CNT, SEED, false> : threads == 64 ? nm_sim_kernel<64, CNT, SEED, false> \
                                                                            : nm_sim_kernel<32, CNT, SEED, false>)
  if (gmap)
    kern = seeded ? nm_sim_kernel<128, uint32_t, true, true> : nm_sim_kernel<128, uint32_t, false, true>;
  else if (small_counts)
    kern = seeded ? NM_SIM_PICK(uint16_t, true) : NM_SIM_PICK(uint16_t, false);
  else
    kern = seeded ? NM_SIM_PICK(uint32_t, true) : NM_SIM_PICK(uint32_t, false);
#undef NM_SIM_PICK
  // the reference's standard five-component model on the common launch shape: the specialised step body
  if (d.comp_word == kStdWord && a.model_kind == 0 && !gmap && !seeded && threads == 128)
    kern = small_counts ? nm_sim_kernel<128, uint16_t, false, false, kStdWord>
                        : nm_sim_kernel<128, uint32_t, false, false, kStdWord>;
  const int64_t blocks = (a.n_mice + threads - 1) / threads;
  NM_CHECK_ARG(blocks <= 2147483647LL, "nm_simulate: too many mice for one launch");
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  kern<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(d);
  e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_sim_kernel launch: %s", cudaGetErrorString(e));
  return NM_OK;
}

extern "C" int nm_sim_args_size(void) { return (int)sizeof(nm_sim_args); }

extern "C" int nm_simulate_fits(int X, int Y, int n_steps) {
  // the narrowest CTA (one warp) of nm_sim_kernel: see smem_for in nm_simulate
  if (X <= 0 || Y <= 0 || n_steps < 0) return 0;
  const size_t nT = (size_t)X * Y, th = 32;
  const size_t cnt_bytes = n_steps < 65535 ? sizeof(uint16_t) : sizeof(uint32_t);
  const size_t per_thread = (size_t)(1 + 2 * kA) + (size_t)(5 + kBcwRows) * sizeof(double);
  const size_t geo = NmSimGeom::smem_bytes(X, Y);
  if (kSimFixedBytes + th * per_thread + nT * th * cnt_bytes + nT * sizeof(double) + geo <= 227 * 1024) return 1;
  return kSimFixedBytes + (size_t)kSimThreads * per_thread + geo <= 227 * 1024 ? 2 : 0;
}

extern "C" int64_t nm_simulate_scratch_columns(int64_t n_mice) {
  return n_mice <= 0 ? 0 : ((n_mice + kSimThreads - 1) / kSimThreads) * kSimThreads;
}

extern "C" int nm_maze_movements_possible_device(const nm_maze *mz, const double *d_segments, int64_t n,
                                                 uint8_t *d_out, void *stream) {
  NM_CHECK_ARG(mz && mz->device >= 0, "nm_maze_movements_possible_device: maze not on a device");
  NM_CHECK_ARG(n >= 0 && (n == 0 || (d_segments && d_out)), "nm_maze_movements_possible_device: NULL argument");
  if (n == 0) return NM_OK;
  if (cudaSetDevice(mz->device) != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaSetDevice(%d) failed", mz->device);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  const size_t geo_bytes = NmSimGeom::smem_bytes(mz->X, mz->Y);
  NM_CHECK_ARG(geo_bytes <= 200 * 1024, "nm_maze_movements_possible_device: maze too large for shared memory");
  cudaError_t e = cudaFuncSetAttribute(nm_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo_bytes);
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  nm_segments_kernel<<<blocks, 256, geo_bytes, (cudaStream_t)stream>>>(mz->d_tiles, mz->X, mz->Y, mz->size_tile,
                                                                       d_segments, n, d_out);
  e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_segments_kernel launch: %s", cudaGetErrorString(e));
  return NM_OK;
}
