// Model-based agent: learned reward model + value-iteration planning + softmax likelihood, swept over a
// parameter grid.  EXTENSION (the reference has no model-based agent) specified in DESIGN.md
// section 2 and composed from reference primitives: the maze geometry as the transition model
// (Maze.is_movement_possible, simulation/maze.py:512-527), the `x += alpha * (target - x)` update form
// (simulation/rl.py:185-187) and the softmax of simulation/exploration_model.py:84-93.
//
// Mapping: one WARP owns one parameter set (alpha, beta, gamma).  Its tables W, V and Rhat live in shared memory; the
// transition table is staged once per CTA as rows of eight 16-bit byte offsets into W (one LDS.128 per state) and
// shared by its warps.
//   * likelihood of a step: lanes 0..A-1 evaluate Q(s, a) = Rhat[next] + gamma * V[next]; max / gather by warp
//     shuffles; log-probability in FP64 with the same operation order as the oracle;
//   * planning: a value-iteration sweep is two passes split over the 32 lanes (states u = lane, lane + 32, ...):
//     (1) W[u] = Rhat[u] + gamma * V[u] once per state -- Q(u, a) depends on (u, a) only through next(u, a), so the
//     S*A action values of a sweep are S distinct numbers, computed with exactly the oracle's two roundings;
//     (2) V[u] = max over the available actions, in action order, of W[next(u, a)].  Unavailable actions point at
//     one of sixteen sentinel slots W[-16..-1] = -inf, which the running compare-select never picks, so the pass is
//     branch-free; V can be overwritten in place because pass 2 reads only W (the sweep stays synchronous / Jacobi).
//     Which sentinel: a 64-bit LDS is served per half-warp, conflict-free when its sixteen lanes touch sixteen
//     different bank pairs (or the same address).  For every group of sixteen consecutive states and every action the
//     prologue picks a sentinel whose bank pair none of the group's available actions uses, and all unavailable
//     lanes of the group share it (a broadcast).  With a single sentinel 22 % of the wavefronts were conflicts.
// The V / Rhat recurrences use _rn intrinsics (no FMA contraction): tables are bit-identical to the numpy oracle.
//
// GROUP variant (nm_mb_kernel<true>): V and Rhat depend on (alpha, gamma) only -- beta enters the likelihood, not the
// model or the planning -- so when every aligned run of 32 consecutive parameter sets shares (alpha, gamma) (a grid
// with beta fastest; checked on the device, nm_group.cuh) a warp owns a RUN: the tables, the reward-model update and
// the value-iteration sweeps -- almost all of the work -- are done once for the 32 sets, and lane l evaluates the
// likelihood of set p0 + l from action values every lane loads as broadcasts.  Same operations per parameter set in
// the same order: results are bit-identical to the warp-per-set kernel.
#include <cuda_runtime.h>

#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "nm_common.h"

namespace {

#include "nm_group.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
#include "nm_exp.cuh"

constexpr unsigned kFull = 0xffffffffu;
constexpr int kExpRegionBytes = 1024;  // room for the 512-byte-aligned 2^(j/64) table of the likelihood exponential
constexpr int kSent = 16;  // sentinel slots in front of W (one per bank pair of a half-warp's 64-bit access)

struct MbArgs {
  const nm_step_rec *records;
  const int64_t *offsets, *online;
  const int16_t *next;  // [S * A]
  int S, A, n_sweeps;
  const double *alpha, *beta, *gamma;
  int64_t P;
  const double *v_init;  // [n_tiles] shared or NULL
  int n_tiles;
  const int32_t *tile_of_state, *state_of_tile;
  double *ll_out, *v_out, *r_out;
  const int *mode;  // device word set by nm_runs_detect_kernel: 0 warp-per-set kernel, 1 warp-per-run kernel
};

__device__ __forceinline__ double dmax2(double a, double b) { return (b > a) ? b : a; }

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// max over the eight row entries, in action order, starting from the sentinel (-inf): identical to "first available
// action initialises, the others compare-select" for every non-NaN table.
__device__ __forceinline__ double row_max(uint32_t w_base, const uint4 row) {
  const double q0 = lds_f64(w_base + (row.x & 0xffffu)), q1 = lds_f64(w_base + (row.x >> 16));
  const double q2 = lds_f64(w_base + (row.y & 0xffffu)), q3 = lds_f64(w_base + (row.y >> 16));
  const double q4 = lds_f64(w_base + (row.z & 0xffffu)), q5 = lds_f64(w_base + (row.z >> 16));
  const double q6 = lds_f64(w_base + (row.w & 0xffffu)), q7 = lds_f64(w_base + (row.w >> 16));
  double best = q0;
  best = dmax2(best, q1);
  best = dmax2(best, q2);
  best = dmax2(best, q3);
  best = dmax2(best, q4);
  best = dmax2(best, q5);
  best = dmax2(best, q6);
  best = dmax2(best, q7);
  return best;
}

template <bool GROUP>
__global__ void __launch_bounds__(512) nm_mb_kernel(const MbArgs a) {
  if (*a.mode != (GROUP ? 1 : 0)) return;  // the other mapping owns this sweep
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int S = a.S, A = a.A;
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows[u] = eight u16 byte offsets into W: 8 * (next(u, a) + kSent), or 8 * t (sentinel slot t < kSent) when
  // unavailable
  uint16_t *rows = reinterpret_cast<uint16_t *>(smem_raw);
  const size_t rows_bytes = (size_t)S * 16;
  const size_t tab_doubles = (size_t)3 * S + kSent;  // W[-kSent..S-1] | V[S] | R[S]
  double *tabs = reinterpret_cast<double *>(smem_raw + rows_bytes) + (size_t)warp * tab_doubles;
  double *W = tabs, *V = tabs + S + kSent, *R = tabs + 2 * S + kSent;
  const int groups = (S + 15) >> 4;
  for (int i = threadIdx.x; i < groups * 8; i += blockDim.x) {  // one (group of 16 states, action) per thread
    const int u0 = (i >> 3) << 4, k = i & 7;
    const int cnt = (S - u0 < 16) ? (S - u0) : 16;
    unsigned used = 0u;
    for (int j = 0; j < cnt; ++j) {
      const int nu = k < A ? (int)a.next[(u0 + j) * A + k] : -1;
      if (nu >= 0) used |= 1u << (nu & 15);  // bank pair of slot nu + kSent, relative to the table base
    }
    const int t = used == 0xffffu ? 0 : __ffs(~used) - 1;
    for (int j = 0; j < cnt; ++j) {
      const int nu = k < A ? (int)a.next[(u0 + j) * A + k] : -1;
      rows[(u0 + j) * 8 + k] = (uint16_t)(8 * (nu >= 0 ? nu + kSent : t));
    }
  }

  const int sess = blockIdx.y;
  // the warp's unit: one parameter set, or (GROUP) a run of 32 sets of which lane l evaluates set p + l
  const int64_t p = ((int64_t)blockIdx.x * warps + warp) * (GROUP ? 32 : 1);
  const bool live = p < a.P;
  const double alpha = live ? a.alpha[p] : 0.0, gamma = live ? a.gamma[p] : 0.0;
  const int64_t pl = GROUP ? p + lane : p;
  const double beta = pl < a.P ? a.beta[pl] : 0.0;
  for (int u = lane; u < S; u += 32) {
    V[u] = a.v_init ? a.v_init[a.tile_of_state[u]] : 0.0;
    R[u] = 0.0;
  }
  if (lane < kSent) W[lane] = -INFINITY;
  ExpSim::fill(smem_raw + rows_bytes + (size_t)warps * tab_doubles * sizeof(double), (int)threadIdx.x);
  __syncthreads();
  const uint32_t w_base = (uint32_t)__cvta_generic_to_shared(W);
  const uint4 *rows4 = reinterpret_cast<const uint4 *>(rows);

  const int64_t rec0 = a.offsets[sess];
  const int64_t n = a.offsets[sess + 1] - rec0, n_on = a.online[sess];
  const uint4 *recs = reinterpret_cast<const uint4 *>(a.records + rec0);
  double ll = 0.0;
  // the likelihood exponentials and the per-chunk logarithm of the fit kernels (nm_exp.cuh: ExpLL, one log per 32
  // steps of the product of the softmax denominators -- each in [1, 8], so the product stays below 2^96)
  ExpLL er;
  double acc_obs = 0.0, prod = 1.0;
  er.load(smem_raw + rows_bytes + (size_t)warps * tab_doubles * sizeof(double));
  for (int64_t i = 0; i < n; ++i) {
    const uint4 w0 = __ldg(recs + 2 * i);  // warp-uniform address: one broadcast transaction
    const double r = __hiloint2double((int)w0.y, (int)w0.x);
    const int s = (int)(w0.z & 0xffffu), s1 = (int)(w0.z >> 16);
    // ---- action values of the current state -------------------------------------------------------------
    if constexpr (!GROUP) {
    const int na = (lane < A) ? ((int)rows[s * 8 + lane] >> 3) - kSent : -1;
    const bool ok = na >= 0;
    const double q = ok ? __dadd_rn(R[na], __dmul_rn(gamma, V[na])) : 0.0;
    const unsigned avail = __ballot_sync(kFull, ok);
    const unsigned hit = __ballot_sync(kFull, ok && na == s1);
    if (i < n_on && hit) {  // observed action = first a with next(s, a) == s1 (warp-uniform branch)
      const int a_obs = __ffs(hit) - 1;
      const double z = __dmul_rn(beta, q);  // exploration_model.py:87
      double m = ok ? z : -INFINITY;
#pragma unroll
      for (int off = 4; off > 0; off >>= 1) m = dmax2(m, __shfl_xor_sync(kFull, m, off));  // lanes 0..7 hold the max
      m = __shfl_sync(kFull, m, 0);
      const double zc = __dsub_rn(z, m);  // :92
      const double e = ok ? er(zc) : 0.0;
      const int K = __popc(avail);
      double sum;
      if (K == NM_MAX_CAND) {  // numpy pairwise sum, n == 8: lanes 0..7 are all available
        double t = __dadd_rn(e, __shfl_xor_sync(kFull, e, 1));   // (e0+e1), (e2+e3), ...
        t = __dadd_rn(t, __shfl_xor_sync(kFull, t, 2));          // ((e0+e1)+(e2+e3)), ...
        sum = __dadd_rn(__shfl_sync(kFull, t, 0), __shfl_sync(kFull, t, 4));
      } else {  // n < 8: sequential from 0.0 over the available actions in order
        sum = 0.0;
        for (int k = 0; k < A; ++k) {
          const double ek = __shfl_sync(kFull, e, k);
          if ((avail >> k) & 1u) sum = __dadd_rn(sum, ek);
        }
      }
      const double zobs = __shfl_sync(kFull, zc, a_obs);
      acc_obs = __dadd_rn(acc_obs, zobs);
      prod = __dmul_rn(prod, sum);
    }
    } else {
    // every lane loads the (at most eight) action values as broadcasts and evaluates its own beta; the action set
    // and the observed action are warp-uniform
    const uint4 rw = rows4[s];
    const uint32_t offs[8] = {rw.x & 0xffffu, rw.x >> 16, rw.y & 0xffffu, rw.y >> 16,
                              rw.z & 0xffffu, rw.z >> 16, rw.w & 0xffffu, rw.w >> 16};
    unsigned avail = 0u;
    int a_obs = -1;
    double q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int na = (int)(offs[k] >> 3) - kSent;
      q[k] = 0.0;
      if (na >= 0) {
        avail |= 1u << k;
        if (a_obs < 0 && na == s1) a_obs = k;
        q[k] = __dadd_rn(R[na], __dmul_rn(gamma, V[na]));
      }
    }
    if (i < n_on && a_obs >= 0) {
      double z[8], m = -INFINITY;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        z[k] = __dmul_rn(beta, q[k]);  // exploration_model.py:87
        if ((avail >> k) & 1u) m = dmax2(m, z[k]);
      }
      double e[8], zobs = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double zc = __dsub_rn(z[k], m);  // :92
        e[k] = ((avail >> k) & 1u) ? er(zc) : 0.0;
        if (k == a_obs) zobs = zc;
      }
      double sum;
      if (avail == 0xffu) {  // numpy pairwise sum, n == 8
        sum = __dadd_rn(__dadd_rn(__dadd_rn(e[0], e[1]), __dadd_rn(e[2], e[3])),
                        __dadd_rn(__dadd_rn(e[4], e[5]), __dadd_rn(e[6], e[7])));
      } else {  // n < 8: sequential from 0.0; an unavailable action's +0.0 leaves the partial sum unchanged
        sum = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum = __dadd_rn(sum, e[k]);
      }
      acc_obs = __dadd_rn(acc_obs, zobs);
      prod = __dmul_rn(prod, sum);
    }
    }
    if ((i & 31) == 31 || i == n - 1) {  // fold the chunk:  ll += sum(zobs) - log(prod of the softmax denominators)
      ll = __dadd_rn(ll, __dsub_rn(acc_obs, log_chunk_product(prod)));
      acc_obs = 0.0;
      prod = 1.0;
    }
    // ---- world model: Rhat[s1] += alpha * (r - Rhat[s1]) ---------------------------------------------------
    if (lane == 0) {
      const double rs = R[s1];
      R[s1] = __dadd_rn(rs, __dmul_rn(alpha, __dsub_rn(r, rs)));
    }
    __syncwarp();
    // ---- planning: n_sweeps synchronous value-iteration sweeps --------------------------------------------
    for (int sw = 0; sw < a.n_sweeps; ++sw) {
      for (int u = lane; u < S; u += 32) W[u + kSent] = __dadd_rn(R[u], __dmul_rn(gamma, V[u]));
      __syncwarp();
      int u = lane;
      for (; u + 32 < S; u += 64) {  // two states per trip: sixteen independent loads in flight
        const uint4 ra = rows4[u], rb = rows4[u + 32];
        const double ba = row_max(w_base, ra), bb = row_max(w_base, rb);
        if (ba != -INFINITY) V[u] = ba;  // no available action (every entry a sentinel): V[u] stays
        if (bb != -INFINITY) V[u + 32] = bb;
      }
      if (u < S) {
        const uint4 ra = rows4[u];
        const double ba = row_max(w_base, ra);
        if (ba != -INFINITY) V[u] = ba;
      }
      __syncwarp();
    }
  }
  if (a.ll_out && (GROUP ? pl < a.P : (live && lane == 0))) a.ll_out[(int64_t)sess * a.P + pl] = ll;
  if (live && (a.v_out || a.r_out)) {
    const int nT = a.n_tiles;
    const int copies = GROUP ? (int)(a.P - p < 32 ? a.P - p : 32) : 1;  // every set of a run receives the run's tables
    for (int c = 0; c < copies; ++c)
      for (int tile = lane; tile < nT; tile += 32) {
        const int st = a.state_of_tile[tile];
        const int64_t o = ((int64_t)sess * a.P + p + c) * nT + tile;
        if (a.v_out) a.v_out[o] = st >= 0 ? V[st] : (a.v_init ? a.v_init[tile] : 0.0);
        if (a.r_out) a.r_out[o] = st >= 0 ? R[st] : 0.0;
      }
  }
}

}  // namespace

extern "C" int nm_modelbased_sweep(const nm_maze *mz, const nm_streams *st, const int16_t *d_next, int n_actions,
                                   int n_sweeps, const double *d_alpha, const double *d_beta, const double *d_gamma,
                                   int64_t P, const double *d_v_init, double *d_ll_out, double *d_v_out,
                                   double *d_r_out, void *stream) {
  NM_CHECK_ARG(mz && st && d_next, "nm_modelbased_sweep: NULL maze / streams / transition table");
  NM_CHECK_ARG(mz->device >= 0 && mz->device == st->device, "nm_modelbased_sweep: maze and streams must live on the "
               "same CUDA device");
  NM_CHECK_ARG(st->num_states == mz->num_states, "nm_modelbased_sweep: streams were built for another maze");
  NM_CHECK_ARG(n_actions >= 1 && n_actions <= NM_MAX_CAND, "nm_modelbased_sweep: need 1..%d actions", NM_MAX_CAND);
  NM_CHECK_ARG(n_sweeps >= 0 && n_sweeps <= 1024, "nm_modelbased_sweep: n_sweeps out of range");
  NM_CHECK_ARG(P >= 0 && (P == 0 || (d_alpha && d_beta && d_gamma && d_ll_out)), "nm_modelbased_sweep: NULL array");
  NM_CHECK_ARG(st->S <= 65535, "nm_modelbased_sweep: at most 65535 sessions per call");
  if (P == 0) return NM_OK;
  if (cudaSetDevice(mz->device) != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaSetDevice(%d) failed", mz->device);
  const int S = mz->num_states;
  NM_CHECK_ARG(S <= 8192 - kSent, "nm_modelbased_sweep: at most %d states (16-bit byte offsets into the value table)",
               8192 - kSent);
  const size_t rows_bytes = (size_t)S * 16;
  const size_t per_warp = ((size_t)3 * S + kSent) * sizeof(double);
  // CTAs per SM x warps per CTA: as many resident warps as 227 KB of shared memory (1 KB reserved per CTA) allow
  int warps = 0, best_total = 0;
  for (int ctas = 1; ctas <= 8; ++ctas) {
    const long avail = (long)(227 * 1024) / ctas - 1024 - (long)rows_bytes - (long)kExpRegionBytes;
    if (avail < (long)per_warp) break;
    int w = (int)(avail / (long)per_warp);
    if (w > 16) w = 16;
    int total = w * ctas;
    if (total > 64) total = 64;
    if (total >= best_total) {
      best_total = total;
      warps = w;
    }
  }
  NM_CHECK_ARG(warps >= 1, "nm_modelbased_sweep: a maze with %d states does not fit in shared memory", S);
  const size_t smem = rows_bytes + per_warp * warps + kExpRegionBytes;
  cudaError_t e = cudaFuncSetAttribute(nm_mb_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(nm_mb_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  MbArgs a;
  a.records = st->d_records;
  a.offsets = st->d_offsets;
  a.online = st->d_online;
  a.next = d_next;
  a.S = S;
  a.A = n_actions;
  a.n_sweeps = n_sweeps;
  a.alpha = d_alpha;
  a.beta = d_beta;
  a.gamma = d_gamma;
  a.P = P;
  a.v_init = d_v_init;
  a.n_tiles = mz->X * mz->Y;
  a.tile_of_state = mz->d_tile_of_state;
  a.state_of_tile = mz->d_state_of_tile;
  a.ll_out = d_ll_out;
  a.v_out = d_v_out;
  a.r_out = d_r_out;
  // Parameter sets that share (alpha, gamma) share V and Rhat: the device decides whether every aligned run of 32
  // consecutive sets does, and exactly one of the two launches does the work (NM_MB_GROUP=0: never group).
  cudaStream_t cs = (cudaStream_t)stream;
  std::lock_guard<std::mutex> launch_lock(runs_launch_mutex());  // decision + launches: one sequence on the stream
  int *mode = nullptr;
  const int rc = runs_mode_word(stream, &mode);
  if (rc != NM_OK) return rc;
  a.mode = mode;
  const char *env = getenv("NM_MB_GROUP");
  const bool group = !(env && atoi(env) == 0);
  runs_decide(group, d_alpha, d_gamma, P, mode, cs);
  const int64_t blocks = (P + warps - 1) / warps;
  NM_CHECK_ARG(blocks <= 2147483647LL, "nm_modelbased_sweep: too many parameter sets for one launch");
  if (group) {
    // a run per warp: spread few runs over the SMs (the sweep is shared-memory bound: fewer warps per SM are faster)
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, mz->device);
    const int64_t runs = (P + 31) / 32;
    int64_t per_cta = (runs * st->S + sms - 1) / sms;
    per_cta = (per_cta + st->S - 1) / st->S;
    const int gw = (int)(per_cta < 1 ? 1 : per_cta > warps ? warps : per_cta);
    nm_mb_kernel<true><<<dim3((unsigned)((runs + gw - 1) / gw), (unsigned)st->S, 1), gw * 32,
                         rows_bytes + per_warp * gw + kExpRegionBytes, cs>>>(a);
  }
  nm_mb_kernel<false><<<dim3((unsigned)blocks, (unsigned)st->S, 1), warps * 32, smem, cs>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_mb_kernel launch: %s", cudaGetErrorString(e));
  return NM_OK;
}

extern "C" int nm_modelbased_sweep_mode(void *stream, int *mode_out) {
  return runs_mode_read(stream, mode_out, "nm_modelbased_sweep_mode");
}
