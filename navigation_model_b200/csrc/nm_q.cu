// Tabular Q-learning agent with an explicit Q[s, a] table, swept over a parameter grid.  EXTENSION (the reference
// has state-value learners only) specified in DESIGN.md section 2 and composed from reference
// primitives: the maze geometry as the action model (Maze.is_movement_possible, simulation/maze.py:512-527), the
// max-backup update form of PositiveConditioning (simulation/rl.py:184-187) and the softmax of
// simulation/exploration_model.py:84-93.
//
// Mapping: eight lanes (an "octet") own one parameter set (alpha, beta, gamma); lane k of the octet owns action k,
// i.e. COLUMN k of the set's table Q[state][8] in shared memory -- a lane only ever reads and writes its own
// column, so the step loop needs no barrier at all.  A warp carries four parameter sets.  All sets walk the same
// stream, so a step reads one 64-byte row per octet; the table stride is kept at 64 (mod 128) bytes so that the two
// octets of a half-warp fall into different halves of the 32 banks (conflict-free 64-bit accesses).
//   * action set of a state = actions whose centre-to-centre tile move is possible (next[s, a] >= 0);
//   * observed action = first a with next[s, a] == s1; a step without one (a jump) is skipped entirely;
//   * log-probability: beta * Q[s, avail], subtract the max, exps, numpy-order sum (sequential for fewer than eight
//     available actions, pairwise for eight), evaluated BEFORE the update; the log is taken once per 32 records on
//     the product of the sums (each in [1, 8]);
//   * update: Q[s, a] += alpha * (r + gamma * max_{a' avail at s1} Q[s1, a'] - Q[s, a]), bootstrap 0 when s1 has no
//     available action.  _rn intrinsics only (no FMA contraction): tables are bit-identical to the numpy oracle.
// Replay records (past the session's online count) update the table without contributing to the likelihood.
#include <cuda_runtime.h>

#include "nm_common.h"

namespace {

constexpr unsigned kFull = 0xffffffffu;

#include "nm_exp.cuh"

struct QArgs {
  const nm_step_rec *records;
  const int64_t *offsets, *online;
  const int16_t *next;  // [S * A]
  int S, A;
  uint32_t stride;  // bytes between the tables of consecutive parameter sets
  const double *alpha, *beta, *gamma;
  int64_t P;
  const double *q_init;  // [n_tiles * A] shared or NULL
  int n_tiles;
  const int32_t *tile_of_state, *state_of_tile;
  double *ll_out, *q_out;
};

__device__ __forceinline__ double dmax2(double a, double b) { return (b > a) ? b : a; }

// max over the eight lanes of an octet, for two independent values at once (the two butterfly chains are interleaved
// by hand: each level issues its four shuffles before the first compare); every lane receives both maxima
__device__ __forceinline__ void octet_max2(double &a, double &b) {
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) {
    const double a2 = __shfl_xor_sync(kFull, a, off), b2 = __shfl_xor_sync(kFull, b, off);
    a = dmax2(a, a2);
    b = dmax2(b, b2);
  }
}

__global__ void __launch_bounds__(512) nm_q_kernel(const QArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = a.S, A = a.A;
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int oct = lane >> 3, k = lane & 7;
  double *exptab = reinterpret_cast<double *>(smem_raw);                 // 128 B
  int16_t *nxt8 = reinterpret_cast<int16_t *>(smem_raw + 128);           // [S][8], -1 = unavailable
  const size_t nxt_bytes = ((size_t)S * 16 + 127) & ~(size_t)127;
  double *Q = reinterpret_cast<double *>(smem_raw + 128 + nxt_bytes + (size_t)(warp * 4 + oct) * a.stride);
  if (threadIdx.x < 16) exptab[threadIdx.x] = c_exp2tab[threadIdx.x];
  for (int i = threadIdx.x; i < S * 8; i += blockDim.x) {
    const int u = i >> 3, kk = i & 7;
    nxt8[i] = kk < A ? a.next[u * A + kk] : (int16_t)-1;
  }
  const int sess = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * (warps * 4) + warp * 4 + oct;
  const bool live = p < a.P;
  const double alpha = live ? a.alpha[p] : 0.0, beta = live ? a.beta[p] : 0.0, gamma = live ? a.gamma[p] : 0.0;
  for (int u = 0; u < S; ++u)
    Q[u * 8 + k] = (a.q_init && k < A) ? a.q_init[(int64_t)a.tile_of_state[u] * A + k] : 0.0;
  __syncthreads();
  ExpRegs er;
  er.load((uint32_t)__cvta_generic_to_shared(exptab));

  const int64_t rec0 = a.offsets[sess];
  const int64_t n = a.offsets[sess + 1] - rec0, n_on = a.online[sess];
  const uint4 *recs = reinterpret_cast<const uint4 *>(a.records + rec0);
  double ll = 0.0, acc_obs = 0.0, prod = 1.0;
  // Everything a step needs besides Q -- the record, the action sets of s and s1, the observed action -- does not
  // depend on the parameters: it is fetched and decoded one (action sets) and two (record) steps ahead, so that the
  // only latency left in a step is the Q-dependent chain.
  struct Step {
    int s, s1;
    double r;
    unsigned avail, hit, avail1;
    bool ok, ok1;
  };
  auto fetch = [&](int64_t i) -> uint4 { return i < n ? __ldg(recs + 2 * i) : make_uint4(0u, 0u, 0u, 0u); };
  auto decode = [&](const uint4 w0) -> Step {
    Step d;
    d.s = (int)(w0.z & 0xffffu);
    d.s1 = (int)(w0.z >> 16);
    d.r = __hiloint2double((int)w0.y, (int)w0.x);
    const int nk = nxt8[d.s * 8 + k], n1 = nxt8[d.s1 * 8 + k];
    d.ok = nk >= 0;
    d.ok1 = n1 >= 0;
    d.avail = __ballot_sync(kFull, d.ok) & 0xffu;  // the four octets agree: octet 0's answer
    d.hit = __ballot_sync(kFull, d.ok && nk == d.s1) & 0xffu;
    d.avail1 = __ballot_sync(kFull, d.ok1) & 0xffu;
    return d;
  };
  uint4 w_next = fetch(1);
  Step cur = decode(fetch(0));
  const int n32 = (int)n, n_on32 = (int)(n_on < n ? n_on : n);  // n < 2^31 (checked on the host)
  for (int i = 0; i < n32; ++i) {
    const Step st = cur;
    cur = decode(w_next);  // step i + 1 (an all-zero record past the end: state 0, harmless)
    w_next = fetch(i + 2);
    if (st.hit) {  // warp-uniform; a step without an observed action is skipped entirely
      const int a_obs = __ffs(st.hit) - 1;
      double *row = Q + st.s * 8 + k;
      const double q = *row, q1 = Q[st.s1 * 8 + k];
      const double z = __dmul_rn(beta, q);  // exploration_model.py:87
      double boot = st.ok1 ? q1 : -INFINITY, m = st.ok ? z : -INFINITY;
      octet_max2(boot, m);
      if (!st.avail1) boot = 0.0;
      if (i < n_on32) {
        const double zc = __dsub_rn(z, m);  // :92
        const double e = st.ok ? er.exp_nonpos(dmax2(zc, -800.0)) : 0.0;  // :93
        // every lane gathers the octet's eight terms (sixteen independent shuffles), then adds them in numpy's order:
        // pairwise for n == 8, else sequentially from 0.0 over the available actions -- an unavailable action's term
        // is +0.0, and x + 0.0 == x bit for bit for the non-negative partial sums, so nothing needs to be skipped
        double ev[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) ev[kk] = __shfl_sync(kFull, e, kk, 8);
        double sum;
        if (st.avail == 0xffu) {
          sum = __dadd_rn(__dadd_rn(__dadd_rn(ev[0], ev[1]), __dadd_rn(ev[2], ev[3])),
                          __dadd_rn(__dadd_rn(ev[4], ev[5]), __dadd_rn(ev[6], ev[7])));
        } else {
          sum = 0.0;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) sum = __dadd_rn(sum, ev[kk]);
        }
        acc_obs = __dadd_rn(acc_obs, __shfl_sync(kFull, zc, a_obs, 8));
        prod = __dmul_rn(prod, sum);
      }
      if (k == a_obs) {
        const double rpe = __dsub_rn(__dadd_rn(st.r, __dmul_rn(gamma, boot)), q);  // update form of rl.py:185
        *row = __dadd_rn(q, __dmul_rn(alpha, rpe));                                 // rl.py:186-187
      }
    }
    if ((i & 31) == 31 || i + 1 == n32) {  // fold: at most 32 sums, each in [1, 8] -> prod <= 2^96
      ll = __dadd_rn(ll, __dsub_rn(acc_obs, log(prod)));
      acc_obs = 0.0;
      prod = 1.0;
    }
  }
  if (live && k == 0 && a.ll_out) a.ll_out[(int64_t)sess * a.P + p] = ll;
  if (live && a.q_out && k < A) {
    const int nT = a.n_tiles;
    double *dst = a.q_out + ((int64_t)sess * a.P + p) * nT * A;
    for (int tile = 0; tile < nT; ++tile) {
      const int st = a.state_of_tile[tile];
      dst[(int64_t)tile * A + k] = st >= 0 ? Q[st * 8 + k] : (a.q_init ? a.q_init[(int64_t)tile * A + k] : 0.0);
    }
  }
}

}  // namespace

extern "C" int nm_qtable_sweep(const nm_maze *mz, const nm_streams *st, const int16_t *d_next, int n_actions,
                               const double *d_alpha, const double *d_beta, const double *d_gamma, int64_t P,
                               const double *d_q_init, double *d_ll_out, double *d_q_out, void *stream) {
  NM_CHECK_ARG(mz && st && d_next, "nm_qtable_sweep: NULL maze / streams / transition table");
  NM_CHECK_ARG(mz->device >= 0 && mz->device == st->device, "nm_qtable_sweep: maze and streams must live on the same "
               "CUDA device");
  NM_CHECK_ARG(st->num_states == mz->num_states, "nm_qtable_sweep: streams were built for another maze");
  NM_CHECK_ARG(n_actions >= 1 && n_actions <= NM_MAX_CAND, "nm_qtable_sweep: need 1..%d actions", NM_MAX_CAND);
  NM_CHECK_ARG(P >= 0 && (P == 0 || (d_alpha && d_beta && d_gamma && d_ll_out)), "nm_qtable_sweep: NULL array");
  NM_CHECK_ARG(st->S <= 65535, "nm_qtable_sweep: at most 65535 sessions per call");
  for (int i = 0; i < st->S; ++i)
    NM_CHECK_ARG(st->offsets[i + 1] - st->offsets[i] < 2147483647LL, "nm_qtable_sweep: session %d is too long", i);
  if (P == 0) return NM_OK;
  if (cudaSetDevice(mz->device) != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaSetDevice(%d) failed", mz->device);
  const int S = mz->num_states;
  NM_CHECK_ARG(S <= 32767, "nm_qtable_sweep: at most 32767 states (16-bit successor ids)");
  size_t stride = (size_t)S * 64;
  if ((stride & 127) == 0) stride += 64;  // 64 (mod 128): the two octets of a half-warp use different bank halves
  const size_t fixed = 128 + (((size_t)S * 16 + 127) & ~(size_t)127);
  const size_t per_warp = 4 * stride;
  // CTAs per SM x warps per CTA: as many resident warps as 227 KB of shared memory (1 KB reserved per CTA) allow
  int warps = 0, best_total = 0;
  for (int ctas = 1; ctas <= 8; ++ctas) {
    const long avail = (long)(227 * 1024) / ctas - 1024 - (long)fixed;
    if (avail < (long)per_warp) break;
    int w = (int)(avail / (long)per_warp);
    if (w > 16) w = 16;
    int total = w * ctas;
    if (total > 64) total = 64;
    if (total > best_total) {
      best_total = total;
      warps = w;
    }
  }
  NM_CHECK_ARG(warps >= 1, "nm_qtable_sweep: the Q-tables of four parameter sets on a maze with %d states do not fit "
               "in shared memory", S);
  const size_t smem = fixed + per_warp * warps;
  cudaError_t e = cudaFuncSetAttribute(nm_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  QArgs a;
  a.records = st->d_records;
  a.offsets = st->d_offsets;
  a.online = st->d_online;
  a.next = d_next;
  a.S = S;
  a.A = n_actions;
  a.stride = (uint32_t)stride;
  a.alpha = d_alpha;
  a.beta = d_beta;
  a.gamma = d_gamma;
  a.P = P;
  a.q_init = d_q_init;
  a.n_tiles = mz->X * mz->Y;
  a.tile_of_state = mz->d_tile_of_state;
  a.state_of_tile = mz->d_state_of_tile;
  a.ll_out = d_ll_out;
  a.q_out = d_q_out;
  const int64_t per_cta = (int64_t)warps * 4;
  const int64_t blocks = (P + per_cta - 1) / per_cta;
  NM_CHECK_ARG(blocks <= 2147483647LL, "nm_qtable_sweep: too many parameter sets for one launch");
  nm_q_kernel<<<dim3((unsigned)blocks, (unsigned)st->S, 1), warps * 32, smem, (cudaStream_t)stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_q_kernel launch: %s", cudaGetErrorString(e));
  return NM_OK;
}
