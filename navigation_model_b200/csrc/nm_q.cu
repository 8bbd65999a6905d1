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
//
// Second mapping (nm_qgroup_kernel), taken when every aligned run of 32 consecutive parameter sets shares (alpha,
// gamma) -- a Cartesian grid with beta fastest; the device checks the arrays itself (nm_runs_detect_kernel): the
// table depends on (alpha, gamma) only, so a WARP keeps ONE table Q[state * 8 + action][warp] for its 32 sets and
// lane l evaluates the softmax of set p0 + l.  Table loads are broadcasts, every lane stores the same updated value.
// A producer warp TMA-stages the raw records (cp.async.bulk + mbarrier), derives the action lists of s and s1 and
// the observed action ONCE per CTA and hands 80-byte decoded records to the consumer warps through mbarriers.  Per
// parameter set the operations are those of the octet kernel, in the same order: results are bit-identical.
#include <cuda_runtime.h>

#include <type_traits>

#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "nm_common.h"

namespace {

constexpr unsigned kFull = 0xffffffffu;

#include "nm_exp.cuh"
#include "nm_group.cuh"
#include "nm_ring.cuh"
#include "nm_step_math.cuh"

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
  const int *mode;  // device word set by nm_runs_detect_kernel: 0 octet kernel, 1 group kernel; NULL: always run
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
  if (a.mode && *a.mode != 0) return;  // the group kernel owns this sweep
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = a.S, A = a.A;
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int oct = lane >> 3, k = lane & 7;
  unsigned char *exptab = smem_raw;                                       // 1024 B (ExpLL table region)
  int16_t *nxt8 = reinterpret_cast<int16_t *>(smem_raw + 1024);          // [S][8], -1 = unavailable
  const size_t nxt_bytes = ((size_t)S * 16 + 127) & ~(size_t)127;
  double *Q = reinterpret_cast<double *>(smem_raw + 1024 + nxt_bytes + (size_t)(warp * 4 + oct) * a.stride);
  ExpLL::fill(exptab, (int)threadIdx.x);
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
  ExpLL er;
  er.load(exptab);

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
        const double zc = fma(beta, q, -m);  // :92, fused like the per-warp-table kernel (bit-identical results)
        const double e = st.ok ? er(zc) : 0.0;  // :93
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
      ll = __dadd_rn(ll, __dsub_rn(acc_obs, log_chunk_product(prod)));
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

// =================================== one table per warp =======================================================
constexpr int kQChunk = 32;     // records per chunk = one per producer lane
constexpr int kQRawStages = 2;  // raw (TMA landing) ring
constexpr int kQDecStages = 4;  // decoded ring
constexpr int kQDecWords = 5;   // uint4 per decoded record (80 bytes)
constexpr int kQBarBytes = 128;
constexpr int kQRawBytes = kQRawStages * kQChunk * 32;
constexpr int kQDecBytes = kQDecStages * kQChunk * kQDecWords * 16;
constexpr int kQTabBytes = 1024;  // room for the 512-byte-aligned 2^(j/64) table of ExpLL (nm_exp.cuh)
constexpr int kQFixedBytes = kQBarBytes + kQRawBytes + kQDecBytes + kQTabBytes;
constexpr uint32_t kQOnline = 1u << 8, kQHit = 1u << 9, kQBoot = 1u << 10;

// decoded record (byte offsets = table row * row_bytes; row = state * 8 + action; the row past the table holds -inf):
//   q0 = {r.lo, r.hi, K1 | kQOnline | kQHit | kQBoot, offset of the updated entry (s, observed action)}
//   q1, q2 = the K1 available actions of s, in action order (softmax candidates)
//   q3, q4 = the eight actions of s1 (unavailable ones point at the -inf row)
struct QRings {
  uint64_t *full_raw, *full_dec, *empty_dec;
  uint4 *raw, *dec;
};

__device__ __forceinline__ void q_producer_loop(const QRings &rg, const nm_step_rec *src, int64_t n, int64_t n_on,
                                                int nchunks, const int16_t *nxt8, uint32_t row_bytes, uint32_t sentinel,
                                                int lane) {
  auto issue = [&](int c) {
    const int slot = c % kQRawStages;
    const int64_t first = (int64_t)c * kQChunk;
    const uint32_t bytes = (uint32_t)((n - first < kQChunk ? n - first : kQChunk) * 32);
    mbar_expect_tx(&rg.full_raw[slot], bytes);
    tma_bulk_g2s(rg.raw + slot * (kQChunk * 2), src + first, bytes, &rg.full_raw[slot]);
  };
  if (lane == 0)
    for (int c = 0; c < kQRawStages && c < nchunks; ++c) issue(c);
  for (int c = 0; c < nchunks; ++c) {
    const int rslot = c % kQRawStages, dslot = c % kQDecStages;
    if (c >= kQDecStages) mbar_wait_relaxed(&rg.empty_dec[dslot], (uint32_t)(((c / kQDecStages) - 1) & 1));
    mbar_wait_relaxed(&rg.full_raw[rslot], (uint32_t)((c / kQRawStages) & 1));
    const int64_t first = (int64_t)c * kQChunk;
    const int cnt = (int)(n - first < kQChunk ? n - first : kQChunk);
    const int src_lane = lane < cnt ? lane : cnt - 1;
    const uint4 w0 = rg.raw[(rslot * kQChunk + src_lane) * 2];
    const int s = (int)(w0.z & 0xffffu), s1 = (int)(w0.z >> 16);
    const uint4 n_s = *reinterpret_cast<const uint4 *>(nxt8 + s * 8), n_s1 = *reinterpret_cast<const uint4 *>(nxt8 + s1 * 8);
    const uint32_t ws[4] = {n_s.x, n_s.y, n_s.z, n_s.w}, ws1[4] = {n_s1.x, n_s1.y, n_s1.z, n_s1.w};
    uint32_t o1[8], o2[8];
    uint32_t K1 = 0, upd = 0, flags = (first + lane < n_on) ? kQOnline : 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) o1[k] = sentinel;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int nk = (int)(int16_t)((k & 1) ? (ws[k >> 1] >> 16) : (ws[k >> 1] & 0xffffu));
      const int n1 = (int)(int16_t)((k & 1) ? (ws1[k >> 1] >> 16) : (ws1[k >> 1] & 0xffffu));
      const uint32_t off = (uint32_t)(s * 8 + k) * row_bytes;
      if (nk >= 0) {
        if (nk == s1 && !(flags & kQHit)) {  // observed action = first a with next[s, a] == s1
          flags |= kQHit;
          upd = off;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if ((uint32_t)j == K1) o1[j] = off;
        ++K1;
      }
      o2[k] = n1 >= 0 ? (uint32_t)(s1 * 8 + k) * row_bytes : sentinel;
      if (n1 >= 0) flags |= kQBoot;
    }
    if (lane < cnt) {
      uint4 *d = rg.dec + (dslot * kQChunk + lane) * kQDecWords;
      d[0] = make_uint4(w0.x, w0.y, K1 | flags, upd);
      d[1] = make_uint4(o1[0], o1[1], o1[2], o1[3]);
      d[2] = make_uint4(o1[4], o1[5], o1[6], o1[7]);
      d[3] = make_uint4(o2[0], o2[1], o2[2], o2[3]);
      d[4] = make_uint4(o2[4], o2[5], o2[6], o2[7]);
    }
    mbar_arrive(&rg.full_dec[dslot]);  // every lane: releases its own stores (count = 32)
    __syncwarp();                      // all lanes have read the raw slot before it is refilled
    if (lane == 0 && c + kQRawStages < nchunks) issue(c + kQRawStages);
  }
}

// Where a table lives (see nm_fit_kernels.cuh): any pointer, or a shared-window address pinned in a register.
struct QGenericCol {
  char *qb;
  __device__ __forceinline__ double ld(uint32_t off) const { return *reinterpret_cast<const double *>(qb + off); }
  __device__ __forceinline__ void st(uint32_t off, double v) const { *reinterpret_cast<double *>(qb + off) = v; }
};
struct QSharedCol {
  uint32_t base;
  __device__ __forceinline__ explicit QSharedCol(const void *col) {
    base = (uint32_t)__cvta_generic_to_shared(col);
    asm volatile("mov.u32 %0, %0;" : "+r"(base));
  }
  __device__ __forceinline__ double ld(uint32_t off) const {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + off) : "memory");
    return v;
  }
  __device__ __forceinline__ void st(uint32_t off, double v) const {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + off), "d"(v) : "memory");
  }
};

struct QAcc {
  double acc_obs, prod;
  ExpLL er;  // likelihood exponentials: nm_exp.cuh (1e-9 bar on the session's log-likelihood, tables never see them)
};

// SIGN: the caller has established once, for the whole warp, that beta >= 0 (the per-step test and its predicated twin
// disappear from the step body; same results).
template <int K1, class Col, bool SIGN = false>
__device__ __forceinline__ void qgroup_step_k(const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3,
                                              const uint4 q4, const Col qb, const double alpha, const double beta,
                                              const double gamma, QAcc &acc) {
  const double r = __hiloint2double((int)q0.y, (int)q0.x);
  const double q = qb.ld(q0.w);  // Q[s, observed action]
  const uint32_t o2[8] = {q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, q4.z, q4.w};
  double w[NM_MAX_CAND];
#pragma unroll
  for (int k = 0; k < 8; ++k) w[k] = qb.ld(o2[k]);
  double boot = extremum<NM_ALG_POSITIVE, 8>(w);  // unavailable actions read -inf
  if (!(q0.z & kQBoot)) boot = 0.0;
  if (q0.z & kQOnline) {
    const uint32_t o1[8] = {q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
    double v[NM_MAX_CAND], e[NM_MAX_CAND];
#pragma unroll
    for (int k = 0; k < K1; ++k) v[k] = qb.ld(o1[k]);
    // beta >= 0: max_k fl(beta * v_k) == fl(beta * max_k v_k), rounding is monotonic   (exploration_model.py:87, 92)
    double m;
    if (SIGN || beta >= 0.0) {
      m = __dmul_rn(beta, extremum<NM_ALG_POSITIVE, K1>(v));
    } else {
      double z[NM_MAX_CAND];
#pragma unroll
      for (int k = 0; k < K1; ++k) z[k] = __dmul_rn(beta, v[k]);
      m = extremum<NM_ALG_POSITIVE, K1>(z);
    }
    const double neg_m = -m;
#pragma unroll
    for (int k = 0; k < K1; ++k) e[k] = acc.er(fma(beta, v[k], neg_m));  // :92-93, one fused argument per action
    acc.acc_obs = __dadd_rn(acc.acc_obs, fma(beta, q, neg_m));
    acc.prod = __dmul_rn(acc.prod, np_sum_fixed<K1>(e));
  }
  const double rpe = __dsub_rn(__dadd_rn(r, __dmul_rn(gamma, boot)), q);  // update form of rl.py:185
  qb.st(q0.w, __dadd_rn(q, __dmul_rn(alpha, rpe)));                        // rl.py:186-187
}

template <class Col, bool SIGN = false>
__device__ __forceinline__ void qgroup_step(const uint4 *__restrict__ rec, const Col qb, const double alpha, const double beta,
                                            const double gamma, QAcc &acc) {
  const uint4 q0 = rec[0];
  if (!(q0.z & kQHit)) return;  // no observed action: the step is skipped entirely (warp-uniform)
  const uint4 q1 = rec[1], q2 = rec[2], q3 = rec[3], q4 = rec[4];
#define NM_CASE(KK) \
  case KK: qgroup_step_k<KK, Col, SIGN>(q0, q1, q2, q3, q4, qb, alpha, beta, gamma, acc); break;
  switch (q0.z & 0xffu) {  // K1, warp-uniform (>= 1 on a hit)
    NM_CASE(1) NM_CASE(2) NM_CASE(3) NM_CASE(4) NM_CASE(5) NM_CASE(6) NM_CASE(7) NM_CASE(8)
    default: break;
  }
#undef NM_CASE
}

// Runs only when *a.mode == 1.  blockDim.x = 32 * (nw + 1): nw consumer warps, each with its own table
// Q[state * 8 + action][warp] (+ one row of -inf), and the producer warp.
__global__ void __launch_bounds__(960, 1) nm_qgroup_kernel(const QArgs a) {
  if (*a.mode != 1) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = a.S, A = a.A;
  QRings rg;
  rg.full_raw = reinterpret_cast<uint64_t *>(smem_raw);
  rg.full_dec = rg.full_raw + kQRawStages;
  rg.empty_dec = rg.full_dec + kQDecStages;
  rg.raw = reinterpret_cast<uint4 *>(smem_raw + kQBarBytes);
  rg.dec = reinterpret_cast<uint4 *>(smem_raw + kQBarBytes + kQRawBytes);
  unsigned char *exptab = smem_raw + kQBarBytes + kQRawBytes + kQDecBytes;
  int16_t *nxt8 = reinterpret_cast<int16_t *>(smem_raw + kQFixedBytes);
  const size_t nxt_bytes = ((size_t)S * 16 + 127) & ~(size_t)127;
  double *Q = reinterpret_cast<double *>(smem_raw + kQFixedBytes + nxt_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nw = (blockDim.x >> 5) - 1;
  ExpLL::fill(exptab, tid);
  for (int i = tid; i < S * 8; i += blockDim.x) {
    const int u = i >> 3, kk = i & 7;
    nxt8[i] = kk < A ? a.next[u * A + kk] : (int16_t)-1;
  }
  const int sess = blockIdx.y;
  const int64_t rec0 = a.offsets[sess];
  const int64_t n = a.offsets[sess + 1] - rec0, n_on = a.online[sess];
  const int nchunks = (int)((n + kQChunk - 1) / kQChunk);
  if (tid == 0) {
    for (int i = 0; i < kQRawStages; ++i) mbar_init(&rg.full_raw[i], 1);
    for (int i = 0; i < kQDecStages; ++i) {
      mbar_init(&rg.full_dec[i], 32);
      mbar_init(&rg.empty_dec[i], nw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int rows = S * 8;
  if (warp >= nw) {
    q_producer_loop(rg, a.records + rec0, n, n_on, nchunks, nxt8, (uint32_t)nw * 8u, (uint32_t)rows * (uint32_t)nw * 8u,
                    lane);
  } else {
    const int64_t p0 = ((int64_t)blockIdx.x * nw + warp) * 32;
    const int64_t p = p0 + lane;
    const bool alive = p0 < a.P;  // warp-uniform
    const double alpha = alive ? a.alpha[p0] : 0.0, gamma = alive ? a.gamma[p0] : 0.0;
    const double beta = p < a.P ? a.beta[p] : 0.0;
    double *Qcol = Q + warp;
    for (int e = lane; e < rows; e += 32) {
      const int u = e >> 3, kk = e & 7;
      Qcol[e * nw] = (a.q_init && kk < A) ? a.q_init[(int64_t)a.tile_of_state[u] * A + kk] : 0.0;
    }
    if (lane == 0) Qcol[rows * nw] = -INFINITY;
    __syncwarp();
    const QSharedCol qb(Qcol);
    double ll = 0.0;
    QAcc acc;
    acc.acc_obs = 0.0;
    acc.prod = 1.0;
    acc.er.load(exptab);
    // is every lane's beta non-negative?  (warp-uniform answer, taken once: the step body is compiled twice)
    const bool sign_ok = __all_sync(0xffffffffu, beta >= 0.0);
    auto sweep_chunks = [&](auto sign_tag) {
      constexpr bool SIGN = decltype(sign_tag)::value;
      for (int c = 0; c < nchunks; ++c) {
        const int dslot = c % kQDecStages;
        mbar_wait(&rg.full_dec[dslot], (uint32_t)((c / kQDecStages) & 1));
        const uint4 *recs = rg.dec + dslot * (kQChunk * kQDecWords);
        const int64_t first = (int64_t)c * kQChunk;
        const int cnt = (int)(n - first < kQChunk ? n - first : kQChunk);
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) qgroup_step<QSharedCol, SIGN>(recs + kQDecWords * i, qb, alpha, beta, gamma, acc);
        ll = __dadd_rn(ll, __dsub_rn(acc.acc_obs, log_chunk_product(acc.prod)));  // the octet kernel folds every 32 records too
        acc.acc_obs = 0.0;
        acc.prod = 1.0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&rg.empty_dec[dslot]);
      }
    };
    if (sign_ok) sweep_chunks(std::true_type{});
    else sweep_chunks(std::false_type{});
    if (p < a.P && a.ll_out) a.ll_out[(int64_t)sess * a.P + p] = ll;
  }
  if (a.q_out) {  // every parameter set of a run receives the run's table
    __syncthreads();
    const int64_t pc = (int64_t)blockIdx.x * nw * 32;
    const int64_t np = (a.P - pc < (int64_t)nw * 32) ? (a.P - pc) : (int64_t)nw * 32;
    const int64_t per_set = (int64_t)a.n_tiles * A;
    double *dst = a.q_out + ((int64_t)sess * a.P + pc) * per_set;
    for (int64_t i = tid; i < np * per_set; i += blockDim.x) {
      const int u = (int)(i / per_set);
      const int rest = (int)(i - (int64_t)u * per_set);
      const int tile = rest / A, kk = rest - tile * A;
      const int st = a.state_of_tile[tile];
      dst[i] = st >= 0 ? Q[(st * 8 + kk) * nw + (u >> 5)] : (a.q_init ? a.q_init[(int64_t)tile * A + kk] : 0.0);
    }
  }
}

// =================================== one table per set, in device memory =====================================
// Arbitrary parameter sets: tables T[state * 8 + action][parameter set] in device memory (rows padded to 32 sets, one
// more row of -inf), one column per THREAD.  All threads of a warp walk the same stream, so the sixteen loads and the
// store of a step hit whole rows at consecutive sets: coalesced 256-byte accesses (504 MB of tables for the
// configs[1] grid: HBM-resident, 136 bytes per update).  Same producer warp, decoded records and step bodies as the
// per-warp kernel (bit-identical results); 64 registers, up to 28 consumer warps per SM; one launch per session.
// Replaces the octet kernel as the default for parameter arrays without uniform runs (NM_Q_TABLES=shared keeps it).
__global__ void __launch_bounds__(960, 1) nm_qglobal_kernel(const QArgs a, double *__restrict__ tables, const int64_t ppad,
                                                            const int sess) {
  if (a.mode && *a.mode != 0) return;  // the per-warp kernel owns this sweep
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = a.S, A = a.A;
  QRings rg;
  rg.full_raw = reinterpret_cast<uint64_t *>(smem_raw);
  rg.full_dec = rg.full_raw + kQRawStages;
  rg.empty_dec = rg.full_dec + kQDecStages;
  rg.raw = reinterpret_cast<uint4 *>(smem_raw + kQBarBytes);
  rg.dec = reinterpret_cast<uint4 *>(smem_raw + kQBarBytes + kQRawBytes);
  unsigned char *exptab = smem_raw + kQBarBytes + kQRawBytes + kQDecBytes;
  int16_t *nxt8 = reinterpret_cast<int16_t *>(smem_raw + kQFixedBytes);
  const int tid = threadIdx.x, lane = tid & 31;
  const int nc = blockDim.x - 32;
  ExpLL::fill(exptab, tid);
  for (int i = tid; i < S * 8; i += blockDim.x) {
    const int u = i >> 3, kk = i & 7;
    nxt8[i] = kk < A ? a.next[u * A + kk] : (int16_t)-1;
  }
  const int64_t rec0 = a.offsets[sess];
  const int64_t n = a.offsets[sess + 1] - rec0, n_on = a.online[sess];
  const int nchunks = (int)((n + kQChunk - 1) / kQChunk);
  if (tid == 0) {
    for (int i = 0; i < kQRawStages; ++i) mbar_init(&rg.full_raw[i], 1);
    for (int i = 0; i < kQDecStages; ++i) {
      mbar_init(&rg.full_dec[i], 32);
      mbar_init(&rg.empty_dec[i], nc >> 5);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int rows = S * 8;
  if (tid >= nc) {
    q_producer_loop(rg, a.records + rec0, n, n_on, nchunks, nxt8, (uint32_t)(ppad * 8), (uint32_t)((int64_t)rows * ppad * 8),
                    lane);
  } else {
    const int64_t p = (int64_t)blockIdx.x * nc + tid;
    const bool live = p < a.P;
    const double alpha = live ? a.alpha[p] : 0.0, beta = live ? a.beta[p] : 0.0, gamma = live ? a.gamma[p] : 0.0;
    double *col = tables + (live ? p : (int64_t)blockIdx.x * nc);
    if (live) {
      for (int e = 0; e < rows; ++e) {
        const int u = e >> 3, kk = e & 7;
        col[(int64_t)e * ppad] = (a.q_init && kk < A) ? a.q_init[(int64_t)a.tile_of_state[u] * A + kk] : 0.0;
      }
      col[(int64_t)rows * ppad] = -INFINITY;
    }
    const QGenericCol qb{reinterpret_cast<char *>(col)};
    double ll = 0.0;
    QAcc acc;
    acc.acc_obs = 0.0;
    acc.prod = 1.0;
    acc.er.load(exptab);
    for (int c = 0; c < nchunks; ++c) {
      const int dslot = c % kQDecStages;
      mbar_wait(&rg.full_dec[dslot], (uint32_t)((c / kQDecStages) & 1));
      const uint4 *recs = rg.dec + dslot * (kQChunk * kQDecWords);
      const int64_t first = (int64_t)c * kQChunk;
      const int cnt = (int)(n - first < kQChunk ? n - first : kQChunk);
      if (live) {
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) qgroup_step(recs + kQDecWords * i, qb, alpha, beta, gamma, acc);
      }
      ll = __dadd_rn(ll, __dsub_rn(acc.acc_obs, log_chunk_product(acc.prod)));
      acc.acc_obs = 0.0;
      acc.prod = 1.0;
      __syncwarp();
      if (lane == 0) mbar_arrive(&rg.empty_dec[dslot]);
    }
    if (live && a.ll_out) a.ll_out[(int64_t)sess * a.P + p] = ll;
    if (live && a.q_out) {
      double *dst = a.q_out + ((int64_t)sess * a.P + p) * a.n_tiles * A;
      for (int tile = 0; tile < a.n_tiles; ++tile) {
        const int st = a.state_of_tile[tile];
        for (int kk = 0; kk < A; ++kk)
          dst[(int64_t)tile * A + kk] = st >= 0 ? col[(int64_t)(st * 8 + kk) * ppad]
                                                : (a.q_init ? a.q_init[(int64_t)tile * A + kk] : 0.0);
      }
    }
  }
}

// grow-only scratch per (device, stream) for the tables in device memory
std::mutex &q_tables_mutex() {
  static std::mutex mu;
  return mu;
}
std::map<std::pair<int, void *>, std::pair<double *, size_t>> &q_tables_map() {
  static std::map<std::pair<int, void *>, std::pair<double *, size_t>> scratch;
  return scratch;
}
int q_global_tables(int device, void *stream, size_t bytes, double **out) {
  std::lock_guard<std::mutex> lock(q_tables_mutex());
  auto &slot = q_tables_map()[std::make_pair(device, stream)];
  if (slot.second < bytes) {
    if (slot.first) {
      cudaDeviceSynchronize();  // earlier sweeps may still use the old block
      cudaFree(slot.first);
      slot.first = nullptr;
      slot.second = 0;
    }
    double *buf = nullptr;
    const cudaError_t e = cudaMalloc((void **)&buf, bytes);
    if (e != cudaSuccess) return nm_fail(NM_ERR_NOMEM, "Q tables in device memory: cudaMalloc(%zu): %s", bytes,
                                         cudaGetErrorString(e));
    slot.first = buf;
    slot.second = bytes;
  }
  *out = slot.first;
  return NM_OK;
}

size_t q_global_capacity(int device, void *stream) {
  std::lock_guard<std::mutex> lock(q_tables_mutex());
  const auto it = q_tables_map().find(std::make_pair(device, stream));
  return it == q_tables_map().end() ? 0 : it->second.second;
}

// 1 when "mode 0" of the last sweep on (device, stream) ran with tables in device memory: nm_qtable_sweep_mode
std::map<std::pair<int, void *>, int> &q_last_global() {
  static std::map<std::pair<int, void *>, int> m;
  return m;
}
std::mutex &q_last_mutex() {
  static std::mutex mu;
  return mu;
}

}  // namespace

void nm_q_scratch_release() {
  std::lock_guard<std::mutex> lock(q_tables_mutex());
  int current = 0;
  cudaGetDevice(&current);
  for (auto &kv : q_tables_map()) {
    if (!kv.second.first) continue;
    cudaSetDevice(kv.first.first);
    cudaDeviceSynchronize();
    cudaFree(kv.second.first);
    kv.second = std::make_pair((double *)nullptr, (size_t)0);
  }
  cudaSetDevice(current);
}

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
  const size_t fixed = 1024 + (((size_t)S * 16 + 127) & ~(size_t)127);  // ExpLL table region + packed next[S][8]
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
  // per-set tables: in device memory by default (any maze size, 3x faster), in shared memory (the octet kernel) with
  // NM_Q_TABLES=shared
  const char *env_tab = getenv("NM_Q_TABLES");
  const bool octet = env_tab && env_tab[0] == 's';
  NM_CHECK_ARG(!octet || warps >= 1, "nm_qtable_sweep: the Q-tables of four parameter sets on a maze with %d states do "
               "not fit in shared memory", S);
  const size_t smem = fixed + per_warp * (warps > 0 ? warps : 1);
  cudaError_t e = cudaSuccess;
  if (octet) e = cudaFuncSetAttribute(nm_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
  cudaStream_t cs = (cudaStream_t)stream;
  // Parameter sets that share (alpha, gamma) share their table: the device decides whether every aligned run of 32
  // consecutive sets does, and exactly one of the two kernels does the work (the other returns at its first
  // instruction) -- no host round trip.  NM_Q_GROUP=0 keeps the octet kernel (A/B runs, tests).
  std::lock_guard<std::mutex> launch_lock(runs_launch_mutex());  // decision + launches: one sequence on the stream
  int *mode = nullptr;
  const int rc = runs_mode_word(stream, &mode);
  if (rc != NM_OK) return rc;
  a.mode = mode;
  const char *env = getenv("NM_Q_GROUP");
  const size_t g_fixed = kQFixedBytes + (((size_t)S * 16 + 127) & ~(size_t)127);
  const size_t g_per_warp = ((size_t)S * 8 + 1) * sizeof(double);  // one table + the -inf row
  int g_max = (int)(((size_t)227 * 1024 - g_fixed) / g_per_warp);
  if (g_max > 29) g_max = 29;
  const bool group = !(env && atoi(env) == 0) && g_max >= 1 && (size_t)(S * 8 + 1) * g_max * 8 <= 0xffffffffull;
  runs_decide(group, d_alpha, d_gamma, P, mode, cs);
  if (group) {
    // consumer warps per CTA: spread few runs over the SMs, else the fullest CTA (same staircase as the value-table
    // group kernel: only the busiest scheduler's warp count matters, ties go to the smaller CTA)
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, mz->device);
    const int64_t runs = (P + 31) / 32;
    static const double kWave[9] = {0.0, 1.0, 1.12, 1.29, 1.45, 1.65, 1.91, 2.19, 2.50};
    int nw = 1;
    double best = 1e300;
    const char *env_nw = getenv("NM_Q_GROUP_NW");
    for (int c = 1; c <= g_max; ++c) {
      const int64_t ctas = (int64_t)st->S * ((runs + c - 1) / c);
      const int64_t waves = (ctas + sms - 1) / sms;
      const double cost = (double)waves * kWave[(c + 3) / 4 > 8 ? 8 : (c + 3) / 4];
      if (cost < best * 0.999) {
        best = cost;
        nw = c;
      }
    }
    if (env_nw && atoi(env_nw) >= 1 && atoi(env_nw) <= g_max) nw = atoi(env_nw);
    const size_t g_smem = g_fixed + g_per_warp * nw;
    e = cudaFuncSetAttribute(nm_qgroup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem);
    if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaFuncSetAttribute(nm_qgroup_kernel): %s", cudaGetErrorString(e));
    nm_qgroup_kernel<<<dim3((unsigned)((runs + nw - 1) / nw), (unsigned)st->S, 1), 32 * (nw + 1), g_smem, cs>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_qgroup_kernel launch: %s", cudaGetErrorString(e));
  }
  {
    std::lock_guard<std::mutex> lock(q_last_mutex());
    q_last_global()[std::make_pair(mz->device, stream)] = octet ? 0 : 1;
  }
  if (octet) {
    const int64_t per_cta = (int64_t)warps * 4;
    const int64_t blocks = (P + per_cta - 1) / per_cta;
    NM_CHECK_ARG(blocks <= 2147483647LL, "nm_qtable_sweep: too many parameter sets for one launch");
    nm_q_kernel<<<dim3((unsigned)blocks, (unsigned)st->S, 1), warps * 32, smem, cs>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_q_kernel launch: %s", cudaGetErrorString(e));
    return NM_OK;
  }
  // tables in device memory: slices of at most one full wave (and of 32-bit row offsets), one launch per session
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, mz->device);
  const int64_t rows1 = (int64_t)S * 8 + 1;
  int64_t max_sets = (((int64_t)0xffffffffLL / 8) / rows1) & ~(int64_t)31;
  NM_CHECK_ARG(max_sets >= 32, "nm_qtable_sweep: a maze with %d states is too large", S);
  if (max_sets > (int64_t)sms * 896) max_sets = (int64_t)sms * 896;
  e = cudaFuncSetAttribute(nm_qglobal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_fixed);
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaFuncSetAttribute(nm_qglobal_kernel): %s", cudaGetErrorString(e));
  if (a.mode) {
    // the device is deciding whether the table-per-warp kernel owns this sweep (then every launch below is a no-op):
    // do not GROW the per-set table scratch (504 MB for the configs[1] grid) for launches that may not need it -- when
    // the block kept for this (device, stream) is too small, wait for the decision once
    const int64_t first = P < max_sets ? P : max_sets;
    const size_t need = sizeof(double) * (size_t)rows1 * (size_t)((first + 31) & ~(int64_t)31);
    if (q_global_capacity(mz->device, stream) < need) {
      int decided = 0;
      e = cudaStreamSynchronize(cs);
      if (e == cudaSuccess) e = cudaMemcpy(&decided, a.mode, sizeof(int), cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_qtable_sweep: reading the sweep mode: %s", cudaGetErrorString(e));
      if (decided == 1) return NM_OK;
    }
  }
  for (int64_t lo = 0; lo < P; lo += max_sets) {
    QArgs b = a;
    b.P = P - lo < max_sets ? P - lo : max_sets;
    b.alpha = d_alpha + lo;
    b.beta = d_beta + lo;
    b.gamma = d_gamma + lo;
    const int64_t ppad = (b.P + 31) & ~(int64_t)31;
    int nc = (int)(((b.P + sms - 1) / sms + 31) / 32) * 32;
    nc = nc < 32 ? 32 : nc > 896 ? 896 : nc;
    double *tables = nullptr;
    const int rc2 = q_global_tables(mz->device, stream, sizeof(double) * (size_t)rows1 * (size_t)ppad, &tables);
    if (rc2 != NM_OK) return rc2;
    for (int sess = 0; sess < st->S; ++sess) {
      QArgs c = b;  // outputs of the whole call are [S, P, ...]: shift to this slice's first parameter set
      if (d_ll_out) c.ll_out = d_ll_out + (int64_t)sess * (P - b.P) + lo;
      if (d_q_out) c.q_out = d_q_out + ((int64_t)sess * (P - b.P) + lo) * a.n_tiles * n_actions;
      nm_qglobal_kernel<<<(unsigned)((b.P + nc - 1) / nc), nc + 32, g_fixed, cs>>>(c, tables, ppad, sess);
      e = cudaGetLastError();
      if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_qglobal_kernel launch: %s", cudaGetErrorString(e));
    }
  }
  return NM_OK;
}

extern "C" int nm_qtable_sweep_mode(void *stream, int *mode_out) {
  const int rc = runs_mode_read(stream, mode_out, "nm_qtable_sweep_mode");
  if (rc == NM_OK && *mode_out == 0) {  // per-set tables: shared memory (0) or device memory (2)?
    int device = 0;
    cudaGetDevice(&device);
    std::lock_guard<std::mutex> lock(q_last_mutex());
    const auto it = q_last_global().find(std::make_pair(device, stream));
    if (it != q_last_global().end() && it->second) *mode_out = 2;
  }
  return rc;
}
