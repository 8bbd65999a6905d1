// Device code of the parameter-grid sweeps.  Included by nm_fit.cu inside its anonymous namespace.
//
// CTA anatomy (both kernels):   blockDim.x = nc consumer threads + ONE producer warp (the last warp)
//   producer warp   issues the TMA bulk copies (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP) of raw
//                   32-byte step records into a 2-deep raw ring, DECODES every landed chunk once (state ids ->
//                   shared-memory byte offsets of table rows; one record per lane) into a 4-deep ring of
//                   64-byte decoded records and hands chunks to the consumers through mbarriers
//                   (full_dec / empty_dec).  No __syncthreads in the step loop.
//   consumer thread owns one parameter set (alpha, beta, gamma); its value table is a column of shared
//                   memory.  All threads walk the SAME stream, so every LDS/STS of a step hits one row ->
//                   conflict-free, warp-uniform control flow.
//
// Two step bodies, both specialised on the (warp-uniform) candidate count K through a switch:
//   EXACT  (nm_fit_exact_kernel)  row = V[state][thread] (8 B).  Softmax log-probability evaluated as the
//          reference's formula reads -- beta * V[cand], subtract the max, K exps, numpy-order sum
//          (exploration_model.py:87-93) -- with a range-restricted exp polynomial and the log taken once per
//          chunk on the PRODUCT of the sums.  Also used without the likelihood (V-table sweeps, replay
//          passes) and as the safety net of MEMO.
//   MEMO   (nm_fit_memo_kernel)   row = {V, W}[state][thread] (16 B), W = exp(beta * V).  A step changes V at
//          exactly one state, so the softmax denominator is sum_k W[cand_k] and only ONE exp per step is
//          needed; log p = beta * V[obs] - log(sum W).  One LDS.128 fetches a candidate's V and W together.
//          The exp of step t is evaluated while step t+1 runs: the producer moves the candidates equal to
//          the previous step's state (whose W is still in flight) to the end of the list and counts them
//          (n_hz), so the consumer adds n_hz * w_inflight instead of selecting per candidate.
//          Mathematically identical to EXACT; |beta * V| is guarded at 700 -- a parameter set that ever
//          leaves the guard writes NaN and is recomputed by the EXACT kernel in "redo" mode (nm_fit.cu).
// In both, the V recurrence is the same sequence of _rn operations (no FMA contraction) as the reference's
// numpy scalars: V-tables are bit-identical (rl.py:184-187, 227-230, 257-259).  (max / min are
// compare-selects: like np.max they keep the first of equal values; NaN tables are out of contract.)

#include <type_traits>

constexpr int kChunk = 32;      // records per chunk = one per producer lane
constexpr int kRawStages = 2;   // raw (TMA landing) ring
#ifndef NM_FIT_DEC_STAGES
#define NM_FIT_DEC_STAGES 4
#endif
constexpr int kDecStages = NM_FIT_DEC_STAGES;   // decoded ring
constexpr int kRawBytes = kRawStages * kChunk * 32;
constexpr int kDecBytes = kDecStages * kChunk * 64;
constexpr int kBarBytes = 256;  // full_raw[2], full_dec[kDecStages], empty_dec[kDecStages]
static_assert((kRawStages + 2 * kDecStages) * 8 <= kBarBytes, "mbarrier block too small");
constexpr int kTabBytes = 1024;  // room for the 512-byte-aligned 2^(j/64) table of ExpLL (nm_exp.cuh)
constexpr int kFixedBytes = kBarBytes + kRawBytes + kDecBytes + kTabBytes;
constexpr uint32_t kNone = 0xFFFFFFFFu;

// decoded record, 64 bytes = 4 x uint4 (offsets = byte offsets of table rows = state * row_bytes):
//   q0 = {r.lo, r.hi, off_s, off_obs}          off_obs = kNone on unobserved / replay steps
//   q1 = {K | Kw << 8 | online << 16, off_s1, off_c0, off_c1}     Kw = candidates whose W is current (MEMO)
//   q2 = {off_c2 .. off_c5}
//   q3 = {off_c6, off_c7, n_hz as a double (lo, hi)}              n_hz = K - Kw (MEMO)

struct Rings {
  uint64_t *full_raw, *full_dec, *empty_dec;
  uint4 *raw, *dec;
};

__device__ __forceinline__ Rings carve_rings(unsigned char *smem) {
  Rings r;
  r.full_raw = reinterpret_cast<uint64_t *>(smem);
  r.full_dec = r.full_raw + kRawStages;
  r.empty_dec = r.full_dec + kDecStages;
  r.raw = reinterpret_cast<uint4 *>(smem + kBarBytes);
  r.dec = reinterpret_cast<uint4 *>(smem + kBarBytes + kRawBytes);
  return r;
}

__device__ __forceinline__ void init_rings(const Rings &r, int n_consumer_warps) {
  for (int i = 0; i < kRawStages; ++i) mbar_init(&r.full_raw[i], 1);
  for (int i = 0; i < kDecStages; ++i) {
    mbar_init(&r.full_dec[i], 32);
    mbar_init(&r.empty_dec[i], n_consumer_warps);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// ---- producer warp ---------------------------------------------------------------------------------
template <bool MEMO>
__device__ __forceinline__ void producer_loop(const Rings &rg, const nm_step_rec *src, int64_t n, int64_t n_on,
                                              int nchunks, uint32_t row_bytes, int lane) {
  auto issue = [&](int c) {
    const int slot = c % kRawStages;
    const int64_t first = (int64_t)c * kChunk;
    const uint32_t bytes = (uint32_t)((n - first < kChunk ? n - first : kChunk) * 32);
    mbar_expect_tx(&rg.full_raw[slot], bytes);
    tma_bulk_g2s(rg.raw + slot * (kChunk * 2), src + first, bytes, &rg.full_raw[slot]);
  };
  if (lane == 0)
    for (int c = 0; c < kRawStages && c < nchunks; ++c) issue(c);
  uint32_t carry_s = kNone;  // state of the last record of the previous chunk (MEMO hazard detection)
  for (int c = 0; c < nchunks; ++c) {
    const int rslot = c % kRawStages, dslot = c % kDecStages;
    if (c >= kDecStages) mbar_wait_relaxed(&rg.empty_dec[dslot], (uint32_t)(((c / kDecStages) - 1) & 1));
    mbar_wait_relaxed(&rg.full_raw[rslot], (uint32_t)((c / kRawStages) & 1));
    const int64_t first = (int64_t)c * kChunk;
    const int cnt = (int)(n - first < kChunk ? n - first : kChunk);
    const int src_lane = lane < cnt ? lane : cnt - 1;
    const uint4 w0 = rg.raw[(rslot * kChunk + src_lane) * 2], w1 = rg.raw[(rslot * kChunk + src_lane) * 2 + 1];
    const uint32_t s = w0.z & 0xffffu, s1 = w0.z >> 16, K = w0.w & 0xffu, obs = (w0.w >> 8) & 0xffu;
    const uint32_t cw[4] = {w1.x, w1.y, w1.z, w1.w};
    uint32_t cs[NM_MAX_CAND];
#pragma unroll
    for (int k = 0; k < NM_MAX_CAND; ++k) cs[k] = (k & 1) ? (cw[k >> 1] >> 16) : (cw[k >> 1] & 0xffffu);
    const bool online = first + lane < n_on;
    uint32_t oobs = kNone;
#pragma unroll
    for (int k = 0; k < NM_MAX_CAND; ++k)
      if ((uint32_t)k == obs) oobs = cs[k] * row_bytes;
    if (!online) oobs = kNone;  // replay passes update V only (rl.py:419-424)
    uint32_t oc[NM_MAX_CAND];
    uint32_t Kw = K;
    if (MEMO) {
      // state updated by the previous step: its W is still being computed when this step runs
      uint32_t prev = __shfl_up_sync(0xffffffffu, s, 1);
      if (lane == 0) prev = carry_s;
      carry_s = __shfl_sync(0xffffffffu, s, cnt - 1);
      if (!online || first + lane == 0) prev = kNone;
      // stable partition: candidates != prev first, the (n_hz) candidates == prev last
      uint32_t head = 0, tail = K;
#pragma unroll
      for (int k = 0; k < NM_MAX_CAND; ++k) oc[k] = 0u;
#pragma unroll
      for (int k = 0; k < NM_MAX_CAND; ++k) {
        if ((uint32_t)k < K) {
          const bool hz = cs[k] == prev;
          const uint32_t pos = hz ? --tail : head++;
#pragma unroll
          for (int j = 0; j < NM_MAX_CAND; ++j)
            if ((uint32_t)j == pos) oc[j] = cs[k] * row_bytes;
        }
      }
      Kw = head;
    } else {
#pragma unroll
      for (int k = 0; k < NM_MAX_CAND; ++k) oc[k] = cs[k] * row_bytes;
      // likelihood steps: the OBSERVED candidate goes first (swapped with candidate 0), so that the consumer's
      // exponential argument of candidate 0 is the observed log-odds.  The value update takes the extremum of the
      // same multiset of table entries: bit-identical tables whatever the order.
      if (oobs != kNone) {
#pragma unroll
        for (int k = 1; k < NM_MAX_CAND; ++k)
          if ((uint32_t)k == obs) {
            oc[k] = oc[0];
            oc[0] = oobs;
          }
      }
    }
    if (lane < cnt) {
      const double nhz = (double)(K - Kw);
      uint4 *d = rg.dec + (dslot * kChunk + lane) * 4;
      d[0] = make_uint4(w0.x, w0.y, s * row_bytes, oobs);
      d[1] = make_uint4(K | (Kw << 8) | (online ? 1u << 16 : 0u), s1 * row_bytes, oc[0], oc[1]);
      d[2] = make_uint4(oc[2], oc[3], oc[4], oc[5]);
      d[3] = make_uint4(oc[6], oc[7], (uint32_t)__double2loint(nhz), (uint32_t)__double2hiint(nhz));
    }
    mbar_arrive(&rg.full_dec[dslot]);  // every lane: releases its own stores (count = 32)
    __syncwarp();                      // all lanes have read the raw slot before it is refilled
    if (lane == 0 && c + kRawStages < nchunks) issue(c + kRawStages);
  }
}

// ---- exponentials ------------------------------------------------------------------------------------
// 2^k * exp(r), |r| <= ln2/2: round x/ln2 to nearest with the magic-constant trick, two-step Cody-Waite
// reduction, degree-13 Taylor/Horner (truncation < 4e-18 relative), exponent add.  Coefficients live in
// constant memory so that every DFMA takes its constant operand from the constant bank (no extra moves).
__constant__ double c_exp[16] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0,
                                 1.0 / 362880.0,     1.0 / 40320.0,     1.0 / 5040.0,     1.0 / 720.0,
                                 1.0 / 120.0,        1.0 / 24.0,        1.0 / 6.0,        0.5,
                                 1.4426950408889634074,          // log2(e)
                                 6755399441055744.0,             // 1.5 * 2^52
                                 -6.93147180369123816490e-01,    // -ln2 (high part)
                                 -1.90821492927058770002e-10};   // -ln2 (low part)

__device__ __forceinline__ double exp_core(double x) {
  const double t = fma(x, c_exp[12], c_exp[13]);
  const int k = __double2loint(t);
  const double n = __dsub_rn(t, c_exp[13]);
  double r = fma(n, c_exp[14], x);
  r = fma(n, c_exp[15], r);
  double p = fma(c_exp[0], r, c_exp[1]);
#pragma unroll
  for (int i = 2; i < 12; ++i) p = fma(p, r, c_exp[i]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
// exp(x) for x <= 0; below -708 the result is flushed to 0 (its share of a sum whose largest term is 1 is
// < 1e-307).
__device__ __forceinline__ double exp_nonpos(double x) { return (x < -708.0) ? 0.0 : exp_core(x); }

#include "nm_exp.cuh"

#include "nm_step_math.cuh"

// Where a value table lives.  GenericCol: any pointer (device-memory tables; shared-memory columns whose address the
// compiler may recompute).  SharedCol: a 32-bit shared-window address pinned in a register through an opaque move --
// under the 64-register cap of the per-warp-table kernel the compiler otherwise REMATERIALISES the column base and
// the exponential table's address in every step (S2R / S2UR / ULEA / shifts: ~10 instructions per step).
struct GenericCol {
  char *vb;
  __device__ __forceinline__ double ld(uint32_t off) const { return *reinterpret_cast<const double *>(vb + off); }
  __device__ __forceinline__ void st(uint32_t off, double v) const { *reinterpret_cast<double *>(vb + off) = v; }
};
struct SharedCol {
  uint32_t base;
  __device__ __forceinline__ explicit SharedCol(const void *col) {
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

// =================================== EXACT body ===========================================================
struct ExactAcc {
  double acc_obs, prod;
  ExpLL er;
};

// SIGN: the caller has established (once, outside the step loop, for the whole warp) that beta has the sign for which
// max_k fl(beta * V_k) == fl(beta * extremum_k V_k) -- non-negative for the max-backup learner, non-positive for the
// min-backup one -- so the per-step test and its predicated twin disappear from the step body.
template <int ALG, bool LL, int K, class Col, bool SIGN = false>
__device__ __forceinline__ void exact_step_k(const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3, const Col vb,
                                             const double alpha, const double beta, const double gamma,
                                             ExactAcc &acc) {
  const double r = __hiloint2double((int)q0.y, (int)q0.x);
  const uint32_t offs[NM_MAX_CAND] = {q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y};
  const double vs = vb.ld(q0.z);
  double v[NM_MAX_CAND];
  double boot = 0.0;
  if constexpr (LL || ALG != NM_ALG_TD0) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = vb.ld(offs[k]);
  }
  if constexpr (ALG == NM_ALG_TD0) boot = vb.ld(q1.y);
  else boot = extremum<ALG, K>(v);
  if (LL && q0.w != kNone) {  // observed step (warp-uniform); the producer put the OBSERVED candidate first
    // val_k = beta * V[cand_k] - max (exploration_model.py:87, 92) in one fused instruction per candidate; the maximum
    // itself is fl(beta * max V) for a non-negative beta (rounding is monotonic), else the largest rounded product
    double m;
    if (SIGN || (ALG == NM_ALG_POSITIVE && beta >= 0.0) || (ALG == NM_ALG_NEGATIVE && beta <= 0.0)) {
      m = __dmul_rn(beta, boot);
    } else {
      double z[NM_MAX_CAND];
#pragma unroll
      for (int k = 0; k < K; ++k) z[k] = __dmul_rn(beta, v[k]);
      m = extremum<NM_ALG_POSITIVE, K>(z);
    }
    const double neg_m = -m;
    double x[NM_MAX_CAND], e[NM_MAX_CAND];
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = fma(beta, v[k], neg_m);
#pragma unroll
    for (int k = 0; k < K; ++k) e[k] = acc.er(x[k]);  // :93
    const double sum = np_sum_fixed<K>(e);
    acc.acc_obs = __dadd_rn(acc.acc_obs, x[0]);  // log p = val_obs - log(sum): val_obs is candidate 0's argument
    acc.prod = __dmul_rn(acc.prod, sum);         // sum in [1, 8]; 32 steps per chunk -> prod <= 2^96
  }
  const double rpe = __dsub_rn(__dadd_rn(r, __dmul_rn(gamma, boot)), vs);  // rl.py:185, 228, 257
  vb.st(q0.z, __dadd_rn(vs, __dmul_rn(alpha, rpe)));                        // rl.py:186-187, 258-259
}

template <int ALG, bool LL, class Col, bool SIGN = false>
__device__ __forceinline__ void exact_step(const uint4 *__restrict__ rec, const Col vb, const double alpha,
                                           const double beta, const double gamma, ExactAcc &acc) {
  const uint4 q0 = rec[0], q1 = rec[1], q2 = rec[2], q3 = rec[3];
  if constexpr (!LL && ALG == NM_ALG_TD0) {  // no candidates needed
    exact_step_k<NM_ALG_TD0, false, 0, Col>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc);
  } else {
#define NM_CASE(KK) \
  case KK: exact_step_k<ALG, LL, KK, Col, SIGN>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc); break;
    switch (q1.x & 0xffu) {  // K, warp-uniform
      NM_CASE(1) NM_CASE(2) NM_CASE(3) NM_CASE(4) NM_CASE(5) NM_CASE(6) NM_CASE(7) NM_CASE(8)
      default:  // K == 0 only occurs in TD0 streams built without candidates (validated on the host)
        if constexpr (ALG == NM_ALG_TD0)
          exact_step_k<NM_ALG_TD0, false, 0, Col>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc);
        break;
    }
#undef NM_CASE
  }
}

// initial table of one parameter set; `stride` = doubles between consecutive states of a column
__device__ __forceinline__ void init_v_column(const FitArgs &a, double *col, int stride, int64_t p, bool live) {
  const int NV = a.num_states;
  if (a.v_init == nullptr) {
    for (int s = 0; s < NV; ++s) col[s * stride] = 0.0;  // rl.py:66
  } else if (!a.v_init_per_unit) {
    for (int s = 0; s < NV; ++s) col[s * stride] = a.v_init[a.tile_of_state[s]];  // rl.py:63-64
  } else {
    for (int s = 0; s < NV; ++s) col[s * stride] = live ? a.v_init[p * a.n_tiles + a.tile_of_state[s]] : 0.0;
  }
}

// dense [S, P, X*Y] tables, written coalesced through a transposed read of the shared-memory columns
// (all threads of the CTA).  V of (state st, unit u) sits at V[st * row + u * unit] doubles.
__device__ __forceinline__ void write_v_out(const FitArgs &a, const double *V, int nc, int row, int unit, int sess) {
  const int64_t p0 = (int64_t)blockIdx.x * nc;
  const int64_t np = (a.P - p0 < nc) ? (a.P - p0) : nc;
  const int nT = a.n_tiles;
  double *dst = a.v_out + ((int64_t)sess * a.P + p0) * nT;
  for (int64_t i = threadIdx.x; i < np * nT; i += blockDim.x) {
    const int u = (int)(i / nT), tile = (int)(i - (int64_t)u * nT);
    const int st = a.state_of_tile[tile];
    double val;
    if (st >= 0) val = V[st * row + u * unit];
    else if (a.v_init == nullptr) val = 0.0;
    else val = a.v_init_per_unit ? a.v_init[(p0 + u) * nT + tile] : a.v_init[tile];
    dst[i] = val;
  }
}

// REDO = true: safety net of the MEMO kernel -- only CTAs holding a parameter set whose ll_out is NaN run,
// and only those parameter sets are rewritten.
template <int ALG, bool LL, bool REDO>
__global__ void __launch_bounds__(512, 1) nm_fit_exact_kernel(const FitArgs a) {
  if (a.mode && *a.mode != 0) return;  // a group kernel owns this sweep
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Rings rg = carve_rings(smem_raw);
  double *V = reinterpret_cast<double *>(smem_raw + kFixedBytes);
  const int tid = threadIdx.x;
  const int nc = blockDim.x - 32;
  const int sess = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * nc + tid;
  const bool live = tid < nc && p < a.P;
  bool wanted = live;
  if (REDO) {
    wanted = live && isnan(a.ll_out[(int64_t)sess * a.P + p]);
    if (!__syncthreads_or(wanted ? 1 : 0)) return;  // whole CTA: nothing to redo
  }
  const int64_t rec0 = a.offsets[sess];
  const int64_t n = a.offsets[sess + 1] - rec0;
  const int64_t n_on = a.online[sess];
  const int nchunks = (int)((n + kChunk - 1) / kChunk);
  if (tid == 0) init_rings(rg, nc >> 5);
  ExpLL::fill(smem_raw + kBarBytes + kRawBytes + kDecBytes, tid);
  __syncthreads();

  if (tid >= nc) {
    producer_loop<false>(rg, a.records + rec0, n, n_on, nchunks, (uint32_t)nc * 8u, tid - nc);
  } else {
    const double alpha = live ? a.alpha[p] : 0.0;
    const double beta = (LL && live) ? a.beta[p] : 0.0;
    const double gamma = live ? a.gamma[p] : 0.0;
    double *Vcol = V + tid;
    init_v_column(a, Vcol, nc, p, live);
    const SharedCol vb(Vcol);
    double ll = 0.0;
    ExactAcc acc;
    acc.acc_obs = 0.0;
    acc.prod = 1.0;
    acc.er.load(smem_raw + kBarBytes + kRawBytes + kDecBytes);
    const int lane = tid & 31;
    // the sign test of beta, once per warp (see the group kernel); lanes without a parameter set carry beta = 0
    const bool sign_ok = LL && ALG != NM_ALG_TD0 &&
                         __all_sync(0xffffffffu, ALG == NM_ALG_POSITIVE ? beta >= 0.0 : beta <= 0.0);
    auto sweep_chunks = [&](auto sign_tag) {
      constexpr bool SIGN = decltype(sign_tag)::value;
      for (int c = 0; c < nchunks; ++c) {
        const int dslot = c % kDecStages;
        mbar_wait(&rg.full_dec[dslot], (uint32_t)((c / kDecStages) & 1));
        const uint4 *recs = rg.dec + dslot * (kChunk * 4);
        const int64_t first = (int64_t)c * kChunk;
        const int cnt = (int)(n - first < kChunk ? n - first : kChunk);
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) exact_step<ALG, LL, SharedCol, SIGN>(recs + 4 * i, vb, alpha, beta, gamma, acc);
        if (LL) {  // fold the chunk:  ll += sum(zobs - m) - log(prod of the softmax denominators)
          ll = __dadd_rn(ll, __dsub_rn(acc.acc_obs, log_chunk_product(acc.prod)));
          acc.acc_obs = 0.0;
          acc.prod = 1.0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&rg.empty_dec[dslot]);
      }
    };
    if constexpr (LL && ALG != NM_ALG_TD0) {
      if (sign_ok) sweep_chunks(std::true_type{});
      else sweep_chunks(std::false_type{});
    } else {
      sweep_chunks(std::false_type{});
    }
    if (LL && wanted && a.ll_out) a.ll_out[(int64_t)sess * a.P + p] = ll;
  }
  if (!REDO && a.v_out) {
    __syncthreads();
    write_v_out(a, V, nc, nc, 1, sess);
  }
}

// =================================== GROUP body ============================================================
// A value table depends on (alpha, gamma) only -- beta enters the likelihood, not the recurrence -- so on a Cartesian
// grid whose beta axis varies fastest, runs of consecutive parameter sets own IDENTICAL tables.  When every aligned
// run of 32 parameter sets shares (alpha, gamma) (checked on the device by nm_group_detect_kernel), a WARP keeps ONE
// table column for its 32 sets: loads of V become broadcasts, the table update is the same value in every lane (all
// lanes store it: same address, same bits) and shared memory stops limiting residency -- the kernel is compiled for
// 64 registers, so up to 29 consumer warps + the producer share an SM (the exact kernel: 14).  Per parameter set the
// operation sequence is the exact body's: results are bit-identical.
//
// mode[0] = 1 if every aligned run of 32 parameter sets shares (alpha, gamma), else 0.
__global__ void nm_group_init_kernel(int *mode) { mode[0] = 1; }
__global__ void __launch_bounds__(256) nm_group_detect_kernel(const double *__restrict__ alpha,
                                                              const double *__restrict__ gamma, int64_t P, int *mode) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int64_t head = p & ~(int64_t)31;
  if (__double_as_longlong(alpha[p]) != __double_as_longlong(alpha[head]) ||
      __double_as_longlong(gamma[p]) != __double_as_longlong(gamma[head]))
    mode[0] = 0;
}
__global__ void nm_group_off_kernel(int *mode) { mode[0] = 0; }

// Runs only when *a.mode == 1 (the exact kernel takes mode 0).  blockDim.x = 32 * (nw + 1): nw consumer warps, each
// with its own table column V[state][warp], and the producer warp.
template <int ALG>
__global__ void __launch_bounds__(960, 1) nm_fit_group_kernel(const FitArgs a) {
  if (*a.mode != 1) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Rings rg = carve_rings(smem_raw);
  double *V = reinterpret_cast<double *>(smem_raw + kFixedBytes);
  const int tid = threadIdx.x;
  const int nw = (blockDim.x >> 5) - 1;
  const int sess = blockIdx.y;
  const int64_t rec0 = a.offsets[sess];
  const int64_t n = a.offsets[sess + 1] - rec0;
  const int64_t n_on = a.online[sess];
  const int nchunks = (int)((n + kChunk - 1) / kChunk);
  if (tid == 0) init_rings(rg, nw);
  ExpLL::fill(smem_raw + kBarBytes + kRawBytes + kDecBytes, tid);
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  if (warp >= nw) {
    producer_loop<false>(rg, a.records + rec0, n, n_on, nchunks, (uint32_t)nw * 8u, lane);
  } else {
    const int64_t p0 = ((int64_t)blockIdx.x * nw + warp) * 32;
    const int64_t p = p0 + lane;
    const bool alive = p0 < a.P;  // warp-uniform
    const double alpha = alive ? a.alpha[p0] : 0.0, gamma = alive ? a.gamma[p0] : 0.0;
    const double beta = p < a.P ? a.beta[p] : 0.0;
    double *Vcol = V + warp;
    for (int s = lane; s < a.num_states; s += 32)
      Vcol[s * nw] = a.v_init ? a.v_init[a.tile_of_state[s]] : 0.0;  // rl.py:63-66 (shared initial table)
    __syncwarp();
    const SharedCol vb(Vcol);
    double ll = 0.0;
    ExactAcc acc;
    acc.acc_obs = 0.0;
    acc.prod = 1.0;
    acc.er.load(smem_raw + kBarBytes + kRawBytes + kDecBytes);
    // does every lane's beta have the sign that lets the softmax maximum be fl(beta * extremum V)?  (warp-uniform answer,
    // taken once: the step body is compiled twice)
    const bool sign_ok = ALG != NM_ALG_TD0 &&
                         __all_sync(0xffffffffu, ALG == NM_ALG_POSITIVE ? beta >= 0.0 : beta <= 0.0);
    auto sweep_chunks = [&](auto sign_tag) {
      constexpr bool SIGN = decltype(sign_tag)::value;
      for (int c = 0; c < nchunks; ++c) {
        const int dslot = c % kDecStages;
        mbar_wait(&rg.full_dec[dslot], (uint32_t)((c / kDecStages) & 1));
        const uint4 *recs = rg.dec + dslot * (kChunk * 4);
        const int64_t first = (int64_t)c * kChunk;
        const int cnt = (int)(n - first < kChunk ? n - first : kChunk);
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) exact_step<ALG, true, SharedCol, SIGN>(recs + 4 * i, vb, alpha, beta, gamma, acc);
        ll = __dadd_rn(ll, __dsub_rn(acc.acc_obs, log_chunk_product(acc.prod)));  // fold the chunk exactly like the exact kernel
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
  if (a.v_out) {  // every parameter set of a run receives the run's table
    __syncthreads();
    const int64_t pc = (int64_t)blockIdx.x * nw * 32;
    const int64_t np = (a.P - pc < (int64_t)nw * 32) ? (a.P - pc) : (int64_t)nw * 32;
    const int nT = a.n_tiles;
    double *dst = a.v_out + ((int64_t)sess * a.P + pc) * nT;
    for (int64_t i = tid; i < np * nT; i += blockDim.x) {
      const int u = (int)(i / nT), tile = (int)(i - (int64_t)u * nT);
      const int st = a.state_of_tile[tile];
      dst[i] = st >= 0 ? V[st * nw + (u >> 5)] : (a.v_init ? a.v_init[tile] : 0.0);
    }
  }
}

// =================================== GLOBAL body ===========================================================
// Value tables in device memory: T[state][parameter set] (a row = all parameter sets of the launch, padded to 32), one
// column per thread.  All threads of a warp walk the same stream, so a step's loads and its store hit one row at
// consecutive parameter sets: coalesced 256-byte accesses, served by L2 while the working set (states x sets x 8 B)
// fits its 126 MB, by HBM beyond.  Nothing limits the maze size, and with 64 registers 28 consumer warps per SM hide
// the L2 latency -- twice the residency the shared-memory layout allows, which makes this the faster kernel for
// arbitrary parameter sets (configs[1]: 4.42 ms against 4.78 ms) and the default for them; NM_FIT_TABLES=shared
// keeps the shared-memory kernel.  One launch per session (the scratch holds one table per set).
// Same producer warp / rings / exact_step bodies: results are bit-identical to the shared-memory kernels.
template <int ALG, bool LL>
__global__ void __launch_bounds__(960, 1) nm_fit_global_kernel(const FitArgs a, double *__restrict__ tables,
                                                               const int64_t ppad, const int sess) {
  if (a.mode && *a.mode != 0) return;  // the group kernel owns this sweep
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Rings rg = carve_rings(smem_raw);
  const int tid = threadIdx.x;
  const int nc = blockDim.x - 32;
  const int64_t p = (int64_t)blockIdx.x * nc + tid;
  const bool live = tid < nc && p < a.P;
  const int64_t rec0 = a.offsets[sess];
  const int64_t n = a.offsets[sess + 1] - rec0;
  const int64_t n_on = a.online[sess];
  const int nchunks = (int)((n + kChunk - 1) / kChunk);
  if (tid == 0) init_rings(rg, nc >> 5);
  ExpLL::fill(smem_raw + kBarBytes + kRawBytes + kDecBytes, tid);
  __syncthreads();

  if (tid >= nc) {
    producer_loop<false>(rg, a.records + rec0, n, n_on, nchunks, (uint32_t)(ppad * 8), tid - nc);
  } else {
    const double alpha = live ? a.alpha[p] : 0.0;
    const double beta = (LL && live) ? a.beta[p] : 0.0;
    const double gamma = live ? a.gamma[p] : 0.0;
    double *col = tables + (live ? p : (int64_t)blockIdx.x * nc);  // dead lanes shadow a live column, store nothing new
    const int NV = a.num_states;
    if (live) {
      for (int s = 0; s < NV; ++s) {
        double v0 = 0.0;  // rl.py:63-66
        if (a.v_init) v0 = a.v_init_per_unit ? a.v_init[p * a.n_tiles + a.tile_of_state[s]] : a.v_init[a.tile_of_state[s]];
        col[(int64_t)s * ppad] = v0;
      }
    }
    const GenericCol vb{reinterpret_cast<char *>(col)};
    double ll = 0.0;
    ExactAcc acc;
    acc.acc_obs = 0.0;
    acc.prod = 1.0;
    acc.er.load(smem_raw + kBarBytes + kRawBytes + kDecBytes);
    const int lane = tid & 31;
    // the sign test of beta, once per warp (see the group kernel); lanes without a parameter set carry beta = 0
    const bool sign_ok = LL && ALG != NM_ALG_TD0 &&
                         __all_sync(0xffffffffu, ALG == NM_ALG_POSITIVE ? beta >= 0.0 : beta <= 0.0);
    auto sweep_chunks = [&](auto sign_tag) {
      constexpr bool SIGN = decltype(sign_tag)::value;
      for (int c = 0; c < nchunks; ++c) {
        const int dslot = c % kDecStages;
        mbar_wait(&rg.full_dec[dslot], (uint32_t)((c / kDecStages) & 1));
        const uint4 *recs = rg.dec + dslot * (kChunk * 4);
        const int64_t first = (int64_t)c * kChunk;
        const int cnt = (int)(n - first < kChunk ? n - first : kChunk);
        if (live) {
#pragma unroll 1
          for (int i = 0; i < cnt; ++i) exact_step<ALG, LL, GenericCol, SIGN>(recs + 4 * i, vb, alpha, beta, gamma, acc);
        }
        if (LL) {
          ll = __dadd_rn(ll, __dsub_rn(acc.acc_obs, log_chunk_product(acc.prod)));
          acc.acc_obs = 0.0;
          acc.prod = 1.0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&rg.empty_dec[dslot]);
      }
    };
    if constexpr (LL && ALG != NM_ALG_TD0) {
      if (sign_ok) sweep_chunks(std::true_type{});
      else sweep_chunks(std::false_type{});
    } else {
      sweep_chunks(std::false_type{});
    }
    if (LL && live && a.ll_out) a.ll_out[(int64_t)sess * a.P + p] = ll;
    if (live && a.v_out) {
      double *dst = a.v_out + ((int64_t)sess * a.P + p) * a.n_tiles;
      for (int tile = 0; tile < a.n_tiles; ++tile) {
        const int st = a.state_of_tile[tile];
        double val;
        if (st >= 0) val = col[(int64_t)st * ppad];
        else if (a.v_init == nullptr) val = 0.0;
        else val = a.v_init_per_unit ? a.v_init[p * a.n_tiles + tile] : a.v_init[tile];
        dst[tile] = val;
      }
    }
  }
}

// =================================== MEMO body =============================================================
struct MemoState {
  double acc_obs;     // sum over observed steps of beta * V[obs]
  double prod;        // running product of the softmax denominators, mantissa kept in [1, 2)
  int eacc;           // ... and its accumulated binary exponent
  double pend_arg;    // beta * V_new of the previous step: its exp is evaluated during the current step
  uint32_t pend_off;  // row offset whose W is pending (kNone: nothing pending)
  bool bad;           // |beta * V| left the guard: recompute this parameter set exactly
};

template <int ALG, int K>
__device__ __forceinline__ void memo_step_k(const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3, char *vb,
                                            const double alpha, const double beta, const double gamma,
                                            MemoState &st) {
  const double r = __hiloint2double((int)q0.y, (int)q0.x);
  const uint32_t offs[NM_MAX_CAND] = {q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y};
  const uint32_t Kw = (q1.x >> 8) & 0xffu;
  const bool online = (q1.x >> 16) != 0u;  // W is maintained over the online pass only (warp-uniform)
  double *ps = reinterpret_cast<double *>(vb + q0.z);
  const double vs = *ps;
  // one LDS.128 per candidate: {V, W}
  double v[NM_MAX_CAND], w[NM_MAX_CAND];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double2 vw = *reinterpret_cast<const double2 *>(vb + offs[k]);
    v[k] = vw.x;
    w[k] = vw.y;
  }
  double boot;
  if constexpr (ALG == NM_ALG_TD0) boot = lds_f64(vb, q1.y);
  else boot = extremum<ALG, K>(v);
  // exp of the previous step's new value: an independent chain the scheduler overlaps with this step
  double wprev = 0.0;
  if (online) wprev = exp_core(st.pend_arg);
  if (q0.w != kNone) {  // observed step: log p = beta * V[obs] - log(sum_k W[cand_k])
    const double nhz = __hiloint2double((int)q3.w, (int)q3.z);
    double s0 = __dmul_rn(nhz, wprev), s1 = 0.0;  // the n_hz candidates whose W is still in flight
#pragma unroll
    for (int k = 0; k < K; k += 2) {
      if ((uint32_t)k < Kw) s0 = __dadd_rn(s0, w[k]);
      if (k + 1 < K && (uint32_t)(k + 1) < Kw) s1 = __dadd_rn(s1, w[k + 1]);
    }
    const double sum = __dadd_rn(s0, s1);
    st.acc_obs = fma(beta, lds_f64(vb, q0.w), st.acc_obs);
    const double pr = __dmul_rn(st.prod, sum);
    const int hi = __double2hiint(pr);
    st.eacc += ((hi >> 20) & 0x7ff) - 1023;
    st.prod = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(pr));
  }
  if (online && st.pend_off != kNone) *reinterpret_cast<double *>(vb + st.pend_off + 8) = wprev;
  const double rpe = __dsub_rn(__dadd_rn(r, __dmul_rn(gamma, boot)), vs);  // rl.py:185, 228, 257
  const double vnew = __dadd_rn(vs, __dmul_rn(alpha, rpe));                // rl.py:186-187, 258-259
  *ps = vnew;
  if (online) {
    const double arg = __dmul_rn(beta, vnew);
    st.pend_arg = arg;
    st.pend_off = q0.z;
    st.bad = st.bad || !(fabs(arg) <= 700.0);
  }
}

template <int ALG>
__device__ __forceinline__ void memo_step(const uint4 *__restrict__ rec, char *vb, const double alpha,
                                          const double beta, const double gamma, MemoState &st) {
  const uint4 q0 = rec[0], q1 = rec[1], q2 = rec[2], q3 = rec[3];
#define NM_CASE(KK) \
  case KK: memo_step_k<ALG, KK>(q0, q1, q2, q3, vb, alpha, beta, gamma, st); break;
  switch (q1.x & 0xffu) {
    NM_CASE(1) NM_CASE(2) NM_CASE(3) NM_CASE(4) NM_CASE(5) NM_CASE(6) NM_CASE(7) NM_CASE(8)
    default: break;  // K == 0 is rejected on the host for likelihood sweeps
  }
#undef NM_CASE
}

template <int ALG>
__global__ void __launch_bounds__(512, 1) nm_fit_memo_kernel(const FitArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Rings rg = carve_rings(smem_raw);
  double *VW = reinterpret_cast<double *>(smem_raw + kFixedBytes);  // [state][thread]{V, W}
  const int tid = threadIdx.x;
  const int nc = blockDim.x - 32;
  const int sess = blockIdx.y;
  const int64_t rec0 = a.offsets[sess];
  const int64_t n = a.offsets[sess + 1] - rec0;
  const int64_t n_on = a.online[sess];
  const int nchunks = (int)((n + kChunk - 1) / kChunk);
  if (tid == 0) init_rings(rg, nc >> 5);
  ExpLL::fill(smem_raw + kBarBytes + kRawBytes + kDecBytes, tid);
  __syncthreads();

  if (tid >= nc) {
    producer_loop<true>(rg, a.records + rec0, n, n_on, nchunks, (uint32_t)nc * 16u, tid - nc);
  } else {
    const int64_t p = (int64_t)blockIdx.x * nc + tid;
    const bool live = p < a.P;
    const double alpha = live ? a.alpha[p] : 0.0;
    const double beta = live ? a.beta[p] : 0.0;
    const double gamma = live ? a.gamma[p] : 0.0;
    double *col = VW + 2 * tid;
    init_v_column(a, col, 2 * nc, p, live);
    const int NV = a.num_states;
    MemoState st;
    st.acc_obs = 0.0;
    st.prod = 1.0;
    st.eacc = 0;
    st.pend_arg = 0.0;
    st.pend_off = kNone;
    st.bad = false;
    for (int s = 0; s < NV; ++s) {  // W = exp(beta * V) of the initial table
      const double arg = __dmul_rn(beta, col[s * 2 * nc]);
      st.bad = st.bad || !(fabs(arg) <= 700.0);
      col[s * 2 * nc + 1] = exp_core(arg);
    }
    char *vb = reinterpret_cast<char *>(col);
    const int lane = tid & 31;
    for (int c = 0; c < nchunks; ++c) {
      const int dslot = c % kDecStages;
      mbar_wait(&rg.full_dec[dslot], (uint32_t)((c / kDecStages) & 1));
      const uint4 *recs = rg.dec + dslot * (kChunk * 4);
      const int64_t first = (int64_t)c * kChunk;
      const int cnt = (int)(n - first < kChunk ? n - first : kChunk);
#pragma unroll 1
      for (int i = 0; i < cnt; ++i) memo_step<ALG>(recs + 4 * i, vb, alpha, beta, gamma, st);
      __syncwarp();
      if (lane == 0) mbar_arrive(&rg.empty_dec[dslot]);
    }
    if (live && a.ll_out) {
      // ll = sum beta*V[obs] - log(prod of denominators) = acc_obs - (log(mantissa) + eacc * ln2)
      const double lg = __dadd_rn(log(st.prod), __dmul_rn((double)st.eacc, 6.93147180559945309417e-01));
      const double ll = __dsub_rn(st.acc_obs, lg);
      a.ll_out[(int64_t)sess * a.P + p] = st.bad ? __longlong_as_double(0x7ff8000000000000LL) : ll;
    }
  }
  if (a.v_out) {
    __syncthreads();
    write_v_out(a, VW, nc, 2 * nc, 2, sess);
  }
}
