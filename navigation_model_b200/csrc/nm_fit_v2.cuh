// Fit kernel, generation 2: warp-specialised producer + decoded step ring + K-specialised step bodies.
//
// Included by nm_fit.cu inside its anonymous namespace (uses FitArgs and the PTX helpers defined there).
//
// What changed against generation 1 (profiles/r01_v1_summary.md: issue-bound at 73 % issue-slot
// utilisation with only 21 % of the instructions on the FP64 pipe):
//   * ONE producer warp per CTA owns the stream: it issues the TMA bulk copies (cp.async.bulk, UBLKCP)
//     of raw 32-byte records into a 2-deep raw ring, DECODES each landed chunk once (state ids ->
//     shared-memory byte offsets of the V rows, 1 record per lane) into a 4-deep ring of 64-byte decoded
//     records, and hands chunks to the consumer warps through mbarriers (full / empty).  No
//     __syncthreads in the step loop: consumer warps drift freely up to the ring depth.
//   * consumers read a decoded record with 4 broadcast LDS.128 and address V rows with one IADD each;
//     the step body is specialised on K (switch on the warp-uniform candidate count) so the candidate
//     loops carry no predicates or selects; the observed candidate's value is re-loaded by its row offset
//     instead of being selected from registers;
//   * exp() is a range-restricted FP64 polynomial for x <= 0 (the softmax subtracts the max): round to
//     nearest multiple of ln2 with the magic-constant trick, two-step Cody-Waite reduction, degree-13
//     Taylor/Horner, exponent add.  |rel err| < 2^-52.  Inputs below -708 flush to 0 (their share of a
//     sum whose largest term is 1 is < 1e-307).
//   * log(): the per-step log(sum_k e_k) is replaced by the log of a running PRODUCT of the sums, taken
//     once per chunk (32 steps: prod <= 8^32 = 2^96, no overflow):  sum_t log(s_t) = log(prod_t s_t).
//   * for the max-backup learner and beta >= 0, max_k(beta * v_k) = beta * max_k v_k (rounding is
//     monotonic), so the max needed by the softmax is the one the backup computes anyway.
// The V recurrence itself is untouched: same _rn operations in the same order -> bit-exact V-tables.

constexpr int k2Chunk = 32;      // records per chunk = one per producer lane
constexpr int k2RawStages = 2;   // raw (TMA landing) ring
constexpr int k2DecStages = 4;   // decoded ring
constexpr int k2RawBytes = k2RawStages * k2Chunk * 32;
constexpr int k2DecRecBytes = 64;
constexpr int k2DecBytes = k2DecStages * k2Chunk * k2DecRecBytes;
constexpr int k2BarBytes = 128;  // full_raw[2], full_dec[4], empty_dec[4]
constexpr int k2FixedBytes = k2BarBytes + k2RawBytes + k2DecBytes;

// decoded record, 64 bytes = 4 x uint4:
//   q0 = {r.lo, r.hi, off_s, off_obs}   q1 = {K, off_s1, off_c0, off_c1}   q2 = {off_c2..off_c5}   q3 = {off_c6, off_c7, -, -}
// offsets are byte offsets of V rows: state * (consumer threads) * 8; off_obs = 0xFFFFFFFF when unobserved.

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// exp(x) for x <= 0 (NaN propagates).  17 FP64 instructions + 3 integer.
__device__ __forceinline__ double exp_nonpos(double x) {
  const double kLog2e = 1.4426950408889634074;
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  const double kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10;
  const double t = fma(x, kLog2e, kMagic);
  const int k = __double2loint(t);
  const double n = __dsub_rn(t, kMagic);
  double r = fma(n, -kLn2Hi, x);
  r = fma(n, -kLn2Lo, r);
  double p = 1.0 / 6227020800.0;                 // 1/13!
  p = fma(p, r, 1.0 / 479001600.0);              // 1/12!
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double scaled = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
  return (x < -708.0) ? 0.0 : scaled;
}

template <int K>
__device__ __forceinline__ double np_sum_fixed(const double (&e)[NM_MAX_CAND]) {
  if (K == NM_MAX_CAND) {
    return __dadd_rn(__dadd_rn(__dadd_rn(e[0], e[1]), __dadd_rn(e[2], e[3])),
                     __dadd_rn(__dadd_rn(e[4], e[5]), __dadd_rn(e[6], e[7])));
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < K; ++k) s = __dadd_rn(s, e[k]);
  return s;
}

// One step with a compile-time candidate count.  `vb` = this thread's column base (bytes, shared window).
template <int ALG, bool LL, int K>
__device__ __forceinline__ void fit2_step_k(const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3,
                                            const char *__restrict__ vb, const double alpha, const double beta,
                                            const double gamma, double &acc_obs, double &prod) {
  const double r = __hiloint2double((int)q0.y, (int)q0.x);
  const uint32_t offs[NM_MAX_CAND] = {q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y};
  double v[NM_MAX_CAND];
  double vext = 0.0;  // max (Positive) / min (Negative) of the candidate values
  if (LL || ALG != NM_ALG_TD0) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = *reinterpret_cast<const double *>(vb + offs[k]);
    if (ALG != NM_ALG_TD0) {
      vext = v[0];
#pragma unroll
      for (int k = 1; k < K; ++k) vext = (ALG == NM_ALG_POSITIVE) ? fmax(vext, v[k]) : fmin(vext, v[k]);
    }
  }
  if (LL && q0.w != 0xFFFFFFFFu) {  // observed step (warp-uniform)
    double z[NM_MAX_CAND];
#pragma unroll
    for (int k = 0; k < K; ++k) z[k] = __dmul_rn(beta, v[k]);
    double m;
    if (ALG == NM_ALG_POSITIVE && beta >= 0.0) {
      m = __dmul_rn(beta, vext);  // == max_k fl(beta * v_k): rounding is monotonic
    } else if (ALG == NM_ALG_NEGATIVE && beta <= 0.0) {
      m = __dmul_rn(beta, vext);
    } else {
      m = z[0];
#pragma unroll
      for (int k = 1; k < K; ++k) m = fmax(m, z[k]);
    }
    double e[NM_MAX_CAND];
#pragma unroll
    for (int k = 0; k < K; ++k) e[k] = exp_nonpos(__dsub_rn(z[k], m));
    const double sum = np_sum_fixed<K>(e);
    const double zobs = __dmul_rn(beta, *reinterpret_cast<const double *>(vb + q0.w));
    acc_obs = __dadd_rn(acc_obs, __dsub_rn(zobs, m));
    prod = __dmul_rn(prod, sum);
  }
  double boot;
  if (ALG == NM_ALG_TD0) boot = *reinterpret_cast<const double *>(vb + q1.y);
  else boot = vext;
  double *ps = reinterpret_cast<double *>(const_cast<char *>(vb) + q0.z);
  const double vs = *ps;
  const double rpe = __dsub_rn(__dadd_rn(r, __dmul_rn(gamma, boot)), vs);  // rl.py:185, 228, 257
  *ps = __dadd_rn(vs, __dmul_rn(alpha, rpe));                               // rl.py:186-187, 258-259
}

template <int ALG, bool LL>
__device__ __forceinline__ void fit2_step(const uint4 *__restrict__ rec, const char *__restrict__ vb,
                                          const double alpha, const double beta, const double gamma, double &acc_obs,
                                          double &prod) {
  const uint4 q0 = rec[0], q1 = rec[1], q2 = rec[2], q3 = rec[3];
  if (!LL && ALG == NM_ALG_TD0) {  // no candidates needed
    fit2_step_k<ALG, LL, 0>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc_obs, prod);
    return;
  }
  switch (q1.x) {  // K, warp-uniform
    case 1: fit2_step_k<ALG, LL, 1>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc_obs, prod); break;
    case 2: fit2_step_k<ALG, LL, 2>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc_obs, prod); break;
    case 3: fit2_step_k<ALG, LL, 3>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc_obs, prod); break;
    case 4: fit2_step_k<ALG, LL, 4>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc_obs, prod); break;
    case 5: fit2_step_k<ALG, LL, 5>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc_obs, prod); break;
    case 6: fit2_step_k<ALG, LL, 6>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc_obs, prod); break;
    case 7: fit2_step_k<ALG, LL, 7>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc_obs, prod); break;
    case 8: fit2_step_k<ALG, LL, 8>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc_obs, prod); break;
    default:  // K == 0 can only happen for TD0 streams built without candidates
      if (ALG == NM_ALG_TD0) fit2_step_k<NM_ALG_TD0, false, 0>(q0, q1, q2, q3, vb, alpha, beta, gamma, acc_obs, prod);
      break;
  }
}

// blockDim.x = consumers + 32; the LAST warp is the producer.
template <int ALG, bool LL>
__global__ void __launch_bounds__(512, 1) nm_fit2_kernel(const FitArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *full_raw = reinterpret_cast<uint64_t *>(smem_raw);   // [k2RawStages]
  uint64_t *full_dec = full_raw + k2RawStages;                  // [k2DecStages]
  uint64_t *empty_dec = full_dec + k2DecStages;                 // [k2DecStages]
  uint4 *raw = reinterpret_cast<uint4 *>(smem_raw + k2BarBytes);
  uint4 *dec = reinterpret_cast<uint4 *>(smem_raw + k2BarBytes + k2RawBytes);
  double *V = reinterpret_cast<double *>(smem_raw + k2FixedBytes);

  const int tid = threadIdx.x;
  const int nc = blockDim.x - 32;  // consumer threads = V columns
  const int n_cwarps = nc >> 5;
  const int sess = blockIdx.y;
  const int64_t rec0 = a.offsets[sess];
  const int64_t n = a.offsets[sess + 1] - rec0;
  const int64_t n_on = a.online[sess];
  const int nchunks = (int)((n + k2Chunk - 1) / k2Chunk);
  const nm_step_rec *src = a.records + rec0;

  if (tid == 0) {
    for (int i = 0; i < k2RawStages; ++i) mbar_init(&full_raw[i], 1);
    for (int i = 0; i < k2DecStages; ++i) {
      mbar_init(&full_dec[i], 32);
      mbar_init(&empty_dec[i], n_cwarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= nc) {
    // ======================= producer warp: TMA + decode ==========================================
    const int lane = tid - nc;
    auto issue = [&](int c) {
      const int slot = c % k2RawStages;
      const int64_t first = (int64_t)c * k2Chunk;
      const uint32_t bytes = (uint32_t)((n - first < k2Chunk ? n - first : k2Chunk) * 32);
      mbar_expect_tx(&full_raw[slot], bytes);
      tma_bulk_g2s(raw + slot * (k2Chunk * 2), src + first, bytes, &full_raw[slot]);
    };
    if (lane == 0)
      for (int c = 0; c < k2RawStages && c < nchunks; ++c) issue(c);
    const uint32_t row = (uint32_t)nc * 8u;  // bytes per V row
    for (int c = 0; c < nchunks; ++c) {
      const int rslot = c % k2RawStages, dslot = c % k2DecStages;
      if (c >= k2DecStages) mbar_wait(&empty_dec[dslot], (uint32_t)(((c / k2DecStages) - 1) & 1));
      mbar_wait(&full_raw[rslot], (uint32_t)((c / k2RawStages) & 1));
      const int cnt = (int)(n - (int64_t)c * k2Chunk < k2Chunk ? n - (int64_t)c * k2Chunk : k2Chunk);
      if (lane < cnt) {
        const uint4 w0 = raw[(rslot * k2Chunk + lane) * 2], w1 = raw[(rslot * k2Chunk + lane) * 2 + 1];
        const uint32_t s = w0.z & 0xffffu, s1 = w0.z >> 16, K = w0.w & 0xffu, obs = (w0.w >> 8) & 0xffu;
        const uint32_t cw[4] = {w1.x, w1.y, w1.z, w1.w};
        uint32_t oc[NM_MAX_CAND];
#pragma unroll
        for (int k = 0; k < NM_MAX_CAND; ++k) oc[k] = ((k & 1) ? (cw[k >> 1] >> 16) : (cw[k >> 1] & 0xffffu)) * row;
        uint32_t oobs = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < NM_MAX_CAND; ++k)
          if ((uint32_t)k == obs) oobs = oc[k];
        const bool online = ((int64_t)c * k2Chunk + lane) < n_on;
        if (!online) oobs = 0xFFFFFFFFu;  // replay passes update V only
        uint4 *d = dec + (dslot * k2Chunk + lane) * 4;
        d[0] = make_uint4(w0.x, w0.y, s * row, oobs);
        d[1] = make_uint4(K, s1 * row, oc[0], oc[1]);
        d[2] = make_uint4(oc[2], oc[3], oc[4], oc[5]);
        d[3] = make_uint4(oc[6], oc[7], 0u, 0u);
      }
      mbar_arrive(&full_dec[dslot]);  // every lane: releases its own stores (count = 32)
      __syncwarp();                   // all lanes have read the raw slot before it is refilled
      if (lane == 0 && c + k2RawStages < nchunks) issue(c + k2RawStages);
    }
  } else {
    // ======================= consumer warps ==========================================================
    const int64_t p = (int64_t)blockIdx.x * nc + tid;
    const bool live = p < a.P;
    const double alpha = live ? a.alpha[p] : 0.0;
    const double beta = (LL && live) ? a.beta[p] : 0.0;
    const double gamma = live ? a.gamma[p] : 0.0;
    double *Vcol = V + tid;
    const int NV = a.num_states;
    if (a.v_init == nullptr) {
      for (int s = 0; s < NV; ++s) Vcol[s * nc] = 0.0;
    } else if (!a.v_init_per_unit) {
      for (int s = 0; s < NV; ++s) Vcol[s * nc] = a.v_init[a.tile_of_state[s]];
    } else {
      for (int s = 0; s < NV; ++s) Vcol[s * nc] = live ? a.v_init[p * a.n_tiles + a.tile_of_state[s]] : 0.0;
    }
    const char *vb = reinterpret_cast<const char *>(Vcol);
    double ll = 0.0, acc_obs = 0.0, prod = 1.0;
    const int lane = tid & 31;
    for (int c = 0; c < nchunks; ++c) {
      const int dslot = c % k2DecStages;
      mbar_wait(&full_dec[dslot], (uint32_t)((c / k2DecStages) & 1));
      const uint4 *recs = dec + dslot * (k2Chunk * 4);
      const int cnt = (int)(n - (int64_t)c * k2Chunk < k2Chunk ? n - (int64_t)c * k2Chunk : k2Chunk);
#pragma unroll 1
      for (int i = 0; i < cnt; ++i) fit2_step<ALG, LL>(recs + 4 * i, vb, alpha, beta, gamma, acc_obs, prod);
      if (LL) {  // fold the chunk: ll += sum(zobs - m) - log(prod of the softmax denominators)
        ll = __dadd_rn(ll, __dsub_rn(acc_obs, log(prod)));
        acc_obs = 0.0;
        prod = 1.0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_dec[dslot]);
    }
    if (LL && live && a.ll_out) a.ll_out[(int64_t)sess * a.P + p] = ll;
  }

  if (a.v_out) {  // dense [S, P, X*Y] tables, written coalesced through a transposed read of V
    __syncthreads();
    const int64_t p0 = (int64_t)blockIdx.x * nc;
    const int64_t np = (a.P - p0 < nc) ? (a.P - p0) : nc;
    const int nT = a.n_tiles;
    double *dst = a.v_out + ((int64_t)sess * a.P + p0) * nT;
    for (int64_t i = tid; i < np * nT; i += blockDim.x) {
      const int u = (int)(i / nT), tile = (int)(i - (int64_t)u * nT);
      const int st = a.state_of_tile[tile];
      double val;
      if (st >= 0) val = V[st * nc + u];
      else if (a.v_init == nullptr) val = 0.0;
      else val = a.v_init_per_unit ? a.v_init[(p0 + u) * nT + tile] : a.v_init[tile];
      dst[i] = val;
    }
  }
}
