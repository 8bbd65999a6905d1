// Parameter-grid sweeps of the conditioning recurrence (V-table learning + softmax log-likelihood).
//
// Mapping (DESIGN.md "fit kernel"):
//   * one THREAD owns one parameter set (alpha, beta, gamma); its V-table lives in shared memory as
//     V[state][thread] (column per thread): every thread of the CTA walks the SAME observed stream, so
//     at each step all lanes touch the same state row -> conflict-free LDS/STS, no divergence;
//   * the step stream (32-byte records, nm_step_rec) is staged global->shared by the TMA engine
//     (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) in a 4-deep ring of 64-record chunks issued
//     by one elected thread; consumers read records with broadcast LDS.128;
//   * grid = (ceil(P / threads), S): CTA (b, j) runs parameter block b on session j;
//   * all arithmetic FP64; the V recurrence uses explicit _rn intrinsics (no FMA contraction) so that
//     V-tables are bit-identical to the reference's numpy scalars (rl.py:184-187, 227-230, 257-259).
//
// Reference semantics restated here:
//   TD0 / PositiveConditioning / NegativeConditioning .update      simulation/rl.py:176-194,219-237,256-262
//   ConditioningExperiment.run loop + ShuffledReplays passes        simulation/rl.py:88-102, 419-424
//   softmax of BehaviouralModel.decision_making                    simulation/exploration_model.py:84-93
#include <cuda_runtime.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "nm_common.h"

namespace {

constexpr int kChunk = 64;   // records per staging chunk
constexpr int kStages = 4;   // ring depth
constexpr int kRecBytes = (int)sizeof(nm_step_rec);
constexpr int kBarBytes = 128;
constexpr int kStageBytes = kStages * kChunk * kRecBytes;
constexpr int kMaxSmem = 232448;  // 227 KB opt-in limit per CTA on sm_100

static_assert(sizeof(nm_step_rec) == 32, "nm_step_rec must be 32 bytes");

struct FitArgs {
  const nm_step_rec *records;
  const int64_t *offsets;
  const int64_t *online;
  const double *alpha, *beta, *gamma;
  int64_t P;
  int num_states, n_tiles;
  const int32_t *tile_of_state, *state_of_tile;
  const double *v_init;
  int v_init_per_unit;
  double *ll_out, *v_out;
};

// ---- PTX helpers: mbarrier + 1-D TMA bulk copy ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "NM_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra NM_DONE;\n"
      "bra NM_WAIT;\n"
      "NM_DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- one step of the recurrence ---------------------------------------------------------------
// w0 = {r.lo, r.hi, s | s1 << 16, ncand | obs << 8}, w1 = cand[0..7] as 4 x (u16, u16)
template <int ALG, bool LL>
__device__ __forceinline__ void fit_step(const uint4 w0, const uint4 w1, double *__restrict__ Vcol, const int nt,
                                         const double alpha, const double beta, const double gamma, double &ll) {
  const double r = __hiloint2double((int)w0.y, (int)w0.x);
  const int s = (int)(w0.z & 0xffffu);
  const int s1 = (int)(w0.z >> 16);
  const int K = (int)(w0.w & 0xffu);
  const int obs = (int)((w0.w >> 8) & 0xffu);

  double v[NM_MAX_CAND];
  if (LL || ALG != NM_ALG_TD0) {
    const uint32_t cw[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int k = 0; k < NM_MAX_CAND; ++k) {
      const int c = (int)((k & 1) ? (cw[k >> 1] >> 16) : (cw[k >> 1] & 0xffffu));
      v[k] = (k < K) ? Vcol[c * nt] : 0.0;
    }
  }

  if (LL) {
    if (obs != NM_OBS_NONE) {  // warp-uniform
      // val = beta * V[cand]; val -= max(val); p = exp(val) / sum(exp(val))   exploration_model.py:87-93
      double z[NM_MAX_CAND];
      double m = -INFINITY;
      double zobs = 0.0;
#pragma unroll
      for (int k = 0; k < NM_MAX_CAND; ++k) {
        z[k] = __dmul_rn(beta, v[k]);
        if (k < K) m = fmax(m, z[k]);
        if (k == obs) zobs = z[k];
      }
      double e[NM_MAX_CAND];
#pragma unroll
      for (int k = 0; k < NM_MAX_CAND; ++k) e[k] = (k < K) ? exp(__dsub_rn(z[k], m)) : 0.0;
      double sum;
      if (K == NM_MAX_CAND) {  // numpy pairwise sum for n == 8
        sum = __dadd_rn(__dadd_rn(__dadd_rn(e[0], e[1]), __dadd_rn(e[2], e[3])),
                        __dadd_rn(__dadd_rn(e[4], e[5]), __dadd_rn(e[6], e[7])));
      } else {  // n < 8: sequential from 0.0
        sum = 0.0;
#pragma unroll
        for (int k = 0; k < NM_MAX_CAND - 1; ++k)
          if (k < K) sum = __dadd_rn(sum, e[k]);
      }
      ll = __dadd_rn(ll, __dsub_rn(__dsub_rn(zobs, m), log(sum)));
    }
  }

  // value update
  double boot;
  if (ALG == NM_ALG_TD0) {
    boot = Vcol[s1 * nt];
  } else {
    boot = v[0];
#pragma unroll
    for (int k = 1; k < NM_MAX_CAND; ++k)
      if (k < K) boot = (ALG == NM_ALG_POSITIVE) ? fmax(boot, v[k]) : fmin(boot, v[k]);
  }
  const double vs = Vcol[s * nt];
  // rpe = r + gamma * boot - V[s];  V[s] += alpha * rpe       (left-to-right, no contraction)
  const double rpe = __dsub_rn(__dadd_rn(r, __dmul_rn(gamma, boot)), vs);
  Vcol[s * nt] = __dadd_rn(vs, __dmul_rn(alpha, rpe));
}

template <int ALG, bool LL>
__global__ void __launch_bounds__(512, 1) nm_fit_kernel(const FitArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
  uint4 *stage = reinterpret_cast<uint4 *>(smem_raw + kBarBytes);
  double *V = reinterpret_cast<double *>(smem_raw + kBarBytes + kStageBytes);

  const int tid = threadIdx.x, nt = blockDim.x;
  const int sess = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * nt + tid;
  const bool live = p < a.P;
  const int64_t rec0 = a.offsets[sess];
  const int64_t n = a.offsets[sess + 1] - rec0;
  const int64_t n_on = a.online[sess];
  const int nchunks = (int)((n + kChunk - 1) / kChunk);
  const nm_step_rec *src = a.records + rec0;

  auto issue = [&](int c) {  // elected thread: arm the barrier, launch the bulk copy of chunk c
    const int stg = c % kStages;
    const int64_t first = (int64_t)c * kChunk;
    const uint32_t bytes = (uint32_t)((n - first < kChunk ? n - first : kChunk) * kRecBytes);
    mbar_expect_tx(&full[stg], bytes);
    tma_bulk_g2s(stage + stg * (kChunk * 2), src + first, bytes, &full[stg]);
  };

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0)
    for (int c = 0; c < kStages && c < nchunks; ++c) issue(c);

  const double alpha = live ? a.alpha[p] : 0.0;
  const double beta = (LL && live) ? a.beta[p] : 0.0;
  const double gamma = live ? a.gamma[p] : 0.0;
  double *Vcol = V + tid;
  const int NV = a.num_states;
  if (a.v_init == nullptr) {
    for (int s = 0; s < NV; ++s) Vcol[s * nt] = 0.0;
  } else if (!a.v_init_per_unit) {
    for (int s = 0; s < NV; ++s) Vcol[s * nt] = a.v_init[a.tile_of_state[s]];
  } else {
    for (int s = 0; s < NV; ++s) Vcol[s * nt] = live ? a.v_init[p * a.n_tiles + a.tile_of_state[s]] : 0.0;
  }

  double ll = 0.0;
  for (int c = 0; c < nchunks; ++c) {
    const int stg = c % kStages;
    mbar_wait(&full[stg], (uint32_t)((c / kStages) & 1));
    const uint4 *recs = stage + stg * (kChunk * 2);
    const int64_t first = (int64_t)c * kChunk;
    const int cnt = (int)(n - first < kChunk ? n - first : kChunk);
    int cnt_on = (int)(n_on - first);
    cnt_on = cnt_on < 0 ? 0 : (cnt_on > cnt ? cnt : cnt_on);
    int i = 0;
    if (LL) {
      for (; i < cnt_on; ++i) fit_step<ALG, true>(recs[2 * i], recs[2 * i + 1], Vcol, nt, alpha, beta, gamma, ll);
    }
    for (; i < cnt; ++i) fit_step<ALG, false>(recs[2 * i], recs[2 * i + 1], Vcol, nt, alpha, beta, gamma, ll);
    __syncthreads();  // every warp is done with this stage -> it can be refilled
    if (tid == 0 && c + kStages < nchunks) issue(c + kStages);
  }

  if (LL && live && a.ll_out) a.ll_out[(int64_t)sess * a.P + p] = ll;

  if (a.v_out) {  // dense [S, P, X*Y] tables, written coalesced through a transposed read of V
    __syncthreads();
    const int64_t p0 = (int64_t)blockIdx.x * nt;
    const int64_t np = (a.P - p0 < nt) ? (a.P - p0) : nt;
    const int nT = a.n_tiles;
    double *dst = a.v_out + ((int64_t)sess * a.P + p0) * nT;
    for (int64_t i = tid; i < np * nT; i += nt) {
      const int u = (int)(i / nT), tile = (int)(i - (int64_t)u * nT);
      const int st = a.state_of_tile[tile];
      double val;
      if (st >= 0) val = V[st * nt + u];
      else if (a.v_init == nullptr) val = 0.0;
      else val = a.v_init_per_unit ? a.v_init[(p0 + u) * nT + tile] : a.v_init[tile];
      dst[i] = val;
    }
  }
}

#include "nm_fit_v2.cuh"

// ---- launch -----------------------------------------------------------------------------------
struct LaunchPlan {
  int nt;       // parameter sets (V columns) per CTA
  int threads;  // launched threads per CTA (generation 2 adds one producer warp)
  int gen;
  size_t smem;
};

int fit_generation() {
  const char *env = getenv("NM_FIT_GEN");
  return (env && atoi(env) == 1) ? 1 : 2;
}

int plan_launch(int num_states, int64_t P, int S, int device, LaunchPlan *plan) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int gen = fit_generation();
  plan->gen = gen;
  const int fixed = gen == 2 ? k2FixedBytes : kBarBytes + kStageBytes;
  int max_nt = (int)(((int64_t)kMaxSmem - fixed) / ((int64_t)num_states * 8));
  max_nt = (max_nt / 32) * 32;
  if (max_nt > (gen == 2 ? 480 : 512)) max_nt = gen == 2 ? 480 : 512;
  if (max_nt < 32)
    return nm_fail(NM_ERR_INVALID, "maze with %d visitable tiles does not fit the shared-memory V layout", num_states);
  const char *env = getenv("NM_FIT_NT");
  if (env) {
    int v = atoi(env);
    if (v >= 32 && v <= max_nt && v % 32 == 0) {
      plan->nt = v;
      plan->threads = v + (gen == 2 ? 32 : 0);
      plan->smem = (size_t)fixed + (size_t)num_states * 8 * v;
      return NM_OK;
    }
  }
  const double sat = 256.0;  // resident threads per SM below which the FP64 pipe is latency-starved
  double best = 1e300;
  int best_nt = 32;
  for (int nt = 32; nt <= max_nt; nt += 32) {
    const size_t smem = (size_t)fixed + (size_t)num_states * 8 * nt;
    int r = (int)(kMaxSmem / (smem + 1024));  // 1 KB per-CTA reservation
    if (r < 1) r = 1;
    if (r > 2048 / nt) r = 2048 / nt;
    if (r > 32) r = 32;
    const int64_t ctas = (int64_t)S * ((P + nt - 1) / nt);
    const int64_t slots = (int64_t)sms * r;
    const int64_t waves = (ctas + slots - 1) / slots;
    int64_t r_eff = (ctas + sms - 1) / sms;
    if (r_eff > r) r_eff = r;
    double load = (double)(r_eff * nt);
    if (load < sat) load = sat;
    const double cost = (double)waves * load;
    if (cost < best * 0.999) {
      best = cost;
      best_nt = nt;
    }
  }
  plan->nt = best_nt;
  plan->threads = best_nt + (gen == 2 ? 32 : 0);
  plan->smem = (size_t)fixed + (size_t)num_states * 8 * best_nt;
  return NM_OK;
}

template <int ALG, bool LL>
int launch_fit(const FitArgs &a, int S, const LaunchPlan &plan, cudaStream_t cs) {
  auto kernel = plan.gen == 2 ? nm_fit2_kernel<ALG, LL> : nm_fit_kernel<ALG, LL>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const dim3 grid((unsigned)((a.P + plan.nt - 1) / plan.nt), (unsigned)S, 1);
  kernel<<<grid, plan.threads, plan.smem, cs>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_fit_kernel launch: %s", cudaGetErrorString(e));
  return NM_OK;
}

int fit_sweep(int alg, bool with_ll, const nm_maze *mz, const nm_streams *st, const double *d_alpha,
              const double *d_beta, const double *d_gamma, int64_t P, const double *d_v_init, int v_init_per_unit,
              double *d_ll_out, double *d_v_out, void *stream) {
  NM_CHECK_ARG(mz && st, "sweep: NULL maze / streams handle");
  NM_CHECK_ARG(alg == NM_ALG_TD0 || alg == NM_ALG_POSITIVE || alg == NM_ALG_NEGATIVE, "sweep: unknown algorithm %d", alg);
  NM_CHECK_ARG(mz->device >= 0 && mz->device == st->device, "sweep: maze and streams must live on the same CUDA device");
  NM_CHECK_ARG(st->num_states == mz->num_states, "sweep: streams were built for %d states, maze has %d",
               st->num_states, mz->num_states);
  NM_CHECK_ARG(P >= 0 && (P == 0 || (d_alpha && d_gamma)), "sweep: NULL parameter array");
  NM_CHECK_ARG(!with_ll || P == 0 || (d_beta && d_ll_out), "loglik sweep: NULL beta / ll_out");
  NM_CHECK_ARG(with_ll || d_v_out, "vtable sweep: NULL v_out");
  NM_CHECK_ARG(st->S <= 65535, "sweep: at most 65535 sessions per call");
  NM_CHECK_ARG(st->total == 0 || st->min_ncand > 0 || (alg == NM_ALG_TD0 && !with_ll),
               "sweep: the streams hold steps without candidate next states; the look-ahead learners and the "
               "log-likelihood need them (np.max([]) raises in the reference)");
  if (P == 0) return NM_OK;
  if (cudaSetDevice(mz->device) != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaSetDevice(%d) failed", mz->device);
  LaunchPlan plan;
  int rc = plan_launch(mz->num_states, P, st->S, mz->device, &plan);
  if (rc != NM_OK) return rc;
  NM_CHECK_ARG((P + plan.nt - 1) / plan.nt <= 2147483647LL, "sweep: too many parameter sets");
  FitArgs a;
  a.records = st->d_records;
  a.offsets = st->d_offsets;
  a.online = st->d_online;
  a.alpha = d_alpha;
  a.beta = d_beta;
  a.gamma = d_gamma;
  a.P = P;
  a.num_states = mz->num_states;
  a.n_tiles = mz->X * mz->Y;
  a.tile_of_state = mz->d_tile_of_state;
  a.state_of_tile = mz->d_state_of_tile;
  a.v_init = d_v_init;
  a.v_init_per_unit = v_init_per_unit;
  a.ll_out = d_ll_out;
  a.v_out = d_v_out;
  cudaStream_t cs = (cudaStream_t)stream;
#define NM_DISPATCH(A)                                             \
  case A:                                                          \
    return with_ll ? launch_fit<A, true>(a, st->S, plan, cs) : launch_fit<A, false>(a, st->S, plan, cs);
  switch (alg) {
    NM_DISPATCH(NM_ALG_TD0)
    NM_DISPATCH(NM_ALG_POSITIVE)
    NM_DISPATCH(NM_ALG_NEGATIVE)
  }
#undef NM_DISPATCH
  return nm_fail(NM_ERR_INVALID, "unreachable");
}

// ---- argmin epilogue ---------------------------------------------------------------------------
struct Best {
  double val;
  long long idx;
};

__device__ __forceinline__ bool better(double v, long long i, double bv, long long bi) {
  if (bi < 0) return i >= 0;
  if (i < 0) return false;
  return v < bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(256) nm_nll_partial_kernel(const double *__restrict__ ll, int S, int64_t P,
                                                             double *__restrict__ nll_out,
                                                             Best *__restrict__ partial) {
  double bv = 0.0;
  long long bi = -1;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < S; ++s) acc = __dadd_rn(acc, ll[(int64_t)s * P + p]);
    const double nll = -acc;
    if (nll_out) nll_out[p] = nll;
    if (nll == nll && better(nll, p, bv, bi)) {
      bv = nll;
      bi = p;
    }
  }
  __shared__ Best red[256];
  red[threadIdx.x].val = bv;
  red[threadIdx.x].idx = bi;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      const Best o = red[threadIdx.x + off];
      if (better(o.val, o.idx, red[threadIdx.x].val, red[threadIdx.x].idx)) red[threadIdx.x] = o;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) nm_nll_final_kernel(const Best *__restrict__ partial, int n,
                                                           Best *__restrict__ out) {
  __shared__ Best red[256];
  Best b;
  b.val = 0.0;
  b.idx = -1;
  for (int i = threadIdx.x; i < n; i += 256)
    if (better(partial[i].val, partial[i].idx, b.val, b.idx)) b = partial[i];
  red[threadIdx.x] = b;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      const Best o = red[threadIdx.x + off];
      if (better(o.val, o.idx, red[threadIdx.x].val, red[threadIdx.x].idx)) red[threadIdx.x] = o;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (red[0].idx < 0) red[0].val = __longlong_as_double(0x7ff8000000000000LL);
    *out = red[0];
  }
}

// ---- FP64 issue-rate probe ----------------------------------------------------------------------
__global__ void __launch_bounds__(1024) nm_fp64_probe_kernel(int iters, double *sink) {
  double x0 = 1.0 + threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3, x4 = x0 + 4e-3, x5 = x0 + 5e-3,
         x6 = x0 + 6e-3, x7 = x0 + 7e-3;
  const double a = 0.999999, b = 1e-7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

}  // namespace

extern "C" int nm_vtable_sweep(int alg, const nm_maze *mz, const nm_streams *st, const double *d_alpha,
                               const double *d_gamma, int64_t P, const double *d_v_init, int v_init_per_unit,
                               double *d_v_out, void *stream) {
  return fit_sweep(alg, false, mz, st, d_alpha, nullptr, d_gamma, P, d_v_init, v_init_per_unit, nullptr, d_v_out,
                   stream);
}

extern "C" int nm_loglik_sweep(int alg, const nm_maze *mz, const nm_streams *st, const double *d_alpha,
                               const double *d_beta, const double *d_gamma, int64_t P, const double *d_v_init,
                               int v_init_per_unit, double *d_ll_out, double *d_v_out, void *stream) {
  return fit_sweep(alg, true, mz, st, d_alpha, d_beta, d_gamma, P, d_v_init, v_init_per_unit, d_ll_out, d_v_out,
                   stream);
}

extern "C" int nm_argmin_nll(const double *d_ll, int S, int64_t P, double *d_nll_out, void *d_best_out,
                             void *stream) {
  NM_CHECK_ARG(d_ll && d_best_out && S >= 1 && P >= 1, "nm_argmin_nll: bad argument");
  cudaStream_t cs = (cudaStream_t)stream;
  int blocks = (int)((P + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  // 16 KB of partials: one persistent scratch buffer per (device, stream), allocated on first use and
  // kept for the life of the process -- no allocator traffic on the hot path (stream-ordered pools hand
  // memory back to the OS at synchronisation points, which showed up as 20-500 ms stalls in bench e2e)
  int device = 0;
  cudaGetDevice(&device);
  Best *partial = nullptr;
  {
    static std::mutex mu;
    static std::map<std::pair<int, void *>, Best *> scratch;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(device, stream);
    auto it = scratch.find(key);
    if (it == scratch.end()) {
      Best *buf = nullptr;
      cudaError_t ea = cudaMalloc((void **)&buf, sizeof(Best) * 1024);
      if (ea != cudaSuccess) return nm_fail(NM_ERR_NOMEM, "nm_argmin_nll: cudaMalloc: %s", cudaGetErrorString(ea));
      it = scratch.emplace(key, buf).first;
    }
    partial = it->second;
  }
  nm_nll_partial_kernel<<<blocks, 256, 0, cs>>>(d_ll, S, P, d_nll_out, partial);
  nm_nll_final_kernel<<<1, 256, 0, cs>>>(partial, blocks, (Best *)d_best_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_argmin_nll launch: %s", cudaGetErrorString(e));
  return NM_OK;
}

extern "C" int nm_fp64_peak_probe(int blocks, int threads, int iters, double *d_sink, int64_t *h_fp64_instr,
                                  void *stream) {
  NM_CHECK_ARG(blocks > 0 && threads > 0 && threads <= 1024 && iters > 0 && d_sink, "nm_fp64_peak_probe: bad argument");
  nm_fp64_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, d_sink);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_fp64_probe launch: %s", cudaGetErrorString(e));
  if (h_fp64_instr) *h_fp64_instr = (int64_t)blocks * threads * (int64_t)iters * 8;
  return NM_OK;
}
