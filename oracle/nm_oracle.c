/*
 * nm_oracle.c -- plain-C CPU restatement of the hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker / the timed CPU baseline.  The product never links it.
 *
 * It follows the reference (paths relative to /root/reference/src/navigation_model):
 *   orc_fit      simulation/rl.py:88-102 (online loop), :257-259 (TD0), :184-187 (Positive),
 *                :227-230 (Negative), :419-424 (shuffled replay passes); the log-likelihood term is the
 *                softmax of simulation/exploration_model.py:84-93 in log space (EXTENSION: parity
 *                unpinned -- the reference has no likelihood).
 * Inputs are the plain transition arrays produced by oracle/np_oracle.py:extract_transitions (its own
 * format, independent of the product's nm_step_rec).  Pinned against the reference's outputs by
 * tests/test_oracle_golden.py (fixtures: tests/golden/).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp; no FMA contraction so that every
 * operation rounds like the reference's numpy scalar arithmetic).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_TD0 0
#define ORC_POSITIVE 1
#define ORC_NEGATIVE 2
#define ORC_OBS_NONE 255

int orc_version(void) { return 1; }

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* numpy.sum of n <= 8 contiguous doubles: sequential from 0.0 for n < 8, pairwise tree for n == 8 */
static double np_sum8(const double *e, int n) {
  if (n < 8) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += e[i];
    return acc;
  }
  return ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
}

static inline void value_update(int alg, double *V, int s, int s1, double r, const int32_t *cand, int K,
                                double alpha, double gamma) {
  double boot;
  if (alg == ORC_TD0) {
    boot = V[s1]; /* rl.py:257 */
  } else {
    boot = V[cand[0]];
    for (int k = 1; k < K; ++k) {
      const double v = V[cand[k]];
      if (alg == ORC_POSITIVE ? (v > boot) : (v < boot)) boot = v; /* np.max / np.min, rl.py:184, 227 */
    }
  }
  const double rpe = r + gamma * boot - V[s]; /* rl.py:185, 228, 257 */
  V[s] += alpha * rpe;                        /* rl.py:186-187, 258-259 */
}

/* One parameter set over one session; returns the log-likelihood (0 when with_ll == 0). */
static double run_one(int alg, int with_ll, const int32_t *s, const int32_t *s1, const double *r,
                      const int32_t *ncand, const int32_t *cand, const int32_t *obs, int64_t n_online,
                      const int64_t *replay, int64_t n_replay, double alpha, double beta, double gamma,
                      double *V) {
  double ll = 0.0;
  for (int64_t i = 0; i < n_online; ++i) {
    const int32_t *c = cand + 8 * i;
    const int K = ncand[i];
    if (with_ll && obs[i] != ORC_OBS_NONE) {
      double val[8] = {0}, e[8] = {0}, m;
      for (int k = 0; k < K; ++k) val[k] = beta * V[c[k]]; /* exploration_model.py:87 */
      m = val[0];
      for (int k = 1; k < K; ++k)
        if (val[k] > m) m = val[k];
      for (int k = 0; k < K; ++k) { /* :92-93 */
        val[k] = val[k] - m;
        e[k] = exp(val[k]);
      }
      ll = ll + (val[obs[i]] - log(np_sum8(e, K)));
    }
    value_update(alg, V, s[i], s1[i], r[i], c, K, alpha, gamma);
  }
  for (int64_t j = 0; j < n_replay; ++j) { /* rl.py:421-424 */
    const int64_t i = replay[j];
    value_update(alg, V, s[i], s1[i], r[i], cand + 8 * i, ncand[i], alpha, gamma);
  }
  return ll;
}

/*
 * P parameter sets over one session.
 *   s, s1, ncand, obs: int32[n]; r: double[n]; cand: int32[n, 8]; replay: int64[n_replay] indices
 *   v0: double[n_states] or NULL (zeros); V_out: double[P, n_states] or NULL; ll_out: double[P] or NULL
 *   threads <= 0: all OpenMP threads.
 */
int orc_fit(int alg, int with_ll, const int32_t *s, const int32_t *s1, const double *r, const int32_t *ncand,
            const int32_t *cand, const int32_t *obs, int64_t n_online, const int64_t *replay,
            int64_t n_replay, const double *alpha, const double *beta, const double *gamma, int64_t P,
            int n_states, const double *v0, double *V_out, double *ll_out, int threads) {
  if (alg < 0 || alg > 2 || n_states <= 0 || P < 0) return -1;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  int err = 0;
#pragma omp parallel num_threads(threads)
  {
    double *V = (double *)malloc(sizeof(double) * (size_t)n_states);
    if (!V) {
#pragma omp atomic write
      err = 1;
    } else {
#pragma omp for schedule(static)
      for (int64_t p = 0; p < P; ++p) {
        if (v0) memcpy(V, v0, sizeof(double) * (size_t)n_states);
        else memset(V, 0, sizeof(double) * (size_t)n_states);
        const double ll = run_one(alg, with_ll, s, s1, r, ncand, cand, obs, n_online, replay, n_replay,
                                  alpha[p], with_ll ? beta[p] : 0.0, gamma[p], V);
        if (ll_out) ll_out[p] = ll;
        if (V_out) memcpy(V_out + (size_t)p * n_states, V, sizeof(double) * (size_t)n_states);
      }
      free(V);
    }
  }
  return err ? -2 : 0;
}
