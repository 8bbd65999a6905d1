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

/* =====================================================================================================
 * orc_simulate -- forward simulation of N independent mice with replayed uniform draws.
 *
 * Plain-C restatement, operation for operation, of (paths relative to /root/reference/src):
 *   StandardExperiment.run                 navigation_model/simulation/experiment.py:33-51
 *   Mouse.get_endpoints / move_to          _navigation_model/mouse.cpp:64-87 (libm cos / sin / atan2, Eigen isApprox)
 *   Maze.cont2disc / is_visitable /
 *   is_movement_possible_cont              navigation_model/simulation/maze.py:380-395, 488-501, 563-611
 *   BehaviouralModel.decision_making       navigation_model/simulation/exploration_model.py:68-109
 *   the five components                    navigation_model/simulation/behavioural_components.py:315-544
 *   Generator.choice(a, p=p)               numpy: cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, 'right')
 * It is the compiled twin of oracle/np_oracle.py:simulate (same structure, ~200x faster) so that the
 * simulation kernel can be checked on 10^6 agent-steps in seconds.  Pinned bit for bit against histories
 * the unmodified reference produced (tests/golden/sim_golden.npz, random_mazes_golden.npz, grid_golden.npz;
 * tests/test_oracle_sim_golden.py).  `exp` is glibc's where the reference calls numpy's: the two differ by
 * 1 ulp on a few per cent of the arguments, which moves a probability in its last bit and can only change a
 * choice when a draw falls within ~1e-16 of a cdf step (none does in the fixtures).
 * ===================================================================================================== */

/* Python's float // float (CPython floatobject.c:float_divmod), then int() */
static int py_floordiv_int(double vx, double wx) {
  double mod = fmod(vx, wx);
  double div = (vx - mod) / wx;
  if (mod != 0.0) {
    if ((wx < 0) != (mod < 0)) {
      mod += wx;
      div -= 1.0;
    }
  }
  double floordiv;
  if (div != 0.0) {
    floordiv = floor(div);
    if (div - floordiv > 0.5) floordiv += 1.0;
  } else {
    floordiv = copysign(0.0, vx / wx);
  }
  return (int)floordiv;
}

typedef struct {
  const uint8_t *tiles; /* [X*Y], x-major; 0 = VOID */
  int X, Y;
  double st;
} orc_maze;

static int mz_visitable(const orc_maze *m, int x, int y) { /* maze.py:499-501 */
  if (x < 0 || x >= m->X || y < 0 || y >= m->Y) return 0;
  return m->tiles[x * m->Y + y] != 0;
}
static int mz_visitable_cont(const orc_maze *m, double x, double y) {
  return mz_visitable(m, py_floordiv_int(x, m->st), py_floordiv_int(y, m->st));
}

static int mz_move_possible(const orc_maze *m, double x1, double y1, double x2, double y2) { /* maze.py:563-611 */
  const double st = m->st;
  if (!mz_visitable_cont(m, x1, y1) || !mz_visitable_cont(m, x2, y2)) return 0;
  if (x1 != x2) {
    if (x1 > x2) {
      double t = x1; x1 = x2; x2 = t;
      t = y1; y1 = y2; y2 = t;
    }
    const double mm = (y2 - y1) / (x2 - x1);
    const int hi = py_floordiv_int(x2, st);
    for (int dx = py_floordiv_int(x1, st); dx < hi; ++dx) {
      const double x = (dx + 1) * st;
      const double y = mm * (x - x1) + y1;
      if (!mz_visitable_cont(m, x - st, y)) return 0;
      if (!mz_visitable_cont(m, x, y)) return 0;
    }
  }
  if (y1 != y2) {
    if (y1 > y2) {
      double t = x1; x1 = x2; x2 = t;
      t = y1; y1 = y2; y2 = t;
    }
    const double mm = (x2 - x1) / (y2 - y1);
    const int hi = py_floordiv_int(y2, st);
    for (int dy = py_floordiv_int(y1, st); dy < hi; ++dy) {
      const double y = (dy + 1) * st;
      const double x = mm * (y - y1) + x1;
      if (!mz_visitable_cont(m, x, y - st)) return 0;
      if (!mz_visitable_cont(m, x, y)) return 0;
    }
  }
  return 1;
}

#define ORC_COMP_ANXIETY 1
#define ORC_COMP_BCOST 2
#define ORC_COMP_EXPERIENCE 3
#define ORC_COMP_BPERSISTENCE 4
#define ORC_COMP_VTABLE 5

/*
 * N mice x T steps.  Per-mouse parameter rows (NULL when the component is absent):
 *   anxiety[N,5] = W_anx,p1,p2,p3,p4   bcost_w[N], bcost_table[N,A]   experience[N,6] = W_exp,xi,kappa_home,
 *   kappa_explo,theta_explo,theta_home   bpers[N,4] = W_bper,bp,Wm,Wnm   vtable_w[N], vtable[X*Y] (normalised, shared)
 *   comp_kind[n_comp]: evaluation order (model order).  init_pose[N,3], init_disc[N,2], beta[N], draws[N,T].
 * Outputs (any may be NULL): hist_pos[N,T+1,2], hist_ori[N,T+1], choice[N,T] (action index, 255 after a stop),
 *   occupancy[N,X*Y] (visits of history[0..T]), status[N] (0 or 1 + the step without a reachable endpoint),
 *   final_pose[N,3].
 */
int orc_simulate(const uint8_t *tiles, int X, int Y, double st, const double *actions, const uint8_t *zero_act, int A,
                 int64_t N, int T, const int *comp_kind, int n_comp, const double *beta, const double *anxiety,
                 const double *bcost_w, const double *bcost_table, const double *experience, const double *bpers,
                 const double *vtable_w, const double *vtable, const double *init_pose, const double *init_disc,
                 const double *draws, double *hist_pos, double *hist_ori, uint8_t *choice, int32_t *occupancy,
                 int32_t *status, double *final_pose, int threads) {
  if (A < 1 || A > 8 || X <= 0 || Y <= 0 || N < 0 || T < 0) return -1;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  const orc_maze mz = {tiles, X, Y, st};
  int err = 0;
#pragma omp parallel num_threads(threads)
  {
    double *exp_map = (double *)malloc(sizeof(double) * (size_t)X * Y); /* behavioural_components.py:445 */
    if (!exp_map) {
#pragma omp atomic write
      err = 1;
    } else {
#pragma omp for schedule(dynamic, 4)
      for (int64_t n = 0; n < N; ++n) {
        double px = init_pose[3 * n], py = init_pose[3 * n + 1], ori = init_pose[3 * n + 2];
        double prevx = init_disc[2 * n], prevy = init_disc[2 * n + 1]; /* exploration_model.py:47 */
        int moving = 0;                                                /* :48 (True when the mouse STAYED) */
        double fam = 0.0;
        int home = 0; /* behavioural_components.py:440, 447 */
        int st_code = 0;
        memset(exp_map, 0, sizeof(double) * (size_t)X * Y);
        if (occupancy) memset(occupancy + n * X * Y, 0, sizeof(int32_t) * (size_t)X * Y);
        if (hist_pos) {
          hist_pos[(n * (T + 1)) * 2] = px;
          hist_pos[(n * (T + 1)) * 2 + 1] = py;
        }
        if (hist_ori) hist_ori[n * (T + 1)] = ori;
        int t = 0;
        for (; t < T; ++t) {
          /* experiment.py:36-41 */
          const double c = cos(ori), s = sin(ori); /* mouse.cpp:66 */
          double ex[8], ey[8];
          int vis[8], idx[8], K = 0;
          for (int i = 0; i < A; ++i) {
            const double ax = actions[2 * i], ay = actions[2 * i + 1];
            ex[i] = (c * ax + (-s) * ay) + px; /* mouse.cpp:67-72: R(ori) * a + pos */
            ey[i] = (s * ax + c * ay) + py;
            vis[i] = mz_move_possible(&mz, px, py, ex[i], ey[i]);
            if (vis[i]) idx[K++] = i;
          }
          const int cx = py_floordiv_int(px, st), cy = py_floordiv_int(py, st);
          if (occupancy && cx >= 0 && cx < X && cy >= 0 && cy < Y) occupancy[n * X * Y + cx * Y + cy] += 1;
          if (K == 0) { /* np.max([]) raises, exploration_model.py:92 */
            st_code = t + 1;
            break;
          }
          int tx[8], ty[8];
          for (int k = 0; k < K; ++k) {
            tx[k] = py_floordiv_int(ex[idx[k]], st);
            ty[k] = py_floordiv_int(ey[idx[k]], st);
          }
          double val[8];
          for (int k = 0; k < K; ++k) val[k] = 0.0; /* exploration_model.py:84 */
          for (int ci = 0; ci < n_comp; ++ci) {
            switch (comp_kind[ci]) {
              case ORC_COMP_ANXIETY: { /* behavioural_components.py:332-341, 364-374 */
                const double *q = anxiety + 5 * n;
                const double w[5] = {-1.0, q[2] * q[1], q[1], q[3] * q[2] * q[1], q[4]};
                for (int k = 0; k < K; ++k) val[k] = val[k] + w[tiles[tx[k] * Y + ty[k]]] * q[0];
                break;
              }
              case ORC_COMP_BCOST: /* :408-413 */
                for (int k = 0; k < K; ++k)
                  val[k] = val[k] + (moving ? bcost_table[n * A + idx[k]] : 0.0) * bcost_w[n];
                break;
              case ORC_COMP_EXPERIENCE: { /* :449-479 */
                const double *q = experience + 6 * n;
                exp_map[cx * Y + cy] += 1;
                fam = (1 - q[1]) * fam + q[1] * exp_map[cx * Y + cy];
                if (home && fam >= q[4]) home = 0;
                else if (!home && fam < q[5]) home = 1;
                for (int k = 0; k < K; ++k) {
                  const double cnt = exp_map[tx[k] * Y + ty[k]];
                  const double v = !home ? exp(-q[3] * cnt) : 1.0 - exp(-q[2] * cnt);
                  val[k] = val[k] + v * q[0];
                }
                break;
              }
              case ORC_COMP_BPERSISTENCE: { /* :502-513 */
                const double *q = bpers + 4 * n;
                for (int k = 0; k < K; ++k) {
                  const int z = zero_act[idx[k]];
                  const double v = moving ? (z ? q[2] * q[1] : q[1]) : (z ? q[1] : q[3] * q[1]);
                  val[k] = val[k] + v * q[0];
                }
                break;
              }
              case ORC_COMP_VTABLE: /* :537-541 */
                for (int k = 0; k < K; ++k) val[k] = val[k] + vtable[tx[k] * Y + ty[k]] * vtable_w[n];
                break;
              default:
                break;
            }
          }
          double m = -INFINITY, e[8];
          for (int k = 0; k < K; ++k) {
            val[k] = beta[n] * val[k]; /* exploration_model.py:87 */
            if (val[k] > m) m = val[k];
          }
          for (int k = 0; k < K; ++k) e[k] = exp(val[k] - m); /* :92-93 */
          const double sum = np_sum8(e, K);
          double cdf[8], run = 0.0;
          for (int k = 0; k < K; ++k) { /* Generator.choice: cumsum(p) */
            const double p = e[k] / sum;
            run = k == 0 ? p : run + p;
            cdf[k] = run;
          }
          const double u = draws[n * T + t];
          int pick = K; /* searchsorted(cdf / cdf[-1], u, side='right') */
          for (int k = 0; k < K; ++k)
            if (cdf[k] / cdf[K - 1] > u) {
              pick = k;
              break;
            }
          if (pick == K) { /* numpy would index past the end and raise */
            st_code = -(t + 1);
            break;
          }
          const double nx = ex[idx[pick]], ny = ey[idx[pick]];
          moving = (prevx == nx) && (prevy == ny); /* :101 */
          prevx = nx;
          prevy = ny;
          /* Mouse.move_to, mouse.cpp:75-87 (Eigen isApprox, precision 1e-12) */
          const double dx = nx - px, dy = ny - py;
          const double na = nx * nx + ny * ny, nb = px * px + py * py;
          if (!((dx * dx + dy * dy) <= (1e-12 * 1e-12) * (na < nb ? na : nb))) {
            ori = atan2(ny - py, nx - px);
            px = nx;
            py = ny;
          }
          if (hist_pos) {
            hist_pos[(n * (T + 1) + t + 1) * 2] = px;
            hist_pos[(n * (T + 1) + t + 1) * 2 + 1] = py;
          }
          if (hist_ori) hist_ori[n * (T + 1) + t + 1] = ori;
          if (choice) choice[n * T + t] = (uint8_t)idx[pick];
        }
        if (st_code == 0 && occupancy) { /* history entry T */
          const int cx = py_floordiv_int(px, st), cy = py_floordiv_int(py, st);
          if (cx >= 0 && cx < X && cy >= 0 && cy < Y) occupancy[n * X * Y + cx * Y + cy] += 1;
        }
        for (int tt = t; tt < T && st_code != 0; ++tt) {
          if (choice) choice[n * T + tt] = 255;
          if (hist_pos) {
            hist_pos[(n * (T + 1) + tt + 1) * 2] = px;
            hist_pos[(n * (T + 1) + tt + 1) * 2 + 1] = py;
          }
          if (hist_ori) hist_ori[n * (T + 1) + tt + 1] = ori;
        }
        if (status) status[n] = st_code;
        if (final_pose) {
          final_pose[3 * n] = px;
          final_pose[3 * n + 1] = py;
          final_pose[3 * n + 2] = ori;
        }
      }
      free(exp_map);
    }
  }
  return err ? -2 : 0;
}
