/* Host restatement of ExpSim::exp_nonpos (navigation_model_b200/csrc/nm_exp.cuh), operation for operation -- every
 * operation there is an IEEE basic operation or an FMA, which round the same on the device -- measured against the
 * long-double exponential.  expsim_tables.h (the table and the constants, copied VERBATIM from the header's initialisers)
 * is written by tests/test_expsim_cpu.py.
 *
 *   expsim_check N   ->  "<max ulp> <fraction within 1 ulp> <fraction correctly rounded> <max ulp of exp(0)..>"  */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "expsim_tables.h" /* static const double tab64[64], k[7], kll[4]; */

static double exp_nonpos(double x) {
  const double t = fma(x, k[0], k[1]);
  uint64_t tb;
  memcpy(&tb, &t, 8);
  int32_t kk = (int32_t)(uint32_t)tb; /* __double2loint */
  if (kk < -1021 * 64) kk = -1021 * 64;
  const uint32_t k8 = (uint32_t)kk << 3;
  const double n = t - k[1];
  double r = fma(n, k[2], x);
  r = fma(n, k[3], r);
  const double s = tab64[(k8 & 0x1f8u) >> 3];
  double p = fma(k[4], r, k[5]);
  p = fma(p, r, k[6]);
  p = fma(p, r, 0.5);
  const double q = fma(p, r * r, r);
  const double v = fma(s, q, s);
  uint64_t vb;
  memcpy(&vb, &v, 8);
  int32_t hi = (int32_t)(vb >> 32);
  hi = (int32_t)((uint32_t)hi + (k8 & 0xfffffe00u) * 2048u); /* (k >> 6) << 20 */
  vb = ((uint64_t)(uint32_t)hi << 32) | (vb & 0xffffffffu);
  double out;
  memcpy(&out, &vb, 8);
  return out;
}

/* ExpLL::operator() (likelihood kernels): one fused reduction step, degree-3 polynomial, product with the table value */
static double exp_ll(double x) {
  const double t = fma(x, kll[0], kll[1]);
  uint64_t tb;
  memcpy(&tb, &t, 8);
  int32_t kk = (int32_t)(uint32_t)tb;
  if (kk < -1021 * 64) kk = -1021 * 64;
  const uint32_t k8 = (uint32_t)kk << 3;
  const double r = fma(t - kll[1], kll[2], x);
  const double s = tab64[(k8 & 0x1f8u) >> 3];
  double p = fma(kll[3], r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p = p * s;
  uint64_t vb;
  memcpy(&vb, &p, 8);
  int32_t hi = (int32_t)(vb >> 32);
  hi = (int32_t)((uint32_t)hi + (k8 & 0xfffffe00u) * 2048u);
  vb = ((uint64_t)(uint32_t)hi << 32) | (vb & 0xffffffffu);
  double out;
  memcpy(&out, &vb, 8);
  return out;
}

static int main_ll(long N) {
  srand48(777);
  double worst = 0.0;
  for (long i = 0; i < N; ++i) {
    const double x = (i & 1) ? -drand48() * 40.0 : -drand48() * 2.0;
    const long double ref = expl((long double)x);
    const double rel = fabs((double)(((long double)exp_ll(x) - ref) / ref));
    if (rel > worst) worst = rel;
  }
  /* exp(0) exact; a rounding error above 0 (fma(beta, Vmax, -fl(beta * Vmax))) stays next to 1; hugely negative
   * arguments give a tiny positive number through the clamped exponent, never garbage */
  const int ok = exp_ll(0.0) == 1.0 && fabs(exp_ll(3e-13) - 1.0) < 1e-12 && exp_ll(-800.0) >= 0.0 && exp_ll(-800.0) < 1e-300 &&
                 exp_ll(-5000.0) >= 0.0 && exp_ll(-5000.0) < 1e-300 && exp_ll(-1e6) >= 0.0 && exp_ll(-1e6) < 1e-300;
  printf("%.4e %d\n", worst, ok);
  return 0;
}

int main(int argc, char **argv) {
  const long N = argc > 1 ? atol(argv[1]) : 1000000;
  if (argc > 2 && !strcmp(argv[2], "ll")) return main_ll(N);
  srand48(12345);
  double worst = 0.0;
  long within = 0, exact = 0;
  for (long i = 0; i < N; ++i) {
    double x;
    switch (i & 3) { /* softmax arguments, component arguments kappa * count, deep tails, tiny arguments */
      case 0: x = -drand48() * 40.0; break;
      case 1: x = -drand48() * 2.0; break;
      case 2: x = -drand48() * 700.0; break;
      default: x = -drand48() * 1e-3; break;
    }
    const long double ref = expl((long double)x);
    const double rd = (double)ref, ulp = nextafter(rd, INFINITY) - rd;
    const double got = exp_nonpos(x);
    const double err = fabs((double)((long double)got - ref)) / ulp;
    if (err > worst) worst = err;
    within += err <= 1.0;
    exact += got == rd;
  }
  /* exp(0) must be exactly 1: the softmax's largest term */
  printf("%.4f %.6f %.6f %d\n", worst, (double)within / N, (double)exact / N, exp_nonpos(0.0) == 1.0 && exp_nonpos(-0.0) == 1.0);
  return 0;
}
