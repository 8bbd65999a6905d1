/* Host check of navigation_model_b200/csrc/nm_libm.h: nml_sin / nml_cos / nml_atan2 against the C library's sin / cos /
 * atan2, bit for bit, on pseudo-random arguments drawn from the regimes the pose chain of a simulated mouse produces
 * (headings in [-pi, pi], displacements of a fraction of a tile) and from wide ranges.
 *     gcc -O2 -ffp-contract=off -mfma -fopenmp tests/libm_port_check.c -lm -o /tmp/libm_port_check
 *     /tmp/libm_port_check <samples per regime>
 * Prints one line per regime "name n mismatches" and exits 1 on any mismatch.  (tests/test_libm_port.py runs it.) */
#include <stdio.h>
#include <stdlib.h>

#include "../navigation_model_b200/csrc/nm_libm.h"

static inline uint64_t splitmix(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline double unit(uint64_t *s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline int same(double a, double b) { return NML_BITS(a) == NML_BITS(b); }

typedef struct {
  const char *name;
  double lo, hi;
  int log_scale;
} regime;

int main(int argc, char **argv) {
  const long n = argc > 1 ? atol(argv[1]) : 1000000;
  long bad_total = 0;
  const regime sincos_regimes[] = {{"sincos [-pi,pi]", -3.141592653589793, 3.141592653589793, 0},
                                   {"sincos [-0.13,0.13]", -0.13, 0.13, 0},
                                   {"sincos [-100,100]", -100, 100, 0},
                                   {"sincos log 1e-9..1e8", 1e-9, 1.0e8, 1},
                                   {"sincos near pi/2 multiples", 0, 0, 2}};
  for (unsigned r = 0; r < sizeof(sincos_regimes) / sizeof(regime); ++r) {
    const regime g = sincos_regimes[r];
    long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long i = 0; i < n; ++i) {
      uint64_t s = 0x1234567ULL * (r + 1) + (uint64_t)i * 0x9E3779B97F4A7C15ULL;
      double x;
      if (g.log_scale == 1) {
        x = exp(log(g.lo) + unit(&s) * (log(g.hi) - log(g.lo)));
        if (splitmix(&s) & 1) x = -x;
      } else if (g.log_scale == 2) {
        x = (double)((long)(unit(&s) * 200) - 100) * 1.5707963267948966 + (unit(&s) - 0.5) * 1e-6 * unit(&s);
      } else {
        x = g.lo + unit(&s) * (g.hi - g.lo);
      }
      if (!nml_sincos_in_scope(x)) continue;
      if (!same(nml_sin(x), sin(x)) || !same(nml_cos(x), cos(x))) {
        if (bad < 3) fprintf(stderr, "  %s: x=%a sin %a vs %a cos %a vs %a\n", g.name, x, nml_sin(x), sin(x), nml_cos(x), cos(x));
        ++bad;
      }
    }
    printf("%s %ld %ld\n", g.name, n, bad);
    bad_total += bad;
  }
  /* atan2: displacement vectors of a mouse (rotated actions of ~0.1 with rounding noise), wide ranges, axis cases */
  const char *atan_names[] = {"atan2 disc", "atan2 wide log", "atan2 near axes", "atan2 |y|~|x|", "atan2 signed zeros / tiny"};
  for (int r = 0; r < 5; ++r) {
    long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long i = 0; i < n; ++i) {
      uint64_t s = 0xABCDEFULL * (r + 1) + (uint64_t)i * 0x9E3779B97F4A7C15ULL;
      double x, y;
      if (r == 0) {
        const double th = (unit(&s) * 2 - 1) * 3.141592653589793, rad = 0.05 + 0.2 * unit(&s);
        x = rad * cos(th);
        y = rad * sin(th);
      } else if (r == 1) {
        x = exp((unit(&s) * 2 - 1) * 60) * ((splitmix(&s) & 1) ? -1 : 1);
        y = exp((unit(&s) * 2 - 1) * 60) * ((splitmix(&s) & 1) ? -1 : 1);
      } else if (r == 2) {
        const double big = 0.111 * (1 + (splitmix(&s) % 3)), small = (unit(&s) - 0.5) * exp(-unit(&s) * 45);
        if (splitmix(&s) & 1) { x = (splitmix(&s) & 1) ? big : -big; y = small; }
        else { y = (splitmix(&s) & 1) ? big : -big; x = small; }
      } else if (r == 3) {
        x = (unit(&s) - 0.5);
        y = x * (1 + (unit(&s) - 0.5) * exp(-unit(&s) * 40)) * ((splitmix(&s) & 1) ? -1 : 1);
      } else {
        const double v[] = {0.0, -0.0, 1e-310, -1e-310, 0.111, -0.111, 1e300, -1e300, 1e-200, -3e-160};
        x = v[splitmix(&s) % 10];
        y = v[splitmix(&s) % 10];
      }
      if (!nml_atan2_in_scope(y, x)) continue;
      if (!same(nml_atan2(y, x), atan2(y, x))) {
        if (bad < 3) fprintf(stderr, "  %s: y=%a x=%a got %a want %a\n", atan_names[r], y, x, nml_atan2(y, x), atan2(y, x));
        ++bad;
      }
    }
    printf("%s %ld %ld\n", atan_names[r], n, bad);
    bad_total += bad;
  }
  return bad_total ? 1 : 0;
}
