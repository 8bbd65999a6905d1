// sin / cos / atan2 that reproduce, BIT FOR BIT, the results of the host C library the reference ran on.
//
// Why.  The reference's C++ Mouse calls std::cos / std::sin / std::atan2 (src/_navigation_model/mouse.cpp:66, 78); a pose
// is a chain of such calls, so a device libm that differs from the host's in the last place of ONE result makes every
// later position of that mouse differ in its last bits.  north_star asks for bit-exact simulated trajectories under
// replayed draws, so the simulation kernel evaluates these three functions with the algorithm of the host library --
// GNU libc 2.39 (the image's libm.so.6; x86-64 FMA variants, which is what every host with FMA3 + AVX2 dispatches to):
//     sysdeps/ieee754/dbl-64/s_sin.c   (__sin, __cos: table-driven, 440-entry sin/cos table on a 1/128 grid,
//                                       Cody-Waite reduction by pi/2 in four pieces for 2.43 < |x| < 1.05e8)
//     sysdeps/ieee754/dbl-64/e_atan2.c (__ieee754_atan2: quotient in double-double, 241-row table of
//                                       atan expansions for |u| >= 1/16, odd polynomial below)
// restated here operation by operation, INCLUDING which multiply-adds the FMA build fuses (taken from the machine code of
// the library: tools/libm_tables.py documents the addresses; tests/test_libm_port.py holds this file to the library on
// 10^8+ arguments on the host).  Every operation below is an IEEE-754 basic operation or a fused multiply-add, which
// round identically on the host and on the device -- so agreement with the library on the host is agreement on the GPU.
//
// Scope: finite arguments with |x| < 105414350 for sin / cos (beyond that the library switches to a multi-precision
// reduction, not restated: nml_sincos_in_scope); atan2 for finite arguments, signed
// zeros, subnormal and huge magnitudes included (the library's own scaling and exponent-gap shortcuts are restated).
// NaN / infinite arguments are out of scope (nml_atan2_in_scope).
//
// The tables (nm_libm_tables.h) are the library's own read-only data, extracted by tools/libm_tables.py.
#ifndef NM_LIBM_H
#define NM_LIBM_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define NML_HD __host__ __device__ __forceinline__
#else
#define NML_HD static inline
#endif
#if defined(__CUDA_ARCH__)
#define NML_FMA(a, b, c) __fma_rn((a), (b), (c))
#define NML_MUL(a, b) __dmul_rn((a), (b))
#define NML_ADD(a, b) __dadd_rn((a), (b))
#define NML_SUB(a, b) __dsub_rn((a), (b))
#define NML_DIV(a, b) __ddiv_rn((a), (b))
#define NML_TAB(t) t##_dev
#define NML_BITS(x) __double_as_longlong(x)
#define NML_FROM_BITS(i) __longlong_as_double(i)
#else
// host: compile with -ffp-contract=off so that nothing but NML_FMA fuses
#define NML_FMA(a, b, c) fma((a), (b), (c))
#define NML_MUL(a, b) ((a) * (b))
#define NML_ADD(a, b) ((a) + (b))
#define NML_SUB(a, b) ((a) - (b))
#define NML_DIV(a, b) ((a) / (b))
#define NML_TAB(t) t##_host
static inline int64_t nml_bits_(double x) {
  int64_t i;
  memcpy(&i, &x, 8);
  return i;
}
static inline double nml_from_bits_(int64_t i) {
  double x;
  memcpy(&x, &i, 8);
  return x;
}
#define NML_BITS(x) nml_bits_(x)
#define NML_FROM_BITS(i) nml_from_bits_(i)
#endif

#include "nm_libm_tables.h"

// ---- the libraries' constants (read from its .rodata; hexadecimal so that no decimal conversion rounds).  On the device
// they live in ONE __constant__ array, so that an FP64 instruction takes its constant straight from the constant bank: as
// literals every use costs two extra instructions that build the 64-bit immediate.
#define NML_CONSTANTS(X) \
  X(NML_S1, -0x1.5555555555555p-3) \
  X(NML_S2, 0x1.1111111110ecep-7) \
  X(NML_S3, -0x1.a01a019db08b8p-13) \
  X(NML_S4, 0x1.71de27b9a7ed9p-19) \
  X(NML_S5, -0x1.addffc2fcdf59p-26) \
  X(NML_SN3, -0x1.5555555555515p-3) \
  X(NML_SN5, 0x1.11110e829872fp-7) \
  X(NML_CS2, 0x1.0000000000000p-1) \
  X(NML_CS4, -0x1.5555555555535p-5) \
  X(NML_CS6, 0x1.6c16bedd9e239p-10) \
  X(NML_BIG, 0x1.8000000000000p+45) \
  X(NML_TOINT, 0x1.8000000000000p+52) \
  X(NML_HPINV, 0x1.45f306dc9c883p-1) \
  X(NML_MP1, 0x1.921fb58000000p+0) \
  X(NML_MP2, -0x1.dde973c000000p-27) \
  X(NML_PP3, -0x1.cb3b398000000p-55) \
  X(NML_PP4, -0x1.d747f23e32ed7p-83) \
  X(NML_HP0, 0x1.921fb54442d18p+0) \
  X(NML_HP1, 0x1.1a62633145c07p-54) \
  X(NML_TAYLOR_LIMIT, 0x1.020c49ba5e354p-3) \
  X(NML_D3, -0x1.5555555555555p-2) \
  X(NML_D5, 0x1.99999999997fdp-3) \
  X(NML_D7, -0x1.24924923f7603p-3) \
  X(NML_D9, 0x1.c71c6e5129a3bp-4) \
  X(NML_D11, -0x1.7458022b13c25p-4) \
  X(NML_D13, 0x1.375f08b31cbcep-4) \
  X(NML_HPI, 0x1.921fb54442d18p+0) \
  X(NML_HPI1, 0x1.1a62633145c07p-54) \
  X(NML_OPI, 0x1.921fb54442d18p+1) \
  X(NML_OPI1, 0x1.1a62633145c07p-53) \
  X(NML_TWO52, 0x1.0p+52) \
  X(NML_TWO8, 0x1.0p+8) \
  X(NML_TWO500, 0x1.0p+500) \
  X(NML_TWOM500, 0x1.0p-500)
enum {
#define NML_X(name, value) name##_IDX,
  NML_CONSTANTS(NML_X)
#undef NML_X
  NML_N_CONSTANTS
};
#if defined(__CUDACC__)
static __constant__ double nml_k_dev[NML_N_CONSTANTS] = {
#define NML_X(name, value) value,
    NML_CONSTANTS(NML_X)
#undef NML_X
};
#endif
static const double nml_k_host[NML_N_CONSTANTS] = {
#define NML_X(name, value) value,
    NML_CONSTANTS(NML_X)
#undef NML_X
};
#if defined(__CUDA_ARCH__)
#define NML_K(name) nml_k_dev[name##_IDX]
#else
#define NML_K(name) nml_k_host[name##_IDX]
#endif
#define NML_S1 NML_K(NML_S1)
#define NML_S2 NML_K(NML_S2)
#define NML_S3 NML_K(NML_S3)
#define NML_S4 NML_K(NML_S4)
#define NML_S5 NML_K(NML_S5)
#define NML_SN3 NML_K(NML_SN3)
#define NML_SN5 NML_K(NML_SN5)
#define NML_CS2 NML_K(NML_CS2)
#define NML_CS4 NML_K(NML_CS4)
#define NML_CS6 NML_K(NML_CS6)
#define NML_BIG NML_K(NML_BIG)
#define NML_TOINT NML_K(NML_TOINT)
#define NML_HPINV NML_K(NML_HPINV)
#define NML_MP1 NML_K(NML_MP1)
#define NML_MP2 NML_K(NML_MP2)
#define NML_PP3 NML_K(NML_PP3)
#define NML_PP4 NML_K(NML_PP4)
#define NML_HP0 NML_K(NML_HP0)
#define NML_HP1 NML_K(NML_HP1)
#define NML_TAYLOR_LIMIT NML_K(NML_TAYLOR_LIMIT)
#define NML_D3 NML_K(NML_D3)
#define NML_D5 NML_K(NML_D5)
#define NML_D7 NML_K(NML_D7)
#define NML_D9 NML_K(NML_D9)
#define NML_D11 NML_K(NML_D11)
#define NML_D13 NML_K(NML_D13)
#define NML_HPI NML_K(NML_HPI)
#define NML_HPI1 NML_K(NML_HPI1)
#define NML_OPI NML_K(NML_OPI)
#define NML_OPI1 NML_K(NML_OPI1)
#define NML_TWO52 NML_K(NML_TWO52)
#define NML_TWO8 NML_K(NML_TWO8)
#define NML_TWO500 NML_K(NML_TWO500)
#define NML_TWOM500 NML_K(NML_TWOM500)


NML_HD double nml_copysign(double mag, double sgn) {
  return NML_FROM_BITS((NML_BITS(mag) & 0x7fffffffffffffffLL) | (NML_BITS(sgn) & (int64_t)0x8000000000000000ULL));
}
NML_HD double nml_fabs(double x) { return NML_FROM_BITS(NML_BITS(x) & 0x7fffffffffffffffLL); }

// TAYLOR_SIN(xx, x, dx): x + ((POLY(xx) * x - 0.5 * dx) * xx + dx), every "* .. +" fused
NML_HD double nml_taylor_sin(double xx, double x, double dx) {
  double p = NML_FMA(NML_S5, xx, NML_S4);
  p = NML_FMA(p, xx, NML_S3);
  p = NML_FMA(p, xx, NML_S2);
  p = NML_FMA(p, xx, NML_S1);
  const double t1 = NML_FMA(p, x, -NML_MUL(0.5, dx));  // vfmsub: p * x - 0.5 * dx
  const double t = NML_FMA(xx, t1, dx);
  return NML_ADD(x, t);
}

// do_sin(a, da) (is_cos = 0) or do_cos(a, da) (is_cos = 1) of s_sin.c in ONE straight line: the two table-driven bodies
// differ in four places, each a select here, so a warp whose lanes need different functions does not diverge (the
// library's own code branches; the operations executed per lane are exactly the library's):
//   do_sin: if (a <= 0) da = -da;  x = |a| - (u - big);       s = x + (x*xx * P1 + da);  c = x * da + xx * P2;
//           cor = s * cs + (-(c * sn) + (s * ccs + ssn));  result = copysign(sn + cor, a)   [|a| >= 0.126, else Taylor]
//   do_cos: if (a <  0) da = -da;  x = |a| - (u - big) + da;  s = x*xx * P1 + x;         c = xx * P2;
//           cor = -(s * sn) + (-(c * cs) + (-(s * ssn) + ccs));  result = cs + cor
NML_HD double nml_do_sin_or_cos(double a, double da, int is_cos) {
  const double aa = nml_fabs(a);
  if (!is_cos && aa < NML_TAYLOR_LIMIT) return nml_taylor_sin(NML_MUL(a, a), a, da);
  if (is_cos ? (a < 0.0) : (a <= 0.0)) da = -da;
  const double u = NML_ADD(NML_BIG, aa);
  const double x0 = NML_SUB(aa, NML_SUB(u, NML_BIG));
  const double x = is_cos ? NML_ADD(x0, da) : x0;
  const int k = ((int)(uint32_t)NML_BITS(u)) << 2;
  const double xx = NML_MUL(x, x);
  const double m = NML_MUL(x, xx);
  const double p1 = NML_FMA(NML_SN5, xx, NML_SN3);
  const double p2 = NML_MUL(xx, NML_FMA(NML_FMA(NML_CS6, xx, NML_CS4), xx, NML_CS2));
  const double s = is_cos ? NML_FMA(m, p1, x) : NML_ADD(x, NML_FMA(m, p1, da));
  const double c = is_cos ? p2 : NML_FMA(x, da, p2);
  const double *T = NML_TAB(nml_sincostab) + k;
  const double sn = T[0], ssn = T[1], cs = T[2], ccs = T[3];
  const double A = is_cos ? cs : sn, dA = is_cos ? ccs : ssn, B = is_cos ? sn : cs, dB = is_cos ? ssn : ccs;
  const double ss = is_cos ? -s : s;
  const double res = NML_ADD(A, NML_FMA(ss, B, NML_FMA(-c, A, NML_FMA(ss, dB, dA))));
  return is_cos ? res : nml_copysign(res, a);
}

// 1 when nml_sin / nml_cos cover x (finite, below the multi-precision range of the library)
NML_HD int nml_sincos_in_scope(double x) { return ((uint32_t)(NML_BITS(x) >> 32) & 0x7fffffffu) < 0x419921FBu; }

// __sin and __cos in one pass.  Each is +-do_sin(a, da) or +-do_cos(a, da) for an (a, da) that depends on the range of |x|
// (s_sin.c): the ranges only SELECT the arguments here; the evaluation is the straight line above, once per function.
//   |x| < 0.855469:            sin = do_sin(x, 0)                       cos = do_cos(x, 0)
//   0.855469 <= |x| < 2.426265: sin = copysign(do_cos(hp0 - |x|, hp1), x)  cos = do_sin(y + hp1, (y - (y + hp1)) + hp1), y = hp0 - |x|
//   2.426265 <= |x| < 1.05e8:  (a, da, n) = reduce_sincos(x);  sin = do_sincos(a, da, n), cos = do_sincos(a, da, n + 1)
NML_HD void nml_sincos(double x, double *sin_out, double *cos_out) {
  const uint32_t k = (uint32_t)(NML_BITS(x) >> 32) & 0x7fffffffu;
  const double ax = nml_fabs(x);
  double sa, sda, ca, cda;
  int n_s, n_c, s_copysign_x = 0;
  if (k < 0x3feb6000u) {
    sa = ca = x;
    sda = cda = 0.0;
    n_s = 0;
    n_c = 1;
  } else if (k < 0x400368fdu) {
    const double y = NML_SUB(NML_HP0, ax);
    sa = y;
    sda = NML_HP1;
    n_s = 1;
    s_copysign_x = 1;
    ca = NML_ADD(y, NML_HP1);
    cda = NML_ADD(NML_SUB(y, ca), NML_HP1);
    n_c = 0;
  } else {  // reduce_sincos: x = n * pi/2 + (a + da)
    const double t = NML_FMA(x, NML_HPINV, NML_TOINT);
    const double xn = NML_SUB(t, NML_TOINT);
    const double y = NML_FMA(-xn, NML_MP2, NML_FMA(-xn, NML_MP1, x));
    const int n = (int)((uint32_t)NML_BITS(t) & 3u);
    const double t2 = NML_FMA(-xn, NML_PP3, y);
    double db = NML_FMA(-xn, NML_PP3, NML_SUB(y, t2));
    const double b = NML_FMA(-xn, NML_PP4, t2);
    db = NML_ADD(db, NML_FMA(-xn, NML_PP4, NML_SUB(t2, b)));
    sa = ca = b;
    sda = cda = db;
    n_s = n;
    n_c = n + 1;
  }
  double rs = nml_do_sin_or_cos(sa, sda, n_s & 1);
  double rc = nml_do_sin_or_cos(ca, cda, n_c & 1);
  if (s_copysign_x) rs = nml_copysign(rs, x);
  else if (n_s & 2) rs = -rs;
  if (n_c & 2) rc = -rc;
  *sin_out = k < 0x3e500000u ? x : rs;    // |x| < 2^-26
  *cos_out = k < 0x3e400000u ? 1.0 : rc;  // |x| < 2^-27
}

NML_HD double nml_sin(double x) {
  double s, c;
  nml_sincos(x, &s, &c);
  return s;
}
NML_HD double nml_cos(double x) {
  double s, c;
  nml_sincos(x, &s, &c);
  return c;
}

// ---- atan2 (e_atan2.c) ---------------------------------------------------------------------------------------------

// d3 + v * (d5 + v * (d7 + v * (d9 + v * (d11 + v * d13)))), every step fused
NML_HD double nml_atan_poly(double v) {
  double p = NML_FMA(NML_D13, v, NML_D11);
  p = NML_FMA(p, v, NML_D9);
  p = NML_FMA(p, v, NML_D7);
  p = NML_FMA(p, v, NML_D5);
  return NML_FMA(p, v, NML_D3);
}

// row of cij for u in [1/16, 1]: i = (int)((TWO52 + TWO8 * u) - TWO52) - 16
NML_HD const double *nml_atan_row(double u) {
  const double t = NML_SUB(NML_FMA(u, NML_TWO8, NML_TWO52), NML_TWO52);
  return NML_TAB(nml_cij) + 7 * ((int)t - 16);
}

// c[2] + v * (c[3] + v * (c[4] + v * (c[5] + v * c[6])))
NML_HD double nml_atan_row_poly(const double *c, double v) {
  double q = NML_FMA(c[6], v, c[5]);
  q = NML_FMA(q, v, c[4]);
  q = NML_FMA(q, v, c[3]);
  return NML_FMA(q, v, c[2]);
}

// finite arguments only (NaN / infinity: the caller uses the platform function)
NML_HD int nml_atan2_in_scope(double y, double x) {
  const uint32_t ex = (uint32_t)(NML_BITS(x) >> 32) & 0x7ff00000u, ey = (uint32_t)(NML_BITS(y) >> 32) & 0x7ff00000u;
  return ex != 0x7ff00000u && ey != 0x7ff00000u;
}

NML_HD double nml_atan2(double y, double x) {
  const int64_t bx = NML_BITS(x), by = NML_BITS(y);
  const int ux = (int)(bx >> 32), uy = (int)(by >> 32);
  if ((by & 0x7fffffffffffffffLL) == 0) return bx < 0 ? nml_copysign(NML_OPI, y) : y;  // y = +-0
  if ((bx & 0x7fffffffffffffffLL) == 0) return nml_copysign(NML_HPI, y);                 // x = +-0
  double ax = nml_fabs(x), ay = nml_fabs(y);
  const int de = (uy & 0x7ff00000) - (ux & 0x7ff00000);
  if (de >= 59768832) return nml_copysign(NML_HPI, y);  // |y / x| >= 2^57
  if (de <= -59768832) return x > 0.0 ? nml_copysign(NML_DIV(ay, ax), y) : nml_copysign(NML_OPI, y);
  if (ax < NML_TWOM500 || ay < NML_TWOM500) {
    ax = NML_MUL(ax, NML_TWO500);
    ay = NML_MUL(ay, NML_TWO500);
  }
  if (ax > NML_TWO500 || ay > NML_TWO500) {
    ax = NML_MUL(ax, NML_TWOM500);
    ay = NML_MUL(ay, NML_TWOM500);
  }
  // the quotient of the smaller by the larger magnitude in double-double: u + du
  double u, du;
  const int y_smaller = ay < ax;
  {
    const double num = y_smaller ? ay : ax, den = y_smaller ? ax : ay;
    u = NML_DIV(num, den);
    const double v = NML_MUL(u, den);
    const double vv = NML_FMA(u, den, -v);
    du = NML_DIV(NML_SUB(NML_SUB(num, v), vv), den);
  }
  // Four cases in the library: (i) x > 0, |y| < |x|: atan(u);  (ii) x > 0, |x| <= |y|: pi/2 - atan(u);
  // (iii) x < 0, |x| < |y|: pi/2 + atan(u);  (iv) x < 0, |y| <= |x|: pi - atan(u).  (ii)-(iv) are one formula with
  // K = pi/2 or pi (in two pieces K0 + K1) and a sign on u / du / the correction, so they are evaluated as ONE straight
  // line with selects (x - y == x + (-y) and fma(-a, b, c) == fma(a, -b, c) exactly); (i) keeps its own, more careful
  // form.  Below 1/16 an odd polynomial, above it the 241-row table of expansions around x_i.
  double z;
  const int small = u < 0x1.0p-4;
  if (x > 0.0 && y_smaller) {  // (i)
    if (small) {
      const double v = NML_MUL(u, u);
      z = NML_ADD(u, NML_FMA(NML_MUL(u, v), nml_atan_poly(v), du));
    } else {
      const double *c = nml_atan_row(u);
      const double t3 = NML_SUB(u, c[0]);
      const double v = NML_ADD(t3, du);
      const double dv = nml_fabs(t3) > nml_fabs(du) ? NML_ADD(NML_SUB(t3, v), du) : NML_ADD(NML_SUB(du, v), t3);
      double p = NML_FMA(c[6], v, c[5]);
      p = NML_FMA(p, v, c[4]);
      p = NML_FMA(p, v, c[3]);
      const double zz = NML_FMA(v, c[2], NML_FMA(dv, c[2], NML_MUL(NML_MUL(v, v), p)));
      z = NML_ADD(zz, c[1]);
    }
  } else {
    const int plus = !(x > 0.0) && !y_smaller;             // (iii): + atan(u); (ii), (iv): - atan(u)
    const int full = !(x > 0.0) && y_smaller;              // (iv): K = pi; (ii), (iii): K = pi/2
    const double K0 = full ? NML_OPI : NML_HPI, K1 = full ? NML_OPI1 : NML_HPI1;
    if (small) {
      const double v = NML_MUL(u, u);
      const double zz = NML_MUL(NML_MUL(u, v), nml_atan_poly(v));
      const double su = plus ? u : -u, sdu = plus ? du : -du, szz = plus ? zz : -zz;
      const double t2 = NML_ADD(K0, su);
      const double cor = NML_ADD(NML_SUB(K0, t2), su);
      z = NML_ADD(NML_ADD(NML_ADD(NML_ADD(cor, K1), sdu), szz), t2);
    } else {
      const double *c = nml_atan_row(u);
      const double v = NML_ADD(NML_SUB(u, c[0]), du);
      const double zz = NML_FMA(plus ? v : -v, nml_atan_row_poly(c, v), K1);
      z = NML_ADD(NML_ADD(K0, plus ? c[1] : -c[1]), zz);
    }
  }
  return nml_copysign(z, y);
}

#if defined(__CUDACC__)
// Device entry points.  Arguments must be in scope: finite, |heading| < 105414350 (checked where poses enter the
// library -- a mouse's heading is its initial one or an atan2 result, its positions are maze coordinates).  There is no
// fallback to the CUDA functions: their argument-reduction slow paths would triple the code a simulation step drags
// through the instruction cache for a case that does not occur.
#ifndef NM_LIBM_CALL
#define NM_LIBM_CALL __forceinline__  // measured on the simulation kernel: 171 ms inline against 187 ms out of line
#endif
static __device__ NM_LIBM_CALL void nm_host_sincos(double x, double *s, double *c) { nml_sincos(x, s, c); }
static __device__ NM_LIBM_CALL double nm_host_atan2(double y, double x) { return nml_atan2(y, x); }
#endif

#endif  // NM_LIBM_H
