// Table-driven FP64 exponentials for non-positive arguments: ExpLL (likelihood kernels: fit, Q-table, model-based) and ExpSim
// (simulation kernel), and the per-chunk logarithm of the likelihood kernels.
// Included inside each translation unit's anonymous namespace (every TU gets its own constant tables).
#ifndef NM_EXP_CUH
#define NM_EXP_CUH

// ---- exponentials of a LIKELIHOOD step (softmax denominator) ---------------------------------------------------------
// Bar: 1e-9 relative on a session's log-likelihood (BASELINE.json north_star), not the last bit -- value tables never see
// an exponential, and the simulation kernel (whose discrete choices must reproduce the reference's) has its own ~1-ulp form, ExpSim below.
// exp(x), x <= ~0:  x = (64 e + j) * ln2/64 + r, |r| <= ln2/128:  exp(x) = 2^e * 2^(j/64) * P3(r)
//   * ONE fused step of range reduction against the double nearest to ln2/64 (error |x| * 1.1e-16 relative -- a term of
//     weight exp(x) <= 1 in a sum >= 1 moves it by < 5e-17);
//   * a 64-entry table (four 128-byte shared-memory rows) and a degree-3 Taylor polynomial: truncation r^4/24 <= 3.6e-11
//     relative per exponential (mean 7e-12), always low: a session's log-likelihood moves by < 1e-11 relative;
//   * the caller forms x = fma(beta, v, -m) in one instruction.
// 7 FP64 instructions + the caller's fma instead of 10 + 2 (round 1's degree-5 / 16-entry form), 5 integer instructions
// instead of 6.  An FP64 warp instruction holds a scheduler's issue port for two cycles on this part (64 lanes / clk / SM),
// so each one saved is worth two of the others.  The binary exponent is clamped (k >= -1021 * 64) so that a hugely
// negative x gives ~2^-1021, not garbage.
__constant__ double c_exp2tab64[64] = {
    1.00000000000000000e+00, 1.01088928605170048e+00, 1.02189714865411663e+00, 1.03302487902122841e+00,
    1.04427378242741375e+00, 1.05564517836055716e+00, 1.06714040067682370e+00, 1.07876079775711986e+00,
    1.09050773266525769e+00, 1.10238258330784089e+00, 1.11438674259589243e+00, 1.12652161860824185e+00,
    1.13878863475669156e+00, 1.15118922995298267e+00, 1.16372485877757748e+00, 1.17639699165028122e+00,
    1.18920711500272103e+00, 1.20215673145270308e+00, 1.21524735998046896e+00, 1.22848053610687002e+00,
    1.24185781207348400e+00, 1.25538075702469110e+00, 1.26905095719173322e+00, 1.28287001607877826e+00,
    1.29683955465100964e+00, 1.31096121152476441e+00, 1.32523664315974132e+00, 1.33966752405330292e+00,
    1.35425554693689265e+00, 1.36900242297459052e+00, 1.38390988196383202e+00, 1.39897967253831124e+00,
    1.41421356237309515e+00, 1.42961333839197002e+00, 1.44518080697704665e+00, 1.46091779418064704e+00,
    1.47682614593949935e+00, 1.49290772829126484e+00, 1.50916442759342284e+00, 1.52559815074453842e+00,
    1.54221082540794074e+00, 1.55900440023783693e+00, 1.57598084510788650e+00, 1.59314215134226700e+00,
    1.61049033194925428e+00, 1.62802742185734783e+00, 1.64575547815396495e+00, 1.66367658032673638e+00,
    1.68179283050742900e+00, 1.70010635371852348e+00, 1.71861929812247793e+00, 1.73733383527370622e+00,
    1.75625216037329945e+00, 1.77537649252652119e+00, 1.79470907500310717e+00, 1.81425217550039886e+00,
    1.83400808640934243e+00, 1.85397912508338547e+00, 1.87416763411029996e+00, 1.89457598158696561e+00,
    1.91520656139714740e+00, 1.93606179349229435e+00, 1.95714412417540018e+00, 1.97845602638795093e+00};
__constant__ double c_expll_k[4] = {9.23324826168936567683e+01,   // 64 * log2(e)
                                    6755399441055744.0,           // 1.5 * 2^52
                                    -1.08304246962491450973e-02,  // -ln2/64 (nearest double)
                                    1.0 / 6.0};

struct ExpLL {
  double c[4];
  uint32_t tab;  // shared-window address of 64 doubles holding c_exp2tab64, 512-BYTE ALIGNED (the entry index is OR-ed in)
  // `region`: 1024 bytes of shared memory; the table sits at the first 512-byte boundary inside it.
  // fill(): threads 0..63 of the CTA write the table -- call it BEFORE the __syncthreads() that precedes the first use.
  static __device__ __forceinline__ void fill(unsigned char *region, int tid) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(region);
    if (tid < 64) reinterpret_cast<double *>(region + (((base + 511u) & ~511u) - base))[tid] = c_exp2tab64[tid];
  }
  __device__ __forceinline__ void load(unsigned char *region) {
#pragma unroll
    for (int i = 0; i < 4; ++i) asm volatile("mov.f64 %0, %1;" : "=d"(c[i]) : "d"(c_expll_k[i]));
    tab = ((uint32_t)__cvta_generic_to_shared(region) + 511u) & ~511u;
    asm volatile("mov.u32 %0, %0;" : "+r"(tab));  // pinned: not recomputed per step under register pressure
  }
  __device__ __forceinline__ double operator()(double x) const {
    const double t = fma(x, c[0], c[1]);
    const uint32_t k8 = (uint32_t)max(__double2loint(t), -1021 * 64) << 3;  // 8 * round(64 x / ln2), clamped
    const double r = fma(__dsub_rn(t, c[1]), c[2], x);
    double s;
    asm("ld.shared.f64 %0, [%1];" : "=d"(s) : "r"(tab | (k8 & 0x1f8u)));
    double p = fma(c[3], r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p = __dmul_rn(p, s);
    // binary exponent: hi += (k >> 6) << 20 = (k8 & ~0x1ff) * 2048 -- one logic op + one integer multiply-add (spelled in
    // PTX: the compiler otherwise turns the product back into shift + mask + add)
    int hi;
    asm("{\n.reg .b32 t;\nand.b32 t, %1, 0xfffffe00;\nmad.lo.s32 %0, t, 2048, %2;\n}" : "=r"(hi) : "r"(k8), "r"(__double2hiint(p)));
    return __hiloint2double(hi, __double2loint(p));
  }
};

// ---- exponentials of the SIMULATION kernel (component values and the softmax of a decision) ------------------------------
// exp(x), x <= 0, to ~1 ulp: x = (64 e + j) * ln2/64 + r, |r| <= ln2/128, two-step Cody-Waite reduction (exact products:
// the high part of ln2/64 has 32 trailing zero bits), the 64-entry table of ExpLL, and
//     exp(r) - 1 = q = r + r^2 * (1/2 + r/6 + r^2/24 + r^3/120)      (truncation r^6/720 <= 3.5e-17)
//     result     = fma(s, q, s),  s = 2^(j/64)                        (ONE rounding after the table value's own)
// 10 FP64 instructions (round 1's 16-entry / degree-7 form: 12) and a shorter dependent chain (r^2 beside the Horner steps).
// Measured on the host against the long-double exponential over [-700, 0] (2e7 arguments): max 1.27 ulp, 99.86 % within
// 1 ulp, 81 % correctly rounded -- the 16-entry form: max 1.96 ulp, 96 % within 1 ulp, 75 % correctly rounded.  Seven
// constants pinned in registers (0.5 is an immediate); the binary exponent is clamped like ExpLL's.
__constant__ double c_expsim_k[7] = {9.23324826168936567683e+01,           // 64 * log2(e)
                                     6755399441055744.0,                  // 1.5 * 2^52
                                     -6.93147180369123816490e-01 / 64.0,  // -ln2/64 (high part, exact scaling)
                                     -1.90821492927058770002e-10 / 64.0,  // -ln2/64 (low part)
                                     1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};

struct ExpSim {
  double c[7];
  uint32_t tab;  // shared-window address of 64 doubles holding c_exp2tab64, 512-BYTE ALIGNED (the entry index is OR-ed in)
  static constexpr int kRegionBytes = 1024;  // the table sits at the first 512-byte boundary inside the region
  // every thread of the CTA calls fill() BEFORE the __syncthreads() that precedes the first use (any CTA size)
  static __device__ __forceinline__ void fill(unsigned char *region, int tid) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(region);
    double *t = reinterpret_cast<double *>(region + (((base + 511u) & ~511u) - base));
    for (int i = tid; i < 64; i += (int)blockDim.x) t[i] = c_exp2tab64[i];
  }
  __device__ __forceinline__ void load(unsigned char *region) {
#pragma unroll
    for (int i = 0; i < 7; ++i) asm volatile("mov.f64 %0, %1;" : "=d"(c[i]) : "d"(c_expsim_k[i]));
    tab = ((uint32_t)__cvta_generic_to_shared(region) + 511u) & ~511u;
    asm volatile("mov.u32 %0, %0;" : "+r"(tab));  // pinned: not recomputed per use under register pressure
  }
  __device__ __forceinline__ double exp_nonpos(double x) const {
    const double t = fma(x, c[0], c[1]);
    const uint32_t k8 = (uint32_t)max(__double2loint(t), -1021 * 64) << 3;  // 8 * round(64 x / ln2), clamped
    const double n = __dsub_rn(t, c[1]);
    double r = fma(n, c[2], x);
    r = fma(n, c[3], r);
    double s;
    asm("ld.shared.f64 %0, [%1];" : "=d"(s) : "r"(tab | (k8 & 0x1f8u)));
    double p = fma(c[4], r, c[5]);
    p = fma(p, r, c[6]);
    p = fma(p, r, 0.5);
    const double q = fma(p, __dmul_rn(r, r), r);
    const double v = fma(s, q, s);
    int hi;  // hi += (k >> 6) << 20 = (k8 & ~0x1ff) * 2048
    asm("{\n.reg .b32 t;\nand.b32 t, %1, 0xfffffe00;\nmad.lo.s32 %0, t, 2048, %2;\n}" : "=r"(hi) : "r"(k8), "r"(__double2hiint(v)));
    return __hiloint2double(hi, __double2loint(v));
  }
};

// log(x) for the per-chunk fold of the likelihood kernels: x = product of 32 softmax denominators, in [1, 2^96] (normal,
// positive -- no special cases).  x = 2^e * m, m in [sqrt(1/2), sqrt(2)):  log x = e * ln2 + 2 atanh(s), s = (m-1)/(m+1),
// |s| <= 0.1716: odd series to s^19 (next term 2 s^21 / 21 < 8e-18).  ~30 instructions against ~200 of the library log
// (whose denormal / special-value paths this range never takes); relative error < 3e-16 -- far inside the 1e-9 bar.
__device__ __forceinline__ double log_chunk_product(double x) {
  int hi = __double2hiint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;  // m in [1, 2)
  if (hi >= 0x3ff6a09f) {               // m >= sqrt(2): halve it
    hi -= 0x00100000;
    e += 1;
  }
  const double m = __hiloint2double(hi, __double2loint(x));
  const double s = __ddiv_rn(__dsub_rn(m, 1.0), __dadd_rn(m, 1.0));
  const double s2 = __dmul_rn(s, s);
  double p = 2.0 / 19.0;
  p = fma(p, s2, 2.0 / 17.0);
  p = fma(p, s2, 2.0 / 15.0);
  p = fma(p, s2, 2.0 / 13.0);
  p = fma(p, s2, 2.0 / 11.0);
  p = fma(p, s2, 2.0 / 9.0);
  p = fma(p, s2, 2.0 / 7.0);
  p = fma(p, s2, 2.0 / 5.0);
  p = fma(p, s2, 2.0 / 3.0);
  p = fma(p, s2, 2.0);
  // e * ln2 in two pieces (hi part exact for |e| < 2^20), then the series
  const double de = (double)e;
  return fma(de, 6.93147180369123816490e-01, fma(de, 1.90821492927058770002e-10, __dmul_rn(p, s)));
}

#endif  // NM_EXP_CUH
