// Small FP64 building blocks of a likelihood step, shared by the value-table (nm_fit.cu) and Q-table (nm_q.cu) sweeps:
// numpy-order sums, first-of-equals extrema, shared-memory loads by byte offset.  Included inside each translation
// unit's anonymous namespace, after NM_MAX_CAND / NM_ALG_* (navmodel_b200.h).
#ifndef NM_STEP_MATH_CUH
#define NM_STEP_MATH_CUH

__device__ __forceinline__ double dmax(double a, double b) { return (b > a) ? b : a; }  // first of equals, like np.max
__device__ __forceinline__ double dmin(double a, double b) { return (b < a) ? b : a; }

template <int ALG, int K>
__device__ __forceinline__ double extremum(const double (&v)[NM_MAX_CAND]) {
  // sequential left-to-right like np.max / np.min over the candidate list (rl.py:184, 227); a balanced tree
  // of compare-selects returns the same value
  double t[NM_MAX_CAND];
#pragma unroll
  for (int k = 0; k < NM_MAX_CAND; ++k) t[k] = v[k < K ? k : 0];
#pragma unroll
  for (int w = 1; w < NM_MAX_CAND; w <<= 1)
#pragma unroll
    for (int k = 0; k + w < NM_MAX_CAND; k += 2 * w)
      if (k + w < K) t[k] = (ALG == NM_ALG_POSITIVE) ? dmax(t[k], t[k + w]) : dmin(t[k], t[k + w]);
  return t[0];
}

template <int K>
__device__ __forceinline__ double np_sum_fixed(const double (&e)[NM_MAX_CAND]) {
  if (K == NM_MAX_CAND) {  // numpy pairwise sum, n == 8
    return __dadd_rn(__dadd_rn(__dadd_rn(e[0], e[1]), __dadd_rn(e[2], e[3])),
                     __dadd_rn(__dadd_rn(e[4], e[5]), __dadd_rn(e[6], e[7])));
  }
  double s = 0.0;  // n < 8: sequential from 0.0
#pragma unroll
  for (int k = 0; k < K; ++k) s = __dadd_rn(s, e[k]);
  return s;
}

__device__ __forceinline__ double lds_f64(const char *base, uint32_t off) {
  return *reinterpret_cast<const double *>(base + off);
}

#endif  // NM_STEP_MATH_CUH
