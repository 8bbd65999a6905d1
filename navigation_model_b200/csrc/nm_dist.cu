// Row-wise distances between metric results of N simulated mice and one set of data metrics: the arithmetic of
// the reference's analysis/statistics.py (normalized_distance :67-92, manhattan_distance :95-103, js_distance
// :10-54) batched on the device, so that a parameter sweep over simulations reduces
// simulation -> metrics -> distance without returning per-mouse histograms to the host.
//
// One WARP per row (coalesced reads of the [N, L] matrix, shuffle reductions in FP64).  Reduction order differs
// from numpy's dot / pairwise sums: results agree to ~1e-15 relative, not bit for bit.
#include <cuda_runtime.h>
#include <math.h>

#include "nm_common.h"

namespace {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFullMask, v, off);
  return v;
}

// scipy.special.rel_entr
__device__ __forceinline__ double rel_entr(double x, double y) {
  if (x > 0.0 && y > 0.0) return x * log(x / y);
  if (x == 0.0 && y >= 0.0) return 0.0;
  return INFINITY;
}

__global__ void __launch_bounds__(256) nm_dist_kernel(int kind, const double *__restrict__ x,
                                                      const double *__restrict__ ref, int64_t N, int L,
                                                      double *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= N) return;
  const double *xr = x + row * L;
  if (kind == NM_DIST_MANHATTAN) {  // np.linalg.norm(x1 - x2, ord=1)
    double acc = 0.0;
    for (int i = lane; i < L; i += 32) acc += fabs(xr[i] - ref[i]);
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc;
    return;
  }
  double c1 = 0.0, c2 = 0.0;
  for (int i = lane; i < L; i += 32) {
    c1 += xr[i];
    c2 += ref[i];
  }
  c1 = warp_sum(c1);
  c2 = warp_sum(c2);
  if (kind == NM_DIST_NORMALIZED) {  // statistics.py:67-92
    double acc = 0.0;
    for (int i = lane; i < L; i += 32) {
      double d;
      if (c1 == 0.0 && c2 == 0.0) d = 0.0;
      else if (c1 == 0.0) d = ref[i] / c2;
      else if (c2 == 0.0) d = xr[i] / c1;
      else d = xr[i] / c1 - ref[i] / c2;
      acc += d * d;
    }
    acc = sqrt(warp_sum(acc));
    if (lane == 0) out[row] = acc > 1.0 ? 1.0 : acc;
    return;
  }
  // Jensen-Shannon distance, statistics.py:10-54 (an all-zero histogram is the zero vector; inf -> 1)
  double left = 0.0, right = 0.0;
  for (int i = lane; i < L; i += 32) {
    const double p = c1 == 0.0 ? 0.0 : xr[i] / c1;
    const double q = c2 == 0.0 ? 0.0 : ref[i] / c2;
    const double m = (p + q) / 2.0;
    left += rel_entr(p, m);
    right += rel_entr(q, m);
  }
  left = warp_sum(left);
  right = warp_sum(right);
  double dist = sqrt((left + right) / 2.0);
  if (isinf(dist)) dist = 1.0;
  if (lane == 0) out[row] = dist;
}

}  // namespace

extern "C" int nm_distance_rows(int kind, const double *d_x, const double *d_ref, int64_t N, int L, double *d_out,
                                void *stream) {
  NM_CHECK_ARG(kind == NM_DIST_NORMALIZED || kind == NM_DIST_MANHATTAN || kind == NM_DIST_JS,
               "nm_distance_rows: unknown distance %d", kind);
  NM_CHECK_ARG(N >= 0 && L >= 1 && (N == 0 || (d_x && d_ref && d_out)), "nm_distance_rows: bad argument");
  if (N == 0) return NM_OK;
  const int64_t blocks = (N * 32 + 255) / 256;
  NM_CHECK_ARG(blocks <= 2147483647LL, "nm_distance_rows: too many rows for one launch");
  nm_dist_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(kind, d_x, d_ref, N, L, d_out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_dist_kernel launch: %s", cudaGetErrorString(e));
  return NM_OK;
}
