// How the FP64 pipe of one SM sub-partition (scheduler) shares between warps on sm_100a.  One CTA on one SM,
// 4 * w warps (w per scheduler).  Per iteration every thread runs
//   phase A: a dependent chain of NA integer multiply-adds (no FP64; pure latency, ~4 cycles each), then
//   phase B: NB FP64 operations in CH independent chains (OPMIX 0: DFMA only; 1: DFMA / DADD / DMUL / compare-select mix)
// and the cycles per iteration are printed for w = 1..4.  Ideal sharing: T(w) = max(A + B/ilp latency, w * NB * 2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/_build/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, int OPMIX>
__global__ void probe(double *out, long long *cycles, int iters, int na, int nb, double a, double b, int seed) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = a + (threadIdx.x + c) * 1e-9;
  int k = seed + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    for (int j = 0; j < na; ++j) k = k * 1664525 + 1013904223;  // dependent IMAD chain
    if (k == 12345) x[0] += 1.0;                                 // phase B depends on phase A (never true)
    for (int j = 0; j < nb; j += CH) {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (OPMIX == 0) x[c] = fma(x[c], b, a);
        else {
          const int sel = (j / CH) & 3;
          if (sel == 0) x[c] = fma(x[c], b, a);
          else if (sel == 1) x[c] = __dadd_rn(x[c], a);
          else if (sel == 2) x[c] = __dmul_rn(x[c], b);
          else x[c] = (x[c] > a) ? x[c] : a;
        }
      }
    }
    if (x[0] == 98765.4321) k += 1;  // phase A of the next iteration depends on phase B (never true)
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  out[threadIdx.x] = s + k;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <int CH, int OPMIX>
void run(const char *name, int na, int nb) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 4096 * sizeof(double)); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  printf("%-34s A=%4d imad  B=%4d fp64:", name, na, nb);
  for (int w = 1; w <= 4; ++w) {
    for (int rep = 0; rep < 2; ++rep) probe<CH, OPMIX><<<1, 128 * w>>>(out, cyc, iters, na, nb, 1.0000001, 0.9999999, 7);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("  w=%d %7.1f", w, (double)h / iters);
  }
  printf("   cycles / iteration\n");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<1, 0>("DFMA, 1 chain", 0, 144);
  run<3, 0>("DFMA, 3 chains", 0, 144);
  run<6, 0>("DFMA, 6 chains", 0, 144);
  run<3, 1>("FP64 mix, 3 chains", 0, 144);
  run<3, 0>("DFMA, 3 chains + latency phase", 100, 144);
  run<6, 0>("DFMA, 6 chains + latency phase", 100, 144);
  run<3, 1>("FP64 mix, 3 chains + latency phase", 100, 144);
  run<3, 0>("latency phase only", 100, 0);
  return 0;
}
