// Dependent-chain latency probe for the FP64 pipe and shared memory on sm_100a (one warp, one CTA).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/lat_probe tools/lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double *out, long long *cycles, int iters, double a, double b) {
  __shared__ double sm[64];
  sm[threadIdx.x] = a;
  __syncthreads();
  double x = a + threadIdx.x * 1e-9;
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (OP == 0) x = fma(x, b, a);
      if (OP == 1) x = __dadd_rn(x, a);
      if (OP == 2) x = __dmul_rn(x, b);
      if (OP == 3) x = fmax(x, a + j);           // DSETP + selects
      if (OP == 4) { idx = (int)sm[idx & 31] ; }  // LDS -> F2I -> LDS (pointer chase through smem)
      if (OP == 5) x = (x < a) ? x + 1.0 : x - 1.0;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + idx;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char *name, int warps) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 4096 * sizeof(double)); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  chain<OP><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 0.9999999);
  chain<OP><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 0.9999999);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps=%2d  %.2f cycles per dependent op\n", name, warps, (double)h / (iters * 16.0));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 16}) {
    run<0>("DFMA chain", w);
  }
  run<1>("DADD chain", 1);
  run<2>("DMUL chain", 1);
  run<3>("fmax(double) chain", 1);
  run<4>("LDS->cvt->LDS chain", 1);
  run<5>("compare+select+add chain", 1);
  return 0;
}
