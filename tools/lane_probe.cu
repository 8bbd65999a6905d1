// What a warp-wide FP64 instruction costs on sm_100a as a function of its ACTIVE lanes.  One CTA on one SM, 4 * w warps
// (w per scheduler); every warp runs 8 independent DFMA chains inside a branch that only the lanes of `mask` take.
// Printed: cycles per DFMA warp-instruction and scheduler (throughput-bound at w = 4 x 8 chains).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/_build/lane_probe tools/lane_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(double *out, long long *cycles, int iters, unsigned mask, double a, double b) {
  double x[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) x[c] = a + (threadIdx.x + c) * 1e-9;
  const bool on = (mask >> (threadIdx.x & 31)) & 1u;
  __syncthreads();
  const long long t0 = clock64();
  if (on) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 8; ++c) x[c] = fma(x[c], b, a);
    }
  }
  __syncwarp();
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += x[c];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 4096 * sizeof(double)); cudaMalloc(&cyc, 8);
  const int iters = 4000;
  struct { const char *name; unsigned mask; } cases[] = {
      {"32 lanes", 0xffffffffu}, {"lower 16", 0x0000ffffu}, {"upper 16", 0xffff0000u}, {"even 16", 0x55555555u},
      {"lower 8", 0x000000ffu}, {"8 spread (every 4th)", 0x11111111u}, {"10 scattered", 0x2490a481u},
      {"17 (lower 16 + lane 16)", 0x0001ffffu}, {"1 lane", 0x00000001u}, {"lanes 0 and 31", 0x80000001u}};
  for (auto &cs : cases) {
    printf("%-26s", cs.name);
    for (int w = 1; w <= 4; ++w) {
      for (int rep = 0; rep < 2; ++rep) probe<<<1, 128 * w>>>(out, cyc, iters, cs.mask, 1.0000001, 0.9999999);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("  w=%d %5.2f", w, (double)h / ((double)iters * 8 * w));
    }
    printf("   cycles per DFMA warp-instruction per scheduler\n");
  }
  return 0;
}
