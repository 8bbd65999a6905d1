// Run detection shared by the sweeps whose per-parameter-set tables depend on (alpha, gamma) only (nm_q.cu, nm_mb.cu;
// nm_fit.cu keeps its copy next to its arg-min scratch): is every aligned run of 32 consecutive parameter sets
// uniform in (alpha, gamma)?  Decided on the device, read by the kernels through one word -- no host round trip.
// Included inside each translation unit's anonymous namespace (needs <map>, <mutex>, <utility>, nm_common.h).
#ifndef NM_GROUP_CUH
#define NM_GROUP_CUH

__global__ void nm_runs_set_kernel(int *mode, int value) { mode[0] = value; }

// mode[0] must have been set to 1; cleared when a parameter set differs from the head of its run (bit compare: NaN or
// -0.0 vs 0.0 break a run, which only costs the shared path)
__global__ void __launch_bounds__(256) nm_runs_detect_kernel(const double *__restrict__ alpha,
                                                             const double *__restrict__ gamma, int64_t P, int *mode) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int64_t head = p & ~(int64_t)31;
  if (__double_as_longlong(alpha[p]) != __double_as_longlong(alpha[head]) ||
      __double_as_longlong(gamma[p]) != __double_as_longlong(gamma[head]))
    mode[0] = 0;
}

// one word of sweep mode per (device, stream), allocated on first use and kept for the life of the process
static int runs_mode_word(void *stream, int **out) {
  int device = 0;
  cudaGetDevice(&device);
  static std::mutex mu;
  static std::map<std::pair<int, void *>, int *> words;
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_pair(device, stream);
  auto it = words.find(key);
  if (it == words.end()) {
    int *buf = nullptr;
    const cudaError_t e = cudaMalloc((void **)&buf, 64);
    if (e != cudaSuccess) return nm_fail(NM_ERR_NOMEM, "sweep mode word: cudaMalloc: %s", cudaGetErrorString(e));
    it = words.emplace(key, buf).first;
  }
  *out = it->second;
  return NM_OK;
}

// The decision and the launches that read it must reach the stream as one uninterrupted sequence: two host threads
// enqueueing sweeps on the SAME stream would otherwise interleave "set / detect / kernels" and read each other's word.
static std::mutex &runs_launch_mutex() {
  static std::mutex mu;
  return mu;
}

// enable = false: mode 0 (one table per parameter set); else 1 iff every aligned run of 32 sets is uniform
static void runs_decide(bool enable, const double *d_alpha, const double *d_gamma, int64_t P, int *mode, cudaStream_t cs) {
  nm_runs_set_kernel<<<1, 1, 0, cs>>>(mode, enable ? 1 : 0);
  if (enable) nm_runs_detect_kernel<<<(unsigned)((P + 255) / 256), 256, 0, cs>>>(d_alpha, d_gamma, P, mode);
}

// synchronous read of the word (diagnostics / tests)
static int runs_mode_read(void *stream, int *mode_out, const char *who) {
  if (!mode_out) return nm_fail(NM_ERR_INVALID, "%s: NULL output", who);
  int *mode = nullptr;
  const int rc = runs_mode_word(stream, &mode);
  if (rc != NM_OK) return rc;
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaMemcpy(mode_out, mode, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
  return NM_OK;
}

#endif  // NM_GROUP_CUH
