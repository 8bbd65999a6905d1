// nm_vtable_trace: ONE parameter set over ONE session, with the per-step outputs of RLAlgorithm.update.
//
// The reference's update returns Update(transition, dv, |rpe|) for every online step and ConditioningExperiment.run puts
// it into the replay buffer (rl.py:97-98, 190-194, 233-237, 260-262); prioritised-replay style strategies read those
// fields.  The grid sweeps return tables only, so the one-set launch behind ConditioningExperiment.run goes through
// this kernel instead: same recurrence, same _rn operations (bit-identical table), plus dv[i] and |rpe|[i] per record.
//
// Mapping: one warp.  The recurrence is strictly sequential (V_t depends on V_{t-1}), so lane 0 walks it on a table in
// shared memory; the other lanes only stage: every lane fetches one 32-byte record of the next chunk of 32 into shared
// memory, and the chunk's 2 x 32 outputs leave through one coalesced store each.  ~50 ns per step.
#include <cuda_runtime.h>

#include "nm_common.h"

namespace {

constexpr int kTraceChunk = 32;

template <int ALG>
__global__ void __launch_bounds__(32, 1)
    nm_trace_kernel(const nm_step_rec *__restrict__ records, int64_t n, double alpha, double gamma, int num_states,
                    int n_tiles, const int32_t *__restrict__ tile_of_state, const int32_t *__restrict__ state_of_tile,
                    const double *__restrict__ v_init, double *__restrict__ v_out, double *__restrict__ dv_out,
                    double *__restrict__ rpe_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4 *chunk = reinterpret_cast<uint4 *>(smem_raw);                       // [32][2] raw records
  double *o_dv = reinterpret_cast<double *>(smem_raw + kTraceChunk * 32);   // [32]
  double *o_rpe = o_dv + kTraceChunk;                                       // [32]
  double *V = o_rpe + kTraceChunk;                                          // [num_states]
  const int lane = threadIdx.x;
  for (int s = lane; s < num_states; s += 32) V[s] = v_init ? v_init[tile_of_state[s]] : 0.0;  // rl.py:63-66
  const uint4 *src = reinterpret_cast<const uint4 *>(records);
  for (int64_t first = 0; first < n; first += kTraceChunk) {
    const int cnt = (int)(n - first < kTraceChunk ? n - first : kTraceChunk);
    if (lane < cnt) {
      chunk[2 * lane] = __ldg(src + 2 * (first + lane));
      chunk[2 * lane + 1] = __ldg(src + 2 * (first + lane) + 1);
    }
    __syncwarp();
    if (lane == 0) {
      for (int i = 0; i < cnt; ++i) {
        const uint4 w0 = chunk[2 * i], w1 = chunk[2 * i + 1];
        const double r = __hiloint2double((int)w0.y, (int)w0.x);
        const int s = (int)(w0.z & 0xffffu), s1 = (int)(w0.z >> 16), K = (int)(w0.w & 0xffu);
        double boot;
        if (ALG == NM_ALG_TD0) {
          boot = V[s1];  // rl.py:257
        } else {
          const uint32_t cw[4] = {w1.x, w1.y, w1.z, w1.w};
          boot = V[cw[0] & 0xffffu];
          for (int k = 1; k < K; ++k) {
            const double v = V[(k & 1) ? (cw[k >> 1] >> 16) : (cw[k >> 1] & 0xffffu)];
            if (ALG == NM_ALG_POSITIVE ? (v > boot) : (v < boot)) boot = v;  // np.max / np.min, rl.py:184, 227
          }
        }
        const double vs = V[s];
        const double rpe = __dsub_rn(__dadd_rn(r, __dmul_rn(gamma, boot)), vs);  // rl.py:185, 228, 257
        const double dv = __dmul_rn(alpha, rpe);                                  // rl.py:186, 229, 258
        V[s] = __dadd_rn(vs, dv);                                                 // update_v_table, rl.py:144
        o_dv[i] = dv;
        o_rpe[i] = fabs(rpe);                                                     // Update(rpe=abs(rpe))
      }
    }
    __syncwarp();
    if (lane < cnt) {
      if (dv_out) dv_out[first + lane] = o_dv[lane];
      if (rpe_out) rpe_out[first + lane] = o_rpe[lane];
    }
    __syncwarp();
  }
  __syncwarp();
  for (int t = lane; t < n_tiles; t += 32) {  // dense table; non-visitable tiles keep their initial value
    const int st = state_of_tile[t];
    v_out[t] = st >= 0 ? V[st] : (v_init ? v_init[t] : 0.0);
  }
}

}  // namespace

extern "C" int nm_vtable_trace_max_states(void) { return (int)((227 * 1024 - kTraceChunk * 48) / 8); }

extern "C" int nm_vtable_trace(int alg, const nm_maze *mz, const nm_streams *st, int session, double alpha, double gamma,
                               const double *d_v_init, double *d_v_out, double *d_dv_out, double *d_rpe_out,
                               void *stream) {
  NM_CHECK_ARG(mz && st && d_v_out, "nm_vtable_trace: NULL argument");
  NM_CHECK_ARG(alg == NM_ALG_TD0 || alg == NM_ALG_POSITIVE || alg == NM_ALG_NEGATIVE, "nm_vtable_trace: unknown algorithm %d", alg);
  NM_CHECK_ARG(mz->device >= 0 && mz->device == st->device, "nm_vtable_trace: maze and streams must live on the same CUDA device");
  NM_CHECK_ARG(st->num_states == mz->num_states, "nm_vtable_trace: streams were built for %d states, maze has %d",
               st->num_states, mz->num_states);
  NM_CHECK_ARG(session >= 0 && session < st->S, "nm_vtable_trace: session %d of %d", session, st->S);
  NM_CHECK_ARG(mz->num_states <= nm_vtable_trace_max_states(), "nm_vtable_trace: at most %d states",
               nm_vtable_trace_max_states());
  NM_CHECK_ARG(st->total == 0 || st->min_ncand > 0 || alg == NM_ALG_TD0,
               "nm_vtable_trace: the streams hold steps without candidate next states (np.max([]) raises in the reference)");
  if (cudaSetDevice(mz->device) != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaSetDevice(%d) failed", mz->device);
  const int64_t rec0 = st->offsets[session], n = st->offsets[session + 1] - rec0;
  const size_t smem = (size_t)kTraceChunk * 48 + (size_t)mz->num_states * 8;
  void (*kern)(const nm_step_rec *, int64_t, double, double, int, int, const int32_t *, const int32_t *, const double *,
               double *, double *, double *) =
      alg == NM_ALG_TD0 ? nm_trace_kernel<NM_ALG_TD0>
                        : alg == NM_ALG_POSITIVE ? nm_trace_kernel<NM_ALG_POSITIVE> : nm_trace_kernel<NM_ALG_NEGATIVE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaFuncSetAttribute(nm_trace_kernel): %s", cudaGetErrorString(e));
  kern<<<1, 32, smem, (cudaStream_t)stream>>>(st->d_records + rec0, n, alpha, gamma, mz->num_states, mz->X * mz->Y,
                                              mz->d_tile_of_state, mz->d_state_of_tile, d_v_init, d_v_out, d_dv_out,
                                              d_rpe_out);
  e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_trace_kernel launch: %s", cudaGetErrorString(e));
  return NM_OK;
}
