// Session ingestion on the device: raw recordings -> subsampled sessions -> step streams,
// for many sessions at once, without the host touching a sample.
//
//   nm_sessions_subsample_device   Session.__init__ time-window subsampling        analysis/data.py:39-74
//   nm_streams_build_device        ConditioningExperiment.run transition extraction simulation/rl.py:85-102
//                                  + PositiveConditioning candidate next states     simulation/rl.py:177-182, 189
//                                  (what nm_stream_build does on the host, one session at a time)
//   nm_sessions_orientations_device compute_orientations                            analysis/metrics.py:653-694
//   nm_positions_adjust_device     adjust_positions_to_maze                         analysis/data.py:372-388
//
// Mapping.  Both jobs have a short sequential spine per session and an embarrassingly parallel bulk:
//   * subsampling: the window boundaries form a chain (each window starts where the previous one ended), so ONE THREAD
//     walks one session; sessions run in parallel (a recording is a few 10^4 samples, a launch takes ~1 ms);
//   * stream building: the heading of the algorithm's mouse at step i is the direction of the last real displacement
//     before it (Mouse.move_to keeps the pose when the new point isApprox the old one, mouse.cpp:75-87) -- again a
//     chain, walked by one thread per session (nm_pose_chain_kernel, 3 doubles per step); everything else -- nearest
//     visitable tile of both end points, eight rotated endpoints, tile lookups, the index of the observed candidate --
//     is independent per step and runs one THREAD PER RECORD (nm_records_kernel) over all sessions.
// cos / sin / atan2 are the host C library's (nm_libm.h: its algorithms restated with their fused operations), so
// headings, rotated endpoints and therefore candidate lists equal the host builder's bit for bit
// (tests/test_session_device_gpu.py compares whole streams).
#include <cuda_runtime.h>

#include <new>
#include <vector>

#include "nm_common.h"
#include "nm_libm.h"
#include "nm_pose.h"

namespace {

__global__ void nm_subsample_kernel(const double *__restrict__ time, const double *__restrict__ traj,
                                    const double *__restrict__ reward, const int64_t *__restrict__ offsets, int S,
                                    double dt, double *__restrict__ traj_out, double *__restrict__ reward_out,
                                    int64_t *__restrict__ counts) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= S) return;
  const int64_t base = offsets[j], n = offsets[j + 1] - base;
  const double *t = time + base, *p = traj + 2 * base;
  const double *r = reward ? reward + base : nullptr;
  double *po = traj_out + 2 * base, *ro = reward_out ? reward_out + base : nullptr;
  int64_t count = 0, idx = 0;
  while (idx < n - 1) {
    const double cur = t[idx];
    if (isnan(cur)) {  // data.py:45-47
      ++idx;
      continue;
    }
    const double limit = cur + dt;
    int64_t nxt = idx + 1;
    while (nxt < n && (isnan(t[nxt]) || t[nxt] <= limit)) ++nxt;  // :48-51
    int64_t k = idx;
    while (k < nxt && !(isfinite(p[2 * k]) && isfinite(p[2 * k + 1]))) ++k;  // first finite point (:54-56)
    if (k < nxt) {
      po[2 * count] = p[2 * k];
      po[2 * count + 1] = p[2 * k + 1];
      if (r) {  // first reward of the window that is neither 0 nor NaN, else 0.0 (:62-70)
        double rv = 0.0;
        for (int64_t q = idx; q < nxt; ++q) {
          const double v = r[q];
          if (v != 0.0 && !isnan(v)) {
            rv = v;
            break;
          }
        }
        ro[count] = rv;
      }
      ++count;
    }
    idx = nxt;
  }
  counts[j] = count;
}

struct BuildArgs {
  const double *traj, *reward;
  const int64_t *src_off, *src_cnt, *rec_off;  // device [S] / [S] / [S+1]
  const double *mouse_init;                    // device [S,3] or NULL
  int S, A;
  double act[NM_MAX_CAND][2];
  const uint8_t *tiles;
  const int32_t *state_of_tile;
  int X, Y;
  double st;
  double *pose;  // [total,3] scratch
  nm_step_rec *rec;
  int *status;  // [0] first failing global record + 1 (0: ok); [1] min ncand
};

// the algorithm's mouse walking through a session exactly like nm_stream_build drives it (rl.py:178-189)
__global__ void nm_pose_chain_kernel(const BuildArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.S) return;
  const int64_t n = a.src_cnt[j];
  if (n < 2) return;
  const double *p = a.traj + 2 * a.src_off[j];
  double *pose = a.pose + 3 * a.rec_off[j];
  double px = p[0], py = p[1], ori = 0.0;  // Mouse(trajectory[0], 0.0, actions) unless told otherwise
  if (a.mouse_init) {
    px = a.mouse_init[3 * j];
    py = a.mouse_init[3 * j + 1];
    ori = a.mouse_init[3 * j + 2];
  }
  auto move_to = [&](double x, double y) {  // mouse.cpp:75-87
    if (!nm_pos_is_approx(x, y, px, py)) {
      ori = nm_host_atan2(y - py, x - px);
      px = x;
      py = y;
    }
  };
  for (int64_t i = 1; i < n; ++i) {
    move_to(p[2 * (i - 1)], p[2 * (i - 1) + 1]);
    pose[3 * (i - 1)] = px;
    pose[3 * (i - 1) + 1] = py;
    pose[3 * (i - 1) + 2] = ori;
    move_to(p[2 * i], p[2 * i + 1]);
  }
}

__global__ void nm_records_kernel(const BuildArgs a, int64_t total) {
  extern __shared__ unsigned char sm_tiles[];
  const int nT = a.X * a.Y;
  for (int i = threadIdx.x; i < nT; i += blockDim.x) sm_tiles[i] = a.tiles[i];
  __syncthreads();
  NmMazeView mv;
  mv.tiles = sm_tiles;
  mv.X = a.X;
  mv.Y = a.Y;
  mv.st = a.st;
  mv.inv = 1.0 / a.st;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    // session of this record: binary search in rec_off
    int lo = 0, hi = a.S;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.rec_off[mid] <= g) lo = mid;
      else hi = mid;
    }
    const int64_t i = g - a.rec_off[lo] + 1;  // transition i-1 -> i of the session
    const double *p = a.traj + 2 * a.src_off[lo];
    int sx = 0, sy = 0, tx = 0, ty = 0;
    nm_closest_visitable(mv, p[2 * (i - 1)], p[2 * (i - 1) + 1], &sx, &sy);
    nm_closest_visitable(mv, p[2 * i], p[2 * i + 1], &tx, &ty);
    const int s = a.state_of_tile[sx * a.Y + sy], s1 = a.state_of_tile[tx * a.Y + ty];
    nm_step_rec rec;
    rec.r = a.reward[a.src_off[lo] + i];
    rec.s = (uint16_t)s;
    rec.s1 = (uint16_t)s1;
    rec.ncand = 0;
    rec.obs = NM_OBS_NONE;
    rec.reserved = 0;
#pragma unroll
    for (int k = 0; k < NM_MAX_CAND; ++k) rec.cand[k] = 0;
    if (a.A > 0) {
      const double px = a.pose[3 * g], py = a.pose[3 * g + 1], ori = a.pose[3 * g + 2];
      double c, sn;
      nm_host_sincos(ori, &sn, &c);
      for (int k = 0; k < a.A; ++k) {
        double ex, ey;
        nm_pose_endpoint(c, sn, px, py, a.act[k][0], a.act[k][1], &ex, &ey);
        const int ix = nm_tile_index_inv(ex, mv.st, mv.inv), iy = nm_tile_index_inv(ey, mv.st, mv.inv);
        if (nm_tile_visitable(mv, ix, iy)) {  // rl.py:180-182
          const int cs = a.state_of_tile[ix * a.Y + iy];
          if (rec.obs == NM_OBS_NONE && cs == s1) rec.obs = rec.ncand;
          rec.cand[rec.ncand++] = (uint16_t)cs;
        }
      }
      if (rec.ncand == 0) atomicMin(&a.status[0], (int)(g < 2147483646LL ? g + 1 : 2147483647LL));
      atomicMin(&a.status[1], (int)rec.ncand);
    }
    a.rec[g] = rec;
  }
}

// compute_orientations (analysis/metrics.py:653-694) for S sessions: one thread walks one session (prev_ori chains
// through the moved steps).  out[j] = heading of step j-1; the entries up to and including the first moved step
// repeat its heading (":694"); first_moved[j] = index of that step, -1 if the session never moves (the reference
// raises IndexError there).
struct OriArgs {
  const double *traj;
  const int64_t *off, *cnt;
  int S, A;  // A = 0: no maze / action filter
  double act[NM_MAX_CAND][2];
  const uint8_t *tiles;
  int X, Y;
  double st;
  double *out;
  int64_t *first_moved;
};

__device__ __forceinline__ bool nm_np_isclose(double a, double b) {  // numpy.isclose defaults (rtol 1e-5, atol 1e-8)
  if (isfinite(a) && isfinite(b)) return fabs(a - b) <= 1e-8 + 1e-5 * fabs(b);
  return a == b;
}

__global__ void nm_orientations_kernel(const OriArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.S) return;
  const int64_t n = a.cnt[j];
  const double *p = a.traj + 2 * a.off[j];
  double *out = a.out + a.off[j];
  const double inv = 1.0 / a.st;
  double prev = 0.0, last = 0.0;
  int64_t first = -1;
  for (int64_t i = 0; i + 1 < n; ++i) {
    const double x1 = p[2 * i], y1 = p[2 * i + 1], x2 = p[2 * i + 2], y2 = p[2 * i + 3];
    bool moved = !(nm_np_isclose(x1, x2) && nm_np_isclose(y1, y2));
    if (moved && a.A > 0) {  // the tile reached must be the tile under one of the rotated endpoints (:668-671)
      const int tx = nm_tile_index_inv(x2, a.st, inv), ty = nm_tile_index_inv(y2, a.st, inv);
      double c, sn;
      nm_host_sincos(prev, &sn, &c);
      bool hit = false;
      for (int k = 0; k < a.A; ++k) {
        double ex, ey;
        nm_pose_endpoint(c, sn, x1, y1, a.act[k][0], a.act[k][1], &ex, &ey);
        hit = hit || (nm_tile_index_inv(ex, a.st, inv) == tx && nm_tile_index_inv(ey, a.st, inv) == ty);
      }
      moved = hit;
    }
    if (moved) {
      last = nm_host_atan2(y2 - y1, x2 - x1);  // always inside [-pi, pi]: the corrections of :676-679 never fire
      prev = last;
      if (first < 0) {
        first = i;
        for (int64_t q = 0; q <= i; ++q) out[q] = last;
      }
    }
    if (first >= 0) out[i + 1] = last;
  }
  if (first < 0)
    for (int64_t q = 0; q < n; ++q) out[q] = nan("");
  a.first_moved[j] = first;
}

// adjust_positions_to_maze (analysis/data.py:372-388): every point outside the visitable tiles moves to the centre of
// the nearest visitable tile; NaN points stay.  One thread per point.
__global__ void nm_adjust_kernel(const double *pts, int64_t n, const uint8_t *__restrict__ tiles, int X, int Y,
                                 double st, double *out) {  // out may alias pts (a thread rewrites its own point)
  extern __shared__ unsigned char sm_tiles[];
  for (int i = threadIdx.x; i < X * Y; i += blockDim.x) sm_tiles[i] = tiles[i];
  __syncthreads();
  NmMazeView mv;
  mv.tiles = sm_tiles;
  mv.X = X;
  mv.Y = Y;
  mv.st = st;
  mv.inv = 1.0 / st;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (int64_t)gridDim.x * blockDim.x) {
    double x = pts[2 * g], y = pts[2 * g + 1];
    if (!isnan(x) && !isnan(y) && !nm_free_at(mv, x, y)) {
      int tx = 0, ty = 0;
      if (nm_closest_visitable(mv, x, y, &tx, &ty)) {
        x = nm_tile_centre(tx, st);
        y = nm_tile_centre(ty, st);
      }
    }
    out[2 * g] = x;
    out[2 * g + 1] = y;
  }
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// dst = online records of every session followed by its replay passes: dst[i] = src[map[i]]
__global__ void __launch_bounds__(256) nm_gather_records_kernel(const uint4 *__restrict__ src, const int64_t *__restrict__ map,
                                                                int64_t n, uint4 *__restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = map[i];
    dst[2 * i] = src[2 * j];
    dst[2 * i + 1] = src[2 * j + 1];
  }
}

}  // namespace

extern "C" int nm_sessions_subsample_device(const double *d_time, const double *d_traj, const double *d_reward,
                                            const int64_t *d_offsets, int S, double sampling_time, double *d_traj_out,
                                            double *d_reward_out, int64_t *d_counts, void *stream) {
  NM_CHECK_ARG(S >= 0 && (S == 0 || (d_time && d_traj && d_offsets && d_traj_out && d_counts)),
               "nm_sessions_subsample_device: NULL argument");
  NM_CHECK_ARG(!d_reward == !d_reward_out, "nm_sessions_subsample_device: give both reward arrays or neither");
  if (S == 0) return NM_OK;
  nm_subsample_kernel<<<(S + 63) / 64, 64, 0, (cudaStream_t)stream>>>(d_time, d_traj, d_reward, d_offsets, S,
                                                                      sampling_time, d_traj_out, d_reward_out, d_counts);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_subsample_kernel launch: %s", cudaGetErrorString(e));
  return NM_OK;
}

extern "C" int nm_sessions_orientations_device(const nm_maze *mz, const double *d_traj, const int64_t *d_offsets,
                                               const int64_t *d_counts, int S, const double *h_actions, int n_actions,
                                               double *d_ori_out, int64_t *d_first_moved, void *stream) {
  NM_CHECK_ARG(S >= 0 && (S == 0 || (d_traj && d_offsets && d_counts && d_ori_out && d_first_moved)),
               "nm_sessions_orientations_device: NULL argument");
  NM_CHECK_ARG(n_actions >= 0 && n_actions <= NM_MAX_CAND, "nm_sessions_orientations_device: bad number of actions");
  NM_CHECK_ARG(n_actions == 0 || (mz && mz->device >= 0 && h_actions),
               "nm_sessions_orientations_device: the action filter needs a device maze and the actions");
  if (S == 0) return NM_OK;
  OriArgs a;
  a.traj = d_traj;
  a.off = d_offsets;
  a.cnt = d_counts;
  a.S = S;
  a.A = n_actions;
  for (int k = 0; k < NM_MAX_CAND; ++k) {
    a.act[k][0] = k < n_actions ? h_actions[2 * k] : 0.0;
    a.act[k][1] = k < n_actions ? h_actions[2 * k + 1] : 0.0;
  }
  a.tiles = n_actions ? mz->d_tiles : nullptr;
  a.X = n_actions ? mz->X : 0;
  a.Y = n_actions ? mz->Y : 0;
  a.st = n_actions ? mz->size_tile : 1.0;
  a.out = d_ori_out;
  a.first_moved = d_first_moved;
  nm_orientations_kernel<<<(S + 31) / 32, 32, 0, (cudaStream_t)stream>>>(a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_orientations_kernel launch: %s", cudaGetErrorString(e));
  return NM_OK;
}

extern "C" int nm_positions_adjust_device(const nm_maze *mz, const double *d_points, int64_t n, double *d_out,
                                          void *stream) {
  NM_CHECK_ARG(mz && mz->device >= 0, "nm_positions_adjust_device: the maze must live on a CUDA device");
  NM_CHECK_ARG(n >= 0 && (n == 0 || (d_points && d_out)), "nm_positions_adjust_device: NULL argument");
  if (n == 0) return NM_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  nm_adjust_kernel<<<blocks, 256, (size_t)mz->X * mz->Y, (cudaStream_t)stream>>>(d_points, n, mz->d_tiles, mz->X, mz->Y,
                                                                               mz->size_tile, d_out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_adjust_kernel launch: %s", cudaGetErrorString(e));
  return NM_OK;
}

extern "C" nm_streams *nm_streams_build_device(const nm_maze *mz, const double *d_traj, const double *d_reward,
                                               const int64_t *h_offsets, const int64_t *h_counts, int S,
                                               const double *h_actions, int n_actions, const double *h_mouse_init,
                                               void *stream) {
  if (!mz || mz->device < 0 || !d_traj || !d_reward || !h_offsets || !h_counts || S <= 0 || n_actions < 0 ||
      n_actions > NM_MAX_CAND || (n_actions > 0 && !h_actions)) {
    nm_fail(NM_ERR_INVALID, "nm_streams_build_device: bad argument (the maze must live on a CUDA device)");
    return nullptr;
  }
  cudaStream_t cs = (cudaStream_t)stream;
  nm_streams *st = new (std::nothrow) nm_streams();
  if (!st) {
    nm_fail(NM_ERR_NOMEM, "nm_streams_build_device: out of memory");
    return nullptr;
  }
  st->S = S;
  st->num_states = mz->num_states;
  st->device = mz->device;
  st->offsets.assign(S + 1, 0);
  st->online.assign(S, 0);
  for (int j = 0; j < S; ++j) {
    if (h_counts[j] < 1) {
      nm_fail(NM_ERR_INVALID, "nm_streams_build_device: session %d is empty", j);
      delete st;
      return nullptr;
    }
    st->online[j] = h_counts[j] - 1;
    st->offsets[j + 1] = st->offsets[j] + st->online[j];
  }
  const int64_t total = st->offsets[S];
  st->total = total;
  char *base = nullptr, *scratch = nullptr;
  int h_status[2] = {2147483647, NM_MAX_CAND};
  cudaError_t e = cudaSetDevice(mz->device);
  const size_t o_rec = 0, o_off = align256(sizeof(nm_step_rec) * (size_t)(total + 1)),
               o_onl = o_off + align256(sizeof(int64_t) * (S + 1)), bytes = o_onl + align256(sizeof(int64_t) * S);
  // scratch: pose [total,3] | src_off [S] | src_cnt [S] | mouse_init [S,3] | status [2]
  const size_t s_pose = 0, s_off = align256(sizeof(double) * 3 * (size_t)(total + 1)),
               s_cnt = s_off + align256(sizeof(int64_t) * S), s_init = s_cnt + align256(sizeof(int64_t) * S),
               s_stat = s_init + align256(sizeof(double) * 3 * S), s_bytes = s_stat + 256;
  if (e == cudaSuccess) e = cudaMalloc((void **)&base, bytes);
  if (e == cudaSuccess) e = cudaMalloc((void **)&scratch, s_bytes);
  if (e == cudaSuccess) {
    st->d_block = base;
    st->d_records = (nm_step_rec *)(base + o_rec);
    st->d_offsets = (int64_t *)(base + o_off);
    st->d_online = (int64_t *)(base + o_onl);
    e = cudaMemcpyAsync(st->d_offsets, st->offsets.data(), sizeof(int64_t) * (S + 1), cudaMemcpyHostToDevice, cs);
  }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(st->d_online, st->online.data(), sizeof(int64_t) * S, cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess) e = cudaMemcpyAsync(scratch + s_off, h_offsets, sizeof(int64_t) * S, cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess) e = cudaMemcpyAsync(scratch + s_cnt, h_counts, sizeof(int64_t) * S, cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess && h_mouse_init)
    e = cudaMemcpyAsync(scratch + s_init, h_mouse_init, sizeof(double) * 3 * S, cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess) e = cudaMemcpyAsync(scratch + s_stat, h_status, sizeof(h_status), cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess && total > 0) {
    BuildArgs a;
    a.traj = d_traj;
    a.reward = d_reward;
    a.src_off = (const int64_t *)(scratch + s_off);
    a.src_cnt = (const int64_t *)(scratch + s_cnt);
    a.rec_off = st->d_offsets;
    a.mouse_init = h_mouse_init ? (const double *)(scratch + s_init) : nullptr;
    a.S = S;
    a.A = n_actions;
    for (int k = 0; k < NM_MAX_CAND; ++k) {
      a.act[k][0] = k < n_actions ? h_actions[2 * k] : 0.0;
      a.act[k][1] = k < n_actions ? h_actions[2 * k + 1] : 0.0;
    }
    a.tiles = mz->d_tiles;
    a.state_of_tile = mz->d_state_of_tile;
    a.X = mz->X;
    a.Y = mz->Y;
    a.st = mz->size_tile;
    a.pose = (double *)(scratch + s_pose);
    a.rec = st->d_records;
    a.status = (int *)(scratch + s_stat);
    if (n_actions > 0) nm_pose_chain_kernel<<<(S + 31) / 32, 32, 0, cs>>>(a);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    nm_records_kernel<<<blocks, 256, (size_t)mz->X * mz->Y, cs>>>(a, total);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_status, scratch + s_stat, sizeof(h_status), cudaMemcpyDeviceToHost, cs);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(cs);  // setup call: the status word decides success
  if (scratch) cudaFree(scratch);
  if (e != cudaSuccess) {
    nm_fail(NM_ERR_CUDA, "nm_streams_build_device: %s", cudaGetErrorString(e));
    if (base) cudaFree(base);
    delete st;
    return nullptr;
  }
  if (h_status[0] != 2147483647) {  // np.max([]) raises in the reference (rl.py:184)
    nm_fail(NM_ERR_INVALID,
            "zero-size array to reduction operation maximum which has no identity (record %d has no visitable candidate)",
            h_status[0] - 1);
    cudaFree(base);
    delete st;
    return nullptr;
  }
  st->min_ncand = (n_actions > 0 && total > 0) ? h_status[1] : 0;
  return st;
}

extern "C" int nm_streams_download(const nm_streams *st, nm_step_rec *h_records, void *stream) {
  NM_CHECK_ARG(st && (st->total == 0 || h_records), "nm_streams_download: NULL argument");
  if (st->total == 0) return NM_OK;
  if (cudaSetDevice(st->device) != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaSetDevice(%d) failed", st->device);
  cudaError_t e = cudaMemcpyAsync(h_records, st->d_records, sizeof(nm_step_rec) * (size_t)st->total,
                                  cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_streams_download: %s", cudaGetErrorString(e));
  return NM_OK;
}

extern "C" nm_streams *nm_streams_with_replays(const nm_streams *src, const int64_t *h_order, const int64_t *h_counts,
                                               void *stream) {
  if (!src || src->device < 0 || !h_counts || src->S <= 0) {
    nm_fail(NM_ERR_INVALID, "nm_streams_with_replays: bad argument (the streams must live on a CUDA device)");
    return nullptr;
  }
  const int S = src->S;
  nm_streams *st = new (std::nothrow) nm_streams();
  if (!st) {
    nm_fail(NM_ERR_NOMEM, "nm_streams_with_replays: out of memory");
    return nullptr;
  }
  st->S = S;
  st->num_states = src->num_states;
  st->device = src->device;
  st->min_ncand = src->min_ncand;
  st->offsets.assign(S + 1, 0);
  st->online.assign(S, 0);
  std::vector<int64_t> map;  // global source index of every destination record
  int64_t cursor = 0;
  for (int j = 0; j < S; ++j) {
    const int64_t n_on = src->online[j], base = src->offsets[j], extra = h_counts[j];
    if (extra < 0 || (extra > 0 && !h_order)) {
      nm_fail(NM_ERR_INVALID, "nm_streams_with_replays: bad replay count for session %d", j);
      delete st;
      return nullptr;
    }
    st->online[j] = n_on;
    st->offsets[j + 1] = st->offsets[j] + n_on + extra;
    for (int64_t i = 0; i < n_on; ++i) map.push_back(base + i);
    for (int64_t i = 0; i < extra; ++i) {
      const int64_t k = h_order[cursor + i];
      if (k < 0 || k >= n_on) {
        nm_fail(NM_ERR_INVALID, "nm_streams_with_replays: replay index %lld of session %d is out of range", (long long)k, j);
        delete st;
        return nullptr;
      }
      map.push_back(base + k);
    }
    cursor += extra;
  }
  const int64_t total = st->offsets[S];
  st->total = total;
  cudaStream_t cs = (cudaStream_t)stream;
  char *base = nullptr;
  int64_t *d_map = nullptr;
  cudaError_t e = cudaSetDevice(src->device);
  const size_t o_off = align256(sizeof(nm_step_rec) * (size_t)(total + 1)),
               o_onl = o_off + align256(sizeof(int64_t) * (S + 1)), bytes = o_onl + align256(sizeof(int64_t) * S);
  if (e == cudaSuccess) e = cudaMalloc((void **)&base, bytes);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_map, sizeof(int64_t) * (size_t)(total + 1));
  if (e == cudaSuccess) {
    st->d_block = base;
    st->d_records = (nm_step_rec *)base;
    st->d_offsets = (int64_t *)(base + o_off);
    st->d_online = (int64_t *)(base + o_onl);
    e = cudaMemcpyAsync(st->d_offsets, st->offsets.data(), sizeof(int64_t) * (S + 1), cudaMemcpyHostToDevice, cs);
  }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(st->d_online, st->online.data(), sizeof(int64_t) * S, cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess && total > 0) {
    e = cudaMemcpyAsync(d_map, map.data(), sizeof(int64_t) * (size_t)total, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess) {
      int blocks = (int)((total + 255) / 256);
      if (blocks > 148 * 8) blocks = 148 * 8;
      nm_gather_records_kernel<<<blocks, 256, 0, cs>>>((const uint4 *)src->d_records, d_map, total, (uint4 *)st->d_records);
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(cs);  // setup call; `map` and d_map die here
  if (d_map) cudaFree(d_map);
  if (e != cudaSuccess) {
    nm_fail(NM_ERR_CUDA, "nm_streams_with_replays: %s", cudaGetErrorString(e));
    if (base) cudaFree(base);
    delete st;
    return nullptr;
  }
  return st;
}
