// Maze and step-stream handles: host copies + device-resident dense tables (HBM).
//
// Layout in HBM (one allocation per handle, 256-byte aligned sub-buffers):
//   nm_maze    : tiles u8[X*Y] | subareas u8[X*Y] | state_of_tile i32[X*Y] | tile_of_state i32[NV]
//   nm_streams : records nm_step_rec[total] (32 B each) | offsets i64[S+1] | online i64[S]
// The tables are tiny (a 9x9 maze is 81 B; 100 sessions x 5 000 steps x 32 B = 16 MB) and are read by
// every CTA, so they live in L2; the kernels stage them into shared memory.
#include <cuda_runtime.h>
#include <string.h>

#include <new>

#include "nm_common.h"
#include "nm_geometry.h"

#define NM_CUDA_OK(call, what)                                                              \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      nm_fail(NM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e_));                         \
      goto fail;                                                                            \
    }                                                                                       \
  } while (0)

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int nm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" nm_maze *nm_maze_upload(const uint8_t *h_tiles, const uint8_t *h_subareas, int X, int Y,
                                   double size_tile, int device, void *stream) {
  if (!h_tiles || X <= 0 || Y <= 0 || !(size_tile > 0.0) || (int64_t)X * Y > 65535) {
    nm_fail(NM_ERR_INVALID, "nm_maze_upload: need tiles, 0 < X*Y <= 65535 and size_tile > 0");
    return nullptr;
  }
  nm_maze *mz = new (std::nothrow) nm_maze();
  if (!mz) {
    nm_fail(NM_ERR_NOMEM, "nm_maze_upload: out of memory");
    return nullptr;
  }
  const int n = X * Y;
  mz->X = X;
  mz->Y = Y;
  mz->size_tile = size_tile;
  mz->device = device;
  mz->tiles.assign(h_tiles, h_tiles + n);
  if (h_subareas) mz->subareas.assign(h_subareas, h_subareas + n);
  mz->state_of_tile.assign(n, -1);
  for (int i = 0; i < n; ++i)
    if (h_tiles[i] != 0) {
      mz->state_of_tile[i] = (int32_t)mz->tile_of_state.size();
      mz->tile_of_state.push_back(i);
    }
  mz->num_states = (int)mz->tile_of_state.size();
  if (device >= 0) {
    cudaStream_t cs = (cudaStream_t)stream;
    const size_t o_tiles = 0, o_sub = align256(n), o_sot = o_sub + align256(n),
                 o_tos = o_sot + align256(sizeof(int32_t) * n),
                 total = o_tos + align256(sizeof(int32_t) * (mz->num_states + 1));
    char *base = nullptr;
    NM_CUDA_OK(cudaSetDevice(device), "nm_maze_upload: cudaSetDevice");
    NM_CUDA_OK(cudaMalloc((void **)&base, total), "nm_maze_upload: cudaMalloc");
    mz->d_block = base;
    mz->d_tiles = (uint8_t *)(base + o_tiles);
    mz->d_subareas = h_subareas ? (uint8_t *)(base + o_sub) : nullptr;
    mz->d_state_of_tile = (int32_t *)(base + o_sot);
    mz->d_tile_of_state = (int32_t *)(base + o_tos);
    NM_CUDA_OK(cudaMemcpyAsync(mz->d_tiles, mz->tiles.data(), n, cudaMemcpyHostToDevice, cs), "maze tiles H2D");
    if (h_subareas)
      NM_CUDA_OK(cudaMemcpyAsync(mz->d_subareas, mz->subareas.data(), n, cudaMemcpyHostToDevice, cs),
                 "maze subareas H2D");
    NM_CUDA_OK(cudaMemcpyAsync(mz->d_state_of_tile, mz->state_of_tile.data(), sizeof(int32_t) * n,
                               cudaMemcpyHostToDevice, cs), "maze state_of_tile H2D");
    if (mz->num_states > 0)
      NM_CUDA_OK(cudaMemcpyAsync(mz->d_tile_of_state, mz->tile_of_state.data(),
                                 sizeof(int32_t) * mz->num_states, cudaMemcpyHostToDevice, cs),
                 "maze tile_of_state H2D");
    // the host vectors are owned by the handle, so the async copies stay valid
  }
  return mz;
fail:
  if (mz->d_block) cudaFree(mz->d_block);
  delete mz;
  return nullptr;
}

extern "C" void nm_maze_free(nm_maze *mz) {
  if (!mz) return;
  if (mz->d_block) {
    cudaSetDevice(mz->device);
    cudaFree(mz->d_block);
  }
  delete mz;
}

extern "C" int nm_maze_num_states(const nm_maze *mz) {
  NM_CHECK_ARG(mz, "nm_maze_num_states: NULL handle");
  return mz->num_states;
}

static inline NmMazeView host_view(const nm_maze *mz) {
  NmMazeView v;
  v.tiles = mz->tiles.data();
  v.X = mz->X;
  v.Y = mz->Y;
  v.st = mz->size_tile;
  v.inv = 1.0 / mz->size_tile;
  return v;
}

extern "C" int nm_maze_cont2disc(const nm_maze *mz, double x, double y, int *dx, int *dy) {
  NM_CHECK_ARG(mz && dx && dy, "nm_maze_cont2disc: NULL argument");
  *dx = nm_tile_index(x, mz->size_tile);
  *dy = nm_tile_index(y, mz->size_tile);
  return NM_OK;
}

extern "C" int nm_maze_movement_possible_cont(const nm_maze *mz, double x1, double y1, double x2, double y2) {
  NM_CHECK_ARG(mz, "nm_maze_movement_possible_cont: NULL handle");
  return nm_move_possible(host_view(mz), x1, y1, x2, y2) ? 1 : 0;
}

extern "C" int nm_maze_closest_visitable_cont(const nm_maze *mz, double x, double y, int *dx, int *dy) {
  NM_CHECK_ARG(mz && dx && dy, "nm_maze_closest_visitable_cont: NULL argument");
  if (!nm_closest_visitable(host_view(mz), x, y, dx, dy))
    return nm_fail(NM_ERR_RUNTIME, "maze has no visitable tile");
  return NM_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" nm_streams *nm_streams_upload(const nm_step_rec *h_records, const int64_t *h_offsets,
                                         const int64_t *h_online, int S, int num_states, int device,
                                         void *stream) {
  if (!h_offsets || !h_online || S <= 0 || num_states <= 0 || num_states > 65535 || device < 0) {
    nm_fail(NM_ERR_INVALID, "nm_streams_upload: bad argument");
    return nullptr;
  }
  if (h_offsets[0] != 0) {
    nm_fail(NM_ERR_INVALID, "nm_streams_upload: offsets[0] must be 0");
    return nullptr;
  }
  for (int j = 0; j < S; ++j) {
    const int64_t len = h_offsets[j + 1] - h_offsets[j];
    if (len < 0 || h_online[j] < 0 || h_online[j] > len) {
      nm_fail(NM_ERR_INVALID, "nm_streams_upload: session %d: need 0 <= online <= length", j);
      return nullptr;
    }
  }
  const int64_t total = h_offsets[S];
  if (total > 0 && !h_records) {
    nm_fail(NM_ERR_INVALID, "nm_streams_upload: NULL records");
    return nullptr;
  }
  int min_ncand = NM_MAX_CAND;
  for (int64_t i = 0; i < total; ++i) {
    const nm_step_rec &r = h_records[i];
    if (r.ncand < min_ncand) min_ncand = r.ncand;
    bool ok = r.s < num_states && r.s1 < num_states && r.ncand <= NM_MAX_CAND &&
              (r.obs == NM_OBS_NONE || r.obs < r.ncand);
    for (int k = 0; ok && k < r.ncand; ++k) ok = r.cand[k] < num_states;
    if (!ok) {
      nm_fail(NM_ERR_INVALID, "nm_streams_upload: record %lld has an out-of-range state / candidate",
              (long long)i);
      return nullptr;
    }
  }
  nm_streams *st = new (std::nothrow) nm_streams();
  if (!st) {
    nm_fail(NM_ERR_NOMEM, "nm_streams_upload: out of memory");
    return nullptr;
  }
  st->S = S;
  st->num_states = num_states;
  st->device = device;
  st->total = total;
  st->min_ncand = min_ncand;
  st->offsets.assign(h_offsets, h_offsets + S + 1);
  st->online.assign(h_online, h_online + S);
  {
    cudaStream_t cs = (cudaStream_t)stream;
    // +1 record of padding so that 16-byte bulk copies of the last chunk never run past the buffer
    const size_t o_rec = 0, o_off = align256(sizeof(nm_step_rec) * (size_t)(total + 1)),
                 o_onl = o_off + align256(sizeof(int64_t) * (S + 1)),
                 bytes = o_onl + align256(sizeof(int64_t) * S);
    char *base = nullptr;
    NM_CUDA_OK(cudaSetDevice(device), "nm_streams_upload: cudaSetDevice");
    NM_CUDA_OK(cudaMalloc((void **)&base, bytes), "nm_streams_upload: cudaMalloc");
    st->d_block = base;
    st->d_records = (nm_step_rec *)(base + o_rec);
    st->d_offsets = (int64_t *)(base + o_off);
    st->d_online = (int64_t *)(base + o_onl);
    if (total > 0)
      NM_CUDA_OK(cudaMemcpyAsync(st->d_records, h_records, sizeof(nm_step_rec) * (size_t)total,
                                 cudaMemcpyHostToDevice, cs), "records H2D");
    NM_CUDA_OK(cudaMemcpyAsync(st->d_offsets, st->offsets.data(), sizeof(int64_t) * (S + 1),
                               cudaMemcpyHostToDevice, cs), "offsets H2D");
    NM_CUDA_OK(cudaMemcpyAsync(st->d_online, st->online.data(), sizeof(int64_t) * S, cudaMemcpyHostToDevice, cs),
               "online H2D");
  }
  return st;
fail:
  if (st->d_block) cudaFree(st->d_block);
  delete st;
  return nullptr;
}

extern "C" int nm_streams_update(nm_streams *st, const nm_step_rec *h_records, void *stream) {
  NM_CHECK_ARG(st && (st->total == 0 || h_records), "nm_streams_update: NULL argument");
  int min_ncand = NM_MAX_CAND;
  for (int64_t i = 0; i < st->total; ++i) {
    const nm_step_rec &r = h_records[i];
    if (r.ncand < min_ncand) min_ncand = r.ncand;
    bool ok = r.s < st->num_states && r.s1 < st->num_states && r.ncand <= NM_MAX_CAND &&
              (r.obs == NM_OBS_NONE || r.obs < r.ncand);
    for (int k = 0; ok && k < r.ncand; ++k) ok = r.cand[k] < st->num_states;
    if (!ok)
      return nm_fail(NM_ERR_INVALID, "nm_streams_update: record %lld has an out-of-range state / candidate",
                     (long long)i);
  }
  st->min_ncand = min_ncand;
  if (st->total == 0) return NM_OK;
  if (cudaSetDevice(st->device) != cudaSuccess) return nm_fail(NM_ERR_CUDA, "cudaSetDevice(%d) failed", st->device);
  cudaError_t e = cudaMemcpyAsync(st->d_records, h_records, sizeof(nm_step_rec) * (size_t)st->total,
                                  cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_streams_update: %s", cudaGetErrorString(e));
  return NM_OK;
}

extern "C" void nm_streams_free(nm_streams *st) {
  if (!st) return;
  if (st->d_block) {
    cudaSetDevice(st->device);
    cudaFree(st->d_block);
  }
  delete st;
}
