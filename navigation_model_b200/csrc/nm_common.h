// Shared host-side helpers of the C-ABI library: thread-local error message, handle layouts.
#ifndef NM_COMMON_H
#define NM_COMMON_H

#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "../../include/navmodel_b200.h"

// --- thread-local error state (nm_last_error) ---------------------------------------------------
char *nm_error_buffer();  // 512 bytes, thread-local
int nm_fail(int code, const char *fmt, ...);

#define NM_CHECK_ARG(cond, ...)                                \
  do {                                                         \
    if (!(cond)) return nm_fail(NM_ERR_INVALID, __VA_ARGS__);  \
  } while (0)

// --- handles ------------------------------------------------------------------------------------
struct nm_mouse {
  double init_pos[2];
  double pos[2];
  double init_ori;
  double ori;
  std::vector<double> actions;  // [A,2]
  bool save_history;
  std::vector<double> hist_pos;  // [n,2]
  std::vector<double> hist_ori;  // [n]
};
// mouse.cpp:75-87 (defined in nm_mouse.cpp); returns 1 if the mouse moved
int nm_mouse_move_impl(nm_mouse *m, double x, double y);

struct nm_maze {
  int X = 0, Y = 0;
  double size_tile = 0.0;
  int num_states = 0;
  int device = -1;  // < 0: host-only
  std::vector<uint8_t> tiles, subareas;       // host copies
  std::vector<int32_t> state_of_tile;         // [X*Y] compact id or -1
  std::vector<int32_t> tile_of_state;         // [NV]
  // device copies (one allocation)
  uint8_t *d_tiles = nullptr;
  uint8_t *d_subareas = nullptr;
  int32_t *d_state_of_tile = nullptr;
  int32_t *d_tile_of_state = nullptr;
  void *d_block = nullptr;
};

struct nm_streams {
  int S = 0;
  int num_states = 0;
  int device = -1;
  int64_t total = 0;
  int min_ncand = NM_MAX_CAND;   // smallest candidate count over all records (0: value-only TD0 streams)
  std::vector<int64_t> offsets;  // [S+1]
  std::vector<int64_t> online;   // [S]
  nm_step_rec *d_records = nullptr;
  int64_t *d_offsets = nullptr;
  int64_t *d_online = nullptr;
  void *d_block = nullptr;
};

// frees the grow-only device scratch of the Q-table sweep (nm_q.cu); called by nm_scratch_release (nm_fit.cu)
void nm_q_scratch_release();

#endif  // NM_COMMON_H
