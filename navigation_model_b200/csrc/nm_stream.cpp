// Host-side extraction of the parameter-independent step stream of a conditioning session.
//
// Restates, for a whole trajectory at once, what the reference does one Python object at a time:
//   ConditioningExperiment.run        simulation/rl.py:85-102  (s, s1 = closest visitable tile, r = reward[i])
//   PositiveConditioning.update       simulation/rl.py:177-182 (the algorithm's own Mouse is moved to the
//                                     start point, its endpoints are discretised and filtered by TILE
//                                     visitability -- not by the ray check), :189 (then moved to the arrival)
// The stream does not depend on (alpha, beta, gamma), so it is built once and broadcast to every
// parameter set on the device.
#include <math.h>

#include "nm_common.h"
#include "nm_pose.h"

namespace {

// One transition (cx, cy) -> (nx, ny) with known compact states s, s1 and reward r: the record, and -- when the
// algorithm's mouse is given -- its candidate next states, driving the mouse like PositiveConditioning.update does
// (rl.py:177-182, 189).  Returns false (error set) when no candidate is visitable.
bool build_record(const nm_maze *mz, const NmMazeView &mv, nm_mouse *mouse, const double *actions, int A, double cx,
                  double cy, double nx, double ny, int s, int s1, double r, long long step, nm_step_rec *out) {
  nm_step_rec rec;
  rec.r = r;
  rec.s = (uint16_t)s;
  rec.s1 = (uint16_t)s1;
  rec.ncand = 0;
  rec.obs = NM_OBS_NONE;
  rec.reserved = 0;
  for (int k = 0; k < NM_MAX_CAND; ++k) rec.cand[k] = 0;
  if (mouse) {
    nm_mouse_move_impl(mouse, cx, cy);  // rl.py:178 (history is pushed like the reference's)
    const double px = mouse->pos[0], py = mouse->pos[1];
    const double c = cos(mouse->ori), sn = sin(mouse->ori);
    for (int a = 0; a < A; ++a) {
      double ex, ey;
      nm_pose_endpoint(c, sn, px, py, actions[2 * a], actions[2 * a + 1], &ex, &ey);
      const int ix = nm_tile_index(ex, mv.st), iy = nm_tile_index(ey, mv.st);
      if (nm_tile_visitable(mv, ix, iy)) {  // rl.py:180-182
        const int cs = mz->state_of_tile[ix * mz->Y + iy];
        if (rec.obs == NM_OBS_NONE && cs == s1) rec.obs = rec.ncand;
        rec.cand[rec.ncand++] = (uint16_t)cs;
      }
    }
    if (rec.ncand == 0) {  // np.max([]) raises in the reference (rl.py:184)
      nm_fail(NM_ERR_INVALID,
              "zero-size array to reduction operation maximum which has no identity "
              "(step %lld has no visitable candidate)", step);
      return false;
    }
    nm_mouse_move_impl(mouse, nx, ny);  // rl.py:189
  }
  *out = rec;
  return true;
}

int check_mouse(const nm_mouse *mouse, const char *who, int *A) {
  *A = mouse ? (int)(mouse->actions.size() / 2) : 0;
  if (mouse) NM_CHECK_ARG(*A >= 1 && *A <= NM_MAX_CAND, "%s: need a mouse with 1..%d relative actions (has %d)", who,
                          NM_MAX_CAND, *A);
  return NM_OK;
}

NmMazeView view_of(const nm_maze *mz) {
  NmMazeView mv;
  mv.tiles = mz->tiles.data();
  mv.X = mz->X;
  mv.Y = mz->Y;
  mv.st = mz->size_tile;
  mv.inv = 1.0 / mz->size_tile;
  return mv;
}

}  // namespace

extern "C" int64_t nm_stream_build(const nm_maze *mz, const double *trajectory, const double *reward, int64_t T,
                                   nm_mouse *mouse, nm_step_rec *out) {
  NM_CHECK_ARG(mz && trajectory && reward && T >= 1, "nm_stream_build: NULL argument or empty trajectory");
  NM_CHECK_ARG(T == 1 || out, "nm_stream_build: NULL output");
  int A;
  if (check_mouse(mouse, "nm_stream_build", &A) != NM_OK) return NM_ERR_INVALID;
  const NmMazeView mv = view_of(mz);
  const double *actions = mouse ? mouse->actions.data() : nullptr;
  int sx, sy;
  if (!nm_closest_visitable(mv, trajectory[0], trajectory[1], &sx, &sy))
    return nm_fail(NM_ERR_RUNTIME, "maze has no visitable tile");
  int s = mz->state_of_tile[sx * mz->Y + sy];
  for (int64_t i = 1; i < T; ++i) {
    int tx = 0, ty = 0;
    nm_closest_visitable(mv, trajectory[2 * i], trajectory[2 * i + 1], &tx, &ty);
    const int s1 = mz->state_of_tile[tx * mz->Y + ty];
    if (!build_record(mz, mv, mouse, actions, A, trajectory[2 * (i - 1)], trajectory[2 * (i - 1) + 1], trajectory[2 * i],
                      trajectory[2 * i + 1], s, s1, reward[i], (long long)i, &out[i - 1]))
      return NM_ERR_INVALID;
    s = s1;
  }
  return T - 1;
}

extern "C" int64_t nm_stream_build_pairs(const nm_maze *mz, const double *starts, const double *arrivals,
                                         const int32_t *start_tiles, const int32_t *arrival_tiles, const double *reward,
                                         int64_t n, nm_mouse *mouse, nm_step_rec *out) {
  NM_CHECK_ARG(mz && n >= 0 && (n == 0 || (starts && arrivals && start_tiles && arrival_tiles && reward && out)),
               "nm_stream_build_pairs: NULL argument");
  int A;
  if (check_mouse(mouse, "nm_stream_build_pairs", &A) != NM_OK) return NM_ERR_INVALID;
  const NmMazeView mv = view_of(mz);
  const double *actions = mouse ? mouse->actions.data() : nullptr;
  for (int64_t i = 0; i < n; ++i) {
    const int sx = start_tiles[2 * i], sy = start_tiles[2 * i + 1];
    const int ax = arrival_tiles[2 * i], ay = arrival_tiles[2 * i + 1];
    NM_CHECK_ARG(nm_tile_visitable(mv, sx, sy) && nm_tile_visitable(mv, ax, ay),
                 "nm_stream_build_pairs: transition %lld starts or ends on a tile that is not visitable", (long long)i);
    if (!build_record(mz, mv, mouse, actions, A, starts[2 * i], starts[2 * i + 1], arrivals[2 * i], arrivals[2 * i + 1],
                      mz->state_of_tile[sx * mz->Y + sy], mz->state_of_tile[ax * mz->Y + ay], reward[i], (long long)i,
                      &out[i]))
      return NM_ERR_INVALID;
  }
  return n;
}
