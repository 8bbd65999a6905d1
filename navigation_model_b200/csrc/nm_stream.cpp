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

extern "C" int64_t nm_stream_build(const nm_maze *mz, const double *trajectory, const double *reward, int64_t T,
                                   nm_mouse *mouse, nm_step_rec *out) {
  NM_CHECK_ARG(mz && trajectory && reward && T >= 1, "nm_stream_build: NULL argument or empty trajectory");
  NM_CHECK_ARG(T == 1 || out, "nm_stream_build: NULL output");
  const bool with_candidates = mouse != nullptr;
  const int A = with_candidates ? (int)(mouse->actions.size() / 2) : 0;
  if (with_candidates) {
    NM_CHECK_ARG(A >= 1 && A <= NM_MAX_CAND, "nm_stream_build: need a mouse with 1..%d relative actions (has %d)",
                 NM_MAX_CAND, A);
  }
  NmMazeView mv;
  mv.tiles = mz->tiles.data();
  mv.X = mz->X;
  mv.Y = mz->Y;
  mv.st = mz->size_tile;
  mv.inv = 1.0 / mz->size_tile;

  const double *actions = with_candidates ? mouse->actions.data() : nullptr;

  int sx, sy;
  if (!nm_closest_visitable(mv, trajectory[0], trajectory[1], &sx, &sy))
    return nm_fail(NM_ERR_RUNTIME, "maze has no visitable tile");
  int s = mz->state_of_tile[sx * mz->Y + sy];
  for (int64_t i = 1; i < T; ++i) {
    const double cx = trajectory[2 * (i - 1)], cy = trajectory[2 * (i - 1) + 1];
    const double nx = trajectory[2 * i], ny = trajectory[2 * i + 1];
    int tx = 0, ty = 0;
    nm_closest_visitable(mv, nx, ny, &tx, &ty);
    const int s1 = mz->state_of_tile[tx * mz->Y + ty];
    nm_step_rec rec;
    rec.r = reward[i];
    rec.s = (uint16_t)s;
    rec.s1 = (uint16_t)s1;
    rec.ncand = 0;
    rec.obs = NM_OBS_NONE;
    rec.reserved = 0;
    for (int k = 0; k < NM_MAX_CAND; ++k) rec.cand[k] = 0;
    if (with_candidates) {
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
      if (rec.ncand == 0)  // np.max([]) raises in the reference (rl.py:184)
        return nm_fail(NM_ERR_INVALID,
                       "zero-size array to reduction operation maximum which has no identity "
                       "(step %lld has no visitable candidate)", (long long)i);
      nm_mouse_move_impl(mouse, nx, ny);  // rl.py:189
    }
    out[i - 1] = rec;
    s = s1;
  }
  return T - 1;
}
