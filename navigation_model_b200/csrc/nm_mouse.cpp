// Native Mouse: continuous pose + relative-action endpoints.
//
// B200-repo replacement of the reference's only native component, the pybind11/Eigen class in
// src/_navigation_model/{mouse.hpp:12-45, mouse.cpp:5-92, bindings.cpp:11-62}.  Eigen-free: the 2x2
// rotation is written out, and Eigen's `isApprox` (default double precision 1e-12,
// |a-b|^2 <= p^2 * min(|a|^2,|b|^2)) is restated in nm_pos_is_approx.  The same pose arithmetic is
// reused by the stream builder (nm_stream.cpp) and, on the device, by the simulation kernel.
#include <math.h>

#include <new>
#include <vector>

#include "nm_common.h"
#include "nm_pose.h"

static void mouse_reset_history(nm_mouse *m) {
  m->hist_pos.clear();
  m->hist_ori.clear();
  if (m->save_history) {
    m->hist_pos.push_back(m->init_pos[0]);
    m->hist_pos.push_back(m->init_pos[1]);
    m->hist_ori.push_back(m->init_ori);
  }
}

extern "C" nm_mouse *nm_mouse_create(const double *init_pos, double init_ori, const double *actions,
                                     int n_actions, int n_cols, int save_history) {
  if (!init_pos || n_actions < 0 || (n_actions > 0 && !actions)) {
    nm_fail(NM_ERR_INVALID, "nm_mouse_create: NULL argument");
    return nullptr;
  }
  if (n_actions > 0 && n_cols != 2) {  // mouse.cpp:6-9
    nm_fail(NM_ERR_INVALID, "The list of possible actions must have shape (N, 2)");
    return nullptr;
  }
  nm_mouse *m = new (std::nothrow) nm_mouse();
  if (!m) {
    nm_fail(NM_ERR_NOMEM, "nm_mouse_create: out of memory");
    return nullptr;
  }
  m->init_pos[0] = m->pos[0] = init_pos[0];
  m->init_pos[1] = m->pos[1] = init_pos[1];
  m->init_ori = m->ori = init_ori;
  m->actions.assign(actions, actions + 2 * (size_t)n_actions);
  m->save_history = save_history != 0;
  mouse_reset_history(m);
  return m;
}

extern "C" void nm_mouse_free(nm_mouse *m) { delete m; }

extern "C" int nm_mouse_state(const nm_mouse *m, double *out6) {
  NM_CHECK_ARG(m && out6, "nm_mouse_state: NULL argument");
  out6[0] = m->pos[0]; out6[1] = m->pos[1]; out6[2] = m->ori;
  out6[3] = m->init_pos[0]; out6[4] = m->init_pos[1]; out6[5] = m->init_ori;
  return NM_OK;
}

extern "C" int nm_mouse_num_actions(const nm_mouse *m) {
  NM_CHECK_ARG(m, "nm_mouse_num_actions: NULL handle");
  return (int)(m->actions.size() / 2);
}

extern "C" int nm_mouse_endpoints(const nm_mouse *m, double *out) {
  NM_CHECK_ARG(m && (out || m->actions.empty()), "nm_mouse_endpoints: NULL argument");
  const double c = cos(m->ori), s = sin(m->ori);
  const size_t n = m->actions.size() / 2;
  for (size_t i = 0; i < n; ++i)
    nm_pose_endpoint(c, s, m->pos[0], m->pos[1], m->actions[2 * i], m->actions[2 * i + 1], &out[2 * i],
                     &out[2 * i + 1]);
  return NM_OK;
}

int nm_mouse_move_impl(nm_mouse *m, double x, double y) {
  int moved = 0;
  if (!nm_pos_is_approx(x, y, m->pos[0], m->pos[1])) {
    m->ori = atan2(y - m->pos[1], x - m->pos[0]);
    m->pos[0] = x;
    m->pos[1] = y;
    moved = 1;
  }
  if (m->save_history) {  // pushed even when not moving (mouse.cpp:82-85)
    m->hist_pos.push_back(m->pos[0]);
    m->hist_pos.push_back(m->pos[1]);
    m->hist_ori.push_back(m->ori);
  }
  return moved;
}

extern "C" int nm_mouse_move_to(nm_mouse *m, double x, double y) {
  NM_CHECK_ARG(m, "nm_mouse_move_to: NULL handle");
  return nm_mouse_move_impl(m, x, y);
}

extern "C" int nm_mouse_move_through(nm_mouse *m, const double *points, int64_t n) {
  NM_CHECK_ARG(m && n >= 0 && (n == 0 || points), "nm_mouse_move_through: bad argument");
  int moved = 0;
  for (int64_t i = 0; i < n; ++i) moved = nm_mouse_move_impl(m, points[2 * i], points[2 * i + 1]);
  return moved;
}

extern "C" int nm_mouse_enable_history(nm_mouse *m, int enable) {
  NM_CHECK_ARG(m, "nm_mouse_enable_history: NULL handle");
  m->save_history = enable != 0;
  mouse_reset_history(m);
  return NM_OK;
}

extern "C" int nm_mouse_reset_history(nm_mouse *m) {
  NM_CHECK_ARG(m, "nm_mouse_reset_history: NULL handle");
  mouse_reset_history(m);
  return NM_OK;
}

extern "C" int nm_mouse_reset(nm_mouse *m, const double *init_pos, const double *init_ori) {
  NM_CHECK_ARG(m, "nm_mouse_reset: NULL handle");
  if (init_pos) {
    m->init_pos[0] = init_pos[0];
    m->init_pos[1] = init_pos[1];
  }
  if (init_ori) m->init_ori = *init_ori;
  m->pos[0] = m->init_pos[0];
  m->pos[1] = m->init_pos[1];
  m->ori = m->init_ori;
  mouse_reset_history(m);
  return NM_OK;
}

extern "C" int64_t nm_mouse_history_len(const nm_mouse *m) {
  NM_CHECK_ARG(m, "nm_mouse_history_len: NULL handle");
  return (int64_t)m->hist_ori.size();
}

extern "C" int nm_mouse_history(const nm_mouse *m, double *pos_out, double *ori_out) {
  NM_CHECK_ARG(m, "nm_mouse_history: NULL handle");
  const size_t n = m->hist_ori.size();
  if (pos_out) for (size_t i = 0; i < 2 * n; ++i) pos_out[i] = m->hist_pos[i];
  if (ori_out) for (size_t i = 0; i < n; ++i) ori_out[i] = m->hist_ori[i];
  return NM_OK;
}

extern "C" int nm_mouse_adopt_trajectory(nm_mouse *m, const double *pos, const double *ori, int64_t n) {
  NM_CHECK_ARG(m && n >= 0 && (n == 0 || (pos && ori)), "nm_mouse_adopt_trajectory: bad argument");
  for (int64_t i = 0; i < n; ++i) {
    m->pos[0] = pos[2 * i];
    m->pos[1] = pos[2 * i + 1];
    m->ori = ori[i];
    if (m->save_history) {
      m->hist_pos.push_back(m->pos[0]);
      m->hist_pos.push_back(m->pos[1]);
      m->hist_ori.push_back(m->ori);
    }
  }
  return NM_OK;
}
