// Pose arithmetic of the Mouse, shared host/device (restates src/_navigation_model/mouse.cpp:64-87).
#ifndef NM_POSE_H
#define NM_POSE_H

#include "nm_geometry.h"

// endpoint = R(ori) * a + pos with R = [c -s; s c]; products and sums in Eigen's evaluation order
// (row of the 2x2 times the action, then + pos), no FMA.   mouse.cpp:65-71
NM_HD void nm_pose_endpoint(double c, double s, double px, double py, double ax, double ay, double *ex,
                            double *ey) {
  *ex = (c * ax + (-s) * ay) + px;
  *ey = (s * ax + c * ay) + py;
}

// Eigen::Vector2d::isApprox(other) with the default precision 1e-12.   mouse.cpp:77
NM_HD bool nm_pos_is_approx(double ax, double ay, double bx, double by) {
  const double dx = ax - bx, dy = ay - by;
  const double na = ax * ax + ay * ay, nb = bx * bx + by * by;
  const double mn = na < nb ? na : nb;
  return (dx * dx + dy * dy) <= (1e-12 * 1e-12) * mn;
}

#endif  // NM_POSE_H
