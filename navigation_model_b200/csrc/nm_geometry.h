// Maze geometry shared by the host stream builder and the device simulation kernel.
//
// Restates the arithmetic of the reference's simulation/maze.py operation for operation (no FMA
// contraction: the library is compiled with -fmad=false and -ffp-contract=off):
//   py_floordiv        CPython float_divmod (Objects/floatobject.c) -- `x // w`        maze.py:393-394
//   nm_cont2disc       int(x // st), int(y // st)                                      maze.py:380-395
//   nm_tile_visitable  bounds check + tile != VOID                                     maze.py:488-501
//   nm_move_possible   ray / AABB walk                                                 maze.py:563-611
//   nm_closest_visitable  nearest visitable tile centre (KD-tree query in the ref)     maze.py:469-484
#ifndef NM_GEOMETRY_H
#define NM_GEOMETRY_H

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define NM_HD __host__ __device__ __forceinline__
#else
#define NM_HD inline
#endif

struct NmMazeView {
  const uint8_t *tiles;  // [X*Y], x-major; 0 = VOID
  int X, Y;
  double st;   // size_tile
  double inv;  // fl(1 / size_tile): only the device fast path uses it
};

// Python `x // w` for floats, w > 0 or w < 0 (w != 0).
NM_HD double py_floordiv(double x, double w) {
  double mod = fmod(x, w);
  double div = (x - mod) / w;
  if (mod != 0.0) {
    if ((w < 0.0) != (mod < 0.0)) div -= 1.0;
  }
  double fl;
  if (div != 0.0) {
    fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
  } else {
    fl = copysign(0.0, x / w);
  }
  return fl;
}

NM_HD bool nm_tile_visitable(const NmMazeView &m, int x, int y) {
  if (x < 0 || x >= m.X || y < 0 || y >= m.Y) return false;
  return m.tiles[x * m.Y + y] != 0;
}

// int() of a float floor-quotient; clamp so that huge / non-finite values stay "outside the maze".
//
// Python's `x // w` equals the mathematical floor of the EXACT quotient (divmod is fmod-based).  On
// the device that integer is obtained without fmod: q0 = floor(fl(x / w)) can only be n or n + 1, and
// the residual x - q0 * w, evaluated exactly-rounded by one FMA, is negative iff q0 = n + 1.
NM_HD int nm_tile_index_inv(double x, double st, double inv) {
#ifdef __CUDA_ARCH__
  // q0 = floor(fl(x * fl(1/st))) is within one of the exact floor n; the FMA residual x - q0*st (exact up to
  // a rounding that preserves its order relative to 0 and st) tells which: < 0 -> n = q0 - 1, >= st -> q0 + 1.
  double q = floor(x * inv);
  const double res = fma(-q, st, x);
  if (res < 0.0) q -= 1.0;
  else if (res >= st) q += 1.0;
#else
  (void)inv;
  double q = py_floordiv(x, st);
#endif
  if (!(q > -1.0e9)) return -1000000000;
  if (!(q < 1.0e9)) return 1000000000;
  return (int)q;
}

NM_HD int nm_tile_index(double x, double st) { return nm_tile_index_inv(x, st, 1.0 / st); }

NM_HD bool nm_free_at(const NmMazeView &m, double x, double y) {
  return nm_tile_visitable(m, nm_tile_index_inv(x, m.st, m.inv), nm_tile_index_inv(y, m.st, m.inv));
}

// `start_ok`: visitability of the tile under (x1, y1), computed by the caller (the simulation checks eight
// endpoints from the same start point).
NM_HD bool nm_move_possible_from(const NmMazeView &m, bool start_ok, double x1, double y1, double x2, double y2) {
  const double st = m.st;
  if (!start_ok || !nm_free_at(m, x2, y2)) return false;
  if (x1 != x2) {
    if (x1 > x2) {
      double t = x1; x1 = x2; x2 = t;
      t = y1; y1 = y2; y2 = t;
    }
    const double slope = (y2 - y1) / (x2 - x1);
    const int lo = nm_tile_index_inv(x1, st, m.inv), hi = nm_tile_index_inv(x2, st, m.inv);
    for (int i = lo; i < hi; ++i) {
      const double gx = (double)(i + 1) * st;
      const double gy = slope * (gx - x1) + y1;
      if (!nm_free_at(m, gx - st, gy)) return false;
      if (!nm_free_at(m, gx, gy)) return false;
    }
  }
  if (y1 != y2) {
    if (y1 > y2) {
      double t = x1; x1 = x2; x2 = t;
      t = y1; y1 = y2; y2 = t;
    }
    const double slope = (x2 - x1) / (y2 - y1);
    const int lo = nm_tile_index_inv(y1, st, m.inv), hi = nm_tile_index_inv(y2, st, m.inv);
    for (int j = lo; j < hi; ++j) {
      const double gy = (double)(j + 1) * st;
      const double gx = slope * (gy - y1) + x1;
      if (!nm_free_at(m, gx, gy - st)) return false;
      if (!nm_free_at(m, gx, gy)) return false;
    }
  }
  return true;
}

NM_HD bool nm_move_possible(const NmMazeView &m, double x1, double y1, double x2, double y2) {
  return nm_move_possible_from(m, nm_free_at(m, x1, y1), x1, y1, x2, y2);
}

// centre of tile i along one axis: st*(i+1) - st/2   (maze.py:262-263)
NM_HD double nm_tile_centre(int i, double st) { return st * (double)(i + 1) - st / 2.0; }

// Tile containing (x, y) if visitable, else the visitable tile whose centre is nearest (squared
// Euclidean distance, first in x-major order on exact ties).  Returns false if the maze has no
// visitable tile.
NM_HD bool nm_closest_visitable(const NmMazeView &m, double x, double y, int *ox, int *oy) {
  int tx = nm_tile_index(x, m.st), ty = nm_tile_index(y, m.st);
  if (nm_tile_visitable(m, tx, ty)) {
    *ox = tx; *oy = ty;
    return true;
  }
  double best = INFINITY;
  int bx = -1, by = -1;
  for (int i = 0; i < m.X; ++i)
    for (int j = 0; j < m.Y; ++j) {
      if (m.tiles[i * m.Y + j] == 0) continue;
      const double dx = nm_tile_centre(i, m.st) - x, dy = nm_tile_centre(j, m.st) - y;
      const double d = dx * dx + dy * dy;
      if (d < best) { best = d; bx = i; by = j; }
    }
  if (bx < 0) return false;
  // the reference re-discretises the centre it found (maze.py:484)
  *ox = nm_tile_index(nm_tile_centre(bx, m.st), m.st);
  *oy = nm_tile_index(nm_tile_centre(by, m.st), m.st);
  return true;
}

#endif  // NM_GEOMETRY_H
