// Device-only fast path of Maze.is_movement_possible_cont (simulation/maze.py:563-611) for the simulation kernel.
//
// Same decisions as nm_geometry.h::nm_move_possible (which restates the reference operation for operation), with the
// parameter-independent parts of the ray / AABB walk tabulated once per CTA and the rest made branch-free, so that
// the 32 mice of a warp stay converged:
//
//   * the walk visits the vertical lines x = fl((dx+1)*st), dx in [tile(x1), tile(x2)), and tests the two tiles
//     cont2disc(x - st, y) and cont2disc(x, y) (maze.py:588-595).  Their x indices depend on dx only, so
//         crossx[dx][ty] = visitable(tile(fl(x - st)), ty) && visitable(tile(x), ty)
//     is a table of X*Y bytes built with the exact floor division of nm_geometry.h; `gl[dx]` holds x itself.
//     Horizontal lines: crossy[dy][tx] likewise (maze.py:602-610);
//   * what remains per crossing is y = m*(x - x1) + y1 (the reference's operations, no FMA) and ONE tile index;
//   * m = (y2-y1)/(x2-x1) does not depend on the swap that orders the end points (maze.py:583-586, 598-601):
//     IEEE subtraction and division are sign-symmetric, (-a)/(-b) == a/b bit for bit;
//   * the caller's per-tile flag bytes (zone / corridor marks of the fused metrics) ride along as a fourth table;
//   * both end points are inside the maze whenever the walk runs, so dx in [0, X-2] and the tables never overflow;
//     lanes whose move is already refused run the same instructions on index 0 and discard the result.
//
// tile index: Python's `x // st` is the floor of the exact quotient.  n = rint(x * fl(1/st)) (one FMA against the
// 1.5*2^52 constant, low word = n) is floor or floor+1 for |x/st| < 2^31; the residual x - n*st, rounded once by an
// FMA, keeps its sign, and is negative exactly when n = floor+1.  No FRND / F2I (quarter-rate XU ops) needed.
#ifndef NM_SIMGEOM_CUH
#define NM_SIMGEOM_CUH

#include "nm_geometry.h"

// Shared-memory reads by 32-bit shared-window address.  The tables below are addressed through four such addresses,
// PINNED in registers by an opaque move: under the simulation kernel's 80-register cap the compiler otherwise
// rematerialises every table pointer at every use (S2R / S2UR / UMOV / ULEA / IMAD chains from the launch parameters --
// 14 % of the kernel's executed instructions in round 1's profile).
// Loads of tables that no longer change: plain asm (the compiler may schedule and merge them freely).  Their addresses
// derive from a pin executed AFTER the table was written, so no load can move above its table's initialisation.
__device__ __forceinline__ uint32_t nm_lds_u8(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double nm_lds_f64(uint32_t addr) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void nm_sts_f64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ uint32_t nm_pin_u32(uint32_t v) {
  asm volatile("mov.u32 %0, %0;" : "+r"(v));
  return v;
}

struct NmSimGeom {
  uint32_t a_tiles;   // [X*Y] shared-window address of the tile table
  uint32_t a_crossx;  // [X][Y]
  uint32_t a_crossy;  // [Y][X]
  uint32_t a_flags;   // [X*Y] the caller's per-tile flag bytes (zone / corridor marks of the fused metrics), 0 without
  uint32_t a_gl;      // [max(X,Y)] doubles  fl((i+1)*st)
  int X, Y;
  double st, inv;

  __device__ __forceinline__ uint32_t tile_code(int idx) const { return nm_lds_u8(a_tiles + (uint32_t)idx); }
  __device__ __forceinline__ uint32_t tile_flags(int idx) const { return nm_lds_u8(a_flags + (uint32_t)idx); }

  // the flag bytes are staged for mazes of up to 4096 tiles (where shared memory is plentiful); larger mazes read them
  // from device memory, so that the tile / grid-line tables alone decide how large a maze fits (~66 000 tiles)
  static __host__ __device__ bool flags_staged(int X, int Y) { return (size_t)X * Y <= 4096; }
  static __host__ __device__ size_t smem_bytes(int X, int Y) {
    const int m = X > Y ? X : Y;
    return (size_t)m * sizeof(double) + (flags_staged(X, Y) ? 4 : 3) * (((size_t)X * Y + 15) & ~(size_t)15);
  }

  // `base` 8-byte aligned, smem_bytes(X, Y) long; every thread of the CTA calls build(), then __syncthreads().
  __device__ __forceinline__ void attach(unsigned char *base, int X_, int Y_, double st_) {
    X = X_; Y = Y_; st = st_; inv = 1.0 / st_;
    const int m = X > Y ? X : Y;
    const uint32_t slab = ((uint32_t)(X * Y) + 15u) & ~15u;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(base);
    a_gl = nm_pin_u32(b);
    a_tiles = nm_pin_u32(b + (uint32_t)m * (uint32_t)sizeof(double));
    a_crossx = nm_pin_u32(a_tiles + slab);
    a_crossy = nm_pin_u32(a_crossx + slab);
    a_flags = nm_pin_u32(a_crossy + slab);
  }

  static __device__ __noinline__ void fill(unsigned char *base, const uint8_t *g_tiles, const uint8_t *g_flags, int X, int Y,
                                           double st) {
    const double inv = 1.0 / st;
    const int nT = X * Y, m = X > Y ? X : Y;
    const size_t slab = ((size_t)nT + 15) & ~(size_t)15;
    double *gl_w = reinterpret_cast<double *>(base);
    uint8_t *t_w = base + (size_t)m * sizeof(double), *cx_w = t_w + slab, *cy_w = cx_w + slab, *f_w = cy_w + slab;
    NmMazeView mv;
    mv.tiles = g_tiles; mv.X = X; mv.Y = Y; mv.st = st; mv.inv = inv;
#pragma unroll 1
    for (int i = threadIdx.x; i < m; i += blockDim.x) gl_w[i] = (double)(i + 1) * st;
#pragma unroll 1
    for (int i = threadIdx.x; i < nT; i += blockDim.x) {
      t_w[i] = g_tiles[i];
      if (flags_staged(X, Y)) f_w[i] = g_flags ? g_flags[i] : 0;
      {  // crossx[dx][ty]
        const int dx = i / Y, ty = i - dx * Y;
        const double gx = (double)(dx + 1) * st;
        const int ta = nm_tile_index_inv(gx - st, st, inv), tb = nm_tile_index_inv(gx, st, inv);
        cx_w[i] = (nm_tile_visitable(mv, ta, ty) && nm_tile_visitable(mv, tb, ty)) ? 1 : 0;
      }
      {  // crossy[dy][tx]
        const int dy = i / X, tx = i - dy * X;
        const double gy = (double)(dy + 1) * st;
        const int ta = nm_tile_index_inv(gy - st, st, inv), tb = nm_tile_index_inv(gy, st, inv);
        cy_w[i] = (nm_tile_visitable(mv, tx, ta) && nm_tile_visitable(mv, tx, tb)) ? 1 : 0;
      }
    }
  }

  // every thread of the CTA: fill the tables, synchronise, THEN take the (pinned) addresses the reads go through
  __device__ __forceinline__ void build(unsigned char *base, const uint8_t *g_tiles, const uint8_t *g_flags, int X_, int Y_,
                                        double st_) {
    fill(base, g_tiles, g_flags, X_, Y_, st_);
    __syncthreads();
    attach(base, X_, Y_, st_);
  }

  __device__ __forceinline__ int tile(double x) const {
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
    const double t = fma(x, inv, kMagic);
    const int n = __double2loint(t);
    const double res = fma(-__dsub_rn(t, kMagic), st, x);
    return res < 0.0 ? n - 1 : n;
  }

  // one pass of the walk: lines crossed between the start tile index s_t and the end tile index e_t along the axis
  // whose coordinates are (pa, ea); (pb, eb) are the coordinates along the other axis.
  __device__ __forceinline__ bool pass(bool ok, double pa, double pb, double ea, double eb, int s_t, int e_t,
                                       const uint32_t cross, int n_other) const {
    const bool sw = pa > ea;                       // order the end points along this axis
    const double a1 = sw ? ea : pa, b1 = sw ? eb : pb;
    const int lo = sw ? e_t : s_t, hi = sw ? s_t : e_t;
    const int n = ok ? hi - lo : 0;                // equal coordinates give equal tiles: the pass is skipped
    const double den = __dsub_rn(ea, pa);
    const double m = __ddiv_rn(__dsub_rn(eb, pb), den != 0.0 ? den : 1.0);
    bool good = true;
#pragma unroll
    for (int c = 0; c < 2; ++c) {                  // the first two crossings, predicated
      const bool on = n > c;
      const int d = on ? lo + c : 0;
      const double g = nm_lds_f64(a_gl + 8u * (uint32_t)d);
      const double o = __dadd_rn(__dmul_rn(m, __dsub_rn(g, a1)), b1);
      const int t = tile(o);
      const bool in = (unsigned)t < (unsigned)n_other;
      const uint32_t f = nm_lds_u8(cross + (uint32_t)(d * n_other + (in ? t : 0)));
      good = good && (!on || (in && f != 0));
    }
#pragma unroll 1
    for (int c = 2; c < n; ++c) {                  // longer moves (actions of more than two tiles)
      const int d = lo + c;
      const double o = __dadd_rn(__dmul_rn(m, __dsub_rn(nm_lds_f64(a_gl + 8u * (uint32_t)d), a1)), b1);
      const int t = tile(o);
      const bool in = (unsigned)t < (unsigned)n_other;
      good = good && in && nm_lds_u8(cross + (uint32_t)(d * n_other + (in ? t : 0))) != 0;
    }
    return good;
  }

  // (px, py) in tile (stx, sty) whose visitability is `start_ok`; end point (ex, ey).  *etile = flat index of the end
  // tile (0 when outside).
  __device__ __forceinline__ bool move_ok(bool start_ok, int stx, int sty, double px, double py, double ex, double ey,
                                          int *etile) const {
    const int etx = tile(ex), ety = tile(ey);
    const bool e_in = (unsigned)etx < (unsigned)X && (unsigned)ety < (unsigned)Y;
    const int et = e_in ? etx * Y + ety : 0;
    *etile = et;
    const bool ok = start_ok && e_in && tile_code(et) != 0;  // maze.py:578-579
    const bool gx = pass(ok, px, py, ex, ey, stx, etx, a_crossx, Y);  // vertical lines   :581-595
    const bool gy = pass(ok, py, px, ey, ex, sty, ety, a_crossy, X);  // horizontal lines :597-610
    return ok && gx && gy;
  }
};

#endif  // NM_SIMGEOM_CUH
