// nncache.cuh — the per-query certificate that lets the NEXT ICP iteration skip the neighbour search, and the
// cell box a search has to cover.
//
// Contract (it replaces pcl::KdTreeFLANN::nearestKSearch(k = 1) inside
// CorrespondenceEstimation::determineCorrespondences, reference src/icpslam/icp_odometer.cpp:198,
// src/icpslam/octree_mapper.cpp:114): the result is the float-arithmetic
// nearest neighbour, ties on d2 to the smallest original index.  What a search leaves behind for the query:
//
//   c0, c1   the nearest and second-nearest target point found (coordinates + original index, 16 B each,
//            stored per query so that the next iteration's test is one coalesced streaming read),
//   L        a lower bound on the Euclidean distance from the query to EVERY OTHER target point (kept in the
//            fourth component of the running point, so it costs no extra load or store).
//
// ICP moves each query a little per iteration.  If the query has moved by `step` since the bound was
// taken, every other point is still at least L - step away (triangle inequality), so whenever
//       min(d(q, c0), d(q, c1))  <  L - step
// the nearest neighbour is one of the two cached points and no search is needed (the idea of Greenspan &
// Godin's cached-neighbour ICP, made exact in float arithmetic by the rounding margins below).  The
// sweep (sweep.cuh) runs that test for every query and sends only the failures to the search (coop.cuh).
//
// The search covers the cells overlapped by the axis-aligned box q +- r, r = sqrt(thr) + margin, where thr is
// the squared distance of the best cached candidate (or a probe radius when there is none).  Cells are ordered
// x-fastest, so each (y, z) row of the box is ONE contiguous run of the sorted target array.
// L = min(third distance, distance to the faces of the scanned box) - rounding.
#pragma once
#include "common.cuh"
#include "grid.cuh"
#include "nn.cuh"

namespace b2 {

constexpr float kRelUp = 1.000002f;    // > 1 + 16 ulp: turns a float distance into an upper bound
constexpr float kRelDown = 0.999998f;  // < 1 - 16 ulp: turns a float distance into a lower bound

// sqrt.approx.f32 (one MUFU, max relative error 2^-23): every use below is widened by kRelUp / kRelDown,
// which is 16 ulp, so the IEEE-rounded (software-assisted) sqrt is not needed.
__device__ __forceinline__ float sqrt_fast(float x) {
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct CellBox {
  int xa, xb, ya, yb, za, zb;
};

// The cells overlapped by the box q +- (sqrt(min(thr, bound2)) + margin), clamped to `max_span` cells
// either side of the query's cell.  cell_coord is monotone and is the expression the build binned with,
// so a target point within that distance of q on an axis lies in [ca, cb] on that axis.
__device__ __forceinline__ CellBox cell_box(const GridView& g, float qx, float qy, float qz, float thr, float bound2,
                                            float margin, int max_span) {
  const int cx = cell_coord(qx, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(qy, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(qz, g.oz, g.inv_cell, g.nz);
  thr = fminf(thr, bound2);
  CellBox b;
  if (thr < INFINITY) {
    const float rs = __fadd_ru(__fadd_ru(__fmul_ru(sqrt_fast(thr), kRelUp), margin), g.slack);
    b.xa = max(cell_coord(__fsub_rd(qx, rs), g.ox, g.inv_cell, g.nx), cx - max_span);
    b.xb = min(cell_coord(__fadd_ru(qx, rs), g.ox, g.inv_cell, g.nx), cx + max_span);
    b.ya = max(cell_coord(__fsub_rd(qy, rs), g.oy, g.inv_cell, g.ny), cy - max_span);
    b.yb = min(cell_coord(__fadd_ru(qy, rs), g.oy, g.inv_cell, g.ny), cy + max_span);
    b.za = max(cell_coord(__fsub_rd(qz, rs), g.oz, g.inv_cell, g.nz), cz - max_span);
    b.zb = min(cell_coord(__fadd_ru(qz, rs), g.oz, g.inv_cell, g.nz), cz + max_span);
  } else {  // unbounded search that found nothing nearby: the ring budget decides
    b.xa = max(cx - max_span, 0); b.xb = min(cx + max_span, g.nx - 1);
    b.ya = max(cy - max_span, 0); b.yb = min(cy + max_span, g.ny - 1);
    b.za = max(cz - max_span, 0); b.zb = min(cz + max_span, g.nz - 1);
  }
  return b;
}

}  // namespace b2
