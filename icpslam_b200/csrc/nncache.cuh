// nncache.cuh — exact 1-NN with a per-query certificate that lets the NEXT ICP iteration skip the search.
//
// Same contract as nn.cuh (it replaces pcl::KdTreeFLANN::nearestKSearch(k = 1) inside
// CorrespondenceEstimation::determineCorrespondences, reference src/icpslam/icp_odometer.cpp:198,
// src/icpslam/octree_mapper.cpp:114): the result is the float-arithmetic
// nearest neighbour, ties on d2 to the smallest original index.  What is new is what the search leaves
// behind for the query:
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
// fused sweep (icp.cuh) runs that test for every query and sends only the failures here.
//
// The search itself: the cells overlapped by the axis-aligned box q +- r, r = sqrt(thr) + margin, where
// thr is the squared distance of the best cached candidate (or of the own-cell / first-ring probe when
// there is none).  Cells are ordered x-fastest, so each (y, z) row of the box is ONE contiguous run of
// the sorted target array: two cell_start loads per row, no per-cell tests.  Runs are queued in shared
// memory and scanned by one flattened loop that keeps the best key, the runner-up and the third
// distance.  L = min(third distance, distance to the faces of the scanned box) - rounding.
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

#ifndef B2_X_REJECT
#define B2_X_REJECT 0  /* measured: 114 vs 105 us per iteration, the extra divergent branch costs more than the flops it saves */
#endif
#ifndef B2_PREFETCH_RUNS
#define B2_PREFETCH_RUNS 0  /* measured: no gain (10 294 vs 10 571 scans/s), the runs are mostly L1 hits already */
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

#ifndef B2_CACHE_K
#define B2_CACHE_K 2 /* measured with 3: 17.0 % of the queries searched instead of 19.9 %, but 18.9k scans/s against 22.1k */
#endif
constexpr int kCacheK = B2_CACHE_K;  // cached neighbours per query (c0 .. c{K-1}); the bound is the (K+1)-th distance

struct Top3 {               // the K nearest candidates of a scan and the (K+1)-th distance ("Top3" from K = 2)
  unsigned long long k0;    // best (d2, original index)
  int p[kCacheK];           // sorted positions: p[0] best, p[1..] runners-up in order, -1 = none
  float b[kCacheK];         // b[j] = (j+2)-th smallest d2 seen (+inf = none); b[K-1] is the bound
};

__device__ __forceinline__ void top3_init(Top3& t) {
  t.k0 = kInfKey;
#pragma unroll
  for (int j = 0; j < kCacheK; ++j) {
    t.p[j] = -1;
    t.b[j] = INFINITY;
  }
}

// branch-free insertion of candidate (d, original index idx, sorted position j)
__device__ __forceinline__ void top3_insert(Top3& t, float d, int idx, int j) {
  const unsigned long long k = pack_key(d, idx);
  const bool nb = k < t.k0;               // new best
  float dl = nb ? key_d2(t.k0) : d;       // the loser of each comparison goes on to the next rank
  int pl = nb ? t.p[0] : j;
  t.k0 = nb ? k : t.k0;
  t.p[0] = nb ? j : t.p[0];
#pragma unroll
  for (int r = 1; r < kCacheK; ++r) {
    const bool ns = dl < t.b[r - 1];
    const float nd = ns ? t.b[r - 1] : dl;
    const int np = ns ? t.p[r] : pl;
    t.b[r - 1] = ns ? dl : t.b[r - 1];
    t.p[r] = ns ? pl : t.p[r];
    dl = nd;
    pl = np;
  }
  t.b[kCacheK - 1] = fminf(t.b[kCacheK - 1], dl);
}

// min d2 over one run of the sorted array (probe only: no candidate bookkeeping)
__device__ __forceinline__ float probe_run(const float4* __restrict__ pts, int s, int e, float qx, float qy, float qz,
                                           float best) {
#pragma unroll 2
  for (int j = s; j < e; ++j) {
    const float4 p = __ldg(pts + j);
    best = fminf(best, sqdist3(qx, qy, qz, p.x, p.y, p.z));
  }
  return best;
}

struct CellBox {
  int xa, xb, ya, yb, za, zb;
};

// First-touch probe (no cached candidate): min d2 over the own cell, then over the 3x3x3 block (nine
// x-runs), so that the real scan starts from a finite radius.
__device__ __forceinline__ float probe_seed(const GridView& g, float qx, float qy, float qz) {
  const int cx = cell_coord(qx, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(qy, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(qz, g.oz, g.inv_cell, g.nz);
  const int* cs = g.cell_start + (cz * g.ny + cy) * g.nx + cx;
  float thr = probe_run(g.pts, __ldg(cs), __ldg(cs + 1), qx, qy, qz, INFINITY);
  if (!(thr < INFINITY)) {
    const int xa = max(cx - 1, 0), xb = min(cx + 1, g.nx - 1);
    for (int z = max(cz - 1, 0); z <= min(cz + 1, g.nz - 1); ++z)
      for (int y = max(cy - 1, 0); y <= min(cy + 1, g.ny - 1); ++y) {
        const int* row = g.cell_start + (z * g.ny + y) * g.nx;
        thr = probe_run(g.pts, __ldg(row + xa), __ldg(row + xb + 1), qx, qy, qz, thr);
      }
  }
  return thr;
}

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

// Exact nearest / second nearest over the cells of `b`.  On return `lrest` is a lower bound on the
// distance from q to every target point that was NOT scanned (those outside the box).
template <int THREADS>
__device__ __forceinline__ void box_search(const GridView& g, float qx, float qy, float qz, const CellBox& b,
                                           NNScratch<THREADS>& sc, Top3& top, float& lrest) {
  top3_init(top);
  const int tid = threadIdx.x;
  const int xa = b.xa, xb = b.xb, ya = b.ya, yb = b.yb, za = b.za, zb = b.zb;
  // distance from q to the outside of the box (faces that have cells beyond them only)
  {
    float gmin = INFINITY;
    if (xa > 0) gmin = fminf(gmin, __fsub_rd(qx, __fadd_ru(g.ox, __fmul_ru((float)xa, g.cell))));
    if (xb < g.nx - 1) gmin = fminf(gmin, __fsub_rd(__fadd_rd(g.ox, __fmul_rd((float)(xb + 1), g.cell)), qx));
    if (ya > 0) gmin = fminf(gmin, __fsub_rd(qy, __fadd_ru(g.oy, __fmul_ru((float)ya, g.cell))));
    if (yb < g.ny - 1) gmin = fminf(gmin, __fsub_rd(__fadd_rd(g.oy, __fmul_rd((float)(yb + 1), g.cell)), qy));
    if (za > 0) gmin = fminf(gmin, __fsub_rd(qz, __fadd_ru(g.oz, __fmul_ru((float)za, g.cell))));
    if (zb < g.nz - 1) gmin = fminf(gmin, __fsub_rd(__fadd_rd(g.oz, __fmul_rd((float)(zb + 1), g.cell)), qz));
    lrest = fmaxf(__fsub_rd(gmin, g.slack), 0.0f);
  }
  // rows of the box -> runs -> one flattened scan per chunk of kListCap runs
  const int nrow = (yb - ya + 1) * (zb - za + 1);
  int y = ya, z = za;
  for (int k = 0; k < nrow;) {
    int nlist = 0;
#pragma unroll 2
    for (; k < nrow && nlist < kListCap; ++k) {
      const int* row = g.cell_start + (z * g.ny + y) * g.nx;
      const int s = __ldg(row + xa), e = __ldg(row + xb + 1);
      if (++y > yb) {
        y = ya;
        ++z;
      }
      if (e > s) {
        sc.start[nlist][tid] = s;
        sc.meta[nlist][tid] = (unsigned)e;
        ++nlist;
#if B2_PREFETCH_RUNS
        // request the run's lines now: the scan below would otherwise pay one L2 round trip per run, in turn
        prefetch_l1(g.pts + s);
        if (e - s > 8) prefetch_l1(g.pts + s + 8);
#endif
      }
    }
    int li = 0, j = 0, e = 0;
    for (;;) {
      if (j >= e) {
        if (li >= nlist) break;
        j = sc.start[li][tid];
        e = (int)sc.meta[li][tid];
        ++li;
      }
      // two candidates per trip, both loads in flight before either is used
      const bool two = j + 1 < e;
      const float4 p = __ldg(g.pts + j);
      const float4 p1 = __ldg(g.pts + (two ? j + 1 : j));
#if B2_X_REJECT
      // d2 = (dx2 + dy2) + dz2 >= dx2 in float arithmetic too: a candidate whose x term alone reaches the
      // third-best distance (and exceeds the best) cannot change the state, whatever its y and z
      const float dx = fsub(qx, p.x), dx1 = fsub(qx, p1.x);
      const float x2 = fmul(dx, dx), x21 = two ? fmul(dx1, dx1) : INFINITY;
      if (fminf(x2, x21) < top.b[kCacheK - 1] || fminf(x2, x21) <= key_d2(top.k0)) {
        const float dy = fsub(qy, p.y), dz = fsub(qz, p.z), dy1 = fsub(qy, p1.y), dz1 = fsub(qz, p1.z);
        const float d = fadd(fadd(x2, fmul(dy, dy)), fmul(dz, dz));
        const float d1 = two ? fadd(fadd(x21, fmul(dy1, dy1)), fmul(dz1, dz1)) : INFINITY;
        if (fminf(d, d1) < top.b[kCacheK - 1] || fminf(d, d1) <= key_d2(top.k0)) {
          top3_insert(top, d, __float_as_int(p.w), j);
          if (two) top3_insert(top, d1, __float_as_int(p1.w), j + 1);
        }
      }
#else
      const float d = sqdist3(qx, qy, qz, p.x, p.y, p.z);
      const float d1 = two ? sqdist3(qx, qy, qz, p1.x, p1.y, p1.z) : INFINITY;
      // only candidates that beat the third-best distance (or tie the best) can change the state
      if (fminf(d, d1) < top.b[kCacheK - 1] || fminf(d, d1) <= key_d2(top.k0)) {
        top3_insert(top, d, __float_as_int(p.w), j);
        if (two) top3_insert(top, d1, __float_as_int(p1.w), j + 1);
      }
#endif
      j += 2;
    }
  }
}

// lower bound on the distance to every target point other than the K cached ones
__device__ __forceinline__ float top3_bound(const Top3& top, float lrest) {
  const float bk = top.b[kCacheK - 1];
  const float l3 = bk < INFINITY ? __fmul_rd(sqrt_fast(bk), kRelDown) : INFINITY;
  return fminf(l3, lrest);
}

}  // namespace b2
