// nn.cuh — K2: exact nearest-neighbour search against the uniform grid, one thread per query.
//
// Replaces pcl::KdTreeFLANN::nearestKSearch(k = 1) as reached from
// CorrespondenceEstimation::determineCorrespondences / GICP::searchForNeighbors inside icp.align()
// (reference src/icpslam/icp_odometer.cpp:198, src/icpslam/octree_mapper.cpp:114; SURVEY.md
// App. A.2 step 3, A.3, A.6) and from getFitnessScore() (icp_odometer.cpp:201).
//
// Search order for one query (thr = min(best d2 so far, bound2) is the pruning threshold):
//   seed      the previous iteration's match, if any: one point load that makes thr tight at once;
//   phase 0   the query's own cell;
//   phase 1   the 26 neighbour cells are TESTED (no memory access) against thr with conservative
//             per-axis face distances; survivors get their [start, end) range loaded and pushed on a
//             small per-thread list in shared memory — all range loads are independent (MLP);
//   phase 2   ONE flattened loop pops ranges and scans candidates, so a warp runs for the longest
//             lane's total work instead of the sum over cells of the per-cell maximum (divergence);
//   rings     cells at Chebyshev distance >= 2 are visited only while the distance from the query to
//             the outside of the block already covered does not exceed thr (rare).
// Lower bounds are distances to cell faces minus `slack`, summed in the association order of the float
// distance itself, so they can never exceed the float d2 of a point inside the cell: the result is
// the exact float-arithmetic nearest neighbour, ties on d2 resolved to the smallest original index —
// bit-identical to the oracle's exhaustive scan.
#pragma once
#include "common.cuh"
#include "grid.cuh"

namespace b2 {

constexpr int kListCap = 8;   // ranges a thread can queue for phase 2; overflow is scanned at once (6 / 12: same speed)

struct NNResult {
  unsigned long long key;  // pack_key(d2, original target index); kInfKey = nothing found
  int pos;                 // position of that point in the sorted array (GridView::pts)
  bool resolved;           // false: ring budget exhausted before the search could be proven exact
};

// Per-CTA scratch of the range lists: [kListCap][threads] words, column = thread (conflict-free).
template <int THREADS>
struct NNScratch {
  int start[kListCap][THREADS];
  unsigned meta[kListCap][THREADS];  // (float bits of lb, truncated) & 0xFFFF0000 | count
};

// conservative distance from coordinate q to the slab of cells [k_lo, k_hi] on one axis
__device__ __forceinline__ float slab_gap(float q, float o, float cell, int k_lo, int k_hi, float slack) {
  float lo = o + (float)k_lo * cell;
  float hi = o + (float)(k_hi + 1) * cell;
  float g = fmaxf(lo - q, q - hi) - slack;
  return fmaxf(g, 0.0f);
}

__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, int s, int e, float qx, float qy,
                                           float qz, unsigned long long& best, int& bpos) {
#pragma unroll 2
  for (int j = s; j < e; ++j) {
    float4 p = __ldg(pts + j);
    float d = sqdist3(qx, qy, qz, p.x, p.y, p.z);
    unsigned long long k = pack_key(d, __float_as_int(p.w));
    if (k < best) {
      best = k;
      bpos = j;
    }
  }
}

// `seed_pos` >= 0: sorted position of a target point already known to be close (the previous
// iteration's match); `seed_key`: the packed (d2, index) of such a point when only its value is known.
// Either only tightens the pruning threshold; the result is still the exact NN.
template <int THREADS>
__device__ __forceinline__ NNResult grid_nn(const GridView& g, float qx, float qy, float qz, float bound2,
                                            int max_rings, int seed_pos, NNScratch<THREADS>& sc,
                                            unsigned long long seed_key = kInfKey) {
  NNResult r;
  r.key = kInfKey;
  r.pos = -1;
  r.resolved = true;
  const int cx = cell_coord(qx, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(qy, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(qz, g.oz, g.inv_cell, g.nz);
  const int* cs_own = g.cell_start + (cz * g.ny + cy) * g.nx + cx;
  // issue the own-cell range loads before the seed's dependent point load
  const int s_own = __ldg(cs_own), e_own = __ldg(cs_own + 1);
  if (seed_pos >= 0) {
    const float4 p = __ldg(g.pts + seed_pos);
    r.key = pack_key(sqdist3(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
    r.pos = seed_pos;
  }
  if (seed_key < r.key) {  // a candidate known by value (d2, index) only: tightens the threshold the same way
    r.key = seed_key;
    r.pos = -1;
  }
  // ---- phase 0: own cell
  scan_range(g.pts, s_own, e_own, qx, qy, qz, r.key, r.pos);
  float thr = fminf(bound2, key_d2(r.key));

  // ---- phase 1: test the 26 neighbours, queue the survivors' ranges
  // squared conservative gaps to the neighbour slabs on each axis (index 0: minus side, 1: plus side)
  float gx2[2], gy2[2], gz2[2];
  {
    const float fxl = g.ox + (float)cx * g.cell, fyl = g.oy + (float)cy * g.cell, fzl = g.oz + (float)cz * g.cell;
    float t;
    t = fmaxf(qx - fxl - g.slack, 0.0f);            gx2[0] = fmul(t, t);
    t = fmaxf(fxl + g.cell - qx - g.slack, 0.0f);   gx2[1] = fmul(t, t);
    t = fmaxf(qy - fyl - g.slack, 0.0f);            gy2[0] = fmul(t, t);
    t = fmaxf(fyl + g.cell - qy - g.slack, 0.0f);   gy2[1] = fmul(t, t);
    t = fmaxf(qz - fzl - g.slack, 0.0f);            gz2[0] = fmul(t, t);
    t = fmaxf(fzl + g.cell - qz - g.slack, 0.0f);   gz2[1] = fmul(t, t);
  }
  const int tid = threadIdx.x;
  int nlist = 0;
  // (rows are a real loop, not unrolled: the kernel is instruction-fetch sensitive)
#pragma unroll 1
  for (int row = 0; row < 9; ++row) {
    const int dz = row / 3 - 1, dy = row % 3 - 1;
    {
      const int z = cz + dz;
      if (z < 0 || z >= g.nz) continue;
    }
    const float lz2 = dz == 0 ? 0.0f : (dz > 0 ? gz2[1] : gz2[0]);
    if (lz2 > thr) continue;
    {
      const int y = cy + dy;
      if (y < 0 || y >= g.ny) continue;
      const float ly2 = dy == 0 ? 0.0f : (dy > 0 ? gy2[1] : gy2[0]);
      const float lyz = fadd(ly2, lz2);  // (0 + ly2) + lz2 <= (dx2 + dy2) + dz2
      if (lyz > thr) continue;
      const int* cs = cs_own + (dz * g.ny + dy) * g.nx;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        if (dx == 0 && dy == 0 && dz == 0) continue;
        const int x = cx + dx;
        if (x < 0 || x >= g.nx) continue;
        const float lx2 = dx == 0 ? 0.0f : gx2[dx > 0];
        const float lb = fadd(fadd(lx2, ly2), lz2);
        if (lb > thr) continue;
        const int s = __ldg(cs + dx), e = __ldg(cs + dx + 1);
        const int cnt = e - s;
        if (cnt <= 0) continue;
        if (nlist < kListCap && cnt <= 0xFFFF) {
          sc.start[nlist][tid] = s;
          sc.meta[nlist][tid] = (__float_as_uint(lb) & 0xFFFF0000u) | (unsigned)cnt;
          ++nlist;
        } else {  // list full (no seed yet / very sparse own cell): scan now
          scan_range(g.pts, s, e, qx, qy, qz, r.key, r.pos);
          thr = fminf(bound2, key_d2(r.key));
        }
      }
    }
  }

  // ---- phase 2: one flattened loop over the queued ranges
  {
    int li = 0, j = 0, e = 0;
    for (;;) {
      if (j >= e) {
        if (li >= nlist) break;
        const int s = sc.start[li][tid];
        const unsigned m = sc.meta[li][tid];
        ++li;
        if (__uint_as_float(m & 0xFFFF0000u) > thr) continue;  // thr has tightened since phase 1
        j = s;
        e = s + (int)(m & 0xFFFFu);
      }
      // two candidates per trip (both loads in flight before either is used); when only one is left
      // the second slot re-reads it, which cannot change the result
      const int j1 = (j + 1 < e) ? j + 1 : j;
      const float4 p = __ldg(g.pts + j);
      const float4 p1 = __ldg(g.pts + j1);
      const float d = sqdist3(qx, qy, qz, p.x, p.y, p.z);
      const float d1 = sqdist3(qx, qy, qz, p1.x, p1.y, p1.z);
      const unsigned long long k = pack_key(d, __float_as_int(p.w));
      const unsigned long long k1 = pack_key(d1, __float_as_int(p1.w));
      if (k < r.key) {
        r.key = k;
        r.pos = j;
      }
      if (k1 < r.key) {
        r.key = k1;
        r.pos = j1;
      }
      thr = fminf(bound2, key_d2(r.key));
      j += 2;
    }
  }

  // ---- further rings, only while something outside the covered block could still win
  for (int rho = 1;; ++rho) {
    // distance from q to the outside of the block [c - rho, c + rho]^3, over directions that still have cells
    float ex = INFINITY;
    if (cx - rho > 0) ex = fminf(ex, qx - (g.ox + (float)(cx - rho) * g.cell));
    if (cx + rho < g.nx - 1) ex = fminf(ex, (g.ox + (float)(cx + rho + 1) * g.cell) - qx);
    if (cy - rho > 0) ex = fminf(ex, qy - (g.oy + (float)(cy - rho) * g.cell));
    if (cy + rho < g.ny - 1) ex = fminf(ex, (g.oy + (float)(cy + rho + 1) * g.cell) - qy);
    if (cz - rho > 0) ex = fminf(ex, qz - (g.oz + (float)(cz - rho) * g.cell));
    if (cz + rho < g.nz - 1) ex = fminf(ex, (g.oz + (float)(cz + rho + 1) * g.cell) - qz);
    if (ex == INFINITY) break;  // the block covers the whole grid
    ex = fmaxf(ex - g.slack, 0.0f);
    if (fmul(ex, ex) > thr) break;
    const int R = rho + 1;
    if (R > max_rings) {
      r.resolved = false;
      break;
    }
    const int z0 = max(cz - R, 0), z1 = min(cz + R, g.nz - 1);
    const int y0 = max(cy - R, 0), y1 = min(cy + R, g.ny - 1);
    const int xa = max(cx - R, 0), xb = min(cx + R, g.nx - 1);
    for (int z = z0; z <= z1; ++z) {
      const float lz = slab_gap(qz, g.oz, g.cell, z, z, g.slack);
      const float lz2 = fmul(lz, lz);
      if (lz2 > thr) continue;
      const bool zface = (z == cz - R) || (z == cz + R);
      for (int y = y0; y <= y1; ++y) {
        const float ly = slab_gap(qy, g.oy, g.cell, y, y, g.slack);
        const float lyz = fadd(fmul(ly, ly), lz2);
        if (lyz > thr) continue;
        const int row = (z * g.ny + y) * g.nx;
        if (zface || y == cy - R || y == cy + R) {
          const int s = __ldg(g.cell_start + row + xa), e = __ldg(g.cell_start + row + xb + 1);
          scan_range(g.pts, s, e, qx, qy, qz, r.key, r.pos);
        } else {
          if (cx - R >= 0) {
            const int s = __ldg(g.cell_start + row + cx - R), e = __ldg(g.cell_start + row + cx - R + 1);
            scan_range(g.pts, s, e, qx, qy, qz, r.key, r.pos);
          }
          if (cx + R <= g.nx - 1) {
            const int s = __ldg(g.cell_start + row + cx + R), e = __ldg(g.cell_start + row + cx + R + 1);
            scan_range(g.pts, s, e, qx, qy, qz, r.key, r.pos);
          }
        }
        thr = fminf(bound2, key_d2(r.key));
      }
    }
  }
  return r;
}

// (The stand-alone search entry point, nn_search_box_kernel, lives in nncache.cuh; grid_nn above serves
// getFitnessScore's uncertified queries and GICP's correspondence kernel, which need seeds by position.)

// Fallback for the queries whose ring budget ran out (far outside the map): exhaustive scan.  CTA x owns one
// chunk of the sorted target array (it stays in L1 while the CTA walks the list of queries, blockIdx.y-strided),
// every warp reduces its share with shuffles and merges it into the query's packed (d2, index) key with a
// 64-bit atomicMin — so one far query is scanned by the whole grid at once instead of by a single warp
// (measured: 27 far queries of a 2.1 M-query launch took 1.76 ms with one warp per query).
__global__ void __launch_bounds__(256) nn_brute_fallback(GridView g, const float4* __restrict__ q,
                                                         const int* __restrict__ list,
                                                         const unsigned int* __restrict__ count,
                                                         unsigned long long* __restrict__ keys) {
  const int lane = threadIdx.x & 31;
  const unsigned int total = *count;
  const int chunk = (g.n + gridDim.x - 1) / gridDim.x;
  const int j0 = blockIdx.x * chunk, j1 = min(j0 + chunk, g.n);
  for (unsigned int w = blockIdx.y; w < total; w += gridDim.y) {
    const float4 p = __ldg(q + list[w]);
    unsigned long long best = 0xFFFFFFFFFFFFFFFFull;
    for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
      const float4 t = __ldg(g.pts + j);
      const unsigned long long k = pack_key(sqdist3(p.x, p.y, p.z, t.x, t.y, t.z), __float_as_int(t.w));
      best = k < best ? k : best;
    }
    best = warp_min_key(best);
    if (lane == 0 && best != 0xFFFFFFFFFFFFFFFFull) atomicMin(keys + w, best);
  }
}

__global__ void __launch_bounds__(256) nn_brute_unpack(const int* __restrict__ list, const unsigned int* __restrict__ count,
                                                       const unsigned long long* __restrict__ keys,
                                                       int* __restrict__ idx, float* __restrict__ d2) {
  const unsigned int total = *count;
  for (unsigned int w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x) {
    const unsigned long long k = keys[w];
    const bool none = k == 0xFFFFFFFFFFFFFFFFull;
    idx[list[w]] = none ? -1 : key_idx(k);
    d2[list[w]] = none ? INFINITY : key_d2(k);
  }
}

}  // namespace b2
