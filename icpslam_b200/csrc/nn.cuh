// nn.cuh — K2: exact nearest-neighbour search against the uniform grid, one thread per query.
//
// Replaces pcl::KdTreeFLANN::nearestKSearch(k = 1) as reached from
// CorrespondenceEstimation::determineCorrespondences / GICP::searchForNeighbors inside icp.align()
// (reference src/icpslam/icp_odometer.cpp:198, src/icpslam/octree_mapper.cpp:114; SURVEY.md
// App. A.2 step 3, A.3, A.6) and from getFitnessScore() (icp_odometer.cpp:201).
//
// Search = ring expansion around the query's cell.  The 3x3x3 block is visited first as nine
// x-runs of three cells (one contiguous range of the sorted point array each, centre row first);
// a row is skipped when its conservative lower bound exceeds the running threshold
// thr = min(best d2, bound2).  Further rings are visited only while the distance from the query to
// the outside of the block already covered does not exceed thr.  Lower bounds are distances to cell
// faces minus `slack`, so they can never exceed the float d2 of a point inside the cell: the result
// is the exact float-arithmetic nearest neighbour with ties on d2 resolved to the smallest original
// index — bit-identical to the oracle's exhaustive scan.
#pragma once
#include "common.cuh"
#include "grid.cuh"

namespace b2 {

struct NNResult {
  unsigned long long key;  // pack_key(d2, original target index); kInfKey = nothing found
  int pos;                 // position of that point in the sorted array (GridView::pts)
  bool resolved;           // false: ring budget exhausted before the search could be proven exact
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
#pragma unroll 4
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
// iteration's match).  It only tightens the pruning threshold; the result is still the exact NN.
__device__ __forceinline__ NNResult grid_nn(const GridView& g, float qx, float qy, float qz, float bound2,
                                            int max_rings, int seed_pos = -1) {
  NNResult r;
  r.key = kInfKey;
  r.pos = -1;
  r.resolved = true;
  const int cx = cell_coord(qx, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(qy, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(qz, g.oz, g.inv_cell, g.nz);
  float thr = bound2;
  if (seed_pos >= 0) {
    const float4 p = __ldg(g.pts + seed_pos);
    r.key = pack_key(sqdist3(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
    r.pos = seed_pos;
    thr = fminf(bound2, key_d2(r.key));
  }

  // ---- 3x3x3 block: centre row, then the 4 edge-adjacent rows, then the 4 corner rows; inside a
  // row the three cells are pruned one by one against thr
  {
    // conservative x-gaps to the left / right neighbour cells (0 for the own cell)
    const float gxl = fmaxf(qx - (g.ox + (float)cx * g.cell) - g.slack, 0.0f);
    const float gxr = fmaxf((g.ox + (float)(cx + 1) * g.cell) - qx - g.slack, 0.0f);
    const float gxl2 = fmul(gxl, gxl), gxr2 = fmul(gxr, gxr);
    const bool has_l = cx > 0, has_r = cx < g.nx - 1;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int dy = (t == 1 || t == 5 || t == 7) ? -1 : ((t == 2 || t == 6 || t == 8) ? 1 : 0);
      const int dz = (t == 3 || t == 5 || t == 6) ? -1 : ((t == 4 || t == 7 || t == 8) ? 1 : 0);
      const int y = cy + dy, z = cz + dz;
      if (y < 0 || y >= g.ny || z < 0 || z >= g.nz) continue;
      const float ly = slab_gap(qy, g.oy, g.cell, y, y, g.slack);
      const float lz = slab_gap(qz, g.oz, g.cell, z, z, g.slack);
      const float ly2 = fmul(ly, ly), lz2 = fmul(lz, lz);
      const float lb = fadd(ly2, lz2);
      if (lb > thr) continue;
      const int* cs = g.cell_start + (z * g.ny + y) * g.nx + cx;
      const int s0 = __ldg(cs), s1 = __ldg(cs + 1);
      scan_range(g.pts, s0, s1, qx, qy, qz, r.key, r.pos);
      thr = fminf(bound2, key_d2(r.key));
      if (has_l && !(fadd(fadd(gxl2, ly2), lz2) > thr)) {
        scan_range(g.pts, __ldg(cs - 1), s0, qx, qy, qz, r.key, r.pos);
        thr = fminf(bound2, key_d2(r.key));
      }
      if (has_r && !(fadd(fadd(gxr2, ly2), lz2) > thr)) {
        scan_range(g.pts, s1, __ldg(cs + 2), qx, qy, qz, r.key, r.pos);
        thr = fminf(bound2, key_d2(r.key));
      }
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

// ---- stand-alone NN sweep (b2icp_nn_search): the roofline kernel ------------------------------
// Algorithmic bytes per launch: 16 n_q (queries) + 16 N_t' (each target point in a touched cell once)
// + 8 n_q (idx + d2).
__global__ void __launch_bounds__(kSweepThreads) nn_search_kernel(GridView g, const float4* __restrict__ q, int n,
                                                                  float bound2, int max_rings,
                                                                  int* __restrict__ idx, float* __restrict__ d2,
                                                                  int* __restrict__ unresolved_list,
                                                                  unsigned int* __restrict__ unresolved_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(q + i);
  NNResult r = grid_nn(g, p.x, p.y, p.z, bound2, max_rings);
  if (!r.resolved) {
    unsigned int slot = atomicAdd(unresolved_count, 1u);
    unresolved_list[slot] = i;
    return;
  }
  int id = key_idx(r.key);
  idx[i] = (r.key == kInfKey) ? -1 : id;
  d2[i] = key_d2(r.key);
}

// Fallback for the queries whose ring budget ran out (far outside the map): one warp per query,
// exhaustive coalesced scan of the sorted target array, warp-shuffle min of the packed keys.
__global__ void __launch_bounds__(256) nn_brute_fallback(GridView g, const float4* __restrict__ q,
                                                         const int* __restrict__ list,
                                                         const unsigned int* __restrict__ count,
                                                         int* __restrict__ idx, float* __restrict__ d2) {
  const int lane = threadIdx.x & 31;
  const unsigned int nw = (gridDim.x * blockDim.x) >> 5;
  const unsigned int total = *count;
  for (unsigned int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += nw) {
    const int qi = list[w];
    const float4 p = __ldg(q + qi);
    unsigned long long best = kInfKey;
    for (int j = lane; j < g.n; j += 32) {
      float4 t = __ldg(g.pts + j);
      unsigned long long k = pack_key(sqdist3(p.x, p.y, p.z, t.x, t.y, t.z), __float_as_int(t.w));
      best = k < best ? k : best;
    }
    best = warp_min_key(best);
    if (lane == 0) {
      idx[qi] = (best == kInfKey) ? -1 : key_idx(best);
      d2[qi] = key_d2(best);
    }
  }
}

}  // namespace b2
