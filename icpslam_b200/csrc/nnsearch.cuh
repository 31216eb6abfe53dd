// nnsearch.cuh — the stand-alone exact 1-NN search (b2icp_nn_search, K9's map_nearest): cooperative groups over
// TMA-staged candidate rows (coop.cuh).  Replaces pcl::KdTreeFLANN::nearestKSearch(k = 1) as the reference reaches
// it through icp.align() (src/icpslam/icp_odometer.cpp:198) and pcl::octree approxNearestSearch
// (src/icpslam/octree_mapper.cpp:84).
#pragma once
#include "common.cuh"
#include "coop.cuh"

namespace b2 {

// Cooperative: 32 consecutive queries per
// group.  Round 0 probes a radius of one cell; queries whose best candidate is not provably the nearest go
// again with the radius they found (or, with nothing found, with boxes of 3 and then 10 cells either side);
// what is still open after that (nothing within ~10 cells) goes to the exhaustive fallback of nn.cuh.
// Algorithmic bytes per launch: 16 n_q (queries) + 16 N_t' (each target point in a touched cell once) + 8 n_q.
// SORTED: `q` is the query cloud counting-sorted by cell of the TARGET grid (grid.cuh: grid_count / scan / grid_scatter
// with the target's GridView), q[i].w = original query index.  32 consecutive sorted queries then share one cell or a
// few adjacent cells of a row, and the union box of a group is about one query's box: ~30 staged candidates per
// group instead of the several hundred of 32 consecutive rays of a sweep.  Results go to the original positions.
template <int W, bool SORTED>
__global__ void __launch_bounds__(kSweepThreads, kSweepMinCtas) nn_search_coop(GridView g, const float4* __restrict__ q, int n,
                                                                                int max_span, int join_d, int* __restrict__ idx,
                                                                                float* __restrict__ d2,
                                                                                int* __restrict__ unresolved_list,
                                                                                unsigned int* __restrict__ unresolved_count) {
  __shared__ __align__(16) float4 s_buf[kSweepThreads / 32][kCoopCap];
  __shared__ __align__(8) unsigned long long s_bar[kSweepThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) mbar_init(s_bar + warp, 1);
  __syncwarp();
  unsigned int phase = 0;
  const int ngroup = (n + 31) / 32;
  for (int grp = blockIdx.x * (kSweepThreads / 32) + warp; grp < ngroup; grp += gridDim.x * (kSweepThreads / 32)) {
    const int i = grp * 32 + lane;
    const bool have = i < n;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (have) p = __ldg(q + i);
    const bool finite = isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
    bool want = have && finite;
    float thr = fmul(g.cell, g.cell);  // round 0: probe radius of one cell
    int span = 0x3FFFFFFF;
    bool exact = false, certain = false;
    CoopTop top;
    coop_init(top);
#pragma unroll 1
    for (int round = 0; round < 4; ++round) {
      const CellBox bx = cell_box(g, p.x, p.y, p.z, thr, INFINITY, 0.01f * g.cell, span);
      coop_search<W>(g, want, p.x, p.y, p.z, bx, join_d, s_buf[warp], s_bar + warp, phase, top);
      if (want) {
        const float best = key_d2(top.k0);
        const bool found = top.k0 != kInfKey;
        exact = found && (!(top.lrest < INFINITY) || best < __fmul_rd(__fmul_rd(top.lrest, top.lrest), kRelDown));
        if (exact || certain) {
          want = false;  // (certain and not exact: the box was clamped to 10 cells — exhaustive fallback)
        } else if (found) {  // the distance found is a certain radius
          thr = best;
          span = 3 * max_span + 1;
          certain = true;
        } else if (span > 3 * max_span + 1) {  // nothing within the probe radius: 3 cells either side
          thr = INFINITY;
          span = max_span;
        } else if (span == max_span) {
          span = 3 * max_span + 1;
        } else {
          want = false;  // nothing within 10 cells: exhaustive fallback
        }
      }
      if (!__any_sync(0xFFFFFFFFu, want)) break;
    }
    if (have) {
      const int o = SORTED ? __float_as_int(p.w) : i;  // where the answer goes
      if (!finite) {
        idx[o] = -1;
        d2[o] = INFINITY;
      } else if (exact) {
        idx[o] = key_idx(top.k0);
        d2[o] = key_d2(top.k0);
      } else {
        unresolved_list[atomicAdd(unresolved_count, 1u)] = o;
      }
    }
  }
}

}  // namespace b2
