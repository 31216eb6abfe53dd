// sweep.cuh — K2+K3 / K4: one ICP iteration of pcl::IterativeClosestPoint for every scan of a
// batch = two launches:
//
//   icp_sweep_coop   over the batch's ENTRY array (all queries of all scans, ordered by target tile: sort.cuh)
//       per entry   q = transformation_ * q   (in-place float chain, exactly PCL's transformCloud(input_transformed))
//                   cached-neighbour certificate (nncache.cuh): most queries keep their neighbour without a search
//       per warp    the failures of its slab, compacted, are searched COOPERATIVELY: union cell box, rows staged by
//                   bulk-async copies (TMA) into shared memory, one broadcast scan per group (coop.cuh)
//   icp_reduce       per scan, in original point order: gate d2 <= max^2 (CorrespondenceEstimation keeps equality),
//                   17 fp64 sums over a FIXED tree (256-query units, lane-strided, xor butterfly, units in order),
//                   then the last CTA of the scan runs Umeyama/SVD, final = T * final and
//                   DefaultConvergenceCriteria (solve.cuh).  `done` turns both kernels into no-ops for the scan.
//
// Reference: the loop inside icp.align(), src/icpslam/icp_odometer.cpp:198 / src/icpslam/octree_mapper.cpp:114.
// The summation tree does not depend on the launch configuration or on what else is in the batch, so a scan's
// result is bit-identical whether it runs alone, in a batch, or in a streamed batch.
//
// HBM bytes per entry per iteration: 1 (scan id) + 48 read (running point with its bound in .w, two cached
// neighbours) + 16 written; + 48 written and the staged target rows when the entry is searched; the reduce reads
// 4 (pos) + 32 (gathered running point and neighbour, L2-resident: the sweep has just written them).
#pragma once
#include "common.cuh"
#include "coop.cuh"
#include "nncache.cuh"
#include "solve.cuh"

namespace b2 {

__device__ __forceinline__ void accumulate_pair(double* acc, const float4& q, const float4& m, float d2) {
  const double sx = q.x, sy = q.y, sz = q.z, dx = m.x, dy = m.y, dz = m.z;
  acc[0] += 1.0;
  acc[1] += sx; acc[2] += sy; acc[3] += sz;
  acc[4] += dx; acc[5] += dy; acc[6] += dz;
  acc[7] += dx * sx; acc[8] += dx * sy; acc[9] += dx * sz;
  acc[10] += dy * sx; acc[11] += dy * sy; acc[12] += dy * sz;
  acc[13] += dz * sx; acc[14] += dz * sy; acc[15] += dz * sz;
  acc[16] += (double)d2;
}

constexpr int kTStride = 13;  // floats per scan in the shared transform table (odd: lanes of different scans spread over the banks)

struct SweepSmem {  // carved out of dynamic shared memory by sweep_smem_bytes()
  float4* buf;                // [warps][kCoopCap]
  unsigned long long* bar;    // [warps]
  float* T;                   // [kMaxScans][kTStride]
  GridView* grid;             // [kMaxScans] (per-scan grids; unused by SHARED kernels)
  unsigned short* wl;         // [warps][32 * QPT]
  int* tcs;                   // [warps][kTileRows * kTileW] cell-table slices of the warp's tile
  unsigned short* toff;       // [warps][kTileRows] first slot of every tile row
  unsigned char* flag;        // [kMaxScans]: bit 0 = done, bit 1 = first sweep
};

__host__ __device__ constexpr size_t sweep_smem_bytes(int qpt, bool shared_grid) {
  return (size_t)(kSweepThreads / 32) * kCoopCap * 16 + (size_t)(kSweepThreads / 32) * 8 + (size_t)kMaxScans * kTStride * 4 +
         (shared_grid ? 0 : (size_t)kMaxScans * sizeof(GridView)) + (size_t)(kSweepThreads / 32) * 32 * qpt * 2 +
         (size_t)(kSweepThreads / 32) * kTileRows * (kTileW * 4 + 2) + kMaxScans + 64;
}

__device__ __forceinline__ SweepSmem sweep_carve(unsigned char* base, int qpt, bool shared_grid) {
  SweepSmem s;
  constexpr int kWarps = kSweepThreads / 32;
  s.buf = reinterpret_cast<float4*>(base);
  base += (size_t)kWarps * kCoopCap * 16;
  s.bar = reinterpret_cast<unsigned long long*>(base);
  base += kWarps * 8;
  s.T = reinterpret_cast<float*>(base);
  base += kMaxScans * kTStride * 4;
  s.grid = reinterpret_cast<GridView*>(base);
  base += shared_grid ? 0 : kMaxScans * sizeof(GridView);
  s.wl = reinterpret_cast<unsigned short*>(base);
  base += (size_t)kWarps * 32 * qpt * 2;
  s.tcs = reinterpret_cast<int*>(base);
  base += (size_t)kWarps * kTileRows * kTileW * 4;
  s.toff = reinterpret_cast<unsigned short*>(base);
  base += (size_t)kWarps * kTileRows * 2;
  s.flag = base;
  return s;
}

struct SweepTune {
  float probe_frac;  // radius of the first, unseeded pass as a fraction of the cell edge
  int join_d;        // cells of slack inside which a lane joins its group's pass
  int use_tiles;     // stage 1 of phase B (slab tiles) on / off
};

// QPT: queries per lane (a warp owns a slab of 32 * QPT consecutive entries).  W: lanes per cooperative group.
// SHARED: every scan of the batch registers against bv.grid (one target: localisation / scan-to-map batches);
// otherwise each scan has its own grid (consecutive-pair batches) and a pass serves one scan at a time.
template <int QPT, int W, bool SHARED>
__global__ void __launch_bounds__(kSweepThreads, kSweepMinCtas) icp_sweep_coop(BatchView bv, const ScanTask* __restrict__ tasks,
                                                                                IcpConfig cfg, SweepTune tune) {
  constexpr int kSlab = 32 * QPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SweepSmem sm = sweep_carve(smem_raw, QPT, SHARED);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < bv.nscan * 12; k += kSweepThreads) {
    const int sid = k / 12, c = k - sid * 12;
    sm.T[sid * kTStride + c] = tasks[sid].state->Tinc[c];
  }
  for (int k = threadIdx.x; k < bv.nscan; k += kSweepThreads) {
    const IcpState* st = tasks[k].state;
    sm.flag[k] = (unsigned char)((st->done ? 1 : 0) | (st->iter == 0 ? 2 : 0));
    if (!SHARED) sm.grid[k] = tasks[k].grid;
  }
  if (lane == 0) mbar_init(sm.bar + warp, 1);
  __syncthreads();
  const int base = (blockIdx.x * (kSweepThreads / 32) + warp) * kSlab;
  if (base >= bv.E) return;
  unsigned short* const wl = sm.wl + warp * kSlab;
  float4* const wbuf = sm.buf + warp * kCoopCap;
  unsigned long long* const bar = sm.bar + warp;
  unsigned int phase = 0;

  // ---- phase A: transform + certificate, failures -> the warp's work list; the union of the failures' cell boxes
  int wc = 0;
  int uxa = 0x7FFFFFFF, uxb = -1, uya = 0x7FFFFFFF, uyb = -1, uza = 0x7FFFFFFF, uzb = -1;  // warp-uniform
  int slab_sid = -1;   // scan of the first failure (per-scan grids: a tile serves one grid)
  bool mixed = false;  // failures of more than one scan
#pragma unroll 1
  for (int qi = 0; qi < QPT; ++qi) {
    const int e = base + qi * 32 + lane;
    bool need = false;
    int sid = 0;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    float seed = INFINITY;
    if (e < bv.E) {
      sid = __ldg(bv.ent_sid + e);
      const unsigned int fl = sm.flag[sid];
      if (!(fl & 1u)) {
        const bool first = (fl & 2u) != 0;
        float T[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) T[c] = sm.T[sid * kTStride + c];
        float4 p, c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
        if (first) {
          p = __ldg(bv.ent_src + e);
        } else {  // independent coalesced loads
          p = ld_stream(bv.cur + e);
          c0 = ld_stream(bv.c0 + e);
          c1 = ld_stream(bv.c1 + e);
        }
        q = xform_f(T, p.x, p.y, p.z);
        q.w = 0.0f;
        if (!(isfinite(q.x) && isfinite(q.y) && isfinite(q.z))) {
          atomicOr(&tasks[sid].state->pad, 1);  // non-finite source point or transform: reported by the reduce
          const float4 none = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
          st_stream(bv.c0 + e, none);
          st_stream(bv.c1 + e, none);
          q = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (first) {
          need = true;
        } else {
          // bound after this iteration's motion (upper-rounded step, lower-rounded difference)
          const float step = __fmul_ru(sqrt_fast(sqdist3(q.x, q.y, q.z, p.x, p.y, p.z)), kRelUp);
          const float L = __fsub_rd(p.w, step);
          const int i0 = __float_as_int(c0.w), i1 = __float_as_int(c1.w);
          const unsigned long long k0 = i0 >= 0 ? pack_key(sqdist3(q.x, q.y, q.z, c0.x, c0.y, c0.z), i0) : kInfKey;
          const unsigned long long k1 = i1 >= 0 ? pack_key(sqdist3(q.x, q.y, q.z, c1.x, c1.y, c1.z), i1) : kInfKey;
          if (k1 < k0) {  // keep c0 = the nearer of the cached points
            st_stream(bv.c0 + e, c1);
            st_stream(bv.c1 + e, c0);
          }
          const float d2 = key_d2(k1 < k0 ? k1 : k0);
          const float L2 = L > 0.0f ? __fmul_rd(__fmul_rd(L, L), kRelDown) : 0.0f;
          if (fminf(d2, cfg.bound2) < L2) {
            q.w = L;  // certificate holds: the NN is c0, or nothing lies within the gate
          } else {
            need = true;
            seed = d2;  // distance to the nearer cached point: a certain search radius (+inf: none cached)
          }
        }
        st_stream(bv.cur + e, q);
      }
    }
    const unsigned int bal = __ballot_sync(0xFFFFFFFFu, need);
    if (bal) {
      if (need) wl[wc + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)(qi * 32 + lane);
      wc += __popc(bal);
      // the box this failure will search (the same expression phase B evaluates), into the slab's union
      if (slab_sid < 0) slab_sid = __shfl_sync(0xFFFFFFFFu, sid, __ffs(bal) - 1);
      mixed = mixed || __any_sync(0xFFFFFFFFu, need && sid != slab_sid);
      const GridView& g = SHARED ? bv.grid : sm.grid[slab_sid];
      const float pr = tune.probe_frac * g.cell;
      const CellBox bx = cell_box(g, q.x, q.y, q.z, seed < INFINITY ? seed : fmul(pr, pr), cfg.bound2, cfg.margin_frac * g.cell,
                                  cfg.max_rings);
      uxa = min(uxa, __reduce_min_sync(0xFFFFFFFFu, need ? bx.xa : 0x7FFFFFFF));
      uxb = max(uxb, __reduce_max_sync(0xFFFFFFFFu, need ? bx.xb : -1));
      uya = min(uya, __reduce_min_sync(0xFFFFFFFFu, need ? bx.ya : 0x7FFFFFFF));
      uyb = max(uyb, __reduce_max_sync(0xFFFFFFFFu, need ? bx.yb : -1));
      uza = min(uza, __reduce_min_sync(0xFFFFFFFFu, need ? bx.za : 0x7FFFFFFF));
      uzb = max(uzb, __reduce_max_sync(0xFFFFFFFFu, need ? bx.zb : -1));
    }
  }
  __syncwarp();
  if (wc == 0) return;
  if (lane == 0) atomicAdd(&tasks[__ldg(bv.ent_sid + base)].state->unresolved, (unsigned int)wc);  // statistics (per batch)

  // ---- phase B, stage 1: the union box of the work list staged once (bulk-async copies of its rows), every lane
  // scanning its own box out of shared memory.  Queries whose probe radius proved too small, and slabs whose
  // union does not fit a tile, go to stage 2.
  int wc2 = wc;
  bool use_cache_always = false;
  if (tune.use_tiles && (SHARED || !mixed)) {
    const GridView& g = SHARED ? bv.grid : sm.grid[slab_sid];
    int* const tcs = sm.tcs + warp * (kTileRows * kTileW);
    unsigned short* const toff = sm.toff + warp * kTileRows;
    if (tile_stage(g, uxa, uxb, uya, uyb, uza, uzb, wbuf, tcs, toff, bar, phase)) {
      Tile tile;
      tile.buf = wbuf;
      tile.cs = tcs;
      tile.off = toff;
      tile.xa = uxa;
      tile.ya = uya;
      tile.za = uza;
      tile.ny = uyb - uya + 1;
      wc2 = 0;
      for (int e0 = 0; e0 < wc; e0 += 32) {
        const bool have = e0 + lane < wc;
        const unsigned short id = have ? wl[e0 + lane] : (unsigned short)0;
        const int e = base + (int)id;
        bool again = false;
        if (have) {
          const int sid = (int)__ldg(bv.ent_sid + e);
          const float4 q = bv.cur[e];
          float seed = INFINITY;
          if (!(sm.flag[sid] & 2u)) {  // radius from the cached pair (the list keeps ids only; both points are L2-hot)
            const float4 a = bv.c0[e], b = bv.c1[e];
            if (__float_as_int(a.w) >= 0) seed = sqdist3(q.x, q.y, q.z, a.x, a.y, a.z);
            if (__float_as_int(b.w) >= 0) seed = fminf(seed, sqdist3(q.x, q.y, q.z, b.x, b.y, b.z));
          }
          const float pr = tune.probe_frac * g.cell;
          const CellBox bx = cell_box(g, q.x, q.y, q.z, seed < INFINITY ? seed : fmul(pr, pr), cfg.bound2,
                                      cfg.margin_frac * g.cell, cfg.max_rings);
          CoopTop top;
          tile_search(g, tile, q.x, q.y, q.z, bx, top);
          float bound = coop_bound(top);
          if (!(seed < INFINITY)) {  // probe radius: exact only if the best candidate beats everything outside the box
            const float best = key_d2(top.k0);
            const bool exact = !(top.lrest < INFINITY) ||
                               (top.k0 != kInfKey && best < __fmul_rd(__fmul_rd(top.lrest, top.lrest), kRelDown));
            if (!exact) {  // what was found becomes the cached pair: stage 2 takes its radius from it
              again = true;
              bound = 0.0f;
            }
          }
          st_stream(bv.c0 + e, coop_c0(top));
          st_stream(bv.c1 + e, coop_c1(top));
          st_stream(bv.cur + e, make_float4(q.x, q.y, q.z, bound));
        }
        const unsigned int bal = __ballot_sync(0xFFFFFFFFu, again);  // (also orders this round's reads of wl before the writes)
        if (again) wl[wc2 + __popc(bal & ((1u << lane) - 1u))] = id;
        wc2 += __popc(bal);
      }
      use_cache_always = true;
      __syncwarp();
    }
  }

  // ---- phase B, stage 2: what is left, 32 entries at a time, searched cooperatively (coop.cuh)
  for (int e0 = 0; e0 < wc2; e0 += 32) {
    const bool have = e0 + lane < wc2;
    const int e = have ? base + (int)wl[e0 + lane] : base;
    const int sid = have ? (int)__ldg(bv.ent_sid + e) : -1;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    float seed = INFINITY;
    if (have) {
      q = bv.cur[e];
      if (use_cache_always || !(sm.flag[sid] & 2u)) {
        const float4 a = bv.c0[e], b = bv.c1[e];
        if (__float_as_int(a.w) >= 0) seed = sqdist3(q.x, q.y, q.z, a.x, a.y, a.z);
        if (__float_as_int(b.w) >= 0) seed = fminf(seed, sqdist3(q.x, q.y, q.z, b.x, b.y, b.z));
      }
    }
    bool seeded = seed < INFINITY;  // thr is a true upper bound of the NN distance (or the gate decides)
    bool want = have;
    CoopTop top;
    coop_init(top);
    // At most three rounds: unseeded queries first try a probe radius; those whose best candidate is not closer
    // than everything outside the scanned box go again with a radius that is certain.
#pragma unroll 1
    for (int round = 0; round < 3; ++round) {
      if (SHARED) {
        const GridView& g = bv.grid;
        const float pr = tune.probe_frac * g.cell;
        const float thr = seeded ? seed : fmul(pr, pr);
        const CellBox bx = cell_box(g, q.x, q.y, q.z, thr, cfg.bound2, cfg.margin_frac * g.cell, cfg.max_rings);
        coop_search<W>(g, want, q.x, q.y, q.z, bx, tune.join_d, wbuf, bar, phase, top);
      } else {  // one scan (= one grid) at a time
        unsigned int todo = __ballot_sync(0xFFFFFFFFu, want);
        while (todo) {
          const int gs = __shfl_sync(0xFFFFFFFFu, sid, __ffs(todo) - 1);
          const GridView& g = sm.grid[gs];
          const bool mine = want && sid == gs;
          const float pr = tune.probe_frac * g.cell;
          const float thr = seeded ? seed : fmul(pr, pr);
          const CellBox bx = cell_box(g, q.x, q.y, q.z, thr, cfg.bound2, cfg.margin_frac * g.cell, cfg.max_rings);
          coop_search<W>(g, mine, q.x, q.y, q.z, bx, tune.join_d, wbuf, bar, phase, top);
          todo &= ~__ballot_sync(0xFFFFFFFFu, mine);
        }
      }
      if (want && !seeded) {
        const float best = key_d2(top.k0);
        const bool exact = !(top.lrest < INFINITY) ||
                           (top.k0 != kInfKey && best < __fmul_rd(__fmul_rd(top.lrest, top.lrest), kRelDown));
        if (exact) want = false;
        else seed = top.k0 != kInfKey ? best : INFINITY;  // next round: a certain radius, the gate, or the whole grid
        seeded = true;
      } else {
        want = false;
      }
      if (!__any_sync(0xFFFFFFFFu, want)) break;
    }
    if (have) {
      st_stream(bv.c0 + e, coop_c0(top));
      st_stream(bv.c1 + e, coop_c1(top));
      st_stream(bv.cur + e, make_float4(q.x, q.y, q.z, coop_bound(top)));
    }
  }
}

// K3 + K4: sums of one iteration of one scan over a fixed tree, then the solve (last CTA of the scan).
constexpr int kReduceUnit = 256;  // queries per warp
__global__ void __launch_bounds__(256) icp_reduce(const ScanTask* __restrict__ tasks, IcpConfig cfg) {
  const ScanTask& t = tasks[blockIdx.y];
  IcpState* st = t.state;
  if (st->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nunit = (t.n + kReduceUnit - 1) / kReduceUnit;
  const int unit = blockIdx.x * 8 + warp;
  if (unit >= nunit) return;
  double acc[kNumSums];
#pragma unroll
  for (int c = 0; c < kNumSums; ++c) acc[c] = 0.0;
  // four queries at a time: their entry positions first, then the eight gathers, so that a lane has all of its
  // loads of a round in flight before the first sum (the gathers are L2 hits: the sweep has just written them)
#pragma unroll 1
  for (int k0 = 0; k0 < kReduceUnit / 32; k0 += 4) {
    int e[4];
    float4 q[4], m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = unit * kReduceUnit + (k0 + k) * 32 + lane;
      e[k] = i < t.n ? __ldg(t.pos + i) : -1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (e[k] >= 0) {
        q[k] = __ldcg(t.cur + e[k]);
        m[k] = __ldcg(t.c0 + e[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (e[k] >= 0 && __float_as_int(m[k].w) >= 0) {
        const float d2 = sqdist3(q[k].x, q[k].y, q[k].z, m[k].x, m[k].y, m[k].z);
        if (!((double)d2 > cfg.max2)) accumulate_pair(acc, q[k], m[k], d2);
      }
    }
  }
  double mine = 0.0;  // lane c < kNumSums ends up with sum c of the unit
#pragma unroll
  for (int c = 0; c < kNumSums; ++c) {
    const double s = warp_sum(acc[c]);
    if (lane == c) mine = s;
  }
  if (lane < kNumSums) t.partials[(size_t)unit * kNumSums + lane] = mine;
  __threadfence();
  __syncwarp();
  unsigned int ticket = 0;
  if (lane == 0) ticket = atomicAdd(&st->ticket, 1u);
  ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
  if (ticket != (unsigned int)(nunit - 1)) return;
  __threadfence();
  // last warp of the scan: lane l adds units l, l + 32, ... in order, then the butterfly
#pragma unroll
  for (int c = 0; c < kNumSums; ++c) acc[c] = 0.0;
  for (int b = lane; b < nunit; b += 32) {
    const double* p = t.partials + (size_t)b * kNumSums;
#pragma unroll
    for (int c = 0; c < kNumSums; ++c) acc[c] += __ldcg(p + c);
  }
  __shared__ double s_sums[8][kNumSums];
#pragma unroll
  for (int c = 0; c < kNumSums; ++c) {
    const double s = warp_sum(acc[c]);
    if (lane == 0) s_sums[warp][c] = s;
  }
  if (lane == 0) {
    st->ticket = 0;
    p2p_finish_iteration(&s_sums[warp][0], st, cfg);
  }
}

// After the loop: the correspondences of the last sweep (what PCL's correspondences_ holds when align()
// returns) from the per-entry state: c0 is the exact nearest neighbour of cur whenever one lies within
// the gate, so  idx = c0.idx if d2 <= max_dist^2 else -1.  Written in the scan's original point order.
__global__ void __launch_bounds__(256) icp_finalize_corr(BatchView bv, const ScanTask* __restrict__ tasks, IcpConfig cfg) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= bv.E) return;
  const ScanTask& t = tasks[__ldg(bv.ent_sid + e)];
  const int i = __ldg(bv.ent_orig + e);
  const float4 q = bv.cur[e];
  const float4 m = bv.c0[e];
  const int id = __float_as_int(m.w);
  const float d2 = id >= 0 ? sqdist3(q.x, q.y, q.z, m.x, m.y, m.z) : INFINITY;
  t.corr_idx[i] = (id >= 0 && !((double)d2 > cfg.max2)) ? id : -1;
  t.corr_d2[i] = d2;
}

// ---- stand-alone exact 1-NN (b2icp_nn_search, K9's map_nearest), cooperative: 32 consecutive queries per
// group.  Round 0 probes a radius of one cell; queries whose best candidate is not provably the nearest go
// again with the radius they found (or, with nothing found, with boxes of 3 and then 10 cells either side);
// what is still open after that (nothing within ~10 cells) goes to the exhaustive fallback of nn.cuh.
// Algorithmic bytes per launch: 16 n_q (queries) + 16 N_t' (each target point in a touched cell once) + 8 n_q.
template <int W>
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
      if (!finite) {
        idx[i] = -1;
        d2[i] = INFINITY;
      } else if (exact) {
        idx[i] = key_idx(top.k0);
        d2[i] = key_d2(top.k0);
      } else {
        unresolved_list[atomicAdd(unresolved_count, 1u)] = i;
      }
    }
  }
}

}  // namespace b2
