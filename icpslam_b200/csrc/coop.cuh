// coop.cuh — warp-cooperative exact nearest-neighbour search over TMA-staged candidate tiles.
//
// The stand-alone search (pcl::KdTreeFLANN::nearestKSearch(k = 1) as reached from icp.align(), reference
// src/icpslam/icp_odometer.cpp:198, src/icpslam/octree_mapper.cpp:114; also K9's
// approxNearestNeighbors in exact mode): nnsearch.cuh drives it.  Inside the ICP loop the per-lane box scan of
// nncache.cuh stays: there the cached-neighbour certificate has already removed 80 % of the searches and what is left
// is too sparse to share candidates (measured: DESIGN.md section 4, profiles/r02_coop_sweep_*).
//
// The queries a warp has to search are close together (large query clouds are counting-sorted by target cell
// first, b2icp.cu: nn_search_impl).  Instead of every lane walking its own cells through dependent global loads, a GROUP of W lanes
// (W = 32, or 8 when queries are sparse against the map)
//   1. takes the union of its lanes' cell boxes.  Cells are x-fastest, so every (y, z) row of the union box is
//      ONE contiguous run of the sorted target array;
//   2. its lanes read the runs' bounds (two cell_start loads per row, one row per lane), prefix-sum the lengths
//      and each lane issues ONE bulk-async copy (cp.async.bulk global -> shared, completion on an mbarrier:
//      the TMA engine, `UBLKCP` + `SYNCS` in SASS) of its run into the group's staging buffer;
//   3. after the mbarrier flips, all lanes scan the SAME staged candidates: one broadcast LDS.128 per candidate
//      per warp, the distance and a three-smallest network per lane (coop_scan_chunk).  Trip counts are uniform,
//      no lane waits on a load.
// Exactness is what it was: float d2 in FLANN's operation order, ties to the smallest original index, and the
// caller's cell box (cell_box(), nncache.cuh) contains every point within its radius; `lrest` is the distance
// from the query to the outside of the box that was actually scanned (the union), a lower bound for every point
// not seen.
//
// Lanes whose boxes lie far from the group's first pending lane wait for a later pass (`join_d` cells of slack),
// so one far-away query cannot make the whole group scan a huge box.
#pragma once
#include "common.cuh"
#include "grid.cuh"
#include "nncache.cuh"
#include "tma.cuh"

namespace b2 {

constexpr int kCoopCap = 256;  // staged candidates per warp and chunk (4 KB of shared memory)

// Per-lane state of a cooperative scan: the two nearest candidates and the third distance.
struct CoopTop {
  unsigned long long k0;  // best (d2, original index); kInfKey = nothing seen
  float b1, b2;           // second and third smallest d2 (+inf = none)
  int p0, p1;             // best / second: >= 0 slot in the chunk being scanned, -1 none, -2 held in m0, -3 held in m1
  float4 m0, m1;          // candidates of earlier chunks: xyz + original index
  float lrest;            // lower bound on the distance to every target point that was not scanned
};

__device__ __forceinline__ void coop_init(CoopTop& t) {
  t.k0 = kInfKey;
  t.b1 = t.b2 = INFINITY;
  t.p0 = t.p1 = -1;
  t.m0 = t.m1 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
  t.lrest = INFINITY;
}

// branch-free insertion of candidate (d, original index idx) staged in slot j; `valid` false = no-op
__device__ __forceinline__ void coop_insert(CoopTop& t, bool valid, float d, int idx, int j) {
  const unsigned long long k = pack_key(d, idx);
  const bool nb = valid && k < t.k0;  // new best
  const float d0 = key_d2(t.k0);
  const float dl = nb ? d0 : d;  // the loser of each comparison goes on to the next rank
  const int pl = nb ? t.p0 : j;
  t.k0 = nb ? k : t.k0;
  t.p0 = nb ? j : t.p0;
  const bool ns = valid && dl < t.b1;  // new second
  const float dl2 = ns ? t.b1 : dl;
  t.p1 = ns ? pl : t.p1;
  t.b1 = ns ? dl : t.b1;
  t.b2 = valid ? fminf(t.b2, dl2) : t.b2;
}

// end of a chunk: candidates that live in the staging buffer move to registers before the buffer is reused
__device__ __forceinline__ void coop_resolve(CoopTop& t, const float4* buf) {
  const float4 n0 = t.p0 >= 0 ? buf[t.p0] : t.m0;
  const float4 n1 = t.p1 >= 0 ? buf[t.p1] : (t.p1 == -2 ? t.m0 : t.m1);
  t.m0 = n0;
  t.m1 = n1;
  t.p0 = t.p0 == -1 ? -1 : -2;
  t.p1 = t.p1 == -1 ? -1 : -3;
}

__device__ __forceinline__ float4 coop_c0(const CoopTop& t) {
  return t.p0 != -1 ? t.m0 : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
}
__device__ __forceinline__ float4 coop_c1(const CoopTop& t) {
  return t.p1 != -1 ? t.m1 : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
}
// lower bound on the distance to every target point other than the two cached ones
__device__ __forceinline__ float coop_bound(const CoopTop& t) {
  const float l3 = t.b2 < INFINITY ? __fmul_rd(sqrt_fast(t.b2), kRelDown) : INFINITY;
  return fminf(l3, t.lrest);
}

// ---- the scan of one staged chunk -----------------------------------------------------------------------------
// Fast path: a three-smallest network over  v = (bits(d2) & ~0xFF) | slot  (15 instructions per candidate instead of
// the 24 of the exact bookkeeping).  v is monotone in d2 (non-negative floats order like their bits; the low 8
// mantissa bits are traded for the slot number), so:
//   * if the two smallest v differ above the low 8 bits, the smallest one is the UNIQUE nearest candidate of the
//     chunk (every other candidate's d2 is at least one 2^-15 bucket larger) — no tie to break;
//   * the second smallest names a candidate whose bucket is minimal among the rest: a valid runner-up to cache
//     (the certificate only needs its bound to hold for every point other than the two cached ones);
//   * (v2 & ~0xFF) as a float is a lower bound of the d2 of every other candidate of the chunk.
// The chunk's two winners are re-evaluated exactly and merged into the running exact state; when the two smallest
// share a bucket (a tie or a near tie: duplicated map points, lattices) the whole warp falls back to the exact scan
// of the chunk, which resolves ties to the smallest original index.
template <int W>
__device__ __forceinline__ void coop_scan_chunk(CoopTop& t, const float4* gbuf, int cnt, int max_cnt, float qx, float qy,
                                                float qz, bool mine) {
  int v0 = 0x7FFFFFFF, v1 = 0x7FFFFFFF, v2 = 0x7FFFFFFF;
#pragma unroll 4
  for (int j = 0; j < max_cnt; ++j) {
    const float4 p = gbuf[j];
    const float d = sqdist3(qx, qy, qz, p.x, p.y, p.z);
    int v = (int)((__float_as_uint(d) & 0xFFFFFF00u) | (unsigned int)j);
    if (W != 32) v = j < cnt ? v : 0x7FFFFFFF;
    const int t1 = max(v0, v);
    v0 = min(v0, v);
    const int t2 = max(v1, t1);
    v1 = min(v1, t1);
    v2 = min(v2, t2);
  }
  const bool tie = mine && v1 != 0x7FFFFFFF && ((v0 ^ v1) >> 8) == 0;
  if (!__any_sync(0xFFFFFFFFu, tie)) {
    if (v0 != 0x7FFFFFFF) {
      const int j0 = v0 & 0xFF;
      const float4 p = gbuf[j0];
      coop_insert(t, true, sqdist3(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w), j0);
    }
    if (v1 != 0x7FFFFFFF) {
      const int j1 = v1 & 0xFF;
      const float4 p = gbuf[j1];
      coop_insert(t, true, sqdist3(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w), j1);
    }
    if (v2 != 0x7FFFFFFF) t.b2 = fminf(t.b2, __uint_as_float((unsigned int)v2 & 0xFFFFFF00u));
  } else {  // exact bookkeeping for every candidate
#pragma unroll 2
    for (int j = 0; j < max_cnt; ++j) {
      const float4 p = gbuf[j];
      const float d = sqdist3(qx, qy, qz, p.x, p.y, p.z);
      coop_insert(t, W == 32 ? true : j < cnt, d, __float_as_int(p.w), j);
    }
  }
  coop_resolve(t, gbuf);
}

// Search for every lane with `want` set: exact nearest / second nearest / a lower bound of the third distance over
// the cells of its box `bx` (at least — the union box of its group is scanned).  `g` is warp-uniform.  `wbuf` is the
// warp's staging buffer (kCoopCap points), `bar` its mbarrier, `phase` the parity the next wait expects (kept by the
// caller).  Lanes without `want` leave `out` untouched.
template <int W>
__device__ __forceinline__ void coop_search(const GridView& g, bool want, float qx, float qy, float qz, const CellBox& bx,
                                            int join_d, float4* wbuf, unsigned long long* bar, unsigned int& phase,
                                            CoopTop& out) {
  constexpr unsigned int kFull = 0xFFFFFFFFu;
  constexpr int kGroups = 32 / W;
  constexpr int kCapG = kCoopCap / kGroups;
  constexpr int R = kGroups;  // rows per lane and step: every group enumerates 32 rows per step
  const int lane = threadIdx.x & 31, sub = lane & (W - 1);
  const unsigned int gmask = W == 32 ? kFull : (((1u << W) - 1u) << (lane - sub));
  float4* const gbuf = wbuf + (lane / W) * kCapG;
  bool pending = want;
  unsigned int pend = __ballot_sync(kFull, pending);
  while (pend) {
    // ---- who searches in this pass: the group's first pending lane and every pending lane whose box is nearby
    const unsigned int gp = pend & gmask;
    const int leader = gp ? __ffs(gp) - 1 : lane;
    const int lxa = __shfl_sync(kFull, bx.xa, leader), lxb = __shfl_sync(kFull, bx.xb, leader);
    const int lya = __shfl_sync(kFull, bx.ya, leader), lyb = __shfl_sync(kFull, bx.yb, leader);
    const int lza = __shfl_sync(kFull, bx.za, leader), lzb = __shfl_sync(kFull, bx.zb, leader);
    const bool join = pending && bx.xa >= lxa - join_d && bx.xb <= lxb + join_d && bx.ya >= lya - join_d &&
                      bx.yb <= lyb + join_d && bx.za >= lza - join_d && bx.zb <= lzb + join_d;
    const int xa = __reduce_min_sync(gmask, join ? bx.xa : 0x7FFFFFFF), xb = __reduce_max_sync(gmask, join ? bx.xb : -1);
    const int ya = __reduce_min_sync(gmask, join ? bx.ya : 0x7FFFFFFF), yb = __reduce_max_sync(gmask, join ? bx.yb : -1);
    const int za = __reduce_min_sync(gmask, join ? bx.za : 0x7FFFFFFF), zb = __reduce_max_sync(gmask, join ? bx.zb : -1);
    const int ny = gp ? yb - ya + 1 : 0, nz = gp ? zb - za + 1 : 0;
    const int nrow = ny * nz;
    const int max_nrow = W == 32 ? nrow : __reduce_max_sync(kFull, nrow);
    CoopTop t;
    coop_init(t);
    if (join) {  // distance from q to the outside of the union box (faces that have cells beyond them only)
      float gmin = INFINITY;
      if (xa > 0) gmin = fminf(gmin, __fsub_rd(qx, __fadd_ru(g.ox, __fmul_ru((float)xa, g.cell))));
      if (xb < g.nx - 1) gmin = fminf(gmin, __fsub_rd(__fadd_rd(g.ox, __fmul_rd((float)(xb + 1), g.cell)), qx));
      if (ya > 0) gmin = fminf(gmin, __fsub_rd(qy, __fadd_ru(g.oy, __fmul_ru((float)ya, g.cell))));
      if (yb < g.ny - 1) gmin = fminf(gmin, __fsub_rd(__fadd_rd(g.oy, __fmul_rd((float)(yb + 1), g.cell)), qy));
      if (za > 0) gmin = fminf(gmin, __fsub_rd(qz, __fadd_ru(g.oz, __fmul_ru((float)za, g.cell))));
      if (zb < g.nz - 1) gmin = fminf(gmin, __fsub_rd(__fadd_rd(g.oz, __fmul_rd((float)(zb + 1), g.cell)), qz));
      t.lrest = fmaxf(__fsub_rd(gmin, g.slack), 0.0f);
    }
    // ---- the rows of the union box, 32 per group and step (R per lane): bounds -> offsets -> bulk copies -> one
    // shared scan
    for (int r0 = 0; r0 < max_nrow; r0 += 32) {
      int s[R], len[R];
      int mine_total = 0;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int r = r0 + sub * R + k;
        s[k] = 0;
        len[k] = 0;
        if (r < nrow) {
          const int y = ya + r % ny, z = za + r / ny;
          const int* row = g.cell_start + (z * g.ny + y) * g.nx;
          s[k] = __ldg(row + xa);
          len[k] = __ldg(row + xb + 1) - s[k];
        }
        mine_total += len[k];
      }
      int incl = mine_total;
#pragma unroll
      for (int o = 1; o < W; o <<= 1) {
        const int up = __shfl_up_sync(kFull, incl, o, W);
        if (sub >= o) incl += up;
      }
      const int total = __shfl_sync(kFull, incl, W - 1, W);  // candidates of this row step in my group
      const int max_total = W == 32 ? total : __reduce_max_sync(kFull, total);
      for (int c0 = 0; c0 < max_total; c0 += kCapG) {
        unsigned int nb = 0;
        int off = incl - mine_total;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int lo = max(off, c0), hi = min(off + len[k], c0 + kCapG);
          if (hi > lo) nb += (unsigned int)(hi - lo) * 16u;
          off += len[k];
        }
        const unsigned int tot = __reduce_add_sync(kFull, nb);  // > 0: c0 < max_total
        if (lane == 0) mbar_expect_tx(bar, tot);
        __syncwarp();
        off = incl - mine_total;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int lo = max(off, c0), hi = min(off + len[k], c0 + kCapG);
          if (hi > lo) bulk_g2s(gbuf + (lo - c0), g.pts + s[k] + (lo - off), (unsigned int)(hi - lo) * 16u, bar);
          off += len[k];
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        coop_scan_chunk<W>(t, gbuf, min(total - c0, kCapG), min(max_total - c0, kCapG), qx, qy, qz, join);
        __syncwarp();  // every lane is done with the buffer before the next copies land in it
      }
    }
    if (join) {
      out = t;
      pending = false;
    }
    pend = __ballot_sync(kFull, pending);
  }
}

}  // namespace b2
