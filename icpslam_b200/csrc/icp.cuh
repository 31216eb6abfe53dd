// icp.cuh — K2+K3+K4 fused: one launch = one ICP iteration of pcl::IterativeClosestPoint
// (SURVEY.md App. A.3), for every scan of a batch (blockIdx.y).
//
//   per query    q = transformation_ * q          (in-place float chain, exactly PCL's
//                                                  transformCloud(input_transformed, ..., in place))
//                exact 1-NN in the target grid    (cached-neighbour certificate, else box search: nncache.cuh)
//                gate  d2 <= max_dist^2           (CorrespondenceEstimation keeps equality)
//                16 running sums + sum(d2)        (fp64 accumulators of the lane that settled the query)
//   per warp     xor-butterfly -> partials[slab][17]
//   last warp    fixed-order sum over the slabs' partials (deterministic), Umeyama/SVD,
//                final = T * final, DefaultConvergenceCriteria -> IcpState (solve.cuh)
// No host round trip: `done` in IcpState turns the remaining launches of the batch into no-ops.
// HBM bytes per query per iteration: 48 read (running point with its bound in .w, two cached neighbours) + 16
// written (running point + bound), + 32 more written and the target points of the scanned cells when the query is searched;
// the reduction adds 136 B per warp.  Also here: getFitnessScore, the correspondence write-out, transformPointCloud.
#pragma once
#include "common.cuh"
#include "nn.cuh"
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

// One ICP iteration = one launch over all scans of the batch (nncache.cuh explains the certificate).
//
// Per-query state carried between iterations (ScanTask): cur (running cloud), c0 / c1 (nearest and
// runner-up target point: xyz + original index in .w, index -1 = none); the distance bound is cur.w.  All of it
// is read and written with coalesced, evict-first 16-byte accesses; nothing in the streaming pass depends
// on a gathered load.  corr_idx / corr_d2 are produced once, after the loop (icp_finalize_corr).
//
// The unit of work is a WARP and its slab of 32 * QPT consecutive queries; warps never wait for each
// other (no CTA barrier anywhere), so the scheduler always has warps in different phases to pick from:
//   A  every query of the slab, 32 at a time: q = T_inc * q (in place), distances to its two cached
//      candidates, certificate test.  Passers add their pair to the lane's 17 fp64 sums at once; the
//      others are compacted (ballot + popc: the order is a function of the data only) into the warp's
//      work list in shared memory;
//   B  the work list, one entry per lane: exact box search, new candidates + bound, pair added to the
//      sums of the lane that searched it;
//   C  xor-butterfly over the lanes -> partials[slab]; the last warp of the scan (atomic ticket) adds the
//      slabs' partials in a fixed order and runs Umeyama + the convergence test (solve.cuh).
// Which lane sums which pair depends only on the data and every reduction has a fixed order, so results
// are bit-reproducible from run to run.
template <int QPT>
__global__ void __launch_bounds__(kSweepThreads, kSweepMinCtas) icp_sweep_p2p(const ScanTask* __restrict__ tasks, IcpConfig cfg) {
  constexpr int kWarps = kSweepThreads / 32;
  constexpr int kWarpSlab = 32 * QPT;
  __shared__ NNScratch<kSweepThreads> sc;             // one column per thread: private to its warp by construction
  __shared__ float sT[kWarps][16];
  __shared__ unsigned short wl_id[kWarps][kWarpSlab];  // work list of the warp: query index inside its slab
  const ScanTask& t = tasks[blockIdx.y];
  // the fields the row loop uses, read once: the loop stores through float4 pointers, which the compiler has to
  // assume may alias the task record and would otherwise reload them after every store
  const int tn = t.n;
  float4* const tcur = t.cur;
  const float4* const tsrc = t.src;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nslab = (tn + kWarpSlab - 1) / kWarpSlab;
  const int slab = blockIdx.x * kWarps + warp;
  if (slab >= nslab) return;
  IcpState* st = t.state;
  if (st->done) return;
  const bool first = st->iter == 0;
  if (lane < 16) sT[warp][lane] = st->Tinc[lane];
  __syncwarp();
  const float* T = sT[warp];
  const int base = slab * kWarpSlab;
  const float margin = cfg.margin_frac * t.grid.cell;

  float4* const cand[3] = {t.c0, t.c1, t.c2};  // the first kCacheK are used

  double acc[kNumSums];
#pragma unroll
  for (int c = 0; c < kNumSums; ++c) acc[c] = 0.0;

  // ---- phase A
  int wc = 0;  // entries in the warp's work list (warp-uniform)
#pragma unroll kPhaseAUnroll
  for (int qi = 0; qi < QPT; ++qi) {
    const int i = base + qi * 32 + lane;
    bool need = false;
    if (i < tn) {
      float4 p, c[kCacheK];
      float lb = 0.0f;
      if (first) {
        p = __ldg(tsrc + i);
      } else {  // independent coalesced loads
        p = ld_stream(tcur + i);
#pragma unroll
        for (int k = 0; k < kCacheK; ++k) c[k] = ld_stream(cand[k] + i);
        lb = p.w;  // the bound travels in the running point's fourth component
      }
      float4 q = xform_f(T, p.x, p.y, p.z);
      q.w = 0.0f;
      if (!(isfinite(q.x) && isfinite(q.y) && isfinite(q.z))) {
        atomicOr(&st->pad, 1);  // non-finite source point or transform: reported by the last warp
#pragma unroll
        for (int k = 0; k < kCacheK; ++k) cand[k][i] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
      } else if (first) {
        need = true;
      } else {
        // bound after this iteration's motion (upper-rounded step, lower-rounded difference)
        const float step = __fmul_ru(sqrt_fast(sqdist3(q.x, q.y, q.z, p.x, p.y, p.z)), kRelUp);
        const float L = __fsub_rd(lb, step);
        unsigned long long k0 = kInfKey;
        int arg = 0;
#pragma unroll
        for (int k = 0; k < kCacheK; ++k) {
          const int id = __float_as_int(c[k].w);
          const unsigned long long kk = id >= 0 ? pack_key(sqdist3(q.x, q.y, q.z, c[k].x, c[k].y, c[k].z), id) : kInfKey;
          if (kk < k0) {
            k0 = kk;
            arg = k;
          }
        }
        float4 c0 = c[0];
#pragma unroll
        for (int k = 1; k < kCacheK; ++k)
          if (arg == k) {  // keep c0 = the nearest of the cached points
            st_stream(cand[0] + i, c[k]);
            st_stream(cand[k] + i, c[0]);
            c0 = c[k];
          }
        const float d2 = key_d2(k0);
        const float L2 = L > 0.0f ? __fmul_rd(__fmul_rd(L, L), kRelDown) : 0.0f;
        if (fminf(d2, cfg.bound2) < L2) {  // certificate holds: the NN is c0, or nothing is within the bound
          q.w = L;
          if ((d2 < L2) && !((double)d2 > cfg.max2)) accumulate_pair(acc, q, c0, d2);
        } else {
          need = true;
        }
      }
      st_stream(tcur + i, q);
    }
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, need);
    if (need) wl_id[warp][wc + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)(qi * 32 + lane);
    wc += __popc(bal);
  }
  __syncwarp();

  // ---- phase B
  for (int e = lane; e < wc; e += 32) {
    const int i = base + (int)wl_id[warp][e];
    const float4 q = tcur[i];
    // search radius: the nearer cached candidate (the list keeps ids only, so that long slabs fit in shared
    // memory; the two points are L2-hot), or the probe when there is none
    float seed = INFINITY;
    if (!first) {
#pragma unroll
      for (int k = 0; k < kCacheK; ++k) {
        const float4 ck = cand[k][i];
        if (__float_as_int(ck.w) >= 0) seed = fminf(seed, sqdist3(q.x, q.y, q.z, ck.x, ck.y, ck.z));
      }
    }
    if (!(seed < INFINITY)) seed = probe_seed(t.grid, q.x, q.y, q.z);
    const CellBox bx = cell_box(t.grid, q.x, q.y, q.z, seed, cfg.bound2, margin, cfg.max_rings);
    Top3 top;
    float lrest;
    box_search<kSweepThreads>(t.grid, q.x, q.y, q.z, bx, sc, top, lrest);
    const float4 none = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    const float4 m0 = top.p[0] >= 0 ? __ldg(t.grid.pts + top.p[0]) : none;
    st_stream(cand[0] + i, m0);
#pragma unroll
    for (int k = 1; k < kCacheK; ++k) st_stream(cand[k] + i, top.p[k] >= 0 ? __ldg(t.grid.pts + top.p[k]) : none);
    st_stream(tcur + i, make_float4(q.x, q.y, q.z, top3_bound(top, lrest)));
    const float d2 = key_d2(top.k0);
    if ((top.k0 != kInfKey) && !((double)d2 > cfg.max2)) accumulate_pair(acc, q, m0, d2);
  }

  // ---- phase C
  double mine = 0.0;  // lane c < kNumSums ends up with sum c of the slab
#pragma unroll
  for (int c = 0; c < kNumSums; ++c) {
    const double s = warp_sum(acc[c]);
    if (lane == c) mine = s;
  }
  if (lane < kNumSums) t.partials[(size_t)slab * kNumSums + lane] = mine;
  __threadfence();
  __syncwarp();
  unsigned int ticket = 0;
  if (lane == 0) {
    if (wc) atomicAdd(&st->unresolved, (unsigned int)wc);
    ticket = atomicAdd(&st->ticket, 1u);
  }
  ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
  if (ticket != (unsigned int)(nslab - 1)) return;
  __threadfence();
  // last warp of the scan: lane l adds slabs l, l + 32, ... in order, then the butterfly
#pragma unroll
  for (int c = 0; c < kNumSums; ++c) acc[c] = 0.0;
  for (int b = lane; b < nslab; b += 32) {
    const double* p = t.partials + (size_t)b * kNumSums;
#pragma unroll
    for (int c = 0; c < kNumSums; ++c) acc[c] += __ldcg(p + c);
  }
  __shared__ double s_sums[kWarps][kNumSums];
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
// returns) from the per-query state: c0 is the exact nearest neighbour of cur whenever one lies within
// the gate, so  idx = c0.idx if d2 <= max_dist^2 else -1.
__global__ void __launch_bounds__(256) icp_finalize_corr(const ScanTask* __restrict__ tasks, IcpConfig cfg) {
  const ScanTask& t = tasks[blockIdx.y];
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= t.n) return;
  const float4 q = t.cur[i];
  const float4 m = t.c0[i];
  const int id = __float_as_int(m.w);
  const float d2 = id >= 0 ? sqdist3(q.x, q.y, q.z, m.x, m.y, m.z) : INFINITY;
  t.corr_idx[i] = (id >= 0 && !((double)d2 > cfg.max2)) ? id : -1;
  t.corr_d2[i] = d2;
}

// ---- getFitnessScore(max_range): transform by final_T, exact unbounded 1-NN, mean of d2 <= range
__global__ void __launch_bounds__(kSweepThreads) fitness_kernel(const ScanTask* __restrict__ task, int max_rings,
                                                                int* __restrict__ unresolved_list,
                                                                unsigned int* __restrict__ unresolved_count,
                                                                float4* __restrict__ q_out, int* __restrict__ idx_out,
                                                                float* __restrict__ d2_out) {
  __shared__ float sT[16];
  __shared__ NNScratch<kSweepThreads> sc;
  const ScanTask& t = *task;
  if (threadIdx.x < 16) sT[threadIdx.x] = t.state->final_T[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * kSweepThreads + threadIdx.x;
  if (i >= t.n) return;
  const float4 p = __ldg(t.src + i);
  const float4 q = xform_f(sT, p.x, p.y, p.z);
  q_out[i] = q;
  unsigned long long seed = kInfKey;
  if (t.pad) {
    // the point-to-point loop left its certificate (nncache.cuh) for cur[i], which is final_T * src[i] up to
    // the last increment and float rounding: the same triangle-inequality test settles most queries here too
    const float4 c = t.cur[i];
    const float step = __fmul_ru(sqrt_fast(sqdist3(q.x, q.y, q.z, c.x, c.y, c.z)), kRelUp);
    const float L = __fsub_rd(c.w, step);
    const float4* const cand[3] = {t.c0, t.c1, t.c2};
#pragma unroll
    for (int k = 0; k < kCacheK; ++k) {
      const float4 ck = cand[k][i];
      const int id = __float_as_int(ck.w);
      const unsigned long long kk = id >= 0 ? pack_key(sqdist3(q.x, q.y, q.z, ck.x, ck.y, ck.z), id) : kInfKey;
      seed = kk < seed ? kk : seed;
    }
    const float L2 = L > 0.0f ? __fmul_rd(__fmul_rd(L, L), kRelDown) : 0.0f;
    if (key_d2(seed) < L2) {
      idx_out[i] = key_idx(seed);
      d2_out[i] = key_d2(seed);
      return;
    }
  }
  NNResult r = grid_nn<kSweepThreads>(t.grid, q.x, q.y, q.z, INFINITY, max_rings, -1, sc, seed);
  if (!r.resolved) {
    unsigned int slot = atomicAdd(unresolved_count, 1u);
    unresolved_list[slot] = i;
    return;
  }
  idx_out[i] = (r.key == kInfKey) ? -1 : key_idx(r.key);
  d2_out[i] = key_d2(r.key);
}

// deterministic mean of d2 over idx >= 0 && d2 <= max_range: single CTA, fixed order
__global__ void __launch_bounds__(1024) fitness_reduce(const int* __restrict__ idx, const float* __restrict__ d2,
                                                       int n, double max_range, IcpState* __restrict__ st) {
  __shared__ double ssum[32];
  __shared__ unsigned long long scnt[32];
  double s = 0;
  unsigned long long c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float d = d2[i];
    if (idx[i] >= 0 && (double)d <= max_range) {
      s += (double)d;
      ++c;
    }
  }
  s = warp_sum(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  if ((threadIdx.x & 31) == 0) {
    ssum[threadIdx.x >> 5] = s;
    scnt[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0;
    unsigned long long tc = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      ts += ssum[w];
      tc += scnt[w];
    }
    st->fitness_sum = ts;
    st->fitness_cnt = tc;
  }
}

// ---- pcl::transformPointCloud (App. A.5) ------------------------------------------------------
__global__ void __launch_bounds__(256) transform_cloud_f(const float4* __restrict__ in, int n, const float* __restrict__ T,
                                                        float4* __restrict__ out) {
  __shared__ float sT[16];
  if (threadIdx.x < 16) sT[threadIdx.x] = T[threadIdx.x];
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(in + i);
  out[i] = xform_f(sT, p.x, p.y, p.z);
}

__global__ void __launch_bounds__(256) transform_cloud_d(const float4* __restrict__ in, int n, const double* __restrict__ T,
                                                        float4* __restrict__ out) {
  __shared__ double sT[16];
  if (threadIdx.x < 16) sT[threadIdx.x] = T[threadIdx.x];
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(in + i);
  double x = p.x, y = p.y, z = p.z;
  float4 o;
  o.x = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sT[0], x), __dmul_rn(sT[1], y)), __dmul_rn(sT[2], z)), sT[3]);
  o.y = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sT[4], x), __dmul_rn(sT[5], y)), __dmul_rn(sT[6], z)), sT[7]);
  o.z = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sT[8], x), __dmul_rn(sT[9], y)), __dmul_rn(sT[10], z)), sT[11]);
  o.w = 1.0f;
  out[i] = o;
}

}  // namespace b2
