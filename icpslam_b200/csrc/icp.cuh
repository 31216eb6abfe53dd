// icp.cuh — K2+K3+K4 fused: one launch = one ICP iteration of pcl::IterativeClosestPoint
// (SURVEY.md App. A.3), for every scan of a batch (blockIdx.y).
//
//   per thread   q = transformation_ * q          (in-place float chain, exactly PCL's
//                                                  transformCloud(input_transformed, ..., in place))
//                exact 1-NN in the target grid    (nn.cuh)
//                gate  d2 <= max_dist^2           (CorrespondenceEstimation keeps equality)
//                16 running sums + sum(d2)        (fp64 accumulators)
//   per CTA      warp-shuffle reduction, then a fixed-order cross-warp sum -> partials[cta][17]
//   last CTA     fixed-order tree over the per-CTA partials (deterministic), Umeyama/SVD,
//                final = T * final, DefaultConvergenceCriteria -> IcpState (solve.cuh)
// No host round trip: `done` in IcpState turns the remaining launches of the batch into no-ops.
// HBM/L2 bytes per query per iteration: 16 read + 16 write (running cloud) + 8 write (idx, d2)
// + the target points of the touched cells; the reduction adds 136 B per CTA.
#pragma once
#include "common.cuh"
#include "nn.cuh"
#include "nncache.cuh"
#include "solve.cuh"

namespace b2 {

// Sum `acc[kNumSums]` over the CTA in a fixed order; result valid in warp 0 lane 0's acc.
__device__ __forceinline__ void cta_reduce_sums(double* acc, double (*smem)[kNumSums]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < kNumSums; ++c) acc[c] = warp_sum(acc[c]);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < kNumSums; ++c) smem[warp][c] = acc[c];
  }
  __syncthreads();
  if (warp == 0) {
    if (lane < kNumSums) {
      double s = 0;
#pragma unroll
      for (int w = 0; w < kSweepThreads / 32; ++w) s += smem[w][lane];
      smem[0][lane] = s;
    }
  }
  __syncthreads();
}

// Called by every thread of every CTA after cta_reduce_sums; returns true in ALL threads of the last
// CTA to arrive (for this scan), with the grid-wide sums left in smem[0][0..kNumSums).
__device__ __forceinline__ bool grid_reduce_last(const ScanTask& t, double (*smem)[kNumSums], int ncta) {
  __shared__ unsigned int s_ticket;
  if (threadIdx.x < kNumSums) t.partials[(size_t)blockIdx.x * kNumSums + threadIdx.x] = smem[0][threadIdx.x];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&t.state->ticket, 1u);
  __syncthreads();
  if (s_ticket != (unsigned int)(ncta - 1)) return false;
  __threadfence();
  // fixed-order reduction over CTAs: thread j sums CTAs j, j+256, ... then the CTA-wide fixed tree
  double acc[kNumSums];
#pragma unroll
  for (int c = 0; c < kNumSums; ++c) acc[c] = 0.0;
  for (int b = threadIdx.x; b < ncta; b += kSweepThreads) {
    const double* p = t.partials + (size_t)b * kNumSums;
#pragma unroll
    for (int c = 0; c < kNumSums; ++c) acc[c] += __ldcg(p + c);
  }
  __syncthreads();
  cta_reduce_sums(acc, smem);
  return true;
}

// QPT = queries per thread: consecutive 256-point slabs handled by the same CTA, so that the prologue,
// the 17-value fp64 reduction and the CTA barriers are paid once per QPT queries, and so that the
// compacted list of queries that need a real search is long enough to fill the CTA's warps.
//
// Three phases per CTA (nncache.cuh explains the certificate):
//   A  every query: q = T_inc * q (in place), distances to its two cached candidates, certificate
//      test; passers get their correspondence at once, the others are compacted (ballot + prefix, so
//      the list order is deterministic) into a shared-memory work list;
//   B  the work list, one entry per thread: exact box search, new candidates + bound;
//   C  every query again, in (thread, slab) order: the 17 fp64 sums from the stored correspondences —
//      the summation order does not depend on which queries were searched, so results are
//      bit-reproducible, and the 34 accumulator registers are not live during the search.
template <int QPT>
__global__ void __launch_bounds__(kSweepThreads, kSweepMinCtas) icp_sweep_p2p(const ScanTask* __restrict__ tasks, IcpConfig cfg) {
  constexpr int kWarps = kSweepThreads / 32;
  constexpr int kSlab = QPT * kSweepThreads;  // queries per CTA
  __shared__ double smem[kWarps][kNumSums];
  __shared__ NNScratch<kSweepThreads> sc;
  __shared__ float sT[16];
  __shared__ int s_flags[2];
  __shared__ float wl_thr[kSlab];            // per local query: squared distance of its best cached candidate / probe
  __shared__ unsigned short wl_sorted[kSlab];  // work list: local query ids, cheapest cost class first
  __shared__ int s_cnt[kCostClasses][kWarps];
  __shared__ int s_base[kCostClasses][kWarps];
  __shared__ int s_total;
  const ScanTask& t = tasks[blockIdx.y];
  const int ncta = (t.n + kSlab - 1) / kSlab;
  if ((int)blockIdx.x >= ncta) return;
  IcpState* st = t.state;
  if (threadIdx.x < 16) sT[threadIdx.x] = st->Tinc[threadIdx.x];
  if (threadIdx.x == 16) s_flags[0] = st->done;
  if (threadIdx.x == 17) s_flags[1] = st->iter;
  __syncthreads();
  if (s_flags[0]) return;
  const bool first = s_flags[1] == 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int base = blockIdx.x * kSlab;
  const float margin = cfg.margin_frac * t.grid.cell;

  // ---- phase A
  unsigned packed = 0;  // 4 bits per slab: cost class of this thread's query, 15 = no search needed
  int mycnt = 0;        // lane c < kCostClasses: entries of class c in this warp
#pragma unroll 1
  for (int qi = 0; qi < QPT; ++qi) {
    const int i = base + qi * kSweepThreads + threadIdx.x;
    int cls = 15;
    if (i < t.n) {
      const float4 p = first ? __ldg(t.src + i) : t.cur[i];
      const float4 q = xform_f(sT, p.x, p.y, p.z);
      t.cur[i] = q;
      float seed = INFINITY;
      bool need = false;
      if (!(isfinite(q.x) && isfinite(q.y) && isfinite(q.z))) {
        atomicOr(&st->pad, 1);  // non-finite source point or transform: reported by the last CTA
        t.corr_idx[i] = -1;
        t.corr_d2[i] = INFINITY;
        t.cand[i] = make_int2(-1, -1);
        t.lb[i] = 0.0f;
      } else if (first) {
        need = true;
        seed = probe_seed(t.grid, q.x, q.y, q.z);
      } else {
        int2 c = t.cand[i];
        // bound after this iteration's motion (upper-rounded step, lower-rounded difference)
        const float step = __fmul_ru(__fsqrt_ru(sqdist3(q.x, q.y, q.z, p.x, p.y, p.z)), kRelUp);
        const float L = __fsub_rd(t.lb[i], step);
        unsigned long long k0 = kInfKey, k1 = kInfKey;
        if (c.x >= 0) {
          const float4 m = __ldg(t.grid.pts + c.x);
          k0 = pack_key(sqdist3(q.x, q.y, q.z, m.x, m.y, m.z), __float_as_int(m.w));
        }
        if (c.y >= 0) {
          const float4 m = __ldg(t.grid.pts + c.y);
          k1 = pack_key(sqdist3(q.x, q.y, q.z, m.x, m.y, m.z), __float_as_int(m.w));
        }
        if (k1 < k0) {  // keep cand.x = the nearer of the two
          k0 = k1;
          c = make_int2(c.y, c.x);
          t.cand[i] = c;
        }
        const float d2 = key_d2(k0);
        const float L2 = L > 0.0f ? __fmul_rd(__fmul_rd(L, L), kRelDown) : 0.0f;
        if (fminf(d2, cfg.bound2) < L2) {  // certificate holds: the NN is cand.x, or nothing is within the bound
          const bool keep = (d2 < L2) && !((double)d2 > cfg.max2);
          t.corr_idx[i] = keep ? key_idx(k0) : -1;
          t.corr_d2[i] = d2;
          t.lb[i] = L;
        } else {
          need = true;
          seed = d2;
        }
      }
      if (need) {
        wl_thr[qi * kSweepThreads + threadIdx.x] = seed;
        cls = cost_class(cell_box(t.grid, q.x, q.y, q.z, seed, cfg.bound2, margin, cfg.max_rings));
      }
    }
    packed |= (unsigned)cls << (4 * qi);
#pragma unroll
    for (int c = 0; c < kCostClasses; ++c) {
      const unsigned bal = __ballot_sync(0xFFFFFFFFu, cls == c);
      if (lane == c) mycnt += __popc(bal);
    }
  }
  if (lane < kCostClasses) s_cnt[lane][warp] = mycnt;
  __syncthreads();
  if (threadIdx.x < kCostClasses * kWarps) {  // exclusive prefix over (class, warp), class-major
    const int c = threadIdx.x / kWarps, w = threadIdx.x % kWarps;
    int b = 0;
    for (int k = 0; k < c * kWarps + w; ++k) b += s_cnt[k / kWarps][k % kWarps];
    s_base[c][w] = b;
    if (threadIdx.x == kCostClasses * kWarps - 1) s_total = b + s_cnt[c][w];
  }
  __syncthreads();
  {
    int run[kCostClasses];
#pragma unroll
    for (int c = 0; c < kCostClasses; ++c) run[c] = s_base[c][warp];
#pragma unroll 1
    for (int qi = 0; qi < QPT; ++qi) {
      const int cls = (packed >> (4 * qi)) & 15;
#pragma unroll
      for (int c = 0; c < kCostClasses; ++c) {
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, cls == c);
        if (cls == c) wl_sorted[run[c] + __popc(bal & lt_mask)] = (unsigned short)(qi * kSweepThreads + threadIdx.x);
        run[c] += __popc(bal);
      }
    }
  }
  __syncthreads();

  // ---- phase B
  const int total = s_total;
  for (int e = threadIdx.x; e < total; e += kSweepThreads) {
    const int loc = wl_sorted[e];
    const int i = base + loc;
    const float4 q = t.cur[i];
    const CellBox bx = cell_box(t.grid, q.x, q.y, q.z, wl_thr[loc], cfg.bound2, margin, cfg.max_rings);
    Top3 top;
    float lrest;
    box_search<kSweepThreads>(t.grid, q.x, q.y, q.z, bx, sc, top, lrest);
    const float d2 = key_d2(top.k0);
    const bool keep = (top.k0 != kInfKey) && !((double)d2 > cfg.max2);
    t.corr_idx[i] = keep ? key_idx(top.k0) : -1;
    t.corr_d2[i] = d2;
    t.cand[i] = make_int2(top.p0, top.p1);
    t.lb[i] = top3_bound(top, lrest);
  }
  __syncthreads();
  if (threadIdx.x == 0 && total) atomicAdd(&st->unresolved, (unsigned int)total);

  // ---- phase C
  double acc[kNumSums];
#pragma unroll
  for (int c = 0; c < kNumSums; ++c) acc[c] = 0.0;
#pragma unroll 1
  for (int qi = 0; qi < QPT; ++qi) {
    const int i = base + qi * kSweepThreads + threadIdx.x;
    if (i >= t.n) break;
    if (t.corr_idx[i] < 0) continue;
    const float4 q = t.cur[i];
    const float4 m = __ldg(t.grid.pts + t.cand[i].x);
    const double sx = q.x, sy = q.y, sz = q.z, dx = m.x, dy = m.y, dz = m.z;
    acc[0] += 1.0;
    acc[1] += sx; acc[2] += sy; acc[3] += sz;
    acc[4] += dx; acc[5] += dy; acc[6] += dz;
    acc[7] += dx * sx; acc[8] += dx * sy; acc[9] += dx * sz;
    acc[10] += dy * sx; acc[11] += dy * sy; acc[12] += dy * sz;
    acc[13] += dz * sx; acc[14] += dz * sy; acc[15] += dz * sz;
    acc[16] += (double)t.corr_d2[i];
  }
  cta_reduce_sums(acc, smem);
  if (!grid_reduce_last(t, smem, ncta)) return;
  if (threadIdx.x == 0) {
    st->ticket = 0;
    p2p_finish_iteration(&smem[0][0], st, cfg);
  }
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
  NNResult r = grid_nn<kSweepThreads>(t.grid, q.x, q.y, q.z, INFINITY, max_rings, -1, sc);
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
