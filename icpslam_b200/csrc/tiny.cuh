// tiny.cuh — the whole ICP loop of ONE small scan pair in ONE launch (BASELINE configs[2]: 1080-point planar scans
// arriving at 40 Hz, scan-to-scan; reference src/icpslam/icp_odometer.cpp:188-209).
//
// For a thin cloud the general path is all overhead: a neighbour grid per new target (two host round trips), one
// launch per iteration, a fitness pass with its own launches.  Here a small cooperative grid (four threads per query,
// 128-thread CTAs) keeps the target in shared memory and runs
//     q = T_inc * q   ->   exhaustive 1-NN over the staged target   ->   gate, 17 fp64 sums   ->   grid barrier   ->
//     Umeyama / SVD + DefaultConvergenceCriteria (solve.cuh; every CTA redundantly, on its own copy of the state)
// for as many iterations as PCL would, then getFitnessScore() on the final transform, and leaves the state, the
// running cloud and the correspondences of the last iteration in global memory.  No grid, no host round trip
// inside the scan.  The search is exhaustive in target-index order with a strict <, which IS the canonical rule
// (float d2 in FLANN's operation order, ties to the smallest index): exact by construction.
// Sums: lane butterfly, the CTA's warps in order, the CTAs in order — a fixed tree, bit-reproducible.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "solve.cuh"

namespace b2 {

constexpr int kTinyThreads = 128;
constexpr int kTinyTpq = 4;     // threads per query: each scans a quarter of the target, the four results are merged
constexpr int kTinyQpc = kTinyThreads / kTinyTpq;  // queries per CTA
constexpr int kTinyMax = 4096;  // largest source / target the single-launch path takes

struct TinyArgs {
  const float4* src;
  const float4* tgt;   // original order: the index of a point is its correspondence index
  float4* cur;         // [ns] running cloud after the last iteration
  int* corr_idx;       // [ns]
  float* corr_d2;      // [ns]
  double* partials;    // [2][ctas][kNumSums]
  IcpState* state;     // in: Tinc = final_T = guess, the rest zero; out: the finished loop
  IcpConfig cfg;
  int ns, nt;
  int with_fitness;
};

// The kTinyTpq lanes of a query scan interleaved quarters of the target (lane s: j = s, s + 4, ...) in ascending
// index with a strict <, then the packed (d2, index) keys are merged with a min: ties stay with the smallest index.
__device__ __forceinline__ void tiny_nn(const float4* __restrict__ s_tgt, int nt, int sub, float qx, float qy, float qz,
                                        float& best, int& bi) {
  float b = INFINITY;
  int k = 0x7FFFFFFF;
#pragma unroll 4
  for (int j = sub; j < nt; j += kTinyTpq) {
    const float4 t = s_tgt[j];
    const float d = sqdist3(qx, qy, qz, t.x, t.y, t.z);
    if (d < b) {
      b = d;
      k = j;
    }
  }
  unsigned long long key = pack_key(b, k);
#pragma unroll
  for (int o = 1; o < kTinyTpq; o <<= 1) {
    const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
    key = other < key ? other : key;
  }
  best = key_d2(key);
  bi = key_idx(key) == 0x7FFFFFFF ? -1 : key_idx(key);
}

__global__ void __launch_bounds__(kTinyThreads) icp_tiny_kernel(TinyArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char tiny_smem[];
  float4* s_tgt = reinterpret_cast<float4*>(tiny_smem);
  __shared__ IcpState s_st;
  __shared__ double s_warp[kTinyThreads / 32][kNumSums];
  __shared__ double s_sum[kNumSums];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = threadIdx.x & (kTinyTpq - 1);
  const int i = blockIdx.x * kTinyQpc + (threadIdx.x / kTinyTpq);
  const int G = gridDim.x;
  for (int j = threadIdx.x; j < a.nt; j += kTinyThreads) s_tgt[j] = __ldg(a.tgt + j);
  if (threadIdx.x == 0) s_st = *a.state;
  __syncthreads();
  const bool have = i < a.ns;
  float4 p = have ? __ldg(a.src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 p_src = p;
  int par = 0, last_idx = -1;
  float last_d2 = INFINITY;
  while (!s_st.done) {
    double acc[kNumSums];
#pragma unroll
    for (int c = 0; c < kNumSums; ++c) acc[c] = 0.0;
    {
      // (every lane runs the search — its shuffles are warp-wide; lanes past the cloud carry a dummy query)
      const float4 q = xform_f(s_st.Tinc, p.x, p.y, p.z);
      p = q;
      const bool finite = isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
      float d2;
      int bi;
      tiny_nn(s_tgt, a.nt, sub, q.x, q.y, q.z, d2, bi);
      if (have) {
        if (!finite) {
          atomicOr(&a.state->pad, 1);
          last_idx = -1;
          last_d2 = INFINITY;
        } else {
          last_d2 = d2;
          last_idx = (bi >= 0 && !((double)d2 > a.cfg.max2)) ? bi : -1;
          if (last_idx >= 0 && sub == 0) {  // one lane of the query carries its pair into the sums
            const float4 m = s_tgt[bi];
            const double sx = q.x, sy = q.y, sz = q.z, dx = m.x, dy = m.y, dz = m.z;
            acc[0] = 1.0;
            acc[1] = sx; acc[2] = sy; acc[3] = sz;
            acc[4] = dx; acc[5] = dy; acc[6] = dz;
            acc[7] = dx * sx; acc[8] = dx * sy; acc[9] = dx * sz;
            acc[10] = dy * sx; acc[11] = dy * sy; acc[12] = dy * sz;
            acc[13] = dz * sx; acc[14] = dz * sy; acc[15] = dz * sz;
            acc[16] = (double)d2;
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kNumSums; ++c) {
      const double s = warp_sum(acc[c]);
      if (lane == 0) s_warp[warp][c] = s;
    }
    __syncthreads();
    if (threadIdx.x < kNumSums) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kTinyThreads / 32; ++w) s += s_warp[w][threadIdx.x];
      a.partials[((size_t)par * G + blockIdx.x) * kNumSums + threadIdx.x] = s;
    }
    __threadfence();
    grid.sync();
    if (threadIdx.x < kNumSums) {
      double s = 0.0;
      for (int b = 0; b < G; ++b) s += __ldcg(a.partials + ((size_t)par * G + b) * kNumSums + threadIdx.x);
      s_sum[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      s_st.pad = *(volatile int*)&a.state->pad;
      p2p_finish_iteration(s_sum, &s_st, a.cfg);
    }
    __syncthreads();
    par ^= 1;
  }
  if (have && sub == 0) {
    a.cur[i] = make_float4(p.x, p.y, p.z, 0.0f);
    a.corr_idx[i] = last_idx;
    a.corr_d2[i] = last_d2;
  }
  // getFitnessScore(): final_T * source, exact unbounded 1-NN, mean of d2 (icp_odometer.cpp:201)
  if (a.with_fitness && s_st.status == 0) {
    double fs = 0.0, fc = 0.0;
    {
      const float4 q = xform_f(s_st.final_T, p_src.x, p_src.y, p_src.z);
      float d2;
      int bi;
      tiny_nn(s_tgt, a.nt, sub, q.x, q.y, q.z, d2, bi);
      if (have && bi >= 0 && sub == 0) {
        fs = (double)d2;
        fc = 1.0;
      }
    }
    fs = warp_sum(fs);
    fc = warp_sum(fc);
    if (lane == 0) {
      s_warp[warp][0] = fs;
      s_warp[warp][1] = fc;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      double s = 0.0;
      for (int w = 0; w < kTinyThreads / 32; ++w) s += s_warp[w][threadIdx.x];
      a.partials[((size_t)par * G + blockIdx.x) * kNumSums + threadIdx.x] = s;
    }
    __threadfence();
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      double ts = 0.0, tc = 0.0;
      for (int b = 0; b < G; ++b) {
        ts += __ldcg(a.partials + ((size_t)par * G + b) * kNumSums);
        tc += __ldcg(a.partials + ((size_t)par * G + b) * kNumSums + 1);
      }
      s_st.fitness_sum = ts;
      s_st.fitness_cnt = (unsigned long long)tc;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    s_st.pad = 0;
    *a.state = s_st;
  }
}

}  // namespace b2
