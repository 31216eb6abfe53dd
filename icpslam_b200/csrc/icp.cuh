// icp.cuh — the kernels around the ICP loop: getFitnessScore and pcl::transformPointCloud.
// (The loop itself — one cooperative sweep over the batch's entry array plus one per-scan reduce / solve per
// iteration — is sweep.cuh; the neighbour search it uses is coop.cuh.)
#pragma once
#include "common.cuh"
#include "nn.cuh"
#include "nncache.cuh"
#include "solve.cuh"
#include "sweep.cuh"

namespace b2 {

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
    const int e = t.pos[i];  // the loop's per-query state lives in the batch's entry arrays
    const float4 c = t.cur[e];
    const float step = __fmul_ru(sqrt_fast(sqdist3(q.x, q.y, q.z, c.x, c.y, c.z)), kRelUp);
    const float L = __fsub_rd(c.w, step);
    const float4* const cand[2] = {t.c0, t.c1};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float4 ck = cand[k][e];
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
