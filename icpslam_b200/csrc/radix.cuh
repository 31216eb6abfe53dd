// radix.cuh — a stable LSD radix sort of (key, value) pairs, 11 bits per pass, written for DETERMINISM: equal keys keep
// their input order, so whatever is derived from the sorted order (the float centroid of a voxel, summed in original
// point order like a stable sort of PCL's would) is a pure function of the input.
// Per pass: sort_hist (per-chunk digit histogram) -> exclusive scan of the [digit][chunk] table (grid.cuh scan_*) ->
// sort_scatter (stable ranks: per-warp histograms + match_any).  Used by the sparse path of the voxel filter
// (pcl::VoxelGrid sorts point indices by voxel: reference src/icpslam/icp_odometer.cpp:96-101).
#pragma once
#include "common.cuh"
#include "grid.cuh"

namespace b2 {

constexpr int kSortBits = 11;
constexpr int kSortBins = 1 << kSortBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortPerWarp = 512;                        // consecutive elements owned by one warp
constexpr int kSortChunk = kSortWarps * kSortPerWarp;    // 4096 elements per CTA

__global__ void __launch_bounds__(kSortThreads) sort_hist(const unsigned int* __restrict__ keys, int n, int shift,
                                                          int nchunk, int* __restrict__ hist) {
  __shared__ int h[kSortBins];
  for (int b = threadIdx.x; b < kSortBins; b += kSortThreads) h[b] = 0;
  __syncthreads();
  const int base = blockIdx.x * kSortChunk;
  for (int k = threadIdx.x; k < kSortChunk; k += kSortThreads) {
    const int i = base + k;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & (kSortBins - 1)], 1);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kSortBins; b += kSortThreads) hist[(size_t)b * nchunk + blockIdx.x] = h[b];
}

// `scanned` = exclusive prefix sums of hist ([digit][chunk] order): the first output slot of (digit, chunk).
__global__ void __launch_bounds__(kSortThreads) sort_scatter(const unsigned int* __restrict__ keys,
                                                             const unsigned int* __restrict__ vals, int n, int shift,
                                                             int nchunk, const int* __restrict__ scanned,
                                                             unsigned int* __restrict__ keys_out,
                                                             unsigned int* __restrict__ vals_out) {
  extern __shared__ int wh[];  // [kSortWarps][kSortBins]: per-warp digit counts, then per-warp running output slots
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = threadIdx.x; b < kSortWarps * kSortBins; b += kSortThreads) wh[b] = 0;
  __syncthreads();
  const int wbase = blockIdx.x * kSortChunk + warp * kSortPerWarp;
  int* mine = wh + warp * kSortBins;
  for (int r = 0; r < kSortPerWarp; r += 32) {
    const int i = wbase + r + lane;
    if (i < n) atomicAdd(&mine[(keys[i] >> shift) & (kSortBins - 1)], 1);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kSortBins; b += kSortThreads) {
    int run = scanned[(size_t)b * nchunk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const int c = wh[w * kSortBins + b];
      wh[w * kSortBins + b] = run;
      run += c;
    }
  }
  __syncthreads();
  // every warp walks its elements in order; inside a round, lanes with the same digit take consecutive slots in
  // lane order (match_any), so equal digits keep their input order: the pass is stable
  for (int r = 0; r < kSortPerWarp; r += 32) {
    const int i = wbase + r + lane;
    const bool valid = i < n;
    const unsigned int key = valid ? keys[i] : 0u;
    const unsigned int digit = valid ? ((key >> shift) & (kSortBins - 1)) : (unsigned int)(kSortBins + lane);
    const unsigned int peers = __match_any_sync(0xFFFFFFFFu, digit);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const int leader = __ffs(peers) - 1;
    int slot = 0;
    if (valid && lane == leader) {
      slot = mine[digit];
      mine[digit] = slot + __popc(peers);
    }
    slot = __shfl_sync(0xFFFFFFFFu, slot, leader);
    __syncwarp();
    if (valid) {
      keys_out[slot + rank] = key;
      vals_out[slot + rank] = vals[i];
    }
  }
}


// ---- sparse voxel filter: leaders and centroids over the SORTED order -------------------------------------------
// key[i] = voxel index of sorted slot i, val[i] = original point index (ascending inside a voxel: the sort is stable)
__global__ void __launch_bounds__(256) voxel_sparse_keys(const float4* __restrict__ p, int n, VoxelParams vp,
                                                         unsigned int* __restrict__ keys, unsigned int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = (unsigned int)voxel_index(__ldg(p + i), vp);
  vals[i] = (unsigned int)i;
}
__global__ void __launch_bounds__(256) voxel_sparse_flags(const unsigned int* __restrict__ keys, int n, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// flags[] holds the exclusive scan (flags[n] = number of voxels): one thread per leader sums its run in order
__global__ void __launch_bounds__(128) voxel_sparse_centroids(const float4* __restrict__ p, const unsigned int* __restrict__ keys,
                                                              const unsigned int* __restrict__ vals, int n,
                                                              const int* __restrict__ slot_of, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (slot_of[i + 1] == slot_of[i]) return;  // not a leader
  const unsigned int key = keys[i];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  int cnt = 0;
  for (int j = i; j < n && keys[j] == key; ++j) {
    const float4 q = __ldg(p + vals[j]);
    sx = fadd(sx, q.x);
    sy = fadd(sy, q.y);
    sz = fadd(sz, q.z);
    ++cnt;
  }
  const float c = (float)cnt;
  out[slot_of[i]] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), 1.0f);
}

}  // namespace b2
