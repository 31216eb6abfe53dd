// grid.cuh — K1: build of the uniform neighbour grid over a target cloud.
//
// Replaces the FLANN k-d tree build that pcl::Registration::initCompute performs after
// icp.setInputTarget(prev_cloud_) (reference src/icpslam/icp_odometer.cpp:194,
// src/icpslam/octree_mapper.cpp:110).
//
// A counting sort in four passes, all streaming and coalesced (HBM-bound, 16 B/point loads):
//   bbox_kernel      min/max corner + non-finite flag               read 16 N
//   grid_count       cell id per point, per-cell histogram (atomics) read 16 N, write 8 N
//   scan_*           in-place exclusive prefix sum of the histogram  read 2 C, write C   (C cells)
//   grid_scatter     points to their sorted slot, .w = original idx  read 24 N, write 16 N
// Order inside a cell is whatever the atomics gave; every consumer breaks distance ties on the
// original index explicitly, so results do not depend on it.
#pragma once
#include "common.cuh"

namespace b2 {

struct BBox {
  int mn[3];  // order-preserving int encoding of the float min corner
  int mx[3];
  int nonfinite;
  int occupied;  // number of non-empty cells (filled by scan_tile_sums)
};

__device__ __forceinline__ int f2ord(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__host__ __device__ __forceinline__ float ord2f(int i) {
  int j = i >= 0 ? i : i ^ 0x7FFFFFFF;
#ifdef __CUDA_ARCH__
  return __int_as_float(j);
#else
  float f;
  memcpy(&f, &j, 4);
  return f;
#endif
}

__global__ void bbox_init(BBox* b) {
  if (threadIdx.x < 3) {
    b->mn[threadIdx.x] = 0x7FFFFFFF;
    b->mx[threadIdx.x] = (int)0x80000000;
  }
  if (threadIdx.x == 3) b->nonfinite = 0;
  if (threadIdx.x == 4) b->occupied = 0;
}

__global__ void __launch_bounds__(256) bbox_kernel(const float4* __restrict__ p, int n, BBox* __restrict__ out) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  int bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 v = __ldg(p + i);
    bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z));
    mn[0] = fminf(mn[0], v.x);
    mn[1] = fminf(mn[1], v.y);
    mn[2] = fminf(mn[2], v.z);
    mx[0] = fmaxf(mx[0], v.x);
    mx[1] = fmaxf(mx[1], v.y);
    mx[2] = fmaxf(mx[2], v.z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xFFFFFFFFu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xFFFFFFFFu, mx[d], o));
    }
    bad |= __shfl_xor_sync(0xFFFFFFFFu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (mn[d] <= mx[d]) {
        atomicMin(&out->mn[d], f2ord(mn[d]));
        atomicMax(&out->mx[d], f2ord(mx[d]));
      }
    }
    if (bad) atomicOr(&out->nonfinite, 1);
  }
}

// Cell coordinate of a coordinate value: the SAME expression bins target points and locates
// queries, so it is monotone in v and consistent between build and search.
__device__ __forceinline__ int cell_coord(float v, float o, float inv_cell, int n) {
  float f = floorf(fmul(fsub(v, o), inv_cell));
  int c = (int)fminf(fmaxf(f, 0.0f), (float)(n - 1));
  return c;
}

__global__ void __launch_bounds__(256) grid_count(const float4* __restrict__ p, int n, GridView g,
                                                  int* __restrict__ cell_of, int* __restrict__ rank,
                                                  int* __restrict__ count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v = __ldg(p + i);
  int cx = cell_coord(v.x, g.ox, g.inv_cell, g.nx);
  int cy = cell_coord(v.y, g.oy, g.inv_cell, g.ny);
  int cz = cell_coord(v.z, g.oz, g.inv_cell, g.nz);
  int c = (cz * g.ny + cy) * g.nx + cx;
  cell_of[i] = c;
  rank[i] = atomicAdd(count + c, 1);
}

__global__ void __launch_bounds__(256) grid_scatter(const float4* __restrict__ p, int n,
                                                    const int* __restrict__ cell_of, const int* __restrict__ rank,
                                                    const int* __restrict__ cell_start, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v = __ldg(p + i);
  int pos = __ldg(cell_start + __ldg(cell_of + i)) + __ldg(rank + i);
  out[pos] = make_float4(v.x, v.y, v.z, __int_as_float(i));
}

// ---- in-place exclusive scan over `n` ints (+ total written to data[n]) ------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096 ints per CTA

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int& total) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (kScanThreads / 32) ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < (kScanThreads / 32)) smem[lane] = winc - w;
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  total = smem[32];
  return inc - v + smem[warp];
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int* __restrict__ data, int n,
                                                               int* __restrict__ tile_sums, BBox* __restrict__ stats) {
  __shared__ int smem[33];
  int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int s = 0, occ = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k += 4) {
    int i = base + k;
    if (i + 3 < n) {
      int4 v = *reinterpret_cast<const int4*>(data + i);
      s += v.x + v.y + v.z + v.w;
      occ += (v.x > 0) + (v.y > 0) + (v.z > 0) + (v.w > 0);
    } else {
      for (int j = i; j < n && j < i + 4; ++j) {
        int v = data[j];
        s += v;
        occ += v > 0;
      }
    }
  }
  int total;
  block_exclusive_scan(s, smem, total);
  int occ_total;
  __syncthreads();
  block_exclusive_scan(occ, smem, occ_total);
  if (threadIdx.x == 0) {
    tile_sums[blockIdx.x] = total;
    if (occ_total) atomicAdd(&stats->occupied, occ_total);
  }
}

// single CTA: exclusive scan of the tile sums, in place
__global__ void __launch_bounds__(kScanThreads) scan_of_sums(int* __restrict__ tile_sums, int ntiles) {
  __shared__ int smem[33];
  int carry = 0;
  for (int base = 0; base < ntiles; base += kScanThreads) {
    int i = base + threadIdx.x;
    int v = i < ntiles ? tile_sums[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, smem, total);
    if (i < ntiles) tile_sums[i] = ex + carry;
    carry += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply(int* __restrict__ data, int n,
                                                           const int* __restrict__ tile_offsets, int grand_total) {
  __shared__ int smem[33];
  int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k += 4) {
    int i = base + k;
    if (i + 3 < n) {
      int4 t = *reinterpret_cast<const int4*>(data + i);
      v[k] = t.x;
      v[k + 1] = t.y;
      v[k + 2] = t.z;
      v[k + 3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[k + j] = (i + j < n) ? data[i + j] : 0;
    }
    s += v[k] + v[k + 1] + v[k + 2] + v[k + 3];
  }
  int total;
  int run = block_exclusive_scan(s, smem, total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int t = v[k];
    v[k] = run;
    run += t;
  }
#pragma unroll
  for (int k = 0; k < kScanItems; k += 4) {
    int i = base + k;
    if (i + 3 < n) {
      *reinterpret_cast<int4*>(data + i) = make_int4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i + j < n) data[i + j] = v[k + j];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) data[n] = grand_total;
}

}  // namespace b2

// ---- K8: pcl::VoxelGrid<PointXYZ>::applyFilter (reference src/icpslam/icp_odometer.cpp:96-101; SURVEY.md
// App. A.8) as a counting sort over the dense voxel table plus one centroid pass ---------------------------
namespace b2 {

struct VoxelParams {
  float inv_leaf;         // 1 / leaf (float division, like Eigen::Array4f::Ones() / leaf_size_)
  int min_b[3];           // floor(min_p * inv_leaf)
  int div[3];             // max_b - min_b + 1
};

// PCL: ijk = static_cast<int>(floor(p * inverse_leaf_size) - static_cast<float>(min_b));
//      idx = ijk0 + ijk1 * dx + ijk2 * dx * dy
__device__ __forceinline__ int voxel_index(float4 v, const VoxelParams& p) {
  const int i0 = (int)fsub(floorf(fmul(v.x, p.inv_leaf)), (float)p.min_b[0]);
  const int i1 = (int)fsub(floorf(fmul(v.y, p.inv_leaf)), (float)p.min_b[1]);
  const int i2 = (int)fsub(floorf(fmul(v.z, p.inv_leaf)), (float)p.min_b[2]);
  return i0 + i1 * p.div[0] + i2 * p.div[0] * p.div[1];
}

__global__ void __launch_bounds__(256) voxel_count(const float4* __restrict__ p, int n, VoxelParams vp,
                                                   int* __restrict__ cell_of, int* __restrict__ rank,
                                                   int* __restrict__ count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = voxel_index(__ldg(p + i), vp);
  cell_of[i] = c;
  rank[i] = atomicAdd(count + c, 1);
}

// leader flag per sorted slot: the first slot of every occupied voxel
__global__ void __launch_bounds__(256) voxel_flag_leaders(int n, const int* __restrict__ cell_of,
                                                          const int* __restrict__ rank,
                                                          const int* __restrict__ cell_start, int* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[__ldg(cell_start + __ldg(cell_of + i)) + __ldg(rank + i)] = (__ldg(rank + i) == 0) ? 1 : 0;
}

// One thread per occupied voxel (its leader slot): float centroid of the voxel's points taken in ORIGINAL
// index order (what a stable sort by voxel index gives), written at the voxel's rank among occupied
// voxels = ascending voxel index, PCL's output order.
__global__ void __launch_bounds__(128) voxel_centroids(const float4* __restrict__ sorted, int n, VoxelParams vp,
                                                       const int* __restrict__ cell_start,
                                                       const int* __restrict__ slot_of, float4* __restrict__ out) {
  int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  if (__ldg(slot_of + pos + 1) == __ldg(slot_of + pos)) return;  // not a leader
  const int c = voxel_index(__ldg(sorted + pos), vp);
  const int s = __ldg(cell_start + c), e = __ldg(cell_start + c + 1);
  float sx = 0.f, sy = 0.f, sz = 0.f;
  int last = -1;
  for (int k = s; k < e; ++k) {  // selection by increasing original index (voxels hold a handful of points)
    int best = 0x7FFFFFFF, bj = s;
    for (int j = s; j < e; ++j) {
      const int id = __float_as_int(__ldg(&sorted[j].w));
      if (id > last && id < best) {
        best = id;
        bj = j;
      }
    }
    const float4 p = __ldg(sorted + bj);
    sx = fadd(sx, p.x);
    sy = fadd(sy, p.y);
    sz = fadd(sz, p.z);
    last = best;
  }
  const float cnt = (float)(e - s);
  out[__ldg(slot_of + pos)] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), 1.0f);
}

}  // namespace b2
