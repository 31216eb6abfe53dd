// map.cuh — K9: the mapper's point map as a device-resident voxel hash set.
//
// Replaces the pcl::octree::OctreePointCloudSearch member of OctreeMapper as the reference uses it
// (reference src/icpslam/octree_mapper.cpp:56-71):
//     for each scan point in order: if (!isVoxelOccupiedAtPoint(p)) addPointToCloud(p, map_cloud_);
// i.e. at most one point per `octree_resolution` voxel, first come wins, insertion order = scan order.
// Here a voxel is a cell of the global lattice floor(p / resolution) (PCL anchors its lattice on the first
// point's bounding box and regrows it, which only shifts which points share a voxel; DESIGN.md lists the
// deviation), and one insert call is three streaming passes over the n new points:
//   map_claim    key -> slot of an open-addressing table (atomicCAS on the key), atomicMin of the point's
//                index on the slot's value: the smallest index of the call wins the voxel; voxels filled by
//                earlier calls hold -1 and can never be won;
//   map_flag     winner flags (value == own index), then the exclusive scan of grid.cuh;
//   map_append   winners go to map[size + rank] in ascending index order (scan order) and commit the slot.
// Bytes per new point: 16 read x 3, 4 + 4 + 4 written, 16 written per winner; the table is 12 B per slot
// at load <= 0.5.  Deterministic: no result depends on the order of the atomics.
#pragma once
#include "common.cuh"

namespace b2 {

constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr int kCommitted = -1;

struct MapTable {
  unsigned long long* keys;  // [cap] voxel key, kEmptyKey = free
  int* vals;                 // [cap] smallest claiming index of the running insert, kCommitted = in the map
  unsigned int mask;         // cap - 1 (cap is a power of two)
  double inv_res;            // 1 / octree_resolution
};

// 21 bits per axis of floor(p / resolution), computed in double like the host restatement
__device__ __forceinline__ bool voxel_key(const float4& p, double inv_res, unsigned long long& key) {
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) return false;
  const long long ix = (long long)floor((double)p.x * inv_res);
  const long long iy = (long long)floor((double)p.y * inv_res);
  const long long iz = (long long)floor((double)p.z * inv_res);
  key = ((unsigned long long)(ix & 0x1FFFFF) << 42) | ((unsigned long long)(iy & 0x1FFFFF) << 21) |
        (unsigned long long)(iz & 0x1FFFFF);
  return true;
}

__device__ __forceinline__ unsigned int hash_key(unsigned long long k) {  // splitmix64 finaliser
  k ^= k >> 30;
  k *= 0xBF58476D1CE4E5B9ull;
  k ^= k >> 27;
  k *= 0x94D049BB133111EBull;
  k ^= k >> 31;
  return (unsigned int)k;
}

// slot of `key`, inserting it if absent
__device__ __forceinline__ unsigned int table_find_or_insert(const MapTable& t, unsigned long long key) {
  unsigned int slot = hash_key(key) & t.mask;
  for (;;) {
    unsigned long long k = t.keys[slot];
    if (k == kEmptyKey) k = atomicCAS(t.keys + slot, kEmptyKey, key);
    if (k == kEmptyKey || k == key) return slot;
    slot = (slot + 1) & t.mask;
  }
}

__global__ void __launch_bounds__(256) map_claim(const float4* __restrict__ pts, int n, MapTable t,
                                                 int* __restrict__ slot_of) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long key;
  if (!voxel_key(__ldg(pts + i), t.inv_res, key)) {
    slot_of[i] = -1;
    return;
  }
  const unsigned int slot = table_find_or_insert(t, key);
  slot_of[i] = (int)slot;
  atomicMin(t.vals + slot, i);  // a committed voxel holds -1 and stays -1
}

__global__ void __launch_bounds__(256) map_flag(int n, MapTable t, const int* __restrict__ slot_of,
                                                int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = slot_of[i];
  flags[i] = (s >= 0 && t.vals[s] == i) ? 1 : 0;
}

// flags[] now holds the exclusive scan (flags[n] = number of winners)
__global__ void __launch_bounds__(256) map_append(const float4* __restrict__ pts, int n, MapTable t,
                                                  const int* __restrict__ slot_of, const int* __restrict__ rank,
                                                  float4* __restrict__ map, int map_size) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = slot_of[i];
  if (s < 0 || rank[i + 1] == rank[i]) return;
  const float4 p = __ldg(pts + i);
  map[map_size + rank[i]] = make_float4(p.x, p.y, p.z, 1.0f);
  t.vals[s] = kCommitted;
}

// table growth: re-enter the voxels of the points already in the map
__global__ void __launch_bounds__(256) map_rehash(const float4* __restrict__ map, int map_size, MapTable t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= map_size) return;
  unsigned long long key;
  if (!voxel_key(map[i], t.inv_res, key)) return;
  t.vals[table_find_or_insert(t, key)] = kCommitted;
}

// nn_cloud of OctreeMapper::approxNearestNeighbors (octree_mapper.cpp:84-87): the map POINT of every query
// that has a neighbour, compacted in query order.  flags/rank as above.
__global__ void __launch_bounds__(256) map_gather_flag(const int* __restrict__ idx, int n, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = idx[i] >= 0 ? 1 : 0;
}
__global__ void __launch_bounds__(256) map_gather(const int* __restrict__ idx, int n, const int* __restrict__ rank,
                                                  const float4* __restrict__ map, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || idx[i] < 0) return;
  const float4 p = map[idx[i]];
  out[rank[i]] = make_float4(p.x, p.y, p.z, 1.0f);
}

}  // namespace b2
