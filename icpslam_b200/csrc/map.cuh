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
  // PCL-compatible lattice (b2icp_map_reset_octree): voxel = floor((p - org) / res), org = first point - res, the
  // corner PCL's octree anchors its leaves on; compat == 0: the global lattice floor(p / res)
  int compat;
  double res, org[3];
};

// 21 bits per axis of the voxel coordinates, computed in double like the host restatement
__device__ __forceinline__ bool voxel_key(const float4& p, const MapTable& t, unsigned long long& key) {
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) return false;
  long long ix, iy, iz;
  if (t.compat) {  // PCL: key = (p - min) / resolution (a division)
    ix = (long long)floor(__ddiv_rn(__dsub_rn((double)p.x, t.org[0]), t.res));
    iy = (long long)floor(__ddiv_rn(__dsub_rn((double)p.y, t.org[1]), t.res));
    iz = (long long)floor(__ddiv_rn(__dsub_rn((double)p.z, t.org[2]), t.res));
  } else {
    ix = (long long)floor((double)p.x * t.inv_res);
    iy = (long long)floor((double)p.y * t.inv_res);
    iz = (long long)floor((double)p.z * t.inv_res);
  }
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
  if (!voxel_key(__ldg(pts + i), t, key)) {
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
  if (!voxel_key(map[i], t, key)) return;
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

// ---- PCL-compatible mode: the octree OctreeMapper really searches (reference src/icpslam/octree_mapper.cpp:73-90) --
// pcl::octree::OctreePointCloudSearch::approxNearestSearch is NOT the nearest neighbour: from the root it follows, at
// every level, the EXISTING child whose voxel centre is nearest to the query, and returns the point stored in the leaf
// it ends in.  Which children exist at which level depends on how the root box grew, so the box
// is tracked exactly as PCL grows it (host, b2icp.cu) and the tree is kept as a hash set of (level, key) prefixes:
//   octree_first_outside   smallest index >= cursor of a finite point outside the current box: the next growth event
//   octree_build           every map point enters its leaf key and all its prefixes (leaf value = point index)
//   octree_approx_nearest  the greedy descent, one thread per query: 8 table probes per level

struct OctreeBox {
  double mn[3], mx[3];  // current root box (PCL's min_x_ .. max_z_)
  double res;
  int depth;            // octree_depth_
  int defined;
};

__global__ void __launch_bounds__(256) octree_first_outside(const float4* __restrict__ pts, int n, int cursor, OctreeBox b,
                                                            int* __restrict__ first) {
  for (int i = cursor + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(pts + i);
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;
    const bool out = !b.defined || (double)p.x < b.mn[0] || (double)p.x >= b.mx[0] || (double)p.y < b.mn[1] ||
                     (double)p.y >= b.mx[1] || (double)p.z < b.mn[2] || (double)p.z >= b.mx[2];
    if (out) atomicMin(first, i);
  }
}

__device__ __forceinline__ unsigned long long tree_key(int level, unsigned int kx, unsigned int ky, unsigned int kz) {
  return ((unsigned long long)level << 57) | ((unsigned long long)(kx & 0x7FFFF) << 38) | ((unsigned long long)(ky & 0x7FFFF) << 19) |
         (unsigned long long)(kz & 0x7FFFF);
}

// genOctreeKeyforPoint: key = (unsigned)((p - min) / resolution) per axis, double arithmetic
__device__ __forceinline__ void octree_leaf_key(const float4& p, const OctreeBox& b, unsigned int* k) {
  k[0] = (unsigned int)__ddiv_rn(__dsub_rn((double)p.x, b.mn[0]), b.res);
  k[1] = (unsigned int)__ddiv_rn(__dsub_rn((double)p.y, b.mn[1]), b.res);
  k[2] = (unsigned int)__ddiv_rn(__dsub_rn((double)p.z, b.mn[2]), b.res);
}

__global__ void __launch_bounds__(256) octree_build(const float4* __restrict__ map, int n, OctreeBox b, MapTable t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned int k[3];
  octree_leaf_key(map[i], b, k);
  for (int level = 1; level <= b.depth; ++level) {
    const int sh = b.depth - level;
    const unsigned int slot = table_find_or_insert(t, tree_key(level, k[0] >> sh, k[1] >> sh, k[2] >> sh));
    if (level == b.depth) atomicMin(t.vals + slot, i);  // a leaf holds one point (first come wins: the smallest index)
  }
}

__device__ __forceinline__ int table_lookup(const MapTable& t, unsigned long long key) {
  unsigned int slot = hash_key(key) & t.mask;
  for (;;) {
    const unsigned long long k = t.keys[slot];
    if (k == key) return (int)slot;
    if (k == kEmptyKey) return -1;
    slot = (slot + 1) & t.mask;
  }
}

__global__ void __launch_bounds__(256) octree_approx_nearest(const float4* __restrict__ q, int n, OctreeBox b, MapTable t,
                                                             int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(q + i);
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) || b.depth < 1) {
    idx[i] = -1;
    return;
  }
  unsigned int key[3] = {0u, 0u, 0u};
  int leaf_slot = -1;
  for (int tree_depth = 1; tree_depth <= b.depth; ++tree_depth) {
    // voxel side at this level; centre = (key + 0.5f) * side + min, in double, stored as float (genVoxelCenterFromOctreeKey)
    const double side = __dmul_rn(b.res, (double)(1u << (b.depth - tree_depth)));
    double best = 1.7976931348623157e308;
    unsigned int bk[3] = {0u, 0u, 0u};
    int bslot = -1;
    for (unsigned int c = 0; c < 8; ++c) {
      const unsigned int nk[3] = {(key[0] << 1) + ((c >> 2) & 1u), (key[1] << 1) + ((c >> 1) & 1u), (key[2] << 1) + (c & 1u)};
      const int slot = table_lookup(t, tree_key(tree_depth, nk[0], nk[1], nk[2]));
      if (slot < 0) continue;
      const float cx = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)nk[0], 0.5), side), b.mn[0]);
      const float cy = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)nk[1], 0.5), side), b.mn[1]);
      const float cz = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)nk[2], 0.5), side), b.mn[2]);
      const float dx = fsub(cx, p.x), dy = fsub(cy, p.y), dz = fsub(cz, p.z);
      const double d = (double)fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
      if (d >= best) continue;
      best = d;
      bslot = slot;
      bk[0] = nk[0]; bk[1] = nk[1]; bk[2] = nk[2];
    }
    if (bslot < 0) {  // cannot happen in a consistent tree
      idx[i] = -1;
      return;
    }
    key[0] = bk[0]; key[1] = bk[1]; key[2] = bk[2];
    leaf_slot = bslot;
  }
  idx[i] = t.vals[leaf_slot];
}

}  // namespace b2
