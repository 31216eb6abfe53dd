/* octree_oracle.cpp — CPU restatement of pcl::octree::OctreePointCloudSearch<PointXYZ> as OctreeMapper uses it
 * (reference src/icpslam/octree_mapper.cpp:56-90).  TEST INFRASTRUCTURE ONLY (see b2icp_oracle.h).
 *
 *   resetMap                new octree at `resolution`, empty cloud                               (:56-60)
 *   addPointsToMap          per point, in order: if (!isVoxelOccupiedAtPoint(p)) addPointToCloud   (:63-71)
 *   approxNearestNeighbors  per query: approxNearestSearch, the MAP POINT is pushed to nn_cloud    (:73-90)
 *
 * PCL is absent from this image (parity unpinned); what is restated, from the published PCL 1.8 sources as
 * SURVEY.md App. A.7 summarises them:
 *   - the first point defines the bounding box p +- resolution/2, which getKeyBitSize() widens to a depth-1 tree
 *     (side 2 * resolution, the point at its centre);
 *   - a point outside the box adds a new root: per axis the old tree becomes the LOWER child iff that axis' upper
 *     bound is violated (otherwise the box grows towards minus), the side doubles, max = min + side - FLT_EPSILON;
 *   - key = (unsigned)((p - min) / resolution) per axis (double arithmetic); a leaf is one voxel of side resolution;
 *   - isVoxelOccupiedAtPoint: inside the box (bounds inclusive) and a leaf exists at the key;
 *   - approxNearestSearch: from the root, at every level among the EXISTING children the one whose voxel centre is
 *     nearest to the query (float squared distance, strict <, child order 0..7 with bit2 = x, bit1 = y, bit0 = z),
 *     down to a leaf, then the nearest point stored in that leaf.  No backtracking: not the exact NN.
 *     key_rule 0 restates PCL 1.8.x literally: the key handed to the next level is the one computed for the LAST
 *     existing child of the loop, not for the chosen child (the voxel centres of the deeper levels are then taken
 *     around that key); key_rule 1 hands down the chosen child's key (later PCL releases).
 */
#include "b2icp_oracle.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

namespace {

struct Node {
  std::unique_ptr<Node> child[8];
  std::vector<int> idx;  /* leaf: point indices */
};

struct Octree {
  double res = 0, mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  unsigned depth = 0;
  bool defined = false;
  size_t leaves = 0;
  std::unique_ptr<Node> root{new Node()};
  std::vector<float> pts; /* xyzw, insertion order */

  static double log2d(double v) { return std::log(v) / std::log(2.0); }

  void get_key_bit_size() {
    const float minValue = std::numeric_limits<float>::epsilon();
    unsigned mk[3];
    for (int d = 0; d < 3; ++d) mk[d] = (unsigned)std::ceil((mx[d] - mn[d] - minValue) / res);
    const unsigned max_voxels = std::max(std::max(std::max(mk[0], mk[1]), mk[2]), 2u);
    depth = std::max(std::min(32u, (unsigned)std::ceil(log2d((double)max_voxels) - minValue)), 0u);
    const double side = (double)(1u << depth) * res;
    if (leaves == 0) {
      for (int d = 0; d < 3; ++d) {
        const double over = (side - (mx[d] - mn[d])) / 2.0;
        if (over > minValue) {
          mn[d] -= over;
          mx[d] += over;
        }
      }
    } else {
      for (int d = 0; d < 3; ++d) mx[d] = mn[d] + side;
    }
  }

  void adopt_bounding_box(const float* p) {
    const float minValue = std::numeric_limits<float>::epsilon();
    for (;;) {
      bool lo[3], hi[3], any = false;
      for (int d = 0; d < 3; ++d) {
        lo[d] = (double)p[d] < mn[d];
        hi[d] = (double)p[d] >= mx[d];
        any = any || lo[d] || hi[d];
      }
      if (!any && defined) break;
      if (defined) {
        const unsigned child_idx = ((!hi[0]) << 2) | ((!hi[1]) << 1) | (!hi[2]);
        std::unique_ptr<Node> nr(new Node());
        nr->child[child_idx] = std::move(root);
        root = std::move(nr);
        double side = (double)(1u << depth) * res;
        for (int d = 0; d < 3; ++d)
          if (!hi[d]) mn[d] -= side;
        ++depth;
        side = (double)(1u << depth) * res - minValue;
        for (int d = 0; d < 3; ++d) mx[d] = mn[d] + side;
      } else {
        for (int d = 0; d < 3; ++d) {
          mn[d] = (double)p[d] - res / 2;
          mx[d] = (double)p[d] + res / 2;
        }
        get_key_bit_size();
        defined = true;
      }
    }
  }

  void key_of(const float* p, unsigned* k) const {
    for (int d = 0; d < 3; ++d) k[d] = (unsigned)(((double)p[d] - mn[d]) / res);
  }

  Node* find_leaf(const unsigned* k, bool create) {
    Node* n = root.get();
    for (unsigned level = depth; level-- > 0;) {
      const unsigned c = (((k[0] >> level) & 1u) << 2) | (((k[1] >> level) & 1u) << 1) | ((k[2] >> level) & 1u);
      if (!n->child[c]) {
        if (!create) return nullptr;
        n->child[c].reset(new Node());
        if (level == 0) ++leaves;
      }
      n = n->child[c].get();
    }
    return n;
  }

  bool voxel_occupied(const float* p) {
    if (!defined) return false;
    for (int d = 0; d < 3; ++d)
      if ((double)p[d] < mn[d] || (double)p[d] > mx[d]) return false;
    unsigned k[3];
    key_of(p, k);
    for (int d = 0; d < 3; ++d)
      if (k[d] >= (1u << depth)) return false; /* p == max bound: no leaf there */
    return find_leaf(k, false) != nullptr;
  }

  void add_point(const float* p) {
    const int id = (int)(pts.size() / 4);
    pts.insert(pts.end(), {p[0], p[1], p[2], 1.0f});
    adopt_bounding_box(p);
    unsigned k[3];
    key_of(p, k);
    find_leaf(k, true)->idx.push_back(id);
  }

  /* genVoxelCenterFromOctreeKey(key, tree_depth) -> float point */
  void center(const unsigned* k, unsigned tree_depth, float* c) const {
    const double w = res * (double)(1u << (depth - tree_depth));
    for (int d = 0; d < 3; ++d) c[d] = (float)(((double)k[d] + 0.5f) * w + mn[d]);
  }

  static float sq(const float* a, const float* b) {
    const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return dx * dx + dy * dy + dz * dz; /* Eigen squaredNorm of a Vector3f */
  }

  int approx_nearest(const float* q, int key_rule) const {
    if (leaves == 0) return -1;
    const Node* n = root.get();
    unsigned key[3] = {0, 0, 0};
    for (unsigned tree_depth = 1;; ++tree_depth) {
      double best = std::numeric_limits<double>::max();
      int bc = -1;
      unsigned nk[3] = {0, 0, 0}, bk[3] = {0, 0, 0};
      for (unsigned c = 0; c < 8; ++c) {
        if (!n->child[c]) continue;
        nk[0] = (key[0] << 1) + ((c >> 2) & 1u);
        nk[1] = (key[1] << 1) + ((c >> 1) & 1u);
        nk[2] = (key[2] << 1) + (c & 1u);
        float ctr[3];
        center(nk, tree_depth, ctr);
        const double d = (double)sq(ctr, q);
        if (d >= best) continue;
        best = d;
        bc = (int)c;
        std::memcpy(bk, nk, sizeof(bk));
      }
      if (bc < 0) return -1;
      n = n->child[bc].get();
      std::memcpy(key, key_rule == 0 ? nk : bk, sizeof(key));
      if (tree_depth >= depth) break;
    }
    double best = std::numeric_limits<double>::max();
    int result = -1;
    for (int id : n->idx) {
      const double d = (double)sq(&pts[4 * (size_t)id], q);
      if (d >= best) continue;
      best = d;
      result = id;
    }
    return result;
  }
};

}  // namespace

extern "C" {

void* b2o_octree_create(double resolution) {
  if (!(resolution > 0)) return nullptr;
  Octree* t = new Octree();
  t->res = resolution;
  return t;
}
void b2o_octree_free(void* tree) { delete static_cast<Octree*>(tree); }

/* OctreeMapper::addPointsToMap: returns the number of points added; non-finite points are skipped. */
size_t b2o_octree_add_points(void* tree, const float* xyzw, size_t n) {
  Octree* t = static_cast<Octree*>(tree);
  size_t added = 0;
  for (size_t i = 0; i < n; ++i) {
    const float* p = xyzw + 4 * i;
    if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
    if (!t->voxel_occupied(p)) {
      t->add_point(p);
      ++added;
    }
  }
  return added;
}
size_t b2o_octree_size(void* tree) { return static_cast<Octree*>(tree)->pts.size() / 4; }
void b2o_octree_points(void* tree, float* out_xyzw) {
  Octree* t = static_cast<Octree*>(tree);
  std::memcpy(out_xyzw, t->pts.data(), t->pts.size() * sizeof(float));
}
/* min corner (3 doubles), depth: the lattice the tree has grown into */
void b2o_octree_box(void* tree, double* min3, int* depth) {
  Octree* t = static_cast<Octree*>(tree);
  for (int d = 0; d < 3; ++d) min3[d] = t->mn[d];
  *depth = (int)t->depth;
}
/* OctreeMapper::approxNearestNeighbors: idx[i] = map index returned by approxNearestSearch (-1: non-finite query or
 * empty map). */
void b2o_octree_approx_nearest(void* tree, const float* q_xyzw, size_t n, int key_rule, int32_t* idx) {
  Octree* t = static_cast<Octree*>(tree);
  for (size_t i = 0; i < n; ++i) {
    const float* q = q_xyzw + 4 * i;
    idx[i] = (std::isfinite(q[0]) && std::isfinite(q[1]) && std::isfinite(q[2])) ? t->approx_nearest(q, key_rule) : -1;
  }
}
}
