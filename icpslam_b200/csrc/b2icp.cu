// b2icp.cu — the C ABI of libb2icp.so (include/b2icp.h): handle, device memory, launch sequences.
//
// Drop-in boundary for the PCL registration object the reference builds on the stack
// (reference src/icpslam/icp_odometer.cpp:188-201, src/icpslam/octree_mapper.cpp:104-117).
// No C++ exception crosses this file's extern "C" functions; there is no CPU fallback.
//
// Execution model: a handle owns one CUDA stream, a pool of SCAN SLOTS (source cloud + running cloud
// + correspondence buffers + per-CTA partial sums) and a pool of GRID SLOTS (target cloud + its
// neighbour grid).  The single-scan entry points use slot 0 / grid 0; b2icp_align_batch fills up to
// kMaxBatch slots and advances all of them with ONE launch of the fused sweep kernel per ICP
// iteration (blockIdx.y = slot), so that a 148-SM part is filled by independent scans.
#include "../../include/b2icp.h"

#include <cuda_runtime.h>
#include "fiber.h"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "gicp.cuh"
#include "grid.cuh"
#include "icp.cuh"
#include "map.cuh"
#include "nn.cuh"
#include "nnsearch.cuh"
#include "radix.cuh"
#include "solve.cuh"
#include "tiny.cuh"

using namespace b2;

namespace {

constexpr int kMaxCells = 1 << 25;        // dense cell table cap (128 MiB of int32)
constexpr int kUnboundedRings = 3;        // ring budget of unbounded searches before the brute-force fallback
// Cell sizing, measured on the bench workload (cell edge vs scans/s: 0.2 m 5655, 0.25 m 6211, 0.3 m 6449,
// 0.35 m 6337, 0.4 m 6084, 0.5 m 5623): the optimum sits near 2.3 points per occupied cell.
constexpr double kTargetOccupancy = 2.5;  // points per occupied cell the auto-sizing aims at
constexpr double kMaxOccupancy = 4.0;     // above this the grid is rebuilt with smaller cells
constexpr float kCacheMarginFrac = 0.05f;  // nncache.cuh: box-search margin as a fraction of the cell edge
constexpr int kStreamedQpt = 32;          // slab length of streamed batches (4: 15.5k, 8: 18.6k, 16: 21.7k, 32: 23.2k scans/s)
constexpr int kFirstSweepQpt = 2;         // slab length (x 32 queries per warp) of the first sweep of a batch
constexpr int kMaxBatch = 64;             // scans advanced together by one sweep launch
constexpr int kGicpGroups = 4;            // groups a GICP batch can be dealt to (device rounds overlapping host updates)
constexpr int kStreamSets = 8;            // streamed batches in flight (b2icp_align_batch_submit / _wait)
constexpr int kSetSlots = 32;             // scans per streamed batch
constexpr int kSlots = kStreamSets * kSetSlots > kMaxBatch ? kStreamSets * kSetSlots : kMaxBatch;  // state / task records

struct DeviceBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

struct Cloud {
  DeviceBuf raw;  // float4[n] as uploaded
  const float4* ext = nullptr;  // a caller's device cloud used in place (streamed device batches); raw is unused then
  size_t n = 0;
  bool valid = false;
  const float4* dev() const { return ext ? ext : raw.as<float4>(); }
};

struct GridSlot {
  Cloud tgt;                     // owned copy of the target (unused when the grid borrows a slot's source)
  const float4* pts = nullptr;   // the cloud the grid indexes
  DeviceBuf sorted, cell_start, cell_of, rank, tile_sums, bbox;
  DeviceBuf cov;                 // GICP: 9 doubles per point (original order), valid while the grid is
  bool cov_valid = false;
  GridView view;
  double cell = 0, min_cell = 0;
  bool lazy = false;      // the target is set but its grid is not built yet (small targets: the single-launch path of
                          // b2icp_align never needs one; ensure_grid() builds it for anything else)
  double force_cell = 0;  // > 0: the cell edge is given (coarse second-pass grids of the GICP covariances), no refinement
  float mn[3], mx[3];
  double occupancy = 0;
  bool valid = false;
  // what the last build of this slot settled on: a similar cloud (the next sweep of a stream) starts from that cell
  // instead of the density estimate and usually needs no refinement pass
  double hint_cell = 0, hint_ext[3] = {0, 0, 0};
  size_t hint_n = 0;
  bool bbox_known = false;  // mn / mx are already set by the caller (a second grid over the same cloud)
  void release() {
    for (DeviceBuf* b : {&tgt.raw, &sorted, &cell_start, &cell_of, &rank, &tile_sums, &bbox, &cov}) b->release();
  }
};

struct ScanSlot {
  Cloud src;
  DeviceBuf cur, corr_idx, corr_d2, corr_pos, c0, c1, c2, partials;
  DeviceBuf cov, mahal, gicp_partials;  // GICP: source covariances, Mahalanobis matrices, per-CTA sums
  int grid = 0;
  void release() {
    for (DeviceBuf* b : {&src.raw, &cur, &corr_idx, &corr_d2, &corr_pos, &c0, &c1, &c2, &partials, &cov, &mahal, &gicp_partials})
      b->release();
  }
};

__global__ void zero_counter(unsigned int* c) { *c = 0; }

// pcl::fromROSMsg for PointXYZ: the three float fields of every point of a PointCloud2 payload (byte-wise loads: the
// offsets of a message need not be 4-byte aligned)
__global__ void __launch_bounds__(256) unpack_pointcloud2(const unsigned char* __restrict__ data, unsigned int width, unsigned int height,
                                                          unsigned int point_step, unsigned int row_step, unsigned int off_x,
                                                          unsigned int off_y, unsigned int off_z, int swap, float4* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)width * height) return;
  const size_t row = i / width, col = i - row * width;
  const unsigned char* p = data + row * row_step + col * point_step;
  auto field = [&](unsigned int off) {
    const unsigned char* b = p + off;
    const unsigned int v = swap ? ((unsigned int)b[0] << 24) | ((unsigned int)b[1] << 16) | ((unsigned int)b[2] << 8) | b[3]
                                : ((unsigned int)b[3] << 24) | ((unsigned int)b[2] << 16) | ((unsigned int)b[1] << 8) | b[0];
    return __uint_as_float(v);
  };
  out[i] = make_float4(field(off_x), field(off_y), field(off_z), 1.0f);
}

// IcpState[B] -> b2icp_record[B] (the record sink of streamed batches: b2icp_set_record_sink)
__global__ void export_records(const IcpState* __restrict__ st, int B, int with_fitness, b2icp_record* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const IcpState& s = st[i];
  b2icp_record r;
  for (int k = 0; k < 16; ++k) r.T[k] = s.final_T[k];
  r.converged = s.converged;
  r.iterations = s.iter;
  r.n_corr_last = s.n_corr;
  r.status_detail = s.status;
  r.mse_last = s.mse;
  r.fitness = (with_fitness && s.status == 0) ? (s.fitness_cnt > 0 ? s.fitness_sum / (double)s.fitness_cnt : 1.7976931348623157e308)
                                              : __longlong_as_double(0x7FF8000000000000ll);
  out[i] = r;
}


}  // namespace

// A captured ICP loop (task / state upload, one sweep per iteration, write-out) and the shape it was captured for.
struct GraphKey {
  int B, iters, qpt, streamed, sched;
  size_t max_n;
  IcpConfig cfg;
};
struct GraphCache {
  cudaGraphExec_t exec = nullptr;
  GraphKey exec_key, last_key;
  bool has_last = false;
};

struct b2icp_handle {
  b2icp_params params;
  IcpConfig cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::mutex mu;
  std::string err;

  std::vector<std::unique_ptr<ScanSlot>> slots;
  std::vector<std::unique_ptr<GridSlot>> grids;
  DeviceBuf states, tasks, unres_list, unres_count, unres_keys;
  // K9: the mapper's point map (map.cuh)
  DeviceBuf map_pts, map_keys, map_vals, map_slot_of, map_flags, map_tiles, map_stats;
  size_t map_size = 0, map_table_cap = 0;
  double map_resolution = 0.0;
  bool map_grid_valid = false;
  // PCL-compatible map mode (b2icp_map_reset_octree): the root box as PCL grows it, the (level, key) prefix table
  bool map_compat = false;
  OctreeBox map_box{};
  double map_org[3] = {0, 0, 0};
  DeviceBuf tree_keys, tree_vals, map_first;
  size_t tree_cap = 0;
  bool tree_valid = false;
  int* h_first = nullptr;     // pinned
  float* h_point = nullptr;   // pinned, one point
  DeviceBuf query, q_idx, q_d2, xf_in, xf_out, mat;
  IcpState* h_states = nullptr;  // pinned [kSlots]: upload (init) and read-back
  ScanTask* h_tasks = nullptr;   // pinned [kSlots]
  BBox* h_bbox = nullptr;        // pinned [kMaxBatch + 1]
  bool aligned = false;          // slot 0 holds a completed align
  int last_batch = 0;

  std::vector<cudaEvent_t> events;
  b2icp_timing timing;
  long long launches = 0;
  // streaming batches (b2icp_align_batch_submit / _wait): kStreamSets sets of kSetSlots slots
  struct Pending {
    int set, B, with_fitness;
  };
  cudaStream_t copy_stream = nullptr;
  cudaStream_t set_stream[kStreamSets] = {};  // one compute stream per slot set: the sparse late sweeps of one
                                              // batch share the machine with the first sweeps of the next ones
  cudaEvent_t set_uploaded[kStreamSets] = {}, set_done[kStreamSets] = {};
  Pending pending[kStreamSets] = {};
  int n_pending = 0, first_pending = 0, next_set = 0;
  int max_in_flight = kStreamSets;  // B2ICP_IN_FLIGHT environment variable (tuning only)
  int qpt_override = 0;  // B2ICP_QPT environment variable (tuning only)
  std::vector<int> qpt_sched;  // B2ICP_QPT_SCHED="2,4,8": slab length per iteration, last value repeats (tuning only)
  bool use_tiny = true;     // B2ICP_NO_TINY switches the single-launch path of small scan pairs off (tuning / debugging)
  bool tiny_last = false;   // the last b2icp_align ran on it (its getFitnessScore is already in the state)
  DeviceBuf tiny_partials;
  GraphCache graphs[kStreamSets + 1];  // one per streamed slot set, the last one for synchronous calls
  cudaStream_t capture_stream = nullptr;
  bool use_graphs = true;  // B2ICP_NO_GRAPH switches the captured loops off (tuning / debugging only)
  long long graph_launches = 0;
  b2icp_record* sink = nullptr;  // b2icp_set_record_sink: device records of streamed batches
  size_t sink_cap = 0, sink_used = 0;
  int w_override = 0;  // B2ICP_W: lanes per cooperative group of the stand-alone search, 8 or 32 (tuning only)
  int nn_sort_override = -1;  // B2ICP_NN_SORT=0/1 (tuning only)
  DeviceBuf qs_sorted, qs_cell_of, qs_rank, qs_count, qs_tiles;  // queries of the stand-alone search, sorted by target cell
  int join_d = 4;      // B2ICP_JOIN: cells of slack inside which a lane joins its group's pass (tuning only)
  double* h_gicp_partials = nullptr;  // pinned read-back of the GICP rounds: 14 sums per scan
  size_t h_gicp_partials_cap = 0;
  void* h_gicp_tasks = nullptr;  // pinned task arrays of a GICP round (gicp_host.inl)
  size_t h_gicp_tasks_cap = 0;
  DeviceBuf gicp_tasks, gicp_tickets, knn_tasks, knn_list2, knn_counts;
  long gicp_evals = 0, gicp_rounds = 0;
  int gicp_groups = 4;  // B2ICP_GICP_GROUPS (tuning only)
  int fitness_rings = kUnboundedRings;  // B2ICP_FITNESS_RINGS (tuning only): ring budget of getFitnessScore's search
  cudaStream_t gicp_streams[kGicpGroups] = {nullptr, nullptr, nullptr, nullptr};
};

namespace {

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                               \
      return B2ICP_ERR_CUDA;                                                                      \
    }                                                                                             \
  } while (0)

int fail(b2icp_handle* h, int code, const char* msg) {
  if (h) h->err = msg;
  return code;
}

ScanSlot& slot(b2icp_handle* h, size_t i) {
  while (h->slots.size() <= i) h->slots.emplace_back(new ScanSlot());
  return *h->slots[i];
}
GridSlot& gslot(b2icp_handle* h, size_t i) {
  while (h->grids.size() <= i) h->grids.emplace_back(new GridSlot());
  return *h->grids[i];
}

void derive_config(b2icp_handle* h) {
  const b2icp_params& p = h->params;
  IcpConfig& c = h->cfg;
  c.max2 = p.max_correspondence_distance * p.max_correspondence_distance;
  c.rot_thresh = 1.0 - p.transformation_epsilon;
  c.trans_thresh = p.transformation_epsilon;
  c.mse_abs = 1e-12;
  c.mse_rel = p.euclidean_fitness_epsilon;
  float b = (float)c.max2;
  if ((double)b < c.max2) b = nextafterf(b, INFINITY);
  if (!(c.max2 < (double)FLT_MAX)) b = INFINITY;
  c.bound2 = b;
  c.max_iterations = p.max_iterations;
  c.min_corr = 3;
  c.max_rings = 1;
  c.margin_frac = kCacheMarginFrac;
  if (const char* e = getenv("B2ICP_MARGIN")) c.margin_frac = (float)atof(e);  // tuning only
}

// The fused sweep, one instantiation per slab length.
template <int QPT>
void launch_sweep_one(dim3 grid, cudaStream_t st, const ScanTask* tasks, const IcpConfig& cfg) {
  icp_sweep_p2p<QPT><<<grid, kSweepThreads, 0, st>>>(tasks, cfg);
}
void launch_sweep(int q, dim3 grid, cudaStream_t st, const ScanTask* tasks, const IcpConfig& cfg) {
  switch (q) {
    case 32: launch_sweep_one<32>(grid, st, tasks, cfg); break;
    case 16: launch_sweep_one<16>(grid, st, tasks, cfg); break;
    case 8: launch_sweep_one<8>(grid, st, tasks, cfg); break;
    case 4: launch_sweep_one<4>(grid, st, tasks, cfg); break;
    case 2: launch_sweep_one<2>(grid, st, tasks, cfg); break;
    default: launch_sweep_one<1>(grid, st, tasks, cfg); break;
  }
}

int ensure_grid(b2icp_handle* h, GridSlot& g);

int rings_for_bound(const b2icp_handle* h, double min_cell) {
  if (!std::isfinite(h->cfg.bound2)) return kUnboundedRings;
  double r = std::sqrt((double)h->cfg.bound2);
  double k = std::ceil(r / min_cell) + 2.0;
  return k > 1e6 ? 1000000 : (int)k;
}

// ---- K1: neighbour-grid build, split in phases so that a batch needs two host syncs in total ------
// phase A: bounding boxes of `count` clouds -> h->h_bbox[0..count)
int grids_bbox(b2icp_handle* h, GridSlot* const* g, const size_t* n, int count) {
  bool any = false;
  for (int i = 0; i < count; ++i) {
    CK(g[i]->bbox.ensure(sizeof(BBox)));
    if (g[i]->bbox_known) continue;
    any = true;
    const int blocks = (int)((n[i] + 255) / 256);
    bbox_init<<<1, 32, 0, h->stream>>>(g[i]->bbox.as<BBox>());
    bbox_kernel<<<std::min(blocks, 148 * 8), 256, 0, h->stream>>>(g[i]->pts, (int)n[i], g[i]->bbox.as<BBox>());
    h->launches += 2;
    CK(cudaMemcpyAsync(&h->h_bbox[i], g[i]->bbox.p, sizeof(BBox), cudaMemcpyDeviceToHost, h->stream));
  }
  if (any) CK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < count; ++i) {
    double ext[3];
    if (!g[i]->bbox_known) {
      if (h->h_bbox[i].nonfinite) return fail(h, B2ICP_ERR_NONFINITE_INPUT, "target cloud holds non-finite coordinates");
      for (int d = 0; d < 3; ++d) {
        g[i]->mn[d] = ord2f(h->h_bbox[i].mn[d]);
        g[i]->mx[d] = ord2f(h->h_bbox[i].mx[d]);
      }
    }
    for (int d = 0; d < 3; ++d) ext[d] = (double)g[i]->mx[d] - (double)g[i]->mn[d];
    const double emax = std::max(ext[0], std::max(ext[1], ext[2]));
    // density-based cell: ~kTargetOccupancy points per cell if the cloud filled its box uniformly;
    // axes thinner than 1e-6 of the largest extent (planar scans) are left out of the estimate
    double vol = 1.0;
    int dims = 0;
    for (int d = 0; d < 3; ++d)
      if (ext[d] > 1e-6 * emax && ext[d] > 0) {
        vol *= ext[d];
        ++dims;
      }
    double cell = 1.0;
    if (dims > 0) cell = std::pow(vol * kTargetOccupancy / (double)n[i], 1.0 / dims);
    const double r = h->params.max_correspondence_distance;
    const bool bounded = r > 0 && std::isfinite(r) && r < 1e8;
    if (bounded) cell = std::min(cell, 0.5 * r);
    if (g[i]->hint_cell > 0 && (double)n[i] > 0.8 * (double)g[i]->hint_n && (double)n[i] < 1.25 * (double)g[i]->hint_n) {
      bool similar = true;  // same kind of cloud as the one this slot indexed last (cell size is a speed matter only)
      for (int d = 0; d < 3; ++d) similar = similar && std::fabs(ext[d] - g[i]->hint_ext[d]) <= 0.25 * std::max(emax, 1e-9);
      if (similar) cell = bounded ? std::min(g[i]->hint_cell, 0.5 * r) : g[i]->hint_cell;
    }
    if (h->params.grid_cell > 0) cell = h->params.grid_cell;
    if (g[i]->force_cell > 0) cell = g[i]->force_cell;
    if (!(cell > 0) || !std::isfinite(cell)) cell = 1.0;
    g[i]->cell = cell;
    g[i]->min_cell = (h->params.grid_cell > 0 || g[i]->force_cell > 0) ? cell : (bounded ? std::min(cell, r / 8.0) : cell / 8.0);
  }
  return B2ICP_OK;
}

// phase B: enqueue the counting sort of one cloud with the cell size chosen in g->cell
int grid_enqueue_build(b2icp_handle* h, GridSlot* g, size_t n_) {
  const int n = (int)n_;
  double ext[3];
  for (int d = 0; d < 3; ++d) ext[d] = (double)g->mx[d] - (double)g->mn[d];
  const double emax = std::max(ext[0], std::max(ext[1], ext[2]));
  long long nx, ny, nz;
  for (;;) {  // respect the dense-table cap
    nx = (long long)std::floor(ext[0] / g->cell) + 1;
    ny = (long long)std::floor(ext[1] / g->cell) + 1;
    nz = (long long)std::floor(ext[2] / g->cell) + 1;
    if ((double)nx * (double)ny * (double)nz <= (double)kMaxCells) break;
    g->cell *= 1.26;
  }
  GridView& v = g->view;
  v.ox = g->mn[0];
  v.oy = g->mn[1];
  v.oz = g->mn[2];
  v.cell = (float)g->cell;
  v.inv_cell = 1.0f / v.cell;
  v.nx = (int)nx;
  v.ny = (int)ny;
  v.nz = (int)nz;
  v.n = n;
  float amax = 0.f;
  for (int d = 0; d < 3; ++d) amax = std::max(amax, std::max(std::fabs(g->mn[d]), std::fabs(g->mx[d]) + v.cell));
  v.slack = std::max(amax, (float)emax) * 9.5367431640625e-7f + 1e-30f;  // 2^-20 of the coordinate range
  const int ncell = v.nx * v.ny * v.nz;
  CK(g->cell_start.ensure((size_t)(ncell + 1 + 4) * sizeof(int)));
  CK(g->sorted.ensure((size_t)n * sizeof(float4)));
  CK(g->cell_of.ensure((size_t)n * sizeof(int)));
  CK(g->rank.ensure((size_t)n * sizeof(int)));
  const int ntiles = (ncell + kScanTile - 1) / kScanTile;
  CK(g->tile_sums.ensure((size_t)ntiles * sizeof(int)));
  int* cs = g->cell_start.as<int>();
  const int blocks = (n + 255) / 256;
  CK(cudaMemsetAsync(cs, 0, (size_t)(ncell + 1) * sizeof(int), h->stream));
  bbox_init<<<1, 32, 0, h->stream>>>(g->bbox.as<BBox>());
  grid_count<<<blocks, 256, 0, h->stream>>>(g->pts, n, v, g->cell_of.as<int>(), g->rank.as<int>(), cs);
  scan_tile_sums<<<ntiles, kScanThreads, 0, h->stream>>>(cs, ncell, g->tile_sums.as<int>(), g->bbox.as<BBox>());
  scan_of_sums<<<1, kScanThreads, 0, h->stream>>>(g->tile_sums.as<int>(), ntiles);
  scan_apply<<<ntiles, kScanThreads, 0, h->stream>>>(cs, ncell, g->tile_sums.as<int>(), n);
  grid_scatter<<<blocks, 256, 0, h->stream>>>(g->pts, n, g->cell_of.as<int>(), g->rank.as<int>(), cs,
                                               g->sorted.as<float4>());
  h->launches += 6;
  v.pts = g->sorted.as<float4>();
  v.cell_start = cs;
  return B2ICP_OK;
}

// Build `count` grids (bbox pass, sort, occupancy check with at most two refinements).
int build_grids(b2icp_handle* h, GridSlot* const* g, const size_t* n, int count) {
  for (int i = 0; i < count; ++i) {
    g[i]->valid = false;
    g[i]->cov_valid = false;
  }
  int rc = grids_bbox(h, g, n, count);
  if (rc) return rc;
  std::vector<int> todo(count);
  for (int i = 0; i < count; ++i) todo[i] = i;
  for (int attempt = 0; attempt < 3 && !todo.empty(); ++attempt) {
    bool refinable = false;  // a given cell edge is final: nothing to read back, the build stays asynchronous
    for (int i : todo) refinable = refinable || !(g[i]->force_cell > 0);
    for (int i : todo) {
      rc = grid_enqueue_build(h, g[i], n[i]);
      if (rc) return rc;
      if (refinable) CK(cudaMemcpyAsync(&h->h_bbox[i], g[i]->bbox.p, sizeof(BBox), cudaMemcpyDeviceToHost, h->stream));
    }
    if (!refinable) break;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    std::vector<int> again;
    for (int i : todo) {
      const int occ_cells = std::max(1, h->h_bbox[i].occupied);
      g[i]->occupancy = (double)n[i] / (double)occ_cells;
      if (g[i]->occupancy > kMaxOccupancy && g[i]->cell > g[i]->min_cell * 1.0001) {
        // surfaces: occupancy scales ~ cell^2
        double shrink = std::sqrt(kTargetOccupancy / g[i]->occupancy);
        g[i]->cell = std::max(g[i]->min_cell, g[i]->cell * std::max(shrink, 0.25));
        again.push_back(i);
      }
    }
    todo.swap(again);
  }
  for (int i = 0; i < count; ++i) {
    g[i]->valid = true;
    g[i]->hint_cell = g[i]->cell;
    g[i]->hint_n = n[i];
    for (int d = 0; d < 3; ++d) g[i]->hint_ext[d] = (double)g[i]->mx[d] - (double)g[i]->mn[d];
  }
  return B2ICP_OK;
}

int upload_cloud(b2icp_handle* h, Cloud& c, const float* xyzw, size_t n, bool from_device, cudaStream_t on = nullptr) {
  if (!xyzw || n == 0) return fail(h, B2ICP_ERR_EMPTY_CLOUD, "empty cloud");
  if (n > (size_t)INT32_MAX / 8) return fail(h, B2ICP_ERR_INVALID_ARG, "cloud too large");
  CK(c.raw.ensure(n * sizeof(float4)));
  CK(cudaMemcpyAsync(c.raw.p, xyzw, n * sizeof(float4), from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                     on ? on : h->stream));
  c.ext = nullptr;
  c.n = n;
  c.valid = true;
  return B2ICP_OK;
}

// The neighbour grid of a target whose build was deferred (GridSlot::lazy).
int ensure_grid(b2icp_handle* h, GridSlot& g) {
  if (!g.lazy) return B2ICP_OK;
  GridSlot* gp = &g;
  size_t n = g.tgt.n;
  const int rc = build_grids(h, &gp, &n, 1);
  if (!rc) g.lazy = false;
  return rc;
}

bool tiny_target_ok(const b2icp_handle* h, size_t n) {
  return h->use_tiny && h->params.mode == B2ICP_MODE_P2P_SVD && h->params.profile == 0 && n >= 1 && n <= (size_t)kTinyMax;
}

int set_target_impl(b2icp_handle* h, int gi, const float* xyzw, size_t n, bool from_device) {
  GridSlot& g = gslot(h, gi);
  g.valid = false;
  g.lazy = false;
  g.tgt.valid = false;
  if (gi == 0) h->aligned = false;
  if (gi == 0 && !from_device && xyzw && tiny_target_ok(h, n)) {
    // a small target of the handle itself (the odometer's previous scan): the grid is built only if something other
    // than the single-launch loop needs it.  What the build would have reported is checked here, on the host.
    for (size_t i = 0; i < n; ++i)
      if (!(std::isfinite(xyzw[4 * i]) && std::isfinite(xyzw[4 * i + 1]) && std::isfinite(xyzw[4 * i + 2])))
        return fail(h, B2ICP_ERR_NONFINITE_INPUT, "target cloud holds non-finite coordinates");
    int rc = upload_cloud(h, g.tgt, xyzw, n, false);
    if (rc) return rc;
    g.pts = g.tgt.raw.as<float4>();
    g.cov_valid = false;
    g.lazy = true;
    g.valid = true;
    return B2ICP_OK;
  }
  int rc = upload_cloud(h, g.tgt, xyzw, n, from_device);
  if (rc) return rc;
  g.pts = g.tgt.raw.as<float4>();
  GridSlot* gp = &g;
  rc = build_grids(h, &gp, &n, 1);
  if (rc) g.tgt.valid = false;
  return rc;
}

int ensure_slot_work(b2icp_handle* h, ScanSlot& s) {
  const size_t n = s.src.n;
  const size_t ncta = (n + 31) / 32;  // per-warp partial sums, worst case one 32-query slab per warp
  CK(s.cur.ensure(n * sizeof(float4)));
  CK(s.corr_idx.ensure(n * sizeof(int)));
  CK(s.corr_d2.ensure(n * sizeof(float)));
  CK(s.c0.ensure(n * sizeof(float4)));
  CK(s.c1.ensure(n * sizeof(float4)));
  if (kCacheK > 2) CK(s.c2.ensure(n * sizeof(float4)));
  CK(s.partials.ensure(ncta * kNumSums * sizeof(double)));
  return B2ICP_OK;
}

void identity_result(b2icp_result* out) {
  std::memset(out, 0, sizeof(*out));
  for (int i = 0; i < 4; ++i) out->T[5 * i] = 1.0;
  out->mse_last = std::nan("");
  out->fitness = std::nan("");
}

void fill_result(const IcpState& s, b2icp_result* out) {
  for (int i = 0; i < 16; ++i) out->T[i] = (double)s.final_T[i];
  out->converged = s.converged;
  out->iterations = s.iter;
  out->n_corr_last = s.n_corr;
  out->status_detail = s.status;
  out->mse_last = s.mse;
  out->fitness = std::nan("");
}

// Advance slots [0, B) to convergence: one fused sweep launch per iteration for the whole batch.
// guesses: B x 16 floats or NULL.  Results are read back by read_states().
int run_batch(b2icp_handle* h, int B, const float* guesses, int slot0 = 0, bool allow_prof = true, bool streamed = false) {
  if (h->params.mode != B2ICP_MODE_P2P_SVD) return fail(h, B2ICP_ERR_INVALID_ARG, "mode not implemented");
  size_t max_n = 0;
  double min_cell = 1e300;
  int max_dim = 1;
  for (int i = 0; i < B; ++i) {
    ScanSlot& s = slot(h, (size_t)(slot0 + i));
    GridSlot& g = gslot(h, s.grid);
    int rc = ensure_grid(h, g);
    if (rc) return rc;
    rc = ensure_slot_work(h, s);
    if (rc) return rc;
    max_n = std::max(max_n, s.src.n);
    min_cell = std::min(min_cell, (double)g.view.cell);
    max_dim = std::max(max_dim, std::max(g.view.nx, std::max(g.view.ny, g.view.nz)));
    ScanTask& t = h->h_tasks[slot0 + i];
    t.grid = g.view;
    t.src = s.src.dev();
    t.cur = s.cur.as<float4>();
    t.corr_idx = s.corr_idx.as<int>();
    t.corr_d2 = s.corr_d2.as<float>();
    t.c0 = s.c0.as<float4>();
    t.c1 = s.c1.as<float4>();
    t.c2 = s.c2.as<float4>();
    t.partials = s.partials.as<double>();
    t.state = h->states.as<IcpState>() + slot0 + i;
    t.n = (int)s.src.n;
    t.pad = 1;  // the loop leaves a certificate per query: getFitnessScore starts from it
    IcpState& st = h->h_states[slot0 + i];
    std::memset(&st, 0, sizeof(st));
    for (int k = 0; k < 16; ++k) {
      const float v = guesses ? guesses[16 * i + k] : ((k % 5 == 0) ? 1.f : 0.f);
      st.Tinc[k] = v;
      st.final_T[k] = v;
    }
    st.mse = std::nan("");
    st.prev_mse = DBL_MAX;
  }
  // An unbounded gate (PCL's own default, sqrt(DBL_MAX)) has no radius to clamp a search to: the box of a query
  // with nothing nearby may then grow to the whole grid — exact, slow, and off the reference's 1.0 m path.
  h->cfg.max_rings = std::isfinite(h->cfg.bound2) ? rings_for_bound(h, min_cell) + (int)std::ceil(h->cfg.margin_frac) + 1 : max_dim;
  const bool prof = allow_prof && h->params.profile != 0;
  const int iters = std::max(h->params.max_iterations, 1);
  if (prof) {
    while ((int)h->events.size() < 2 * iters + 2) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      h->events.push_back(e);
    }
  }
  // queries per lane (a warp owns 32 * qpt consecutive queries): long slabs amortise the reduction and keep
  // the lanes of the search phase busy; short ones are for launches that could not fill the SMs otherwise
  // (measured on B200, 32 x 64k sweeps: qpt 2 -> 144 us, 4 -> 113 us, 8 -> 109 us per iteration)
  auto ctas_at = [&](int qpt) { return (long long)B * (long long)((max_n + (size_t)kSweepThreads * qpt - 1) / ((size_t)kSweepThreads * qpt)); };
  const long long want = 2LL * 148 * kSweepMinCtas;
  int qpt = h->qpt_override > 0 ? h->qpt_override
                                : (ctas_at(8) >= want ? 8 : (ctas_at(4) >= want ? 4 : (ctas_at(2) >= want ? 2 : 1)));
  if (streamed && h->qpt_override <= 0 && ctas_at(kStreamedQpt) >= 128) qpt = kStreamedQpt;
  // The slab length can change from one launch to the next (the work list lives inside a launch).  The first
  // sweeps search most queries, so their warps are long-running whatever the slab: shorter slabs there keep
  // the last wave of CTAs from running on a third of the machine.
  auto qpt_at = [&](int it) {
    if (!h->qpt_sched.empty()) return h->qpt_sched[std::min<size_t>((size_t)it, h->qpt_sched.size() - 1)];
    // measured on B200 (32 x 64k sweeps, scans/s): 8 everywhere 10 144; 2,8.. 10 451; 2,4,4,8.. 10 533; 2,4,4,4,4,8.. 10 571
    // A streamed batch shares the device with the other batches in flight, which fill its tails: long slabs
    // everywhere are best there (18.6k scans/s against 17.2k with the schedule below).
    if (streamed) return qpt;
    return it == 0 ? std::min(qpt, kFirstSweepQpt) : (it <= 4 ? std::min(qpt, 4) : qpt);
  };
  // task / state upload + one sweep per iteration + the correspondence write-out, in order on `st`
  auto enqueue_loop = [&](cudaStream_t st, bool with_events) -> cudaError_t {
    cudaError_t e = cudaMemcpyAsync(h->tasks.as<ScanTask>() + slot0, h->h_tasks + slot0, sizeof(ScanTask) * B, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(h->states.as<IcpState>() + slot0, h->h_states + slot0, sizeof(IcpState) * B, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    for (int it = 0; it < iters; ++it) {
      if (with_events) cudaEventRecord(h->events[2 + 2 * it], st);
      const int q = qpt_at(it);
      const dim3 grid((unsigned)((max_n + (size_t)kSweepThreads * q - 1) / ((size_t)kSweepThreads * q)), (unsigned)B, 1);
      launch_sweep(q, grid, st, h->tasks.as<ScanTask>() + slot0, h->cfg);
      if (with_events) cudaEventRecord(h->events[3 + 2 * it], st);
    }
    // correspondences of the last sweep, for b2icp_get_correspondences
    const dim3 fgrid((unsigned)((max_n + 255) / 256), (unsigned)B, 1);
    icp_finalize_corr<<<fgrid, 256, 0, st>>>(h->tasks.as<ScanTask>() + slot0, h->cfg);
    return cudaSuccess;
  };
  // The loop is the same 32 operations every time a batch of this shape comes back (streamed replay, the
  // odometer's scan after scan): the second time a shape is seen its loop is captured into a CUDA graph, and from
  // then on one graph launch replaces the 32 stream operations — the host cost per batch no longer grows with the
  // iteration count, which is what kept 8 ranks on one 16-core host from scaling (profiles/r02_scaling.md).
  bool launched = false;
  if (!prof && h->use_graphs) {
    GraphKey key;
    std::memset(&key, 0, sizeof(key));
    key.B = B;
    key.iters = iters;
    key.qpt = qpt;
    key.streamed = streamed ? 1 : 0;
    key.max_n = max_n;
    key.cfg = h->cfg;
    key.sched = h->qpt_sched.empty() ? 0 : 1;
    GraphCache& gc = h->graphs[streamed ? slot0 / kSetSlots : kStreamSets];
    if (!(gc.exec && std::memcmp(&gc.exec_key, &key, sizeof(key)) == 0) && gc.has_last &&
        std::memcmp(&gc.last_key, &key, sizeof(key)) == 0) {
      if (!h->capture_stream) CK(cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
      if (gc.exec) cudaGraphExecDestroy(gc.exec);
      gc.exec = nullptr;
      cudaGraph_t graph = nullptr;
      CK(cudaStreamBeginCapture(h->capture_stream, cudaStreamCaptureModeRelaxed));
      const cudaError_t ce = enqueue_loop(h->capture_stream, false);
      const cudaError_t ee = cudaStreamEndCapture(h->capture_stream, &graph);
      if (ce == cudaSuccess && ee == cudaSuccess && graph && cudaGraphInstantiate(&gc.exec, graph, 0) == cudaSuccess) {
        gc.exec_key = key;
      } else {
        gc.exec = nullptr;
        cudaGetLastError();
      }
      if (graph) cudaGraphDestroy(graph);
    }
    gc.last_key = key;
    gc.has_last = true;
    if (gc.exec && std::memcmp(&gc.exec_key, &key, sizeof(key)) == 0) {
      CK(cudaGraphLaunch(gc.exec, h->stream));
      h->graph_launches += 1;
      launched = true;
    }
  }
  if (!launched) {
    if (prof) CK(cudaEventRecord(h->events[0], h->stream));
    CK(enqueue_loop(h->stream, prof));
    if (prof) CK(cudaEventRecord(h->events[1], h->stream));
  }
  h->launches += iters + 1;
  h->last_batch = B;
  return B2ICP_OK;
}

int read_states(b2icp_handle* h, int B) {
  CK(cudaMemcpyAsync(h->h_states, h->states.p, sizeof(IcpState) * B, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  if (h->params.profile != 0 && h->events.size() >= 2) {
    b2icp_timing& t = h->timing;
    std::memset(&t, 0, sizeof(t));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->events[0], h->events[1]);
    t.total_ms = ms;
    int ran = 1;
    for (int i = 0; i < B; ++i) ran = std::max(ran, h->h_states[i].iter);
    ran = std::min(ran, std::max(h->params.max_iterations, 1));
    for (int it = 0; it < ran; ++it) {
      cudaEventElapsedTime(&ms, h->events[2 + 2 * it], h->events[3 + 2 * it]);
      t.nn_sweep_ms += ms;
    }
    if (getenv("B2ICP_DUMP_ITERS")) {  // tuning only: per-iteration device time of the last batch
      const int all = std::max(h->params.max_iterations, 1);
      fprintf(stderr, "[b2icp] total %.1f us; per iteration (us):", 1e3 * t.total_ms);
      for (int it = 0; it < all; ++it) {
        cudaEventElapsedTime(&ms, h->events[2 + 2 * it], h->events[3 + 2 * it]);
        int active = 0;
        for (int i = 0; i < B; ++i) active += h->h_states[i].iter > it;
        fprintf(stderr, " %d:%.1f(%d)", it, 1e3 * ms, active);
      }
      fprintf(stderr, "\n");
    }
    t.nn_sweep_launches = ran;
    for (int i = 0; i < B; ++i) t.nn_searches += h->h_states[i].unresolved;
  }
  h->timing.kernel_launches = h->launches;
  return B2ICP_OK;
}

const char* status_message(int s) {
  switch (s) {
    case B2ICP_ERR_NONFINITE_INPUT: return "non-finite source point or transform";
    case B2ICP_ERR_NOT_ENOUGH_CORRESPONDENCES: return "ICP loop ended early: not enough correspondences";
    default: return "ICP loop failed";
  }
}

int launch_brute_fallback(b2icp_handle* h, const GridView& view, const float4* d_q, size_t n, int* d_idx, float* d_d2);

// Enqueue getFitnessScore for slot i (result lands in states[i].fitness_*).
int enqueue_fitness(b2icp_handle* h, int i, double max_range) {
  ScanSlot& s = slot(h, i);
  GridSlot& g = gslot(h, s.grid);
  if (g.lazy) {  // the loop ran without a grid (single-launch path): build it and point the scan's task at it
    int rc = ensure_grid(h, g);
    if (rc) return rc;
    h->h_tasks[i].grid = g.view;
    h->h_tasks[i].pad = 0;
    CK(cudaMemcpyAsync(h->tasks.as<ScanTask>() + i, h->h_tasks + i, sizeof(ScanTask), cudaMemcpyHostToDevice, h->stream));
  }
  const size_t n = s.src.n;
  CK(h->query.ensure(n * sizeof(float4)));
  CK(h->q_idx.ensure(n * sizeof(int)));
  CK(h->q_d2.ensure(n * sizeof(float)));
  CK(h->unres_list.ensure(n * sizeof(int)));
  zero_counter<<<1, 1, 0, h->stream>>>(h->unres_count.as<unsigned int>());
  dim3 grid((unsigned)((n + kSweepThreads - 1) / kSweepThreads), 1, 1);
  fitness_kernel<<<grid, kSweepThreads, 0, h->stream>>>(h->tasks.as<ScanTask>() + i, h->fitness_rings,
                                                        h->unres_list.as<int>(), h->unres_count.as<unsigned int>(),
                                                        h->query.as<float4>(), h->q_idx.as<int>(), h->q_d2.as<float>());
  {
    int rc = launch_brute_fallback(h, g.view, h->query.as<float4>(), n, h->q_idx.as<int>(), h->q_d2.as<float>());
    if (rc) return rc;
  }
  fitness_reduce<<<1, 1024, 0, h->stream>>>(h->q_idx.as<int>(), h->q_d2.as<float>(), (int)n, max_range,
                                            h->states.as<IcpState>() + i);
  h->launches += 4;
  return B2ICP_OK;
}

double fitness_value(const IcpState& s) {
  return s.fitness_cnt > 0 ? s.fitness_sum / (double)s.fitness_cnt : DBL_MAX;
}

// exhaustive scan of the queries listed in unres_list (nn.cuh): keys reset, scan, unpack
int launch_brute_fallback(b2icp_handle* h, const GridView& view, const float4* d_q, size_t n, int* d_idx, float* d_d2) {
  CK(h->unres_keys.ensure(n * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(h->unres_keys.p, 0xFF, n * sizeof(unsigned long long), h->stream));
  nn_brute_fallback<<<dim3(148 * 2, 8, 1), 256, 0, h->stream>>>(view, d_q, h->unres_list.as<int>(),
                                                             h->unres_count.as<unsigned int>(),
                                                             h->unres_keys.as<unsigned long long>());
  nn_brute_unpack<<<148, 256, 0, h->stream>>>(h->unres_list.as<int>(), h->unres_count.as<unsigned int>(),
                                              h->unres_keys.as<unsigned long long>(), d_idx, d_d2);
  h->launches += 2;
  return B2ICP_OK;
}

int nn_search_impl(b2icp_handle* h, const float4* d_q, size_t n, int* d_idx, float* d_d2, int grid_index = 0) {
  GridSlot& g = gslot(h, (size_t)grid_index);
  {
    int rc = ensure_grid(h, g);
    if (rc) return rc;
  }
  CK(h->unres_list.ensure(n * sizeof(int)));
  zero_counter<<<1, 1, 0, h->stream>>>(h->unres_count.as<unsigned int>());
  // 32 consecutive queries per cooperative group (coop.cuh); a grid that is a multiple of the SM count
  const int groups = (int)((n + 31) / 32);
  const int ctas = std::max(1, std::min((groups + kSweepThreads / 32 - 1) / (kSweepThreads / 32), 148 * kSweepMinCtas * 4));
  const int ncell = g.view.nx * g.view.ny * g.view.nz;
  // Large query clouds are first counting-sorted by target cell (three streaming passes over the queries and one over
  // the cell table): groups of consecutive queries then share their candidates.  Small ones are searched in place.
  const bool sort_queries = h->nn_sort_override >= 0 ? h->nn_sort_override != 0 : (n >= 65536 && n >= (size_t)ncell / 8);
  const float4* q_in = d_q;
  if (sort_queries) {
    CK(h->qs_sorted.ensure(n * sizeof(float4)));
    CK(h->qs_cell_of.ensure(n * sizeof(int)));
    CK(h->qs_rank.ensure(n * sizeof(int)));
    CK(h->qs_count.ensure((size_t)(ncell + 1 + 4) * sizeof(int)));
    const int ntiles = (ncell + kScanTile - 1) / kScanTile;
    CK(h->qs_tiles.ensure((size_t)ntiles * sizeof(int)));
    CK(h->map_stats.ensure(sizeof(BBox)));
    int* cs = h->qs_count.as<int>();
    const int blocks = (int)((n + 255) / 256);
    CK(cudaMemsetAsync(cs, 0, (size_t)(ncell + 1) * sizeof(int), h->stream));
    grid_count<<<blocks, 256, 0, h->stream>>>(d_q, (int)n, g.view, h->qs_cell_of.as<int>(), h->qs_rank.as<int>(), cs);
    scan_tile_sums<<<ntiles, kScanThreads, 0, h->stream>>>(cs, ncell, h->qs_tiles.as<int>(), h->map_stats.as<BBox>());
    scan_of_sums<<<1, kScanThreads, 0, h->stream>>>(h->qs_tiles.as<int>(), ntiles);
    scan_apply<<<ntiles, kScanThreads, 0, h->stream>>>(cs, ncell, h->qs_tiles.as<int>(), (int)n);
    grid_scatter<<<blocks, 256, 0, h->stream>>>(d_q, (int)n, h->qs_cell_of.as<int>(), h->qs_rank.as<int>(), cs,
                                                 h->qs_sorted.as<float4>());
    h->launches += 5;
    q_in = h->qs_sorted.as<float4>();
  }
#define B2_NN_LAUNCH(W, S)                                                                                                       \
  nn_search_coop<W, S><<<ctas, kSweepThreads, 0, h->stream>>>(g.view, q_in, (int)n, kUnboundedRings, h->join_d, d_idx, d_d2,      \
                                                              h->unres_list.as<int>(), h->unres_count.as<unsigned int>())
  if (h->w_override == 8) {
    if (sort_queries) B2_NN_LAUNCH(8, true); else B2_NN_LAUNCH(8, false);
  } else {
    if (sort_queries) B2_NN_LAUNCH(32, true); else B2_NN_LAUNCH(32, false);
  }
#undef B2_NN_LAUNCH
  h->launches += 2;
  return launch_brute_fallback(h, g.view, d_q, n, d_idx, d_d2);
}

#include "gicp_host.inl"

}  // namespace
extern "C" {  // defined with the streaming entry points below
static int submit_impl(b2icp_handle* h, const float* const* src, const size_t* n_src, size_t batch, int with_fitness,
                       bool from_device);
static int wait_impl(b2icp_handle* h, b2icp_result* out, size_t capacity, size_t* n_out);
}
namespace {

// A synchronous batch against the handle's target is streamed internally: chunks of scans on the slot sets /
// streams of b2icp_align_batch_submit, up to kStreamSets in flight, so that one call gets the overlap a caller of
// the streaming pair gets (sparse late iterations of one chunk under the dense first ones of the next, uploads under
// sweeps).  Results come back in input order.
int streamed_sync_batch(b2icp_handle* h, const float* const* src, const size_t* n_src, size_t batch, int with_fitness,
                        b2icp_result* out, bool from_device) {
  const size_t chunk = (size_t)kSetSlots;
  size_t submitted = 0, done = 0;
  h->next_set = 0;  // nothing is in flight: start from the first slot set, so that short batches reuse the same sets
  int worst = B2ICP_OK;
  while (done < batch) {
    while (submitted < batch && h->n_pending < kStreamSets) {
      const size_t len = std::min(chunk, batch - submitted);
      int rc = submit_impl(h, src + submitted, n_src + submitted, len, with_fitness, from_device);
      if (rc) {
        b2icp_result scratch[kSetSlots];
        while (h->n_pending) wait_impl(h, scratch, kSetSlots, nullptr);
        return rc;
      }
      submitted += len;
    }
    size_t got = 0;
    int rc = wait_impl(h, out + done, batch - done, &got);
    if (rc > worst || (rc < 0 && worst == B2ICP_OK)) worst = rc;
    if (got == 0) return rc ? rc : B2ICP_ERR_CUDA;
    done += got;
  }
  return worst;
}

// GICP over a batch: chunks of up to 32 scans whose BFGS recursions advance in lockstep rounds (gicp_host.inl), with
// the same target conventions as the point-to-point batch.
int gicp_batch_impl(b2icp_handle* h, const float* const* src, const size_t* n_src, const float* const* tgt,
                    const size_t* n_tgt, size_t batch, int with_fitness, b2icp_result* out, bool from_device) {
  const bool shared_target = (tgt == nullptr);
  if ((shared_target || !tgt[0]) && !gslot(h, 0).valid) return fail(h, B2ICP_ERR_NO_TARGET, "no target cloud set");
  int worst = B2ICP_OK;
  h->aligned = false;
  constexpr int kChunk = kSetSlots;
  for (size_t base = 0; base < batch; base += kChunk) {
    const int B = (int)std::min<size_t>(kChunk, batch - base);
    const bool carry = !shared_target && base > 0 && !tgt[base];
    if (carry) {  // consecutive sweeps across a chunk border: the last source of the previous chunk is the target
      GridSlot& g1 = gslot(h, 1);
      ScanSlot& last = slot(h, kChunk - 1);
      std::swap(g1.tgt.raw, last.src.raw);
      g1.tgt.n = last.src.n;
      g1.tgt.valid = true;
      last.src.valid = false;
    }
    for (int i = 0; i < B; ++i) {
      int rc = upload_cloud(h, slot(h, i).src, src[base + i], n_src[base + i], from_device);
      if (rc) return rc;
    }
    if (shared_target) {
      for (int i = 0; i < B; ++i) slot(h, i).grid = 0;
    } else {
      std::vector<GridSlot*> gl;
      std::vector<size_t> nl;
      for (int i = 0; i < B; ++i) {
        const size_t gi = 1 + (size_t)i;  // grid 0 stays the handle's own target
        GridSlot& g = gslot(h, gi);
        const float* tp = tgt[base + i];
        size_t tn = 0;
        if (tp) {
          tn = n_tgt ? n_tgt[base + i] : 0;
          int rc = upload_cloud(h, g.tgt, tp, tn, from_device);
          if (rc) return rc;
          g.pts = g.tgt.raw.as<float4>();
        } else if (i > 0) {
          g.pts = slot(h, i - 1).src.raw.as<float4>();
          tn = slot(h, i - 1).src.n;
        } else if (carry) {
          g.pts = g.tgt.raw.as<float4>();
          tn = g.tgt.n;
        } else {
          slot(h, i).grid = 0;  // pair 0 without a target: the handle's current target
          continue;
        }
        slot(h, i).grid = (int)gi;
        gl.push_back(&g);
        nl.push_back(tn);
      }
      if (!gl.empty()) {
        int rc = build_grids(h, gl.data(), nl.data(), (int)gl.size());
        if (rc) return rc;
      }
    }
    int rc = run_gicp_batch(h, B, nullptr);
    if (rc == B2ICP_ERR_CUDA) return rc;
    if (rc && worst == B2ICP_OK) worst = rc;
    for (int i = 0; i < B; ++i) {
      fill_result(h->h_states[i], &out[base + i]);
      const int status = h->h_states[i].status;
      if (status != 0 && worst == B2ICP_OK) {
        worst = status;
        h->err = status_message(worst);
      }
    }
    if (with_fitness) {
      bool any = false;
      for (int i = 0; i < B; ++i)
        if (h->h_states[i].status == 0) {
          rc = enqueue_fitness(h, i, DBL_MAX);
          if (rc) return rc;
          any = true;
        }
      if (any) {
        CK(cudaMemcpyAsync(h->h_states, h->states.p, sizeof(IcpState) * B, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < B; ++i)
          if (h->h_states[i].status == 0) out[base + i].fitness = fitness_value(h->h_states[i]);
      }
    }
  }
  CK(cudaGetLastError());
  return worst;
}

int batch_impl(b2icp_handle* h, const float* const* src, const size_t* n_src, const float* const* tgt,
               const size_t* n_tgt, size_t batch, int with_fitness, b2icp_result* out, bool from_device) {
  for (size_t i = 0; i < batch; ++i) identity_result(&out[i]);
  if (batch == 0) return B2ICP_OK;
  const bool shared_target = (tgt == nullptr);
  if (h->params.mode == B2ICP_MODE_GICP_BFGS) return gicp_batch_impl(h, src, n_src, tgt, n_tgt, batch, with_fitness, out, from_device);
  if (shared_target && !gslot(h, 0).valid) return fail(h, B2ICP_ERR_NO_TARGET, "no target cloud set");
  if (!shared_target && !tgt[0] && !gslot(h, 0).valid) return fail(h, B2ICP_ERR_NO_TARGET, "pair 0 has no target");
  // (measured, 128 host sweeps per call: 15.2k scans/s streamed internally against 10.9k as one launch sequence)
  if (shared_target && batch >= 64 && h->params.profile == 0 && h->n_pending == 0 && getenv("B2ICP_NO_INTERNAL_STREAMING") == nullptr)
    return streamed_sync_batch(h, src, n_src, batch, with_fitness, out, from_device);
  int worst = B2ICP_OK;
  h->aligned = false;
  // In consecutive-sweep mode pair i (tgt[i] == NULL) registers against src[i-1]; across chunk borders the
  // predecessor's device copy is handed over to the next chunk's first grid slot.
  for (size_t base = 0; base < batch; base += kMaxBatch) {
    const int B = (int)std::min<size_t>(kMaxBatch, batch - base);
    const bool carry = !shared_target && base > 0 && !tgt[base];
    if (carry) {
      GridSlot& g1 = gslot(h, 1);
      ScanSlot& last = slot(h, kMaxBatch - 1);
      std::swap(g1.tgt.raw, last.src.raw);
      g1.tgt.n = last.src.n;
      g1.tgt.valid = true;
      last.src.valid = false;
    }
    // 1. upload sources
    for (int i = 0; i < B; ++i) {
      int rc = upload_cloud(h, slot(h, i).src, src[base + i], n_src[base + i], from_device);
      if (rc) return rc;
    }
    // 2. targets / grids
    if (shared_target) {
      for (int i = 0; i < B; ++i) slot(h, i).grid = 0;
    } else {
      std::vector<GridSlot*> gl;
      std::vector<size_t> nl;
      for (int i = 0; i < B; ++i) {
        const size_t gi = 1 + (size_t)i;  // grid 0 stays the handle's own target
        GridSlot& g = gslot(h, gi);
        const float* tp = tgt[base + i];
        size_t tn = 0;
        if (tp) {
          tn = n_tgt ? n_tgt[base + i] : 0;
          int rc = upload_cloud(h, g.tgt, tp, tn, from_device);
          if (rc) return rc;
          g.pts = g.tgt.raw.as<float4>();
        } else if (i > 0) {
          g.pts = slot(h, i - 1).src.raw.as<float4>();
          tn = slot(h, i - 1).src.n;
        } else if (carry) {
          g.pts = g.tgt.raw.as<float4>();
          tn = g.tgt.n;
        } else {
          slot(h, i).grid = 0;  // pair 0 without a target: the handle's current target
          continue;
        }
        slot(h, i).grid = (int)gi;
        gl.push_back(&g);
        nl.push_back(tn);
      }
      if (!gl.empty()) {
        int rc = build_grids(h, gl.data(), nl.data(), (int)gl.size());
        if (rc) return rc;
      }
    }
    // 3. the ICP loops of the whole chunk
    int rc = run_batch(h, B, nullptr);
    if (rc) return rc;
    if (with_fitness)
      for (int i = 0; i < B; ++i) {
        rc = enqueue_fitness(h, i, DBL_MAX);
        if (rc) return rc;
      }
    rc = read_states(h, B);
    if (rc) return rc;
    for (int i = 0; i < B; ++i) {
      fill_result(h->h_states[i], &out[base + i]);
      if (with_fitness && h->h_states[i].status == 0) out[base + i].fitness = fitness_value(h->h_states[i]);
      if (h->h_states[i].status != 0 && worst == B2ICP_OK) {
        worst = h->h_states[i].status;
        h->err = status_message(worst);
      }
    }
  }
  return worst;
}

}  // namespace

extern "C" {

int b2icp_default_params(b2icp_params* p, int preset) {
  if (!p) return B2ICP_ERR_INVALID_ARG;
  std::memset(p, 0, sizeof(*p));
  p->mode = B2ICP_MODE_P2P_SVD;
  p->max_iterations = preset == B2ICP_PRESET_MAPPER ? 30 : 10;  // octree_mapper.h:56 / icp_odometer.h:65
  p->transformation_epsilon = 1e-6;                             // ICP_EPSILON
  p->max_correspondence_distance = 1.0;                         // ICP_MAX_CORR_DIST
  p->euclidean_fitness_epsilon = -DBL_MAX;                      // PCL default
  p->rotation_epsilon = 2e-3;
  p->gicp_epsilon = 1e-3;
  p->k_correspondences = 20;
  p->max_inner_iterations = 20;
  p->device = 0;
  p->profile = 0;
  p->grid_cell = 0.0f;
  return B2ICP_OK;
}

static int validate_params(const b2icp_params* p) {
  if (!p) return B2ICP_ERR_INVALID_ARG;
  if (p->mode != B2ICP_MODE_P2P_SVD && p->mode != B2ICP_MODE_GICP_BFGS) return B2ICP_ERR_INVALID_ARG;
  if (p->max_iterations < 1 || p->max_iterations > 100000) return B2ICP_ERR_INVALID_ARG;
  if (!(p->max_correspondence_distance > 0)) return B2ICP_ERR_INVALID_ARG;
  if (!(p->transformation_epsilon >= 0)) return B2ICP_ERR_INVALID_ARG;
  return B2ICP_OK;
}

int b2icp_create(const b2icp_params* p, b2icp_handle** out) {
  if (!out) return B2ICP_ERR_INVALID_ARG;
  *out = nullptr;
  if (validate_params(p)) return B2ICP_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || p->device < 0 || p->device >= ndev) {
    cudaGetLastError();
    return B2ICP_ERR_CUDA;  // no CUDA device: there is no CPU path behind this ABI
  }
  b2icp_handle* h = new (std::nothrow) b2icp_handle();
  if (!h) return B2ICP_ERR_CUDA;
  h->params = *p;
  h->device = p->device;
  std::memset(&h->timing, 0, sizeof(h->timing));
  derive_config(h);
  if (const char* e = getenv("B2ICP_QPT")) h->qpt_override = atoi(e);
  if (getenv("B2ICP_NO_GRAPH")) h->use_graphs = false;
  if (const char* e = getenv("B2ICP_GICP_GROUPS")) h->gicp_groups = std::max(1, atoi(e));
  if (const char* e = getenv("B2ICP_FITNESS_RINGS")) h->fitness_rings = std::max(1, atoi(e));
  if (getenv("B2ICP_NO_TINY")) h->use_tiny = false;
  if (const char* e = getenv("B2ICP_W")) h->w_override = atoi(e);
  if (const char* e = getenv("B2ICP_JOIN")) h->join_d = atoi(e);
  if (const char* e = getenv("B2ICP_NN_SORT")) h->nn_sort_override = atoi(e);
  if (const char* e = getenv("B2ICP_IN_FLIGHT")) h->max_in_flight = std::max(1, std::min(kStreamSets, atoi(e)));
  if (const char* e = getenv("B2ICP_QPT_SCHED"))
    for (const char* p = e; *p;) {
      const int v = atoi(p);
      h->qpt_sched.push_back(v >= 32 ? 32 : v >= 16 ? 16 : (v >= 8 ? 8 : (v >= 4 ? 4 : (v >= 2 ? 2 : 1))));
      while (*p && *p != ',') ++p;
      if (*p == ',') ++p;
    }
  if (const char* e = getenv("B2ICP_CARVEOUT")) {  // tuning only: shared-memory carve-out of the sweep, percent
    const int pct = atoi(e);
    cudaFuncSetAttribute(icp_sweep_p2p<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(icp_sweep_p2p<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(icp_sweep_p2p<4>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(icp_sweep_p2p<8>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  }
  slot(h, 0);
  gslot(h, 0);
  bool ok = cudaSetDevice(h->device) == cudaSuccess &&
            cudaFuncSetAttribute(icp_tiny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTinyMax * 16) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMallocHost((void**)&h->h_states, sizeof(IcpState) * kSlots) == cudaSuccess &&
            cudaMallocHost((void**)&h->h_tasks, sizeof(ScanTask) * kSlots) == cudaSuccess &&
            cudaMallocHost((void**)&h->h_bbox, sizeof(BBox) * (kMaxBatch + 1)) == cudaSuccess &&
            cudaMallocHost((void**)&h->h_first, sizeof(int)) == cudaSuccess &&
            cudaMallocHost((void**)&h->h_point, 4 * sizeof(float)) == cudaSuccess &&
            h->states.ensure(sizeof(IcpState) * kSlots) == cudaSuccess &&
            h->tasks.ensure(sizeof(ScanTask) * kSlots) == cudaSuccess &&
            h->unres_count.ensure(sizeof(unsigned int)) == cudaSuccess && h->mat.ensure(16 * sizeof(double)) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    b2icp_destroy(h);
    return B2ICP_ERR_CUDA;
  }
  *out = h;
  return B2ICP_OK;
}

int b2icp_destroy(b2icp_handle* h) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& s : h->slots) s->release();
  for (auto& g : h->grids) g->release();
  for (DeviceBuf* b : {&h->map_pts, &h->map_keys, &h->map_vals, &h->map_slot_of, &h->map_flags, &h->map_tiles, &h->map_stats})
    b->release();
  for (DeviceBuf* b : {&h->states, &h->tasks, &h->unres_list, &h->unres_count, &h->unres_keys, &h->query, &h->q_idx, &h->q_d2, &h->xf_in,
                       &h->xf_out, &h->mat})
    b->release();
  for (cudaEvent_t e : h->events) cudaEventDestroy(e);
  for (int k = 0; k < kStreamSets; ++k) {
    if (h->set_uploaded[k]) cudaEventDestroy(h->set_uploaded[k]);
    if (h->set_done[k]) cudaEventDestroy(h->set_done[k]);
    if (h->set_stream[k]) {
      cudaStreamSynchronize(h->set_stream[k]);
      cudaStreamDestroy(h->set_stream[k]);
    }
  }
  if (h->copy_stream) {
    cudaStreamSynchronize(h->copy_stream);
    cudaStreamDestroy(h->copy_stream);
  }
  for (GraphCache& gc : h->graphs)
    if (gc.exec) cudaGraphExecDestroy(gc.exec);
  if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
  for (int k = 0; k < kGicpGroups; ++k)
    if (h->gicp_streams[k]) cudaStreamDestroy(h->gicp_streams[k]);
  if (h->h_states) cudaFreeHost(h->h_states);
  if (h->h_tasks) cudaFreeHost(h->h_tasks);
  if (h->h_bbox) cudaFreeHost(h->h_bbox);
  if (h->h_first) cudaFreeHost(h->h_first);
  if (h->h_point) cudaFreeHost(h->h_point);
  for (DeviceBuf* b : {&h->tree_keys, &h->tree_vals, &h->map_first, &h->qs_sorted, &h->qs_cell_of, &h->qs_rank, &h->qs_count, &h->qs_tiles})
    b->release();
  if (h->h_gicp_partials) cudaFreeHost(h->h_gicp_partials);
  if (h->h_gicp_tasks) cudaFreeHost(h->h_gicp_tasks);
  h->tiny_partials.release();
  for (DeviceBuf* b : {&h->gicp_tasks, &h->gicp_tickets, &h->knn_tasks, &h->knn_list2, &h->knn_counts}) b->release();
  if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return B2ICP_OK;
}

int b2icp_set_stream(b2icp_handle* h, void* cuda_stream) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  h->own_stream = false;
  return B2ICP_OK;
}

int b2icp_set_params(b2icp_handle* h, const b2icp_params* p) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (validate_params(p) || p->device != h->device) return fail(h, B2ICP_ERR_INVALID_ARG, "invalid parameters");
  h->params = *p;
  derive_config(h);
  return B2ICP_OK;
}

int b2icp_set_target(b2icp_handle* h, const float* xyzw, size_t n) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return set_target_impl(h, 0, xyzw, n, false);
}
int b2icp_set_target_device(b2icp_handle* h, const float* d_xyzw, size_t n) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return set_target_impl(h, 0, d_xyzw, n, true);
}
int b2icp_set_source(b2icp_handle* h, const float* xyzw, size_t n) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  h->aligned = false;
  slot(h, 0).src.valid = false;
  slot(h, 0).grid = 0;
  return upload_cloud(h, slot(h, 0).src, xyzw, n, false);
}
int b2icp_set_source_device(b2icp_handle* h, const float* d_xyzw, size_t n) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  h->aligned = false;
  slot(h, 0).src.valid = false;
  slot(h, 0).grid = 0;
  return upload_cloud(h, slot(h, 0).src, d_xyzw, n, true);
}

int b2icp_promote_source_to_target(b2icp_handle* h) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  ScanSlot& s = slot(h, 0);
  GridSlot& g = gslot(h, 0);
  if (!s.src.valid || s.src.ext) return fail(h, B2ICP_ERR_NO_SOURCE, "no source cloud set");
  // a small source that has just been registered successfully is known to be finite: its grid can wait
  const bool defer = h->aligned && h->h_states[0].status == 0 && tiny_target_ok(h, s.src.n);
  std::swap(s.src.raw, g.tgt.raw);
  g.tgt.n = s.src.n;
  g.tgt.valid = true;
  g.pts = g.tgt.raw.as<float4>();
  s.src.valid = false;
  s.src.n = 0;
  h->aligned = false;
  if (defer) {
    g.cov_valid = false;
    g.lazy = true;
    g.valid = true;
    return B2ICP_OK;
  }
  g.lazy = false;
  GridSlot* gp = &g;
  size_t n = g.tgt.n;
  int rc = build_grids(h, &gp, &n, 1);
  if (rc) g.tgt.valid = false;
  return rc;
}

// The whole loop of one small scan pair in one cooperative launch (tiny.cuh); the caller reads the state back.
static int run_tiny(b2icp_handle* h, const float* guess) {
  ScanSlot& s = slot(h, 0);
  GridSlot& g = gslot(h, 0);
  const int ns = (int)s.src.n, nt = (int)g.tgt.n;
  const int G = (ns + kTinyQpc - 1) / kTinyQpc;
  int rc = ensure_slot_work(h, s);
  if (rc) return rc;
  CK(h->tiny_partials.ensure((size_t)2 * G * kNumSums * sizeof(double)));
  IcpState& st = h->h_states[0];
  std::memset(&st, 0, sizeof(st));
  for (int k = 0; k < 16; ++k) {
    const float v = guess ? guess[k] : ((k % 5 == 0) ? 1.f : 0.f);
    st.Tinc[k] = v;
    st.final_T[k] = v;
  }
  st.mse = std::nan("");
  st.prev_mse = DBL_MAX;
  ScanTask& t = h->h_tasks[0];  // what getFitnessScore / the aligned cloud read afterwards
  t.grid = g.view;
  t.src = s.src.dev();
  t.cur = s.cur.as<float4>();
  t.corr_idx = s.corr_idx.as<int>();
  t.corr_d2 = s.corr_d2.as<float>();
  t.c0 = s.c0.as<float4>();
  t.c1 = s.c1.as<float4>();
  t.c2 = s.c2.as<float4>();
  t.partials = s.partials.as<double>();
  t.state = h->states.as<IcpState>();
  t.n = ns;
  t.pad = 0;
  CK(cudaMemcpyAsync(h->tasks.p, h->h_tasks, sizeof(ScanTask), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->states.p, h->h_states, sizeof(IcpState), cudaMemcpyHostToDevice, h->stream));
  TinyArgs a;
  a.src = s.src.dev();
  a.tgt = g.pts;
  a.cur = s.cur.as<float4>();
  a.corr_idx = s.corr_idx.as<int>();
  a.corr_d2 = s.corr_d2.as<float>();
  a.partials = h->tiny_partials.as<double>();
  a.state = h->states.as<IcpState>();
  a.cfg = h->cfg;
  a.ns = ns;
  a.nt = nt;
  a.with_fitness = 1;
  void* args[] = {&a};
  CK(cudaLaunchCooperativeKernel((const void*)icp_tiny_kernel, dim3((unsigned)G), dim3(kTinyThreads), args, (size_t)nt * sizeof(float4),
                                 h->stream));
  h->launches += 1;
  h->last_batch = 1;
  return B2ICP_OK;
}

int b2icp_align(b2icp_handle* h, const float* guess, b2icp_result* out, float* aligned_xyzw) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (!out) return fail(h, B2ICP_ERR_INVALID_ARG, "out == NULL");
  identity_result(out);
  ScanSlot& s = slot(h, 0);
  s.grid = 0;
  if (!s.src.valid) return fail(h, B2ICP_ERR_NO_SOURCE, "no source cloud set");
  if (!gslot(h, 0).valid) return fail(h, B2ICP_ERR_NO_TARGET, "no target cloud set");
  h->aligned = false;
  const bool gicp = h->params.mode == B2ICP_MODE_GICP_BFGS;
  // small pairs (the odometer's 1080-point scans): the whole loop and getFitnessScore in one launch, no grid
  const bool tiny = !gicp && !s.src.ext && tiny_target_ok(h, s.src.n) && tiny_target_ok(h, gslot(h, 0).tgt.n) &&
                    gslot(h, 0).pts == gslot(h, 0).tgt.raw.as<float4>();
  h->tiny_last = false;
  int rc = gicp ? run_gicp(h, guess) : (tiny ? run_tiny(h, guess) : run_batch(h, 1, guess));
  if (rc) return rc;
  h->tiny_last = tiny;
  const size_t n = s.src.n;
  if (aligned_xyzw) {
    CK(h->xf_out.ensure(n * sizeof(float4)));
    transform_cloud_f<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(
        s.src.dev(), (int)n, h->states.as<IcpState>()->final_T, h->xf_out.as<float4>());
    h->launches += 1;
    CK(cudaMemcpyAsync(aligned_xyzw, h->xf_out.p, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
  }
  if (gicp) {  // run_gicp left the final state in h_states[0] and mirrored it to the device
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
  } else {
    rc = read_states(h, 1);
    if (rc) return rc;
  }
  fill_result(h->h_states[0], out);
  h->aligned = true;
  if (h->h_states[0].status != 0) {
    h->err = status_message(h->h_states[0].status);
    return h->h_states[0].status;
  }
  return B2ICP_OK;
}

int b2icp_fitness(b2icp_handle* h, double max_range, double* out) {
  if (!h || !out) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (!h->aligned) return fail(h, B2ICP_ERR_NOT_ALIGNED, "b2icp_fitness needs a completed b2icp_align");
  if (h->tiny_last && !(max_range < DBL_MAX) && h->h_states[0].status == 0) {  // computed by the single-launch loop
    *out = fitness_value(h->h_states[0]);
    return B2ICP_OK;
  }
  int rc = enqueue_fitness(h, 0, max_range);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->h_states, h->states.p, sizeof(IcpState), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  *out = fitness_value(h->h_states[0]);
  return B2ICP_OK;
}

int b2icp_get_correspondences(b2icp_handle* h, int32_t* tgt_idx, float* sqdist) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (!h->aligned) return fail(h, B2ICP_ERR_NOT_ALIGNED, "no completed b2icp_align");
  ScanSlot& s = slot(h, 0);
  const size_t n = s.src.n;
  if (tgt_idx) CK(cudaMemcpyAsync(tgt_idx, s.corr_idx.p, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (sqdist) CK(cudaMemcpyAsync(sqdist, s.corr_d2.p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return B2ICP_OK;
}

int b2icp_nn_search(b2icp_handle* h, const float* q_xyzw, size_t n, int32_t* idx, float* sqdist) {
  if (!h || !idx) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (!gslot(h, 0).valid) return fail(h, B2ICP_ERR_NO_TARGET, "no target cloud set");
  if (n == 0) return B2ICP_OK;
  if (!q_xyzw) return fail(h, B2ICP_ERR_INVALID_ARG, "q_xyzw == NULL");
  CK(h->query.ensure(n * sizeof(float4)));
  CK(h->q_idx.ensure(n * sizeof(int)));
  CK(h->q_d2.ensure(n * sizeof(float)));
  CK(cudaMemcpyAsync(h->query.p, q_xyzw, n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  int rc = nn_search_impl(h, h->query.as<float4>(), n, h->q_idx.as<int>(), h->q_d2.as<float>());
  if (rc) return rc;
  CK(cudaMemcpyAsync(idx, h->q_idx.p, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (sqdist) CK(cudaMemcpyAsync(sqdist, h->q_d2.p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return B2ICP_OK;
}

int b2icp_nn_search_device(b2icp_handle* h, const float* d_q_xyzw, size_t n, int32_t* d_idx, float* d_sqdist) {
  if (!h || !d_idx || !d_sqdist) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (!gslot(h, 0).valid) return fail(h, B2ICP_ERR_NO_TARGET, "no target cloud set");
  if (n == 0) return B2ICP_OK;
  if (!d_q_xyzw) return fail(h, B2ICP_ERR_INVALID_ARG, "d_q_xyzw == NULL");
  const bool prof = h->params.profile != 0;
  if (prof) {
    while (h->events.size() < 2) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      h->events.push_back(e);
    }
    CK(cudaEventRecord(h->events[0], h->stream));
  }
  int rc = nn_search_impl(h, reinterpret_cast<const float4*>(d_q_xyzw), n, d_idx, d_sqdist);
  if (rc) return rc;
  if (prof) CK(cudaEventRecord(h->events[1], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  if (prof) {
    float ms = 0;
    cudaEventElapsedTime(&ms, h->events[0], h->events[1]);
    std::memset(&h->timing, 0, sizeof(h->timing));
    h->timing.nn_sweep_launches = 1;
    h->timing.nn_sweep_ms = ms;
    h->timing.total_ms = ms;
  }
  h->timing.kernel_launches = h->launches;
  return B2ICP_OK;
}

static int transform_impl(b2icp_handle* h, const float* in, size_t n, const void* T, bool dbl, float* out) {
  if (n == 0) return B2ICP_OK;
  if (!in || !T || !out) return fail(h, B2ICP_ERR_INVALID_ARG, "NULL argument");
  CK(h->xf_in.ensure(n * sizeof(float4)));
  CK(h->xf_out.ensure(n * sizeof(float4)));
  CK(cudaMemcpyAsync(h->xf_in.p, in, n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->mat.p, T, 16 * (dbl ? sizeof(double) : sizeof(float)), cudaMemcpyHostToDevice, h->stream));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (dbl)
    transform_cloud_d<<<blocks, 256, 0, h->stream>>>(h->xf_in.as<float4>(), (int)n, h->mat.as<double>(), h->xf_out.as<float4>());
  else
    transform_cloud_f<<<blocks, 256, 0, h->stream>>>(h->xf_in.as<float4>(), (int)n, h->mat.as<float>(), h->xf_out.as<float4>());
  h->launches += 1;
  CK(cudaMemcpyAsync(out, h->xf_out.p, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return B2ICP_OK;
}

int b2icp_transform_cloud(b2icp_handle* h, const float* in_xyzw, size_t n, const double* T, float* out_xyzw) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return transform_impl(h, in_xyzw, n, T, true, out_xyzw);
}
int b2icp_transform_cloud_f(b2icp_handle* h, const float* in_xyzw, size_t n, const float* T, float* out_xyzw) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return transform_impl(h, in_xyzw, n, T, false, out_xyzw);
}

int b2icp_align_batch(b2icp_handle* h, const float* const* src, const size_t* n_src, const float* const* tgt,
                      const size_t* n_tgt, size_t batch, int with_fitness, b2icp_result* out) {
  if (!h || !src || !n_src || !out) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return batch_impl(h, src, n_src, tgt, n_tgt, batch, with_fitness, out, false);
}

// ---- streaming batches: two slot sets, uploads of batch k+1 overlap the sweeps of batch k ----------------
static int wait_impl(b2icp_handle* h, b2icp_result* out, size_t capacity, size_t* n_out);
int b2icp_align_batch_submit(b2icp_handle* h, const float* const* src, const size_t* n_src, size_t batch, int with_fitness) {
  if (!h || !src || !n_src) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return submit_impl(h, src, n_src, batch, with_fitness, false);
}
int b2icp_align_batch_submit_device(b2icp_handle* h, const float* const* d_src, const size_t* n_src, size_t batch,
                                    int with_fitness) {
  if (!h || !d_src || !n_src) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return submit_impl(h, d_src, n_src, batch, with_fitness, true);
}
static int submit_impl(b2icp_handle* h, const float* const* src, const size_t* n_src, size_t batch, int with_fitness,
                       bool from_device) {
  if (h->params.mode != B2ICP_MODE_P2P_SVD) return fail(h, B2ICP_ERR_INVALID_ARG, "streaming batches run the point-to-point mode only");
  if (batch == 0 || batch > (size_t)kSetSlots) return fail(h, B2ICP_ERR_INVALID_ARG, "a streamed batch holds 1..32 scans");
  if (!gslot(h, 0).valid) return fail(h, B2ICP_ERR_NO_TARGET, "no target cloud set");
  if (h->n_pending >= h->max_in_flight) return fail(h, B2ICP_ERR_INVALID_ARG, "too many batches in flight: call b2icp_align_batch_wait");
  if (!h->copy_stream) {
    CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < kStreamSets; ++k) {
      CK(cudaEventCreateWithFlags(&h->set_uploaded[k], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->set_done[k], cudaEventDisableTiming));
      CK(cudaStreamCreateWithFlags(&h->set_stream[k], cudaStreamNonBlocking));
    }
  }
  const int set = h->next_set;
  const int slot0 = set * kSetSlots;
  const int B = (int)batch;
  h->aligned = false;
  // uploads go on the batch's own stream when it has one (measured: 16.2k scans/s end to end, against 12.9k with
  // a shared copy stream + event, whose copies did not overlap the other sets' sweeps)
  cudaStream_t const up = with_fitness ? h->copy_stream : h->set_stream[set];
  for (int i = 0; i < B; ++i) {
    ScanSlot& s = slot(h, (size_t)(slot0 + i));
    s.grid = 0;
    if (from_device) {  // the caller's device cloud is read in place: it must stay valid until the batch's _wait
      if (!src[i] || n_src[i] == 0) return fail(h, B2ICP_ERR_EMPTY_CLOUD, "empty cloud");
      if (n_src[i] > (size_t)INT32_MAX / 8) return fail(h, B2ICP_ERR_INVALID_ARG, "cloud too large");
      s.src.ext = reinterpret_cast<const float4*>(src[i]);
      s.src.n = n_src[i];
      s.src.valid = true;
      continue;
    }
    int rc = upload_cloud(h, s.src, src[i], n_src[i], from_device, up);
    if (rc) {
      cudaStreamSynchronize(up);
      return rc;
    }
  }
  CK(cudaEventRecord(h->set_uploaded[set], up));
  // The batch runs on its set's own stream (fitness shares scratch buffers between the sets, so a batch that asks
  // for it stays on the handle's stream, behind the other set).  Everything below is stream-ordered on `cs`.
  const bool own = !with_fitness && getenv("B2ICP_ONE_STREAM") == nullptr;
  cudaStream_t const saved = h->stream;
  cudaStream_t const cs = own ? h->set_stream[set] : saved;
  if (!own)  // behind every batch still in flight
    for (int k = 0; k < h->n_pending; ++k) CK(cudaStreamWaitEvent(cs, h->set_done[h->pending[(h->first_pending + k) % kStreamSets].set], 0));
  CK(cudaStreamWaitEvent(cs, h->set_uploaded[set], 0));
  h->stream = cs;
  int rc = run_batch(h, B, nullptr, slot0, false, true);
  if (!rc && with_fitness)
    for (int i = 0; i < B && !rc; ++i) rc = enqueue_fitness(h, slot0 + i, DBL_MAX);
  h->stream = saved;
  if (rc) return rc;
  if (h->sink && h->sink_used + (size_t)B <= h->sink_cap) {
    export_records<<<1, 64, 0, cs>>>(h->states.as<IcpState>() + slot0, B, with_fitness, h->sink + h->sink_used);
    h->sink_used += (size_t)B;
    h->launches += 1;
  }
  CK(cudaMemcpyAsync(h->h_states + slot0, h->states.as<IcpState>() + slot0, sizeof(IcpState) * B, cudaMemcpyDeviceToHost, cs));
  CK(cudaEventRecord(h->set_done[set], cs));
  h->pending[(h->first_pending + h->n_pending) % kStreamSets] = {set, B, with_fitness};
  h->n_pending += 1;
  h->next_set = (h->next_set + 1) % kStreamSets;
  return B2ICP_OK;
}

int b2icp_align_batch_wait(b2icp_handle* h, b2icp_result* out, size_t capacity, size_t* n_out) {
  if (!h || !out) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return wait_impl(h, out, capacity, n_out);
}
static int wait_impl(b2icp_handle* h, b2icp_result* out, size_t capacity, size_t* n_out) {
  if (n_out) *n_out = 0;
  if (h->n_pending == 0) return fail(h, B2ICP_ERR_INVALID_ARG, "no streamed batch in flight");
  const b2icp_handle::Pending pd = h->pending[h->first_pending];
  if (capacity < (size_t)pd.B) return fail(h, B2ICP_ERR_INVALID_ARG, "result buffer too small");
  CK(cudaEventSynchronize(h->set_done[pd.set]));
  CK(cudaGetLastError());
  h->first_pending = (h->first_pending + 1) % kStreamSets;
  h->n_pending -= 1;
  const int slot0 = pd.set * kSetSlots;
  int worst = B2ICP_OK;
  for (int i = 0; i < pd.B; ++i) {
    const IcpState& st = h->h_states[slot0 + i];
    fill_result(st, &out[i]);
    if (pd.with_fitness && st.status == 0) out[i].fitness = fitness_value(st);
    if (st.status != 0 && worst == B2ICP_OK) {
      worst = st.status;
      h->err = status_message(worst);
    }
  }
  h->timing.kernel_launches = h->launches;
  if (n_out) *n_out = (size_t)pd.B;
  return worst;
}

int b2icp_set_record_sink(b2icp_handle* h, b2icp_record* d_records, size_t capacity) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->n_pending) return fail(h, B2ICP_ERR_INVALID_ARG, "record sink changed while batches are in flight");
  h->sink = d_records;
  h->sink_cap = d_records ? capacity : 0;
  h->sink_used = 0;
  return B2ICP_OK;
}
int b2icp_record_sink_count(b2icp_handle* h, size_t* n) {
  if (!h || !n) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  *n = h->sink_used;
  return B2ICP_OK;
}

int b2icp_align_batch_device(b2icp_handle* h, const float* const* d_src, const size_t* n_src,
                             const float* const* d_tgt, const size_t* n_tgt, size_t batch, int with_fitness,
                             b2icp_result* out) {
  if (!h || !d_src || !n_src || !out) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return batch_impl(h, d_src, n_src, d_tgt, n_tgt, batch, with_fitness, out, true);
}

int b2icp_compute_covariances(b2icp_handle* h, const float* xyzw, size_t n, double* cov9) {
  if (!h || !cov9) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  GridSlot& g = gslot(h, kMaxBatch + 2);  // scratch grid: leaves the handle's target untouched
  int rc = upload_cloud(h, g.tgt, xyzw, n, false);
  if (rc) return rc;
  g.pts = g.tgt.raw.as<float4>();
  GridSlot* gp = &g;
  rc = build_grids(h, &gp, &n, 1);
  if (rc) return rc;
  rc = compute_covariances(h, g, n, g.cov);
  if (rc) return rc;
  CK(cudaMemcpyAsync(cov9, g.cov.p, n * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return B2ICP_OK;
}

int b2icp_voxel_filter(b2icp_handle* h, const float* in_xyzw, size_t n, float leaf, float* out_xyzw, size_t* n_out) {
  if (!h || !n_out) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  *n_out = 0;
  if (n == 0) return B2ICP_OK;
  if (!in_xyzw || !out_xyzw || !(leaf > 0)) return fail(h, B2ICP_ERR_INVALID_ARG, "voxel_filter: bad argument");
  GridSlot& g = gslot(h, kMaxBatch + 3);  // scratch slot
  int rc = upload_cloud(h, g.tgt, in_xyzw, n, false);
  if (rc) return rc;
  const float4* pts = g.tgt.raw.as<float4>();
  CK(g.bbox.ensure(sizeof(BBox)));
  const int blocks = (int)((n + 255) / 256);
  bbox_init<<<1, 32, 0, h->stream>>>(g.bbox.as<BBox>());
  bbox_kernel<<<std::min(blocks, 148 * 8), 256, 0, h->stream>>>(pts, (int)n, g.bbox.as<BBox>());
  CK(cudaMemcpyAsync(&h->h_bbox[0], g.bbox.p, sizeof(BBox), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_bbox[0].nonfinite) return fail(h, B2ICP_ERR_NONFINITE_INPUT, "voxel_filter: non-finite coordinates");
  VoxelParams vp;
  vp.inv_leaf = 1.0f / leaf;
  long long cells = 1;
  for (int d = 0; d < 3; ++d) {
    const float mn = ord2f(h->h_bbox[0].mn[d]), mx = ord2f(h->h_bbox[0].mx[d]);
    vp.min_b[d] = (int)std::floor(mn * vp.inv_leaf);
    const int max_b = (int)std::floor(mx * vp.inv_leaf);
    vp.div[d] = max_b - vp.min_b[d] + 1;
    cells *= (long long)vp.div[d];
    if (cells > (long long)INT32_MAX) {
      // pcl::VoxelGrid: "Leaf size is too small for the input dataset. Integer indices would overflow." -> output = input
      std::memcpy(out_xyzw, in_xyzw, n * sizeof(float4));
      *n_out = n;
      return B2ICP_OK;
    }
  }
  const int ncell = (int)cells;
  // A dense per-voxel table pays while voxels are about as many as points.  A fine leaf over a wide scan (the
  // reference's code default, 0.05 m, over a 60 m x 60 m x 10 m sweep is 2.9e8 voxels) would allocate, clear and scan
  // gigabytes per call: such boxes take the SPARSE path — what pcl::VoxelGrid itself does — a stable radix sort of
  // (voxel index, point index), leaders where the key changes, centroids over the runs.  Same bits as the dense path.
  if ((size_t)ncell > 8 * n + 4096 || ncell > kMaxCells || getenv("B2ICP_VOXEL_SPARSE")) {
    const int N = (int)n;
    const int nchunk = (N + kSortChunk - 1) / kSortChunk, nh = kSortBins * nchunk;
    const int stiles = (nh + kScanTile - 1) / kScanTile, ftiles = (N + kScanTile - 1) / kScanTile;
    CK(g.cell_of.ensure((n + 8) * 4));
    CK(g.rank.ensure((n + 8) * 4));
    CK(g.sorted.ensure((n + 8) * 8));  // second (key, value) buffer pair
    CK(g.cell_start.ensure(((size_t)nh + 8) * sizeof(int)));
    CK(g.tile_sums.ensure((size_t)std::max(stiles, ftiles) * sizeof(int)));
    CK(h->xf_out.ensure(n * sizeof(float4)));
    CK(h->q_idx.ensure((n + 8) * sizeof(int)));
    unsigned int* k0 = g.cell_of.as<unsigned int>();
    unsigned int* v0 = g.rank.as<unsigned int>();
    unsigned int* k1 = g.sorted.as<unsigned int>();
    unsigned int* v1 = k1 + (n + 8);
    int* hist = g.cell_start.as<int>();
    int* flags = h->q_idx.as<int>();
    static bool configured = false;
    if (!configured) {
      CK(cudaFuncSetAttribute(sort_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSortWarps * kSortBins * sizeof(int))));
      configured = true;
    }
    voxel_sparse_keys<<<blocks, 256, 0, h->stream>>>(pts, N, vp, k0, v0);
    int bits = 1;
    while (bits < 31 && (1ll << bits) < cells) ++bits;
    for (int shift = 0; shift < bits; shift += kSortBits) {
      sort_hist<<<nchunk, kSortThreads, 0, h->stream>>>(k0, N, shift, nchunk, hist);
      scan_tile_sums<<<stiles, kScanThreads, 0, h->stream>>>(hist, nh, g.tile_sums.as<int>(), g.bbox.as<BBox>());
      scan_of_sums<<<1, kScanThreads, 0, h->stream>>>(g.tile_sums.as<int>(), stiles);
      scan_apply<<<stiles, kScanThreads, 0, h->stream>>>(hist, nh, g.tile_sums.as<int>(), N);
      sort_scatter<<<nchunk, kSortThreads, kSortWarps * kSortBins * sizeof(int), h->stream>>>(k0, v0, N, shift, nchunk, hist, k1, v1);
      h->launches += 5;
      std::swap(k0, k1);
      std::swap(v0, v1);
    }
    voxel_sparse_flags<<<blocks, 256, 0, h->stream>>>(k0, N, flags);
    bbox_init<<<1, 32, 0, h->stream>>>(g.bbox.as<BBox>());
    scan_tile_sums<<<ftiles, kScanThreads, 0, h->stream>>>(flags, N, g.tile_sums.as<int>(), g.bbox.as<BBox>());
    scan_of_sums<<<1, kScanThreads, 0, h->stream>>>(g.tile_sums.as<int>(), ftiles);
    CK(cudaMemcpyAsync(&h->h_bbox[1], g.bbox.p, sizeof(BBox), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const int n_vox = h->h_bbox[1].occupied;
    scan_apply<<<ftiles, kScanThreads, 0, h->stream>>>(flags, N, g.tile_sums.as<int>(), n_vox);
    voxel_sparse_centroids<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(pts, k0, v0, N, flags, h->xf_out.as<float4>());
    h->launches += 7;
    CK(cudaMemcpyAsync(out_xyzw, h->xf_out.p, (size_t)n_vox * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    *n_out = (size_t)n_vox;
    g.valid = false;
    return B2ICP_OK;
  }
  CK(g.cell_start.ensure(((size_t)ncell + 8) * sizeof(int)));
  CK(g.sorted.ensure(n * sizeof(float4)));
  CK(g.cell_of.ensure((n + 8) * sizeof(int)));
  CK(g.rank.ensure(n * sizeof(int)));
  CK(h->xf_out.ensure(n * sizeof(float4)));
  CK(h->q_idx.ensure((n + 8) * sizeof(int)));  // leader flags -> output slots
  const int ntiles = (ncell + kScanTile - 1) / kScanTile, ftiles = ((int)n + kScanTile - 1) / kScanTile;
  CK(g.tile_sums.ensure((size_t)std::max(ntiles, ftiles) * sizeof(int)));
  int* cs = g.cell_start.as<int>();
  int* flags = h->q_idx.as<int>();
  CK(cudaMemsetAsync(cs, 0, ((size_t)ncell + 1) * sizeof(int), h->stream));
  voxel_count<<<blocks, 256, 0, h->stream>>>(pts, (int)n, vp, g.cell_of.as<int>(), g.rank.as<int>(), cs);
  scan_tile_sums<<<ntiles, kScanThreads, 0, h->stream>>>(cs, ncell, g.tile_sums.as<int>(), g.bbox.as<BBox>());
  scan_of_sums<<<1, kScanThreads, 0, h->stream>>>(g.tile_sums.as<int>(), ntiles);
  scan_apply<<<ntiles, kScanThreads, 0, h->stream>>>(cs, ncell, g.tile_sums.as<int>(), (int)n);
  grid_scatter<<<blocks, 256, 0, h->stream>>>(pts, (int)n, g.cell_of.as<int>(), g.rank.as<int>(), cs, g.sorted.as<float4>());
  voxel_flag_leaders<<<blocks, 256, 0, h->stream>>>((int)n, g.cell_of.as<int>(), g.rank.as<int>(), cs, flags);
  bbox_init<<<1, 32, 0, h->stream>>>(g.bbox.as<BBox>());  // reset the non-zero counter: it now counts leaders
  scan_tile_sums<<<ftiles, kScanThreads, 0, h->stream>>>(flags, (int)n, g.tile_sums.as<int>(), g.bbox.as<BBox>());
  scan_of_sums<<<1, kScanThreads, 0, h->stream>>>(g.tile_sums.as<int>(), ftiles);
  CK(cudaMemcpyAsync(&h->h_bbox[1], g.bbox.p, sizeof(BBox), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const int n_vox = h->h_bbox[1].occupied;  // leaders = occupied voxels = output points
  scan_apply<<<ftiles, kScanThreads, 0, h->stream>>>(flags, (int)n, g.tile_sums.as<int>(), n_vox);
  voxel_centroids<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(g.sorted.as<float4>(), (int)n, vp, cs, flags,
                                                                      h->xf_out.as<float4>());
  h->launches += 12;
  CK(cudaMemcpyAsync(out_xyzw, h->xf_out.p, (size_t)n_vox * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  *n_out = (size_t)n_vox;
  g.valid = false;
  return B2ICP_OK;
}

// ---- K9: device-resident point map (map.cuh) ---------------------------------------------------------
namespace {
constexpr size_t kMapGrid = kMaxBatch + 4;  // grid slot that indexes the map for b2icp_map_nearest

MapTable map_table(b2icp_handle* h) {
  MapTable t;
  t.keys = h->map_keys.as<unsigned long long>();
  t.vals = h->map_vals.as<int>();
  t.mask = (unsigned int)(h->map_table_cap - 1);
  t.inv_res = 1.0 / h->map_resolution;
  t.compat = h->map_compat ? 1 : 0;
  t.res = h->map_resolution;
  for (int d = 0; d < 3; ++d) t.org[d] = h->map_org[d];
  return t;
}

// ---- PCL-compatible mode: the root box exactly as pcl::octree::OctreePointCloud grows it ------
// first point: box = p +- resolution / 2, which getKeyBitSize() widens to a depth-1 tree with the point at its centre
void octree_define_box(OctreeBox& b, const float* p) {
  const double res = b.res;
  const float minValue = FLT_EPSILON;
  for (int d = 0; d < 3; ++d) {
    b.mn[d] = (double)p[d] - res / 2;
    b.mx[d] = (double)p[d] + res / 2;
  }
  unsigned mk[3];
  for (int d = 0; d < 3; ++d) mk[d] = (unsigned)std::ceil((b.mx[d] - b.mn[d] - minValue) / res);
  const unsigned max_voxels = std::max(std::max(std::max(mk[0], mk[1]), mk[2]), 2u);
  b.depth = (int)std::max(std::min(32u, (unsigned)std::ceil(std::log((double)max_voxels) / std::log(2.0) - minValue)), 0u);
  const double side = (double)(1u << b.depth) * res;
  for (int d = 0; d < 3; ++d) {
    const double over = (side - (b.mx[d] - b.mn[d])) / 2.0;
    if (over > minValue) {
      b.mn[d] -= over;
      b.mx[d] += over;
    }
  }
  b.defined = 1;
}
// adoptBoundingBoxToPoint for a point outside the box: new roots until it fits.  Per axis the old tree becomes the
// lower child iff that axis' UPPER bound is violated; otherwise the box grows towards minus.
void octree_grow_box(OctreeBox& b, const float* p) {
  const float minValue = FLT_EPSILON;
  for (;;) {
    bool hi[3], any = false;
    for (int d = 0; d < 3; ++d) {
      hi[d] = (double)p[d] >= b.mx[d];
      any = any || hi[d] || (double)p[d] < b.mn[d];
    }
    if (!any) return;
    double side = (double)(1u << b.depth) * b.res;
    for (int d = 0; d < 3; ++d)
      if (!hi[d]) b.mn[d] -= side;
    b.depth += 1;
    side = (double)(1u << b.depth) * b.res - minValue;
    for (int d = 0; d < 3; ++d) b.mx[d] = b.mn[d] + side;
  }
}

// smallest index >= cursor of a finite point of pts[0, n) outside the current box (n: none); the point in h->h_point
int octree_next_outside(b2icp_handle* h, const float4* pts, size_t n, int cursor, int* found) {
  CK(h->map_first.ensure(sizeof(int)));
  *h->h_first = INT32_MAX;
  CK(cudaMemcpyAsync(h->map_first.p, h->h_first, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  const int blocks = (int)std::min<size_t>((n - (size_t)cursor + 255) / 256, 148 * 8);
  octree_first_outside<<<std::max(blocks, 1), 256, 0, h->stream>>>(pts, (int)n, cursor, h->map_box, h->map_first.as<int>());
  h->launches += 1;
  CK(cudaMemcpyAsync(h->h_first, h->map_first.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *found = *h->h_first;
  if (*found != INT32_MAX) {
    CK(cudaMemcpyAsync(h->h_point, pts + *found, 4 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return B2ICP_OK;
}

// the (level, key) prefix table of the current map under the current box
int octree_ensure_tree(b2icp_handle* h) {
  if (h->tree_valid) return B2ICP_OK;
  const size_t n = h->map_size;
  size_t entries = 0, pow8 = 1;
  for (int level = 1; level <= h->map_box.depth; ++level) {
    pow8 = pow8 > n ? pow8 : pow8 * 8;
    entries += std::min(n, pow8);
  }
  size_t cap = 1024;
  while (cap < 2 * entries) cap *= 2;
  if (cap > (1ull << 31)) return fail(h, B2ICP_ERR_INVALID_ARG, "map too large for the octree table");
  CK(h->tree_keys.ensure(cap * sizeof(unsigned long long)));
  CK(h->tree_vals.ensure(cap * sizeof(int)));
  h->tree_cap = cap;
  CK(cudaMemsetAsync(h->tree_keys.p, 0xFF, cap * sizeof(unsigned long long), h->stream));
  CK(cudaMemsetAsync(h->tree_vals.p, 0x7F, cap * sizeof(int), h->stream));
  MapTable t{};
  t.keys = h->tree_keys.as<unsigned long long>();
  t.vals = h->tree_vals.as<int>();
  t.mask = (unsigned int)(cap - 1);
  octree_build<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->map_pts.as<float4>(), (int)n, h->map_box, t);
  h->launches += 1;
  h->tree_valid = true;
  return B2ICP_OK;
}

int map_ensure_grid(b2icp_handle* h);

// the map point OctreeMapper::approxNearestNeighbors pairs with every query: exact nearest neighbour (default), or
// PCL's greedy octree descent in compat mode
int map_nearest_impl(b2icp_handle* h, const float4* d_q, size_t n, int* d_idx, float* d_d2) {
  if (!h->map_compat) {
    int rc = map_ensure_grid(h);
    if (rc) return rc;
    return nn_search_impl(h, d_q, n, d_idx, d_d2, (int)(kMaxBatch + 4));
  }
  if (h->map_size == 0) return fail(h, B2ICP_ERR_NO_TARGET, "the map is empty");
  int rc = octree_ensure_tree(h);
  if (rc) return rc;
  MapTable t{};
  t.keys = h->tree_keys.as<unsigned long long>();
  t.vals = h->tree_vals.as<int>();
  t.mask = (unsigned int)(h->tree_cap - 1);
  octree_approx_nearest<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d_q, (int)n, h->map_box, t, d_idx);
  h->launches += 1;
  (void)d_d2;
  return B2ICP_OK;
}

// table able to hold `want` voxels at load <= 0.5; growth re-enters the voxels of the current map
int map_reserve(b2icp_handle* h, size_t want) {
  size_t cap = std::max<size_t>(h->map_table_cap, 1024);
  while (cap < 2 * want) cap *= 2;
  if (cap == h->map_table_cap) return B2ICP_OK;
  if (cap > (1ull << 31)) return fail(h, B2ICP_ERR_INVALID_ARG, "map too large");
  h->map_keys.release();
  h->map_vals.release();
  CK(h->map_keys.ensure(cap * sizeof(unsigned long long)));
  CK(h->map_vals.ensure(cap * sizeof(int)));
  h->map_table_cap = cap;
  CK(cudaMemsetAsync(h->map_keys.p, 0xFF, cap * sizeof(unsigned long long), h->stream));
  CK(cudaMemsetAsync(h->map_vals.p, 0x7F, cap * sizeof(int), h->stream));
  if (h->map_size) {
    map_rehash<<<(unsigned)((h->map_size + 255) / 256), 256, 0, h->stream>>>(h->map_pts.as<float4>(), (int)h->map_size,
                                                                         map_table(h));
    h->launches += 1;
  }
  return B2ICP_OK;
}

// exclusive scan of flags[0..n) in place, flags[n] = total; returns the total through pinned memory
int scan_flags(b2icp_handle* h, int* flags, size_t n, int* total) {
  const int tiles = (int)((n + kScanTile - 1) / kScanTile);
  CK(h->map_tiles.ensure((size_t)tiles * sizeof(int)));
  CK(h->map_stats.ensure(sizeof(BBox)));
  bbox_init<<<1, 32, 0, h->stream>>>(h->map_stats.as<BBox>());
  scan_tile_sums<<<tiles, kScanThreads, 0, h->stream>>>(flags, (int)n, h->map_tiles.as<int>(), h->map_stats.as<BBox>());
  scan_of_sums<<<1, kScanThreads, 0, h->stream>>>(h->map_tiles.as<int>(), tiles);
  CK(cudaMemcpyAsync(&h->h_bbox[0], h->map_stats.p, sizeof(BBox), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *total = h->h_bbox[0].occupied;  // flags are 0/1: non-zero entries = their sum
  scan_apply<<<tiles, kScanThreads, 0, h->stream>>>(flags, (int)n, h->map_tiles.as<int>(), *total);
  h->launches += 4;
  return B2ICP_OK;
}

int map_insert_impl(b2icp_handle* h, const float* xyzw, size_t n, bool from_device, size_t* n_added) {
  if (n_added) *n_added = 0;
  if (!(h->map_resolution > 0)) return fail(h, B2ICP_ERR_INVALID_ARG, "b2icp_map_reset has not been called");
  if (n == 0) return B2ICP_OK;
  if (!xyzw) return fail(h, B2ICP_ERR_INVALID_ARG, "xyzw == NULL");
  if (n > (size_t)INT32_MAX / 8 || h->map_size + n > (size_t)INT32_MAX / 8) return fail(h, B2ICP_ERR_INVALID_ARG, "map too large");
  int rc = map_reserve(h, h->map_size + n);
  if (rc) return rc;
  if (h->map_pts.cap < (h->map_size + n) * sizeof(float4)) {  // grow, keeping the points
    DeviceBuf bigger;
    CK(bigger.ensure((h->map_size + n) * 2 * sizeof(float4)));
    if (h->map_size) CK(cudaMemcpyAsync(bigger.p, h->map_pts.p, h->map_size * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->map_pts.release();
    h->map_pts = bigger;
    h->map_grid_valid = false;
  }
  const float4* pts = reinterpret_cast<const float4*>(xyzw);
  if (!from_device) {
    CK(h->xf_in.ensure(n * sizeof(float4)));
    CK(cudaMemcpyAsync(h->xf_in.p, xyzw, n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    pts = h->xf_in.as<float4>();
  }
  CK(h->map_slot_of.ensure(n * sizeof(int)));
  CK(h->map_flags.ensure((n + 8) * sizeof(int)));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (h->map_compat && !h->map_box.defined) {  // PCL anchors its lattice on the first point it is given
    int first = INT32_MAX;
    rc = octree_next_outside(h, pts, n, 0, &first);
    if (rc) return rc;
    if (first == INT32_MAX) return B2ICP_OK;  // no finite point in the cloud
    octree_define_box(h->map_box, h->h_point);
    for (int d = 0; d < 3; ++d) h->map_org[d] = h->map_box.mn[d];
    h->tree_valid = false;
  }
  const MapTable t = map_table(h);
  map_claim<<<blocks, 256, 0, h->stream>>>(pts, (int)n, t, h->map_slot_of.as<int>());
  map_flag<<<blocks, 256, 0, h->stream>>>((int)n, t, h->map_slot_of.as<int>(), h->map_flags.as<int>());
  int added = 0;
  rc = scan_flags(h, h->map_flags.as<int>(), n, &added);
  if (rc) return rc;
  map_append<<<blocks, 256, 0, h->stream>>>(pts, (int)n, t, h->map_slot_of.as<int>(), h->map_flags.as<int>(),
                                            h->map_pts.as<float4>(), (int)h->map_size);
  h->launches += 3;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  h->map_size += (size_t)added;
  if (added) h->map_grid_valid = false;
  if (added) h->tree_valid = false;
  if (n_added) *n_added = (size_t)added;
  if (h->map_compat) {  // the growth events of this call, in input order: every point outside the box adds roots
    for (int cursor = 0; cursor < (int)n;) {
      int first = INT32_MAX;
      rc = octree_next_outside(h, pts, n, cursor, &first);
      if (rc) return rc;
      if (first == INT32_MAX) break;
      octree_grow_box(h->map_box, h->h_point);
      if (h->map_box.depth > 19) return fail(h, B2ICP_ERR_INVALID_ARG, "octree deeper than 19 levels");
      h->tree_valid = false;
      cursor = first + 1;
    }
  }
  return B2ICP_OK;
}

int map_ensure_grid(b2icp_handle* h) {
  if (h->map_size == 0) return fail(h, B2ICP_ERR_NO_TARGET, "the map is empty");
  GridSlot& g = gslot(h, kMapGrid);
  if (h->map_grid_valid && g.valid) return B2ICP_OK;
  g.pts = h->map_pts.as<float4>();
  GridSlot* gp = &g;
  size_t n = h->map_size;
  int rc = build_grids(h, &gp, &n, 1);
  if (rc) return rc;
  h->map_grid_valid = true;
  return B2ICP_OK;
}
}  // namespace

static int map_reset_impl(b2icp_handle* h, double resolution, bool compat) {
  if (!(resolution > 0) || !std::isfinite(resolution)) return fail(h, B2ICP_ERR_INVALID_ARG, "resolution must be > 0");
  h->map_resolution = resolution;
  h->map_size = 0;
  h->map_grid_valid = false;
  h->map_compat = compat;
  std::memset(&h->map_box, 0, sizeof(h->map_box));
  h->map_box.res = resolution;
  h->map_org[0] = h->map_org[1] = h->map_org[2] = 0.0;
  h->tree_valid = false;
  if (h->map_table_cap) {
    CK(cudaMemsetAsync(h->map_keys.p, 0xFF, h->map_table_cap * sizeof(unsigned long long), h->stream));
    CK(cudaMemsetAsync(h->map_vals.p, 0x7F, h->map_table_cap * sizeof(int), h->stream));
  }
  return B2ICP_OK;
}

int b2icp_map_reset(b2icp_handle* h, double resolution) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return map_reset_impl(h, resolution, false);
}

int b2icp_map_reset_octree(b2icp_handle* h, double resolution) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return map_reset_impl(h, resolution, true);
}

int b2icp_map_insert(b2icp_handle* h, const float* xyzw, size_t n, size_t* n_added) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return map_insert_impl(h, xyzw, n, false, n_added);
}

int b2icp_map_insert_device(b2icp_handle* h, const float* d_xyzw, size_t n, size_t* n_added) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return map_insert_impl(h, d_xyzw, n, true, n_added);
}

int b2icp_map_size(b2icp_handle* h, size_t* n) {
  if (!h || !n) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  *n = h->map_size;
  return B2ICP_OK;
}

int b2icp_map_download(b2icp_handle* h, float* out_xyzw, size_t capacity, size_t* n) {
  if (!h || !n) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  *n = h->map_size;
  if (!out_xyzw || h->map_size == 0) return B2ICP_OK;
  if (capacity < h->map_size) return fail(h, B2ICP_ERR_INVALID_ARG, "output buffer too small");
  CK(cudaMemcpyAsync(out_xyzw, h->map_pts.p, h->map_size * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return B2ICP_OK;
}

int b2icp_map_nearest(b2icp_handle* h, const float* q_xyzw, size_t n, int32_t* idx, float* nn_xyzw, size_t* n_nn) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (n_nn) *n_nn = 0;
  if (n == 0) return B2ICP_OK;
  if (!q_xyzw) return fail(h, B2ICP_ERR_INVALID_ARG, "q_xyzw == NULL");
  if (h->map_size == 0) return fail(h, B2ICP_ERR_NO_TARGET, "the map is empty");
  CK(h->query.ensure(n * sizeof(float4)));
  CK(h->q_idx.ensure((n + 8) * sizeof(int)));
  CK(h->q_d2.ensure(n * sizeof(float)));
  CK(cudaMemcpyAsync(h->query.p, q_xyzw, n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  int rc = map_nearest_impl(h, h->query.as<float4>(), n, h->q_idx.as<int>(), h->q_d2.as<float>());
  if (rc) return rc;
  if (idx) CK(cudaMemcpyAsync(idx, h->q_idx.p, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (nn_xyzw) {
    CK(h->map_flags.ensure((n + 8) * sizeof(int)));
    CK(h->xf_out.ensure(n * sizeof(float4)));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    map_gather_flag<<<blocks, 256, 0, h->stream>>>(h->q_idx.as<int>(), (int)n, h->map_flags.as<int>());
    int found = 0;
    rc = scan_flags(h, h->map_flags.as<int>(), n, &found);
    if (rc) return rc;
    map_gather<<<blocks, 256, 0, h->stream>>>(h->q_idx.as<int>(), (int)n, h->map_flags.as<int>(), h->map_pts.as<float4>(),
                                              h->xf_out.as<float4>());
    h->launches += 2;
    if (found) CK(cudaMemcpyAsync(nn_xyzw, h->xf_out.p, (size_t)found * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    if (n_nn) *n_nn = (size_t)found;
  }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return B2ICP_OK;
}

// OctreeMapper::refineTransformAndGrowMap (octree_mapper.cpp:133-173) with the scan uploaded once and every
// intermediate cloud left in device memory.
int b2icp_mapper_register(b2icp_handle* h, const float* xyzw, size_t n, const float* T_raw, const float* T_raw_inv,
                          b2icp_result* out) {
  if (!h || !out) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  identity_result(out);
  if (!xyzw || n == 0) return fail(h, B2ICP_ERR_EMPTY_CLOUD, "empty cloud");
  if (!T_raw || !T_raw_inv) return fail(h, B2ICP_ERR_INVALID_ARG, "NULL pose matrix");
  if (h->map_size == 0) return fail(h, B2ICP_ERR_NO_TARGET, "the map is empty");
  h->aligned = false;
  ScanSlot& s = slot(h, 0);
  s.grid = 0;
  int rc = upload_cloud(h, s.src, xyzw, n, false);  // icp.setInputSource(curr_cloud), kept for b2icp_mapper_grow
  if (rc) return rc;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  CK(h->query.ensure(n * sizeof(float4)));
  CK(h->q_idx.ensure((n + 8) * sizeof(int)));
  CK(h->q_d2.ensure(n * sizeof(float)));
  CK(h->map_flags.ensure((n + 8) * sizeof(int)));
  CK(h->xf_in.ensure(n * sizeof(float4)));
  CK(h->xf_out.ensure(n * sizeof(float4)));
  CK(h->mat.ensure(32 * sizeof(float)));
  CK(cudaMemcpyAsync(h->mat.p, T_raw, 16 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->mat.as<float>() + 16, T_raw_inv, 16 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  // line 136: cloud_in_map = raw_pose (x) cloud
  transform_cloud_f<<<blocks, 256, 0, h->stream>>>(s.src.raw.as<float4>(), (int)n, h->mat.as<float>(), h->query.as<float4>());
  // lines 145 / 73-90: the map point nearest to every scan point, compacted in scan order
  rc = map_nearest_impl(h, h->query.as<float4>(), n, h->q_idx.as<int>(), h->q_d2.as<float>());
  if (rc) return rc;
  map_gather_flag<<<blocks, 256, 0, h->stream>>>(h->q_idx.as<int>(), (int)n, h->map_flags.as<int>());
  int found = 0;
  rc = scan_flags(h, h->map_flags.as<int>(), n, &found);
  if (rc) return rc;
  if (found == 0) return fail(h, B2ICP_ERR_NO_TARGET, "no map point near the scan");
  map_gather<<<blocks, 256, 0, h->stream>>>(h->q_idx.as<int>(), (int)n, h->map_flags.as<int>(), h->map_pts.as<float4>(),
                                            h->xf_in.as<float4>());
  // line 149: nn_cloud = raw_pose^-1 (x) nn_cloud_in_map
  transform_cloud_f<<<(unsigned)((found + 255) / 256), 256, 0, h->stream>>>(h->xf_in.as<float4>(), found, h->mat.as<float>() + 16,
                                                                          h->xf_out.as<float4>());
  h->launches += 4;
  // lines 104-117: icp(source = cloud, target = nn_cloud)
  rc = set_target_impl(h, 0, h->xf_out.as<float>(), (size_t)found, true);
  if (rc) return rc;
  const bool gicp = h->params.mode == B2ICP_MODE_GICP_BFGS;
  rc = gicp ? run_gicp(h, nullptr) : run_batch(h, 1, nullptr);
  if (rc) return rc;
  if (gicp) {
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
  } else {
    rc = read_states(h, 1);
    if (rc) return rc;
  }
  fill_result(h->h_states[0], out);
  h->aligned = true;
  if (h->h_states[0].status != 0) {
    h->err = status_message(h->h_states[0].status);
    return h->h_states[0].status;
  }
  return B2ICP_OK;
}

int b2icp_mapper_grow(b2icp_handle* h, const float* xyzw, size_t n, const float* T, size_t* n_added) {
  if (!h || !T) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (n_added) *n_added = 0;
  ScanSlot& s = slot(h, 0);
  if (xyzw) {
    int rc = upload_cloud(h, s.src, xyzw, n, false);
    if (rc) return rc;
    h->aligned = false;
  } else if (!s.src.valid) {
    return fail(h, B2ICP_ERR_NO_SOURCE, "no scan retained: pass the cloud or call b2icp_mapper_register first");
  }
  const size_t m = s.src.n;
  CK(h->xf_out.ensure(m * sizeof(float4)));
  CK(h->mat.ensure(32 * sizeof(float)));
  CK(cudaMemcpyAsync(h->mat.p, T, 16 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  // lines 157-158: cloud_in_map = refined_pose (x) cloud; addPointsToMap(cloud_in_map)
  transform_cloud_f<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(s.src.raw.as<float4>(), (int)m, h->mat.as<float>(),
                                                                      h->xf_out.as<float4>());
  h->launches += 1;
  return map_insert_impl(h, h->xf_out.as<float>(), m, true, n_added);
}

int b2icp_set_target_map(b2icp_handle* h) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (h->map_size == 0) return fail(h, B2ICP_ERR_NO_TARGET, "the map is empty");
  return set_target_impl(h, 0, h->map_pts.as<float>(), h->map_size, true);
}

int b2icp_pointcloud2_to_xyzw(b2icp_handle* h, const uint8_t* data, size_t data_bytes, uint32_t width, uint32_t height,
                              uint32_t point_step, uint32_t row_step, uint32_t off_x, uint32_t off_y, uint32_t off_z,
                              int is_bigendian, float* out_xyzw) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  const size_t n = (size_t)width * height;
  if (n == 0) return B2ICP_OK;
  if (!data || !out_xyzw) return fail(h, B2ICP_ERR_INVALID_ARG, "pointcloud2: NULL argument");
  const uint32_t last = std::max(off_x, std::max(off_y, off_z)) + 4;
  if (point_step < last || (size_t)row_step < (size_t)width * point_step || data_bytes < (size_t)height * row_step)
    return fail(h, B2ICP_ERR_INVALID_ARG, "pointcloud2: steps / offsets do not fit the payload");
  if (n > (size_t)INT32_MAX / 8) return fail(h, B2ICP_ERR_INVALID_ARG, "cloud too large");
  CK(h->xf_in.ensure(data_bytes));
  CK(h->xf_out.ensure(n * sizeof(float4)));
  CK(cudaMemcpyAsync(h->xf_in.p, data, data_bytes, cudaMemcpyHostToDevice, h->stream));
  unpack_pointcloud2<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->xf_in.as<unsigned char>(), width, height, point_step, row_step,
                                                                         off_x, off_y, off_z, is_bigendian ? 1 : 0, h->xf_out.as<float4>());
  h->launches += 1;
  CK(cudaMemcpyAsync(out_xyzw, h->xf_out.p, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return B2ICP_OK;
}

int b2icp_get_timing(b2icp_handle* h, b2icp_timing* out) {
  if (!h || !out) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  h->timing.kernel_launches = h->launches;
  *out = h->timing;
  return B2ICP_OK;
}

int b2icp_get_grid_info(b2icp_handle* h, float* cell, int32_t* dims3, double* occupancy) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  GridSlot& g = gslot(h, 0);
  if (!g.valid) return fail(h, B2ICP_ERR_NO_TARGET, "no target cloud set");
  {
    CK(cudaSetDevice(h->device));
    int rc = ensure_grid(h, g);
    if (rc) return rc;
  }
  if (cell) *cell = g.view.cell;
  if (dims3) {
    dims3[0] = g.view.nx;
    dims3[1] = g.view.ny;
    dims3[2] = g.view.nz;
  }
  if (occupancy) *occupancy = g.occupancy;
  return B2ICP_OK;
}

int b2icp_host_alloc(size_t bytes, void** out) {
  if (!out) return B2ICP_ERR_INVALID_ARG;
  *out = nullptr;
  if (cudaMallocHost(out, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return B2ICP_ERR_CUDA;
  }
  return B2ICP_OK;
}
int b2icp_host_free(void* p) {
  if (p && cudaFreeHost(p) != cudaSuccess) {
    cudaGetLastError();
    return B2ICP_ERR_CUDA;
  }
  return B2ICP_OK;
}

const char* b2icp_last_error(const b2icp_handle* h) { return h ? h->err.c_str() : "null handle"; }

const char* b2icp_status_string(int s) {
  switch (s) {
    case B2ICP_OK: return "ok";
    case B2ICP_ERR_INVALID_ARG: return "invalid argument";
    case B2ICP_ERR_EMPTY_CLOUD: return "empty cloud";
    case B2ICP_ERR_TOO_FEW_POINTS: return "fewer points than k_correspondences";
    case B2ICP_ERR_NOT_ENOUGH_CORRESPONDENCES: return "not enough correspondences";
    case B2ICP_ERR_SOLVER_FAILED: return "solver failed";
    case B2ICP_ERR_NONFINITE_INPUT: return "non-finite input";
    case B2ICP_ERR_CUDA: return "CUDA error / no CUDA device";
    case B2ICP_ERR_NO_TARGET: return "no target cloud";
    case B2ICP_ERR_NO_SOURCE: return "no source cloud";
    case B2ICP_ERR_NOT_ALIGNED: return "align has not run";
    default: return "unknown status";
  }
}

int b2icp_version(void) { return B2ICP_VERSION_MAJOR * 1000 + B2ICP_VERSION_MINOR; }

}  // extern "C"
