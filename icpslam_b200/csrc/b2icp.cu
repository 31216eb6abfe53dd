// b2icp.cu — the C ABI of libb2icp.so (include/b2icp.h): handle, device memory, launch sequences.
//
// Drop-in boundary for the PCL registration object the reference builds on the stack
// (reference src/icpslam/icp_odometer.cpp:188-201, src/icpslam/octree_mapper.cpp:104-117).
// No C++ exception crosses this file's extern "C" functions; there is no CPU fallback.
#include "../../include/b2icp.h"

#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "grid.cuh"
#include "icp.cuh"
#include "nn.cuh"
#include "solve.cuh"

using namespace b2;

namespace {

constexpr int kMaxCells = 1 << 25;        // dense cell table cap (128 MiB of int32)
constexpr int kUnboundedRings = 3;        // ring budget of unbounded searches before the brute-force fallback
constexpr double kTargetOccupancy = 6.0;  // points per occupied cell the auto-sizing aims at
constexpr double kMaxOccupancy = 24.0;    // above this the grid is rebuilt with smaller cells

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
  size_t n = 0;
  bool valid = false;
};

struct Grid {
  DeviceBuf sorted, cell_start, cell_of, rank;
  GridView view;
  double occupancy = 0;  // mean points per occupied cell
  bool valid = false;
};

struct SetupArgs {
  ScanTask task;
  float guess[16];
};

__global__ void icp_setup_kernel(ScanTask* tasks, IcpState* st, SetupArgs a) {
  if (threadIdx.x == 0) {
    tasks[0] = a.task;
    for (int i = 0; i < 16; ++i) {
      st->Tinc[i] = a.guess[i];
      st->final_T[i] = a.guess[i];
    }
    st->mse = nan("");
    st->prev_mse = DBL_MAX;
    st->fitness_sum = 0;
    st->fitness_cnt = 0;
    st->iter = 0;
    st->done = 0;
    st->converged = 0;
    st->status = 0;
    st->n_corr = 0;
    st->ticket = 0;
    st->unresolved = 0;
    st->pad = 0;
  }
}

__global__ void zero_counter(unsigned int* c) { *c = 0; }

}  // namespace

struct b2icp_handle {
  b2icp_params params;
  IcpConfig cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  std::string err;

  Cloud src, tgt;
  Grid grid;
  DeviceBuf cur, corr_idx, corr_d2, partials, state, tasks, bbox, tile_sums, unres_list, unres_count;
  DeviceBuf query, q_idx, q_d2, xf_in, xf_out, mat;
  IcpState* h_state = nullptr;  // pinned
  BBox* h_bbox = nullptr;       // pinned
  bool aligned = false;

  std::vector<cudaEvent_t> events;
  b2icp_timing timing;
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
}

int rings_for_bound(const b2icp_handle* h) {
  if (!h->grid.valid || !std::isfinite(h->cfg.bound2)) return kUnboundedRings;
  double r = std::sqrt((double)h->cfg.bound2);
  double k = std::ceil(r / (double)h->grid.view.cell) + 2.0;
  return k > 1e6 ? 1000000 : (int)k;
}

// K1: build the neighbour grid over h->tgt.raw
int build_grid(b2icp_handle* h) {
  const int n = (int)h->tgt.n;
  const float4* pts = h->tgt.raw.as<float4>();
  const int blocks = (n + 255) / 256;
  bbox_init<<<1, 32, 0, h->stream>>>(h->bbox.as<BBox>());
  bbox_kernel<<<std::min(blocks, 148 * 8), 256, 0, h->stream>>>(pts, n, h->bbox.as<BBox>());
  CK(cudaMemcpyAsync(h->h_bbox, h->bbox.p, sizeof(BBox), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_bbox->nonfinite) return fail(h, B2ICP_ERR_NONFINITE_INPUT, "target cloud holds non-finite coordinates");
  float mn[3], mx[3];
  double ext[3];
  for (int d = 0; d < 3; ++d) {
    mn[d] = ord2f(h->h_bbox->mn[d]);
    mx[d] = ord2f(h->h_bbox->mx[d]);
    ext[d] = (double)mx[d] - (double)mn[d];
  }
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
  if (dims > 0) cell = std::pow(vol * kTargetOccupancy / (double)n, 1.0 / dims);
  const double r = h->params.max_correspondence_distance;
  if (r > 0 && std::isfinite(r)) cell = std::min(cell, 0.5 * r);
  if (h->params.grid_cell > 0) cell = h->params.grid_cell;
  if (!(cell > 0) || !std::isfinite(cell)) cell = 1.0;
  const double min_cell = (h->params.grid_cell > 0) ? cell : ((r > 0 && std::isfinite(r)) ? std::min(cell, r / 8.0) : cell / 8.0);

  for (int attempt = 0; attempt < 3; ++attempt) {
    // respect the dense-table cap
    long long nx, ny, nz;
    for (;;) {
      nx = (long long)std::floor(ext[0] / cell) + 1;
      ny = (long long)std::floor(ext[1] / cell) + 1;
      nz = (long long)std::floor(ext[2] / cell) + 1;
      if ((double)nx * (double)ny * (double)nz <= (double)kMaxCells) break;
      cell *= 1.26;
    }
    GridView& g = h->grid.view;
    g.ox = mn[0];
    g.oy = mn[1];
    g.oz = mn[2];
    g.cell = (float)cell;
    g.inv_cell = 1.0f / g.cell;
    g.nx = (int)nx;
    g.ny = (int)ny;
    g.nz = (int)nz;
    g.n = n;
    float amax = 0.f;
    for (int d = 0; d < 3; ++d) amax = std::max(amax, std::max(std::fabs(mn[d]), std::fabs(mx[d]) + g.cell));
    g.slack = std::max(amax, (float)emax) * 9.5367431640625e-7f + 1e-30f;  // 2^-20 relative
    const int ncell = g.nx * g.ny * g.nz;
    CK(h->grid.cell_start.ensure((size_t)(ncell + 1 + 4) * sizeof(int)));
    CK(h->grid.sorted.ensure((size_t)n * sizeof(float4)));
    CK(h->grid.cell_of.ensure((size_t)n * sizeof(int)));
    CK(h->grid.rank.ensure((size_t)n * sizeof(int)));
    const int ntiles = (ncell + kScanTile - 1) / kScanTile;
    CK(h->tile_sums.ensure((size_t)ntiles * sizeof(int)));
    int* cs = h->grid.cell_start.as<int>();
    CK(cudaMemsetAsync(cs, 0, (size_t)(ncell + 1) * sizeof(int), h->stream));
    bbox_init<<<1, 32, 0, h->stream>>>(h->bbox.as<BBox>());
    grid_count<<<blocks, 256, 0, h->stream>>>(pts, n, g, h->grid.cell_of.as<int>(), h->grid.rank.as<int>(), cs);
    scan_tile_sums<<<ntiles, kScanThreads, 0, h->stream>>>(cs, ncell, h->tile_sums.as<int>(), h->bbox.as<BBox>());
    scan_of_sums<<<1, kScanThreads, 0, h->stream>>>(h->tile_sums.as<int>(), ntiles);
    scan_apply<<<ntiles, kScanThreads, 0, h->stream>>>(cs, ncell, h->tile_sums.as<int>(), n);
    grid_scatter<<<blocks, 256, 0, h->stream>>>(pts, n, h->grid.cell_of.as<int>(), h->grid.rank.as<int>(), cs,
                                                 h->grid.sorted.as<float4>());
    CK(cudaMemcpyAsync(h->h_bbox, h->bbox.p, sizeof(BBox), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    g.pts = h->grid.sorted.as<float4>();
    g.cell_start = cs;
    const int occ_cells = std::max(1, h->h_bbox->occupied);
    h->grid.occupancy = (double)n / (double)occ_cells;
    if (h->grid.occupancy <= kMaxOccupancy || cell <= min_cell * 1.0001) break;
    // surfaces: occupancy scales ~ cell^2
    double shrink = std::sqrt(kTargetOccupancy / h->grid.occupancy);
    cell = std::max(min_cell, cell * std::max(shrink, 0.25));
  }
  h->grid.valid = true;
  return B2ICP_OK;
}

int upload_cloud(b2icp_handle* h, Cloud& c, const float* xyzw, size_t n, bool from_device) {
  if (!xyzw || n == 0) return fail(h, B2ICP_ERR_EMPTY_CLOUD, "empty cloud");
  if (n > (size_t)INT32_MAX / 8) return fail(h, B2ICP_ERR_INVALID_ARG, "cloud too large");
  CK(c.raw.ensure(n * sizeof(float4)));
  CK(cudaMemcpyAsync(c.raw.p, xyzw, n * sizeof(float4), from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                     h->stream));
  c.n = n;
  c.valid = true;
  return B2ICP_OK;
}

int set_target_impl(b2icp_handle* h, const float* xyzw, size_t n, bool from_device) {
  h->grid.valid = false;
  h->tgt.valid = false;
  h->aligned = false;
  int rc = upload_cloud(h, h->tgt, xyzw, n, from_device);
  if (rc) return rc;
  rc = build_grid(h);
  if (rc) h->tgt.valid = false;
  return rc;
}

int set_source_impl(b2icp_handle* h, const float* xyzw, size_t n, bool from_device) {
  h->src.valid = false;
  h->aligned = false;
  return upload_cloud(h, h->src, xyzw, n, from_device);
}

int ensure_work(b2icp_handle* h, size_t n) {
  const size_t ncta = (n + kSweepThreads - 1) / kSweepThreads;
  CK(h->cur.ensure(n * sizeof(float4)));
  CK(h->corr_idx.ensure(n * sizeof(int)));
  CK(h->corr_d2.ensure(n * sizeof(float)));
  CK(h->partials.ensure(ncta * kNumSums * sizeof(double)));
  CK(h->unres_list.ensure(n * sizeof(int)));
  return B2ICP_OK;
}

void fill_result(const b2icp_handle* h, b2icp_result* out) {
  const IcpState& s = *h->h_state;
  for (int i = 0; i < 16; ++i) out->T[i] = (double)s.final_T[i];
  out->converged = s.converged;
  out->iterations = s.iter;
  out->n_corr_last = s.n_corr;
  out->status_detail = s.status;
  out->mse_last = s.mse;
  out->fitness = std::nan("");
}

void identity_result(b2icp_result* out) {
  std::memset(out, 0, sizeof(*out));
  for (int i = 0; i < 4; ++i) out->T[5 * i] = 1.0;
  out->mse_last = std::nan("");
  out->fitness = std::nan("");
}

int align_impl(b2icp_handle* h, const float* guess, b2icp_result* out, float* aligned_xyzw) {
  if (!out) return fail(h, B2ICP_ERR_INVALID_ARG, "out == NULL");
  identity_result(out);
  if (!h->src.valid) return fail(h, B2ICP_ERR_NO_SOURCE, "no source cloud set");
  if (!h->tgt.valid || !h->grid.valid) return fail(h, B2ICP_ERR_NO_TARGET, "no target cloud set");
  if (h->params.mode != B2ICP_MODE_P2P_SVD) return fail(h, B2ICP_ERR_INVALID_ARG, "mode not implemented");
  const size_t n = h->src.n;
  int rc = ensure_work(h, n);
  if (rc) return rc;
  h->cfg.max_rings = rings_for_bound(h);

  SetupArgs a;
  a.task.grid = h->grid.view;
  a.task.src = h->src.raw.as<float4>();
  a.task.cur = h->cur.as<float4>();
  a.task.corr_idx = h->corr_idx.as<int>();
  a.task.corr_d2 = h->corr_d2.as<float>();
  a.task.partials = h->partials.as<double>();
  a.task.state = h->state.as<IcpState>();
  a.task.n = (int)n;
  a.task.pad = 0;
  for (int i = 0; i < 16; ++i) a.guess[i] = guess ? guess[i] : ((i % 5 == 0) ? 1.f : 0.f);
  icp_setup_kernel<<<1, 32, 0, h->stream>>>(h->tasks.as<ScanTask>(), h->state.as<IcpState>(), a);

  const bool prof = h->params.profile != 0;
  const int iters = std::max(h->params.max_iterations, 1);
  if (prof) {
    while ((int)h->events.size() < 2 * iters + 2) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      h->events.push_back(e);
    }
    CK(cudaEventRecord(h->events[0], h->stream));
  }
  dim3 grid((unsigned)((n + kSweepThreads - 1) / kSweepThreads), 1, 1);
  for (int it = 0; it < iters; ++it) {
    if (prof) CK(cudaEventRecord(h->events[2 + 2 * it], h->stream));
    icp_sweep_p2p<<<grid, kSweepThreads, 0, h->stream>>>(h->tasks.as<ScanTask>(), h->cfg);
    if (prof) CK(cudaEventRecord(h->events[3 + 2 * it], h->stream));
  }
  if (prof) CK(cudaEventRecord(h->events[1], h->stream));
  if (aligned_xyzw) {
    CK(h->xf_out.ensure(n * sizeof(float4)));
    transform_cloud_f<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(
        h->src.raw.as<float4>(), (int)n, h->state.as<IcpState>()->final_T, h->xf_out.as<float4>());
    CK(cudaMemcpyAsync(aligned_xyzw, h->xf_out.p, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaMemcpyAsync(h->h_state, h->state.p, sizeof(IcpState), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  fill_result(h, out);
  h->aligned = true;
  if (prof) {
    b2icp_timing& t = h->timing;
    std::memset(&t, 0, sizeof(t));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->events[0], h->events[1]);
    t.total_ms = ms;
    const int ran = std::min(iters, std::max(h->h_state->iter, 1));
    for (int it = 0; it < ran; ++it) {
      cudaEventElapsedTime(&ms, h->events[2 + 2 * it], h->events[3 + 2 * it]);
      t.nn_sweep_ms += ms;
    }
    t.nn_sweep_launches = ran;
  }
  if (h->h_state->status != 0) {
    h->err = h->h_state->status == B2ICP_ERR_NONFINITE_INPUT ? "non-finite source point or transform"
                                                              : "ICP loop ended early: not enough correspondences";
    return h->h_state->status;
  }
  return B2ICP_OK;
}

int nn_search_impl(b2icp_handle* h, const float4* d_q, size_t n, int* d_idx, float* d_d2) {
  CK(h->unres_list.ensure(n * sizeof(int)));
  zero_counter<<<1, 1, 0, h->stream>>>(h->unres_count.as<unsigned int>());
  nn_search_kernel<<<(unsigned)((n + kSweepThreads - 1) / kSweepThreads), kSweepThreads, 0, h->stream>>>(
      h->grid.view, d_q, (int)n, INFINITY, kUnboundedRings, d_idx, d_d2, h->unres_list.as<int>(),
      h->unres_count.as<unsigned int>());
  nn_brute_fallback<<<148 * 2, 256, 0, h->stream>>>(h->grid.view, d_q, h->unres_list.as<int>(),
                                                    h->unres_count.as<unsigned int>(), d_idx, d_d2);
  return B2ICP_OK;
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
  bool ok = cudaSetDevice(h->device) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMallocHost((void**)&h->h_state, sizeof(IcpState)) == cudaSuccess &&
            cudaMallocHost((void**)&h->h_bbox, sizeof(BBox)) == cudaSuccess &&
            h->state.ensure(sizeof(IcpState)) == cudaSuccess && h->tasks.ensure(sizeof(ScanTask)) == cudaSuccess &&
            h->bbox.ensure(sizeof(BBox)) == cudaSuccess && h->unres_count.ensure(sizeof(unsigned int)) == cudaSuccess &&
            h->mat.ensure(16 * sizeof(double)) == cudaSuccess;
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
  for (DeviceBuf* b : {&h->src.raw, &h->tgt.raw, &h->grid.sorted, &h->grid.cell_start, &h->grid.cell_of, &h->grid.rank,
                       &h->cur, &h->corr_idx, &h->corr_d2, &h->partials, &h->state, &h->tasks, &h->bbox, &h->tile_sums,
                       &h->unres_list, &h->unres_count, &h->query, &h->q_idx, &h->q_d2, &h->xf_in, &h->xf_out, &h->mat})
    b->release();
  for (cudaEvent_t e : h->events) cudaEventDestroy(e);
  if (h->h_state) cudaFreeHost(h->h_state);
  if (h->h_bbox) cudaFreeHost(h->h_bbox);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
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
  return set_target_impl(h, xyzw, n, false);
}
int b2icp_set_target_device(b2icp_handle* h, const float* d_xyzw, size_t n) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return set_target_impl(h, d_xyzw, n, true);
}
int b2icp_set_source(b2icp_handle* h, const float* xyzw, size_t n) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return set_source_impl(h, xyzw, n, false);
}
int b2icp_set_source_device(b2icp_handle* h, const float* d_xyzw, size_t n) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return set_source_impl(h, d_xyzw, n, true);
}

int b2icp_promote_source_to_target(b2icp_handle* h) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (!h->src.valid) return fail(h, B2ICP_ERR_NO_SOURCE, "no source cloud set");
  std::swap(h->src.raw, h->tgt.raw);
  h->tgt.n = h->src.n;
  h->tgt.valid = true;
  h->src.valid = false;
  h->src.n = 0;
  h->aligned = false;
  h->grid.valid = false;
  int rc = build_grid(h);
  if (rc) h->tgt.valid = false;
  return rc;
}

int b2icp_align(b2icp_handle* h, const float* guess, b2icp_result* out, float* aligned_xyzw) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  return align_impl(h, guess, out, aligned_xyzw);
}

int b2icp_fitness(b2icp_handle* h, double max_range, double* out) {
  if (!h || !out) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (!h->aligned) return fail(h, B2ICP_ERR_NOT_ALIGNED, "b2icp_fitness needs a completed b2icp_align");
  const size_t n = h->src.n;
  CK(h->query.ensure(n * sizeof(float4)));
  CK(h->q_idx.ensure(n * sizeof(int)));
  CK(h->q_d2.ensure(n * sizeof(float)));
  zero_counter<<<1, 1, 0, h->stream>>>(h->unres_count.as<unsigned int>());
  dim3 grid((unsigned)((n + kSweepThreads - 1) / kSweepThreads), 1, 1);
  fitness_kernel<<<grid, kSweepThreads, 0, h->stream>>>(h->tasks.as<ScanTask>(), kUnboundedRings, max_range,
                                                        h->unres_list.as<int>(), h->unres_count.as<unsigned int>(),
                                                        h->query.as<float4>(), h->q_idx.as<int>(), h->q_d2.as<float>());
  nn_brute_fallback<<<148 * 2, 256, 0, h->stream>>>(h->grid.view, h->query.as<float4>(), h->unres_list.as<int>(),
                                                    h->unres_count.as<unsigned int>(), h->q_idx.as<int>(),
                                                    h->q_d2.as<float>());
  fitness_reduce<<<1, 1024, 0, h->stream>>>(h->q_idx.as<int>(), h->q_d2.as<float>(), (int)n, max_range,
                                            h->state.as<IcpState>());
  CK(cudaMemcpyAsync(h->h_state, h->state.p, sizeof(IcpState), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  *out = h->h_state->fitness_cnt > 0 ? h->h_state->fitness_sum / (double)h->h_state->fitness_cnt : DBL_MAX;
  return B2ICP_OK;
}

int b2icp_get_correspondences(b2icp_handle* h, int32_t* tgt_idx, float* sqdist) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (!h->aligned) return fail(h, B2ICP_ERR_NOT_ALIGNED, "no completed b2icp_align");
  const size_t n = h->src.n;
  if (tgt_idx) CK(cudaMemcpyAsync(tgt_idx, h->corr_idx.p, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (sqdist) CK(cudaMemcpyAsync(sqdist, h->corr_d2.p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return B2ICP_OK;
}

int b2icp_nn_search(b2icp_handle* h, const float* q_xyzw, size_t n, int32_t* idx, float* sqdist) {
  if (!h || !idx) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->device));
  if (!h->tgt.valid || !h->grid.valid) return fail(h, B2ICP_ERR_NO_TARGET, "no target cloud set");
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
  if (!h->tgt.valid || !h->grid.valid) return fail(h, B2ICP_ERR_NO_TARGET, "no target cloud set");
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
  int worst = B2ICP_OK;
  for (size_t i = 0; i < batch; ++i) {
    identity_result(&out[i]);
    int rc = B2ICP_OK;
    if (tgt && tgt[i]) {
      rc = set_target_impl(h, tgt[i], n_tgt ? n_tgt[i] : 0, false);
    } else if (i > 0 && h->src.valid) {
      // consecutive-sweep odometry: the previous source becomes the target (icp_odometer.cpp:209)
      std::swap(h->src.raw, h->tgt.raw);
      h->tgt.n = h->src.n;
      h->tgt.valid = true;
      h->src.valid = false;
      h->grid.valid = false;
      rc = build_grid(h);
    } else if (!h->tgt.valid) {
      rc = fail(h, B2ICP_ERR_NO_TARGET, "pair 0 has no target");
    }
    if (!rc) rc = set_source_impl(h, src[i], n_src[i], false);
    if (!rc) rc = align_impl(h, nullptr, &out[i], nullptr);
    if (!rc && with_fitness) {
      // inline fitness (same sequence as b2icp_fitness)
      const size_t n = h->src.n;
      CK(h->query.ensure(n * sizeof(float4)));
      CK(h->q_idx.ensure(n * sizeof(int)));
      CK(h->q_d2.ensure(n * sizeof(float)));
      zero_counter<<<1, 1, 0, h->stream>>>(h->unres_count.as<unsigned int>());
      dim3 grid((unsigned)((n + kSweepThreads - 1) / kSweepThreads), 1, 1);
      fitness_kernel<<<grid, kSweepThreads, 0, h->stream>>>(h->tasks.as<ScanTask>(), kUnboundedRings, DBL_MAX,
                                                            h->unres_list.as<int>(), h->unres_count.as<unsigned int>(),
                                                            h->query.as<float4>(), h->q_idx.as<int>(), h->q_d2.as<float>());
      nn_brute_fallback<<<148 * 2, 256, 0, h->stream>>>(h->grid.view, h->query.as<float4>(), h->unres_list.as<int>(),
                                                        h->unres_count.as<unsigned int>(), h->q_idx.as<int>(),
                                                        h->q_d2.as<float>());
      fitness_reduce<<<1, 1024, 0, h->stream>>>(h->q_idx.as<int>(), h->q_d2.as<float>(), (int)n, DBL_MAX,
                                                h->state.as<IcpState>());
      CK(cudaMemcpyAsync(h->h_state, h->state.p, sizeof(IcpState), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      out[i].fitness = h->h_state->fitness_cnt > 0 ? h->h_state->fitness_sum / (double)h->h_state->fitness_cnt : DBL_MAX;
    }
    if (rc && !worst) worst = rc;
    out[i].status_detail = rc;
  }
  return worst;
}

int b2icp_get_timing(b2icp_handle* h, b2icp_timing* out) {
  if (!h || !out) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  *out = h->timing;
  return B2ICP_OK;
}

int b2icp_get_grid_info(b2icp_handle* h, float* cell, int32_t* dims3, double* occupancy) {
  if (!h) return B2ICP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  if (!h->grid.valid) return fail(h, B2ICP_ERR_NO_TARGET, "no target cloud set");
  if (cell) *cell = h->grid.view.cell;
  if (dims3) {
    dims3[0] = h->grid.view.nx;
    dims3[1] = h->grid.view.ny;
    dims3[2] = h->grid.view.nz;
  }
  if (occupancy) *occupancy = h->grid.occupancy;
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
