/*
 * b2icp.h — C ABI of the Blackwell-native ICP scan-matching engine (libb2icp.so).
 *
 * This is the drop-in boundary for the hot path of YoshuaNava/icpslam.  The reference has no
 * FFI seam: it builds a PCL registration object on the stack and calls it directly
 * (reference src/icpslam/icp_odometer.cpp:188-201, src/icpslam/octree_mapper.cpp:104-117).
 * Every entry point below names the reference lines (or the PCL call made from them) that it
 * replaces.  All symbols are extern "C", take plain pointers and sizes, never throw, and return
 * 0 on success or a negative b2icp_status.  Clouds are arrays of 16-byte points {x,y,z,w}
 * (layout-identical to pcl::PointXYZ, so `reinterpret_cast<const float*>(cloud->points.data())`
 * is a zero-copy argument).  The 4th float is ignored on input and written as 1.0f on output,
 * exactly as pcl::Registration::align / setInputSource force data[3] = 1.
 *
 * Matrices: every 4x4 in this ABI is ROW-MAJOR (T[4*r+c]); p_target = T * p_source.
 *
 * Threading: calls on one handle are serialised by an internal mutex; distinct handles are
 * independent (own CUDA streams, device buffers and pinned memory) and may be driven from
 * different host threads at the same time.  Entry points are synchronous (results are in the
 * caller's buffers on return) because the reference consumes T immediately
 * (icp_odometer.cpp:199-206); the exception is the pair b2icp_align_batch_submit[_device] /
 * b2icp_align_batch_wait, which keeps several batches in flight for offline replay.
 *
 * There is NO CPU backend behind these symbols: without a CUDA device b2icp_create fails with
 * B2ICP_ERR_CUDA.  (The CPU restatement under oracle/ is test infrastructure only.)
 */
#ifndef B2ICP_H_
#define B2ICP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2ICP_VERSION_MAJOR 0
#define B2ICP_VERSION_MINOR 1

typedef struct b2icp_handle b2icp_handle;

/* Error codes (returned negative). PCL prints PCL_ERROR and leaves converged_=false; we return. */
typedef enum b2icp_status {
  B2ICP_OK = 0,
  B2ICP_ERR_INVALID_ARG = -1,
  B2ICP_ERR_EMPTY_CLOUD = -2,                /* setInputSource rejects an empty cloud (PCL gicp.h) */
  B2ICP_ERR_TOO_FEW_POINTS = -3,             /* N < k_correspondences: PCL computeCovariances bails out */
  B2ICP_ERR_NOT_ENOUGH_CORRESPONDENCES = -4, /* < 3 (P2P) / < 4 (GICP) matches: converged=false */
  B2ICP_ERR_SOLVER_FAILED = -5,
  B2ICP_ERR_NONFINITE_INPUT = -6,
  B2ICP_ERR_CUDA = -7,
  B2ICP_ERR_NO_TARGET = -8,
  B2ICP_ERR_NO_SOURCE = -9,
  B2ICP_ERR_NOT_ALIGNED = -10                /* fitness / correspondences requested before align */
} b2icp_status;

/* Per-iteration solver. */
typedef enum b2icp_mode {
  /* pcl::IterativeClosestPoint pipeline (header included at octree_mapper.cpp:8): 1-NN,
   * keep d^2 <= max^2, Umeyama/SVD, in-place float transform, DefaultConvergenceCriteria.
   * This is the pipeline BASELINE.json's north_star spells out. */
  B2ICP_MODE_P2P_SVD = 0,
  /* pcl::GeneralizedIterativeClosestPoint — what the reference instantiates
   * (icp_odometer.cpp:188, octree_mapper.cpp:104): k=20 covariances, Mahalanobis, BFGS. */
  B2ICP_MODE_GICP_BFGS = 1
} b2icp_mode;

/* Presets for b2icp_default_params: the two constant blocks of the reference. */
typedef enum b2icp_preset {
  B2ICP_PRESET_ODOMETER = 0, /* include/icpslam/icp_odometer.h:62-65  (10 iterations) */
  B2ICP_PRESET_MAPPER = 1    /* include/icpslam/octree_mapper.h:53-56 (30 iterations) */
} b2icp_preset;

typedef struct b2icp_params {
  int32_t mode;                       /* b2icp_mode */
  int32_t max_iterations;             /* setMaximumIterations: 10 odometer / 30 mapper */
  double transformation_epsilon;      /* setTransformationEpsilon: 1e-6 */
  double max_correspondence_distance; /* setMaxCorrespondenceDistance: 1.0 m */
  double euclidean_fitness_epsilon;   /* PCL default -DBL_MAX (never set by the reference) */
  double rotation_epsilon;            /* GICP default 2e-3 */
  double gicp_epsilon;                /* GICP default 1e-3 */
  int32_t k_correspondences;          /* GICP default 20 */
  int32_t max_inner_iterations;       /* GICP default 20 */
  int32_t device;                     /* CUDA device ordinal */
  int32_t profile;                    /* !=0: record CUDA events around every nn_sweep launch */
  float grid_cell;                    /* neighbour-grid cell edge in metres; <=0 = auto */
  int32_t reserved[5];
} b2icp_params;

typedef struct b2icp_result {
  double T[16];          /* getFinalTransformation().cast<double>() — row-major, source -> target */
  int32_t converged;     /* hasConverged() */
  int32_t iterations;    /* outer iterations executed (nr_iterations_) */
  int32_t n_corr_last;   /* correspondences that passed the distance gate in the last iteration */
  int32_t status_detail; /* b2icp_status of the solver loop (0, NOT_ENOUGH_CORRESPONDENCES, ...) */
  double mse_last;       /* mean gated d^2 of the last iteration (DefaultConvergenceCriteria's MSE) */
  double fitness;        /* NaN unless filled by b2icp_align_fitness / b2icp_fitness */
} b2icp_result;

/* Device timings of the last b2icp_align / b2icp_align_batch chunk / b2icp_nn_search_device on this
 * handle (CUDA events on the handle's stream, recorded only when params.profile != 0). */
typedef struct b2icp_timing {
  int32_t nn_sweep_launches; /* ICP iterations that did work (each = check + search + sums launches) */
  int32_t reserved;
  double nn_sweep_ms;        /* sum of their CUDA-event durations */
  double build_ms;           /* reserved */
  double total_ms;           /* first launch -> last launch, CUDA events */
  int64_t kernel_launches;   /* every kernel this handle has launched since b2icp_create (always filled) */
  uint64_t nn_searches;      /* queries of those launches that needed a real neighbour search (the rest were
                              * settled by their cached-neighbour certificate, nncache.cuh) */
} b2icp_timing;

/* Fill `p` with the reference's constants for one of its two call sites. */
int b2icp_default_params(b2icp_params* p, int preset);

/* Lifetime of what the reference does with `GeneralizedIterativeClosestPoint icp;` on the stack
 * (icp_odometer.cpp:188-192, octree_mapper.cpp:104-108: ctor + the four set* calls). */
int b2icp_create(const b2icp_params* p, b2icp_handle** out);
int b2icp_destroy(b2icp_handle* h);
/* Change solver parameters on a live handle (keeps device buffers). */
int b2icp_set_params(b2icp_handle* h, const b2icp_params* p);

/* icp.setInputTarget(prev_cloud_) (icp_odometer.cpp:194, octree_mapper.cpp:110) + the k-d tree
 * build PCL does inside align(): uploads the cloud and builds the neighbour grid. */
int b2icp_set_target(b2icp_handle* h, const float* xyzw, size_t n);
/* icp.setInputSource(curr_cloud_) (icp_odometer.cpp:193, octree_mapper.cpp:109). */
int b2icp_set_source(b2icp_handle* h, const float* xyzw, size_t n);
/* Same, for clouds that already live in device memory of the handle's device
 * (device-resident pipelines and the HBM-resident leg of bench.py). */
int b2icp_set_target_device(b2icp_handle* h, const float* d_xyzw, size_t n);
int b2icp_set_source_device(b2icp_handle* h, const float* d_xyzw, size_t n);
/* `*prev_cloud_ = *curr_cloud_;` (icp_odometer.cpp:209): the current source becomes the target,
 * keeping its device buffers (and, in GICP mode, its covariances). */
int b2icp_promote_source_to_target(b2icp_handle* h);

/* icp.align(out) + getFinalTransformation() + hasConverged()
 * (icp_odometer.cpp:198-201, octree_mapper.cpp:114-117).  guess: 16 floats row-major or NULL
 * (identity — the reference never passes one).  aligned_xyzw: n_source points or NULL. */
int b2icp_align(b2icp_handle* h, const float* guess, b2icp_result* out, float* aligned_xyzw);

/* icp.getFitnessScore(max_range) (icp_odometer.cpp:201; PCL default max_range = DBL_MAX):
 * mean squared distance from every aligned source point to its exact nearest target point,
 * counting only pairs with d^2 <= max_range.  Requires a completed b2icp_align. */
int b2icp_fitness(b2icp_handle* h, double max_range, double* out);

/* Correspondences of the last executed iteration: tgt_idx[i] = index into the target cloud as
 * passed to b2icp_set_target, or -1 if gated out; sqdist[i] = float32 squared distance
 * (undefined where idx = -1).  Either pointer may be NULL. */
int b2icp_get_correspondences(b2icp_handle* h, int32_t* tgt_idx, float* sqdist);

/* Stand-alone exact 1-NN of n query points against the current target (KdTreeFLANN::
 * nearestKSearch(k=1), unbounded radius).  Ties in float32 d^2 resolve to the SMALLEST target
 * index (canonical rule).  This is the roofline kernel. */
int b2icp_nn_search(b2icp_handle* h, const float* q_xyzw, size_t n, int32_t* idx, float* sqdist);
int b2icp_nn_search_device(b2icp_handle* h, const float* d_q_xyzw, size_t n, int32_t* d_idx,
                           float* d_sqdist);

/* pcl::transformPointCloud(in, out, Matrix4d) (icp_odometer.cpp:205): double math, float store. */
int b2icp_transform_cloud(b2icp_handle* h, const float* in_xyzw, size_t n, const double* T,
                          float* out_xyzw);
/* pcl_ros::transformPointCloud(in, out, tf::Transform) (octree_mapper.cpp:96) after its
 * conversion to Eigen::Matrix4f: float math. */
int b2icp_transform_cloud_f(b2icp_handle* h, const float* in_xyzw, size_t n, const float* T,
                            float* out_xyzw);

/* Offline replay: `batch` independent registrations, each what one laserCloudCallback does at
 * icp_odometer.cpp:188-201 (or one estimateTransformICP at octree_mapper.cpp:104-117).  Up to 64 scans
 * advance together: one launch of the fused sweep kernel per ICP iteration serves all of them.
 *   tgt == NULL          every source registers against the handle's current target (b2icp_set_target):
 *                        scan-to-map localisation against a resident map;
 *   tgt[i] != NULL       pair i has its own target cloud of n_tgt[i] points;
 *   tgt[i] == NULL, i>0  pair i registers against src[i-1] (consecutive-sweep odometry: the uploaded
 *                        cloud is indexed in place, `*prev_cloud_ = *curr_cloud_` costs nothing);
 *   tgt[0] == NULL       pair 0 registers against the handle's current target.
 * with_fitness != 0 also fills out[i].fitness (getFitnessScore()).  Returns the first non-zero
 * per-scan status (each out[i].status_detail holds its own). */
int b2icp_align_batch(b2icp_handle* h, const float* const* src, const size_t* n_src,
                      const float* const* tgt, const size_t* n_tgt, size_t batch,
                      int with_fitness, b2icp_result* out);
/* Same with every cloud pointer in device memory of the handle's device. */
int b2icp_align_batch_device(b2icp_handle* h, const float* const* d_src, const size_t* n_src,
                             const float* const* d_tgt, const size_t* n_tgt, size_t batch,
                             int with_fitness, b2icp_result* out);

/* Run all work of this handle on the caller's CUDA stream (a cudaStream_t) instead of the private one,
 * so that device-resident pipelines can order their own kernels and events around the ABI calls. */
int b2icp_set_stream(b2icp_handle* h, void* cuda_stream);

/* GeneralizedIterativeClosestPoint::computeCovariances (reached from icp.align() at icp_odometer.cpp:198,
 * octree_mapper.cpp:114): for each of the n points the covariance of its k_correspondences nearest
 * neighbours (itself included), eigenvalues replaced by (1, 1, gicp_epsilon).  cov9 = n x 9 doubles,
 * row-major 3x3, in input order.  Returns B2ICP_ERR_TOO_FEW_POINTS when n < k_correspondences. */
int b2icp_compute_covariances(b2icp_handle* h, const float* xyzw, size_t n, double* cov9);

/* IcpOdometer::voxelFilterCloud -> pcl::VoxelGrid<PointXYZ>::filter (icp_odometer.cpp:96-101) with leaf size
 * `leaf` on all three axes: one centroid per occupied leaf, output in ascending voxel index (x fastest),
 * float accumulation in input order inside a leaf.  out_xyzw must hold n points; *n_out receives the count.
 * Like PCL, a leaf so small that the voxel count overflows an int returns the input unchanged. */
int b2icp_voxel_filter(b2icp_handle* h, const float* in_xyzw, size_t n, float leaf, float* out_xyzw,
                       size_t* n_out);

/* Streaming form of b2icp_align_batch for host clouds against the handle's current target (point-to-point
 * mode): _submit enqueues the uploads of up to 32 scans on a copy stream and their ICP loops behind them and
 * returns at once; _wait blocks until the OLDEST submitted batch is done and writes its results (*n_out of them,
 * in submission order).  Up to B2ICP_MAX_IN_FLIGHT batches may be in flight (each on its own stream), so the PCIe
 * transfer of the next batches overlaps the sweeps of the current one and the sparse late iterations of one batch
 * share the device with the first iterations of the next.  The host clouds of a batch must stay valid (and should be page-locked, b2icp_host_alloc) until its
 * _wait returns; the DEVICE clouds of _submit_device are read in place (no copy) and must stay valid and unchanged
 * until then too.  The synchronous calls must not be mixed in while batches are in flight. */
#define B2ICP_MAX_IN_FLIGHT 8
int b2icp_align_batch_submit(b2icp_handle* h, const float* const* src, const size_t* n_src, size_t batch, int with_fitness);
int b2icp_align_batch_submit_device(b2icp_handle* h, const float* const* d_src, const size_t* n_src, size_t batch,
                                    int with_fitness);
int b2icp_align_batch_wait(b2icp_handle* h, b2icp_result* out, size_t capacity, size_t* n_out);

/* The fixed-size per-scan record that travels between ranks in offline replay (the ONLY
 * exchange of the path is the gather of the per-scan rigid transforms back to rank 0, which then composes
 * pose_i = pose_{i-1} o T_i as icp_odometer.cpp:111-113 does). */
typedef struct b2icp_record {
  float T[16];           /* getFinalTransformation(): Matrix4f, row-major, source -> target */
  int32_t converged;     /* hasConverged() */
  int32_t iterations;
  int32_t n_corr_last;
  int32_t status_detail;
  double mse_last;
  double fitness;        /* getFitnessScore(); NaN when the batch did not ask for it */
} b2icp_record;
/* Record sink in DEVICE memory: every streamed batch submitted after this call appends one b2icp_record per scan
 * at d_records (submission order, stream-ordered before the batch signals completion), so that a replay can hand
 * the records of many batches to ONE collective (ncclAllGather) without a host round trip or a per-batch barrier.
 * d_records == NULL switches the sink off.  b2icp_record_sink_count: records written by the batches waited for
 * so far plus those still in flight. */
int b2icp_set_record_sink(b2icp_handle* h, b2icp_record* d_records, size_t capacity);
int b2icp_record_sink_count(b2icp_handle* h, size_t* n);

/* ---- the mapper's point map (OctreeMapper::map_octree_ / map_cloud_, octree_mapper.h:82-83) --------------
 * b2icp_map_reset          OctreeMapper::resetMap (octree_mapper.cpp:56-60): empty map at `resolution`.
 * b2icp_map_insert         OctreeMapper::addPointsToMap (octree_mapper.cpp:63-71): a point enters the map iff
 *                          no earlier point (of the map or of this call, in input order) lies in its voxel;
 *                          new points are appended in input order.  *n_added (nullable) = points added.
 *                          A voxel is a cell of the lattice floor(p / resolution) (see csrc/map.cuh).
 * b2icp_map_nearest        OctreeMapper::approxNearestNeighbors (octree_mapper.cpp:73-90) with the engine's
 *                          EXACT nearest neighbour instead of PCL's greedy octree descent: idx[i] (nullable) =
 *                          index in the map of the point nearest to query i, -1 if the query is not finite;
 *                          nn_xyzw (nullable, room for n points) = those map points compacted in query order
 *                          (the reference's nn_cloud, duplicates included), *n_nn their number.
 * b2icp_set_target_map     the map itself becomes the registration target (device to device): scan-to-map
 *                          localisation without the gather (BASELINE configs[1] and [4]).
 * The map lives in device memory; download / size are for the caller's bookkeeping and publishing. */
int b2icp_map_reset(b2icp_handle* h, double resolution);
/* PCL-COMPATIBLE map mode ("offer two NN modes"): like b2icp_map_reset, but the map then
 * behaves as pcl::octree::OctreePointCloudSearch does in OctreeMapper (octree_mapper.cpp:56-90):
 *   - voxels are the leaves of PCL's octree: a lattice anchored on the FIRST point inserted (its corner is that
 *     point minus one resolution), not the global lattice floor(p / resolution);
 *   - the root box grows as PCL grows it (new roots towards the violated bounds), tracked exactly;
 *   - b2icp_map_nearest / b2icp_mapper_register pair every query with the map point of the leaf that PCL's
 *     approxNearestSearch reaches — the greedy descent by voxel-centre distance, NOT the nearest neighbour —
 *     so that nn_cloud is the reference's nn_cloud.  Exact mode (b2icp_map_reset) stays the default. */
int b2icp_map_reset_octree(b2icp_handle* h, double resolution);
int b2icp_map_insert(b2icp_handle* h, const float* xyzw, size_t n, size_t* n_added);
int b2icp_map_insert_device(b2icp_handle* h, const float* d_xyzw, size_t n, size_t* n_added);
int b2icp_map_size(b2icp_handle* h, size_t* n);
int b2icp_map_download(b2icp_handle* h, float* out_xyzw, size_t capacity, size_t* n);
int b2icp_map_nearest(b2icp_handle* h, const float* q_xyzw, size_t n, int32_t* idx, float* nn_xyzw, size_t* n_nn);
int b2icp_set_target_map(b2icp_handle* h);
/* OctreeMapper::refineTransformAndGrowMap (octree_mapper.cpp:133-173) with ONE upload of the scan and no cloud coming
 * back: every intermediate cloud stays in device memory.
 *   b2icp_mapper_register  cloud_in_map = T_raw * cloud (line 136); nn_cloud_in_map = the map point nearest to every
 *                          scan point (line 145); nn_cloud = T_raw_inv * nn_cloud_in_map (line 149);
 *                          estimateTransformICP(cloud, nn_cloud) (line 152) -> *out.  T_raw / T_raw_inv: row-major
 *                          float 4x4 of raw_pose and its inverse (what pcl_ros::transformPointCloud builds from the
 *                          tf::Transform).  B2ICP_ERR_NO_TARGET while the map is empty (the caller then only grows).
 *   b2icp_mapper_grow      cloud_in_map = T * cloud; addPointsToMap(cloud_in_map) (lines 139-140, 157-158).  xyzw == NULL:
 *                          the scan of the last b2icp_mapper_register call, still on the device. */
int b2icp_mapper_register(b2icp_handle* h, const float* xyzw, size_t n, const float* T_raw, const float* T_raw_inv,
                          b2icp_result* out);
int b2icp_mapper_grow(b2icp_handle* h, const float* xyzw, size_t n, const float* T, size_t* n_added);

/* pcl::fromROSMsg(sensor_msgs::PointCloud2, pcl::PointCloud<pcl::PointXYZ>) (icp_odometer.cpp:168,173): the x, y, z
 * FLOAT32 fields of a PointCloud2 payload -> 16-byte {x, y, z, 1} points, in message order (row-major over height x
 * width, row_step bytes per row, point_step bytes per point, off_* = byte offsets of the three fields inside a point).
 * The payload is uploaded once and unpacked on the device (a strided gather); NaN / Inf coordinates are copied as
 * they are, like fromROSMsg does (VoxelGrid / the map skip them later).  A payload whose points already are
 * {x, y, z, pad} floats (point_step 16, offsets 0 / 4 / 8) needs no call at all: pass msg.data.data() to the cloud
 * entry points.  out_xyzw must hold width * height points.  is_bigendian != 0 swaps the bytes of every float. */
int b2icp_pointcloud2_to_xyzw(b2icp_handle* h, const uint8_t* data, size_t data_bytes, uint32_t width, uint32_t height,
                              uint32_t point_step, uint32_t row_step, uint32_t off_x, uint32_t off_y, uint32_t off_z,
                              int is_bigendian, float* out_xyzw);

int b2icp_get_timing(b2icp_handle* h, b2icp_timing* out);
/* Neighbour grid of the current target (the structure that replaces the FLANN k-d tree): cell edge,
 * dims3 = {nx, ny, nz}, mean points per occupied cell.  Any pointer may be NULL. */
int b2icp_get_grid_info(b2icp_handle* h, float* cell, int32_t* dims3, double* occupancy);
/* Page-locked host buffers: clouds handed to the set / search calls from such memory are copied
 * with true asynchronous DMA (pcl::PointCloud's allocator can be pointed here). */
int b2icp_host_alloc(size_t bytes, void** out);
int b2icp_host_free(void* p);
const char* b2icp_last_error(const b2icp_handle* h);
const char* b2icp_status_string(int status);
int b2icp_version(void);

#ifdef __cplusplus
}
static_assert(sizeof(b2icp_params) == 88, "b2icp_params layout is part of the ABI");
static_assert(sizeof(b2icp_result) == 160, "b2icp_result layout is part of the ABI");
static_assert(sizeof(b2icp_timing) == 48, "b2icp_timing layout is part of the ABI");
static_assert(sizeof(b2icp_record) == 96, "b2icp_record layout is part of the ABI");
#endif
#endif /* B2ICP_H_ */
