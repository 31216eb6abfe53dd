/*
 * b2icp_oracle.h — CPU oracle for the ICP hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libb2icp.so) never links, loads or calls it.
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in PCL (nominally 1.8.1,
 * un-vendored: reference CMakeLists.txt:7, package.xml:13-14) with FLANN and Eigen underneath;
 * none of them is present in this image and the reference ships no tests, fixtures or golden
 * vectors.  This file restates PCL's published algorithms (SURVEY.md
 * Appendix A) anchored on the reference's call sites; it is cross-checked against independent
 * implementations (scipy cKDTree, numpy SVD, scipy.optimize) and analytic known-answer tests in
 * tests/, never against PCL itself.
 */
#ifndef B2ICP_ORACLE_H_
#define B2ICP_ORACLE_H_

#include "../include/b2icp.h" /* parameter / result structs only */

#ifdef __cplusplus
extern "C" {
#endif

/* Per-stage wall-clock of one b2o_align call, milliseconds. */
typedef struct b2o_stage_ms {
  double build;       /* k-d tree builds (setInputTarget / initComputeReciprocal) */
  double covariances; /* GICP computeCovariances x2 */
  double nn;          /* all correspondence sweeps */
  double solve;       /* Umeyama or BFGS */
  double transform;   /* cloud transforms */
  double total;
} b2o_stage_ms;

/* OpenMP threads used over queries / correspondences (1 = PCL-faithful serial). Returns the
 * number actually in force. */
int b2o_set_threads(int n);
int b2o_get_max_threads(void);

/* Exact 1-NN by exhaustive scan.  FLANN L2_Simple arithmetic: d2 = ((dx*dx)+(dy*dy))+(dz*dz) in
 * float32; canonical tie rule = smallest target index among equal d2. */
int b2o_nn_brute(const float* tgt_xyzw, size_t nt, const float* q_xyzw, size_t nq, int32_t* idx,
                 float* d2);

/* pcl::KdTreeFLANN restatement (Appendix A.6): exact k-NN, results sorted by (d2, index). */
void* b2o_kdtree_build(const float* tgt_xyzw, size_t nt);
void b2o_kdtree_free(void* tree);
int b2o_kdtree_nn(const void* tree, const float* q_xyzw, size_t nq, int32_t* idx, float* d2);
int b2o_kdtree_knn(const void* tree, const float* q_xyzw, size_t nq, int k, int32_t* idx,
                   float* d2);

/* Appendix A.5: out = T*p, in T's scalar type, stored as float; w forced to 1. */
int b2o_transform_cloud_d(const float* in_xyzw, size_t n, const double* T, float* out_xyzw);
int b2o_transform_cloud_f(const float* in_xyzw, size_t n, const float* T, float* out_xyzw);

/* GICP computeCovariances (Appendix A.2): cov9 = n * 9 doubles, row-major 3x3 per point. */
int b2o_covariances(const float* xyzw, size_t n, int k, double gicp_epsilon, double* cov9);
/* tests only: every GICP cost-functor evaluation appends x[6], f, |g| (NaN = not asked for) to buf (8 doubles each) */
void b2o_gicp_trace(double* buf, size_t cap_records);
size_t b2o_gicp_trace_count(void);

/* TransformationEstimationSVD / Umeyama on already-matched pairs (Appendix A.3).
 * T16 row-major double (before the cast to float PCL would store). */
int b2o_umeyama(const float* src_xyzw, const float* dst_xyzw, size_t n, double* T16);
/* 3x3 SVD A = U diag(s) V^T (row-major 3x3, s descending; JacobiSVD FullU|FullV semantics). */
int b2o_svd3(const double* A9, double* U9, double* s3, double* V9);
/* the same recipe with the pair-skip threshold of the GICP covariances (|cos| <= 1e-15) */
int b2o_svd3_cov(const double* A9, double* U9, double* s3, double* V9);

/* Registration::align (Appendix A.1) with the solver selected by p->mode:
 * P2P_SVD = IterativeClosestPoint (A.3), GICP_BFGS = GeneralizedIterativeClosestPoint (A.2/A.4).
 * record_iter >= 0: copy the correspondences of that (0-based) outer iteration into
 * corr_idx[n_src] (-1 = gated out) / corr_d2[n_src]; record_iter = -1 records the last one. */
int b2o_align(const b2icp_params* p, const float* src_xyzw, size_t n_src, const float* tgt_xyzw,
              size_t n_tgt, const float* guess16, b2icp_result* out, float* aligned_xyzw,
              int record_iter, int32_t* corr_idx, float* corr_d2, b2o_stage_ms* stages);

/* Registration::getFitnessScore(max_range) for the float transform T (row-major, 16 floats). */
int b2o_fitness(const float* src_xyzw, size_t n_src, const float* tgt_xyzw, size_t n_tgt,
                const float* T16, double max_range, double* out);

/* pcl::VoxelGrid with a cubic leaf (Appendix A.8; reference icp_odometer.cpp:96-101). out holds n points. */
int b2o_voxel_filter(const float* in_xyzw, size_t n, float leaf, float* out_xyzw, size_t* n_out);
/* OctreeMapper::addPointsToMap (reference src/icpslam/octree_mapper.cpp:63-71) on the global voxel lattice
 * floor(p / resolution): the points of `in` that enter a map already holding `map` (first come wins, input
 * order), written to out_added (room for n points). */
int b2o_map_insert(const float* map_xyzw, size_t n_map, const float* in_xyzw, size_t n, double resolution,
                   float* out_added, size_t* n_added);

/* Pose6DOF algebra (reference src/utils/pose6DOF.cpp:98-105 compose, :117-122 inverse,
 * :185-190 fromEigenMatrix).  pose7 = {px,py,pz,qw,qx,qy,qz}. */
void b2o_pose_compose(const double* a7, const double* b7, double* out7);
void b2o_pose_inverse(const double* a7, double* out7);
void b2o_pose_from_matrix(const double* T16, double* out7);

/* pcl::octree::OctreePointCloudSearch as OctreeMapper uses it (octree_oracle.cpp) */
void* b2o_octree_create(double resolution);
void b2o_octree_free(void* tree);
size_t b2o_octree_add_points(void* tree, const float* xyzw, size_t n);
size_t b2o_octree_size(void* tree);
void b2o_octree_points(void* tree, float* out_xyzw);
void b2o_octree_box(void* tree, double* min3, int* depth);
void b2o_octree_approx_nearest(void* tree, const float* q_xyzw, size_t n, int key_rule, int32_t* idx);

#ifdef __cplusplus
}
#endif
#endif
