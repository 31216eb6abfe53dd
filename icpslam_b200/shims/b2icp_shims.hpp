// b2icp_shims.hpp — ROS-free C++ host side above the C ABI (include/b2icp.h).
//
// Keeps the call surface of the reference's two estimator classes so that the existing node can link
// against them:
//   IcpOdometer   reference include/icpslam/icp_odometer.h:30-58, src/icpslam/icp_odometer.cpp
//   OctreeMapper  reference include/icpslam/octree_mapper.h:23-49, src/icpslam/octree_mapper.cpp
//   Pose6DOF      reference include/utils/pose6DOF.h, src/utils/pose6DOF.cpp (compose :98-105,
//                 inverse :117-122, fromEigenMatrix :185-190) — the SE(3) algebra the path uses
// Without ROS / PCL / Eigen headers (this image has none):  ros::Time -> double seconds,
// pcl::PointCloud<pcl::PointXYZ>::Ptr -> std::shared_ptr<b2::Cloud> whose `points` is a contiguous vector
// of 16-byte {x,y,z,w} (layout-identical to pcl::PointXYZ: a PCL build passes
// reinterpret_cast<const float*>(cloud->points.data()) with zero copies), Eigen::Matrix4d -> double[16]
// row-major.  ROS publishers / subscribers / TF are out of scope; the methods that only did that are
// kept as no-ops so call sites compile unchanged.
//
// Every registration, transform, filter, map operation and search below is a call of the C ABI: there is no CPU
// path.  The map of OctreeMapper (pcl::octree in the reference) lives in device memory behind b2icp_map_*
// (csrc/map.cuh): one point per voxel, first come wins, scan order kept; neighbours come from the engine's EXACT
// search by default (never farther than PCL's answer) or, with OctreeMapperParams::pcl_octree, from a restatement
// of PCL's greedy approxNearestSearch on PCL's own lattice; refineTransformAndGrowMap uploads the scan once and
// keeps every intermediate cloud on the device (b2icp_mapper_register / b2icp_mapper_grow).
// Both classes default to the estimator the reference instantiates (GICP); mode = B2ICP_MODE_P2P_SVD selects the
// point-to-point pipeline of the north star.  b2icp_ros_adapter.hpp restores the original ROS signatures.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b2icp.h"

namespace b2 {

struct PointXYZ {
  float x, y, z, w;
};
static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ layout");

struct Cloud {
  using Ptr = std::shared_ptr<Cloud>;
  std::vector<PointXYZ> points;
  const float* data() const { return reinterpret_cast<const float*>(points.data()); }
  float* data() { return reinterpret_cast<float*>(points.data()); }
  size_t size() const { return points.size(); }
};

// ---- Pose6DOF: position + unit quaternion (w,x,y,z) + stamp ------------------------------------
struct Pose6DOF {
  double time_stamp = 0;
  double pos[3] = {0, 0, 0};
  double rot[4] = {1, 0, 0, 0};

  Pose6DOF() {}
  Pose6DOF(const double* T16, double stamp) : time_stamp(stamp) { fromEigenMatrix(T16); }

  void setIdentity() { *this = Pose6DOF(); }

  static void quatMul(const double* a, const double* b, double* o) {
    double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
  }
  void normalize() {
    double n = std::sqrt(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2] + rot[3] * rot[3]);
    if (n > 0) for (double& v : rot) v /= n;
  }
  // rot.toRotationMatrix(), row-major 3x3
  void rotationMatrix(double* R) const {
    const double w = rot[0], x = rot[1], y = rot[2], z = rot[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
  }
  // pose6DOF.cpp:98-105
  static Pose6DOF compose(const Pose6DOF& p1, const Pose6DOF& p2) {
    Pose6DOF p3;
    p3.time_stamp = p2.time_stamp;
    double R[9];
    p1.rotationMatrix(R);
    for (int i = 0; i < 3; ++i) p3.pos[i] = p1.pos[i] + R[3 * i] * p2.pos[0] + R[3 * i + 1] * p2.pos[1] + R[3 * i + 2] * p2.pos[2];
    quatMul(p1.rot, p2.rot, p3.rot);
    p3.normalize();
    return p3;
  }
  // pose6DOF.cpp:117-122
  static Pose6DOF inverse(const Pose6DOF& pose) {
    Pose6DOF inv;
    const double n2 = pose.rot[0] * pose.rot[0] + pose.rot[1] * pose.rot[1] + pose.rot[2] * pose.rot[2] + pose.rot[3] * pose.rot[3];
    inv.rot[0] = pose.rot[0] / n2;
    for (int i = 1; i < 4; ++i) inv.rot[i] = -pose.rot[i] / n2;
    double R[9];
    inv.rotationMatrix(R);
    for (int i = 0; i < 3; ++i) inv.pos[i] = -(R[3 * i] * pose.pos[0] + R[3 * i + 1] * pose.pos[1] + R[3 * i + 2] * pose.pos[2]);
    return inv;
  }
  Pose6DOF inverse() const { return inverse(*this); }
  Pose6DOF operator+(const Pose6DOF& p2) const { return compose(*this, p2); }  // pose6DOF.h:53-56
  Pose6DOF& operator+=(const Pose6DOF& p2) { *this = compose(*this, p2); return *this; }

  // pose6DOF.cpp:185-190: pos = T[0:3,3]; rot = Quaterniond(T[0:3,0:3]) (Eigen's branches); normalize
  void fromEigenMatrix(const double* T) {
    pos[0] = T[3]; pos[1] = T[7]; pos[2] = T[11];
    const double m[3][3] = {{T[0], T[1], T[2]}, {T[4], T[5], T[6]}, {T[8], T[9], T[10]}};
    double t = m[0][0] + m[1][1] + m[2][2];
    if (t > 0) {
      t = std::sqrt(t + 1.0);
      rot[0] = 0.5 * t;
      t = 0.5 / t;
      rot[1] = (m[2][1] - m[1][2]) * t; rot[2] = (m[0][2] - m[2][0]) * t; rot[3] = (m[1][0] - m[0][1]) * t;
    } else {
      int i = 0;
      if (m[1][1] > m[0][0]) i = 1;
      if (m[2][2] > m[i][i]) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
      rot[1 + i] = 0.5 * t;
      t = 0.5 / t;
      rot[0] = (m[k][j] - m[j][k]) * t; rot[1 + j] = (m[j][i] + m[i][j]) * t; rot[1 + k] = (m[k][i] + m[i][k]) * t;
    }
    normalize();
  }
  // what pcl_ros::transformPointCloud builds from toTFTransform(): Eigen::Matrix4f, row-major here
  void toMatrix4f(float* T) const {
    double R[9];
    rotationMatrix(R);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) T[4 * r + c] = (float)R[3 * r + c];
      T[4 * r + 3] = (float)pos[r];
    }
    T[12] = T[13] = T[14] = 0.f;
    T[15] = 1.f;
  }
  double norm() const { return std::sqrt(pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2]); }
};

// RAII wrapper of one b2icp handle
class Engine {
 public:
  Engine(int preset, int mode, int device = 0) {
    b2icp_default_params(&params_, preset);
    params_.mode = mode;
    params_.device = device;
    int rc = b2icp_create(&params_, &h_);
    if (rc) throw std::runtime_error(std::string("b2icp_create: ") + b2icp_status_string(rc));
  }
  ~Engine() { if (h_) b2icp_destroy(h_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  b2icp_handle* get() const { return h_; }
  const b2icp_params& params() const { return params_; }
 private:
  b2icp_params params_;
  b2icp_handle* h_ = nullptr;
};

// pcl::VoxelGrid<PointXYZ>::filter with a cubic leaf -> b2icp_voxel_filter (K8 on the device).
inline int voxelGridFilter(b2icp_handle* h, const Cloud& in, double leaf, Cloud& out) {
  out.points.resize(in.points.size());
  size_t n_out = 0;
  int rc = b2icp_voxel_filter(h, in.data(), in.size(), (float)leaf, out.data(), &n_out);
  out.points.resize(rc == B2ICP_OK ? n_out : 0);
  return rc;
}

// ---- IcpOdometer ----------------------------------------------------------------------------------
struct IcpOdometerParams {
  int num_clouds_skip = 0;        // icp_odometer.cpp:45
  double voxel_leaf_size = 0.05;  // icp_odometer.cpp:46 (YAML: 0.2); <= 0 disables the filter
  int verbosity_level = 1;
  // The estimator.  Default = what the reference instantiates (pcl::GeneralizedIterativeClosestPoint,
  // icp_odometer.cpp:188): a node that links the shims gets the reference's estimator.  B2ICP_MODE_P2P_SVD is the
  // north-star point-to-point pipeline: ~20x the throughput, a different answer.
  int mode = B2ICP_MODE_GICP_BFGS;
  int device = 0;
};

class IcpOdometer {
 public:
  using Ptr = std::shared_ptr<IcpOdometer>;
  static constexpr double ICP_FITNESS_ACCEPT = 20.0;  // the literal at icp_odometer.cpp:201

  explicit IcpOdometer(const IcpOdometerParams& p = IcpOdometerParams())
      : prm_(p), engine_(B2ICP_PRESET_ODOMETER, p.mode, p.device), prev_cloud_(new Cloud()), curr_cloud_(new Cloud()) {
    init();
  }
  void init() { loadParameters(); advertisePublishers(); registerSubscribers(); }
  void loadParameters() {}        // parameters arrive through IcpOdometerParams
  void advertisePublishers() {}   // ROS topics: out of scope
  void registerSubscribers() {}

  b2icp_handle* engine() const { return engine_.get(); }
  bool isOdomReady() const { return odom_inited_; }
  void setInitialPose(const Pose6DOF& initial_pose) { icp_odom_poses_.push_back(initial_pose); initial_pose_set_ = true; }
  Pose6DOF getFirstPose() const { return icp_odom_poses_.front(); }
  Pose6DOF getLatestPose() const { return icp_odom_poses_.back(); }

  // icp_odometer.cpp:82-94
  void getEstimates(double& stamp, Cloud::Ptr& cloud, Pose6DOF& latest_icp_transform, Pose6DOF& icp_pose, bool& new_transform) {
    cloud = prev_cloud_;
    stamp = latest_stamp;
    latest_icp_transform = icp_latest_transform_;
    icp_pose = getLatestPose();
    new_transform = new_transform_;
    icp_latest_transform_.setIdentity();
    new_transform_ = false;
  }
  // icp_odometer.cpp:96-101
  void voxelFilterCloud(Cloud::Ptr* input, Cloud::Ptr* output) {
    if (prm_.voxel_leaf_size > 0) last_status = voxelGridFilter(engine_.get(), **input, prm_.voxel_leaf_size, **output);
    else **output = **input;
  }
  void publishPath(double) {}
  // icp_odometer.cpp:109-145
  bool updateICPOdometry(double stamp, const double* T16) {
    Pose6DOF transform(T16, stamp);
    Pose6DOF prev_pose = getLatestPose();
    Pose6DOF new_pose = prev_pose + transform;
    icp_latest_transform_ += transform;
    new_transform_ = true;
    icp_odom_poses_.push_back(new_pose);
    return true;
  }
  // icp_odometer.cpp:147-221 with the message already converted to a cloud in the robot frame
  void laserCloudCallback(double stamp, const Cloud::Ptr& input_cloud) {
    if (!initial_pose_set_) return;
    if (clouds_skipped_ < prm_.num_clouds_skip) { ++clouds_skipped_; return; }
    clouds_skipped_ = 0;
    Cloud::Ptr in = input_cloud;
    curr_cloud_.reset(new Cloud());
    voxelFilterCloud(&in, &curr_cloud_);
    if (curr_cloud_->points.empty()) return;
    b2icp_handle* h = engine_.get();
    if (prev_cloud_->points.empty()) {  // first cloud: prev = curr
      *prev_cloud_ = *curr_cloud_;
      last_status = b2icp_set_target(h, prev_cloud_->data(), prev_cloud_->size());
      return;
    }
    latest_stamp = stamp;
    // icp.setInputSource(curr_cloud_); the target (prev_cloud_) is already resident with its grid
    last_status = b2icp_set_source(h, curr_cloud_->data(), curr_cloud_->size());
    if (last_status) return;
    b2icp_result res;
    last_status = b2icp_align(h, nullptr, &res, nullptr);
    last_result = res;
    double fitness = std::numeric_limits<double>::max();
    if (last_status == B2ICP_OK) b2icp_fitness(h, std::numeric_limits<double>::max(), &fitness);
    last_fitness = fitness;
    if (last_status == B2ICP_OK && res.converged && fitness < ICP_FITNESS_ACCEPT) {
      // (the reference also transforms prev_cloud_ by T^-1 into a cloud nobody reads: skipped)
      if (updateICPOdometry(latest_stamp, res.T)) {
        odom_inited_ = true;
        *prev_cloud_ = *curr_cloud_;                       // icp_odometer.cpp:209 ...
        last_status = b2icp_promote_source_to_target(h);   // ... and the same on the device
      }
    }
  }

  int last_status = 0;
  b2icp_result last_result{};
  double last_fitness = 0;

 protected:
  IcpOdometerParams prm_;
  Engine engine_;
  bool initial_pose_set_ = false, odom_inited_ = false, new_transform_ = false;
  int clouds_skipped_ = 0;
  double latest_stamp = 0;
  Pose6DOF icp_latest_transform_;
  std::vector<Pose6DOF> icp_odom_poses_;
  Cloud::Ptr prev_cloud_, curr_cloud_;
};

// ---- OctreeMapper ---------------------------------------------------------------------------------
struct OctreeMapperParams {
  double octree_resolution = 0.5;  // octree_mapper.cpp:42 (YAML: 0.2)
  int verbosity_level = 1;
  int mode = B2ICP_MODE_GICP_BFGS;  // octree_mapper.cpp:104 instantiates GICP; B2ICP_MODE_P2P_SVD = the fast pipeline
  // false (default): neighbours from the engine's exact search on the global voxel lattice.  true: PCL-compatible
  // map (b2icp_map_reset_octree): lattice anchored on the first point, root box grown as pcl::octree grows it,
  // approxNearestNeighbors = PCL's greedy approxNearestSearch descent — nn_cloud as the reference builds it.
  bool pcl_octree = false;
  int device = 0;
};

class OctreeMapper {
 public:
  using Ptr = std::shared_ptr<OctreeMapper>;
  explicit OctreeMapper(const OctreeMapperParams& p = OctreeMapperParams())
      : prm_(p), icp_(B2ICP_PRESET_MAPPER, p.mode, p.device), search_(B2ICP_PRESET_MAPPER, p.mode, p.device) {
    init();
  }
  void init() { loadParameters(); advertisePublishers(); registerSubscribers(); resetMap(); }
  void loadParameters() {}
  void advertisePublishers() {}
  void registerSubscribers() {}

  // octree_mapper.cpp:56-60.  The map lives in device memory behind the C ABI (b2icp_map_*, csrc/map.cuh).
  void resetMap() {
    last_status = prm_.pcl_octree ? b2icp_map_reset_octree(search_.get(), prm_.octree_resolution)
                                  : b2icp_map_reset(search_.get(), prm_.octree_resolution);
    map_cloud_.reset(new Cloud());
    map_cloud_stale_ = false;
  }
  // octree_mapper.cpp:63-71: at most one point per octree_resolution voxel, first come wins, scan order kept
  void addPointsToMap(const Cloud::Ptr& input_cloud) {
    size_t added = 0;
    last_status = b2icp_map_insert(search_.get(), input_cloud->data(), input_cloud->size(), &added);
    if (added) map_cloud_stale_ = true;
  }
  // octree_mapper.cpp:73-90 with the engine's exact nearest neighbour (see the header comment)
  bool approxNearestNeighbors(const Cloud::Ptr& cloud, Cloud::Ptr& nearest_neighbors) {
    nearest_neighbors->points.clear();
    if (mapSize() == 0 || cloud->points.empty()) return false;
    nearest_neighbors->points.resize(cloud->size());
    size_t n_nn = 0;
    last_status = b2icp_map_nearest(search_.get(), cloud->data(), cloud->size(), nullptr, nearest_neighbors->data(), &n_nn);
    if (last_status) n_nn = 0;
    nearest_neighbors->points.resize(n_nn);
    return n_nn > 0;
  }
  // octree_mapper.cpp:92-99: pcl_ros::transformPointCloud(tf::Transform) == float 4x4
  void transformCloudToPoseFrame(const Cloud::Ptr& in_cloud, const Pose6DOF& pose, Cloud::Ptr& out_cloud) {
    float T[16];
    pose.toMatrix4f(T);
    out_cloud->points.resize(in_cloud->size());
    last_status = b2icp_transform_cloud_f(icp_.get(), in_cloud->data(), in_cloud->size(), T, out_cloud->data());
  }
  // octree_mapper.cpp:101-124
  bool estimateTransformICP(const Cloud::Ptr& curr_cloud, const Cloud::Ptr& nn_cloud, Pose6DOF& transform, double stamp = 0) {
    b2icp_handle* h = icp_.get();
    last_status = b2icp_set_source(h, curr_cloud->data(), curr_cloud->size());
    if (!last_status) last_status = b2icp_set_target(h, nn_cloud->data(), nn_cloud->size());
    if (last_status) return false;
    b2icp_result res;
    last_status = b2icp_align(h, nullptr, &res, nullptr);
    last_result = res;
    if (last_status == B2ICP_OK && res.converged) {
      transform = Pose6DOF(res.T, stamp);
      return true;
    }
    return false;
  }
  void publishPath(const Pose6DOF&) {}
  // octree_mapper.cpp:133-173.  The scan is uploaded once (b2icp_mapper_register) and every intermediate cloud
  // (cloud_in_map, nn_cloud_in_map, nn_cloud) stays on the device; only the 4x4 comes back.  The step-by-step
  // members above remain for callers that want the intermediate clouds.
  bool refineTransformAndGrowMap(double stamp, const Cloud::Ptr& cloud, const Pose6DOF& raw_pose, Pose6DOF& transform) {
    float Tr[16], Tri[16];
    raw_pose.toMatrix4f(Tr);
    b2icp_handle* h = search_.get();
    if (mapSize() == 0) {  // lines 138-142: the first scan seeds the map
      size_t added = 0;
      last_status = b2icp_mapper_grow(h, cloud->data(), cloud->size(), Tr, &added);
      if (added) map_cloud_stale_ = true;
      return false;
    }
    raw_pose.inverse().toMatrix4f(Tri);
    b2icp_result res;
    last_status = b2icp_mapper_register(h, cloud->data(), cloud->size(), Tr, Tri, &res);
    last_result = res;
    if (last_status == B2ICP_OK && res.converged) {
      transform = Pose6DOF(res.T, stamp);
      Pose6DOF refined_pose = raw_pose + transform;
      float Tf[16];
      refined_pose.toMatrix4f(Tf);
      size_t added = 0;
      last_status = b2icp_mapper_grow(h, nullptr, 0, Tf, &added);  // the scan is still on the device
      if (added) map_cloud_stale_ = true;
      return true;
    }
    return false;
  }
  size_t mapSize() {
    size_t n = 0;
    b2icp_map_size(search_.get(), &n);
    return n;
  }
  // host copy of map_cloud_ (what the reference publishes, octree_mapper.cpp:45-49): fetched on demand
  const Cloud::Ptr& mapCloud() {
    if (map_cloud_stale_) {
      size_t n = mapSize();
      map_cloud_->points.resize(n);
      last_status = b2icp_map_download(search_.get(), map_cloud_->data(), n, &n);
      map_cloud_stale_ = false;
    }
    return map_cloud_;
  }

  int last_status = 0;
  b2icp_result last_result{};

 protected:
  OctreeMapperParams prm_;
  Engine icp_, search_;
  Cloud::Ptr map_cloud_;
  bool map_cloud_stale_ = false;
};

}  // namespace b2
