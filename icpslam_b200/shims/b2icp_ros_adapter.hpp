// b2icp_ros_adapter.hpp — the ROS-facing side of the drop-in: the reference's ORIGINAL class signatures
// (reference include/icpslam/icp_odometer.h:30-58, include/icpslam/octree_mapper.h:23-49) on top of the ROS-free
// shims of b2icp_shims.hpp, so that icpslam.cpp / icpslam_node.cpp compile unchanged against libb2icp.so:
//
//     // CMakeLists.txt of the node:  target_link_libraries(icpslam b2icp)  and, instead of the two reference headers,
//     #include "b2icp_ros_adapter.hpp"      // defines ::IcpOdometer and ::OctreeMapper
//
// This image has no ROS / PCL / Eigen, so everything that needs their headers sits behind
// `#if __has_include(<ros/ros.h>)`; what the adapter DOES with a message — the PointCloud2 -> cloud conversion of
// icp_odometer.cpp:147-175 — is split out into the ROS-free part below (PointCloud2View, fromROSMsg) and is what
// tests/cpp/shim_driver.cpp and tests/test_gpu_shims.py exercise.
#pragma once
#include "b2icp_shims.hpp"

namespace b2 {

// The fields of sensor_msgs::PointCloud2 that pcl::fromROSMsg reads.
struct PointCloud2View {
  const uint8_t* data = nullptr;
  size_t data_bytes = 0;
  uint32_t width = 0, height = 0, point_step = 0, row_step = 0;
  uint32_t off_x = 0, off_y = 4, off_z = 8;  // byte offsets of the FLOAT32 fields "x", "y", "z"
  bool is_bigendian = false;
};

// pcl::fromROSMsg(msg, pcl::PointCloud<pcl::PointXYZ>) (icp_odometer.cpp:168,173).  A message whose points already are
// {x, y, z, pad} floats is taken over with one memcpy (it is layout-identical to pcl::PointXYZ); any other layout is
// unpacked on the device by b2icp_pointcloud2_to_xyzw.
inline int fromROSMsg(b2icp_handle* h, const PointCloud2View& m, Cloud& out) {
  const size_t n = (size_t)m.width * m.height;
  out.points.resize(n);
  if (n == 0) return B2ICP_OK;
  if (m.point_step == 16 && m.off_x == 0 && m.off_y == 4 && m.off_z == 8 && !m.is_bigendian &&
      m.row_step == m.width * 16u && m.data_bytes >= n * 16) {
    std::memcpy(out.points.data(), m.data, n * 16);
    for (auto& p : out.points) p.w = 1.0f;  // pcl::PointXYZ::data[3] = 1
    return B2ICP_OK;
  }
  return b2icp_pointcloud2_to_xyzw(h, m.data, m.data_bytes, m.width, m.height, m.point_step, m.row_step, m.off_x, m.off_y,
                                   m.off_z, m.is_bigendian ? 1 : 0, out.data());
}

}  // namespace b2

#if defined(__has_include)
#if __has_include(<ros/ros.h>) && __has_include(<pcl/point_cloud.h>) && __has_include(<Eigen/Dense>)
#define B2ICP_HAVE_ROS 1
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>
#include <tf2_ros/transform_listener.h>
#include <tf2_sensor_msgs/tf2_sensor_msgs.h>
#include <Eigen/Dense>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

namespace b2 {

inline PointCloud2View view_of(const sensor_msgs::PointCloud2& msg) {
  PointCloud2View v;
  v.data = msg.data.data();
  v.data_bytes = msg.data.size();
  v.width = msg.width;
  v.height = msg.height;
  v.point_step = msg.point_step;
  v.row_step = msg.row_step;
  v.is_bigendian = msg.is_bigendian;
  for (const auto& f : msg.fields) {
    if (f.name == "x") v.off_x = f.offset;
    if (f.name == "y") v.off_y = f.offset;
    if (f.name == "z") v.off_z = f.offset;
  }
  return v;
}
// pcl::PointXYZ and b2::PointXYZ are both four floats: the clouds are exchanged with one memcpy either way
inline Cloud::Ptr from_pcl(const pcl::PointCloud<pcl::PointXYZ>& c) {
  Cloud::Ptr out(new Cloud());
  out->points.resize(c.points.size());
  if (!c.points.empty()) std::memcpy(out->points.data(), c.points.data(), c.points.size() * 16);
  return out;
}
inline void to_pcl(const Cloud& c, pcl::PointCloud<pcl::PointXYZ>& out) {
  out.points.resize(c.points.size());
  if (!c.points.empty()) std::memcpy(out.points.data(), c.points.data(), c.points.size() * 16);
  out.width = (uint32_t)c.points.size();
  out.height = 1;
  out.is_dense = true;
}
inline void to_rowmajor(const Eigen::Matrix4d& T, double* T16) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) T16[4 * r + c] = T(r, c);
}

}  // namespace b2

// ---- ::IcpOdometer with the reference's signatures (include/icpslam/icp_odometer.h:30-58) ---------------------
class IcpOdometer : public b2::IcpOdometer {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  using Ptr = std::shared_ptr<IcpOdometer>;
  using PclCloud = pcl::PointCloud<pcl::PointXYZ>;

  IcpOdometer(const ros::NodeHandle& nh, const ros::NodeHandle& pnh)
      : b2::IcpOdometer(load(pnh)), nh_(nh), pnh_(pnh), tf_listener_(tf_buffer_) {
    pnh_.param("robot_frame", robot_frame_, std::string("base_link"));  // icp_odometer.cpp loadParameters
  }
  void getEstimates(ros::Time& stamp, PclCloud::Ptr& cloud, b2::Pose6DOF& latest_icp_transform, b2::Pose6DOF& icp_pose,
                    bool& new_transform) {
    double s = 0;
    b2::Cloud::Ptr c;
    b2::IcpOdometer::getEstimates(s, c, latest_icp_transform, icp_pose, new_transform);
    stamp.fromSec(s);
    if (!cloud) cloud.reset(new PclCloud());
    b2::to_pcl(*c, *cloud);
  }
  void voxelFilterCloud(PclCloud::Ptr* input, PclCloud::Ptr* output) {
    b2::Cloud::Ptr in = b2::from_pcl(**input), out(new b2::Cloud());
    b2::IcpOdometer::voxelFilterCloud(&in, &out);
    b2::to_pcl(*out, **output);
  }
  void publishPath(const ros::Time&) {}
  bool updateICPOdometry(const ros::Time& stamp, const Eigen::Matrix4d& T) {
    double T16[16];
    b2::to_rowmajor(T, T16);
    return b2::IcpOdometer::updateICPOdometry(stamp.toSec(), T16);
  }
  // icp_odometer.cpp:147-175: into the robot frame if need be (tf2::doTransform on the message), then fromROSMsg
  void laserCloudCallback(const sensor_msgs::PointCloud2::ConstPtr& cloud_msg) {
    b2::Cloud::Ptr input(new b2::Cloud());
    if (cloud_msg->header.frame_id != robot_frame_) {
      try {
        geometry_msgs::TransformStamped t =
            tf_buffer_.lookupTransform(robot_frame_, cloud_msg->header.frame_id, cloud_msg->header.stamp, ros::Duration(0.03));
        sensor_msgs::PointCloud2 cloud_out;
        tf2::doTransform(*cloud_msg, cloud_out, t);
        b2::fromROSMsg(engine(), b2::view_of(cloud_out), *input);
      } catch (tf2::TransformException& ex) {
        ROS_WARN("%s", ex.what());
      }
    } else {
      b2::fromROSMsg(engine(), b2::view_of(*cloud_msg), *input);
    }
    b2::IcpOdometer::laserCloudCallback(cloud_msg->header.stamp.toSec(), input);
  }

 private:
  static b2::IcpOdometerParams load(const ros::NodeHandle& pnh) {
    b2::IcpOdometerParams p;
    int skip = p.num_clouds_skip;
    double leaf = p.voxel_leaf_size;
    pnh.param("num_clouds_skip", skip, skip);          // config/icpslam.yaml
    pnh.param("voxel_leaf_size", leaf, leaf);
    p.num_clouds_skip = skip;
    p.voxel_leaf_size = leaf;
    return p;
  }
  ros::NodeHandle nh_, pnh_;
  std::string robot_frame_;
  tf2_ros::Buffer tf_buffer_;
  tf2_ros::TransformListener tf_listener_;
};

// ---- ::OctreeMapper with the reference's signatures (include/icpslam/octree_mapper.h:23-49) -------------------
class OctreeMapper : public b2::OctreeMapper {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  using Ptr = std::shared_ptr<OctreeMapper>;
  using PclCloud = pcl::PointCloud<pcl::PointXYZ>;

  OctreeMapper(const ros::NodeHandle& nh, const ros::NodeHandle& pnh) : b2::OctreeMapper(load(pnh)), nh_(nh), pnh_(pnh) {}
  void addPointsToMap(PclCloud::Ptr input_cloud) { b2::OctreeMapper::addPointsToMap(b2::from_pcl(*input_cloud)); }
  bool approxNearestNeighbors(const PclCloud::Ptr& cloud, PclCloud::Ptr& nearest_neighbors) {
    b2::Cloud::Ptr nn(new b2::Cloud());
    const bool ok = b2::OctreeMapper::approxNearestNeighbors(b2::from_pcl(*cloud), nn);
    if (!nearest_neighbors) nearest_neighbors.reset(new PclCloud());
    b2::to_pcl(*nn, *nearest_neighbors);
    return ok;
  }
  void transformCloudToPoseFrame(const PclCloud::Ptr& in_cloud, const b2::Pose6DOF& pose, PclCloud::Ptr& out_cloud) {
    b2::Cloud::Ptr out(new b2::Cloud());
    b2::OctreeMapper::transformCloudToPoseFrame(b2::from_pcl(*in_cloud), pose, out);
    if (!out_cloud) out_cloud.reset(new PclCloud());
    b2::to_pcl(*out, *out_cloud);
  }
  bool estimateTransformICP(const PclCloud::Ptr& curr_cloud, const PclCloud::Ptr& nn_cloud, b2::Pose6DOF& transform) {
    return b2::OctreeMapper::estimateTransformICP(b2::from_pcl(*curr_cloud), b2::from_pcl(*nn_cloud), transform);
  }
  bool refineTransformAndGrowMap(const ros::Time& stamp, const PclCloud::Ptr& cloud, const b2::Pose6DOF& prev_pose,
                                 b2::Pose6DOF& transform) {
    return b2::OctreeMapper::refineTransformAndGrowMap(stamp.toSec(), b2::from_pcl(*cloud), prev_pose, transform);
  }

 private:
  static b2::OctreeMapperParams load(const ros::NodeHandle& pnh) {
    b2::OctreeMapperParams p;
    double res = p.octree_resolution;
    pnh.param("octree_resolution", res, res);          // config/icpslam.yaml:17
    p.octree_resolution = res;
    return p;
  }
  ros::NodeHandle nh_, pnh_;
};

using Pose6DOF = b2::Pose6DOF;
#endif
#endif
