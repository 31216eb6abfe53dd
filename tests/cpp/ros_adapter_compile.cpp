// Compile (and, on a GPU box, run) check of the ROS-facing half of icpslam_b200/shims/b2icp_ros_adapter.hpp against the
// stand-in headers under tests/cpp/ros_stubs: ::IcpOdometer / ::OctreeMapper are instantiated with the reference's
// constructor signatures and every adapter method is referenced with the reference's argument types.
#include "../../icpslam_b200/shims/b2icp_ros_adapter.hpp"

#ifndef B2ICP_HAVE_ROS
#error "the stand-in ROS headers were not found: compile with -I tests/cpp/ros_stubs"
#endif

#include <cstdio>
#include <cstring>

static sensor_msgs::PointCloud2::ConstPtr make_msg(int n, const char* frame, double stamp) {
  auto m = std::make_shared<sensor_msgs::PointCloud2>();
  m->header.frame_id = frame;
  m->header.stamp.fromSec(stamp);
  m->width = (uint32_t)n;
  m->point_step = 16;
  m->row_step = 16 * (uint32_t)n;
  const char* names[3] = {"x", "y", "z"};
  for (int k = 0; k < 3; ++k) {
    sensor_msgs::PointField f;
    f.name = names[k];
    f.offset = 4u * (uint32_t)k;
    m->fields.push_back(f);
  }
  m->data.resize(16u * (size_t)n);
  for (int i = 0; i < n; ++i) {
    const float p[4] = {0.01f * (float)(i % 97), 0.02f * (float)(i % 53), 0.005f * (float)(i % 31), 0.f};
    std::memcpy(m->data.data() + 16 * (size_t)i, p, 16);
  }
  return m;
}

int main(int argc, char** argv) {
  // without arguments: compile / link check only (no GPU needed).  "run": the glue end to end on a GPU box (by hand:
  // the round's GPU budget ended before this mode could be made a -m gpu test)
  if (argc < 2 || std::strcmp(argv[1], "run") != 0) {
    std::printf("compiled\n");
    return 0;
  }
  ros::NodeHandle nh, pnh("~");
  try {
    IcpOdometer::Ptr odom(new IcpOdometer(nh, pnh));
    OctreeMapper::Ptr mapper(new OctreeMapper(nh, pnh));
    Pose6DOF start;
    start.setIdentity();
    odom->setInitialPose(start);                                 // the node does this from the first robot odometry message
    odom->laserCloudCallback(make_msg(2000, "unknown", 0.05));   // lookupTransform throws: warned, the empty scan is dropped
    for (int k = 0; k < 16 && !odom->isOdomReady(); ++k)         // (num_clouds_skip scans are skipped between registrations)
      odom->laserCloudCallback(make_msg(2000, k % 2 ? "velodyne" : "base_link", 0.1 + 0.1 * k));  // odd ones through tf2::doTransform
    if (!odom->isOdomReady()) {
      std::fprintf(stderr, "adapter: the odometer never accepted a scan (status %d)\n", odom->last_status);
      return 4;
    }
    ros::Time stamp;
    IcpOdometer::PclCloud::Ptr cloud;
    Pose6DOF t, pose;
    bool fresh = false;
    odom->getEstimates(stamp, cloud, t, pose, fresh);
    odom->updateICPOdometry(ros::Time(0.4), Eigen::Matrix4d::Identity());
    IcpOdometer::PclCloud::Ptr filtered(new IcpOdometer::PclCloud());
    odom->voxelFilterCloud(&cloud, &filtered);
    mapper->addPointsToMap(cloud);
    OctreeMapper::PclCloud::Ptr nn, moved;
    mapper->approxNearestNeighbors(cloud, nn);
    mapper->transformCloudToPoseFrame(cloud, pose, moved);
    Pose6DOF refined;
    mapper->estimateTransformICP(cloud, nn, refined);
    mapper->refineTransformAndGrowMap(ros::Time(0.5), cloud, pose, refined);
    std::printf("ran: %zu points back, %zu after the voxel filter, %zu neighbours\n", cloud->points.size(),
                filtered->points.size(), nn->points.size());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "adapter: %s\n", e.what());
    return 3;
  }
  return 0;
}
