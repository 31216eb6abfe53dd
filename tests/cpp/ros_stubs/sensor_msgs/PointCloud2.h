#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include <ros/ros.h>
namespace std_msgs { struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; }; }
namespace sensor_msgs {
struct PointField { std::string name; uint32_t offset = 0; uint8_t datatype = 7; uint32_t count = 1; };
struct PointCloud2 {
  typedef std::shared_ptr<PointCloud2 const> ConstPtr;
  std_msgs::Header header;
  uint32_t height = 1, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = true;
};
}  // namespace sensor_msgs
