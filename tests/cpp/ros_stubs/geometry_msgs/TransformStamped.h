#pragma once
#include <sensor_msgs/PointCloud2.h>
namespace geometry_msgs { struct TransformStamped { std_msgs::Header header; std::string child_frame_id; double t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}; }; }
