#pragma once
#include <stdexcept>
#include <geometry_msgs/TransformStamped.h>
namespace tf2 { struct TransformException : std::runtime_error { using std::runtime_error::runtime_error; }; }
namespace tf2_ros {
class Buffer {
 public:
  geometry_msgs::TransformStamped lookupTransform(const std::string& target, const std::string& source, const ros::Time&,
                                                  const ros::Duration&) const {
    if (source == "unknown") throw tf2::TransformException("no transform from " + source + " to " + target);
    return geometry_msgs::TransformStamped();
  }
};
class TransformListener { public: explicit TransformListener(Buffer&) {} };
}  // namespace tf2_ros
