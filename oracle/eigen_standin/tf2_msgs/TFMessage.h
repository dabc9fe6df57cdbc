// oracle/eigen_standin: tf2_msgs/TFMessage (test infrastructure, not ROS); ConstPtr is a std::shared_ptr here
#pragma once
#include <memory>
#include <vector>
#include "geometry_msgs/TransformStamped.h"
namespace tf2_msgs {
struct TFMessage {
  std::vector<geometry_msgs::TransformStamped> transforms;
  typedef std::shared_ptr<const TFMessage> ConstPtr;
};
}  // namespace tf2_msgs
