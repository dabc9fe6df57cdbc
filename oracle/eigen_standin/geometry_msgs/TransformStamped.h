// oracle/eigen_standin: geometry_msgs/TransformStamped as the reference reads it (test infrastructure, not ROS)
#pragma once
#include <cstdint>
#include <memory>
#include <string>
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct Stamp { uint32_t sec = 0, nsec = 0; };
struct Header { uint32_t seq = 0; Stamp stamp; std::string frame_id; };
struct TransformStamped { Header header; std::string child_frame_id; Transform transform; };
}  // namespace geometry_msgs
