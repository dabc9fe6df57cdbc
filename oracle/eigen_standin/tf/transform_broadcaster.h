// oracle/eigen_standin/tf/transform_broadcaster.h -- TEST INFRASTRUCTURE, not tf.
// What src/target_manager_ros.cpp uses of tf: Vector3, Quaternion (normalize() = divide by the length, as tf / Bullet do),
// Transform (setOrigin / setRotation), StampedTransform, TransformBroadcaster::sendTransform -- which here appends to a log the
// test reads back (the filtered poses the node would publish).
#pragma once
#include <cmath>
#include <string>
#include <vector>
#include "geometry_msgs/TransformStamped.h"
#include "ros/ros.h"
namespace tf {
struct Vector3 {
  double x_ = 0, y_ = 0, z_ = 0;
  Vector3() {}
  Vector3(double x, double y, double z) : x_(x), y_(y), z_(z) {}
};
struct Quaternion {
  double x_ = 0, y_ = 0, z_ = 0, w_ = 1;
  Quaternion() {}
  Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
  Quaternion& normalize() {   // tf: *this /= length()
    const double s = 1.0 / std::sqrt(x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_);
    x_ *= s; y_ *= s; z_ *= s; w_ *= s;
    return *this;
  }
};
struct Transform {
  Vector3 origin;
  Quaternion rotation;
  void setOrigin(const Vector3& v) { origin = v; }
  void setRotation(const Quaternion& q) { rotation = q; }
};
struct StampedTransform : Transform {
  ros::Time stamp;
  std::string frame_id, child_frame_id;
  StampedTransform(const Transform& t, const ros::Time& s, const std::string& f, const std::string& c) : Transform(t), stamp(s), frame_id(f), child_frame_id(c) {}
};
class TransformBroadcaster {
 public:
  void sendTransform(const StampedTransform& t) { log().push_back(t); }
  static std::vector<StampedTransform>& log() { static std::vector<StampedTransform> l; return l; }
};
}  // namespace tf
