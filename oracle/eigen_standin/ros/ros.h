// oracle/eigen_standin/ros/ros.h -- TEST INFRASTRUCTURE, not ROS.
// roscpp is absent from this image.  The reference's ROS adapter (src/target_manager_ros.cpp) uses: ros::NodeHandle
// (getParam of a double list / a string, subscribe with a member callback), ros::Subscriber, ros::Time (now(), toSec()) and
// the ROS_*_STREAM log macros.  This stand-in keeps parameters in a map, remembers the subscribed callback so that a test can
// deliver messages to it (the reference's callback is private), and takes "now" from a test-settable clock.  ros::Time::toSec()
// is ROS's own formula, (double)sec + 1e-9 * (double)nsec.
#pragma once
#include <cstdint>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_WARN_STREAM(x) do { } while (0)
#define ROS_ERROR_STREAM(x) do { } while (0)
#define ROS_DEBUG_STREAM(x) do { } while (0)

namespace ros {

struct Time {
  uint32_t sec = 0, nsec = 0;
  Time() {}
  Time(uint32_t s, uint32_t n) : sec(s), nsec(n) {}
  double toSec() const { return static_cast<double>(sec) + 1e-9 * static_cast<double>(nsec); }
  static Time& clock() { static Time t; return t; }     // set by the test
  static Time now() { return clock(); }
};

class Subscriber {};

class NodeHandle {
 public:
  struct State {
    std::map<std::string, std::vector<double>> lists;
    std::map<std::string, std::string> strings;
    std::function<void(const std::shared_ptr<const void>&)> callback;   // type-erased subscriber
  };
  NodeHandle() : st_(std::make_shared<State>()) {}
  bool getParam(const std::string& key, std::vector<double>& out) const {
    auto it = st_->lists.find(key);
    if (it == st_->lists.end()) return false;
    out = it->second;
    return true;
  }
  bool getParam(const std::string& key, std::string& out) const {
    auto it = st_->strings.find(key);
    if (it == st_->strings.end()) return false;
    out = it->second;
    return true;
  }
  template <class M, class T>
  Subscriber subscribe(const std::string& /*topic*/, uint32_t /*queue*/, void (T::*fp)(const std::shared_ptr<const M>&), T* obj) {
    st_->callback = [fp, obj](const std::shared_ptr<const void>& m) { (obj->*fp)(std::static_pointer_cast<const M>(m)); };
    return Subscriber();
  }
  std::shared_ptr<State> st_;   // shared between copies, like a real NodeHandle's connection to the node
};

}  // namespace ros
