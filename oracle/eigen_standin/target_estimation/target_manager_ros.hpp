// oracle/eigen_standin/target_estimation/target_manager_ros.hpp -- TEST INFRASTRUCTURE.
// intersection_solver.hpp includes the ROS adapter header (ros/ros.h, tf, ...) although it only uses TargetManager; ROS is
// absent from this image.  This header shadows it for that translation unit (oracle/Makefile puts this directory first on the
// include path); target_manager.hpp itself comes from /root/reference/include.
#pragma once
#include "target_estimation/target_manager.hpp"
