// oracle/eigen_standin: included by target_manager_ros.hpp, nothing of it is used (test infrastructure, not ROS)
#pragma once
