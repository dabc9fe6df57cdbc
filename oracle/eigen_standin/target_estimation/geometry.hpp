// oracle/eigen_standin/target_estimation/geometry.hpp -- TEST INFRASTRUCTURE.
// The reference's src/kalman.cpp includes target_estimation/geometry.hpp but uses nothing from it; the real header needs
// Eigen's Geometry module, which the stand-in of Eigen/Dense in this directory does not provide.  This empty header shadows
// it for that one translation unit only (oracle/Makefile puts this directory first on the include path); kalman.hpp itself
// comes from /root/reference/include.
#pragma once
