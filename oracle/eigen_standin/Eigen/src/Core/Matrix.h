// oracle/eigen_standin: intersection_solver.hpp includes this internal Eigen header directly; everything is in Eigen/Dense here
#pragma once
#include "../../Dense"
