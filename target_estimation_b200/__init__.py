"""B200-native batched Kalman-filter hot path of graiola/target_estimation.

The product is the CUDA library `lib/libte_pool.so` (include/te_pool.h) and the reference-facing
host surface `lib/libtarget_c.so` (include/target_manager_c.h, include/target_manager.hpp).  This
package is only the Python harness over their C-ABIs (ctypes) for tests and bench.py: it never
computes on the CPU and raises at import of a symbol if the CUDA library is missing.
"""
from ._lib import lib, lib_path, TeError  # noqa: F401
from .pool import (TargetPool, IntersectionSolver, load_model, MODEL_TYPES, ANGULAR_RATES, ANGULAR_VELOCITIES,  # noqa: F401
                   UNIFORM_ACCELERATION, UNIFORM_VELOCITY, ACT_NONE, ACT_PREDICT, ACT_UPDATE, model_dims,
                   bytes_per_step)
