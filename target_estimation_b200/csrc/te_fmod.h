// te_fmod.h -- exact fmod without the iterative reduction (host + device; the host build exists for tests/test_fmod_exact.py)
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define TE_HD __host__ __device__ __forceinline__
#else
#define TE_HD inline
#endif

namespace te {

// fmod(x, y), bit-identical to the IEEE result, for the angle arithmetic of geometry.hpp:31-88.  The remainder
// r = x - q*y with q = trunc(x / y) is exactly representable, so one FMA returns it exactly; q can only be wrong when
// x / y rounds up to an integer the true quotient does not reach, which shows as a remainder of the wrong sign (or
// |r| >= |y|) and is repaired by one more exact FMA.  Quotients beyond 2^51, infinities and NaNs go to the library
// routine.  C's fmod is exact, so agreeing with it bit for bit keeps every unwrap / wrap branch identical to the
// reference's.
TE_HD double fmod_exact(double x, double y) {
  const double ax = fabs(x), ay = fabs(y);
  if (!(ax < 2.0e15 * ay) || !(ay > 1.0e-300) || !(ay < 1.0e300)) return fmod(x, y);
#if defined(__CUDA_ARCH__) || defined(TE_FMOD_RECIPROCAL)
  // device: y is a compile-time constant (2 pi, pi) at every call site, so 1 / y folds and the quotient costs one
  // multiplication instead of a ~30-instruction division; an estimate that is off by one either way is repaired below
  double q = trunc(x * (1.0 / y));
#else
  double q = trunc(x / y);
#endif
  double r = fma(-q, y, x);
  const double toward_zero = ((x < 0.0) == (y < 0.0)) ? 1.0 : -1.0;   // sign of the true quotient
  if (r != 0.0 && ((r < 0.0) != (x < 0.0))) {   // q one step too far from zero
    q -= toward_zero;
    r = fma(-q, y, x);
  } else if (fabs(r) >= ay) {                   // q one step short (possible with the reciprocal estimate)
    q += toward_zero;
    r = fma(-q, y, x);
  }
  if (r == 0.0) r = copysign(0.0, x);           // fmod returns a zero with the sign of x
  return r;
}

}  // namespace te
