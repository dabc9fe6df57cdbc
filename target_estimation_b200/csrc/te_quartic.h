// te_quartic.h -- smallest real root of the interception quartic (src/intersection_solver.cpp:4-17,66-70).
//
// The reference hands the five coefficients (a0 first) to Eigen's PolynomialSolver (companion matrix -> balance ->
// shifted QR) and takes smallestRealRoot(found, 1e-10): the smallest real part among the roots with |imag| < 1e-10;
// -1 if there is none, if the t^4 coefficient is 0, or if that root is negative (NB: a negative smallest real root gives
// -1 even when a positive real root exists -- kept).
//
// Device path, one thread per query, real FP64 only on the fast path:
//   1. Ferrari: depress the monic quartic (x = y - b3/4), take the LARGEST real root z = s^2 of the resolvent cubic
//      z^3 + 2p z^2 + (p^2 - 4r) z - q^2 (closed form + two Newton steps), split
//      y^4 + p y^2 + q y + r = (y^2 + s y + u)(y^2 - s y + v),  u, v = (p + z -+ q/s) / 2  (the smaller one as r / larger);
//   2. the sign of each quadratic's discriminant classifies its pair of roots as real or complex; the smallest real
//      root is polished by Newton steps on the ORIGINAL coefficients until its residual is at rounding level, so the
//      closed form only has to land in the right basin;
//   3. anything suspicious (non-finite intermediate, resolvent residual above 1e-8, polish not converging) falls back
//      to Aberth-Ehrlich simultaneous iteration in complex FP64 with a backward-error stopping rule.
// Both paths and the reference's QR are backward stable; they can classify near-double roots differently
// (SURVEY.md H9).  Compiles for the host too (tests/quartic_check.cpp compares it with the oracle's QR restatement).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define TE_QHD __host__ __device__ __forceinline__
#define TE_QHD_NOINLINE static __host__ __device__ __noinline__
#else
#define TE_QHD inline
#define TE_QHD_NOINLINE inline
#endif
#ifndef TE_QUARTIC_COUNT_FALLBACK
#define TE_QUARTIC_COUNT_FALLBACK() ((void)0)   // tests/quartic_check.cpp counts how often the complex path runs
#endif

namespace te {

struct Cplx { double re, im; };
TE_QHD Cplx cadd(Cplx a, Cplx b) { return {a.re + b.re, a.im + b.im}; }
TE_QHD Cplx csub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
TE_QHD Cplx cmul(Cplx a, Cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
TE_QHD Cplx cdiv(Cplx a, Cplx b) {
  // Smith's algorithm
  if (fabs(b.re) >= fabs(b.im)) {
    double r = b.im / b.re, d = b.re + b.im * r;
    return {(a.re + a.im * r) / d, (a.im - a.re * r) / d};
  }
  double r = b.re / b.im, d = b.re * r + b.im;
  return {(a.re * r + a.im) / d, (a.im * r - a.re) / d};
}
TE_QHD double cabs2(Cplx a) { return a.re * a.re + a.im * a.im; }

// (Eigen) poly_eval: Horner for |x| <= 1, reversed Horner otherwise
TE_QHD double poly_abs4(const double c[5], Cplx x) {
  if (cabs2(x) <= 1.0) {
    Cplx v{c[4], 0.0};
#pragma unroll
    for (int i = 3; i >= 0; --i) v = cadd(cmul(v, x), Cplx{c[i], 0.0});
    return sqrt(cabs2(v));
  }
  Cplx inv = cdiv(Cplx{1.0, 0.0}, x);
  Cplx v{c[0], 0.0};
#pragma unroll
  for (int i = 1; i <= 4; ++i) v = cadd(cmul(v, inv), Cplx{c[i], 0.0});
  Cplx x2 = cmul(x, x), x4 = cmul(x2, x2);
  return sqrt(cabs2(cmul(x4, v)));
}

// the reference's selection rule over four roots, with Eigen 3.4's imaginary-noise clean-up first
TE_QHD double select_smallest_real(const double c[5], Cplx z[4]) {
  const double coarse_prec = 4096.0 * 2.220446049250313e-16;   // 4^(5+1) * eps
  bool found = false;
  double best = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (z[k].im != 0.0 && fabs(z[k].im) <= fabs(z[k].re) * coarse_prec) {
      Cplx r{z[k].re, 0.0};
      if (poly_abs4(c, r) <= poly_abs4(c, z[k])) z[k] = r;
    }
    if (fabs(z[k].im) < 1e-10) {
      if (!found) { found = true; best = z[k].re; }
      else if (z[k].re < best) best = z[k].re;
    }
  }
  return found ? best : -1.0;
}

// Fallback: Aberth-Ehrlich simultaneous iteration.  A root stops moving once its residual is at rounding level
// (|p(z)| <= 4 eps sum |c_i| |z|^i) or its correction is below 2 eps |z|.
TE_QHD_NOINLINE double lowest_real_root4_aberth(const double c[5]) {
  TE_QUARTIC_COUNT_FALLBACK();
  const double b3 = c[3] / c[4], b2 = c[2] / c[4], b1 = c[1] / c[4], b0 = c[0] / c[4];
  // Fujiwara bound for the starting circle
  double rad = fabs(b3);
  rad = fmax(rad, sqrt(fabs(b2)));
  rad = fmax(rad, cbrt(fabs(b1)));
  rad = fmax(rad, sqrt(sqrt(fabs(b0) * 0.5)));
  rad = 2.0 * rad;
  if (!(rad > 0.0) || !(rad < 1e300)) rad = 1.0;
  const double ctr = -b3 * 0.25;
  Cplx z[4];
  // exp(i (0.7 + k pi/2))
  const double cs0 = 0.7648421872844885, sn0 = 0.644217687237691;
  z[0] = {ctr + 0.5 * rad * cs0, 0.5 * rad * sn0};
  z[1] = {ctr - 0.5 * rad * sn0, 0.5 * rad * cs0};
  z[2] = {ctr - 0.5 * rad * cs0, -0.5 * rad * sn0};
  z[3] = {ctr + 0.5 * rad * sn0, -0.5 * rad * cs0};
  const double ac0 = fabs(c[0]), ac1 = fabs(c[1]), ac2 = fabs(c[2]), ac3 = fabs(c[3]), ac4 = fabs(c[4]);
  unsigned done = 0u;
  for (int iter = 0; iter < 100 && done != 15u; ++iter) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if ((done >> k) & 1u) continue;
      Cplx p{c[4], 0.0}, dp{0.0, 0.0};
#pragma unroll
      for (int i = 3; i >= 0; --i) {
        dp = cadd(cmul(dp, z[k]), p);
        p = cadd(cmul(p, z[k]), Cplx{c[i], 0.0});
      }
      const double az = sqrt(cabs2(z[k]));
      const double bound = (((ac4 * az + ac3) * az + ac2) * az + ac1) * az + ac0;
      if (cabs2(p) <= (8.9e-16 * bound) * (8.9e-16 * bound)) { done |= 1u << k; continue; }
      if (dp.re == 0.0 && dp.im == 0.0) { z[k].re += 1e-3 * rad; z[k].im += 1e-3 * rad; continue; }
      Cplx w = cdiv(p, dp);
      Cplx ssum{0.0, 0.0};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j == k) continue;
        Cplx d = csub(z[k], z[j]);
        if (d.re == 0.0 && d.im == 0.0) d = {1e-300, 1e-300};
        ssum = cadd(ssum, cdiv(Cplx{1.0, 0.0}, d));
      }
      Cplx den = csub(Cplx{1.0, 0.0}, cmul(w, ssum));
      Cplx dz = (den.re == 0.0 && den.im == 0.0) ? w : cdiv(w, den);
      z[k] = csub(z[k], dz);
      if (cabs2(dz) <= 1.9e-31 * cabs2(z[k])) done |= 1u << k;
    }
  }
  return select_smallest_real(c, z);
}

// one real quadratic factor y^2 + a y + b of the depressed quartic: updates the running minimum over real roots
// (in x = y - h) and notes a complex pair whose imaginary part is below the reference's 1e-10 threshold
TE_QHD void quad_factor_roots(double a, double b, double h, bool& found, double& best, bool& near_real) {
  const double disc = a * a - 4.0 * b;
  if (disc >= 0.0) {
    const double sq = sqrt(disc);
    const double t = -0.5 * (a + (a >= 0.0 ? sq : -sq));   // the root of larger magnitude
    double y1 = t, y2 = (t != 0.0) ? b / t : 0.0;
    const double lo = fmin(y1, y2) - h;
    if (!found || lo < best) best = lo;
    found = true;
  } else {
    const double im = 0.5 * sqrt(-disc);
    const double re = -0.5 * a - h;
    if (im < 1e-10 || im <= fabs(re) * (4096.0 * 2.220446049250313e-16)) near_real = true;   // H9 zone: let the complex path decide
  }
}

TE_QHD double lowest_real_root4(const double c[5]) {
  if (!(fabs(c[4]) > 0.0)) return -1.0;   // src/intersection_solver.cpp:9
  const double b3 = c[3] / c[4], b2 = c[2] / c[4], b1 = c[1] / c[4], b0 = c[0] / c[4];
  const double h = 0.25 * b3, h2 = h * h;
  const double p = b2 - 6.0 * h2;
  const double q = b1 - 2.0 * b2 * h + 8.0 * h2 * h;
  const double r = b0 - b1 * h + b2 * h2 - 3.0 * h2 * h2;
  // resolvent cubic g(z) = z^3 + A z^2 + B z + C, largest real root (>= 0 because g(0) = -q^2 <= 0)
  const double A = 2.0 * p, B = p * p - 4.0 * r, C = -q * q;
  const double third = 1.0 / 3.0;
  const double P = B - A * A * third;
  const double Qc = (2.0 / 27.0) * A * A * A - third * A * B + C;
  const double disc = 0.25 * Qc * Qc + (P * third) * (P * third) * (P * third);
  double t;
  if (disc <= 0.0 && P < 0.0) {   // three real roots: the k = 0 branch of the trigonometric form is the largest
    const double m = sqrt(-P * third);
    double arg = (3.0 * Qc) / (2.0 * P * m);   // = (3 Qc / 2P) sqrt(-3 / P)
    arg = fmin(1.0, fmax(-1.0, arg));
    t = 2.0 * m * cos(acos(arg) * third);
  } else {                         // one real root (Cardano, cancellation-free form)
    const double sd = sqrt(fmax(disc, 0.0));
    const double u3 = -0.5 * Qc + (Qc > 0.0 ? -sd : sd);
    const double u = cbrt(u3);
    t = (u != 0.0) ? u - P / (3.0 * u) : 0.0;
  }
  double z = t - A * third;
#pragma unroll
  for (int it = 0; it < 2; ++it) {   // Newton on g
    const double g = ((z + A) * z + B) * z + C;
    const double dg = (3.0 * z + 2.0 * A) * z + B;
    if (dg > 0.0) z -= g / dg;
  }
  if (!(z > 0.0)) z = 0.0;
  const double s = sqrt(z);
  const double pz = p + z;
  const double w2 = pz * pz - 4.0 * r;   // = (q / s)^2 when z solves the cubic
  double w;
  if (z > 1e-12 * (fabs(p) + sqrt(fabs(r)))) {
    w = q / s;
    // resolvent residual: the closed form must satisfy the cubic to ~1e-8, otherwise do not trust the split
    if (!(fabs(w * w - w2) <= 1e-8 * (pz * pz + 4.0 * fabs(r) + w * w))) return lowest_real_root4_aberth(c);
  } else {
    w = (q >= 0.0) ? sqrt(fmax(w2, 0.0)) : -sqrt(fmax(w2, 0.0));   // (near-)biquadratic: s -> 0, q / s -> +-sqrt(w2)
    if (!(w2 >= -1e-8 * (pz * pz + 4.0 * fabs(r)))) return lowest_real_root4_aberth(c);
  }
  // u v = r: the larger of the two by the sum formula, the other by division
  double u, v;
  if (pz * w <= 0.0) {   // |pz - w| >= |pz + w|
    u = 0.5 * (pz - w);
    v = (u != 0.0) ? r / u : 0.5 * (pz + w);
  } else {
    v = 0.5 * (pz + w);
    u = (v != 0.0) ? r / v : 0.5 * (pz - w);
  }
  if (!(fabs(u) < 1e300) || !(fabs(v) < 1e300) || !(fabs(s) < 1e300)) return lowest_real_root4_aberth(c);   // also catches NaN
  bool found = false, near_real = false;
  double best = 0.0;
  quad_factor_roots(s, u, h, found, best, near_real);
  quad_factor_roots(-s, v, h, found, best, near_real);
  if (near_real) return lowest_real_root4_aberth(c);
  if (!found) return -1.0;
  // Newton polish on the original coefficients
  const double ac0 = fabs(c[0]), ac1 = fabs(c[1]), ac2 = fabs(c[2]), ac3 = fabs(c[3]), ac4 = fabs(c[4]);
  double x = best;
  bool ok = false;
#pragma unroll 1
  for (int it = 0; it < 6; ++it) {
    const double f = (((c[4] * x + c[3]) * x + c[2]) * x + c[1]) * x + c[0];
    const double ax = fabs(x);
    const double bound = (((ac4 * ax + ac3) * ax + ac2) * ax + ac1) * ax + ac0;
    if (fabs(f) <= 8.9e-16 * bound) { ok = true; break; }
    const double df = ((4.0 * c[4] * x + 3.0 * c[3]) * x + 2.0 * c[2]) * x + c[1];
    if (df == 0.0) break;
    const double dx = f / df;
    x -= dx;
    if (fabs(dx) <= 2.3e-16 * fabs(x)) { ok = true; break; }
  }
  if (!ok || !(fabs(x - best) <= 1e-6 * (fabs(best) + fabs(h) + s))) return lowest_real_root4_aberth(c);
  return x;
}

}  // namespace te
