// host check of te::lowest_real_root4 (csrc/te_quartic.h: Ferrari + Newton polish, Aberth fallback) against the oracle's
// restatement of Eigen's PolynomialSolver (companion matrix -> balance -> QR), built and run by tests/test_quartic.py
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <initializer_list>
static long long g_fallbacks = 0;
#define TE_QUARTIC_COUNT_FALLBACK() (++g_fallbacks)
#include "te_quartic.h"

extern "C" double orc_lowest_real_root(const double* coeffs, int ncoef);
extern "C" int orc_poly_roots(const double* coeffs, int ncoef, double* re, double* im);

static uint64_t s = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double uni() { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }
static inline double uab(double a, double b) { return a + (b - a) * uni(); }

// separation of the reference's roots: min distance between two distinct roots (conjugates count), relative to their size.
// Near-multiple roots are ill-conditioned: two backward-stable solvers may legitimately disagree there.
static double root_separation(const double c[5]) {
  double re[4], im[4];
  int n = orc_poly_roots(c, 5, re, im);
  double sep = 1e300, mag = 0;
  for (int i = 0; i < n; ++i) {
    mag = std::fmax(mag, std::hypot(re[i], im[i]));
    for (int j = i + 1; j < n; ++j) sep = std::fmin(sep, std::hypot(re[i] - re[j], im[i] - im[j]));
  }
  return sep / std::fmax(mag, 1e-300);
}

struct Stat { long long n = 0, bad = 0, bad_illcond = 0, boundary = 0, found = 0; double worst = 0; };

static void check(const double c[5], Stat& st, const char* what) {
  const double got = te::lowest_real_root4(c), ref = orc_lowest_real_root(c, 5);
  // raw values are compared (stricter than what the solver returns: the caller reports every negative root as -1,
  // src/intersection_solver.cpp:83-86); "no real root" is -1 in both
  const double g = got, r = ref;
  ++st.n;
  if (r >= 0) ++st.found;
  const double err = std::fabs(g - r) / std::fmax(1.0, std::fabs(r));
  if (err <= 1e-9) { st.worst = std::fmax(st.worst, err); return; }
  // a root AT zero (c0 = 0): rounding noise alone decides between "0" and "negative -> -1" in either solver
  if (std::fabs(ref) < 1e-9 && std::fabs(got - ref) <= 1e-9) { ++st.boundary; return; }
  // near-multiple roots: neither the classification real / complex nor the value is determined at 1e-9.  Two signs of
  // it: the reference's own roots lie within 1e-3 of each other (generator 3 makes such pairs on purpose), or the
  // condition number of the returned root, sum |c_i| |x|^i / (|x| |p'(x)|), exceeds 1e-11 / eps
  const double sep = root_separation(c);
  bool ill = sep < 1e-3;
  for (double x : {got, ref}) {
    if (x == -1.0) continue;
    const double ax = std::fabs(x);
    const double bound = (((std::fabs(c[4]) * ax + std::fabs(c[3])) * ax + std::fabs(c[2])) * ax + std::fabs(c[1])) * ax + std::fabs(c[0]);
    const double dp = std::fabs(((4 * c[4] * x + 3 * c[3]) * x + 2 * c[2]) * x + c[1]);
    if (2.2e-16 * bound > 1e-11 * ax * dp) ill = true;
  }
  if (ill) { ++st.bad_illcond; return; }
  if (st.bad < 10) std::printf("MISMATCH (%s) c=[%.17g, %.17g, %.17g, %.17g, %.17g] got=%.17g want=%.17g sep=%.3g\n", what, c[0], c[1], c[2], c[3], c[4], got, ref, sep);
  ++st.bad;
}

static void from_roots(const double re[4], const double im[4], double lead, double c[5]) {
  // two real quadratics (x^2 + a1 x + b1)(x^2 + a2 x + b2); im[0] = -im[1], im[2] = -im[3] or all real
  double a1, b1, a2, b2;
  if (im[0] != 0) { a1 = -2 * re[0]; b1 = re[0] * re[0] + im[0] * im[0]; } else { a1 = -(re[0] + re[1]); b1 = re[0] * re[1]; }
  if (im[2] != 0) { a2 = -2 * re[2]; b2 = re[2] * re[2] + im[2] * im[2]; } else { a2 = -(re[2] + re[3]); b2 = re[2] * re[3]; }
  c[4] = lead; c[3] = lead * (a1 + a2); c[2] = lead * (b1 + b2 + a1 * a2); c[1] = lead * (a1 * b2 + a2 * b1); c[0] = lead * b1 * b2;
}

int main(int argc, char** argv) {
  const long long N = argc > 1 ? std::atoll(argv[1]) : 200000;
  Stat phys, rnds, roots, edge;
  // 1. the solver's own coefficients (src/intersection_solver.cpp:66-70) for plausible target states
  for (long long i = 0; i < N; ++i) {
    double p[3], v[3], a[3];
    for (int k = 0; k < 3; ++k) { p[k] = uab(-3, 3); v[k] = uab(-2, 2); a[k] = uab(-0.5, 0.5); }
    a[2] += (i % 3 == 0) ? 0.0 : -9.81;
    if (i % 5 == 0) for (int k = 0; k < 3; ++k) a[k] *= 1e-3;      // tiny t^4 coefficient
    if (i % 7 == 0) for (int k = 0; k < 3; ++k) v[k] *= 1e-4;      // nearly at rest
    const double R = uab(0.1, (i % 2) ? 1.0 : 4.0);
    double c[5];
    c[4] = 0.25 * (a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    c[3] = v[0] * a[0] + v[1] * a[1] + v[2] * a[2];
    c[2] = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + p[0] * a[0] + p[1] * a[1] + p[2] * a[2];
    c[1] = 2 * (p[0] * v[0] + p[1] * v[1] + p[2] * v[2]);
    c[0] = p[0] * p[0] + p[1] * p[1] + p[2] * p[2] - R * R;
    check(c, phys, "physical");
  }
  // 2. random coefficients over several decades
  for (long long i = 0; i < N; ++i) {
    double c[5];
    for (int k = 0; k < 5; ++k) c[k] = (uni() - 0.5) * std::pow(10.0, uab(-3, 3));
    if (i % 9 == 0) c[3] = 0; if (i % 13 == 0) c[1] = 0; if (i % 17 == 0) { c[3] = 0; c[1] = 0; }   // incl. biquadratics
    check(c, rnds, "random");
  }
  // 3. prescribed roots: 4 real / 2 real + pair / 2 pairs, incl. close and repeated ones
  for (long long i = 0; i < N; ++i) {
    double re[4], im[4] = {0, 0, 0, 0}, c[5];
    for (int k = 0; k < 4; ++k) re[k] = uab(-3, 6) * ((i % 4 == 0) ? 100.0 : 1.0);
    const int kind = (int)(i % 3);
    if (kind >= 1) { re[1] = re[0]; im[0] = uab(1e-3, 2.0); im[1] = -im[0]; }
    if (kind == 2) { re[3] = re[2]; im[2] = uab(1e-3, 2.0); im[3] = -im[2]; }
    if (i % 11 == 0 && kind == 0) re[1] = re[0] + uab(0, 1e-3);          // close real pair
    if (i % 23 == 0 && kind == 0) re[1] = re[0];                           // double root
    if (i % 29 == 0 && kind == 0) { re[1] = re[0]; re[3] = re[2]; }        // two double roots
    if (i % 31 == 0) { re[0] = 0; }                                        // root at zero
    from_roots(re, im, uab(0.1, 30.0), c);
    check(c, roots, "from roots");
  }
  // 4. edge cases
  const double E[][5] = {{1, 0, 0, 0, 0}, {0, 0, 0, 0, 1}, {-1, 0, 0, 0, 1}, {1, 0, 0, 0, 1}, {1, 0, 2, 0, 1}, {1, 0, -2, 0, 1}, {0, 0, 0, 1, 1}, {0, 1, 0, 0, 1},
                         {24, -50, 35, -10, 1}, {1e-30, 0, 0, 0, 1}, {5, 1, 1, 1, 0}, {5, 1, 1, 1, 1e-300}, {-4, 0, 0, 0, 1e-12}, {3, 2, 1, 1e-9, 24.06}};
  for (auto& e : E) check(e, edge, "edge");
  {   // quadruple root (x - 1)^4: the computed roots scatter by eps^(1/4); any answer within that of 1, or -1, is legitimate
    const double c4[5] = {1, -4, 6, -4, 1};
    const double g = te::lowest_real_root4(c4);
    if (!(g == -1.0 || std::fabs(g - 1.0) < 1e-3)) { std::printf("MISMATCH quadruple root: %.17g\n", g); ++edge.bad; }
  }
  const Stat* all[] = {&phys, &rnds, &roots, &edge};
  const char* names[] = {"physical", "random", "from-roots", "edge"};
  long long bad = 0;
  for (int k = 0; k < 4; ++k) {
    std::printf("%-10s: %lld cases, %lld with a real root >= 0, %lld mismatches, %lld ill-conditioned (near-multiple root) disagreements, %lld roots at 0, worst agreeing err %.3g\n",
                names[k], all[k]->n, all[k]->found, all[k]->bad, all[k]->bad_illcond, all[k]->boundary, all[k]->worst);
    bad += all[k]->bad;
  }
  std::printf("fallbacks to the complex iteration: %lld\n", g_fallbacks);
  std::printf("%lld mismatches\n", bad);
  return bad ? 1 : 0;
}
