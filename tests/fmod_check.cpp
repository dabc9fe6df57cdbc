// host check of te::fmod_exact against libm fmod (bit-exact), built and run by tests/test_fmod_exact.py
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include "te_fmod.h"

static uint64_t s = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double uni() { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t bits(double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; }

int main() {
  long long bad = 0, n = 0;
  const double PI = 3.14159265358979323846;
  const double ys[] = {2 * PI, PI, -2 * PI, 1.0, 0.1, 3.0, 1e-3, 7.25};
  for (double y : ys) {
    for (int i = 0; i < 4000000; ++i) {
      double mag = std::pow(10.0, -6.0 + 14.0 * uni());
      double x = (uni() - 0.5) * 2.0 * mag;
      if (i % 7 == 0) x = std::nearbyint(x / y) * y;                       // near exact multiples
      if (i % 11 == 0) x = std::nextafter(std::nearbyint(x / y) * y, (i & 1) ? 1e300 : -1e300);
      double a = te::fmod_exact(x, y), b = std::fmod(x, y);
      ++n;
      if (bits(a) != bits(b)) { if (bad < 10) std::printf("MISMATCH x=%.17g y=%.17g got=%.17g want=%.17g\n", x, y, a, b); ++bad; }
    }
  }
  const double edge[][2] = {{0.0, 1.0}, {-0.0, 1.0}, {5.0, 5.0}, {-5.0, 5.0}, {1e300, 3.0}, {INFINITY, 2.0}, {1.0, INFINITY}, {1.0, 0.0}, {NAN, 1.0},
                            {4.0e15, 2 * PI}, {-9.007199254740992e15, 6.283185307179586}, {1e-310, 1e-320}};
  for (auto& e : edge) {
    double a = te::fmod_exact(e[0], e[1]), b = std::fmod(e[0], e[1]);
    ++n;
    bool same = (std::isnan(a) && std::isnan(b)) || bits(a) == bits(b);
    if (!same) { std::printf("EDGE MISMATCH x=%.17g y=%.17g got=%.17g want=%.17g\n", e[0], e[1], a, b); ++bad; }
  }
  std::printf("checked %lld pairs, %lld mismatches\n", n, bad);
  return bad ? 1 : 0;
}
