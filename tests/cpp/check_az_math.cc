// tests/cpp/check_az_math.cc — pins az_math.h (product header, host build) against the live glibc.
// usage: check_az_math [quick|full]
//   logf : every float in (0, inf)            (full)  / every 97th (quick)
//   expf : every float in [-110, 90]          (full)  / every 97th (quick)
//   powf : x in (0,1] sampled, y from the exponents the reference uses + random y in (0, 1e5)
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "az_math.h"

using namespace b2az;
static bool same(float a, float b) {
  if (std::isnan(a) && std::isnan(b)) return true;
  return f2u(a) == f2u(b);
}
int main(int argc, char** argv) {
  const bool full = argc > 1 && !strcmp(argv[1], "full");
  const unsigned step = full ? 1 : 97;
  const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  std::atomic<unsigned long long> bad_log{0}, bad_exp{0}, bad_pow{0}, n_log{0}, n_exp{0}, n_pow{0};
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) {
    th.emplace_back([&, t] {
      unsigned long long bl = 0, be = 0, bp = 0, nl = 0, ne = 0, np = 0;
      // logf over all non-negative bit patterns incl. 0, subnormals, inf, nan, plus a few negatives
      for (unsigned long long u = t * (unsigned long long)step; u <= 0x7fffffffull + 1000; u += nt * (unsigned long long)step) {
        float x = u2f((u32)u);
        if (!same(az_logf(x), logf(x))) { if (bl < 3) fprintf(stderr, "logf mismatch x=%a got %a want %a\n", x, az_logf(x), logf(x)); ++bl; }
        ++nl;
      }
      // expf over [-110, 90]
      for (int sign = 0; sign < 2; ++sign) {
        const u32 hi = f2u(sign ? 110.0f : 90.0f);
        for (unsigned long long u = t * (unsigned long long)step; u <= hi; u += nt * (unsigned long long)step) {
          float x = u2f((u32)u | (sign ? 0x80000000u : 0u));
          if (!same(az_expf(x), expf(x))) { if (be < 3) fprintf(stderr, "expf mismatch x=%a got %a want %a\n", x, az_expf(x), expf(x)); ++be; }
          ++ne;
        }
      }
      // powf
      std::mt19937_64 g(1234 + t);
      const float ys[] = {0.8f, 1.0f / 1.25f, 1.0f / 1.4f, 1.0f, 2.0f, 5.0f, 1.0f / 0.2f, 1.0f / 0.5f, 1.0f / 1.547f, 7.0f / 10.83f, 1.0f / (10.83f / 7.0f), 0.5f, 3.0f, 0.0f, -1.0f, 92336.1f};
      const unsigned long long iters = full ? 40000000ull : 2000000ull;
      for (unsigned long long it = 0; it < iters; ++it) {
        u32 bits = (u32)g();
        float x = u2f(bits % 0x3f800001u);  // [0, 1]
        float y = (it & 1) ? ys[(it >> 1) % (sizeof(ys) / sizeof(ys[0]))] : u2f((u32)(g() % 0x47c35000u));  // [0, 1e5)
        if (!same(az_powf(x, y), powf(x, y))) { if (bp < 3) fprintf(stderr, "powf mismatch x=%a y=%a got %a want %a\n", x, y, az_powf(x, y), powf(x, y)); ++bp; }
        // also arbitrary finite x (incl. negative, >1)
        float x2 = u2f(bits);
        if (!same(az_powf(x2, y), powf(x2, y))) { if (bp < 6) fprintf(stderr, "powf mismatch x=%a y=%a got %a want %a\n", x2, y, az_powf(x2, y), powf(x2, y)); ++bp; }
        np += 2;
      }
      bad_log += bl; bad_exp += be; bad_pow += bp; n_log += nl; n_exp += ne; n_pow += np;
    });
  }
  for (auto& x : th) x.join();
  printf("logf checked %llu mismatches %llu\nexpf checked %llu mismatches %llu\npowf checked %llu mismatches %llu\n",
         (unsigned long long)n_log, (unsigned long long)bad_log, (unsigned long long)n_exp, (unsigned long long)bad_exp,
         (unsigned long long)n_pow, (unsigned long long)bad_pow);
  return (bad_log || bad_exp || bad_pow) ? 1 : 0;
}
