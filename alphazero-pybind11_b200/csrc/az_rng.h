// az_rng.h — the reference's random number machinery, restated so it can run on the device and
// produce the SAME numbers: pcg32 (vendored in the reference: src/pcg/pcg_random.hpp:1663, 484-531)
// driving libstdc++ 13's std::shuffle / uniform_int (Lemire) / generate_canonical /
// normal (polar) / gamma (Marsaglia-Tsang) / extreme_value algorithms
// (/usr/include/c++/13/bits/{stl_algo.h:3742-3805, uniform_int_dist.h:257-331, random.tcc:1811-1844,
// 2340-2392, 2582-2591, 3349-3381}).
//
// Where the reference draws (mcts.cc): add_children shuffle :100, add_root_noise gamma :430-439,
// init_gumbel_state extreme_value :205-209, pick_move uniform_real :718-719.
//
// Float subtleties that are reproduced on purpose (they change low bits):
//   * libstdc++ mixes double literals into float expressions (`- 1.0`, `0.0331 * n*n*n*n`,
//     `0.5 * n * n`), so parts of the gamma acceptance test run in double.
//   * normal_distribution caches its second variate inside the distribution object; the reference
//     constructs one gamma_distribution per add_root_noise call (plain) or per child (shaped).
#pragma once

#include "az_common.h"
#include "az_math.h"

namespace b2az {

struct Pcg32 {
  u64 state;
  u64 inc;
};

#define AZ_PCG_MULT 6364136223846793005ULL
#define AZ_PCG_DEFAULT_INC 1442695040888963407ULL

// engine(seed): state = bump(seed + inc)  (pcg_random.hpp:484-490); default stream
AZ_HD void pcg32_seed(Pcg32& r, u64 seed) {
  r.inc = AZ_PCG_DEFAULT_INC;
  r.state = (seed + r.inc) * AZ_PCG_MULT + r.inc;
}
// engine(seed, stream): inc = (stream << 1) | 1  (pcg_random.hpp:494-501, specific_stream)
AZ_HD void pcg32_seed_stream(Pcg32& r, u64 seed, u64 stream) {
  r.inc = (stream << 1) | 1ULL;
  r.state = (seed + r.inc) * AZ_PCG_MULT + r.inc;
}
// XSH-RR output of the PREVIOUS state (output_previous = true for 64-bit state)
AZ_HD u32 pcg32_next(Pcg32& r) {
  const u64 old = r.state;
  r.state = old * AZ_PCG_MULT + r.inc;
  const u32 xorshifted = (u32)(((old >> 18) ^ old) >> 27);
  const u32 rot = (u32)(old >> 59);
  return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
}

// uniform_int_distribution<unsigned long>{0, range-1} with a 32-bit URBG: Lemire's method on
// 32x32->64 products (uniform_int_dist.h:257-281 via :324-331). range in [1, 2^32).
AZ_HD u32 rng_below(Pcg32& r, u32 range) {
  u64 product = (u64)pcg32_next(r) * (u64)range;
  u32 low = (u32)product;
  if (low < range) {
    const u32 threshold = (0u - range) % range;
    while (low < threshold) {
      product = (u64)pcg32_next(r) * (u64)range;
      low = (u32)product;
    }
  }
  return (u32)(product >> 32);
}

// std::shuffle(first, last, pcg32) for n <= 65535 (pairwise path, stl_algo.h:3766-3799).
// x / d and x % d for x < 2^22, d < 2^11 (the pairwise shuffle's x < (i+1)(i+2)): float quotient estimate + one
// correction step instead of the 32-bit integer division sequence; exact (float(x) and float(d) are exact, the
// estimate is off by at most one).
AZ_HD void small_divmod(u32 x, u32 d, u32& q, u32& rem) {
#if defined(__CUDA_ARCH__)
  u32 qq = (u32)__fdividef((float)x, (float)d);
  int rr = (int)(x - qq * d);
  if (rr < 0) { --qq; rr += (int)d; }
  else if (rr >= (int)d) { ++qq; rr -= (int)d; }
  q = qq; rem = (u32)rr;
#else
  q = x / d; rem = x % d;
#endif
}
template <typename T>
AZ_HD void rng_shuffle(Pcg32& r, T* a, u32 n) {
  if (n == 0) return;
  u32 i = 1;
  if ((n % 2u) == 0u) {
    const u32 j = rng_below(r, 2u);
    T t = a[i]; a[i] = a[j]; a[j] = t;
    ++i;
  }
  while (i < n) {
    const u32 swap_range = i + 1u;
    const u32 x = rng_below(r, swap_range * (swap_range + 1u));
    u32 p0, p1;
    if (n <= 1024u) small_divmod(x, swap_range + 1u, p0, p1);
    else { p0 = x / (swap_range + 1u); p1 = x % (swap_range + 1u); }
    T t = a[i]; a[i] = a[p0]; a[p0] = t;
    ++i;
    t = a[i]; a[i] = a[p1]; a[p1] = t;
    ++i;
  }
}
// std::shuffle over a list of <= 8 small values packed as 4-bit nibbles in one register (element i =
// bits 4i..4i+3): the same draw sequence and the same swaps as rng_shuffle, without an addressable
// array (an indexed local array would live in local memory on the device).
AZ_HD void nib_swap(u32& a, u32 i, u32 j) {
  const u32 x = ((a >> (4u * i)) ^ (a >> (4u * j))) & 15u;
  a ^= (x << (4u * i)) | (x << (4u * j));
}
AZ_HD void rng_shuffle_nib(Pcg32& r, u32& a, u32 n) {
  if (n == 0) return;
  u32 i = 1;
  if ((n % 2u) == 0u) {
    const u32 j = rng_below(r, 2u);
    nib_swap(a, i, j);
    ++i;
  }
  while (i < n) {
    const u32 swap_range = i + 1u;
    const u32 x = rng_below(r, swap_range * (swap_range + 1u));
    u32 p0, p1;
    small_divmod(x, swap_range + 1u, p0, p1);
    nib_swap(a, i, p0);
    ++i;
    nib_swap(a, i, p1);
    ++i;
  }
}
// Same draw sequence as rng_shuffle, result discarded (update_root on a never-expanded root:
// mcts.cc:154-156 shuffles children that are thrown away on the next line).
AZ_HD void rng_shuffle_discard(Pcg32& r, u32 n) {
  if (n == 0) return;
  u32 i = 1;
  if ((n % 2u) == 0u) { (void)rng_below(r, 2u); ++i; }
  while (i < n) {
    const u32 swap_range = i + 1u;
    (void)rng_below(r, swap_range * (swap_range + 1u));
    i += 2;
  }
}

// generate_canonical<float, 24>(pcg32): one 32-bit draw, float(x) / 2^32, clamped below 1.
AZ_HD float rng_canonical(Pcg32& r) {
  const float sum = (float)pcg32_next(r);       // u32 -> float, round to nearest
  float ret = fdiv(sum, 4294967296.0f);
  if (ret >= 1.0f) ret = u2f(0x3f7fffffu);      // nextafter(1, 0)
  return ret;
}
// uniform_real_distribution<float>{0, 1}: canonical * (b - a) + a
AZ_HD float rng_uniform01(Pcg32& r) { return fadd(fmul(rng_canonical(r), 1.0f), 0.0f); }

struct NormalState {
  float saved;
  int available;
};
// normal_distribution<float>{0, 1} (polar Box-Muller; random.tcc:1811-1844)
AZ_HD float rng_normal(Pcg32& r, NormalState& st) {
  float ret;
  if (st.available) {
    st.available = 0;
    ret = st.saved;
  } else {
    float x, y, r2;
    do {
      x = fsub(fmul(2.0f, rng_canonical(r)), 1.0f);  // (double)(2u) - 1.0 rounds to the same float
      y = fsub(fmul(2.0f, rng_canonical(r)), 1.0f);
      r2 = fadd(fmul(x, x), fmul(y, y));
    } while (r2 > 1.0f || r2 == 0.0f);
    const float mult = fsqrt(fdiv(fmul(-2.0f, az_logf(r2)), r2));
    st.saved = fmul(x, mult);
    st.available = 1;
    ret = fmul(y, mult);
  }
  return fadd(fmul(ret, 1.0f), 0.0f);
}

struct GammaDist {  // gamma_distribution<float>{alpha, 1.0f} with its embedded normal_distribution
  float alpha, malpha, a2;
  NormalState nd;
};
AZ_HD void gamma_init(GammaDist& g, float alpha) {  // param_type::_M_initialize (random.tcc:2335-2343)
  g.alpha = alpha;
  g.malpha = (alpha < 1.0f) ? fadd(alpha, 1.0f) : alpha;
  const float a1 = fsub(g.malpha, fdiv(1.0f, 3.0f));
  g.a2 = fdiv(1.0f, fsqrt(fmul(9.0f, a1)));
  g.nd.saved = 0.0f;
  g.nd.available = 0;
}
AZ_HD float gamma_draw(Pcg32& r, GammaDist& g) {  // random.tcc:2351-2392
  float u, v, n;
  const float a1 = fsub(g.malpha, fdiv(1.0f, 3.0f));
  for (;;) {
    do {
      n = rng_normal(r, g.nd);
      v = fadd(1.0f, fmul(g.a2, n));
    } while (v <= 0.0f);
    v = fmul(fmul(v, v), v);
    u = rng_canonical(r);
    // u > 1.0f - 0.0331 * n*n*n*n   (double)
    const double dn = (double)n;
    const double lim = dsub(1.0, dmul(dmul(dmul(dmul(0.0331, dn), dn), dn), dn));
    if (!((double)u > lim)) break;
    // log(u) > 0.5*n*n + a1*(1.0 - v + log(v))   (double, logs are float)
    const double rhs = dadd(dmul(dmul(0.5, dn), dn),
                            dmul((double)a1, dadd(dsub(1.0, (double)v), (double)az_logf(v))));
    if (!((double)az_logf(u) > rhs)) break;
  }
  if (g.alpha == g.malpha) return fmul(fmul(a1, v), 1.0f);
  do {
    u = rng_canonical(r);
  } while (u == 0.0f);
  return fmul(fmul(fmul(az_powf(u, fdiv(1.0f, g.alpha)), a1), v), 1.0f);
}

// extreme_value_distribution<float>{0, 1}: a - b * log(-log(1 - u))
AZ_HD float rng_gumbel(Pcg32& r) {
  const float u = rng_canonical(r);
  return fsub(0.0f, fmul(1.0f, az_logf(-az_logf(fsub(1.0f, u)))));
}

}  // namespace b2az
