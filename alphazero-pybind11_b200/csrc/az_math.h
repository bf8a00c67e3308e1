// az_math.h — single-precision log/exp/pow that are BIT-IDENTICAL to what the reference gets
// from std::log/std::exp/std::pow(float) on the x86-64 hosts it runs on.
//
// Why: the reference calls these at the root of every search (root-policy temperature
// mcts.cc:114,452; shaped Dirichlet mcts.cc:412-424; Gumbel mcts.cc:217,249; temperature decay
// play_manager.cc:302) and inside libstdc++'s gamma/normal/extreme_value distributions. Visit
// counts are only bit-exact if those values are. glibc >= 2.28 implements logf/expf/powf with the
// ARM "optimized routines" algorithms (double-precision table + polynomial, one final rounding);
// on CPUs with FMA the ifunc picks a variant compiled with -mfma, so the exact result depends on
// WHICH multiply-adds are fused. The sequences below restate the FMA variants of this image's
// glibc 2.39 (decoded from libm.so.6; the non-FMA variant differs in a fraction of a percent of
// inputs). tests/test_shared_headers.py checks them against the live libm: exhaustively for
// logf/expf over the ranges used, and on >10^8 sampled pairs for powf.
//
// Only IEEE double add/mul/fma and integer ops are used, which the B200's FP64 units implement
// exactly, so host (unit test) and device builds agree by construction.
#pragma once

#include "az_common.h"
#include "az_math_tables.h"

namespace b2az {

AZ_COLD float az_logf(float x) {
  const double T[32] = AZ_LOGF_TAB_INIT;
  u32 ix = f2u(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2u == 0u) return -INFINITY;             // log(+-0) = -inf
    if (ix == 0x7f800000u) return x;                 // log(inf) = inf
    if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return u2f(0x7fc00000u) * 1.0f + (x - x);  // nan
    ix = f2u(fmul(x, 8388608.0f));                   // subnormal: normalise
    ix -= 23u << 23;
  }
  const u32 tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const int k = (int)tmp >> 23;
  const u32 iz = ix - (tmp & 0xff800000u);
  const double invc = T[2 * i], logc = T[2 * i + 1];
  const double z = (double)u2f(iz);
  const double y0 = dfma((double)k, AZ_LOGF_LN2, logc);
  const double r = dfma(z, invc, -1.0);
  double y = dfma(AZ_LOGF_A1, r, AZ_LOGF_A2);
  const double r2 = dmul(r, r);
  y = dfma(AZ_LOGF_A0, r2, y);
  y = dfma(y, r2, dadd(r, y0));
  return (float)y;
}

AZ_COLD float az_expf(float x) {
  const unsigned long long T[32] = AZ_EXP2_TAB_INIT;
  const double xd = (double)x;
  const u32 ux = f2u(x);
  const u32 abstop = (ux >> 20) & 0x7ffu;
  if (abstop > 0x42au) {  // |x| >= 88 or nan/inf
    if (ux == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8u) return x + x;
    if (x > 0x1.62e42ep6f) return INFINITY;
    if (x < -0x1.9fe368p6f) return 0.0f;
    if (x < -0x1.9d1d9ep6f) return u2f(1u);  // __math_may_uflowf(0): 0x1.4p-75f squared -> 2^-149
  }
  const double kds = dfma(AZ_EXP_INVLN2_SCALED, xd, AZ_EXP_SHIFT);
  const u64 ki = d2u(kds);
  const double kd = dsub(kds, AZ_EXP_SHIFT);
  const double r = dfma(AZ_EXP_INVLN2_SCALED, xd, -kd);
  const u64 t = T[ki & 31u] + (ki << 47);
  const double s = u2d(t);
  const double z = dfma(AZ_EXP_C0, r, AZ_EXP_C1);
  const double r2 = dmul(r, r);
  double y = dfma(AZ_EXP_C2, r, 1.0);
  y = dfma(z, r2, y);
  y = dmul(y, s);
  return (float)y;
}

// returns 0 if y is not an integer, 1 if odd, 2 if even (glibc e_powf.c checkint)
AZ_HD int az_checkint(u32 iy) {
  const int e = (int)((iy >> 23) & 0xffu);
  if (e < 0x7f) return 0;
  if (e > 0x7f + 23) return 2;
  if (iy & ((1u << (0x7f + 23 - e)) - 1u)) return 0;
  if (iy & (1u << (0x7f + 23 - e))) return 1;
  return 2;
}

AZ_COLD float az_powf(float x, float y) {
  const double TL[32] = AZ_POWLOG2_TAB_INIT;
  const unsigned long long TE[32] = AZ_EXP2_TAB_INIT;
  u64 sign_bias = 0;
  u32 ix = f2u(x);
  const u32 iy = f2u(y);
  const bool y_zin = (2u * iy - 1u) >= (2u * 0x7f800000u - 1u);  // zeroinfnan(iy)
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || y_zin) {
    if (y_zin) {
      if (2u * iy == 0u) return (((ix ^ 0x00400000u) & 0x7fffffffu) > 0x7fc00000u) ? x + y : 1.0f;  // sNaN^0
      if (ix == 0x3f800000u) return (((iy ^ 0x00400000u) & 0x7fffffffu) > 0x7fc00000u) ? x + y : 1.0f;
      if (2u * ix > 2u * 0x7f800000u || 2u * iy > 2u * 0x7f800000u) return x + y;
      if (2u * ix == 2u * 0x3f800000u) return 1.0f;
      if ((2u * ix < 2u * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f;
      return y * y;
    }
    if ((2u * ix - 1u) >= (2u * 0x7f800000u - 1u)) {  // zeroinfnan(ix)
      float x2 = fmul(x, x);
      if ((ix & 0x80000000u) && az_checkint(iy) == 1) x2 = -x2;
      if (2u * ix == 0u && (iy & 0x80000000u)) return (x2 < 0.0f || (f2u(x2) >> 31)) ? -INFINITY : INFINITY;
      return (iy & 0x80000000u) ? fdiv(1.0f, x2) : x2;
    }
    if (ix & 0x80000000u) {
      const int yint = az_checkint(iy);
      if (yint == 0) return u2f(0x7fc00000u);
      if (yint == 1) sign_bias = 1ull << (5 + 11);
      ix &= 0x7fffffffu;
    }
    if (ix < 0x00800000u) {
      ix = f2u(fmul(x, 8388608.0f));
      ix &= 0x7fffffffu;
      ix -= 23u << 23;
    }
  }
  // log2_inline
  const u32 tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const u32 top = tmp & 0xff800000u;
  const u32 iz = ix - top;
  const int k = (int)top >> 23;
  const double invc = TL[2 * i], logc = TL[2 * i + 1];
  const double z = (double)u2f(iz);
  const double r = dfma(z, invc, -1.0);
  const double y0 = dadd((double)k, logc);
  const double yy = dfma(AZ_POWLOG2_A0, r, AZ_POWLOG2_A1);
  const double p = dfma(AZ_POWLOG2_A2, r, AZ_POWLOG2_A3);
  const double r2 = dmul(r, r);
  double q = dfma(AZ_POWLOG2_A4, r, y0);
  const double r4 = dmul(r2, r2);
  q = dfma(r2, p, q);
  const double logx = dfma(yy, r4, q);
  const double ylogx = dmul((double)y, logx);
  if (((d2u(ylogx) >> 47) & 0xffffu) > 0x80beu) {  // |y*log2(x)| >= 126
    if (ylogx > 0x1.fffffffd1d571p+6) return sign_bias ? -INFINITY : INFINITY;
    if (ylogx <= -150.0) return sign_bias ? -0.0f : 0.0f;
    if (ylogx < -149.0) return sign_bias ? -u2f(1u) : u2f(1u);
  }
  // exp2_inline
  const double kds = dadd(ylogx, AZ_EXP2_SHIFT_SCALED);
  const u64 ki = d2u(kds);
  const double kd = dsub(kds, AZ_EXP2_SHIFT_SCALED);
  const double rr = dsub(ylogx, kd);
  const u64 t = TE[ki & 31u] + ((ki + sign_bias) << 47);
  const double s = u2d(t);
  const double zz = dfma(AZ_EXP2_C0_SCALED, rr, AZ_EXP2_C1_SCALED);
  const double rr2 = dmul(rr, rr);
  double e = dfma(AZ_EXP2_C2_SCALED, rr, 1.0);
  e = dfma(zz, rr2, e);
  e = dmul(e, s);
  return (float)e;
}

}  // namespace b2az
