// az_common.h — shared host/device plumbing for the B200 self-play engine.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define AZ_HD __host__ __device__ __forceinline__
#define AZ_D __device__ __forceinline__
// cold paths (move making, root noise, compaction): kept out of line so the per-simulation loop stays small
#define AZ_COLD __host__ __device__ __noinline__
// large rule functions that are called from several places of a kernel: one out-of-line copy (instruction-cache footprint)
#define AZ_HD_CALL __host__ __device__ __noinline__
// experiment builds only ("the view is part of the stack frame"): a kernel's view parameter stays in the
// constant bank even though out-of-line helpers take it by reference — measured slower than the by-value copies, not used
#define AZ_GRID_CONSTANT __grid_constant__
#else
#define AZ_HD inline
#define AZ_D inline
#define AZ_COLD inline
#define AZ_HD_CALL inline
#define AZ_GRID_CONSTANT
#endif

namespace b2az {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;

// Exactly-rounded, never-contracted float ops. The reference is built without FMA (plain
// x86-64 SSE2: SURVEY.md Appendix A "Float order"), so every float formula that feeds a
// selection decision is spelled with these to keep nvcc from fusing a*b+c. The host branch
// exists only so the shared headers can be unit-tested on a CPU (tests/test_shared_headers.py);
// `volatile` stops gcc from contracting if someone builds those tests with -march=native.
AZ_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
AZ_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
AZ_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
#if defined(__CUDACC__) && defined(B2AZ_FDIV_NOINLINE)
// experiment knob: one out-of-line copy of the IEEE division sequence instead of one per call site (code size)
static __device__ __noinline__ float fdiv_out_of_line(float a, float b) { return __fdiv_rn(a, b); }
#endif
AZ_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__) && defined(B2AZ_FDIV_NOINLINE)
  return fdiv_out_of_line(a, b);
#elif defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
AZ_HD float fsqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}
AZ_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
AZ_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
AZ_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  volatile double r = a - b;
  return r;
#endif
}
AZ_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return fma(a, b, c);
#endif
}
AZ_HD u32 f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; u32 u; } x;
  x.f = f;
  return x.u;
#endif
}
AZ_HD float u2f(u32 u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; u32 u; } x;
  x.u = u;
  return x.f;
#endif
}
AZ_HD u64 d2u(double d) {
#if defined(__CUDA_ARCH__)
  return (u64)__double_as_longlong(d);
#else
  union { double d; u64 u; } x;
  x.d = d;
  return x.u;
#endif
}
AZ_HD double u2d(u64 u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  union { double d; u64 u; } x;
  x.u = u;
  return x.d;
#endif
}

}  // namespace b2az
