// Shared device/host math helpers for the sm_100a Cooper-Frye kernels.
//
// Everything marked IS3D_HD is plain arithmetic that also compiles on the host, so tests/hostcheck can run the
// exact per-cell / per-momentum formulas through g++ as a development sanity check.  The product never takes
// that route: the C ABI (api.cu) only launches the CUDA kernels.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define IS3D_HD __host__ __device__ __forceinline__
#define IS3D_D __device__ __forceinline__
#else
#define IS3D_HD inline
#define IS3D_D inline
#endif

namespace is3d {

// reference src/cpp/iS3D.h:14-17
constexpr double kHbarC = 0.197327053;
constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 2.0 * kPi;
constexpr double kTwoPi2HbarC3 = 2.0 * kPi * kPi * kHbarC * kHbarC * kHbarC;
constexpr double kFourPi2HbarC3 = 4.0 * kPi * kPi * kHbarC * kHbarC * kHbarC;
// (2 pi hbarc)^-3, reference MomentumSpectra.cpp:38
constexpr double kCooperFryePrefactor = 1.0 / (kTwoPi * kHbarC * kTwoPi * kHbarC * kTwoPi * kHbarC);

IS3D_HD double as_double(int64_t i)
{
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(i);
#else
  double d; std::memcpy(&d, &i, 8); return d;
#endif
}
IS3D_HD int64_t as_int64(double d)
{
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(d);
#else
  int64_t i; std::memcpy(&i, &d, 8); return i;
#endif
}

// NaN or +-inf (portable between nvcc and g++)
IS3D_HD bool not_finite(double x) { return !(fabs(x) <= 1.7976931348623157e308); }

// exp(x) for the Bose/Fermi factor, kept entirely in the FP64 FMA pipe: Cody-Waite reduction x = n ln2 + r,
// |r| <= ln2/2, degree-11 polynomial (Chebyshev-node fit, max relative error 1.7e-17 before rounding), exponent
// patched by integer add.  x is clamped to [-700, 700] so the result stays normal: 1/(e^700 + s) ~ 1e-304 stands
// in for the reference's exact 0 of 1/inf, far below any bin's rounding error.
IS3D_HD double fast_exp(double x)
{
  x = fmin(fmax(x, -700.0), 700.0);
  const double kMagic = 6755399441055744.0;   // 1.5 * 2^52: rounds to nearest integer in the low mantissa bits
  double t = fma(x, 1.4426950408889634, kMagic);
  int64_t n = as_int64(t) - as_int64(kMagic);  // integer value of the rounded quotient (small, fits low bits)
  t -= kMagic;
  double r = fma(t, -6.93147180559945286e-01, x);
  r = fma(t, -2.31904681384629956e-17, r);
  double p = 2.5110049204818658e-08;
  p = fma(p, r, 2.763265472252779e-07);
  p = fma(p, r, 2.755724088722987e-06);
  p = fma(p, r, 2.4801485441561313e-05);
  p = fma(p, r, 0.00019841269890076403);
  p = fma(p, r, 0.0013888888952352863);
  p = fma(p, r, 0.008333333333319589);
  p = fma(p, r, 0.04166666666648795);
  p = fma(p, r, 0.1666666666666668);
  p = fma(p, r, 0.5000000000000019);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return as_double(as_int64(p) + (n << 52));
}

// 1/d for d in the normal range (here d = e^x +- 1 >= ~0.1): hardware seed + Newton steps in the FMA pipe,
// no division slow path.
IS3D_HD double fast_rcp(double d)
{
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  return y;
#else
  return 1.0 / d;
#endif
}

}  // namespace is3d
