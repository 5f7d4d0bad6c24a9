// Shared device/host math helpers for the sm_100a Cooper-Frye kernels.
//
// Everything marked IS3D_HD is plain arithmetic that also compiles on the host: the host layer reuses the table /
// spline helpers (dftables.cuh, gauss_thermal.cuh) for its start-up tables.  The compute entry points never take that
// route: the C ABI (api.cu) only launches the CUDA kernels.
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

// exp(x) for the Bose/Fermi factor, kept in the FP64 FMA pipe and short enough to leave the pipe to the physics:
//   x = k (ln2/1024) + r,  |r| <= ln2/2048,   e^x = 2^(k >> 10) * T[k & 1023] * (1 + r + r^2/2 + r^3/6)
// T[m] = 2^(m/1024) is a 1024-entry table (8 KB; shared memory on the device, filled by load_exp_table).
// 7 FP64 instructions: 3 reduction + 3 polynomial + 1 scaling (the first version, a degree-11 Horner form with FP64
// clamps, needed 19).  Errors: truncation r^4/24 < 5.5e-16; the reduction uses ln2/1024 rounded to double in ONE fma, so
// r carries |x| 1.1e-16 (3e-15 of e^x at x = 30, where the Bose/Fermi factor is already 1e-13; the continuous paths
// promise 1e-10).  The exponent is patched by an integer multiply-add on the high word.  Range: the high word of x is
// clamped (one integer min) so that x never exceeds 680: beyond that e^x is "huge" (>= 2e295) and 1/(e^x + s) is the
// reference's 0 to more than 200 decades; the margin to the double range lets callers multiply e^x by E/T before
// taking one reciprocal.  x <= -708 cannot occur (x = (E - b mu_B)/T >= -mu_B/T > -10); NaN input behaves like 680.
// tools/gen_exp_table.py derives the constants and scans the error of the whole construction in 60-digit arithmetic.
constexpr int kExpTableBits = 10;
constexpr int kExpTableSize = 1 << kExpTableBits;

IS3D_HD int hi_word(double d) { return (int)(as_int64(d) >> 32); }
// d with (k << 20) added to its high word, i.e. d * 2^k
IS3D_HD double scale_by_pow2(double d, int k)
{
#if defined(__CUDA_ARCH__)
  return __hiloint2double(k * 0x100000 + __double2hiint(d), __double2loint(d));     // one IMAD
#else
  return as_double(as_int64(d) + ((int64_t)k << 52));
#endif
}
// x with its high word limited to that of 680.0 (signed compare: negative x is never touched)
IS3D_HD double clamp_hi_word_680(double x)
{
#if defined(__CUDA_ARCH__)
  return __hiloint2double(min(__double2hiint(x), 0x40854000), __double2loint(x));
#else
  const int hx = hi_word(x);
  return hx > 0x40854000 ? as_double(((int64_t)0x40854000 << 32) | (as_int64(x) & 0xffffffffll)) : x;
#endif
}

// constants with a non-zero low word live in the constant bank on the device: DFMA reads a c[bank][offset] operand
// directly, whereas an immediate costs two IMAD.MOV per use inside the register-starved inner loops
#if defined(__CUDACC__)
static __constant__ double c_exp_consts[3] = {1477.3197218702985, -6.769015435155716e-04, 1.6666666666666666e-01};
#endif

// CLAMP = false: the caller has already passed x through clamp_hi_word_680 (and keeps using the clamped value), or knows
// x <= 680 some other way (the spectra kernels' per-tile range flag)
//
// The table is stored PRE-COMPENSATED for the exponent patch: entry m holds the bits of 2^(m/1024) with (m << 10) subtracted
// from the high word, so that ONE integer multiply-add  hi = k * 1024 + hi'  restores the mantissa and adds (k >> 10) to the
// exponent field in the same instruction (k * 1024 = (k >> 10) << 20 + (k & 1023) << 10; the arithmetic is modulo 2^32).
// The first version patched the exponent of the RESULT with shift + multiply-add.
template <bool CLAMP = true>
IS3D_HD double fast_exp_k(double x, const double *__restrict__ tab, int &spare)
{
#if defined(__CUDA_ARCH__)
  const double kInv = c_exp_consts[0], kStep = c_exp_consts[1], kSixth = c_exp_consts[2];
#else
  const double kInv = 1477.3197218702985, kStep = -6.769015435155716e-04, kSixth = 1.6666666666666666e-01;
#endif
  const double kMagic = 6755399441055744.0;          // 1.5 * 2^52: rounds to nearest integer in the low mantissa bits
  if (CLAMP) x = clamp_hi_word_680(x);
  double t = fma(x, kInv, kMagic);                   // 1024 / ln2
  const int k = (int)as_int64(t);                    // low word of t = the integer (two's complement), |k| < 2^21 here
  t -= kMagic;
  const double r = fma(t, kStep, x);                 // x - k ln2/1024
  double q = fma(r, kSixth, 0.5);
  q = fma(q, r, 1.0);
  q = q * r;                                         // r + r^2/2 + r^3/6
#if defined(__CUDA_ARCH__) && !defined(IS3D_EXP_TABLE_GENERIC)
  // the table lives in shared memory (load_exp_table): mask, then ONE multiply-add forms the 32-bit shared address
  // (nvcc's own sequence for tab[k & 1023] is shift + mask + add)
  double Tc;
  {
    const unsigned base = (unsigned)__cvta_generic_to_shared(tab);
    unsigned addr;
    asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(addr) : "r"((unsigned)k & (unsigned)(kExpTableSize - 1)), "r"(base));
    asm("ld.shared.f64 %0, [%1];" : "=d"(Tc) : "r"(addr));
    spare = (int)addr;
  }
#else
  const double Tc = tab[k & (kExpTableSize - 1)];
  spare = 0;
#endif
#if defined(__CUDA_ARCH__)
  const double T = __hiloint2double(k * (1 << kExpTableBits) + __double2hiint(Tc), __double2loint(Tc));     // one IMAD
#else
  const double T = as_double(as_int64(Tc) + (int64_t)((uint64_t)(int64_t)k << (32 + kExpTableBits)));
#endif
  return fma(T, q, T);
}
template <bool CLAMP = true>
IS3D_HD double fast_exp(double x, const double *__restrict__ tab)
{
  int spare;
  return fast_exp_k<CLAMP>(x, tab, spare);
}

// host-side construction of the (pre-compensated, see fast_exp) table, uploaded once per context
inline void fill_exp_table(double *tab)
{
  for (int m = 0; m < kExpTableSize; m++) {
    const double v = (double)exp2l((long double)m / (long double)kExpTableSize);
    tab[m] = as_double(as_int64(v) - ((int64_t)m << (32 + kExpTableBits)));
  }
}

#if defined(__CUDACC__)
// copies the table from global memory into the block's shared-memory copy; the caller synchronises
__device__ __forceinline__ void load_exp_table(double *smem_tab, const double *__restrict__ gmem_tab)
{
  for (int m = threadIdx.x; m < kExpTableSize; m += blockDim.x) smem_tab[m] = gmem_tab[m];
}
#endif

// 1/d for d in the normal range (here d = e^x +- 1 >= ~0.1, or d = E/T): hardware seed + one cubically convergent
// step in the FMA pipe, no division slow path.  The seed (MUFU.RCP64H) sees only the upper 32 bits of d, i.e. it
// carries ~20 bits: e = 1 - d y0 <= 1.5 * 2^-19, y = y0 (1 + e + e^2) leaves e^3 <= 2^-55.
// lo_any: any 32-bit value the caller has at hand (fast_exp_k's spare word): only the high word of the seed carries
// information, and taking its low word from a register that is already live saves the register clear (one IMAD.MOV per
// reciprocal in the inner loops) that the architectural "low word = 0" costs; the seed moves by < 2^-20.
IS3D_HD double fast_rcp(double d, int lo_any)
{
#if defined(__CUDA_ARCH__)
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  double y = __hiloint2double(__double2hiint(y0), lo_any);
  double e = fma(-d, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  return y;
#else
  (void)lo_any;
  return 1.0 / d;
#endif
}
IS3D_HD double fast_rcp(double d) { return fast_rcp(d, 0); }

}  // namespace is3d
