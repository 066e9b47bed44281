// sinf / cosf as glibc >= 2.28 computes them (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h - the ARM
// optimized-routines algorithm): argument widened to double, fast range reduction n = round(x * 2/pi) for |x| < 120,
// degree-7 / degree-8 polynomials evaluated in double in a fixed operation order, result rounded to float once.
// Those results are within 0.56 ulp but NOT always the correctly rounded value, so "evaluate in fp64 and round"
// (what sim.cu does by default) disagrees with the reference's libm in about one call in 10^4 - harmless on free
// motion, visible (0.7 mm after 70 steps) once a contact solver amplifies it.
//
// Plain IEEE double arithmetic, no contraction: compile the including unit with -fmad=false (device) or
// -ffp-contract=off (host).  The same source is compiled for the host by tools/trig_check.cpp, which compares it
// with the libm of the machine over 10^8 arguments, and by the oracle variant libsim_oracle_glibcport.so.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GLT_FN __host__ __device__ inline
#else
#define GLT_FN static inline
#endif

namespace glibc_trig {

// x86-64 glibc selects, at load time, a build of the same C source compiled with -mfma -mavx2 on CPUs that have FMA
// (sysdeps/x86_64/fpu/multiarch/s_sinf.c): there every a + b * c below is one fused operation.  GLT_FMA = 1 restates
// that build, GLT_FMA = 0 the generic one; they differ in about one call in 10^7.
#ifndef GLT_FMA
#define GLT_FMA 1
#endif
#if GLT_FMA
#define GLT_MADD(a, b, c) fma((a), (b), (c))   /* a * b + c, one rounding */
#else
#define GLT_MADD(a, b, c) ((a) * (b) + (c))
#endif

GLT_FN uint32_t abstop12(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  return (u >> 20) & 0x7ff;
}

// polynomial coefficients of sincosf_data.c; sgn = -1 selects the second table (used when n & 2)
GLT_FN float poly(double x, double x2, double sgn, int n) {
  const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
               c4 = 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  if ((n & 1) == 0) {
    const double x3 = x * x2;
    const double t1 = GLT_MADD(x2, s3, s2);
    const double x7 = x3 * x2;
    const double s = GLT_MADD(x3, s1, x);
    return (float)GLT_MADD(x7, t1, s);
  }
  const double x4 = x2 * x2;
  const double k2 = GLT_MADD(x2, sgn * c4, sgn * c3);
  const double k1 = GLT_MADD(x2, sgn * c2, sgn * c1);
  const double x6 = x4 * x2;
  const double c = GLT_MADD(x2, k1, sgn * c0);
  return (float)GLT_MADD(x6, k2, c);
}

GLT_FN double reduce_fast(double x, int* np) {
  const double hpi_inv = 0x1.45F306DC9C883p+23;  // 2/pi * 2^24
  const double hpi = 0x1.921FB54442D18p0;
  const double r = x * hpi_inv;
  const int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return GLT_MADD(-(double)n, hpi, x);
}

// returns false when the argument is outside the fast path (|y| >= 120, inf, nan): the caller falls back
GLT_FN bool sinf_fast(float y, float* out) {
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    if (abstop12(y) < abstop12(0x1p-12f)) { *out = y; return true; }
    *out = poly(x, x * x, 1.0, 0);
    return true;
  }
  if (abstop12(y) < abstop12(120.0f)) {
    int n;
    x = reduce_fast(x, &n);
    const double sign[4] = {1.0, -1.0, -1.0, 1.0};
    const double s = sign[n & 3];
    *out = poly(x * s, x * x, (n & 2) ? -1.0 : 1.0, n);
    return true;
  }
  return false;
}
GLT_FN bool cosf_fast(float y, float* out) {
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    if (abstop12(y) < abstop12(0x1p-12f)) { *out = 1.0f; return true; }
    *out = poly(x, x * x, 1.0, 1);
    return true;
  }
  if (abstop12(y) < abstop12(120.0f)) {
    int n;
    x = reduce_fast(x, &n);
    const double sign[4] = {1.0, -1.0, -1.0, 1.0};
    const double s = sign[n & 3];
    *out = poly(x * s, x * x, (n & 2) ? -1.0 : 1.0, n ^ 1);
    return true;
  }
  return false;
}

}  // namespace glibc_trig
