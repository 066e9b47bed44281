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

// tanf for |x| <= pi/4 (the steering range is [-0.7, 0.7]) as glibc computes it: the float-arithmetic fdlibm kernel
// (sysdeps/ieee754/flt-32/k_tanf.c with y = 0, iy = 1; s_tanf.c).  No fused operations: this file of glibc has no
// FMA build.  Returns false outside the range.
GLT_FN bool tanf_fast(float x, float* out) {
  uint32_t ux;
  memcpy(&ux, &x, 4);
  const int32_t hx = (int32_t)ux;
  const int32_t ix = hx & 0x7fffffff;
  if (ix > 0x3f490fda) return false;  // |x| > pi/4: range reduction not restated
  const float T[13] = {3.3333334327e-01f, 1.3333334029e-01f, 5.3968254477e-02f, 2.1869488060e-02f, 8.8632395491e-03f,
                       3.5920790397e-03f, 1.4562094584e-03f, 5.8804126456e-04f, 2.4646313977e-04f, 7.8179444245e-05f,
                       7.1407252108e-05f, -1.8558637748e-05f, 2.5907305826e-05f};
  const float pio4 = 7.8539812565e-01f, pio4lo = 3.7748947079e-08f;
  float y = 0.0f, z, r, v, w, s;
  if (ix < 0x39000000) {  // |x| < 2**-13
    if ((int)x == 0) { *out = x; return true; }
  }
  if (ix >= 0x3f2ca140) {  // |x| >= 0.6744
    if (hx < 0) { x = -x; y = -y; }
    z = pio4 - x;
    w = pio4lo - y;
    x = z + w;
    y = 0.0f;
    if (fabsf(x) < 0x1p-13f) { *out = (float)(1 - ((hx >> 30) & 2)) * 1 * (1.0f - 2 * 1 * x); return true; }
  }
  z = x * x;
  w = z * z;
  r = T[1] + w * (T[3] + w * (T[5] + w * (T[7] + w * (T[9] + w * T[11]))));
  v = z * (T[2] + w * (T[4] + w * (T[6] + w * (T[8] + w * (T[10] + w * T[12])))));
  s = z * x;
  r = y + z * (s * (r + v) + y);
  r += T[0] * s;
  w = x + r;
  if (ix >= 0x3f2ca140) {
    v = 1.0f;
    *out = (float)(1 - ((hx >> 30) & 2)) * (v - 2.0f * (x - (w * w / (w + v) - r)));
    return true;
  }
  *out = w;
  return true;
}

}  // namespace glibc_trig
