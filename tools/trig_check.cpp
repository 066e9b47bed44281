// Compares ctrlsim_b200/csrc/glibc_trig.h (compiled here for the host, same source as the device build) with the
// sinf / cosf of this machine's libm.   g++ -O2 -ffp-contract=off tools/trig_check.cpp -o /tmp/trig_check && /tmp/trig_check
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../ctrlsim_b200/csrc/glibc_trig.h"

int main(int argc, char** argv) {
  const long N = argc > 1 ? atol(argv[1]) : 100000000L;
  std::mt19937_64 rng(12345);
  long bad_s = 0, bad_c = 0, bad_s64 = 0, bad_c64 = 0, miss = 0;
  const float ranges[3] = {8.0f, 0.9f, 119.0f};
  for (long i = 0; i < N; ++i) {
    const float r = ranges[i % 3];
    const float x = (float)((double)(rng() >> 11) * (1.0 / 9007199254740992.0) * 2.0 * r - r);
    float s, c;
    if (!glibc_trig::sinf_fast(x, &s) || !glibc_trig::cosf_fast(x, &c)) { ++miss; continue; }
    const float rs = sinf(x), rc = cosf(x);
    bad_s += s != rs;
    bad_c += c != rc;
    bad_s64 += (float)sin((double)x) != rs;
    bad_c64 += (float)cos((double)x) != rc;
    if ((s != rs || c != rc) && bad_s + bad_c <= 5) printf("  x=%a port sin %a cos %a  libm sin %a cos %a\n", x, s, c, rs, rc);
  }
  long bad_t = 0, bad_t64 = 0;
  for (long i = 0; i < N; ++i) {
    const float x = (float)((double)(rng() >> 11) * (1.0 / 9007199254740992.0) * 1.5707 - 0.78535);
    float t;
    if (!glibc_trig::tanf_fast(x, &t)) continue;
    const float rt = tanf(x);
    bad_t += t != rt;
    bad_t64 += (float)tan((double)x) != rt;
    if (t != rt && bad_t <= 5) printf("  x=%a port tan %a libm tan %a\n", x, t, rt);
  }
  printf("tanf on [-pi/4, pi/4]: port differs from libm in %ld of %ld; fp64-and-round differs in %ld\n", bad_t, N, bad_t64);
  if (bad_t) return 1;
  printf("%ld arguments: glibc port differs from libm in %ld sinf / %ld cosf; fp64-and-round differs in %ld / %ld; %ld outside the fast path\n",
         N, bad_s, bad_c, bad_s64, bad_c64, miss);
  return (bad_s || bad_c) ? 1 : 0;
}
