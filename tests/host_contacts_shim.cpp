// Host build of the PRODUCT's contact code (ctrlsim_b200/csrc/sim_contacts.cuh, the file sim.cu includes for the GPU) so that
// tests can run it on the CPU against the oracle: g++ -std=c++17 -O2 -ffp-contract=off -shared -fPIC.  The CUDA qualifiers
// are defined away; the body layout enums and the trig entry points are what sim.cu provides (here: this machine's libm).
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
enum { B_PX, B_PY, B_CX, B_CY, B_LCX, B_LCY, B_ANG, B_VX, B_VY, B_OM, B_SLEEP, B_THR, B_BRK, B_STEER, B_AWAKE, B_PAD, B_FIELDS };
#define B2_PI 3.14159265359f
static inline float cr_sinf(float x) { return sinf(x); }
static inline float cr_cosf(float x) { return cosf(x); }
#include "../ctrlsim_b200/csrc/sim_contacts.cuh"

static int g_scratch[CS_SCRATCH_WORDS];
static CsScratch g_solver;

extern "C" {
int hc_words(int N) { return cs_words(N); }
int hc_body_fields(void) { return B_FIELDS; }
// sim_reset_kernel's contact part
void hc_init(float* body, const float* len, const float* wid, int N, int n, float* cstate) {
  SV sv{body, len, wid, N, n};
  SC c = cs_view(cstate, N, n, g_scratch);
  *c.new_contacts = 1; *c.n_contacts = 0; *c.inv_dt0 = 0.0f;
  for (int k = 0; k < n; ++k) cs_init_body(sv, c, k);
}
// sim_step_kernel's serial part: teleports first (SetTransform happened before the FreeCar steps), then the world step
void hc_world_step(float* body, const float* len, const float* wid, int N, int n, float* cstate, const uint8_t* tele, float dt) {
  SV sv{body, len, wid, N, n};
  SC c = cs_view(cstate, N, n, g_scratch);
  for (int k = 0; k < n; ++k)
    if (tele[k]) { cs_teleport(sv, c, k); *c.new_contacts = 1; }
  world_step(sv, c, dt, &g_solver);
}
int hc_num_touching(float* cstate, int N, int n) {
  SC c = cs_view(cstate, N, n, g_scratch);
  int t = 0;
  for (int k = 0; k < *c.n_contacts; ++k) t += c.ct[k].touching;
  return t;
}
}
