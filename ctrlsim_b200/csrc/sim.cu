// GPU-resident scenario batch: vehicle dynamics, collision checks, rewards, focal grouping and metrics.
// THIS FILE IS COMPILED WITH -fmad=false: the reference simulator is built without FMA contraction
// (nocturne/CMakeLists.txt:3-9, -std=c++17) and positions live at Waymo-scale fp32 coordinates, so every multiply
// and add is rounded separately, in the reference's operation order.
//
// Restated (reference file:line):
//   S1 FreeCar::Throttle/Brake/Turn/Step        nocturne/cpp/src/physics/FreeCar.cpp:66-86,88-186 ; defines.h:4-11
//   S2 b2Island::Solve, contact-free subset     third_party/box2d/src/dynamics/b2_island.cpp:188-388
//      b2Body setters / sleep, mass centre      include/box2d/b2_body.h:501-530,637-658,859-863 ; b2_body.cpp:290-354,420-445
//      b2PolygonShape::ComputeMass              src/collision/b2_polygon_shape.cpp:357-431 ; constants b2_common.h:41,95-119
//   S3 Vehicle::Step / CreatePhysicsBody        nocturne/cpp/src/vehicle.cc:25-66,75-88,137-179 ; object.cc:14-28
//   S4 Scenario::UpdateCollision                nocturne/cpp/src/scenario.cc:294-328 ; geometry/polygon.cc:19-27,84-98 ;
//      geometry/intersection.cc:200-233 ; include/geometry/aabb.h:47-50
//   S5 compute_reward / update_vehicle_data_dict utils/sim.py:83-141 ; evaluators/policy_evaluator.py:99-159 ;
//      nearest-vehicle distance                 datasets/rl_waymo/dataset.py:202-236 ; evaluators/evaluator.py:87-104
//   T1 Policy.update_state                      policies/policy.py:68-105
//   T2 greedy focal grouping                    policies/autoregressive_policy.py:86-138 ; dataset.py:278-319
//   S6 act / apply_gt_action / inverse bicycle  policies/autoregressive_policy.py:256-274 ; evaluators/evaluator.py:160-193 ;
//      nocturne/bicycle_model.py:51-109
//   S7 update_running_statistics / histograms   evaluators/policy_evaluator.py:162-305
// Box2D's contact response for vehicle-vehicle contacts: sim_contacts.cuh (one thread per scene between the parallel
// FreeCar step and the parallel Vehicle::Step / collision flags); CTRLSIM_CONTACTS=0 falls back to the contact-free subset.
//
// libm: the reference calls glibc sinf/cosf/tanf (results are the correctly rounded fp32 value in all but ~1e-9 of
// cases).  CUDA's fp32 versions are 1-2 ulp, so those calls are evaluated in fp64 and rounded once.
#include "common.cuh"
#include "glibc_trig.h"
#include "kernels.h"
#include "model.h"

namespace ctrlsim {

// 0: sinf / cosf evaluated in fp64 and rounded once (correctly rounded; differs from glibc's in ~1 % of calls by 1 ulp)
// 1: glibc's own algorithm (glibc_trig.h; identical to the x86-64 FMA build of glibc in 1.2e8 of 1.2e8 arguments)
__constant__ int g_trig_glibc = 1;  // default since round 2: bit-exact with the reference through contacts
__device__ __forceinline__ float cr_sinf(float x) {
  float r;
  if (g_trig_glibc && glibc_trig::sinf_fast(x, &r)) return r;
  return (float)sin((double)x);
}
__device__ __forceinline__ float cr_cosf(float x) {
  float r;
  if (g_trig_glibc && glibc_trig::cosf_fast(x, &r)) return r;
  return (float)cos((double)x);
}
__device__ __forceinline__ float cr_tanf(float x) {
  float r;
  if (g_trig_glibc && glibc_trig::tanf_fast(x, &r)) return r;
  return (float)tan((double)x);
}


#define B2_PI 3.14159265359f
enum { B_PX, B_PY, B_CX, B_CY, B_LCX, B_LCY, B_ANG, B_VX, B_VY, B_OM, B_SLEEP, B_THR, B_BRK, B_STEER, B_AWAKE, B_PAD, B_FIELDS };
enum { O_X, O_Y, O_HEAD, O_SPEED, O_FIELDS };

struct Body {
  float px, py, cx, cy, lcx, lcy, ang, vx, vy, om, sleep_t, thr, brk, steer, awake;
};

__device__ __forceinline__ void body_load(const float* base, int N, int i, Body& B) {
  B.px = base[B_PX * N + i]; B.py = base[B_PY * N + i]; B.cx = base[B_CX * N + i]; B.cy = base[B_CY * N + i];
  B.lcx = base[B_LCX * N + i]; B.lcy = base[B_LCY * N + i]; B.ang = base[B_ANG * N + i]; B.vx = base[B_VX * N + i];
  B.vy = base[B_VY * N + i]; B.om = base[B_OM * N + i]; B.sleep_t = base[B_SLEEP * N + i]; B.thr = base[B_THR * N + i];
  B.brk = base[B_BRK * N + i]; B.steer = base[B_STEER * N + i]; B.awake = base[B_AWAKE * N + i];
}
__device__ __forceinline__ void body_store(float* base, int N, int i, const Body& B) {
  base[B_PX * N + i] = B.px; base[B_PY * N + i] = B.py; base[B_CX * N + i] = B.cx; base[B_CY * N + i] = B.cy;
  base[B_LCX * N + i] = B.lcx; base[B_LCY * N + i] = B.lcy; base[B_ANG * N + i] = B.ang; base[B_VX * N + i] = B.vx;
  base[B_VY * N + i] = B.vy; base[B_OM * N + i] = B.om; base[B_SLEEP * N + i] = B.sleep_t; base[B_THR * N + i] = B.thr;
  base[B_BRK * N + i] = B.brk; base[B_STEER * N + i] = B.steer; base[B_AWAKE * N + i] = B.awake;
}

__device__ __forceinline__ void set_awake(Body& B, bool flag) {
  B.awake = flag ? 1.f : 0.f;
  B.sleep_t = 0.f;
  if (!flag) { B.vx = B.vy = 0.f; B.om = 0.f; }
}
__device__ __forceinline__ void set_linvel(Body& B, float vx, float vy) {
  if (vx * vx + vy * vy > 0.0f) set_awake(B, true);
  B.vx = vx; B.vy = vy;
}
__device__ __forceinline__ void set_angvel(Body& B, float w) {
  if (w * w > 0.0f) set_awake(B, true);
  B.om = w;
}
__device__ __forceinline__ void set_transform(Body& B, float x, float y, float angle) {
  const float qs = cr_sinf(angle), qc = cr_cosf(angle);
  B.px = x; B.py = y;
  B.cx = (qc * B.lcx - qs * B.lcy) + x;
  B.cy = (qs * B.lcx + qc * B.lcy) + y;
  B.ang = angle;
}
__device__ void box_local_center(float hx, float hy, float& lcx, float& lcy) {
  const float vx[4] = {-hx, hx, hx, -hx}, vy[4] = {-hy, -hy, hy, hy};
  float cx = 0.0f, cy = 0.0f, area = 0.0f;
  const float sx = vx[0], sy = vy[0];
  const float k_inv3 = 1.0f / 3.0f;
  for (int i = 0; i < 4; ++i) {
    const float e1x = vx[i] - sx, e1y = vy[i] - sy;
    const int j = (i + 1 < 4) ? i + 1 : 0;
    const float e2x = vx[j] - sx, e2y = vy[j] - sy;
    const float D = e1x * e2y - e1y * e2x;
    const float tri = 0.5f * D;
    area += tri;
    const float k = tri * k_inv3;
    cx += k * (e1x + e2x);
    cy += k * (e1y + e2y);
  }
  const float mass = 20.f * area;
  const float inv_area = 1.0f / area;
  cx *= inv_area; cy *= inv_area;
  const float mcx = cx + sx, mcy = cy + sy;
  const float lx = mass * mcx, ly = mass * mcy;
  const float inv_mass = 1.0f / mass;
  lcx = lx * inv_mass; lcy = ly * inv_mass;
}

#include "sim_contacts.cuh"

__device__ __forceinline__ float dampen(float speed, float target, float damping, float dt) {
  const float red = damping * dt;
  if (speed - target > red) return speed - red;
  if (speed - target < -red) return speed + red;
  return target;
}

__device__ void freecar_step(Body& B, float length, float dt) {
  float target, acc;
  const float thr = B.thr, brk = B.brk, st = B.steer;
  if (thr > 0.0f) {
    if (thr > brk) { target = 50.0f; acc = thr - brk; } else { target = 0.0f; acc = brk - thr; }
  } else {
    if (thr < -brk) { target = -5.0f; acc = -thr - brk; } else { target = 0.0f; acc = brk + thr; }
  }
  float w = B.om;
  const float beta = (float)atan(0.5 * (double)cr_tanf(st));
  const float c = cr_cosf(B.ang + beta), sn = cr_sinf(B.ang + beta);
  const float fx = -sn, fy = c, rx = c, ry = sn;
  float vf = B.vx * fx + B.vy * fy;
  float vr = B.vx * rx + B.vy * ry;
  const float dv = acc * dt;
  if (vf < target) vf = fminf(vf + dv, target); else vf = fmaxf(vf - dv, target);
  float w_steer = 0.0f;
  if (fabs((double)st) > 0.0000001) {
    const float ray = 1.f / cr_tanf(st) * length / cr_cosf(beta);
    w_steer = vf / ray;
  }
  vr = dampen(vr, 0.0f, 25.f, dt);
  w = dampen(w, w_steer, 10.f, dt);
  const float sx = rx * vr + fx * vf, sy = ry * vr + fy * vf;
  set_linvel(B, sx, sy);
  set_angvel(B, w);
}

__device__ void island_solve(Body& B, float h) {
  if (B.awake == 0.f) return;
  float vx = B.vx, vy = B.vy, w = B.om;
  const float tx = h * vx, ty = h * vy;
  if (tx * tx + ty * ty > 5.0f * 5.0f) {
    const float ratio = 5.0f / sqrtf(tx * tx + ty * ty);
    vx *= ratio; vy *= ratio;
  }
  const float rot = h * w;
  const float max_rot = 0.5f * B2_PI;
  if (rot * rot > max_rot * max_rot) {
    const float ratio = max_rot / fabsf(rot);
    w *= ratio;
  }
  B.cx += h * vx; B.cy += h * vy; B.ang += h * w;
  const float qs = cr_sinf(B.ang), qc = cr_cosf(B.ang);
  B.px = B.cx - (qc * B.lcx - qs * B.lcy);
  B.py = B.cy - (qs * B.lcx + qc * B.lcy);
  B.vx = vx; B.vy = vy; B.om = w;
  const float ang_tol = 2.0f / 180.0f * B2_PI, lin_tol = 0.01f;
  if (w * w > ang_tol * ang_tol || vx * vx + vy * vy > lin_tol * lin_tol) {
    B.sleep_t = 0.0f;
  } else {
    B.sleep_t += h;
    if (B.sleep_t >= 0.5f) set_awake(B, false);
  }
}

// ---- geometry ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cross2(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }

__device__ bool separates(float e0x, float e0y, float e1x, float e1y, int n, const float* X, const float* Y) {
  const float dx = e1x - e0x, dy = e1y - e0y;
  for (int k = 0; k < n; ++k)
    if (cross2(X[k] - e0x, Y[k] - e0y, dx, dy) <= 0.0f) return false;
  return true;
}
__device__ bool poly_intersects(int n1, const float* X1, const float* Y1, int n2, const float* X2, const float* Y2) {
  for (int k = 0; k < n1; ++k) {
    const int k1 = (k == n1 - 1) ? 0 : k + 1;
    if (separates(X1[k], Y1[k], X1[k1], Y1[k1], n2, X2, Y2)) return false;
  }
  for (int k = 0; k < n2; ++k) {
    const int k1 = (k == n2 - 1) ? 0 : k + 1;
    if (separates(X2[k], Y2[k], X2[k1], Y2[k1], n1, X1, Y1)) return false;
  }
  return true;
}
__device__ bool poly_contains(int n, const float* X, const float* Y, float px, float py) {
  for (int i = 1; i < n; ++i)
    if (cross2(px - X[i - 1], py - Y[i - 1], X[i] - X[i - 1], Y[i] - Y[i - 1]) > 0.0f) return false;
  return cross2(px - X[n - 1], py - Y[n - 1], X[0] - X[n - 1], Y[0] - Y[n - 1]) <= 0.0f;
}
__device__ bool poly_segment_intersects(int n, const float* X, const float* Y, float ax, float ay, float bx, float by) {
  if (ax == bx && ay == by) return poly_contains(n, X, Y, ax, ay);
  const float dx = bx - ax, dy = by - ay;
  float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
  for (int k = 0; k < n; ++k) {
    const float cur = cross2(X[k] - ax, Y[k] - ay, dx, dy);
    mn = fminf(mn, cur); mx = fmaxf(mx, cur);
  }
  if (mx < 0.0f || mn > 0.0f) return false;
  for (int k = 0; k < n; ++k) {
    const int k1 = (k == n - 1) ? 0 : k + 1;
    const float ex = X[k1] - X[k], ey = Y[k1] - Y[k];
    const float v0 = cross2(ax - X[k], ay - Y[k], ex, ey);
    const float v1 = cross2(bx - X[k], by - Y[k], ex, ey);
    if (v0 > 0.0f && v1 > 0.0f) return false;
  }
  return true;
}

__global__ void geom_poly_poly_kernel(const float* xy1, int n1, const float* xy2, int n2, int* out) {
  float X1[8], Y1[8], X2[8], Y2[8];
  for (int i = 0; i < n1; ++i) { X1[i] = xy1[2 * i]; Y1[i] = xy1[2 * i + 1]; }
  for (int i = 0; i < n2; ++i) { X2[i] = xy2[2 * i]; Y2[i] = xy2[2 * i + 1]; }
  out[0] = poly_intersects(n1, X1, Y1, n2, X2, Y2);
}
__global__ void geom_poly_seg_kernel(const float* xy, int n, const float* seg, int* out) {
  float X[8], Y[8];
  for (int i = 0; i < n; ++i) { X[i] = xy[2 * i]; Y[i] = xy[2 * i + 1]; }
  out[0] = poly_segment_intersects(n, X, Y, seg[0], seg[1], seg[2], seg[3]);
}
int launch_geom_poly_poly(const float* xy1, int n1, const float* xy2, int n2, int* out, cudaStream_t st) {
  if (n1 > 8 || n2 > 8) return set_error(-2, "geom: at most 8 vertices");
  geom_poly_poly_kernel<<<1, 1, 0, st>>>(xy1, n1, xy2, n2, out);
  CS_CHECK_LAUNCH("geom_poly_poly");
  return 0;
}
int launch_geom_poly_seg(const float* xy, int n, const float* seg, int* out, cudaStream_t st) {
  if (n > 8) return set_error(-2, "geom: at most 8 vertices");
  geom_poly_seg_kernel<<<1, 1, 0, st>>>(xy, n, seg, out);
  CS_CHECK_LAUNCH("geom_poly_seg");
  return 0;
}

constexpr int SIM_THREADS = 64;  // == CTRLSIM_MAX_VEH
constexpr int SEG_TILE = 256;

struct Obb { float X[4], Y[4], bb[4]; };

__device__ void make_obb(float ox, float oy, float heading, float len, float wid, Obb& o) {
  const float sh = cr_sinf(heading), ch = cr_cosf(heading);
  const float hl = len * 0.5f, hw = wid * 0.5f;
  const float lx[4] = {hl, -hl, -hl, hl}, ly[4] = {hw, hw, -hw, -hw};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o.X[k] = (lx[k] * ch - ly[k] * sh) + ox;
    o.Y[k] = (lx[k] * sh + ly[k] * ch) + oy;
  }
  o.bb[0] = o.bb[2] = o.X[0]; o.bb[1] = o.bb[3] = o.Y[0];
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    o.bb[0] = fminf(o.bb[0], o.X[k]); o.bb[2] = fmaxf(o.bb[2], o.X[k]);
    o.bb[1] = fminf(o.bb[1], o.Y[k]); o.bb[3] = fmaxf(o.bb[3], o.Y[k]);
  }
}
__device__ __forceinline__ bool aabb_hit(const float* a, const float* b) {
  return a[0] < b[2] && a[2] > b[0] && a[1] < b[3] && a[3] > b[1];
}

// Block = one scene, thread = one vehicle. Collision flags of the current object poses.
__device__ void update_collision(const CtrlSimBatch& b, int s, int i, int n, bool present, float ox, float oy,
                                 float heading, float len, float wid, Obb* sh_obb, float4* sh_seg) {
  Obb me;
  if (present) { make_obb(ox, oy, heading, len, wid, me); sh_obb[i] = me; }
  __syncthreads();
  bool cv = false, ce = false;
  if (present) {
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      if (!aabb_hit(me.bb, sh_obb[j].bb)) continue;
      if (poly_intersects(4, me.X, me.Y, 4, sh_obb[j].X, sh_obb[j].Y)) cv = true;
    }
  }
  const int nseg = b.n_seg[s];
  const float4* segs = reinterpret_cast<const float4*>(b.segs + (size_t)s * b.max_seg * 4);
  for (int k0 = 0; k0 < nseg; k0 += SEG_TILE) {
    const int nk = min(SEG_TILE, nseg - k0);
    __syncthreads();
    for (int k = threadIdx.x; k < nk; k += blockDim.x) sh_seg[k] = segs[k0 + k];
    __syncthreads();
    if (present && !ce) {
      for (int k = 0; k < nk; ++k) {
        const float4 g = sh_seg[k];
        const float bs[4] = {fminf(g.x, g.z), fminf(g.y, g.w), fmaxf(g.x, g.z), fmaxf(g.y, g.w)};
        if (!aabb_hit(me.bb, bs)) continue;
        if (poly_segment_intersects(4, me.X, me.Y, g.x, g.y, g.z, g.w)) { ce = true; break; }
      }
    }
  }
  if (present) {
    b.coll[((size_t)s * 2 + 0) * b.max_veh + i] = cv;
    b.coll[((size_t)s * 2 + 1) * b.max_veh + i] = ce;
  }
}

__global__ void __launch_bounds__(SIM_THREADS)
sim_reset_kernel(CtrlSimBatch b, int T1) {  // T1 = steps + 1
  __shared__ Obb sh_obb[SIM_THREADS];
  __shared__ float4 sh_seg[SEG_TILE];
  const int s = blockIdx.x, i = threadIdx.x, N = b.max_veh;
  const int n = b.n_veh[s];
  const bool present = i < n;
  float x = 0, y = 0, heading = 0, speed = 0, len = 1, wid = 1;
  if (present) {
    const double* g0 = b.gt + (((size_t)s * N + i) * T1 + 0) * 4;  // float32-valued (expert states of the simulator)
    x = (float)g0[0]; y = (float)g0[1]; heading = (float)g0[2]; speed = (float)g0[3];
    len = b.veh_len[(size_t)s * N + i]; wid = b.veh_wid[(size_t)s * N + i];
    Body B;
    B.thr = B.brk = B.steer = 0.f; B.awake = 1.f; B.sleep_t = 0.f; B.om = 0.f; B.vx = B.vy = 0.f;
    box_local_center(wid / 2, len / 2, B.lcx, B.lcy);
    set_transform(B, 0.0f, 0.0f, (float)((double)heading - 3.14159265358979323846 * 0.5f));
    set_transform(B, x, y, B.ang);
    set_linvel(B, speed * cr_cosf(heading), speed * cr_sinf(heading));
    body_store(b.body + (size_t)s * B_FIELDS * N, N, i, B);
    float* o = b.obj + (size_t)s * O_FIELDS * N;
    o[O_X * N + i] = x; o[O_Y * N + i] = y; o[O_HEAD * N + i] = heading; o[O_SPEED * N + i] = speed;
  }
  __syncthreads();
  if (i == 0 && b.cstate) {  // contact state: every proxy freshly inserted, no contacts, first step has dtRatio 0
    __shared__ int scratch[CS_SCRATCH_WORDS];
    float* w = b.cstate + (size_t)s * cs_words(N);
    SV sv{b.body + (size_t)s * B_FIELDS * N, b.veh_len + (size_t)s * N, b.veh_wid + (size_t)s * N, N, n};
    SC c = cs_view(w, N, n, scratch);
    *c.new_contacts = 1; *c.n_contacts = 0; *c.inv_dt0 = 0.0f; *c.overflow = 0;
    for (int k = 0; k < n; ++k) cs_init_body(sv, c, k);
  }
  update_collision(b, s, i, n, present, x, y, heading, len, wid, sh_obb, sh_seg);
}

int set_trig_mode(int glibc) {
  const int v = glibc ? 1 : 0;
  const cudaError_t e = cudaMemcpyToSymbol(g_trig_glibc, &v, sizeof(int));
  if (e != cudaSuccess) return set_error(-5, "set_trig_mode: %s", cudaGetErrorString(e));
  return 0;
}

int launch_sim_reset(const CtrlSimBatch& b, const ModelCfg& mc, cudaStream_t st) {
  if (b.max_veh > SIM_THREADS) return set_error(-2, "sim: at most %d vehicles per scene", SIM_THREADS);
  sim_reset_kernel<<<b.n_scenes, SIM_THREADS, 0, st>>>(b, mc.steps + 1);
  CS_CHECK_LAUNCH("sim_reset");
  return 0;
}

__device__ __forceinline__ double py_mod(double a, double m) {
  double r = fmod(a, m);
  if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
  return r;
}
__device__ __forceinline__ double angle_sub_d(double cur, double tgt) {
  const double two_pi = 6.283185307179586, pi = 3.141592653589793;
  double d = py_mod(tgt - cur, two_pi);
  if (d > pi) d = -(two_pi - d);
  return d;
}

// ---- S5 + T1 ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SIM_THREADS)
observe_kernel(CtrlSimBatch b, int t, ModelCfg mc) {
  __shared__ double sx[SIM_THREADS], sy[SIM_THREADS], gx[SIM_THREADS], gy[SIM_THREADS];
  __shared__ uint8_t sex[SIM_THREADS];
  const int s = blockIdx.x, i = threadIdx.x, N = b.max_veh, T1 = mc.steps + 1;
  const int n = b.n_veh[s];
  const bool present = i < n;
  const size_t vi = (size_t)s * N + i;
  double px = 0, py = 0, head = 0, spd = 0;
  uint8_t ex = 0;
  if (present) {
    const float* o = b.obj + (size_t)s * O_FIELDS * N;
    const float fx = o[O_X * N + i], fy = o[O_Y * N + i], fh = o[O_HEAD * N + i], fs = o[O_SPEED * N + i];
    px = fx; py = fy; head = fh; spd = fs;
    const float vx = fs * cr_cosf(fh), vy = fs * cr_sinf(fh);  // Object::Velocity (include/object.h:152-154)
    ex = b.gt_valid[vi * T1 + t];
    if (t > 0 && !b.tr_exist[vi * T1 + t - 1]) ex = 0;
    b.tr_pos[(vi * T1 + t) * 2] = fx; b.tr_pos[(vi * T1 + t) * 2 + 1] = fy;
    b.tr_vel[(vi * T1 + t) * 2] = vx; b.tr_vel[(vi * T1 + t) * 2 + 1] = vy;
    b.tr_heading[vi * T1 + t] = fh;
    b.tr_exist[vi * T1 + t] = ex;
    // reward (utils/sim.py:83-141), float64 like numpy
    const double* go = b.goal + vi * 4;
    const bool prev = t > 0 && b.tr_reward[(vi * T1 + t - 1) * 8] != 0.f;
    const double ddx = go[0] - px, ddy = go[1] - py;
    const double dist = sqrt(ddx * ddx + ddy * ddy);
    float* rw = b.tr_reward + (vi * T1 + t) * 8;
    rw[0] = prev ? 1.f : (dist < mc.pos_tol ? 1.f : 0.f);
    rw[1] = fabs(angle_sub_d(go[2], head)) < mc.heading_tol ? 1.f : 0.f;
    rw[2] = fabs(go[3] - spd) < mc.speed_tol ? 1.f : 0.f;
    double nz = b.goal_norm[vi];
    if (nz == 0.0) nz = 1.0;
    const double gds = mc.goal_dist_scaling, rs = mc.reward_scaling;
    rw[3] = (float)(prev ? gds / rs : gds * (1 - dist / nz) / rs);
    rw[4] = (float)(gds * (1 - fabs(spd - go[3]) / 40.0) / rs);
    rw[5] = (float)(gds * (1 - fabs(angle_sub_d(head, go[2])) / (2 * 3.141592653589793)) / rs);
    rw[6] = b.coll[((size_t)s * 2 + 0) * N + i] ? 1.f : 0.f;
    rw[7] = b.coll[((size_t)s * 2 + 1) * N + i] ? 1.f : 0.f;
    if (t < mc.steps) {  // Policy.update_state (policy.py:68-105); actions/rtgs of t-1 were written when they were made
      double* hs = b.hist_state + (vi * mc.steps + t) * 8;
      hs[0] = px; hs[1] = py; hs[2] = vx; hs[3] = vy; hs[4] = head;
      hs[5] = b.veh_len[vi]; hs[6] = b.veh_wid[vi]; hs[7] = ex;
    }
    const double* g = b.gt + (vi * T1 + t) * 4;
    gx[i] = g[0]; gy[i] = g[1];
  }
  sx[i] = px; sy[i] = py; sex[i] = ex;
  __syncthreads();
  if (present) {  // nearest existing other vehicle, 0 when undefined (evaluator.py:87-104, dataset.py:202-236)
    double best = 1e300, bestg = 1e300;
    if (ex) {
      for (int j = 0; j < n; ++j) {
        if (j == i || !sex[j]) continue;
        const double dx = sx[i] - sx[j], dy = sy[i] - sy[j];
        best = fmin(best, dx * dx + dy * dy);
        const double ex2 = gx[i] - gx[j], ey2 = gy[i] - gy[j];
        bestg = fmin(bestg, ex2 * ex2 + ey2 * ey2);
      }
    }
    b.tr_nearest[(vi * T1 + t) * 2] = (ex && best < 1e299) ? sqrt(best) : 0.0;
    b.tr_nearest[(vi * T1 + t) * 2 + 1] = (ex && bestg < 1e299) ? sqrt(bestg) : 0.0;
  }
}

int launch_observe(const CtrlSimBatch& b, int t, const ModelCfg& mc, cudaStream_t st) {
  observe_kernel<<<b.n_scenes, SIM_THREADS, 0, st>>>(b, t, mc);
  CS_CHECK_LAUNCH("observe");
  return 0;
}

// ---- S5': real-time dense reward + RTG bookkeeping (real_time_rewards policies) -------------------------------------
// evaluators/evaluator.py:106-140 (compute_dense_reward), utils/data.py:152-290 (signed distance to the road-edge
// polylines, a numpy port of the Waymo off-road metric), datasets/rl_waymo/dataset.py:239-275 (compute_rewards),
// evaluators/policy_evaluator.py:123-149 (RTG series).  float64 in numpy's operation order (this file is built with
// -fmad=false).  One block per scene, one thread per vehicle; every thread walks all road-edge points (a warp reads the
// same point: broadcast loads).
__device__ __forceinline__ double sgn_d(double x) { return (double)(x > 0.0) - (double)(x < 0.0); }

// signed distance of (px, py) to one polyline of m + 1 points (utils/data.py:214-290)
__device__ double signed_distance_polyline(double px, double py, const double* __restrict__ q, int m) {
  const bool cyclic = ((q[0] - q[2 * m]) * (q[0] - q[2 * m]) + (q[1] - q[2 * m + 1]) * (q[1] - q[2 * m + 1])) < 1.0;
  // pass 1: nearest segment (first minimum, like np.argmin)
  int best = 0;
  double best_d = 0.0;
  for (int k = 0; k < m; ++k) {
    const double sx = q[2 * k], sy = q[2 * k + 1];
    const double ex = q[2 * k + 2] - sx, ey = q[2 * k + 3] - sy;      // start_to_end
    const double ax = px - sx, ay = py - sy;                            // start_to_point
    const double den = ex * ex + ey * ey;
    double t = (ax * ex + ay * ey) / den;
    if (t != t) t = 0.0;                                                // nan_to_num (0 / 0 of a degenerate segment)
    else if (t > 1.7976931348623157e308) t = 1.7976931348623157e308;    // +-inf -> +-DBL_MAX
    else if (t < -1.7976931348623157e308) t = -1.7976931348623157e308;
    const double tc = fmin(fmax(t, 0.0), 1.0);
    const double dx = ax - ex * tc, dy = ay - ey * tc;
    const double d = sqrt(dx * dx + dy * dy);
    if (k == 0 || d < best_d) { best_d = d; best = k; }
  }
  // pass 2: sign at the nearest segment from its own side value and, beyond its ends, the neighbour's and the convexity
  auto seg = [&](int k, double& ex, double& ey) { ex = q[2 * k + 2] - q[2 * k]; ey = q[2 * k + 3] - q[2 * k + 1]; };
  auto side = [&](int k) {
    double ex, ey;
    seg(k, ex, ey);
    const double ax = px - q[2 * k], ay = py - q[2 * k + 1];
    return sgn_d(ax * ey - ay * ex);
  };
  const int k = best;
  double ex, ey;
  seg(k, ex, ey);
  const double ax = px - q[2 * k], ay = py - q[2 * k + 1];
  double t = (ax * ex + ay * ey) / (ex * ex + ey * ey);
  if (t != t) t = 0.0;
  const double n = sgn_d(ax * ey - ay * ex);
  double sign;
  if (t < 0.0) {
    // convexity at vertex k: cross(previous segment, this segment), the previous of the first one is the LAST segment
    double pxe, pye;
    seg(k == 0 ? m - 1 : k - 1, pxe, pye);
    const bool convex = (pxe * ey - pye * ex) > 0.0;
    const double n_prior = k == 0 ? (cyclic ? side(m - 1) : n) : side(k - 1);
    sign = convex ? fmax(n, n_prior) : fmin(n, n_prior);
  } else if (t < 1.0) {
    sign = n;
  } else {
    double nxe, nye;
    seg(k == m - 1 ? 0 : k + 1, nxe, nye);
    const bool convex = (ex * nye - ey * nxe) > 0.0;
    const double n_next = k == m - 1 ? (cyclic ? side(0) : n) : side(k + 1);
    sign = convex ? fmax(n, n_next) : fmin(n, n_next);
  }
  return sign * best_d;
}

__global__ void __launch_bounds__(SIM_THREADS)
dense_reward_kernel(CtrlSimBatch b, CtrlSimRewardParams rp, int t, ModelCfg mc) {
  const int s = blockIdx.x, i = threadIdx.x, N = b.max_veh, T1 = mc.steps + 1;
  if (i >= b.n_veh[s]) return;
  const size_t vi = (size_t)s * N + i;
  // RTG of step t: start value, or the previous one minus the dense reward of step t-1 (policy_evaluator.py:123-149)
  if (t < mc.steps) {
    double* rt = b.rt_rtg + (vi * mc.steps + t) * 3;
    if (rp.return_mode == 3) {  // use_rtg = False: the policy's RTG buffer is never written (policies/policy.py:89-95)
      rt[0] = 0.0; rt[1] = 0.0; rt[2] = 0.0;
    } else if (t == 0) {
      if (rp.return_mode == 0) { rt[0] = b.rtg_init[vi * 3]; rt[1] = b.rtg_init[vi * 3 + 1]; rt[2] = b.rtg_init[vi * 3 + 2]; }
      else if (rp.return_mode == 2 && b.evaluated[vi]) { rt[0] = 0.0; rt[1] = -10.0; rt[2] = -10.0; }
      else { rt[0] = 10.0; rt[1] = 90.0; rt[2] = 90.0; }
    } else {
      const double* pr = b.rt_rtg + (vi * mc.steps + t - 1) * 3;
      const double* pd = b.tr_dense + (vi * T1 + t - 1) * 3;
      rt[0] = pr[0] - pd[0]; rt[1] = pr[1] - pd[1]; rt[2] = pr[2] - pd[2];
    }
  }
  const double e = b.tr_exist[vi * T1 + t] ? 1.0 : 0.0;
  const double px = b.tr_pos[(vi * T1 + t) * 2], py = b.tr_pos[(vi * T1 + t) * 2 + 1];
  // nearest road-edge polyline by |signed distance| (utils/data.py:208-211; first minimum, degenerate polylines skipped)
  const double* pts = b.edge_xy + (size_t)s * b.max_edge_pts * 2;
  const int* off = b.edge_off + (size_t)s * (b.max_edge_poly + 1);
  const int ne = b.n_edge[s];
  double sd = 0.0;
  bool have = false;
  for (int p = 0; p < ne; ++p) {
    const int m = off[p + 1] - off[p] - 1;
    if (m < 1) continue;
    const double d = signed_distance_polyline(px, py, pts + 2 * (size_t)off[p], m);
    if (!have || fabs(d) < fabs(sd)) { sd = d; have = true; }
  }
  const double edge = (-sd / rp.dist_to_road_edge_scaling_factor) * e;
  // nearest vehicle: tr_nearest[t] is the normalize=False distance x existence (observe_kernel)
  double* nr = b.tr_nearest + (vi * T1 + t) * 2;
  const double vv = fmin(fmax(nr[0], 0.0), rp.max_veh_veh_distance) / rp.max_veh_veh_distance;
  nr[0] *= rp.max_veh_veh_distance;  // what this mode records as 'nearest_dist' / 'gt_nearest_dist' (evaluator.py:126-127)
  nr[1] *= rp.max_veh_veh_distance;
  // QUIRK: goal / collision terms of step 0 - compute_rewards gets the reward HISTORY and evaluator.py:138 reads index 0
  const float* r0 = b.tr_reward + (vi * T1 + 0) * 8;
  const double rg = (double)r0[0] * e, rs = (double)r0[3] * e, rv = (double)r0[6] * e, re = (double)r0[7] * e;
  double goal = rg * rp.pos_target_achieved_rew_multiplier;
  if (!rp.remove_shaped_goal)
    goal = goal + (fmin(fmax(rs, rp.pos_goal_shaped_min), rp.pos_goal_shaped_max) - rp.pos_goal_shaped_max) * (1 / rp.pos_goal_shaped_max);
  const double veh = rp.remove_shaped_veh_reward ? -1 * rv * rp.veh_veh_collision_rew_multiplier
                                                 : vv - rv * rp.veh_veh_collision_rew_multiplier;
  const double road = rp.remove_shaped_edge_reward
                          ? -1 * re * rp.veh_edge_collision_rew_multiplier
                          : fmin(fmax(fabs(edge) * rp.dist_to_road_edge_scaling_factor, 0.0), 5.0) / 5.0 - re * rp.veh_edge_collision_rew_multiplier;
  double* d = b.tr_dense + (vi * T1 + t) * 3;
  d[0] = goal * e; d[1] = veh * e; d[2] = road * e;
}

int launch_dense_reward(const CtrlSimBatch& b, const CtrlSimRewardParams& rp, int t, const ModelCfg& mc, cudaStream_t st) {
  if (b.n_scenes <= 0) return 0;
  if (!b.edge_xy || !b.edge_off || !b.n_edge || !b.rt_rtg || !b.tr_dense || (rp.return_mode == 0 && !b.rtg_init))
    return set_error(-2, "dense_reward: the batch lacks the real-time-reward arrays (edge_xy / edge_off / n_edge / rt_rtg / tr_dense / rtg_init)");
  if (t < 0 || t > mc.steps) return set_error(-2, "dense_reward: step %d outside 0..%d", t, mc.steps);
  dense_reward_kernel<<<b.n_scenes, SIM_THREADS, 0, st>>>(b, rp, t, mc);
  CS_CHECK_LAUNCH("dense_reward");
  return 0;
}

// ---- T2: greedy focal grouping ------------------------------------------------------------------------------------
// One warp per scene. Context sets are 64-bit masks over vehicle ids (lists in the reference are always ascending).
__global__ void __launch_bounds__(32)
plan_groups_kernel(CtrlSimBatch b, int t, ModelCfg mc) {
  __shared__ double X[SIM_THREADS], Y[SIM_THREADS];
  __shared__ int unacc[SIM_THREADS];
  __shared__ unsigned long long sh_mask;
  const int s = blockIdx.x, lane = threadIdx.x, N = b.max_veh;
  const int n = b.n_veh[s];
  const int t0 = t < T ? 0 : t - (T - 1);
  const double* hs = b.hist_state + (size_t)s * N * mc.steps * 8;
  for (int i = lane; i < SIM_THREADS; i += 32) {
    if (i < n) { X[i] = hs[((size_t)i * mc.steps + t0) * 8]; Y[i] = hs[((size_t)i * mc.steps + t0) * 8 + 1]; }
    else { X[i] = 1e30; Y[i] = 1e30; }
  }
  int n_un = 0;
  for (int i = 0; i < N; ++i) {
    const int v = b.eval_order[(size_t)s * N + i];
    if (v < 0) break;
    if (lane == 0) unacc[n_un] = v;
    ++n_un;
  }
  __syncwarp();
  int ng = 0;
  while (n_un > 0) {
    const int focal = unacc[0];
    __syncwarp();
    if (lane == 0) for (int i = 1; i < n_un; ++i) unacc[i - 1] = unacc[i];
    --n_un;
    __syncwarp();
    const bool alive = hs[((size_t)focal * mc.steps + t) * 8 + 7] != 0.0;
    if (!alive || b.n_poly[s] == 0) {  // dead focal: action (0, 0) (autoregressive_policy.py:101-104,249-251)
      if (lane == 0) { b.next_action[((size_t)s * N + focal) * 2] = 0.0; b.next_action[((size_t)s * N + focal) * 2 + 1] = 0.0; }
      continue;
    }
    // distances of every vehicle to the focal at window step 0 (dataset.py:279-281)
    double d[2];
    bool valid[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = lane + 32 * h;
      const double dx = X[focal] - X[i], dy = Y[focal] - Y[i];
      d[h] = sqrt(dx * dx + dy * dy);
      valid[h] = i < n && d[h] < mc.agent_dist;
    }
    unsigned long long rel = (t == 0) ? 0ull : b.relevant[(size_t)s * N + focal];
    unsigned long long vmask = (unsigned long long)__ballot_sync(0xffffffffu, valid[0]) |
                               ((unsigned long long)__ballot_sync(0xffffffffu, valid[1]) << 32);
    unsigned long long closest;
    if (rel == 0ull) {  // the (up to) 24 nearest, then those within range, ascending ids (dataset.py:290-292)
      bool take[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        int rank = 0;
        for (int j = 0; j < n; ++j) {
          const double dxj = X[focal] - X[j], dyj = Y[focal] - Y[j];
          const double dj = sqrt(dxj * dxj + dyj * dyj);
          rank += (dj < d[h]) || (dj == d[h] && j < i);
        }
        take[h] = i < n && rank < A;
      }
      const unsigned long long near = (unsigned long long)__ballot_sync(0xffffffffu, take[0]) |
                                      ((unsigned long long)__ballot_sync(0xffffffffu, take[1]) << 32);
      closest = near & vmask;
    } else {
      closest = rel & vmask;  // sticky membership, pruned by range (dataset.py:296-302)
    }
    // served vehicles: the reference removes from the list it is iterating, so the element after a hit is skipped
    unsigned long long served_v = 1ull << focal;
    if (lane == 0) {
      int i = 0, cnt = n_un;
      while (i < cnt) {
        const int u = unacc[i];
        if ((closest >> u) & 1ull) {
          served_v |= 1ull << u;
          for (int k = i + 1; k < cnt; ++k) unacc[k - 1] = unacc[k];
          --cnt;
        }
        ++i;
      }
      sh_mask = served_v;
      unacc[SIM_THREADS - 1] = cnt;  // slot N-1 is never a live entry once at least one vehicle was popped
    }
    __syncwarp();
    served_v = sh_mask;
    n_un = unacc[SIM_THREADS - 1];
    __syncwarp();
    // relevant_agent_idxs of every served vehicle <- this group's context (autoregressive_policy.py:129-137)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = lane + 32 * h;
      if ((served_v >> i) & 1ull) b.relevant[(size_t)s * N + i] = closest;
    }
    // emit the group
    if (lane == 0) {
      int* mem = b.group_members + ((size_t)s * N + ng) * A;
      unsigned long long served_slots = 0;
      int k = 0;
      for (int v = 0; v < n && k < A; ++v)
        if ((closest >> v) & 1ull) {
          mem[k] = v;
          if ((served_v >> v) & 1ull) served_slots |= 1ull << k;
          ++k;
        }
      for (; k < A; ++k) mem[k] = -1;
      b.group_focal[(size_t)s * N + ng] = focal;
      b.group_served[(size_t)s * N + ng] = served_slots;
    }
    ++ng;
  }
  if (lane == 0) b.n_groups[s] = ng;
}

__global__ void scan_groups_kernel(CtrlSimBatch b, int* n_total) {
  // single block: exclusive scan of n_groups + compact group tables
  __shared__ int sh[1024];
  const int S = b.n_scenes, N = b.max_veh;
  int carry = 0;
  for (int base = 0; base < S; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < S ? b.n_groups[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int tmp = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += tmp;
      __syncthreads();
    }
    if (i < S) {
      const int off = carry + sh[threadIdx.x] - v;
      b.group_off[i] = off;
      for (int lg = 0; lg < v; ++lg) { b.group_scene[off + lg] = i; b.group_local[off + lg] = lg; }
    }
    carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) { b.group_off[S] = carry; *n_total = carry; }
  (void)N;
}

int launch_plan_groups(const CtrlSimBatch& b, int t, const ModelCfg& mc, int* n_total, cudaStream_t st) {
  if (b.max_veh > SIM_THREADS) return set_error(-2, "plan_groups: at most %d vehicles per scene", SIM_THREADS);
  plan_groups_kernel<<<b.n_scenes, 32, 0, st>>>(b, t, mc);
  CS_CHECK_LAUNCH("plan_groups");
  scan_groups_kernel<<<1, 1024, 0, st>>>(b, n_total);
  CS_CHECK_LAUNCH("scan_groups");
  return 0;
}

// ---- S6 + S1-S4 ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SIM_THREADS)
sim_step_kernel(CtrlSimBatch b, int t, ModelCfg mc) {
  __shared__ Obb sh_obb[SIM_THREADS];
  __shared__ float4 sh_seg[SEG_TILE];
  __shared__ int sh_scratch[CS_SCRATCH_WORDS];
  __shared__ CsScratch sh_solver;
  __shared__ bool sh_tele[SIM_THREADS];
  const int s = blockIdx.x, i = threadIdx.x, N = b.max_veh, T1 = mc.steps + 1;
  const int n = b.n_veh[s];
  const bool present = i < n;
  const bool use_contacts = mc.contacts && b.cstate != nullptr;
  const size_t vi = (size_t)s * N + i;
  float ox = 0, oy = 0, heading = 0, len = 1, wid = 1;
  if (present) {
    Body B;
    body_load(b.body + (size_t)s * B_FIELDS * N, N, i, B);
    float* o = b.obj + (size_t)s * O_FIELDS * N;
    len = b.veh_len[vi]; wid = b.veh_wid[vi];
    const bool exists_t = b.tr_exist[vi * T1 + t] != 0;
    double acc = 0.0, steer = 0.0;
    bool teleport = false;
    if (t >= mc.hist_steps - 1 && b.evaluated[vi]) {  // policy.act (autoregressive_policy.py:256-274)
      if (!exists_t) teleport = true;
      else { acc = b.next_action[vi * 2]; steer = b.next_action[vi * 2 + 1]; }
    } else {  // apply_gt_action (evaluators/evaluator.py:160-193)
      bool ex = b.gt_valid[vi * T1 + t] && b.gt_valid[vi * T1 + t + 1];
      if (t > 0 && !exists_t) ex = false;
      if (!ex) teleport = true;
      else {
        const double* g1 = b.gt + (vi * T1 + t + 1) * 4;  // float64: a scripted target need not be float32-valued
        const double vel_gt = g1[3], theta_gt = g1[2], L = len;
        const double sim_vel = o[O_SPEED * N + i], sim_theta = o[O_HEAD * N + i];
        acc = (vel_gt - sim_vel) / mc.dt_d;
        const double w = angle_sub_d(sim_theta, theta_gt) / mc.dt_d;
        const double C = 2.0 * L * w / (vel_gt + sim_vel + 1e-10);
        double stv = atan(2.0 * C / sqrt(4 - C * C));
        if (isnan(stv)) stv = 0.0;
        steer = fmin(fmax(stv, -0.7), 0.7);
      }
    }
    if (teleport) {  // veh.setPosition(-1e6, -1e6) -> b2Body::SetTransform keeps the angle
      o[O_X * N + i] = -1000000.f; o[O_Y * N + i] = -1000000.f;
      set_transform(B, -1000000.f, -1000000.f, B.ang);
    }
    b.tr_action[(vi * T1 + t) * 2] = acc; b.tr_action[(vi * T1 + t) * 2 + 1] = steer;
    b.hist_action[(vi * mc.steps + t) * 2] = acc; b.hist_action[(vi * mc.steps + t) * 2 + 1] = steer;
    // latch (python compares the float64, pybind narrows to float)
    if (acc > 0.0) {
      const float v = (float)acc;
      B.thr = v > 0.f ? 1.0f * v : 0.f * v;
      B.brk = 0.f;
    } else {
      const float v = (float)fabs(acc);
      if (!((double)fabsf(v) < 0.001)) { B.thr = 0.f; B.brk = 1.0f * v; }
    }
    B.steer = (float)steer;
    freecar_step(B, len, mc.dt);
    if (!use_contacts) island_solve(B, mc.dt);
    body_store(b.body + (size_t)s * B_FIELDS * N, N, i, B);
    sh_tele[i] = teleport;
  }
  if (use_contacts) {  // b2World::Step with contacts: one thread per scene (sim_contacts.cuh)
    __syncthreads();
    if (i == 0) {
      float* w = b.cstate + (size_t)s * cs_words(N);
      SV sv{b.body + (size_t)s * B_FIELDS * N, b.veh_len + (size_t)s * N, b.veh_wid + (size_t)s * N, N, n};
      SC c = cs_view(w, N, n, sh_scratch);
      for (int k = 0; k < n; ++k)
        if (sh_tele[k]) { cs_teleport(sv, c, k); *c.new_contacts = 1; }  // SetTransform happened before the FreeCar steps
      world_step(sv, c, mc.dt, &sh_solver);
    }
    __syncthreads();
  }
  if (present) {
    Body B;
    body_load(b.body + (size_t)s * B_FIELDS * N, N, i, B);
    float* o = b.obj + (size_t)s * O_FIELDS * N;
    ox = B.px; oy = B.py;
    heading = (float)((double)B.ang + 3.14159265358979323846 * 0.5f);
    o[O_X * N + i] = ox; o[O_Y * N + i] = oy; o[O_HEAD * N + i] = heading;
    o[O_SPEED * N + i] = sqrtf(B.vx * B.vx + B.vy * B.vy);
  }
  update_collision(b, s, i, n, present, ox, oy, heading, len, wid, sh_obb, sh_seg);
}

int launch_sim_step(const CtrlSimBatch& b, int t, const ModelCfg& mc, cudaStream_t st) {
  if (b.max_veh > SIM_THREADS) return set_error(-2, "sim: at most %d vehicles per scene", SIM_THREADS);
  sim_step_kernel<<<b.n_scenes, SIM_THREADS, 0, st>>>(b, t, mc);
  CS_CHECK_LAUNCH("sim_step");
  return 0;
}

// ---- S7 -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void hist_add(long long* h, double x, double scale, double offset) {
  // np.histogram with edges e_k = k*scale + offset, k = 0..200: [e_k, e_k+1) and the last bin closed on the right
  if (!(x >= offset) || !(x <= 200.0 * scale + offset)) return;
  int k = (int)floor((x - offset) / scale);
  k = max(0, min(199, k));
  while (k > 0 && x < (double)k * scale + offset) --k;
  while (k < 199 && x >= (double)(k + 1) * scale + offset) ++k;
  atomicAdd(reinterpret_cast<unsigned long long*>(h + k), 1ull);
}

__global__ void __launch_bounds__(SIM_THREADS)
metrics_kernel(CtrlSimBatch b, ModelCfg mc, double* __restrict__ out_scene, long long* __restrict__ out_hist) {
  __shared__ double red[5][SIM_THREADS];
  const int s = blockIdx.x, i = threadIdx.x, N = b.max_veh, T1 = mc.steps + 1;
  const int n = b.n_veh[s];
  const size_t vi = (size_t)s * N + i;
  double goal = 0, has = 0, coll = 0, off = 0, ade = 0, fde = 0;
  if (i < n && b.evaluated[vi]) {
    int cnt = 0, first = -1, last = -1;
    for (int t = mc.hist_steps; t < T1; ++t)
      if (b.tr_exist[vi * T1 + t]) { ++cnt; if (first < 0) first = t; last = t; }
    if (cnt > 0) {
      has = 1;
      double sum = 0;
      const double lin_w = 0.5 * (100.0 / 30.0), nd_w = 0.5 * (100.0 / 40.0);
      for (int t = mc.hist_steps; t < T1; ++t) {
        if (!b.tr_exist[vi * T1 + t]) continue;
        const float* rw = b.tr_reward + (vi * T1 + t) * 8;
        if (rw[0] == 1.f) goal = 1;
        if (rw[6] == 1.f) coll = 1;
        if (rw[7] == 1.f) off = 1;
        const double* g = b.gt + (vi * T1 + t) * 4;
        const double dx = (double)b.tr_pos[(vi * T1 + t) * 2] - (double)g[0];
        const double dy = (double)b.tr_pos[(vi * T1 + t) * 2 + 1] - (double)g[1];
        const double d = sqrt(dx * dx + dy * dy);
        sum += d;
        if (t == last) fde = d;
        const double vx = b.tr_vel[(vi * T1 + t) * 2], vy = b.tr_vel[(vi * T1 + t) * 2 + 1];
        hist_add(out_hist + 0 * 200, fmin(fmax(sqrt(vx * vx + vy * vy), 0.0), 30.0), lin_w, 0.0);
        hist_add(out_hist + 1 * 200, fmin(fmax((double)g[3], 0.0), 30.0), lin_w, 0.0);
        hist_add(out_hist + 2 * 200, fmin(fmax((double)b.tr_heading[vi * T1 + t] / mc.dt_d, -50.0), 50.0), 0.5, -50.0);
        hist_add(out_hist + 3 * 200, fmin(fmax((double)g[2] / mc.dt_d, -50.0), 50.0), 0.5, -50.0);
        if (t != first && t != last) {  // accel: end points of the masked sequence are dropped (policy_evaluator.py:225-232)
          double ga = 0.0;
          if (t > 0 && t < mc.steps - 1)
            ga = ((double)b.gt[(vi * T1 + t + 1) * 4 + 3] - (double)b.gt[(vi * T1 + t - 1) * 4 + 3]) / (2 * mc.dt_d);
          double gn = (fmin(fmax(ga, mc.min_accel), mc.max_accel) - mc.min_accel) / (mc.max_accel - mc.min_accel);
          const int nb = N_ACT / mc.n_steer;
          gn = rint(gn * (nb - 1)) / (nb - 1);
          gn = gn * (mc.max_accel - mc.min_accel) + mc.min_accel;
          const double sa = (t < mc.steps) ? b.tr_action[(vi * T1 + t) * 2] : 0.0;
          // edges arange(21)*2-20: reuse hist_add with 20 live bins (x <= 20 enforced by the 200-bin guard below)
          if (sa >= -20.0 && sa <= 20.0) { int k = (int)floor((sa + 20.0) / 2.0); k = min(19, max(0, k)); atomicAdd(reinterpret_cast<unsigned long long*>(out_hist + 4 * 200 + k), 1ull); }
          if (gn >= -20.0 && gn <= 20.0) { int k = (int)floor((gn + 20.0) / 2.0); k = min(19, max(0, k)); atomicAdd(reinterpret_cast<unsigned long long*>(out_hist + 5 * 200 + k), 1ull); }
        }
        hist_add(out_hist + 6 * 200, fmin(fmax(b.tr_nearest[(vi * T1 + t) * 2], 0.0), 40.0), nd_w, 0.0);
        hist_add(out_hist + 7 * 200, fmin(fmax(b.tr_nearest[(vi * T1 + t) * 2 + 1], 0.0), 40.0), nd_w, 0.0);
      }
      ade = sum / cnt;
    }
  }
  red[0][i] = goal; red[1][i] = has; red[2][i] = coll; red[3][i] = off; red[4][i] = ade;
  __shared__ double redf[SIM_THREADS];
  redf[i] = fde;
  __syncthreads();
  if (i == 0) {
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < SIM_THREADS; ++j) {  // fixed order: deterministic
      a[0] += red[0][j]; a[1] += red[1][j]; a[2] += red[2][j]; a[3] += red[3][j]; a[4] += red[4][j]; a[5] += redf[j];
    }
    double* o = out_scene + (size_t)s * 8;
    o[0] = a[0]; o[1] = a[1];
    o[2] = a[1] > 0 ? a[2] / a[1] : 0.0;  // per-scene mean (policy_evaluator.py:246-248)
    o[3] = a[1] > 0 ? a[3] / a[1] : 0.0;
    o[4] = a[1] > 0 ? 1.0 : 0.0;
    o[5] = a[4]; o[6] = a[5]; o[7] = 0.0;
  }
}

int launch_metrics(const CtrlSimBatch& b, const ModelCfg& mc, double* out_scene, long long* out_hist, cudaStream_t st) {
  metrics_kernel<<<b.n_scenes, SIM_THREADS, 0, st>>>(b, mc, out_scene, out_hist);
  CS_CHECK_LAUNCH("metrics");
  return 0;
}

}  // namespace ctrlsim
