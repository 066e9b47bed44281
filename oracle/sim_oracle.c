/* ORACLE (test infrastructure only; never linked or loaded by the product path).
 *
 * Plain-C restatement of the reference simulator step for the closed-loop rollout, contact-free subset:
 *   S1  FreeCar::Throttle/Brake/Turn/Step      nocturne/cpp/src/physics/FreeCar.cpp:66-86,88-186, defines.h:4-11
 *   S2  b2World::Step -> b2Island::Solve        third_party/box2d/src/dynamics/b2_island.cpp:188-388 (gravity 0,
 *       damping 0, no contacts/joints: velocity clamp, c += h*v, a += h*w, sleep bookkeeping),
 *       b2Body::SetLinearVelocity/SetAngularVelocity/SetAwake  include/box2d/b2_body.h:501-530,637-658,
 *       constants include/box2d/b2_common.h:41,95-119 (b2_maxTranslation patched to 5.0)
 *   S3  Vehicle::Step, CreatePhysicsBody, set_position  nocturne/cpp/src/vehicle.cc:25-66,75-88,137-179;
 *       Object::BoundingPolygon nocturne/cpp/src/object.cc:14-28; Velocity include/object.h:152-154
 *   S4  Scenario::UpdateCollision nocturne/cpp/src/scenario.cc:294-328; SAT polygon.cc:19-27,84-98;
 *       polygon-segment intersection.cc:200-233; AABB strict prefilter include/geometry/aabb.h:47-50
 *       (the BVH is only a broad phase whose leaf test is that AABB test, bvh.h:176-190 -> brute force here).
 * Box2D's contact solver is NOT restated (DESIGN.md "out of scope / staged"): bodies pass through each other.
 *
 * Floating point: fp32 with separately rounded multiply/add exactly where the reference has them (compile with
 * -ffp-contract=off), the same libm entry points the reference calls (sinf, cosf, tanf, double atan, sqrtf), and
 * the two places where the reference computes in double because of an M_PI literal.
 * Pinned against the real nocturne_cpp (oracle/_ref) by tests/test_sim_oracle.py and tests/golden/sim_*.npz.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define B2_PI 3.14159265359f
static const float kMaxTranslation = 5.0f;
static const float kMaxRotation = 0.5f * B2_PI;
static const float kTimeToSleep = 0.5f;
static const float kLinSleepTol = 0.01f;
static const float kAngSleepTol = 2.0f / 180.0f * B2_PI;

typedef struct {
  int n;
  /* Box2D body */
  float *px, *py, *ang, *vx, *vy, *om, *sleep_t;
  float *cx, *cy, *lcx, *lcy; /* b2Sweep::c and b2Sweep::localCenter (the fp32 centroid of the box is not exactly 0) */
  uint8_t* awake;
  /* FreeCar */
  float *thr, *brk, *steer, *len, *wid;
  /* Nocturne object */
  float *ox, *oy, *heading, *speed;
  uint8_t *coll_veh, *coll_edge;
} SimO;

static void body_set_awake(SimO* s, int i, int flag) {
  if (flag) {
    s->awake[i] = 1;
    s->sleep_t[i] = 0.0f;
  } else {
    s->awake[i] = 0;
    s->sleep_t[i] = 0.0f;
    s->vx[i] = s->vy[i] = 0.0f;
    s->om[i] = 0.0f;
  }
}
static void body_set_linvel(SimO* s, int i, float vx, float vy) {
  if (vx * vx + vy * vy > 0.0f) body_set_awake(s, i, 1);
  s->vx[i] = vx;
  s->vy[i] = vy;
}
static void body_set_angvel(SimO* s, int i, float w) {
  if (w * w > 0.0f) body_set_awake(s, i, 1);
  s->om[i] = w;
}

/* b2PolygonShape::ComputeMass (b2_polygon_shape.cpp:357-431) for SetAsBox(hx, hy) with density 20, then
 * b2Body::ResetMassData (b2_body.cpp:290-354): localCenter = (mass * centroid) * (1 / mass). Only the centre is
 * needed (no forces/contacts on this path). */
static void box_local_center(float hx, float hy, float* lcx, float* lcy) {
  const float vx[4] = {-hx, hx, hx, -hx}, vy[4] = {-hy, -hy, hy, hy};
  float cx = 0.0f, cy = 0.0f, area = 0.0f;
  const float sx = vx[0], sy = vy[0];
  const float k_inv3 = 1.0f / 3.0f;
  for (int i = 0; i < 4; ++i) {
    float e1x = vx[i] - sx, e1y = vy[i] - sy;
    int j = (i + 1 < 4) ? i + 1 : 0;
    float e2x = vx[j] - sx, e2y = vy[j] - sy;
    float D = e1x * e2y - e1y * e2x;
    float tri = 0.5f * D;
    area += tri;
    float k = tri * k_inv3;
    cx += k * (e1x + e2x);
    cy += k * (e1y + e2y);
  }
  float mass = 20.f * area;
  float inv_area = 1.0f / area;
  cx *= inv_area;
  cy *= inv_area;
  float mcx = cx + sx, mcy = cy + sy;
  float lx = mass * mcx, ly = mass * mcy;
  float inv_mass = 1.0f / mass;
  *lcx = lx * inv_mass;
  *lcy = ly * inv_mass;
}

/* b2Body::SetTransform (b2_body.cpp:420-445): p = position, c = q * localCenter + p */
static void body_set_transform(SimO* s, int i, float x, float y, float angle) {
  float qs = sinf(angle), qc = cosf(angle);
  s->px[i] = x;
  s->py[i] = y;
  s->cx[i] = (qc * s->lcx[i] - qs * s->lcy[i]) + x;
  s->cy[i] = (qs * s->lcx[i] + qc * s->lcy[i]) + y;
  s->ang[i] = angle;
}

/* Vehicle::CreatePhysicsBody (vehicle.cc:137-179): angle = heading - pi/2 in double, v = speed*(cosf, sinf). */
void simo_spawn(SimO* s, int i, float x, float y, float heading, float speed, float length, float width) {
  s->len[i] = length;
  s->wid[i] = width;
  s->thr[i] = s->brk[i] = s->steer[i] = 0.0f;
  s->awake[i] = 1;
  s->sleep_t[i] = 0.0f;
  s->om[i] = 0.0f;
  box_local_center(width / 2, length / 2, &s->lcx[i], &s->lcy[i]); /* shape.SetAsBox(m_Width/2, m_Length/2) */
  body_set_transform(s, i, 0.0f, 0.0f, (float)((double)heading - M_PI * 0.5f)); /* SetAngle */
  body_set_transform(s, i, x, y, s->ang[i]);                                     /* SetPosition */
  s->vx[i] = s->vy[i] = 0.0f;
  body_set_linvel(s, i, speed * cosf(heading), speed * sinf(heading));
  s->ox[i] = x;
  s->oy[i] = y;
  s->heading[i] = heading;
  s->speed[i] = speed;
  s->coll_veh[i] = s->coll_edge[i] = 0;
}

/* The evaluator's action latch (policies/autoregressive_policy.py:268-272, evaluators/evaluator.py:188-192):
 * accel > 0 -> Throttle(accel); else Brake(|accel|) which ignores |value| < 1e-3; then Turn(steer). */
void simo_set_action(SimO* s, int i, float accel, float steer) {
  if (accel > 0.0f) {
    s->thr[i] = 1.0f * accel;
    s->brk[i] = 0.0f;
  } else {
    float b = fabsf(accel);
    if (!((double)fabsf(b) < 0.001)) {
      s->thr[i] = 0.0f;
      s->brk[i] = 1.0f * b;
    }
  }
  s->steer[i] = steer;
}

/* veh.setPosition(x, y) (vehicle.cc:75-81 -> BaseCar::SetPosition -> b2Body::SetTransform) */
void simo_teleport(SimO* s, int i, float x, float y) {
  s->ox[i] = x;
  s->oy[i] = y;
  body_set_transform(s, i, x, y, s->ang[i]);
}

static float dampen(float speed, float target, float damping, float dt) {
  float red = damping * dt;
  if (speed - target > red) return speed - red;
  if (speed - target < -red) return speed + red;
  return target;
}

static void freecar_step(SimO* s, int i, float dt) {
  float target, acc;
  float thr = s->thr[i], brk = s->brk[i], st = s->steer[i];
  if (thr > 0.0f) {
    if (thr > brk) { target = 50.0f; acc = thr - brk; }
    else { target = 0.0f; acc = brk - thr; }
  } else {
    if (thr < -brk) { target = -5.0f; acc = -thr - brk; }
    else { target = 0.0f; acc = brk + thr; }
  }
  float w = s->om[i];
  float beta = (float)atan(0.5 * (double)tanf(st));
  float c = cosf(s->ang[i] + beta);
  float sn = sinf(s->ang[i] + beta);
  float fx = -sn, fy = c, rx = c, ry = sn;
  float vf = s->vx[i] * fx + s->vy[i] * fy;
  float vr = s->vx[i] * rx + s->vy[i] * ry;
  float dv = acc * dt;
  if (vf < target) vf = fminf(vf + dv, target);
  else vf = fmaxf(vf - dv, target);
  float w_steer = 0.0f;
  if (fabs((double)st) > 0.0000001) {
    float ray = 1.f / tanf(st) * s->len[i] / cosf(beta);
    w_steer = vf / ray;
  }
  vr = dampen(vr, 0.0f, 25.f, dt);
  w = dampen(w, w_steer, 10.f, dt);
  float sx = rx * vr + fx * vf;
  float sy = ry * vr + fy * vf;
  body_set_linvel(s, i, sx, sy);
  body_set_angvel(s, i, w);
}

static void island_solve(SimO* s, int i, float h) {
  if (!s->awake[i]) return; /* b2World::Solve only seeds islands from awake bodies */
  float vx = s->vx[i], vy = s->vy[i], w = s->om[i];
  /* dynamic body, zero force/gravity/damping: v += h*invMass*(0) ; v *= 1/(1+h*0) leave v, w bit-identical */
  float tx = h * vx, ty = h * vy;
  if (tx * tx + ty * ty > kMaxTranslation * kMaxTranslation) {
    float ratio = kMaxTranslation / sqrtf(tx * tx + ty * ty);
    vx *= ratio;
    vy *= ratio;
  }
  float rot = h * w;
  if (rot * rot > kMaxRotation * kMaxRotation) {
    float ratio = kMaxRotation / fabsf(rot);
    w *= ratio;
  }
  s->cx[i] += h * vx;
  s->cy[i] += h * vy;
  s->ang[i] += h * w;
  { /* b2Body::SynchronizeTransform (b2_body.h:859-863) */
    float qs = sinf(s->ang[i]), qc = cosf(s->ang[i]);
    s->px[i] = s->cx[i] - (qc * s->lcx[i] - qs * s->lcy[i]);
    s->py[i] = s->cy[i] - (qs * s->lcx[i] + qc * s->lcy[i]);
  }
  s->vx[i] = vx;
  s->vy[i] = vy;
  s->om[i] = w;
  if (w * w > kAngSleepTol * kAngSleepTol || vx * vx + vy * vy > kLinSleepTol * kLinSleepTol) {
    s->sleep_t[i] = 0.0f;
  } else {
    s->sleep_t[i] += h;
    if (s->sleep_t[i] >= kTimeToSleep) body_set_awake(s, i, 0);
  }
}

/* ---- geometry -------------------------------------------------------------------------------------------- */
static float cross2(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }

static void obb(const SimO* s, int i, float* X, float* Y) {
  float sh = sinf(s->heading[i]), ch = cosf(s->heading[i]);
  float hl = s->len[i] * 0.5f, hw = s->wid[i] * 0.5f;
  const float lx[4] = {hl, -hl, -hl, hl};
  const float ly[4] = {hw, hw, -hw, -hw};
  for (int k = 0; k < 4; ++k) {
    X[k] = (lx[k] * ch - ly[k] * sh) + s->ox[i];
    Y[k] = (lx[k] * sh + ly[k] * ch) + s->oy[i];
  }
}

/* Separates(edge, polygon): all vertices strictly to the right of the edge (polygon.cc:19-27) */
static int separates(float e0x, float e0y, float e1x, float e1y, int n, const float* X, const float* Y) {
  float dx = e1x - e0x, dy = e1y - e0y;
  for (int k = 0; k < n; ++k)
    if (cross2(X[k] - e0x, Y[k] - e0y, dx, dy) <= 0.0f) return 0;
  return 1;
}

int simo_poly_intersects(int n1, const float* X1, const float* Y1, int n2, const float* X2, const float* Y2) {
  for (int k = 0; k < n1; ++k) {
    int k0 = (k == n1 - 1) ? n1 - 1 : k, k1 = (k == n1 - 1) ? 0 : k + 1;
    /* Edges(): (v0,v1),(v1,v2),...,(v_{n-1},v0) (polygon.cc:46-55) */
    if (separates(X1[k0], Y1[k0], X1[k1], Y1[k1], n2, X2, Y2)) return 0;
  }
  for (int k = 0; k < n2; ++k) {
    int k0 = (k == n2 - 1) ? n2 - 1 : k, k1 = (k == n2 - 1) ? 0 : k + 1;
    if (separates(X2[k0], Y2[k0], X2[k1], Y2[k1], n1, X1, Y1)) return 0;
  }
  return 1;
}

static int poly_contains(int n, const float* X, const float* Y, float px, float py) {
  for (int i = 1; i < n; ++i)
    if (cross2(px - X[i - 1], py - Y[i - 1], X[i] - X[i - 1], Y[i] - Y[i - 1]) > 0.0f) return 0;
  return cross2(px - X[n - 1], py - Y[n - 1], X[0] - X[n - 1], Y[0] - Y[n - 1]) <= 0.0f;
}

int simo_poly_segment_intersects(int n, const float* X, const float* Y, float ax, float ay, float bx, float by) {
  if (ax == bx && ay == by) return poly_contains(n, X, Y, ax, ay);
  float dx = bx - ax, dy = by - ay;
  float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
  for (int k = 0; k < n; ++k) {
    float cur = cross2(X[k] - ax, Y[k] - ay, dx, dy);
    mn = fminf(mn, cur);
    mx = fmaxf(mx, cur);
  }
  if (mx < 0.0f || mn > 0.0f) return 0;
  for (int k = 0; k < n; ++k) {
    int k1 = (k == n - 1) ? 0 : k + 1;
    float ex = X[k1] - X[k], ey = Y[k1] - Y[k];
    float v0 = cross2(ax - X[k], ay - Y[k], ex, ey);
    float v1 = cross2(bx - X[k], by - Y[k], ex, ey);
    if (v0 > 0.0f && v1 > 0.0f) return 0;
  }
  return 1;
}

static void aabb4(const float* X, const float* Y, float* b) {
  b[0] = b[2] = X[0];
  b[1] = b[3] = Y[0];
  for (int k = 1; k < 4; ++k) {
    b[0] = fminf(b[0], X[k]); b[2] = fmaxf(b[2], X[k]);
    b[1] = fminf(b[1], Y[k]); b[3] = fmaxf(b[3], Y[k]);
  }
}
static int aabb_hit(const float* a, const float* b) {
  return a[0] < b[2] && a[2] > b[0] && a[1] < b[3] && a[3] > b[1];
}

/* segs: [nseg][4] = (x0, y0, x1, y1) for every consecutive point pair of every road_edge polyline
 * (scenario.cc:1037-1042). Flags are reset every step (scenario.cc:275). */
void simo_update_collision(SimO* s, const float* segs, int nseg) {
  int n = s->n;
  for (int i = 0; i < n; ++i) s->coll_veh[i] = s->coll_edge[i] = 0;
  for (int i = 0; i < n; ++i) {
    float Xi[4], Yi[4], bi[4];
    obb(s, i, Xi, Yi);
    aabb4(Xi, Yi, bi);
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      float Xj[4], Yj[4], bj[4];
      obb(s, j, Xj, Yj);
      aabb4(Xj, Yj, bj);
      if (!aabb_hit(bi, bj)) continue;
      if (simo_poly_intersects(4, Xi, Yi, 4, Xj, Yj)) s->coll_veh[i] = 1;
    }
    for (int k = 0; k < nseg; ++k) {
      const float* g = segs + 4 * k;
      float bs[4] = {fminf(g[0], g[2]), fminf(g[1], g[3]), fmaxf(g[0], g[2]), fmaxf(g[1], g[3])};
      if (!aabb_hit(bi, bs)) continue;
      if (simo_poly_segment_intersects(4, Xi, Yi, g[0], g[1], g[2], g[3])) s->coll_edge[i] = 1;
    }
  }
}

/* Scenario::Step (scenario.cc:266-292): all FreeCar::Step, then the world step, then Vehicle::Step, then collisions */
void simo_step(SimO* s, float dt, const float* segs, int nseg) {
  for (int i = 0; i < s->n; ++i) freecar_step(s, i, dt);
  for (int i = 0; i < s->n; ++i) island_solve(s, i, dt);
  for (int i = 0; i < s->n; ++i) {
    s->ox[i] = s->px[i];
    s->oy[i] = s->py[i];
    s->speed[i] = sqrtf(s->vx[i] * s->vx[i] + s->vy[i] * s->vy[i]);
    s->heading[i] = (float)((double)s->ang[i] + M_PI * 0.5f);
  }
  simo_update_collision(s, segs, nseg);
}

/* Object::Velocity(): speed * (cosf(heading), sinf(heading)) (include/object.h:152-154) */
void simo_velocity(const SimO* s, int i, float* vx, float* vy) {
  *vx = s->speed[i] * cosf(s->heading[i]);
  *vy = s->speed[i] * sinf(s->heading[i]);
}
