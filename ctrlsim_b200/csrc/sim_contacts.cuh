// Box2D contact response for the vehicle bodies of one scene (SURVEY 7.3-1 stage 2), run by ONE thread per scene inside
// sim_step_kernel.  Included by sim.cu only (compiled with -fmad=false: every multiply and add rounds separately, in the
// reference's operation order).
//
// Restated from third_party/box2d 2.4.1 as vendored by the reference (file:line):
//   mass data            b2PolygonShape::ComputeMass b2_polygon_shape.cpp:357-431, b2Body::ResetMassData b2_body.cpp:290-354
//   fat AABBs / pairs    b2PolygonShape::ComputeAABB :338-355, b2Fixture::Synchronize b2_fixture.cpp:156-178,
//                        b2DynamicTree::CreateProxy / MoveProxy b2_dynamic_tree.cpp:110-192, b2BroadPhase::UpdatePairs,
//                        b2ContactManager::AddPair / Collide b2_contact_manager.cpp:109-293
//   manifold             b2CollidePolygons b2_collide_polygon.cpp:25-243, b2ClipSegmentToLine b2_collision.cpp:205-237
//   contact update       b2Contact::Update b2_contact.cpp:163-252 (warm-start impulses matched by feature id)
//   islands + solver     b2World::Solve b2_world.cpp:394-582, b2Island::Solve b2_island.cpp:188-388, b2ContactSolver
//                        b2_contact_solver.cpp:51-755 (block solver; 8 velocity / 3 position iterations,
//                        PhysicsSimulation.cpp:22-24), b2WorldManifold::Initialize b2_collision.cpp:26-90
// Same statement of the algorithm as the CPU oracle (oracle/sim_oracle.c, which is pinned bit for bit against the real
// nocturne_cpp); the limits of both are the same: fixture "A" of a pair is the vehicle created first (what the reference's
// dynamic tree gives, also after the evaluator's per-file world wipe), pairs found by one broad-phase update become
// contacts in (moved body, other body) order, continuous collision never acts on two non-bullet dynamic bodies, pairs of vehicles
// parked at (-1e6, -1e6) are skipped (never observed, re-teleported before every step).  Capacity: CS_MAX_CONTACTS
// broad-phase pairs per scene and CS_MAX_ISLAND_CONTACTS touching contacts per island; beyond that the newest are dropped.
#pragma once
// (included inside namespace ctrlsim, after the body layout enums and cr_sinf / cr_cosf of sim.cu)

constexpr int CS_MAX_BODIES = 64;            // == CTRLSIM_MAX_VEH
constexpr int CS_MAX_CONTACTS = 128;
constexpr int CS_MAX_ISLAND_CONTACTS = 32;
constexpr int CS_CONTACT_WORDS = 20;
#define B2_LINEAR_SLOP 0.005f
#define B2_POLY_RADIUS (2.0f * B2_LINEAR_SLOP)
#define B2_AABB_EXT 0.1f
#define B2_AABB_MULT 4.0f
#define B2_BAUMGARTE 0.2f
#define B2_MAX_LIN_CORR 0.2f

// view of one scene's bodies in CtrlSimBatch.body ([16][N] floats) and sizes
struct SV {
  float* base; const float* len_; const float* wid_; int N; int n;
  __device__ float& px(int i) const { return base[B_PX * N + i]; }
  __device__ float& py(int i) const { return base[B_PY * N + i]; }
  __device__ float& cx(int i) const { return base[B_CX * N + i]; }
  __device__ float& cy(int i) const { return base[B_CY * N + i]; }
  __device__ float& lcx(int i) const { return base[B_LCX * N + i]; }
  __device__ float& lcy(int i) const { return base[B_LCY * N + i]; }
  __device__ float& ang(int i) const { return base[B_ANG * N + i]; }
  __device__ float& vx(int i) const { return base[B_VX * N + i]; }
  __device__ float& vy(int i) const { return base[B_VY * N + i]; }
  __device__ float& om(int i) const { return base[B_OM * N + i]; }
  __device__ float& sleep_t(int i) const { return base[B_SLEEP * N + i]; }
  __device__ float& awake(int i) const { return base[B_AWAKE * N + i]; }
  __device__ float len(int i) const { return len_[i]; }
  __device__ float wid(int i) const { return wid_[i]; }
};
__device__ static void sv_set_awake(const SV& s, int i, bool flag) {
  s.awake(i) = flag ? 1.f : 0.f;
  s.sleep_t(i) = 0.f;
  if (!flag) { s.vx(i) = 0.f; s.vy(i) = 0.f; s.om(i) = 0.f; }
}

typedef struct { float x, y; } V2;
typedef struct { V2 p; float s, c; } XF; /* b2Transform: p, q = (sin, cos) */
typedef struct { float lx, ly, ni, ti; uint32_t key; } MPoint;
typedef struct {
  int a, b;       /* body indices, a = fixture A */
  int touching;
  int type;       /* 1 = e_faceA, 2 = e_faceB */
  V2 ln, lp;      /* localNormal, localPoint */
  int count;
  MPoint pt[2];
  int pad;
} Contact;

typedef struct {
  int n;
  float *mass, *inv_mass, *inv_i;
  float* fat;      /* [n][4] lower.x, lower.y, upper.x, upper.y : the proxy AABB held by the tree */
  int* moved;
  float *c0x, *c0y, *a0;
  int* n_contacts;
  Contact* ct;
  int* new_contacts;
  float* inv_dt0;
  int* overflow;  // broad-phase pairs / island contacts dropped because a capacity was exceeded (parity no longer holds)
  float friction;
  int *island_flag, *stack, *ibody, *iidx; /* island scratch */
} SC;


__device__ static V2 v2(float x, float y) { V2 r = {x, y}; return r; }
__device__ static V2 v2add(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
__device__ static V2 v2sub(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
__device__ static V2 v2mul(float s, V2 a) { return v2(s * a.x, s * a.y); }
__device__ static float v2dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ static float v2cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
__device__ static V2 v2cross_vs(V2 a, float s) { return v2(s * a.y, -s * a.x); }  /* b2Cross(vec, scalar) */
__device__ static V2 v2cross_sv(float s, V2 a) { return v2(-s * a.y, s * a.x); }  /* b2Cross(scalar, vec) */
__device__ static V2 rot_mul(float qs, float qc, V2 v) { return v2(qc * v.x - qs * v.y, qs * v.x + qc * v.y); }
__device__ static V2 rot_mulT(float qs, float qc, V2 v) { return v2(qc * v.x + qs * v.y, -qs * v.x + qc * v.y); }
__device__ static V2 xf_mul(XF T, V2 v) {
  float x = (T.c * v.x - T.s * v.y) + T.p.x;
  float y = (T.s * v.x + T.c * v.y) + T.p.y;
  return v2(x, y);
}
__device__ static V2 xf_mulT(XF T, V2 v) {
  float px = v.x - T.p.x, py = v.y - T.p.y;
  return v2(T.c * px + T.s * py, -T.s * px + T.c * py);
}
__device__ static XF body_xf(const SV& s, int i) {
  XF T;
  T.p = v2(s.px(i), s.py(i));
  T.s = cr_sinf(s.ang(i));
  T.c = cr_cosf(s.ang(i));
  return T;
}
__device__ static void box_verts(const SV& s, int i, V2* v, V2* nrm) {
  float hx = s.wid(i) / 2, hy = s.len(i) / 2; /* shape.SetAsBox(m_Width/2, m_Length/2) */
  v[0] = v2(-hx, -hy); v[1] = v2(hx, -hy); v[2] = v2(hx, hy); v[3] = v2(-hx, hy);
  nrm[0] = v2(0.0f, -1.0f); nrm[1] = v2(1.0f, 0.0f); nrm[2] = v2(0.0f, 1.0f); nrm[3] = v2(-1.0f, 0.0f);
}

/* b2PolygonShape::ComputeMass + b2Body::ResetMassData for the single box fixture, density 20 */
__device__ static void box_mass(float hx, float hy, float* mass, float* inv_mass, float* inv_i) {
  const V2 vs[4] = {{-hx, -hy}, {hx, -hy}, {hx, hy}, {-hx, hy}};
  V2 center = {0.0f, 0.0f};
  float area = 0.0f, I = 0.0f;
  V2 s0 = vs[0];
  const float k_inv3 = 1.0f / 3.0f;
  for (int i = 0; i < 4; ++i) {
    V2 e1 = v2sub(vs[i], s0);
    V2 e2 = i + 1 < 4 ? v2sub(vs[i + 1], s0) : v2sub(vs[0], s0);
    float D = v2cross(e1, e2);
    float tri = 0.5f * D;
    area += tri;
    V2 e12 = v2add(e1, e2);
    float k = tri * k_inv3;
    center.x += k * e12.x;
    center.y += k * e12.y;
    float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
    float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
    float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
    I += (0.25f * k_inv3 * D) * (intx2 + inty2);
  }
  float m = 20.f * area;
  float inv_area = 1.0f / area;
  center.x *= inv_area;
  center.y *= inv_area;
  V2 mc = v2add(center, s0);
  float Io = 20.f * I;
  Io += m * (v2dot(mc, mc) - v2dot(center, center));
  /* ResetMassData */
  V2 lc = v2(m * mc.x, m * mc.y);
  float im = 1.0f / m;
  lc.x *= im;
  lc.y *= im;
  Io -= m * v2dot(lc, lc);
  *mass = m;
  *inv_mass = im;
  *inv_i = 1.0f / Io;
}

__device__ static void shape_aabb(const SV& s, int i, XF T, float* out) {
  V2 v[4], nr[4];
  box_verts(s, i, v, nr);
  V2 lo = xf_mul(T, v[0]), hi = lo;
  for (int k = 1; k < 4; ++k) {
    V2 w = xf_mul(T, v[k]);
    lo = v2(fminf(lo.x, w.x), fminf(lo.y, w.y));
    hi = v2(fmaxf(hi.x, w.x), fmaxf(hi.y, w.y));
  }
  out[0] = lo.x - B2_POLY_RADIUS; out[1] = lo.y - B2_POLY_RADIUS;
  out[2] = hi.x + B2_POLY_RADIUS; out[3] = hi.y + B2_POLY_RADIUS;
}
__device__ static int aabb_contains(const float* a, const float* b) { /* a.Contains(b) */
  return a[0] <= b[0] && a[1] <= b[1] && b[2] <= a[2] && b[3] <= a[3];
}
__device__ static int aabb_overlap(const float* a, const float* b) { /* b2TestOverlap(aabb, aabb) */
  float d1x = b[0] - a[2], d1y = b[1] - a[3], d2x = a[0] - b[2], d2y = a[1] - b[3];
  if (d1x > 0.0f || d1y > 0.0f) return 0;
  if (d2x > 0.0f || d2y > 0.0f) return 0;
  return 1;
}
/* b2Fixture::Synchronize(xf1, xf2) -> b2DynamicTree::MoveProxy */
__device__ static void proxy_sync(const SV& s, SC& c, int i, XF T1, XF T2) {
  float a1[4], a2[4], ab[4];
  shape_aabb(s, i, T1, a1);
  shape_aabb(s, i, T2, a2);
  ab[0] = fminf(a1[0], a2[0]); ab[1] = fminf(a1[1], a2[1]);
  ab[2] = fmaxf(a1[2], a2[2]); ab[3] = fmaxf(a1[3], a2[3]);
  V2 c1 = v2(0.5f * (a1[0] + a1[2]), 0.5f * (a1[1] + a1[3]));
  V2 c2 = v2(0.5f * (a2[0] + a2[2]), 0.5f * (a2[1] + a2[3]));
  V2 disp = v2sub(c2, c1);
  float fat[4] = {ab[0] - B2_AABB_EXT, ab[1] - B2_AABB_EXT, ab[2] + B2_AABB_EXT, ab[3] + B2_AABB_EXT};
  V2 d = v2mul(B2_AABB_MULT, disp);
  if (d.x < 0.0f) fat[0] += d.x; else fat[2] += d.x;
  if (d.y < 0.0f) fat[1] += d.y; else fat[3] += d.y;
  float* tree = c.fat + 4 * i;
  if (aabb_contains(tree, ab)) {
    const float r4 = 4.0f * B2_AABB_EXT;
    float huge[4] = {fat[0] - r4, fat[1] - r4, fat[2] + r4, fat[3] + r4};
    if (aabb_contains(huge, tree)) return;
  }
  tree[0] = fat[0]; tree[1] = fat[1]; tree[2] = fat[2]; tree[3] = fat[3];
  c.moved[i] = 1;
}

__device__ static int find_contact(const SC& c, int a, int b) {
  for (int k = 0; k < (*c.n_contacts); ++k)
    if (c.ct[k].a == a && c.ct[k].b == b) return k;
  return -1;
}
__device__ static void find_new_contacts(const SV& s, SC& c) {
  int n = c.n;
  for (int i = 0; i < n; ++i) {
    if (!c.moved[i]) continue;
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      if (c.moved[j] && j > i) continue; /* both moved: the pair is reported once (b2_broad_phase.cpp QueryCallback) */
      if (!aabb_overlap(c.fat + 4 * i, c.fat + 4 * j)) continue;
      int a = i < j ? i : j, b = i < j ? j : i;
      if (s.px(a) < -500000.f && s.px(b) < -500000.f) continue; /* parked pairs, see header */
      if (find_contact(c, a, b) >= 0) continue;
      if ((*c.n_contacts) == CS_MAX_CONTACTS) { ++(*c.overflow); continue; }
      Contact* k = &c.ct[(*c.n_contacts)++];
      *k = Contact{};
      k->a = a; k->b = b;
    }
  }
  for (int i = 0; i < n; ++i) c.moved[i] = 0;
}

/* ---- b2CollidePolygons ---------------------------------------------------------------------------------------- */
typedef struct { V2 v; uint8_t ia, ib, ta, tb; } ClipV; /* id.cf: indexA, indexB, typeA, typeB (0 = vertex, 1 = face) */
__device__ static uint32_t cf_key(uint8_t ia, uint8_t ib, uint8_t ta, uint8_t tb) {
  return (uint32_t)ia | ((uint32_t)ib << 8) | ((uint32_t)ta << 16) | ((uint32_t)tb << 24);
}
__device__ static float find_max_separation(int* edge, const V2* v1s, const V2* n1s, XF xf1, const V2* v2s, XF xf2) {
  XF T; /* b2MulT(xf2, xf1) */
  T.s = xf2.c * xf1.s - xf2.s * xf1.c;
  T.c = xf2.c * xf1.c + xf2.s * xf1.s;
  T.p = rot_mulT(xf2.s, xf2.c, v2sub(xf1.p, xf2.p));
  int best = 0;
  float max_sep = -3.402823466e+38f;
  for (int i = 0; i < 4; ++i) {
    V2 n = rot_mul(T.s, T.c, n1s[i]);
    V2 v1 = xf_mul(T, v1s[i]);
    float si = 3.402823466e+38f;
    for (int j = 0; j < 4; ++j) {
      float sij = v2dot(n, v2sub(v2s[j], v1));
      if (sij < si) si = sij;
    }
    if (si > max_sep) { max_sep = si; best = i; }
  }
  *edge = best;
  return max_sep;
}
__device__ static int clip_segment(ClipV* out, const ClipV* in, V2 normal, float offset, int vertex_index_a) {
  int count = 0;
  float d0 = v2dot(normal, in[0].v) - offset;
  float d1 = v2dot(normal, in[1].v) - offset;
  if (d0 <= 0.0f) out[count++] = in[0];
  if (d1 <= 0.0f) out[count++] = in[1];
  if (d0 * d1 < 0.0f) {
    float interp = d0 / (d0 - d1);
    V2 e = v2sub(in[1].v, in[0].v);
    out[count].v = v2add(in[0].v, v2mul(interp, e));
    out[count].ia = (uint8_t)vertex_index_a;
    out[count].ib = in[0].ib;
    out[count].ta = 0;
    out[count].tb = 1;
    ++count;
  }
  return count;
}
__device__ static void collide_polygons(Contact* m, const V2* vA, const V2* nA, XF xfA, const V2* vB, const V2* nB, XF xfB) {
  m->count = 0;
  const float total_radius = B2_POLY_RADIUS + B2_POLY_RADIUS;
  int edgeA = 0, edgeB = 0;
  float sepA = find_max_separation(&edgeA, vA, nA, xfA, vB, xfB);
  if (sepA > total_radius) return;
  float sepB = find_max_separation(&edgeB, vB, nB, xfB, vA, xfA);
  if (sepB > total_radius) return;
  const V2 *v1s, *n1s, *v2s, *n2s;
  XF xf1, xf2;
  int edge1, flip;
  const float k_tol = 0.1f * B2_LINEAR_SLOP;
  if (sepB > sepA + k_tol) { v1s = vB; n1s = nB; v2s = vA; n2s = nA; xf1 = xfB; xf2 = xfA; edge1 = edgeB; m->type = 2; flip = 1; }
  else { v1s = vA; n1s = nA; v2s = vB; n2s = nB; xf1 = xfA; xf2 = xfB; edge1 = edgeA; m->type = 1; flip = 0; }
  /* b2FindIncidentEdge */
  ClipV inc[2];
  {
    V2 normal1 = rot_mulT(xf2.s, xf2.c, rot_mul(xf1.s, xf1.c, n1s[edge1]));
    int index = 0;
    float min_dot = 3.402823466e+38f;
    for (int i = 0; i < 4; ++i) {
      float d = v2dot(normal1, n2s[i]);
      if (d < min_dot) { min_dot = d; index = i; }
    }
    int i1 = index, i2 = i1 + 1 < 4 ? i1 + 1 : 0;
    inc[0].v = xf_mul(xf2, v2s[i1]); inc[0].ia = (uint8_t)edge1; inc[0].ib = (uint8_t)i1; inc[0].ta = 1; inc[0].tb = 0;
    inc[1].v = xf_mul(xf2, v2s[i2]); inc[1].ia = (uint8_t)edge1; inc[1].ib = (uint8_t)i2; inc[1].ta = 1; inc[1].tb = 0;
  }
  int iv1 = edge1, iv2 = edge1 + 1 < 4 ? edge1 + 1 : 0;
  V2 v11 = v1s[iv1], v12 = v1s[iv2];
  V2 lt = v2sub(v12, v11);
  { /* b2Vec2::Normalize */
    float len = sqrtf(lt.x * lt.x + lt.y * lt.y);
    if (!(len < 1.19209290e-07f)) { float inv = 1.0f / len; lt.x *= inv; lt.y *= inv; }
  }
  V2 ln = v2cross_vs(lt, 1.0f);
  V2 plane = v2mul(0.5f, v2add(v11, v12));
  V2 tangent = rot_mul(xf1.s, xf1.c, lt);
  V2 normal = v2cross_vs(tangent, 1.0f);
  v11 = xf_mul(xf1, v11);
  v12 = xf_mul(xf1, v12);
  float front = v2dot(normal, v11);
  float side1 = -v2dot(tangent, v11) + total_radius;
  float side2 = v2dot(tangent, v12) + total_radius;
  ClipV cp1[2], cp2[2];
  int np = clip_segment(cp1, inc, v2(-tangent.x, -tangent.y), side1, iv1);
  if (np < 2) return;
  np = clip_segment(cp2, cp1, tangent, side2, iv2);
  if (np < 2) return;
  m->ln = ln;
  m->lp = plane;
  int pc = 0;
  for (int i = 0; i < 2; ++i) {
    float sep = v2dot(normal, cp2[i].v) - front;
    if (sep <= total_radius) {
      MPoint* p = &m->pt[pc];
      V2 l = xf_mulT(xf2, cp2[i].v);
      p->lx = l.x; p->ly = l.y;
      p->key = flip ? cf_key(cp2[i].ib, cp2[i].ia, cp2[i].tb, cp2[i].ta) : cf_key(cp2[i].ia, cp2[i].ib, cp2[i].ta, cp2[i].tb);
      ++pc;
    }
  }
  m->count = pc;
}

/* b2ContactManager::Collide + b2Contact::Update */
__device__ static void collide(SV& s, SC& c) {
  int k = 0;
  while (k < (*c.n_contacts)) {
    Contact* ct = &c.ct[k];
    int a = ct->a, b = ct->b;
    if (s.awake(a) == 0.f && s.awake(b) == 0.f) { ++k; continue; }
    if (!aabb_overlap(c.fat + 4 * a, c.fat + 4 * b)) { /* Destroy */
      if (ct->count > 0) { sv_set_awake(s, a, true); sv_set_awake(s, b, true); }
      for (int j = k + 1; j < (*c.n_contacts); ++j) c.ct[j - 1] = c.ct[j];
      --(*c.n_contacts);
      continue;
    }
    Contact old = *ct;
    V2 vA[4], nA[4], vB[4], nB[4];
    box_verts(s, a, vA, nA);
    box_verts(s, b, vB, nB);
    collide_polygons(ct, vA, nA, body_xf(s, a), vB, nB, body_xf(s, b));
    int touching = ct->count > 0;
    for (int i = 0; i < ct->count; ++i) {
      ct->pt[i].ni = 0.0f;
      ct->pt[i].ti = 0.0f;
      for (int j = 0; j < old.count; ++j)
        if (old.pt[j].key == ct->pt[i].key) { ct->pt[i].ni = old.pt[j].ni; ct->pt[i].ti = old.pt[j].ti; break; }
    }
    if (touching != old.touching) { sv_set_awake(s, a, true); sv_set_awake(s, b, true); }
    ct->touching = touching;
    ++k;
  }
}

/* ---- b2ContactSolver over one island --------------------------------------------------------------------------- */
typedef struct { V2 rA, rB; float ni, ti, nmass, tmass, vbias; } VCP;
typedef struct {
  VCP p[2];
  V2 normal;
  float nm[4], K[4]; /* b2Mat22 as ex.x, ex.y, ey.x, ey.y */
  int ia, ib, count, ci;
  float mA, mB, iA, iB, friction;
} VC;

struct CsScratch {  // working set of the one solving thread, kept in shared memory
  V2 pc_[CS_MAX_BODIES]; float pa[CS_MAX_BODIES];
  V2 vv[CS_MAX_BODIES];  float vw[CS_MAX_BODIES];
  VC vcs[CS_MAX_ISLAND_CONTACTS];
  uint8_t cflag[CS_MAX_CONTACTS];
  int icon[CS_MAX_CONTACTS];
};

__device__ static void solve_island(SV& s, SC& c, const int* bodies, int nb, const int* cidx, int nc, float h, float dt_ratio, CsScratch* sc) {
  V2* pc_ = sc->pc_; float* pa = sc->pa;  // shared memory of the block (only this thread uses it)
  V2* vv = sc->vv;   float* vw = sc->vw;
  VC* vcs = sc->vcs;
  if (nc > CS_MAX_ISLAND_CONTACTS) { (*c.overflow) += nc - CS_MAX_ISLAND_CONTACTS; nc = CS_MAX_ISLAND_CONTACTS; }  // see the header: larger islands drop their newest contacts
  for (int i = 0; i < nb; ++i) {
    int b = bodies[i];
    c.iidx[b] = i;
    c.c0x[b] = s.cx(b); c.c0y[b] = s.cy(b); c.a0[b] = s.ang(b);
    pc_[i] = v2(s.cx(b), s.cy(b)); pa[i] = s.ang(b);
    /* v += h * invMass * (gravityScale * mass * gravity + force) with zero gravity/force; damping factors are 1 */
    float hm = h * c.inv_mass[b];
    V2 f = v2add(v2mul(1.0f * c.mass[b], v2(0.0f, 0.0f)), v2(0.0f, 0.0f));
    V2 v = v2(s.vx(b) + hm * f.x, s.vy(b) + hm * f.y);
    float w = s.om(b) + h * c.inv_i[b] * 0.0f;
    v.x *= 1.0f / (1.0f + h * 0.0f); v.y *= 1.0f / (1.0f + h * 0.0f);
    w *= 1.0f / (1.0f + h * 0.0f);
    vv[i] = v; vw[i] = w;
  }
  /* constructor */
  for (int k = 0; k < nc; ++k) {
    Contact* ct = &c.ct[cidx[k]];
    VC* vc = &vcs[k];
    *vc = VC{};
    vc->friction = c.friction;
    vc->ia = c.iidx[ct->a]; vc->ib = c.iidx[ct->b];
    vc->mA = c.inv_mass[ct->a]; vc->mB = c.inv_mass[ct->b];
    vc->iA = c.inv_i[ct->a]; vc->iB = c.inv_i[ct->b];
    vc->ci = cidx[k];
    vc->count = ct->count;
    for (int j = 0; j < ct->count; ++j) {
      vc->p[j].ni = dt_ratio * ct->pt[j].ni;
      vc->p[j].ti = dt_ratio * ct->pt[j].ti;
    }
  }
  /* InitializeVelocityConstraints */
  for (int k = 0; k < nc; ++k) {
    VC* vc = &vcs[k];
    Contact* ct = &c.ct[vc->ci];
    float mA = vc->mA, mB = vc->mB, iA = vc->iA, iB = vc->iB;
    V2 lcA = v2(s.lcx(ct->a), s.lcy(ct->a)), lcB = v2(s.lcx(ct->b), s.lcy(ct->b));
    V2 cA = pc_[vc->ia], cB = pc_[vc->ib];
    float aA = pa[vc->ia], aB = pa[vc->ib];
    V2 vA = vv[vc->ia], vB = vv[vc->ib];
    float wA = vw[vc->ia], wB = vw[vc->ib];
    XF xfA, xfB;
    xfA.s = cr_sinf(aA); xfA.c = cr_cosf(aA); xfB.s = cr_sinf(aB); xfB.c = cr_cosf(aB);
    xfA.p = v2sub(cA, rot_mul(xfA.s, xfA.c, lcA));
    xfB.p = v2sub(cB, rot_mul(xfB.s, xfB.c, lcB));
    /* b2WorldManifold::Initialize */
    V2 wn, wp[2];
    const float rA_ = B2_POLY_RADIUS, rB_ = B2_POLY_RADIUS;
    if (ct->type == 1) {
      wn = rot_mul(xfA.s, xfA.c, ct->ln);
      V2 plane = xf_mul(xfA, ct->lp);
      for (int j = 0; j < ct->count; ++j) {
        V2 clip = xf_mul(xfB, v2(ct->pt[j].lx, ct->pt[j].ly));
        V2 ca = v2add(clip, v2mul(rA_ - v2dot(v2sub(clip, plane), wn), wn));
        V2 cb = v2sub(clip, v2mul(rB_, wn));
        wp[j] = v2mul(0.5f, v2add(ca, cb));
      }
    } else {
      wn = rot_mul(xfB.s, xfB.c, ct->ln);
      V2 plane = xf_mul(xfB, ct->lp);
      for (int j = 0; j < ct->count; ++j) {
        V2 clip = xf_mul(xfA, v2(ct->pt[j].lx, ct->pt[j].ly));
        V2 cb = v2add(clip, v2mul(rB_ - v2dot(v2sub(clip, plane), wn), wn));
        V2 ca = v2sub(clip, v2mul(rA_, wn));
        wp[j] = v2mul(0.5f, v2add(ca, cb));
      }
      wn = v2(-wn.x, -wn.y);
    }
    vc->normal = wn;
    for (int j = 0; j < vc->count; ++j) {
      VCP* p = &vc->p[j];
      p->rA = v2sub(wp[j], cA);
      p->rB = v2sub(wp[j], cB);
      float rnA = v2cross(p->rA, wn), rnB = v2cross(p->rB, wn);
      float kN = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
      p->nmass = kN > 0.0f ? 1.0f / kN : 0.0f;
      V2 tangent = v2cross_vs(wn, 1.0f);
      float rtA = v2cross(p->rA, tangent), rtB = v2cross(p->rB, tangent);
      float kT = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
      p->tmass = kT > 0.0f ? 1.0f / kT : 0.0f;
      p->vbias = 0.0f;
      V2 rel = v2sub(v2sub(v2add(vB, v2cross_sv(wB, p->rB)), vA), v2cross_sv(wA, p->rA));
      float vrel = v2dot(wn, rel);
      if (vrel < -1.0f) p->vbias = -0.0f * vrel; /* restitution 0, threshold 1 m/s */
    }
    if (vc->count == 2) {
      VCP *p1 = &vc->p[0], *p2 = &vc->p[1];
      float rn1A = v2cross(p1->rA, wn), rn1B = v2cross(p1->rB, wn);
      float rn2A = v2cross(p2->rA, wn), rn2B = v2cross(p2->rB, wn);
      float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
      float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
      float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
      if (k11 * k11 < 1000.0f * (k11 * k22 - k12 * k12)) {
        vc->K[0] = k11; vc->K[1] = k12; vc->K[2] = k12; vc->K[3] = k22;
        float a = k11, b = k12, cc = k12, d = k22; /* GetInverse: a = ex.x, b = ey.x, c = ex.y, d = ey.y */
        float det = a * d - b * cc;
        if (det != 0.0f) det = 1.0f / det;
        vc->nm[0] = det * d; vc->nm[2] = -det * b;
        vc->nm[1] = -det * cc; vc->nm[3] = det * a;
      } else {
        vc->count = 1;
      }
    }
  }
  /* WarmStart */
  for (int k = 0; k < nc; ++k) {
    VC* vc = &vcs[k];
    V2 vA = vv[vc->ia], vB = vv[vc->ib];
    float wA = vw[vc->ia], wB = vw[vc->ib];
    V2 normal = vc->normal, tangent = v2cross_vs(normal, 1.0f);
    for (int j = 0; j < vc->count; ++j) {
      VCP* p = &vc->p[j];
      V2 P = v2add(v2mul(p->ni, normal), v2mul(p->ti, tangent));
      wA -= vc->iA * v2cross(p->rA, P);
      vA = v2sub(vA, v2mul(vc->mA, P));
      wB += vc->iB * v2cross(p->rB, P);
      vB = v2add(vB, v2mul(vc->mB, P));
    }
    vv[vc->ia] = vA; vw[vc->ia] = wA; vv[vc->ib] = vB; vw[vc->ib] = wB;
  }
  /* 8 velocity iterations */
  for (int it = 0; it < 8; ++it) {
    for (int k = 0; k < nc; ++k) {
      VC* vc = &vcs[k];
      float mA = vc->mA, iA = vc->iA, mB = vc->mB, iB = vc->iB;
      V2 vA = vv[vc->ia], vB = vv[vc->ib];
      float wA = vw[vc->ia], wB = vw[vc->ib];
      V2 normal = vc->normal, tangent = v2cross_vs(normal, 1.0f);
      for (int j = 0; j < vc->count; ++j) {
        VCP* p = &vc->p[j];
        V2 dv = v2sub(v2sub(v2add(vB, v2cross_sv(wB, p->rB)), vA), v2cross_sv(wA, p->rA));
        float vt = v2dot(dv, tangent) - 0.0f;
        float lambda = p->tmass * (-vt);
        float maxf = vc->friction * p->ni;
        float ni_ = p->ti + lambda;
        float newi = ni_ < -maxf ? -maxf : (ni_ > maxf ? maxf : ni_); /* b2Clamp = b2Max(low, b2Min(a, high)) */
        newi = fmaxf(-maxf, fminf(ni_, maxf));
        lambda = newi - p->ti;
        p->ti = newi;
        V2 P = v2mul(lambda, tangent);
        vA = v2sub(vA, v2mul(mA, P));
        wA -= iA * v2cross(p->rA, P);
        vB = v2add(vB, v2mul(mB, P));
        wB += iB * v2cross(p->rB, P);
      }
      if (vc->count == 1) {
        VCP* p = &vc->p[0];
        V2 dv = v2sub(v2sub(v2add(vB, v2cross_sv(wB, p->rB)), vA), v2cross_sv(wA, p->rA));
        float vn = v2dot(dv, normal);
        float lambda = -p->nmass * (vn - p->vbias);
        float newi = fmaxf(p->ni + lambda, 0.0f);
        lambda = newi - p->ni;
        p->ni = newi;
        V2 P = v2mul(lambda, normal);
        vA = v2sub(vA, v2mul(mA, P));
        wA -= iA * v2cross(p->rA, P);
        vB = v2add(vB, v2mul(mB, P));
        wB += iB * v2cross(p->rB, P);
      } else {
        VCP *p1 = &vc->p[0], *p2 = &vc->p[1];
        V2 a = v2(p1->ni, p2->ni);
        V2 dv1 = v2sub(v2sub(v2add(vB, v2cross_sv(wB, p1->rB)), vA), v2cross_sv(wA, p1->rA));
        V2 dv2 = v2sub(v2sub(v2add(vB, v2cross_sv(wB, p2->rB)), vA), v2cross_sv(wA, p2->rA));
        float vn1 = v2dot(dv1, normal), vn2 = v2dot(dv2, normal);
        V2 b = v2(vn1 - p1->vbias, vn2 - p2->vbias);
        /* b -= K * a  (b2Mul(Mat22, Vec2): ex.x*v.x + ey.x*v.y, ex.y*v.x + ey.y*v.y) */
        V2 Ka = v2(vc->K[0] * a.x + vc->K[2] * a.y, vc->K[1] * a.x + vc->K[3] * a.y);
        b = v2sub(b, Ka);
        V2 x;
        int solved = 0;
        for (;;) {
          V2 nb_ = v2(vc->nm[0] * b.x + vc->nm[2] * b.y, vc->nm[1] * b.x + vc->nm[3] * b.y);
          x = v2(-nb_.x, -nb_.y);
          if (x.x >= 0.0f && x.y >= 0.0f) { solved = 1; break; }
          x.x = -p1->nmass * b.x; x.y = 0.0f;
          vn2 = vc->K[1] * x.x + b.y;
          if (x.x >= 0.0f && vn2 >= 0.0f) { solved = 1; break; }
          x.x = 0.0f; x.y = -p2->nmass * b.y;
          vn1 = vc->K[2] * x.y + b.x;
          if (x.y >= 0.0f && vn1 >= 0.0f) { solved = 1; break; }
          x.x = 0.0f; x.y = 0.0f;
          vn1 = b.x; vn2 = b.y;
          if (vn1 >= 0.0f && vn2 >= 0.0f) { solved = 1; break; }
          break;
        }
        if (solved) {
          V2 d = v2sub(x, a);
          V2 P1 = v2mul(d.x, normal), P2 = v2mul(d.y, normal);
          V2 P12 = v2add(P1, P2);
          vA = v2sub(vA, v2mul(mA, P12));
          wA -= iA * (v2cross(p1->rA, P1) + v2cross(p2->rA, P2));
          vB = v2add(vB, v2mul(mB, P12));
          wB += iB * (v2cross(p1->rB, P1) + v2cross(p2->rB, P2));
          p1->ni = x.x;
          p2->ni = x.y;
        }
      }
      vv[vc->ia] = vA; vw[vc->ia] = wA; vv[vc->ib] = vB; vw[vc->ib] = wB;
    }
  }
  /* StoreImpulses */
  for (int k = 0; k < nc; ++k) {
    VC* vc = &vcs[k];
    Contact* ct = &c.ct[vc->ci];
    for (int j = 0; j < vc->count; ++j) { ct->pt[j].ni = vc->p[j].ni; ct->pt[j].ti = vc->p[j].ti; }
  }
  /* integrate positions */
  for (int i = 0; i < nb; ++i) {
    V2 v = vv[i];
    float w = vw[i];
    V2 tr = v2mul(h, v);
    if (v2dot(tr, tr) > 5.0f * 5.0f) {
      float ratio = 5.0f / sqrtf(tr.x * tr.x + tr.y * tr.y);
      v.x *= ratio; v.y *= ratio;
    }
    float rot = h * w;
    if (rot * rot > (0.5f * B2_PI) * (0.5f * B2_PI)) {
      float ratio = (0.5f * B2_PI) / fabsf(rot);
      w *= ratio;
    }
    pc_[i].x += h * v.x; pc_[i].y += h * v.y;
    pa[i] += h * w;
    vv[i] = v; vw[i] = w;
  }
  /* 3 position iterations */
  int position_solved = 0;
  for (int it = 0; it < 3; ++it) {
    float min_sep = 0.0f;
    for (int k = 0; k < nc; ++k) {
      VC* vc = &vcs[k];
      Contact* ct = &c.ct[vc->ci];
      V2 lcA = v2(s.lcx(ct->a), s.lcy(ct->a)), lcB = v2(s.lcx(ct->b), s.lcy(ct->b));
      float mA = vc->mA, iA = vc->iA, mB = vc->mB, iB = vc->iB;
      V2 cA = pc_[vc->ia], cB = pc_[vc->ib];
      float aA = pa[vc->ia], aB = pa[vc->ib];
      for (int j = 0; j < ct->count; ++j) { /* pc->pointCount = manifold->pointCount (not reduced by the block-solver test) */
        XF xfA, xfB;
        xfA.s = cr_sinf(aA); xfA.c = cr_cosf(aA); xfB.s = cr_sinf(aB); xfB.c = cr_cosf(aB);
        xfA.p = v2sub(cA, rot_mul(xfA.s, xfA.c, lcA));
        xfB.p = v2sub(cB, rot_mul(xfB.s, xfB.c, lcB));
        V2 normal, point;
        float sep;
        if (ct->type == 1) {
          normal = rot_mul(xfA.s, xfA.c, ct->ln);
          V2 plane = xf_mul(xfA, ct->lp);
          V2 clip = xf_mul(xfB, v2(ct->pt[j].lx, ct->pt[j].ly));
          sep = v2dot(v2sub(clip, plane), normal) - B2_POLY_RADIUS - B2_POLY_RADIUS;
          point = clip;
        } else {
          normal = rot_mul(xfB.s, xfB.c, ct->ln);
          V2 plane = xf_mul(xfB, ct->lp);
          V2 clip = xf_mul(xfA, v2(ct->pt[j].lx, ct->pt[j].ly));
          sep = v2dot(v2sub(clip, plane), normal) - B2_POLY_RADIUS - B2_POLY_RADIUS;
          point = clip;
          normal = v2(-normal.x, -normal.y);
        }
        V2 rA = v2sub(point, cA), rB = v2sub(point, cB);
        min_sep = fminf(min_sep, sep);
        float C = fmaxf(-B2_MAX_LIN_CORR, fminf(B2_BAUMGARTE * (sep + B2_LINEAR_SLOP), 0.0f));
        float rnA = v2cross(rA, normal), rnB = v2cross(rB, normal);
        float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        float impulse = K > 0.0f ? -C / K : 0.0f;
        V2 P = v2mul(impulse, normal);
        cA = v2sub(cA, v2mul(mA, P));
        aA -= iA * v2cross(rA, P);
        cB = v2add(cB, v2mul(mB, P));
        aB += iB * v2cross(rB, P);
      }
      pc_[vc->ia] = cA; pa[vc->ia] = aA; pc_[vc->ib] = cB; pa[vc->ib] = aB;
    }
    if (min_sep >= -3.0f * B2_LINEAR_SLOP) { position_solved = 1; break; }
  }
  /* copy back + SynchronizeTransform */
  for (int i = 0; i < nb; ++i) {
    int b = bodies[i];
    s.cx(b) = pc_[i].x; s.cy(b) = pc_[i].y; s.ang(b) = pa[i];
    s.vx(b) = vv[i].x; s.vy(b) = vv[i].y; s.om(b) = vw[i];
    float qs = cr_sinf(s.ang(b)), qc = cr_cosf(s.ang(b));
    s.px(b) = s.cx(b) - (qc * s.lcx(b) - qs * s.lcy(b));
    s.py(b) = s.cy(b) - (qs * s.lcx(b) + qc * s.lcy(b));
  }
  /* sleep */
  float min_sleep = 3.402823466e+38f;
  for (int i = 0; i < nb; ++i) {
    int b = bodies[i];
    if (s.om(b) * s.om(b) > (2.0f / 180.0f * B2_PI) * (2.0f / 180.0f * B2_PI) || s.vx(b) * s.vx(b) + s.vy(b) * s.vy(b) > 0.01f * 0.01f) {
      s.sleep_t(b) = 0.0f;
      min_sleep = 0.0f;
    } else {
      s.sleep_t(b) += h;
      min_sleep = fminf(min_sleep, s.sleep_t(b));
    }
  }
  if (min_sleep >= 0.5f && position_solved)
    for (int i = 0; i < nb; ++i) sv_set_awake(s, bodies[i], false);
}

/* b2World::Step(dt, 8, 3) */
__device__ static void world_step(SV& s, SC& c, float dt, CsScratch* sc) {
  int n = s.n;
  if ((*c.new_contacts)) { find_new_contacts(s, c); (*c.new_contacts) = 0; }
  float dt_ratio = (*c.inv_dt0) * dt;
  collide(s, c);
  /* Solve: islands seeded from the body list (newest body first), depth-first over touching contacts; a body's
   * contact edges are visited newest contact first */
  for (int i = 0; i < n; ++i) c.island_flag[i] = 0;
  uint8_t* cflag = sc->cflag;
  int* icon = sc->icon;
  for (int k = 0; k < (*c.n_contacts); ++k) cflag[k] = 0;
  for (int seed = n - 1; seed >= 0; --seed) {
    if (c.island_flag[seed] || s.awake(seed) == 0.f) continue;
    int nb = 0, nc = 0, sp = 0;
    c.stack[sp++] = seed;
    c.island_flag[seed] = 1;
    while (sp > 0) {
      int b = c.stack[--sp];
      c.ibody[nb++] = b;
      s.awake(b) = 1.f; /* woken without resetting the sleep timer */
      for (int k = (*c.n_contacts) - 1; k >= 0; --k) {
        Contact* ct = &c.ct[k];
        if (ct->a != b && ct->b != b) continue;
        if (cflag[k] || !ct->touching) continue;
        icon[nc++] = k;
        cflag[k] = 1;
        int other = ct->a == b ? ct->b : ct->a;
        if (c.island_flag[other]) continue;
        c.stack[sp++] = other;
        c.island_flag[other] = 1;
      }
    }
    solve_island(s, c, c.ibody, nb, icon, nc, dt, dt_ratio, sc);
  }
  /* synchronise the fixtures of every body that was in an island (newest body first), then look for new pairs */
  for (int b = n - 1; b >= 0; --b) {
    if (!c.island_flag[b]) continue;
    XF T2 = body_xf(s, b);
    if (s.awake(b) != 0.f) {
      XF T1;
      T1.s = cr_sinf(c.a0[b]); T1.c = cr_cosf(c.a0[b]);
      T1.p = v2sub(v2(c.c0x[b], c.c0y[b]), rot_mul(T1.s, T1.c, v2(s.lcx(b), s.lcy(b))));
      proxy_sync(s, c, b, T1, T2);
    } else {
      proxy_sync(s, c, b, T2, T2);
    }
  }
  find_new_contacts(s, c);
  (*c.inv_dt0) = 1.0f / dt;
}


// persistent contact state of one scene inside CtrlSimBatch.cstate (float words; ints stored bit-wise):
//   [0] new_contacts  [1] n_contacts  [2] inv_dt0  [3] overflow count   | fat[4N] | moved[N] | mass, inv_mass, inv_i [3N] | contacts
__host__ __device__ inline int cs_words(int N) { return 4 + 8 * N + CS_CONTACT_WORDS * CS_MAX_CONTACTS; }
static_assert(sizeof(Contact) == CS_CONTACT_WORDS * 4, "Contact must stay 20 words");

__device__ inline SC cs_view(float* w, int N, int n, int* scratch) {
  SC c;
  c.n = n;
  c.new_contacts = reinterpret_cast<int*>(w + 0);
  c.n_contacts = reinterpret_cast<int*>(w + 1);
  c.inv_dt0 = w + 2;
  c.overflow = reinterpret_cast<int*>(w + 3);
  c.fat = w + 4;
  c.moved = reinterpret_cast<int*>(w + 4 + 4 * N);
  c.mass = w + 4 + 5 * N;
  c.inv_mass = w + 4 + 6 * N;
  c.inv_i = w + 4 + 7 * N;
  c.ct = reinterpret_cast<Contact*>(w + 4 + 8 * N);
  c.friction = sqrtf(0.2f * 0.2f);  // b2MixFriction of two default fixtures
  c.c0x = reinterpret_cast<float*>(scratch);
  c.c0y = c.c0x + CS_MAX_BODIES;
  c.a0 = c.c0y + CS_MAX_BODIES;
  c.island_flag = scratch + 3 * CS_MAX_BODIES;
  c.stack = c.island_flag + CS_MAX_BODIES;
  c.ibody = c.stack + CS_MAX_BODIES;
  c.iidx = c.ibody + CS_MAX_BODIES;
  return c;
}
constexpr int CS_SCRATCH_WORDS = 7 * CS_MAX_BODIES;

// Vehicle::CreatePhysicsBody: mass data and the proxy as CreateProxy / MoveProxy leave it once the body sits at its pose
__device__ static void cs_init_body(const SV& s, SC& c, int i) {
  box_mass(s.wid(i) / 2, s.len(i) / 2, &c.mass[i], &c.inv_mass[i], &c.inv_i[i]);
  float a[4];
  shape_aabb(s, i, body_xf(s, i), a);
  float* tree = c.fat + 4 * i;
  tree[0] = a[0] - B2_AABB_EXT; tree[1] = a[1] - B2_AABB_EXT; tree[2] = a[2] + B2_AABB_EXT; tree[3] = a[3] + B2_AABB_EXT;
  c.moved[i] = 1;
}
// b2Body::SetTransform (veh.setPosition): fixture synchronised with (m_xf, m_xf); the caller raises new_contacts
__device__ static void cs_teleport(const SV& s, SC& c, int i) {
  const XF T = body_xf(s, i);
  proxy_sync(s, c, i, T, T);
}

