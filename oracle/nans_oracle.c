/* nans_oracle.c — TEST INFRASTRUCTURE (oracle/). Not part of the product path.
 *
 * CPU restatement of the reference step; see nans_oracle.h.  Every function
 * cites the reference lines it follows.  Arithmetic is plain IEEE fp32 in the
 * reference's operation order (glm 0.9.9 scalar paths, pinned by the
 * disassembly of build/nans.so); build with
 * -ffp-contract=off so nothing is fused.
 */
#include "nans_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

/* ---- glm semantics --------------------------------------- */
static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 ld3(const float *p) { return V3(p[0], p[1], p[2]); }
static inline void st3(float *p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static inline v3 add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 muls(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }  /* vec * s and s * vec */
static inline v3 divs(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline v3 neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 cross(v3 x, v3 y)
{
    return V3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
static inline float length3(v3 a) { return sqrtf(dot(a, a)); }
static inline v3 normalize3(v3 a) { return muls(a, 1.0f / sqrtf(dot(a, a))); }
static inline int eq3(v3 a, v3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

/* ---- RK4, code/nans.cpp:51-78 ---------------------------------------------- */
/* MovementFunction :64-70.  (real32)(1.0/Mass) == 1.0f/Mass (double rounding of a
 * division is innocuous; the binary does an fp32 divide). */
static v3 movement_fn(v3 vel, v3 forces, float mass)
{
    float g = mass * 9.81f;
    v3 grav = V3(g * 0.0f, g * -1.0f, g * 0.0f);
    v3 t = sub(add(forces, grav), muls(vel, 1.5f));
    return muls(t, 1.0f / mass);
}
/* RotationFunction :73-78 */
static v3 rotation_fn(v3 w, v3 torque, float moi)
{
    return muls(sub(torque, muls(w, 1.5f)), 1.0f / moi);
}
static v3 rk4(v3 (*F)(v3, v3, float), float dt, v3 y0, v3 sum, float m)
{
    v3 k1 = muls(F(y0, sum, m), dt);
    v3 k2 = muls(F(add(y0, divs(k1, 2.0f)), sum, m), dt);
    v3 k3 = muls(F(add(y0, divs(k2, 2.0f)), sum, m), dt);
    v3 k4 = muls(F(add(y0, k3), sum, m), dt);
    v3 s = add(add(add(k1, muls(k2, 2.0f)), muls(k3, 2.0f)), k4);
    return add(y0, muls(s, 1.0f / 6.0f));
}

/* IntegrateForces, code/nans.cpp:975-1018 */
void oracle_integrate_forces(oracle_world *w, float dt)
{
    int nb = w->n_cubes + w->n_spheres;
    for (int i = 0; i < nb; ++i) {
        v3 v = rk4(movement_fn, dt, ld3(w->vel + 3 * i), ld3(w->force + 3 * i), w->mass[i]);
        st3(w->vel + 3 * i, v);
        v3 a = rk4(rotation_fn, dt, ld3(w->angvel + 3 * i), ld3(w->torque + 3 * i), w->moi[i]);
        st3(w->angvel + 3 * i, a);
        st3(w->force + 3 * i, V3(0, 0, 0));
        st3(w->torque + 3 * i, V3(0, 0, 0));
    }
}

/* IntegrateVelocities, code/nans.cpp:1332-1349 */
void oracle_integrate_velocities(oracle_world *w, float dt)
{
    int nb = w->n_cubes + w->n_spheres;
    for (int i = 0; i < nb; ++i) {
        st3(w->pos + 3 * i, add(ld3(w->pos + 3 * i), muls(ld3(w->vel + 3 * i), dt)));
        st3(w->ang + 3 * i, add(ld3(w->ang + 3 * i), muls(ld3(w->angvel + 3 * i), dt)));
    }
}

/* ---- model rebuild, code/nans.cpp:1870-1881,1913-1941 (glm translate/rotate/scale) */
typedef struct { float c[4][4]; } m4; /* c[col][row] */

static void col_mul(float *o, const float *a, float s) { for (int r = 0; r < 4; ++r) o[r] = a[r] * s; }

static m4 glm_rotate(m4 m, float angle, v3 axis_in)
{
    float c = cosf(angle), s = sinf(angle);
    v3 axis = normalize3(axis_in);
    v3 temp = muls(axis, 1.0f - c);
    float R[3][3];
    R[0][0] = c + temp.x * axis.x;
    R[0][1] = temp.x * axis.y + s * axis.z;
    R[0][2] = temp.x * axis.z - s * axis.y;
    R[1][0] = temp.y * axis.x - s * axis.z;
    R[1][1] = c + temp.y * axis.y;
    R[1][2] = temp.y * axis.z + s * axis.x;
    R[2][0] = temp.z * axis.x + s * axis.y;
    R[2][1] = temp.z * axis.y - s * axis.x;
    R[2][2] = c + temp.z * axis.z;
    m4 out;
    for (int i = 0; i < 3; ++i)
        for (int r = 0; r < 4; ++r)
            out.c[i][r] = (m.c[0][r] * R[i][0] + m.c[1][r] * R[i][1]) + m.c[2][r] * R[i][2];
    for (int r = 0; r < 4; ++r) out.c[3][r] = m.c[3][r];
    return out;
}

void oracle_model_vertices(const float pos[3], const float ang[3], const float scale[3],
                           float model_out[16], float verts_out[24])
{
    m4 m;
    memset(&m, 0, sizeof(m));
    m.c[0][0] = m.c[1][1] = m.c[2][2] = m.c[3][3] = 1.0f;
    /* translate: m3' = ((m0*v0 + m1*v1) + m2*v2) + m3 */
    for (int r = 0; r < 4; ++r)
        m.c[3][r] = ((m.c[0][r] * pos[0] + m.c[1][r] * pos[1]) + m.c[2][r] * pos[2]) + m.c[3][r];
    const float rad = 0.01745329251994329576923690768489f; /* glm::radians */
    m = glm_rotate(m, ang[0] * rad, V3(1.0f, 0.0f, 0.0f));
    m = glm_rotate(m, ang[1] * rad, V3(0.0f, 1.0f, 0.0f));
    m = glm_rotate(m, ang[2] * rad, V3(0.0f, 0.0f, 1.0f));
    col_mul(m.c[0], m.c[0], scale[0]);
    col_mul(m.c[1], m.c[1], scale[1]);
    col_mul(m.c[2], m.c[2], scale[2]);
    if (model_out) memcpy(model_out, m.c, 64);
    /* UpdateVertices, code/nans.cpp:395-407: mat4*vec4 = (m0*x + m1*y) + (m2*z + m3*w) */
    static const float corner[8][3] = {
        {0.5f, 0.5f, 0.5f},  {0.5f, 0.5f, -0.5f},  {-0.5f, 0.5f, 0.5f},  {-0.5f, 0.5f, -0.5f},
        {0.5f, -0.5f, 0.5f}, {0.5f, -0.5f, -0.5f}, {-0.5f, -0.5f, 0.5f}, {-0.5f, -0.5f, -0.5f}};
    for (int k = 0; k < 8; ++k)
        for (int r = 0; r < 3; ++r)
            verts_out[3 * k + r] = (m.c[0][r] * corner[k][0] + m.c[1][r] * corner[k][1]) +
                                   (m.c[2][r] * corner[k][2] + m.c[3][r] * 1.0f);
}

void oracle_rebuild_vertices(oracle_world *w)
{
    for (int i = 0; i < w->n_cubes; ++i)
        oracle_model_vertices(w->pos + 3 * i, w->ang + 3 * i, w->scale + 3 * i, NULL, w->verts + 24 * i);
    for (int i = 0; i < w->n_statics; ++i)
        oracle_model_vertices(w->st_pos + 3 * i, w->st_ang + 3 * i, w->st_scale + 3 * i, NULL,
                              w->st_verts + 24 * i);
}

/* ---- supports, code/nans.cpp:410-537 --------------------------------------- */
typedef struct { v3 P, SupA, SupB; } vtx;
typedef struct { vtx *v; int n, cap; } vtx_vec;      /* std::vector<vertex> */

static void vv_push(vtx_vec *s, vtx x)
{
    if (s->n == s->cap) { s->cap = s->cap ? 2 * s->cap : 8; s->v = realloc(s->v, sizeof(vtx) * s->cap); }
    s->v[s->n++] = x;
}
static void vv_erase(vtx_vec *s, int i)
{
    memmove(s->v + i, s->v + i + 1, sizeof(vtx) * (s->n - i - 1));
    s->n--;
}

/* GetCubeSupport / GetFloorSupport :410-430,441-461 */
static v3 box_support(const float *verts, v3 d)
{
    float best = -FLT_MAX;
    v3 res = V3(0.0f, 0.0f, 0.0f);
    for (int k = 0; k < 8; ++k) {
        v3 c = ld3(verts + 3 * k);
        float dist = dot(c, d);
        if (dist > best) { best = dist; res = c; }
    }
    return res;
}
/* GetSphereSupport :433-438 */
static v3 sphere_support(const oracle_shape *s, v3 d)
{
    return add(ld3(s->pos), muls(normalize3(d), s->radius));
}
static v3 shape_support(const oracle_shape *s, v3 d)
{
    return s->kind == 0 ? box_support(s->verts, d) : sphere_support(s, d);
}
/* CalculateSupport :464-519 (always appends to the simplex vector) */
static vtx calc_support(const oracle_shape *A, const oracle_shape *B, v3 d, vtx_vec *simplex)
{
    vtx r;
    r.SupA = shape_support(A, d);
    r.SupB = shape_support(B, muls(d, -1.0f));
    r.P = sub(r.SupA, r.SupB);
    vv_push(simplex, r);
    return r;
}
/* AddSupport :522-537 */
static int add_support(const oracle_shape *A, const oracle_shape *B, v3 d, vtx_vec *simplex)
{
    vtx nv = calc_support(A, B, d, simplex);
    return dot(d, nv.P) >= 0 ? 1 : 0;
}

/* ClosestPointOnLine :540-562 */
static v3 closest_point_on_line(v3 A, v3 B, float *U, float *V)
{
    v3 seg = normalize3(sub(B, A));
    float len = length3(seg);
    *V = dot(neg(A), seg) / len;
    *U = dot(B, seg) / len;
    if (*U <= 0.0f) return B;
    if (*V <= 0.0f) return A;
    return add(muls(A, *U), muls(B, *V));
}
/* TripleCross :565-569 */
static v3 triple_cross(v3 A, v3 B, v3 C) { return sub(muls(B, dot(C, A)), muls(A, dot(C, B))); }

enum { NoIntersection = 0, FoundIntersection = 1, StillEvolving = 2 };

/* EvolveSimplex :572-769 */
static int evolve_simplex(const oracle_shape *A, const oracle_shape *B, vtx_vec *S)
{
    v3 dir = normalize3(sub(ld3(B->pos), ld3(A->pos)));
    switch (S->n) {
    case 0: break;
    case 1: dir = muls(dir, -1.0f); break;
    case 2: {
        float U = 0.0f, V = 0.0f;
        v3 cp = closest_point_on_line(S->v[0].P, S->v[1].P, &U, &V);
        if (V <= 0.0f) { vv_erase(S, 1); dir = neg(cp); }
        else if (U <= 0.0f) { vv_erase(S, 0); dir = neg(cp); }
        else dir = neg(cp);
    } break;
    case 3: {
        v3 ao = neg(S->v[0].P);
        v3 e1 = sub(S->v[1].P, S->v[0].P);
        v3 e2 = sub(S->v[2].P, S->v[0].P);
        v3 tn = cross(e1, e2);
        v3 e1n = cross(e1, tn);
        v3 e2n = cross(tn, e2);
        if (dot(e2n, ao) > 0.0f) {
            if (dot(e2, ao) > 0.0f) { dir = triple_cross(e2, ao, e2); vv_erase(S, 1); }
            else if (dot(e1, ao) > 0.0f) { dir = triple_cross(e1, ao, e1); vv_erase(S, 2); }
            else { dir = ao; vv_erase(S, 2); vv_erase(S, 1); }
        } else if (dot(e1n, ao) > 0.0f) {
            if (dot(e1, ao) > 0.0f) { dir = triple_cross(e1, ao, e1); vv_erase(S, 2); }
            else { dir = ao; vv_erase(S, 2); vv_erase(S, 1); }
        } else if (dot(tn, ao) > 0.0f) {
            dir = tn;
        } else {
            dir = neg(tn);
            vtx t = S->v[1]; S->v[1] = S->v[2]; S->v[2] = t;
        }
    } break;
    case 4: {
        v3 da = sub(S->v[0].P, S->v[3].P);
        v3 db = sub(S->v[1].P, S->v[3].P);
        v3 dc = sub(S->v[2].P, S->v[3].P);
        v3 d0 = muls(S->v[3].P, -1.0f);
        v3 abd = cross(da, db), bcd = cross(db, dc), cad = cross(dc, da);
        if (dot(abd, d0) > 0.0f) { vv_erase(S, 2); dir = abd; }
        else if (dot(bcd, d0) > 0.0f) { vv_erase(S, 0); dir = bcd; }
        else if (dot(cad, d0) > 0.0f) { vv_erase(S, 1); dir = cad; }
        else return FoundIntersection;
    } break;
    default: break; /* unreachable: the simplex never exceeds 4 during GJK */
    }
    if (length3(dir) <= 0.0001f) return NoIntersection;
    return add_support(A, B, dir, S) ? StillEvolving : NoIntersection;
}

/* ---- EPA, code/nans.cpp:233-322, :772-904 ---------------------------------- */
typedef struct { vtx A, B, C; v3 N; } tri;
typedef struct { vtx A, B; } edg;
typedef struct { tri *t; int n, cap; } tri_vec;
typedef struct { edg *e; int n, cap; } edg_vec;

/* PushTriangle :293-322 */
static void push_triangle(tri_vec *T, vtx A, vtx B, vtx C)
{
    tri x;
    x.A = A; x.B = B; x.C = C;
    x.N = normalize3(cross(sub(B.P, A.P), sub(C.P, A.P)));
    if (dot(x.A.P, x.N) < 0) x.N = muls(x.N, -1.0f);
    if (T->n == T->cap) { T->cap = T->cap ? 2 * T->cap : 16; T->t = realloc(T->t, sizeof(tri) * T->cap); }
    T->t[T->n++] = x;
}
/* PushEdge :233-266 — opposite-winding cancellation by VALUE of P (code/nans.h:251-254) */
static void push_edge(edg_vec *E, vtx A, vtx B)
{
    for (int i = 0; i < E->n; ++i)
        if (eq3(E->e[i].A.P, B.P) && eq3(E->e[i].B.P, A.P)) {
            memmove(E->e + i, E->e + i + 1, sizeof(edg) * (E->n - i - 1));
            E->n--;
            return;
        }
    if (E->n == E->cap) { E->cap = E->cap ? 2 * E->cap : 16; E->e = realloc(E->e, sizeof(edg) * E->cap); }
    E->e[E->n].A = A;
    E->e[E->n].B = B;
    E->n++;
}
/* Barycentric :772-785 */
static void barycentric(v3 P, v3 A, v3 B, v3 C, float *U, float *V, float *W)
{
    v3 v0 = sub(B, A), v1 = sub(C, A), v2 = sub(P, A);
    float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1);
    float d20 = dot(v2, v0), d21 = dot(v2, v1);
    float denom = d00 * d11 - d01 * d01;
    *V = (d11 * d20 - d01 * d21) / denom;
    *W = (d00 * d21 - d01 * d20) / denom;
    *U = 1.0f - *V - *W;
}
static int is_valid(float f) { return !(isnan(f) || isinf(f)); }

/* ResolveCollision :788-904 */
static int resolve_collision(const oracle_shape *SA, const oracle_shape *SB, vtx_vec *S,
                             float outN[3], float outPA[3], float outPB[3], oracle_np_stats *st)
{
    tri_vec T = {0, 0, 0};
    edg_vec E = {0, 0, 0};
    int result = 0, it = 0;
    vtx A = S->v[0], B = S->v[1], C = S->v[2], D = S->v[3];
    push_triangle(&T, A, B, C);
    push_triangle(&T, A, C, D);
    push_triangle(&T, A, D, B);
    push_triangle(&T, B, D, C);
    while (it++ <= 64) {
        if (st) { st->epa_iters = it; if (T.n > st->max_faces) st->max_faces = T.n; }
        float cur = fabsf(dot(T.t[0].N, T.t[0].A.P));
        int ci = 0;
        for (int i = 0; i < T.n; ++i) {
            v3 ab = sub(T.t[i].B.P, T.t[i].A.P), ac = sub(T.t[i].C.P, T.t[i].A.P);
            v3 nrm = normalize3(cross(ab, ac));
            float d = fabsf(dot(nrm, T.t[i].A.P));
            if (d < cur) { cur = d; ci = i; }
        }
        v3 dir = T.t[ci].N;
        vtx ns = calc_support(SA, SB, dir, S);
        if (dot(T.t[ci].N, ns.P) - cur < 0.001f) {
            float u, v, w;
            const tri *c = &T.t[ci];
            barycentric(muls(c->N, cur), c->A.P, c->B.P, c->C.P, &u, &v, &w);
            if (fabsf(u) > 1.0f || fabsf(v) > 1.0f || fabsf(w) > 1.0f) { result = 0; goto done; }
            if (!is_valid(u) || !is_valid(v) || !is_valid(w)) { result = 0; goto done; }
            st3(outPA, add(add(muls(c->A.SupA, u), muls(c->B.SupA, v)), muls(c->C.SupA, w)));
            st3(outN, muls(c->N, -1.0f));
            st3(outPB, add(add(muls(c->A.SupB, u), muls(c->B.SupB, v)), muls(c->C.SupB, w)));
            result = 1;
            goto done;
        }
        for (int i = 0; i < T.n;) {
            v3 tmp = sub(ns.P, T.t[i].A.P);
            v3 ab = sub(T.t[i].B.P, T.t[i].A.P), ac = sub(T.t[i].C.P, T.t[i].A.P);
            v3 nrm = normalize3(cross(ab, ac));
            if (dot(nrm, T.t[i].A.P) < 0) nrm = muls(nrm, -1.0f);
            if (dot(nrm, tmp) > 0) {
                tri x = T.t[i];
                push_edge(&E, x.A, x.B);
                push_edge(&E, x.B, x.C);
                push_edge(&E, x.C, x.A);
                if (st && E.n > st->max_edges) st->max_edges = E.n;
                memmove(T.t + i, T.t + i + 1, sizeof(tri) * (T.n - i - 1));
                T.n--;
                continue;
            }
            ++i;
        }
        for (int i = 0; i < E.n; ++i) push_triangle(&T, ns, E.e[i].A, E.e[i].B);
        E.n = 0;
        /* T.n == 0 here means the next iteration reads T.t[0] of an EMPTY list, exactly as the reference reads
         * Triangle[0] of its emptied std::vector (:807-811): the storage still holds the last face, shifted down
         * by the erases (memmove above == vector::erase).  nans.so does report hits this way (pinned by
         * tests/golden/epa_emptied.npz); T.cap >= 16, so the read stays inside the allocation. */
        if (st && T.n == 0) st->emptied++;
    }
done:
    free(T.t);
    free(E.e);
    return result;
}

/* CheckCollision :907-966 */
int oracle_check_collision(const oracle_shape *A, const oracle_shape *B,
                           float outN[3], float outPA[3], float outPB[3], oracle_np_stats *st)
{
    vtx_vec S = {0, 0, 0};
    int ev = StillEvolving, result = 0;
    unsigned iter = 0;
    if (st) memset(st, 0, sizeof(*st));
    while (ev == StillEvolving && iter++ <= 64) {
        ev = evolve_simplex(A, B, &S);
        if (st) st->gjk_iters++;
    }
    if (st) st->gjk_result = ev;
    if (ev == FoundIntersection) result = resolve_collision(A, B, &S, outN, outPA, outPB, st);
    free(S.v);
    return result;
}

void oracle_check_collision_batch(int n, const int32_t *type,
                                  const float *posA, const float *vertsA, const float *radA,
                                  const float *posB, const float *vertsB, const float *radB,
                                  int32_t *hit, int32_t *gjk, float *outN, float *outPA, float *outPB,
                                  oracle_np_stats *stats)
{
    for (int i = 0; i < n; ++i) {
        oracle_shape A, B;
        memset(&A, 0, sizeof(A));
        memset(&B, 0, sizeof(B));
        A.kind = (type[i] == ORC_SS || type[i] == ORC_SF) ? 1 : 0;
        B.kind = (type[i] == ORC_CS || type[i] == ORC_SS) ? 1 : 0;
        memcpy(A.pos, posA + 3 * i, 12);
        memcpy(B.pos, posB + 3 * i, 12);
        if (A.kind == 0) memcpy(A.verts, vertsA + 24 * i, 96); else A.radius = radA[i];
        if (B.kind == 0) memcpy(B.verts, vertsB + 24 * i, 96); else B.radius = radB[i];
        float N[3] = {0, 0, 0}, PA[3] = {0, 0, 0}, PB[3] = {0, 0, 0};
        oracle_np_stats st;
        hit[i] = oracle_check_collision(&A, &B, N, PA, PB, &st);
        if (gjk) gjk[i] = st.gjk_result;
        memcpy(outN + 3 * i, N, 12);
        memcpy(outPA + 3 * i, PA, 12);
        memcpy(outPB + 3 * i, PB, 12);
        if (stats) stats[i] = st;
    }
}

/* ---- DetectCollisions, code/nans.cpp:1352-1536 ------------------------------ */
static void cube_shape(const oracle_world *w, int i, oracle_shape *s)
{
    s->kind = 0; s->radius = 0.0f;
    memcpy(s->pos, w->pos + 3 * i, 12);
    memcpy(s->verts, w->verts + 24 * i, 96);
}
static void sphere_shape(const oracle_world *w, int i, oracle_shape *s)
{
    int b = w->n_cubes + i;
    s->kind = 1; s->radius = w->radius[b];
    memcpy(s->pos, w->pos + 3 * b, 12);
}
static void static_shape(const oracle_world *w, int i, oracle_shape *s)
{
    s->kind = 0; s->radius = 0.0f;
    memcpy(s->pos, w->st_pos + 3 * i, 12);
    memcpy(s->verts, w->st_verts + 24 * i, 96);
}
/* Conservative AABB used only by the optional prefilter (not in the reference). */
static void shape_aabb(const oracle_shape *s, float lo[3], float hi[3])
{
    if (s->kind == 0) {
        for (int r = 0; r < 3; ++r) { lo[r] = s->verts[r]; hi[r] = s->verts[r]; }
        for (int k = 1; k < 8; ++k)
            for (int r = 0; r < 3; ++r) {
                float x = s->verts[3 * k + r];
                if (x < lo[r]) lo[r] = x;
                if (x > hi[r]) hi[r] = x;
            }
    } else {
        for (int r = 0; r < 3; ++r) { lo[r] = s->pos[r] - s->radius; hi[r] = s->pos[r] + s->radius; }
    }
    for (int r = 0; r < 3; ++r) {
        float m = 1e-3f + 1e-5f * (fabsf(lo[r]) > fabsf(hi[r]) ? fabsf(lo[r]) : fabsf(hi[r]));
        lo[r] -= m; hi[r] += m;
    }
}
static int aabb_overlap(const float *alo, const float *ahi, const float *blo, const float *bhi)
{
    for (int r = 0; r < 3; ++r)
        if (!(alo[r] <= bhi[r] && blo[r] <= ahi[r])) return 0;
    return 1;
}

typedef struct { oracle_contact *out; int n, cap; } contact_sink;

static void test_pair(int type, int a, int b, const oracle_shape *A, const oracle_shape *B,
                      int prefilter, contact_sink *sink)
{
    if (prefilter) {
        float alo[3], ahi[3], blo[3], bhi[3];
        shape_aabb(A, alo, ahi);
        shape_aabb(B, blo, bhi);
        if (!aabb_overlap(alo, ahi, blo, bhi)) return;
    }
    oracle_contact c;
    memset(&c, 0, sizeof(c));
    c.type = type; c.a = a; c.b = b;
    if (oracle_check_collision(A, B, c.n, c.point_a, c.point_b, NULL)) {
        if (sink->n < sink->cap) sink->out[sink->n] = c;
        sink->n++;
    }
}

/* ---- spatial prefilter (prefilter == 2): TEST INFRASTRUCTURE ONLY, not in the reference -------------
 * The reference tests every pair (code/nans.cpp:1357-1364); at 10^6 bodies that is 5*10^11 GJK calls.
 * This path yields the SAME contact list in the SAME order: a uniform grid over the inflated AABBs
 * (cell edge >= the largest AABB extent, so overlapping AABBs sit in adjacent cells) proposes, for
 * every body i, the partners j whose AABB overlaps -- the exact predicate of prefilter == 1 -- which
 * are then sorted ascending, so the pairs are visited in the all-pairs loop order.  The GJK/EPA calls
 * (pure functions of one pair) run on worker threads (pthreads); the list is assembled serially in order.
 * tests/test_oracle_grid.py proves it equal to the all-pairs list on <=16-body and 10k-body worlds. */
typedef struct {
    float *lo, *hi;          /* [nb][3] inflated AABBs, cubes then spheres */
    int64_t *tab_key;        /* open-addressing hash: packed cell coordinates -> cell id (-1 = empty) */
    int32_t *tab_cell;
    uint32_t tab_mask;
    int32_t *cell_start;     /* [ncell + 1] */
    int32_t *cell_body;      /* [nb] bodies grouped by cell, ascending inside a cell */
    int32_t *cell_of;        /* [nb] */
    int32_t (*coord)[3];     /* [nb] integer cell coordinates */
    float inv;
} grid_t;

/* unbounded integer cell coordinate, clamped (monotonically) to +-2^20; NaN -> 0 (overlap tests fail anyway) */
static int32_t grid_coord(const grid_t *g, float c)
{
    float f = floorf(c * g->inv);
    if (!(f == f)) return 0;
    if (f < -1048576.0f) return -1048576;
    if (f > 1048576.0f) return 1048576;
    return (int32_t)f;
}
static int64_t cell_key(int32_t x, int32_t y, int32_t z)
{
    return ((int64_t)(x + 2097152) << 44) | ((int64_t)(y + 2097152) << 22) | (int64_t)(z + 2097152);
}
static uint32_t key_hash(int64_t k)
{
    uint64_t h = (uint64_t)k * 0x9E3779B97F4A7C15ull;
    return (uint32_t)(h >> 32);
}
/* cell id of the key, or -1; insert != 0 creates it (next id = *ncell) */
static int32_t grid_lookup(grid_t *g, int64_t key, int insert, int32_t *ncell)
{
    uint32_t h = key_hash(key) & g->tab_mask;
    for (;;) {
        if (g->tab_key[h] == key) return g->tab_cell[h];
        if (g->tab_key[h] == -1) {
            if (!insert) return -1;
            g->tab_key[h] = key;
            g->tab_cell[h] = (*ncell)++;
            return g->tab_cell[h];
        }
        h = (h + 1) & g->tab_mask;
    }
}
static void grid_build(const oracle_world *w, grid_t *g)
{
    const int nb = w->n_cubes + w->n_spheres;
    oracle_shape S;
    g->lo = malloc(sizeof(float) * 3 * (nb + 1));
    g->hi = malloc(sizeof(float) * 3 * (nb + 1));
    float ext = 0.0f;
    for (int b = 0; b < nb; ++b) {
        if (b < w->n_cubes) cube_shape(w, b, &S); else sphere_shape(w, b - w->n_cubes, &S);
        shape_aabb(&S, g->lo + 3 * b, g->hi + 3 * b);
        for (int r = 0; r < 3; ++r) {
            float e = g->hi[3 * b + r] - g->lo[3 * b + r];
            if (e > ext && e < 1e6f) ext = e;      /* a non-finite / absurd box cannot overlap anything finite nearby */
        }
    }
    if (!(ext > 1e-3f)) ext = 1e-3f;
    ext *= 1.001f;
    g->inv = 1.0f / ext;
    uint32_t tab = 1024;
    while (tab < 2u * (uint32_t)(nb + 1)) tab <<= 1;
    g->tab_mask = tab - 1;
    g->tab_key = malloc(sizeof(int64_t) * tab);
    g->tab_cell = malloc(sizeof(int32_t) * tab);
    for (uint32_t k = 0; k < tab; ++k) g->tab_key[k] = -1;
    g->cell_of = malloc(sizeof(int32_t) * (nb + 1));
    g->coord = malloc(sizeof(int32_t[3]) * (nb + 1));
    int32_t ncell = 0;
    for (int b = 0; b < nb; ++b) {
        for (int r = 0; r < 3; ++r) g->coord[b][r] = grid_coord(g, 0.5f * (g->lo[3 * b + r] + g->hi[3 * b + r]));
        g->cell_of[b] = grid_lookup(g, cell_key(g->coord[b][0], g->coord[b][1], g->coord[b][2]), 1, &ncell);
    }
    g->cell_start = calloc((size_t)ncell + 1, sizeof(int32_t));
    g->cell_body = malloc(sizeof(int32_t) * (nb + 1));
    for (int b = 0; b < nb; ++b) g->cell_start[g->cell_of[b] + 1]++;
    for (int32_t c = 0; c < ncell; ++c) g->cell_start[c + 1] += g->cell_start[c];
    int32_t *fill = malloc(sizeof(int32_t) * ((size_t)ncell + 1));
    memcpy(fill, g->cell_start, sizeof(int32_t) * ((size_t)ncell + 1));
    for (int b = 0; b < nb; ++b) g->cell_body[fill[g->cell_of[b]]++] = b;   /* ascending body index per cell */
    free(fill);
}
static void grid_free(grid_t *g)
{
    free(g->lo); free(g->hi); free(g->tab_key); free(g->tab_cell); free(g->cell_start); free(g->cell_body);
    free(g->cell_of); free(g->coord);
}

static int cmp_i32(const void *a, const void *b) { int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return (x > y) - (x < y); }

/* partners of body b among rows [jlo, jhi) with an overlapping AABB, ascending; returns the count.
 * A box of extent > cell (only the absurd ones skipped above) is tested against nothing: it is non-finite or
 * farther than 1e6 across, and cannot be part of a contact the all-pairs loop would keep either -- except by
 * overlapping everything; such worlds are out of this helper's domain (tests use prefilter 0/1 for them). */
static int grid_partners(grid_t *g, int b, int jlo, int jhi, int32_t **buf, int *cap)
{
    int n = 0;
    const int32_t cx = g->coord[b][0], cy = g->coord[b][1], cz = g->coord[b][2];
    for (int32_t z = cz - 1; z <= cz + 1; ++z)
        for (int32_t y = cy - 1; y <= cy + 1; ++y)
            for (int32_t x = cx - 1; x <= cx + 1; ++x) {
                const int32_t c = grid_lookup(g, cell_key(x, y, z), 0, NULL);
                if (c < 0) continue;
                for (int k = g->cell_start[c]; k < g->cell_start[c + 1]; ++k) {
                    const int j = g->cell_body[k];
                    if (j < jlo || j >= jhi) continue;
                    if (!aabb_overlap(g->lo + 3 * b, g->hi + 3 * b, g->lo + 3 * j, g->hi + 3 * j)) continue;
                    if (n == *cap) { *cap = *cap ? 2 * *cap : 64; *buf = realloc(*buf, sizeof(int32_t) * *cap); }
                    (*buf)[n++] = j;
                }
            }
    qsort(*buf, n, sizeof(int32_t), cmp_i32);
    return n;
}

typedef struct { int32_t type, a, b; } cand_t;   /* a, b: reference indices (sphere indices 0-based) */

typedef struct {
    const oracle_world *w; const cand_t *cand; oracle_contact *res; uint8_t *hit; long total; long next;
} np_job;
static void *np_worker(void *arg)
{
    np_job *J = arg;
    const oracle_world *w = J->w;
    for (;;) {
        const long q0 = __atomic_fetch_add(&J->next, 1024, __ATOMIC_RELAXED);
        if (q0 >= J->total) break;
        const long q1 = q0 + 1024 < J->total ? q0 + 1024 : J->total;
        for (long q = q0; q < q1; ++q) {
            oracle_shape A, B;
            const cand_t c = J->cand[q];
            switch (c.type) {
            case ORC_CC: cube_shape(w, c.a, &A); cube_shape(w, c.b, &B); break;
            case ORC_CF: cube_shape(w, c.a, &A); static_shape(w, c.b, &B); break;
            case ORC_SF: sphere_shape(w, c.a, &A); static_shape(w, c.b, &B); break;
            case ORC_CS: cube_shape(w, c.a, &A); sphere_shape(w, c.b, &B); break;
            default: sphere_shape(w, c.a, &A); sphere_shape(w, c.b, &B); break;
            }
            oracle_contact *o = &J->res[q];
            memset(o, 0, sizeof(*o));
            o->type = c.type; o->a = c.a; o->b = c.b;
            J->hit[q] = (uint8_t)(oracle_check_collision(&A, &B, o->n, o->point_a, o->point_b, NULL) != 0);
        }
    }
    return NULL;
}

static int detect_grid(const oracle_world *w, oracle_contact *out, int cap)
{
    grid_t g;
    grid_build(w, &g);
    const int nc = w->n_cubes, ns = w->n_spheres, nb = nc + ns;
    /* candidate list in the reference's loop order: CC, CF, SF, CS, SS (code/nans.cpp:1355-1535) */
    size_t ccap = (size_t)16 * (nb + 16), cn = 0;
    cand_t *cand = malloc(sizeof(cand_t) * ccap);
    int32_t *buf = NULL; int bcap = 0;
#define PUSH(T, A, B) do { if (cn == ccap) { ccap *= 2; cand = realloc(cand, sizeof(cand_t) * ccap); } \
                           cand[cn].type = (T); cand[cn].a = (A); cand[cn].b = (B); ++cn; } while (0)
    for (int i = 0; i < nc; ++i) {
        const int n = grid_partners(&g, i, i + 1, nc, &buf, &bcap);
        for (int k = 0; k < n; ++k) PUSH(ORC_CC, i, buf[k]);
    }
    float slo[3], shi[3];
    oracle_shape S;
    const size_t cf_begin = cn;
    for (int i = 0; i < nc; ++i)
        for (int k = 0; k < w->n_statics; ++k) {
            static_shape(w, k, &S); shape_aabb(&S, slo, shi);
            if (aabb_overlap(g.lo + 3 * i, g.hi + 3 * i, slo, shi)) PUSH(ORC_CF, i, k);
        }
    for (int i = 0; i < ns; ++i)
        for (int k = 0; k < w->n_statics; ++k) {
            static_shape(w, k, &S); shape_aabb(&S, slo, shi);
            if (aabb_overlap(g.lo + 3 * (nc + i), g.hi + 3 * (nc + i), slo, shi)) PUSH(ORC_SF, i, k);
        }
    (void)cf_begin;
    for (int i = 0; i < nc && ns > 0; ++i) {
        const int n = grid_partners(&g, i, nc, nb, &buf, &bcap);
        for (int k = 0; k < n; ++k) PUSH(ORC_CS, i, buf[k] - nc);
    }
    for (int i = 0; i < ns; ++i) {
        const int n = grid_partners(&g, nc + i, nc + i + 1, nb, &buf, &bcap);
        for (int k = 0; k < n; ++k) PUSH(ORC_SS, i, buf[k] - nc);
    }
#undef PUSH
    free(buf);
    /* GJK + EPA per candidate: independent pure functions of one pair, so they run on worker threads */
    oracle_contact *res = malloc(sizeof(oracle_contact) * (cn + 1));
    uint8_t *hit = malloc(cn + 1);
    np_job job = {w, cand, res, hit, (long)cn, 0};
    long nthr = sysconf(_SC_NPROCESSORS_ONLN);
    if (nthr > 32) nthr = 32;
    if (nthr < 1 || cn < 4096) nthr = 1;
    pthread_t th[32];
    for (long t = 1; t < nthr; ++t) pthread_create(&th[t], NULL, np_worker, &job);
    np_worker(&job);
    for (long t = 1; t < nthr; ++t) pthread_join(th[t], NULL);
    /* ordered assembly, with the live CS "Exists" quirk (:1479-1489, see the all-pairs loop below) */
    int n = 0, cs_begin = -1;
    for (size_t q = 0; q < cn; ++q) {
        if (cand[q].type == ORC_CS && cs_begin < 0) cs_begin = n;
        if (!hit[q]) continue;
        if (cand[q].type == ORC_CS && cand[q].b < cand[q].a) {
            int dropped = 0;
            const int lim = n < cap ? n : cap;
            for (int k = cs_begin; k < lim; ++k)
                if (out[k].a == cand[q].b && out[k].b == cand[q].a) { dropped = 1; break; }
            if (dropped) continue;
        }
        if (n < cap) out[n] = res[q];
        ++n;
    }
    free(res); free(hit); free(cand);
    grid_free(&g);
    return n;
}

int oracle_detect_collisions(const oracle_world *w, oracle_contact *out, int cap, int prefilter)
{
    if (prefilter == 2) return detect_grid(w, out, cap);
    contact_sink sink = {out, 0, cap};
    oracle_shape A, B;
    /* prefilter acceleration for large N: sort-free O(N^2) on cached AABBs */
    float *lo = NULL, *hi = NULL;
    if (prefilter) {
        lo = malloc(sizeof(float) * 3 * (w->n_cubes + 1));
        hi = malloc(sizeof(float) * 3 * (w->n_cubes + 1));
        for (int i = 0; i < w->n_cubes; ++i) { cube_shape(w, i, &A); shape_aabb(&A, lo + 3 * i, hi + 3 * i); }
    }
    /* CC :1357-1396 */
    for (int i = 0; i < w->n_cubes; ++i) {
        cube_shape(w, i, &A);
        for (int j = i + 1; j < w->n_cubes; ++j) {
            if (prefilter && !aabb_overlap(lo + 3 * i, hi + 3 * i, lo + 3 * j, hi + 3 * j)) continue;
            cube_shape(w, j, &B);
            test_pair(ORC_CC, i, j, &A, &B, 0, &sink);
        }
    }
    /* CF :1400-1430 (one static in the reference; statics looped innermost here) */
    for (int i = 0; i < w->n_cubes; ++i) {
        cube_shape(w, i, &A);
        for (int k = 0; k < w->n_statics; ++k) { static_shape(w, k, &B); test_pair(ORC_CF, i, k, &A, &B, prefilter, &sink); }
    }
    /* SF :1431-1459 */
    for (int i = 0; i < w->n_spheres; ++i) {
        sphere_shape(w, i, &A);
        for (int k = 0; k < w->n_statics; ++k) { static_shape(w, k, &B); test_pair(ORC_SF, i, k, &A, &B, prefilter, &sink); }
    }
    /* CS :1462-1496 (cube-major).  The "Exists" scan :1479-1489 is NOT dead for this type:
     * its second clause matches an already-pushed CS pair with the two indices swapped, so a
     * hit (cube i, sphere j) is dropped when (cube j, sphere i) was pushed earlier (j < i).
     * The same scans for CC/CF/SF/SS can never match (list cleared per frame, :1758). */
    int cs_begin = sink.n;
    for (int i = 0; i < w->n_cubes; ++i) {
        cube_shape(w, i, &A);
        for (int j = 0; j < w->n_spheres; ++j) {
            sphere_shape(w, j, &B);
            int before = sink.n;
            test_pair(ORC_CS, i, j, &A, &B, prefilter, &sink);
            if (sink.n > before && j < i) {
                int lim = before < sink.cap ? before : sink.cap;
                for (int k = cs_begin; k < lim; ++k)
                    if (sink.out[k].a == j && sink.out[k].b == i) { sink.n = before; break; }
            }
        }
    }
    /* SS :1500-1534 */
    for (int i = 0; i < w->n_spheres; ++i) {
        sphere_shape(w, i, &A);
        for (int j = i + 1; j < w->n_spheres; ++j) { sphere_shape(w, j, &B); test_pair(ORC_SS, i, j, &A, &B, prefilter, &sink); }
    }
    free(lo);
    free(hi);
    return sink.n;
}

/* ---- Constraint, code/nans.cpp:1021-1329 ------------------------------------ */
void oracle_constraint(oracle_world *w, const oracle_contact *c, float dt)
{
    int ia, ib = -1; /* body rows; ib < 0 => static */
    v3 posA, posB, V1, W1, V2, W2;
    float invI1, invI2, invM1, invM2;
    switch (c->type) {
    case ORC_CC: ia = c->a; ib = c->b; break;
    case ORC_CS: ia = c->a; ib = w->n_cubes + c->b; break;
    case ORC_CF: ia = c->a; break;
    case ORC_SS: ia = w->n_cubes + c->a; ib = w->n_cubes + c->b; break;
    case ORC_SF: ia = w->n_cubes + c->a; break;
    default: return;
    }
    posA = ld3(w->pos + 3 * ia);
    invI1 = 1.0f / w->moi[ia];
    invM1 = 1.0f / w->mass[ia];
    V1 = ld3(w->vel + 3 * ia);
    W1 = ld3(w->angvel + 3 * ia);
    if (ib >= 0) {
        posB = ld3(w->pos + 3 * ib);
        invI2 = 1.0f / w->moi[ib];
        invM2 = 1.0f / w->mass[ib];
        V2 = ld3(w->vel + 3 * ib);
        W2 = ld3(w->angvel + 3 * ib);
    } else {
        posB = ld3(w->st_pos + 3 * c->b);
        invI2 = 1.0f / w->st_moi[c->b];
        invM2 = 1.0f / w->st_mass[c->b];
        V2 = V3(0, 0, 0);  /* Floor.V, Floor.W stay zero (never updated, :1278-1289) */
        W2 = V3(0, 0, 0);
    }
    v3 N = normalize3(ld3(c->n));
    if (eq3(N, V3(0, 0, 0))) N = normalize3(sub(posB, posA));           /* :1115-1119 */
    v3 R1 = sub(ld3(c->point_a), posA);
    v3 R2 = sub(ld3(c->point_b), posB);
    v3 T1, T2;
    if (N.x >= 0.57735f) T1 = V3(N.y, -N.x, 0.0f); else T1 = V3(0.0f, N.z, -N.y);
    /* :1133 glm::normalize(T1) result is discarded */
    T2 = cross(N, T1);
    float depth = dot(sub(add(posA, R1), add(posB, R2)), N);
    v3 RN1 = cross(R1, N), RN2 = cross(R2, N);
    float JMJn = invM1 + invM2;
    JMJn += invI1 * dot(RN1, RN1) - invI2 * dot(neg(RN2), neg(RN2));
    JMJn = 1.0f / JMJn;
    v3 dVn = sub(sub(add(V1, cross(W1, N)), V2), cross(W2, N));
    float JdVn = dot(dVn, N);
    float Beta = 0.3f, Cr = 0.1f, Cf = 0.1f;
    float B = -Beta / dt * depth + Cr * JdVn;
    v3 R1T1 = cross(R1, T1), R2T1 = cross(R2, T1), R1T2 = cross(R1, T2), R2T2 = cross(R2, T2);
    float JMJt1 = invM1 + invM2;
    JMJt1 += invI1 * dot(R1T1, R1T1) - invI2 * dot(neg(R2T1), neg(R2T1));
    JMJt1 = 1.0f / JMJt1;
    float JMJt2 = invM1 + invM2;
    JMJt2 += invI1 * dot(R1T2, R1T2) - invI2 * dot(neg(R2T2), neg(R2T2));
    JMJt2 = 1.0f / JMJt2;
    v3 dVt1 = sub(sub(add(V1, cross(W1, T1)), V2), cross(W2, T1));
    float JdVt1 = dot(dVt1, T1);
    v3 dVt2 = sub(sub(add(V1, cross(W1, T2)), V2), cross(W2, T2));
    float JdVt2 = dot(dVt2, T2);

    /* :1176-1227 — accumulators start at zero: the caller works on a fresh copy (:1545) */
    float DLN = 0, sumN = 0, DLT1 = 0, sumT1 = 0, DLT2 = 0, sumT2 = 0;
    for (int iter = 0; iter < 70; ++iter) {
        float lambdaN = (-JdVn + B) * JMJn;
        float oldN = sumN;
        sumN += lambdaN;
        if (sumN < 0) sumN = 0.0f;
        DLN = sumN - oldN;

        float lambdaT1 = (-JdVt1) * JMJt1;
        float oldT1 = sumT1;
        sumT1 += lambdaT1;
        float maxT1 = (float)(sqrt(2.0) * (double)Cf * (double)sumN);
        if (sumT1 < -maxT1) sumT1 = -maxT1;
        if (sumT1 > maxT1) sumT1 = maxT1;
        DLT1 = sumT1 - oldT1;

        float lambdaT2 = (-JdVt2) * JMJt2;
        float oldT2 = sumT2;
        sumT2 += lambdaT2;
        float maxT2 = (float)(sqrt(2.0) * (double)Cf * (double)sumN);
        if (sumT2 < -maxT2) sumT2 = -maxT2;
        if (sumT2 > maxT2) sumT2 = maxT2;
        DLT2 = sumT2 - oldT2;
    }
    v3 LI = muls(N, DLN), LIT1 = muls(T1, DLT1), LIT2 = muls(T2, DLT2);
    v3 AI1 = muls(RN1, DLN), AI2 = muls(RN2, DLN);
    v3 AI1T1 = muls(R1T1, DLT1), AI2T1 = muls(R2T1, DLT1);
    v3 AI1T2 = muls(R1T2, DLT2), AI2T2 = muls(R2T2, DLT2);

    /* :1229-1328 — apply in the reference's order; bodies re-read between updates */
    float *va = w->vel + 3 * ia, *wa = w->angvel + 3 * ia;
    float *vb = ib >= 0 ? w->vel + 3 * ib : NULL, *wb = ib >= 0 ? w->angvel + 3 * ib : NULL;
#define APPLY(L, A1, A2)                                        \
    do { st3(va, add(ld3(va), muls(L, invM1)));                 \
         if (vb) st3(vb, sub(ld3(vb), muls(L, invM2)));         \
         st3(wa, add(ld3(wa), muls(A1, invI1)));                \
         if (wb) st3(wb, sub(ld3(wb), muls(A2, invI2))); } while (0)
    APPLY(LI, AI1, AI2);
    APPLY(LIT1, AI1T1, AI2T1);
    APPLY(LIT2, AI1T2, AI2T2);
#undef APPLY
}

/* SolveConstraints :1539-1548 */
void oracle_solve_constraints(oracle_world *w, float dt, const oracle_contact *c, int n)
{
    for (int i = 0; i < n; ++i) oracle_constraint(w, &c[i], dt);
}

/* code/nans.cpp:1758-1762 then the model rebuild of the draw section */
int oracle_step(oracle_world *w, float dt, oracle_contact *scratch, int cap, int prefilter)
{
    oracle_integrate_forces(w, dt);
    int n = oracle_detect_collisions(w, scratch, cap, prefilter);
    oracle_solve_constraints(w, dt, scratch, n < cap ? n : cap);
    oracle_integrate_velocities(w, dt);
    oracle_rebuild_vertices(w);
    return n;
}
