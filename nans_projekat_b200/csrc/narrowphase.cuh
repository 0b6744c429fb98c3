// narrowphase.cuh — GJK intersection + EPA penetration, one pair per thread.
//
// Restructured from the reference's CheckCollision / EvolveSimplex / ResolveCollision
// (code/nans.cpp:907-966, 572-769, 788-904) with bit-identical arithmetic (nans_math.cuh):
//   * both shapes' support vertices are staged in SHARED memory, transposed ([48][threads], bank
//     conflict free), so the support scans index them dynamically at no register cost and a box
//     support is remembered as a 1-byte vertex index instead of a copied vec3;
//   * the GJK simplex (std::vector<vertex>, <= 4 entries while GJK runs) lives in registers;
//     erase()/swap are predicated register moves;
//   * the EPA polytope is index-based and lives in a per-thread arena: vertices P plus support
//     indices (or the sphere support point), faces as three byte indices plus a CACHED unflipped
//     unit normal n and signed plane offset d = dot(n, A.P).  The reference recomputes
//     normalize(cross(AB,AC)) for every face twice per iteration (closest-face scan :813-821 and
//     visibility test :873-881); both are pure functions of the face's vertices, so they are
//     computed once at face creation: |d| is the scan distance, the flipped normal is
//     (d < 0 ? -n : n) = PushTriangle's stored N (:316-320);
//   * std::vector<edge>/<triangle> erase/push_back order is preserved exactly (the closest-face
//     tie-break is "first minimum", so face order is observable);
//   * the horizon's by-value edge cancellation compares precomputed vertex equivalence classes, the
//     closest face is tracked while the face list is rebuilt, and visible faces are listed first and
//     turned into edges afterwards (keeps the warp converged; profiles/README.md).
#pragma once
#include "nans_math.cuh"
#include "world.cuh"

namespace nans {

enum { kNoIntersection = 0, kFoundIntersection = 1, kStillEvolving = 2 };  // evolve_result, code/nans.h:89-94
constexpr int kEpaOutOfBudget = 2;   // epa_resolve with an iteration budget: not finished (never returned for 65)

#ifndef NANS_NP_THREADS
#define NANS_NP_THREADS 128
#endif
constexpr int kNpThreads = NANS_NP_THREADS;
constexpr int kNoVertex = 8;

// Both shapes' box vertices, transposed: g_np_verts[(24*side + 3*k + r) * kNpThreads + tid].  File-scope __shared__
// so every access compiles to LDS/STS (a pointer carried through the call chain degrades to generic loads).
__shared__ float g_np_verts[48 * kNpThreads];

// Per-thread view of the two shapes.
struct NpShapes {
    vec3 posA, posB;     // body centres (GJK start direction; sphere support)
    vec3 dir0;           // normalize(posB - posA): EvolveSimplex recomputes it every call (:575); hoisted
    float radA, radB;    // spheres
    // box vertex k of a side; k = kNoVertex (every support compare failed) is vec3(0)
    __device__ __forceinline__ vec3 vertex(int side, int k) const
    {
        if (k >= 8) return V3(0.f, 0.f, 0.f);
        const float *p = g_np_verts + (24 * side + 3 * k) * kNpThreads + threadIdx.x;
        return V3(p[0], p[kNpThreads], p[2 * kNpThreads]);
    }
    // stage one box: v24 = its 8 packed vec3 (reference vertex order)
    __device__ __forceinline__ static void store_box(int side, const float (&v24)[24])
    {
#pragma unroll
        for (int q = 0; q < 24; ++q) g_np_verts[(24 * side + q) * kNpThreads + threadIdx.x] = v24[q];
    }
};

// GetCubeSupport / GetFloorSupport, code/nans.cpp:410-430,441-461: first vertex with strictly
// greater dot; vec3(0) if every compare fails (NaN direction).  Keeps (best, index) only while
// scanning and fetches the winner afterwards.
__device__ __forceinline__ vec3 box_support(const NpShapes &S, int side, vec3 d, int &idx)
{
    float dist[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float *p = g_np_verts + 24 * side * kNpThreads + threadIdx.x;
        const vec3 c = V3(p[(3 * k) * kNpThreads], p[(3 * k + 1) * kNpThreads], p[(3 * k + 2) * kNpThreads]);
        dist[k] = dot(c, d);
    }
    float best = -FLT_MAX;
    idx = kNoVertex;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (dist[k] > best) { best = dist[k]; idx = k; }
    return S.vertex(side, idx);
}
// GetSphereSupport, code/nans.cpp:433-438
__device__ __forceinline__ vec3 sphere_support(vec3 pos, float radius, vec3 d)
{
    return pos + radius * normalize(d);
}

// A support record: vertex index for a box, the point itself for a sphere.
template <bool SPHERE> struct SupRec;
template <> struct SupRec<false> { int idx; };
template <> struct SupRec<true> { vec3 v; };

template <bool A_SPHERE, bool B_SPHERE>
struct GjkVertex {           // struct vertex, code/nans.h:245-255
    vec3 P;
    SupRec<A_SPHERE> a;
    SupRec<B_SPHERE> b;
};

template <bool SPHERE>
__device__ __forceinline__ vec3 support_of(const NpShapes &S, int side, vec3 d, SupRec<SPHERE> &rec)
{
    if constexpr (SPHERE) {
        rec.v = sphere_support(side ? S.posB : S.posA, side ? S.radB : S.radA, d);
        return rec.v;
    } else {
        return box_support(S, side, d, rec.idx);
    }
}
// CalculateSupport, code/nans.cpp:464-519
template <bool AS, bool BS>
__device__ __forceinline__ GjkVertex<AS, BS> calc_support(const NpShapes &S, vec3 d)
{
    GjkVertex<AS, BS> r;
    const vec3 sa = support_of<AS>(S, 0, d, r.a);
    const vec3 sb = support_of<BS>(S, 1, -1.0f * d, r.b);
    r.P = sa - sb;
    return r;
}

template <typename V> __device__ __forceinline__ void simplex_erase(V (&s)[4], int &n, int i)
{
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (k >= i) s[k] = s[k + 1];
    --n;
}
template <typename V> __device__ __forceinline__ void simplex_push(V (&s)[4], int &n, const V &v)
{
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k == n) s[k] = v;
    ++n;
}

// TripleCross, code/nans.cpp:565-569
__device__ __forceinline__ vec3 triple_cross(vec3 A, vec3 B, vec3 C) { return (B * dot(C, A)) - (A * dot(C, B)); }

// EvolveSimplex, code/nans.cpp:572-769
template <bool AS, bool BS>
__device__ __forceinline__ int evolve_simplex(const NpShapes &S, GjkVertex<AS, BS> (&s)[4], int &n)
{
    vec3 dir = S.dir0;
    bool found = false;
    if (n == 1) {
        dir = dir * -1.0f;
    } else if (n == 2) {
        // ClosestPointOnLine, code/nans.cpp:540-562 (normalised segment: U,V are not barycentrics)
        const vec3 a = s[0].P, b = s[1].P;
        const vec3 seg = normalize(b - a);
        const float len = length(seg);
        const float V = fdiv(dot(-a, seg), len);
        const float U = fdiv(dot(b, seg), len);
        vec3 cp;
        if (U <= 0.0f) cp = b;
        else if (V <= 0.0f) cp = a;
        else cp = (U * a) + (V * b);
        if (V <= 0.0f) simplex_erase(s, n, 1);
        else if (U <= 0.0f) simplex_erase(s, n, 0);
        dir = -cp;
    } else if (n == 3) {
        const vec3 ao = -s[0].P;
        const vec3 e1 = s[1].P - s[0].P;
        const vec3 e2 = s[2].P - s[0].P;
        const vec3 tn = cross(e1, e2);
        const vec3 e1n = cross(e1, tn);
        const vec3 e2n = cross(tn, e2);
        if (dot(e2n, ao) > 0.0f) {
            if (dot(e2, ao) > 0.0f) { dir = triple_cross(e2, ao, e2); simplex_erase(s, n, 1); }
            else if (dot(e1, ao) > 0.0f) { dir = triple_cross(e1, ao, e1); simplex_erase(s, n, 2); }
            else { dir = ao; simplex_erase(s, n, 2); simplex_erase(s, n, 1); }
        } else if (dot(e1n, ao) > 0.0f) {
            if (dot(e1, ao) > 0.0f) { dir = triple_cross(e1, ao, e1); simplex_erase(s, n, 2); }
            else { dir = ao; simplex_erase(s, n, 2); simplex_erase(s, n, 1); }
        } else if (dot(tn, ao) > 0.0f) {
            dir = tn;
        } else {
            dir = -tn;
            const GjkVertex<AS, BS> t = s[1]; s[1] = s[2]; s[2] = t;
        }
    } else if (n == 4) {
        const vec3 da = s[0].P - s[3].P;
        const vec3 db = s[1].P - s[3].P;
        const vec3 dc = s[2].P - s[3].P;
        const vec3 d0 = -1.0f * s[3].P;
        const vec3 abd = cross(da, db), bcd = cross(db, dc), cad = cross(dc, da);
        if (dot(abd, d0) > 0.0f) { simplex_erase(s, n, 2); dir = abd; }
        else if (dot(bcd, d0) > 0.0f) { simplex_erase(s, n, 0); dir = bcd; }
        else if (dot(cad, d0) > 0.0f) { simplex_erase(s, n, 1); dir = cad; }
        else found = true;
    }
    // (the origin-enclosed exit is taken HERE, after the case chain, so that the chain's branches reconverge
    // before the support call below instead of each running it on its own)
    if (found) return kFoundIntersection;
    if (length(dir) <= 0.0001f) return kNoIntersection;
    // AddSupport, code/nans.cpp:522-537
    const GjkVertex<AS, BS> nv = calc_support<AS, BS>(S, dir);
    simplex_push(s, n, nv);
    return dot(dir, nv.P) >= 0.0f ? kStillEvolving : kNoIntersection;
}

// ---- EPA --------------------------------------------------------------------------------------
// Per-thread arena (local memory; only the touched part costs traffic).  A face record is 20 bytes:
// the unflipped unit normal n with d = dot(n, A.P), and the packed vertex indices a | b<<8 | c<<16.
// (Carrying A.P in the record as well saved a dependent load but grew the arena: 640 threads/SM x
// the touched part already exceeds L1, and the smaller record measured 6 % faster.)
// The polytope's first kEpaShP vertices live in SHARED memory, transposed like the shapes ([3 * k + r][thread]: any
// per-lane index is bank-conflict free): every P access feeds arithmetic at once (the face scan's `ns.P - P[a]`
// with a per-lane index, the class scan of a new vertex, a new face's edges), and in the local-memory arena those
// loads were 45 % of the EPA kernel's stall samples (profiles/r3_np_split_ncu_summary.txt).  Later vertices (only
// long EPA runs have them) stay in the arena.
#ifndef NANS_EPA_SHP
#define NANS_EPA_SHP 0
#endif
constexpr int kEpaShP = NANS_EPA_SHP;
#ifndef NANS_EPA_EDGE4
#define NANS_EPA_EDGE4 0
#endif
#ifndef NANS_EPA_CLS_NOBREAK
#define NANS_EPA_CLS_NOBREAK 0
#endif
#ifndef NANS_EPA_FACE_PIPE2
#define NANS_EPA_FACE_PIPE2 0
#endif
#if NANS_EPA_SHP > 0
__shared__ float g_np_P[3 * kEpaShP * kNpThreads];
#endif

struct EpaGenericArena {
    vec3 Pl[kEpaMaxVerts - kEpaShP];
    __device__ __forceinline__ vec3 getP(int i) const
    {
#if NANS_EPA_SHP > 0
        if (i < kEpaShP) {
            const float *p = g_np_P + 3 * i * kNpThreads + threadIdx.x;
            return V3(p[0], p[kNpThreads], p[2 * kNpThreads]);
        }
#endif
        return Pl[i - kEpaShP];
    }
    __device__ __forceinline__ void setP(int i, vec3 v)
    {
#if NANS_EPA_SHP > 0
        if (i < kEpaShP) {
            float *p = g_np_P + 3 * i * kNpThreads + threadIdx.x;
            p[0] = v.x; p[kNpThreads] = v.y; p[2 * kNpThreads] = v.z;
            return;
        }
#endif
        Pl[i - kEpaShP] = v;
    }
    vec3 SA[kEpaMaxVerts], SB[kEpaMaxVerts];    // sphere sides only
    uint8_t ia[kEpaMaxVerts], ib[kEpaMaxVerts]; // box sides only
    uint8_t cid[kEpaMaxVerts];                  // lowest vertex index with an equal P (kCidNaN: equal to nothing)
    float4 fnd[kEpaMaxFaces];
    uint32_t fidx[kEpaMaxFaces];
    uint32_t vis[kEpaMaxFaces];                 // packed indices of the faces dissolved this iteration
    uint32_t edge[kEpaMaxEdges];                // a | b<<8 | cid[a]<<16 | cid[b]<<24
};
using EpaArena = EpaGenericArena;
constexpr int kCidNaN = 254;

// The reference's edge cancels an opposite-winding edge BY VALUE of P (code/nans.h:251-254).  Float
// equality is an equivalence on non-NaN vectors (+0 == -0 included), so every vertex gets the lowest
// index of its class once, when it is stored, and the edge compares become one integer compare.
template <bool AS, bool BS>
__device__ __forceinline__ void epa_store_vertex(EpaGenericArena &E, int i, const GjkVertex<AS, BS> &v)
{
    E.setP(i, v.P);
    if constexpr (AS) E.SA[i] = v.a.v; else E.ia[i] = (uint8_t)v.a.idx;
    if constexpr (BS) E.SB[i] = v.b.v; else E.ib[i] = (uint8_t)v.b.idx;
    int c = i;
    if (!equal(v.P, v.P)) {
        c = kCidNaN;
    } else {
#if NANS_EPA_CLS_NOBREAK
        // lowest equal index without an early exit: the loads do not wait for each other's compare
        for (int j = i - 1; j >= 0; --j)
            if (equal(E.getP(j), v.P)) c = j;
#else
        for (int j = 0; j < i; ++j)
            if (equal(E.getP(j), v.P)) { c = j; break; }
#endif
    }
    E.cid[i] = (uint8_t)c;
}
template <bool AS> __device__ __forceinline__ vec3 epa_sup_a(const EpaGenericArena &E, const NpShapes &S, int i)
{
    if constexpr (AS) return E.SA[i]; else return S.vertex(0, E.ia[i]);
}
template <bool BS> __device__ __forceinline__ vec3 epa_sup_b(const EpaGenericArena &E, const NpShapes &S, int i)
{
    if constexpr (BS) return E.SB[i]; else return S.vertex(1, E.ib[i]);
}

// closest face = FIRST strict minimum of |d| in face order (:807-822).  The reference rescans every
// iteration; here the minimum is carried along while the face list is rebuilt in the same order.
__device__ __forceinline__ void epa_track_min(float d, int slot, float &cur, int &ci)
{
    const float dist = fabsf(d);
    if (slot == 0 || dist < cur) { cur = dist; ci = slot; }
}

__device__ __forceinline__ void epa_push_face(EpaGenericArena &E, int &nf, int a, int b, int c, vec3 pa, float &cur, int &ci)
{
    // PushTriangle, code/nans.cpp:293-322 (flip folded into the sign of d); pa == E.P[a]
    const vec3 n = normalize(cross(E.getP(b) - pa, E.getP(c) - pa));
    const float d = dot(pa, n);
    E.fnd[nf] = make_float4(n.x, n.y, n.z, d);
    E.fidx[nf] = (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16);
    epa_track_min(d, nf, cur, ci);
    ++nf;
}
__device__ __forceinline__ vec3 face_normal_flipped(const float4 &nd)
{
    const vec3 n = V3(nd);
    return nd.w < 0.0f ? n * -1.0f : n;
}

// PushEdge, code/nans.cpp:233-266: an opposite-winding edge already in the list is erased (order of
// the rest kept), otherwise the edge is appended
__device__ __forceinline__ void epa_push_edge(EpaGenericArena &E, int &ne, int a, int b, int &ovf)
{
    const uint32_t ca = E.cid[a], cb = E.cid[b];
    // a NaN vertex equals nothing, itself included: its id on the probing side never matches a stored one
    const uint32_t want = (cb == kCidNaN ? 255u : cb) | ((ca == kCidNaN ? 255u : ca) << 8);
    int i = 0;
#if NANS_EPA_EDGE4
    // four entries per round trip (the one-at-a-time walk is a chain of dependent local loads); entries past ne are
    // read (inside the array) and ignored
    int at = -1;
    for (int b = 0; b < ne && at < 0; b += 4) {
        const uint32_t e0 = E.edge[b], e1 = E.edge[b + 1], e2 = E.edge[b + 2], e3 = E.edge[b + 3];
        if ((e0 >> 16) == want) at = b;
        else if (b + 1 < ne && (e1 >> 16) == want) at = b + 1;
        else if (b + 2 < ne && (e2 >> 16) == want) at = b + 2;
        else if (b + 3 < ne && (e3 >> 16) == want) at = b + 3;
    }
    i = at < 0 ? ne : at;
#else
    while (i < ne && (E.edge[i] >> 16) != want) ++i;
#endif
    if (i < ne) {
        for (int k = i; k < ne - 1; ++k) E.edge[k] = E.edge[k + 1];
        --ne;
        return;
    }
    if (ne >= kEpaMaxEdges) { ovf |= OVF_EPA_EDGES; return; }
    E.edge[ne++] = (uint32_t)a | ((uint32_t)b << 8) | (ca << 16) | (cb << 24);
}

// ResolveCollision, code/nans.cpp:788-904.  Returns the bool32 result; fills PointA/PointB/N.
template <bool AS, bool BS>
__device__ __forceinline__ int epa_resolve(const NpShapes &S, const GjkVertex<AS, BS> (&s)[4], EpaGenericArena &E,
                                           vec3 &outPA, vec3 &outPB, vec3 &outN, int &ovf, int &max_faces,
                                           int max_iters = 65)
{
    // the simplex becomes the first four vertices and faces (:791-802)
#pragma unroll
    for (int k = 0; k < 4; ++k) epa_store_vertex<AS, BS>(E, k, s[k]);
    int nv = 4, nf = 0, ne = 0, ci = 0, it = 0;
    float cur = 0.f;
    epa_push_face(E, nf, 0, 1, 2, s[0].P, cur, ci);  // ABC
    epa_push_face(E, nf, 0, 2, 3, s[0].P, cur, ci);  // ACD
    epa_push_face(E, nf, 0, 3, 1, s[0].P, cur, ci);  // ADB
    epa_push_face(E, nf, 1, 3, 2, s[1].P, cur, ci);  // BDC
    while (it++ <= 64) {            // MAX_EPA_ITERATIONS, code/nans.h:56
        // a BUDGETED run (max_iters < 65: the multi-pass batch path) gives up before iteration max_iters + 1; the
        // caller runs the pair again, from its simplex, with a larger budget
        if (it > max_iters) return kEpaOutOfBudget;
        max_faces = max(max_faces, nf);
        const float4 cnd = E.fnd[ci];
        const vec3 N = face_normal_flipped(cnd);
        const GjkVertex<AS, BS> ns = calc_support<AS, BS>(S, N);
        if (fsub(dot(N, ns.P), cur) < 0.001f) {   // MAX_EPA_ERROR, code/nans.h:55
            const uint32_t f = E.fidx[ci];
            const int a = f & 255, b = (f >> 8) & 255, c = (f >> 16) & 255;
            // Barycentric, code/nans.cpp:772-785
            const vec3 Pp = N * cur;
            const vec3 A0 = E.getP(a);
            const vec3 v0 = E.getP(b) - A0, v1 = E.getP(c) - A0, v2 = Pp - A0;
            const float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1);
            const float d20 = dot(v2, v0), d21 = dot(v2, v1);
            const float denom = fsub(fmul(d00, d11), fmul(d01, d01));
            const float bv = fdiv(fsub(fmul(d11, d20), fmul(d01, d21)), denom);
            const float bw = fdiv(fsub(fmul(d00, d21), fmul(d01, d20)), denom);
            const float bu = fsub(fsub(1.0f, bv), bw);
            if (fabsf(bu) > 1.0f || fabsf(bv) > 1.0f || fabsf(bw) > 1.0f) return 0;
            if (!isfinite(bu) || !isfinite(bv) || !isfinite(bw)) return 0;   // IsValid, :4-17
            outPA = ((bu * epa_sup_a<AS>(E, S, a)) + (bv * epa_sup_a<AS>(E, S, b))) + (bw * epa_sup_a<AS>(E, S, c));
            outN = -1.0f * N;
            outPB = ((bu * epa_sup_b<BS>(E, S, a)) + (bv * epa_sup_b<BS>(E, S, b))) + (bw * epa_sup_b<BS>(E, S, c));
            return 1;
        }
        if (nv >= kEpaMaxVerts) { ovf |= OVF_EPA_FACES; return 0; }
        epa_store_vertex<AS, BS>(E, nv, ns);
        // dissolve every face the new point can see (:869-891); survivors keep their order.  The
        // dissolved faces are only LISTED here; their edges are pushed in a second loop, so the warp
        // stays converged over the face scan.
        const int nf_old = nf;
        int keep = 0, nvis = 0;
#if NANS_EPA_FACE_PIPE2
        // two faces in flight: face i's record AND its vertex A (a dependent, per-lane-indexed load) are fetched one
        // iteration ahead, face i + 2's record two ahead
        float4 nd_next = E.fnd[0];
        uint32_t f_next = E.fidx[0];
        vec3 pa_next = E.getP(f_next & 255);
        float4 nd_nn = nd_next;
        uint32_t f_nn = f_next;
        if (nf > 1) { nd_nn = E.fnd[1]; f_nn = E.fidx[1]; }
        for (int i = 0; i < nf; ++i) {
            const float4 nd = nd_next;
            const uint32_t f = f_next;
            const vec3 pa_cur = pa_next;
            nd_next = nd_nn; f_next = f_nn;
            if (i + 1 < nf) pa_next = E.getP(f_next & 255);
            if (i + 2 < nf) { nd_nn = E.fnd[i + 2]; f_nn = E.fidx[i + 2]; }
            const vec3 tmp = ns.P - pa_cur;
#else
        float4 nd_next = E.fnd[0];
        uint32_t f_next = E.fidx[0];
        for (int i = 0; i < nf; ++i) {
            const float4 nd = nd_next;
            const uint32_t f = f_next;
            if (i + 1 < nf) { nd_next = E.fnd[i + 1]; f_next = E.fidx[i + 1]; }
            const vec3 tmp = ns.P - E.getP(f & 255);
#endif
            if (dot(face_normal_flipped(nd), tmp) > 0.0f) {
                E.vis[nvis++] = f;
            } else {
                if (keep != i) { E.fnd[keep] = nd; E.fidx[keep] = f; }
                epa_track_min(nd.w, keep, cur, ci);
                ++keep;
            }
        }
        nf = keep;
        for (int j = 0; j < nvis; ++j) {
            uint32_t f = E.vis[j];
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {           // AB, BC, CA
                epa_push_edge(E, ne, f & 255, (f >> 8) & 255, ovf);
                f = (f >> 8) | ((f & 255) << 16);
            }
        }
        // one new face per horizon edge, in edge-list order (:894-901)
        if (nf + ne > kEpaMaxFaces) { ovf |= OVF_EPA_FACES; return 0; }
        for (int i = 0; i < ne; ++i) {
            const uint32_t ed = E.edge[i];
            epa_push_face(E, nf, nv, ed & 255, (ed >> 8) & 255, ns.P, cur, ci);
        }
        ne = 0;
        ++nv;
        // The EMPTIED polytope.  When the new point sees every face and every horizon edge cancels (the origin
        // lies on a face plane of a flat start tetrahedron: seen once per ~10^6 pairs of a settling pile), the
        // reference's std::vector<triangle> is empty, and its next iteration still reads Triangle[0] (:807-811,
        // 824-866).  erase() shifted the list down one element at a time, so that storage slot holds the LAST
        // face of the dissolved list; the reference goes on with it as the closest face (direction, distance,
        // barycentrics) in every remaining iteration.  The prebuilt nans.so behaves exactly so (tests/golden/
        // epa_emptied.npz); kept here: slot 0 := that face, ci = 0, cur = its |d|, nf stays 0.
        if (nf == 0 && nf_old > 0) {
            const float4 last = E.fnd[nf_old - 1];
            E.fnd[0] = last;
            E.fidx[0] = E.fidx[nf_old - 1];
            cur = fabsf(last.w);
            ci = 0;
        }
    }
    return 0;
}

struct NpResult { int hit, gjk; vec3 PA, PB, N; };

// the GJK loop of CheckCollision (:957-965): the evolve_result it ended with; s = the final simplex
template <bool AS, bool BS>
__device__ __forceinline__ int gjk_run(NpShapes &S, GjkVertex<AS, BS> (&s)[4])
{
    int n = 0, ev = kStillEvolving, iter = 0;
    S.dir0 = normalize(S.posB - S.posA);
    while (ev == kStillEvolving && iter++ <= 64)   // MAX_GJK_ITERATIONS, code/nans.h:54
        ev = evolve_simplex<AS, BS>(S, s, n);
    return ev;
}

// The same loop, resumable: at most `budget` more evolutions; n and iter carry the progress.  Finished when the
// result is not kStillEvolving or iter > 64 (the reference's own limit).
template <bool AS, bool BS>
__device__ __forceinline__ int gjk_resume(NpShapes &S, GjkVertex<AS, BS> (&s)[4], int &n, int &iter, int budget)
{
    int ev = kStillEvolving;
    S.dir0 = normalize(S.posB - S.posA);
    while (ev == kStillEvolving && budget-- > 0 && iter++ <= 64) ev = evolve_simplex<AS, BS>(S, s, n);
    return ev;
}

// CheckCollision, code/nans.cpp:907-966
template <bool AS, bool BS>
__device__ __noinline__ NpResult check_collision(NpShapes &S, EpaArena &E, int &ovf, int &max_faces)
{
    GjkVertex<AS, BS> s[4];
#ifdef NANS_NP_GJK_CAPPED   // the capped-then-continued GJK of the split batch path (host check of gjk_resume)
    int n_ = 0, iter_ = 0;
    int ev = gjk_resume<AS, BS>(S, s, n_, iter_, NANS_NP_GJK_CAPPED);
    if (ev == kStillEvolving && iter_ <= 64) ev = gjk_resume<AS, BS>(S, s, n_, iter_, 1000);
#else
    const int ev = gjk_run<AS, BS>(S, s);
#endif
    NpResult r;
    r.gjk = ev;
    r.hit = 0;
    r.PA = r.PB = r.N = V3(0.f, 0.f, 0.f);
    if (ev == kFoundIntersection) r.hit = epa_resolve<AS, BS>(S, s, E, r.PA, r.PB, r.N, ovf, max_faces);
    return r;
}

}  // namespace nans
