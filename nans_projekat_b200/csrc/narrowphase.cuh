// narrowphase.cuh — GJK intersection + EPA penetration, one pair per thread.
//
// Restructured from the reference's CheckCollision / EvolveSimplex / ResolveCollision
// (code/nans.cpp:907-966, 572-769, 788-904) with bit-identical arithmetic (nans_math.cuh):
//   * both shapes' support vertices are staged in SHARED memory, transposed ([48][threads], bank
//     conflict free), so the support scans index them dynamically at no register cost and a box
//     support is remembered as a 1-byte vertex index instead of a copied vec3;
//   * the GJK simplex (std::vector<vertex>, <= 4 entries while GJK runs) lives in registers;
//     erase()/swap are predicated register moves;
//   * the EPA polytope is index-based and lives in a per-thread arena kept as small as it can be (the
//     arenas of the resident threads compete for the L1): a box-pair vertex is its two support
//     indices (P = A[ia] - B[ib] is re-formed from shared memory), a face is three byte indices plus
//     a CACHED unflipped unit normal n; the signed plane offset d = dot(n, A.P) is re-formed next to
//     the visibility test; the face also carries the support indices of its first vertex, so that P[a] is
//     re-formed without a second dependent load.  The reference recomputes normalize(cross(AB,AC)) for every face twice
//     per iteration (closest-face scan :813-821 and visibility test :873-881); it is a pure function
//     of the face's vertices, so it is computed once at face creation: |d| is the scan distance,
//     the flipped normal is (d < 0 ? -n : n) = PushTriangle's stored N (:316-320);
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
// Per-thread arena in local memory; only the touched part costs traffic, and that part is what bounds the kernel:
// the arenas of the resident threads (640 per SM) do not fit the L1 next to the staged shapes, fewer resident
// threads run slower, so every byte the arena does not hold is a byte that cannot miss.  Hence:
//   * a vertex is ONE 32-bit word: its two box-support indices, its equivalence class and a 16-bit hash of P (a
//     sphere side adds its support point, 12 bytes).  P itself is never stored: P = SupA - SupB (CalculateSupport,
//     :464-519) is re-formed from the same operands wherever it is used -- for a box pair six conflict-free
//     shared-memory loads and three subtractions (round 1 stored 15 bytes per vertex);
//   * a face record is 16 bytes {unflipped unit normal n, packed indices a | b<<8 | c<<16 | support indices of a << 24}; the plane offset
//     d = dot(P[a], n) is re-formed next to the visibility test, which needs P[a] anyway (round 1: 20 bytes;
//     carrying P[a] in the record as well was measured then: slower).
// Measured, step by step (1 M-cube pile 100^3 / flat 250x250x16 / config C3): narrowphase stage 0.367 -> 0.355 ->
// 0.347 ms, 0.723 -> 0.681 -> 0.661 ms, 21.85 -> 21.05 -> 19.45 ms; L1 hit rate of the EPA kernel 67 -> 80 %.  Also
// measured, no gain, removed: the first 8-12 vertices in shared memory (the larger carve-out halves the L1), a 4-wide
// edge search, a class scan without early exit, the face scan's vertex fetched one face ahead (profiles/README.md).
struct EpaGenericArena {
    uint32_t vw[kEpaMaxVerts];                  // ia | ib << 4 | class << 8 | hash16(P) << 16
    vec3 SA[kEpaMaxVerts], SB[kEpaMaxVerts];    // sphere sides only: the support points
    float4 fnd[kEpaMaxFaces];                   // {n.xyz, packed indices}
    uint32_t vis[kEpaMaxFaces];                 // packed indices of the faces dissolved this iteration
    uint32_t edge[kEpaMaxEdges];                // a | b<<8 | class[a]<<16 | class[b]<<24
};
using EpaArena = EpaGenericArena;
constexpr int kCidNaN = 254;

template <bool AS, bool BS> __device__ __forceinline__ vec3 epa_sup_a(const EpaGenericArena &E, const NpShapes &S, int i)
{
    if constexpr (AS) return E.SA[i]; else return S.vertex(0, E.vw[i] & 15u);
}
template <bool AS, bool BS> __device__ __forceinline__ vec3 epa_sup_b(const EpaGenericArena &E, const NpShapes &S, int i)
{
    if constexpr (BS) return E.SB[i]; else return S.vertex(1, (E.vw[i] >> 4) & 15u);
}
// vertex i of the polytope, re-formed
template <bool AS, bool BS>
__device__ __forceinline__ vec3 epa_P(const EpaGenericArena &E, const NpShapes &S, int i)
{
    if constexpr (!AS && !BS) {
        const uint32_t w = E.vw[i];             // one load for both indices
        return S.vertex(0, w & 15u) - S.vertex(1, (w >> 4) & 15u);
    } else {
        return epa_sup_a<AS, BS>(E, S, i) - epa_sup_b<AS, BS>(E, S, i);
    }
}
__device__ __forceinline__ uint32_t epa_cid(const EpaGenericArena &E, int i) { return (E.vw[i] >> 8) & 255u; }

// A face word is a | b << 8 | c << 16 | (box-support indices of vertex a: ia | ib << 4) << 24: the face scan needs P[a]
// of every face, and with the indices in the word it re-forms it without the dependent load of vw[a] (that load was
// the kernel's first stall line: 10 % of the samples).  With it, the closest face keeps the sign of its d (no
// re-forming of P[a] at the head of an iteration) and the visibility test takes one dot product (the flipped
// normal is never formed).  Measured (A/B, profiles/r4_ab_*.json): narrowphase stage 0.336 -> 0.324 ms on the 100^3
// pile, 0.647 -> 0.623 ms on the flat pile, config C3 8.39e8 -> 8.58e8 pairs/s, config C4 unchanged.
template <bool AS, bool BS> __device__ __forceinline__ uint32_t epa_sup_byte(const GjkVertex<AS, BS> &v)
{
    uint32_t w = 0;
    if constexpr (!AS) w |= (uint32_t)v.a.idx;
    if constexpr (!BS) w |= (uint32_t)v.b.idx << 4;
    return w;
}
// P of a face's first vertex (same operands as epa_P: the same bits)
template <bool AS, bool BS>
__device__ __forceinline__ vec3 epa_P_face(const EpaGenericArena &E, const NpShapes &S, uint32_t fw)
{
    vec3 sa, sb;
    if constexpr (AS) sa = E.SA[fw & 255u]; else sa = S.vertex(0, (fw >> 24) & 15u);
    if constexpr (BS) sb = E.SB[fw & 255u]; else sb = S.vertex(1, fw >> 28);
    return sa - sb;
}

// equal vectors have equal hashes: x + 0 maps -0 to +0, the one pair of different bit patterns that compare equal
__device__ __forceinline__ uint32_t epa_hash16(vec3 p)
{
    uint32_t h = __float_as_uint(fadd(p.x, 0.0f)) ^ (__float_as_uint(fadd(p.y, 0.0f)) * 0x9E3779B1u) ^
                 (__float_as_uint(fadd(p.z, 0.0f)) * 0x85EBCA6Bu);
    h ^= h >> 16;
    return h & 0xffffu;
}

// The reference's edge cancels an opposite-winding edge BY VALUE of P (code/nans.h:251-254).  Float
// equality is an equivalence on non-NaN vectors (+0 == -0 included), so every vertex gets the lowest
// index of its class once, when it is stored, and the edge compares become one integer compare.  The stored hash
// rules an earlier vertex out without re-forming its P.
template <bool AS, bool BS>
__device__ __forceinline__ void epa_store_vertex(EpaGenericArena &E, const NpShapes &S, int i, const GjkVertex<AS, BS> &v)
{
    uint32_t w = 0;
    if constexpr (AS) E.SA[i] = v.a.v; else w |= (uint32_t)v.a.idx;
    if constexpr (BS) E.SB[i] = v.b.v; else w |= (uint32_t)v.b.idx << 4;
    const uint32_t h = epa_hash16(v.P);
    int c = i;
    if (!equal(v.P, v.P)) {
        c = kCidNaN;
    } else {
        for (int j = 0; j < i; ++j)
            if ((E.vw[j] >> 16) == h && equal(epa_P<AS, BS>(E, S, j), v.P)) { c = j; break; }
    }
    E.vw[i] = w | ((uint32_t)c << 8) | (h << 16);
}

// closest face = FIRST strict minimum of |d| in face order (:807-822).  The reference rescans every
// iteration; here the minimum is carried along while the face list is rebuilt in the same order.
__device__ __forceinline__ void epa_track_min(float d, int slot, float &cur, int &ci, bool &cneg)
{
    const float dist = fabsf(d);
    if (slot == 0 || dist < cur) { cur = dist; ci = slot; cneg = d < 0.0f; }
}

template <bool AS, bool BS>
__device__ __forceinline__ void epa_push_face(EpaGenericArena &E, const NpShapes &S, int &nf, int a, int b, int c, vec3 pa, uint32_t sup_a,
                                              float &cur, int &ci, bool &cneg)
{
    // PushTriangle, code/nans.cpp:293-322 (flip folded into the sign of d); pa == P[a]
    const vec3 n = normalize(cross(epa_P<AS, BS>(E, S, b) - pa, epa_P<AS, BS>(E, S, c) - pa));
    const float d = dot(pa, n);
    E.fnd[nf] = make_float4(n.x, n.y, n.z, __uint_as_float((uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | (sup_a << 24)));
    epa_track_min(d, nf, cur, ci, cneg);
    ++nf;
}

// PushEdge, code/nans.cpp:233-266: an opposite-winding edge already in the list is erased (order of
// the rest kept), otherwise the edge is appended
__device__ __forceinline__ void epa_push_edge(EpaGenericArena &E, int &ne, int a, int b, uint32_t ca, uint32_t cb, int &ovf)
{
    // a NaN vertex equals nothing, itself included: its id on the probing side never matches a stored one
    const uint32_t want = (cb == kCidNaN ? 255u : cb) | ((ca == kCidNaN ? 255u : ca) << 8);
    int i = 0;
    while (i < ne && (E.edge[i] >> 16) != want) ++i;
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
    for (int k = 0; k < 4; ++k) epa_store_vertex<AS, BS>(E, S, k, s[k]);
    int nv = 4, nf = 0, ne = 0, ci = 0, it = 0;
    float cur = 0.f;
    bool cneg = false;              // the closest face's d is negative (its stored N is -n)
    const uint32_t sb0 = epa_sup_byte<AS, BS>(s[0]), sb1 = epa_sup_byte<AS, BS>(s[1]);
    epa_push_face<AS, BS>(E, S, nf, 0, 1, 2, s[0].P, sb0, cur, ci, cneg);  // ABC
    epa_push_face<AS, BS>(E, S, nf, 0, 2, 3, s[0].P, sb0, cur, ci, cneg);  // ACD
    epa_push_face<AS, BS>(E, S, nf, 0, 3, 1, s[0].P, sb0, cur, ci, cneg);  // ADB
    epa_push_face<AS, BS>(E, S, nf, 1, 3, 2, s[1].P, sb1, cur, ci, cneg);  // BDC
    while (it++ <= 64) {            // MAX_EPA_ITERATIONS, code/nans.h:56
        // a BUDGETED run (max_iters < 65: the multi-pass batch path) gives up before iteration max_iters + 1; the
        // caller runs the pair again, from its simplex, with a larger budget
        if (it > max_iters) return kEpaOutOfBudget;
        max_faces = max(max_faces, nf);
        const float4 cnd = E.fnd[ci];
        const uint32_t cf = __float_as_uint(cnd.w);
        const vec3 N = cneg ? V3(cnd) * -1.0f : V3(cnd);     // the sign of d was kept when the face became the closest
        const GjkVertex<AS, BS> ns = calc_support<AS, BS>(S, N);
        if (fsub(dot(N, ns.P), cur) < 0.001f) {   // MAX_EPA_ERROR, code/nans.h:55
            const int a = cf & 255, b = (cf >> 8) & 255, c = (cf >> 16) & 255;
            // Barycentric, code/nans.cpp:772-785
            const vec3 Pp = N * cur;
            const vec3 A0 = epa_P_face<AS, BS>(E, S, cf);
            const vec3 v0 = epa_P<AS, BS>(E, S, b) - A0, v1 = epa_P<AS, BS>(E, S, c) - A0, v2 = Pp - A0;
            const float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1);
            const float d20 = dot(v2, v0), d21 = dot(v2, v1);
            const float denom = fsub(fmul(d00, d11), fmul(d01, d01));
            const float bv = fdiv(fsub(fmul(d11, d20), fmul(d01, d21)), denom);
            const float bw = fdiv(fsub(fmul(d00, d21), fmul(d01, d20)), denom);
            const float bu = fsub(fsub(1.0f, bv), bw);
            if (fabsf(bu) > 1.0f || fabsf(bv) > 1.0f || fabsf(bw) > 1.0f) return 0;
            if (!isfinite(bu) || !isfinite(bv) || !isfinite(bw)) return 0;   // IsValid, :4-17
            outPA = ((bu * epa_sup_a<AS, BS>(E, S, a)) + (bv * epa_sup_a<AS, BS>(E, S, b))) + (bw * epa_sup_a<AS, BS>(E, S, c));
            outN = -1.0f * N;
            outPB = ((bu * epa_sup_b<AS, BS>(E, S, a)) + (bv * epa_sup_b<AS, BS>(E, S, b))) + (bw * epa_sup_b<AS, BS>(E, S, c));
            return 1;
        }
        if (nv >= kEpaMaxVerts) { ovf |= OVF_EPA_FACES; return 0; }
        epa_store_vertex<AS, BS>(E, S, nv, ns);
        // dissolve every face the new point can see (:869-891); survivors keep their order.  The
        // dissolved faces are only LISTED here; their edges are pushed in a second loop, so the warp
        // stays converged over the face scan.
        const int nf_old = nf;
        int keep = 0, nvis = 0;
        float4 nd_next = E.fnd[0];
        for (int i = 0; i < nf; ++i) {
            const float4 nd = nd_next;
            const uint32_t f = __float_as_uint(nd.w);
            if (i + 1 < nf) nd_next = E.fnd[i + 1];
            const vec3 pa = epa_P_face<AS, BS>(E, S, f);
            const float d = dot(pa, V3(nd));
            const vec3 tmp = ns.P - pa;
            // dot(-n, tmp) is -dot(n, tmp) bit for bit (or both are zeros), so the flipped normal is never formed
            const float tv = dot(V3(nd), tmp);
            const bool sees = d < 0.0f ? tv < 0.0f : tv > 0.0f;
            if (sees) {
                E.vis[nvis++] = f;
            } else {
                if (keep != i) E.fnd[keep] = nd;
                epa_track_min(d, keep, cur, ci, cneg);
                ++keep;
            }
        }
        nf = keep;
        for (int j = 0; j < nvis; ++j) {
            uint32_t f = E.vis[j] & 0xffffffu;      // the three vertex indices (the support byte is not rotated)
            // the three vertices' classes, fetched once per face (every vertex is on two of its edges)
            uint32_t cl = epa_cid(E, f & 255) | (epa_cid(E, (f >> 8) & 255) << 8) | (epa_cid(E, (f >> 16) & 255) << 16);
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {           // AB, BC, CA
                epa_push_edge(E, ne, f & 255, (f >> 8) & 255, cl & 255, (cl >> 8) & 255, ovf);
                f = (f >> 8) | ((f & 255) << 16);
                cl = (cl >> 8) | ((cl & 255) << 16);
            }
        }
        // one new face per horizon edge, in edge-list order (:894-901)
        if (nf + ne > kEpaMaxFaces) { ovf |= OVF_EPA_FACES; return 0; }
        const uint32_t sbn = epa_sup_byte<AS, BS>(ns);
        for (int i = 0; i < ne; ++i) {
            const uint32_t ed = E.edge[i];
            epa_push_face<AS, BS>(E, S, nf, nv, ed & 255, (ed >> 8) & 255, ns.P, sbn, cur, ci, cneg);
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
            const float dl = dot(epa_P_face<AS, BS>(E, S, __float_as_uint(last.w)), V3(last));
            cur = fabsf(dl);
            cneg = dl < 0.0f;
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
