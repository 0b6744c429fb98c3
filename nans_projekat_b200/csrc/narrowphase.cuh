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

#ifndef NANS_NP_THREADS
#define NANS_NP_THREADS 128
#endif
constexpr int kNpThreads = NANS_NP_THREADS;
constexpr int kNoVertex = 8;

#ifndef NANS_NP_OPT_C
#define NANS_NP_OPT_C 0   // equivalence-class scan without early exit: +4 % (worse)
#endif   // box support when every compare failed (NaN direction): vec3(0)

#ifndef NANS_NP_V4
#define NANS_NP_V4 0      // 1: box vertices in shared memory as one float4 per vertex (LDS.128); measured 7-15 % slower
#endif

#if NANS_NP_V4
// Both shapes' box vertices, one float4 (x, y, z, -) per vertex: g_np_verts4[(8*side + k) * kNpThreads + tid].
// A warp's LDS.128 of the same vertex is contiguous (conflict free) and a vertex fetch is one instruction.
// File-scope __shared__ so every access compiles to LDS/STS (a pointer carried through the call chain
// degrades to generic loads).
__shared__ float4 g_np_verts4[16 * kNpThreads];
constexpr int kNpSmemBytes = 16 * 16 * kNpThreads;
#else
// Both shapes' box vertices, transposed: g_np_verts[(24*side + 3*k + r) * kNpThreads + tid].
__shared__ float g_np_verts[48 * kNpThreads];
#endif

// Per-thread view of the two shapes.
struct NpShapes {
    vec3 posA, posB;     // body centres (GJK start direction; sphere support)
    vec3 dir0;           // normalize(posB - posA): EvolveSimplex recomputes it every call (:575); hoisted
    float radA, radB;    // spheres
    // box vertex k of a side; k = kNoVertex (every support compare failed) is vec3(0)
    __device__ __forceinline__ vec3 vertex(int side, int k) const
    {
#if NANS_NP_V4
        const float4 v = g_np_verts4[(8 * side + (k & 7)) * kNpThreads + threadIdx.x];
        return k >= 8 ? V3(0.f, 0.f, 0.f) : V3(v.x, v.y, v.z);
#else
        if (k >= 8) return V3(0.f, 0.f, 0.f);
        const float *p = g_np_verts + (24 * side + 3 * k) * kNpThreads + threadIdx.x;
        return V3(p[0], p[kNpThreads], p[2 * kNpThreads]);
#endif
    }
    __device__ __forceinline__ float vertex_x(int side, int k) const
    {
#if NANS_NP_V4
        return k >= 8 ? 0.f : g_np_verts4[(8 * side + (k & 7)) * kNpThreads + threadIdx.x].x;
#else
        return k >= 8 ? 0.f : g_np_verts[(24 * side + 3 * (k & 7)) * kNpThreads + threadIdx.x];
#endif
    }
    // stage one box: v24 = its 8 packed vec3 (reference vertex order)
    __device__ __forceinline__ static void store_box(int side, const float (&v24)[24])
    {
#if NANS_NP_V4
#pragma unroll
        for (int k = 0; k < 8; ++k)
            g_np_verts4[(8 * side + k) * kNpThreads + threadIdx.x] = make_float4(v24[3 * k], v24[3 * k + 1], v24[3 * k + 2], 0.f);
#else
#pragma unroll
        for (int q = 0; q < 24; ++q) g_np_verts[(24 * side + q) * kNpThreads + threadIdx.x] = v24[q];
#endif
    }
};

// GetCubeSupport / GetFloorSupport, code/nans.cpp:410-430,441-461: first vertex with strictly
// greater dot; vec3(0) if every compare fails (NaN direction).  Keeps (best, index) only while
// scanning and fetches the winner afterwards.
#ifndef NANS_NP_TREE_SUPPORT
#define NANS_NP_TREE_SUPPORT 0   // 1: tournament instead of the reference's sequential scan (same result); measured slower: 0.943 vs 0.902 ms
#endif
__device__ __forceinline__ vec3 box_support(const NpShapes &S, int side, vec3 d, int &idx)
{
    float dist[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#if NANS_NP_V4
        const float4 c4 = g_np_verts4[(8 * side + k) * kNpThreads + threadIdx.x];
        const vec3 c = V3(c4.x, c4.y, c4.z);
#else
        const float *p = g_np_verts + 24 * side * kNpThreads + threadIdx.x;
        const vec3 c = V3(p[(3 * k) * kNpThreads], p[(3 * k + 1) * kNpThreads], p[(3 * k + 2) * kNpThreads]);
#endif
        dist[k] = dot(c, d);
    }
#if NANS_NP_TREE_SUPPORT
    // "first vertex whose dot is strictly greater than everything before it, starting from -FLT_MAX" is the
    // lowest-index maximum over the dots that compare greater than -FLT_MAX (a NaN dot never does); as a
    // tournament the dependent chain is 3 compares deep instead of 8.  b wins only if strictly greater, so
    // ties keep the lower index; an invalid entry is (-FLT_MAX, 8) and loses to every valid one.
    float v[8]; int ix[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { const bool ok = dist[k] > -FLT_MAX; v[k] = ok ? dist[k] : -FLT_MAX; ix[k] = ok ? k : kNoVertex; }
#pragma unroll
    for (int w = 1; w < 8; w <<= 1)
#pragma unroll
        for (int k = 0; k < 8; k += 2 * w) {
            const bool take = v[k + w] > v[k];
            v[k] = take ? v[k + w] : v[k];
            ix[k] = take ? ix[k + w] : ix[k];
        }
    idx = ix[0];
#else
    float best = -FLT_MAX;
    idx = kNoVertex;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (dist[k] > best) { best = dist[k]; idx = k; }
#endif
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
struct EpaGenericArena {
    vec3 P[kEpaMaxVerts];
    vec3 SA[kEpaMaxVerts], SB[kEpaMaxVerts];    // sphere sides only
    uint8_t ia[kEpaMaxVerts], ib[kEpaMaxVerts]; // box sides only
    uint8_t cid[kEpaMaxVerts];                  // lowest vertex index with an equal P (kCidNaN: equal to nothing)
    float4 fnd[kEpaMaxFaces];
    uint32_t fidx[kEpaMaxFaces];
    uint32_t vis[kEpaMaxFaces];                 // packed indices of the faces dissolved this iteration
    uint32_t edge[kEpaMaxEdges];                // a | b<<8 | cid[a]<<16 | cid[b]<<24
};
// Box-box pairs (the bulk of every pile / drop world): a polytope vertex is fully described by the two box
// vertex indices it came from (P = vertsA[ia] - vertsB[ib], both in shared memory), so the arena holds no
// per-vertex arrays at all and nothing in the face scan is a dependent local-memory load.  A face/edge corner
// is 16 bits: code = ia | ib<<4 (4 bits each: 8 = "no vertex", vec3(0)) and the vertex's equivalence class.
struct EpaBoxArena {
    float4 fnd[kEpaMaxFaces];                   // flipped unit normal (PushTriangle's N), |dot(n, A.P)|
    uint2 fcr[kEpaMaxFaces];                    // x = code0 | code1<<8 | code2<<16, y = cid0 | cid1<<8 | cid2<<16
    uint2 vis[kEpaMaxFaces];                    // corners of the faces dissolved this iteration
    uint32_t edge[kEpaMaxEdges];                // codeA | codeB<<8 | cidA<<16 | cidB<<24
    uint32_t vcw[kEpaMaxVerts / 4];             // vertex codes in vertex order, 4 per word (equivalence-class scan)
};
union EpaArena {
    EpaGenericArena g;
    EpaBoxArena b;
};
constexpr int kCidNaN = 254;

// The reference's edge cancels an opposite-winding edge BY VALUE of P (code/nans.h:251-254).  Float
// equality is an equivalence on non-NaN vectors (+0 == -0 included), so every vertex gets the lowest
// index of its class once, when it is stored, and the edge compares become one integer compare.
template <bool AS, bool BS>
__device__ __forceinline__ void epa_store_vertex(EpaGenericArena &E, int i, const GjkVertex<AS, BS> &v)
{
    E.P[i] = v.P;
    if constexpr (AS) E.SA[i] = v.a.v; else E.ia[i] = (uint8_t)v.a.idx;
    if constexpr (BS) E.SB[i] = v.b.v; else E.ib[i] = (uint8_t)v.b.idx;
    int c = i;
    if (!equal(v.P, v.P)) {
        c = kCidNaN;
    } else {
#if NANS_NP_OPT_C
        for (int j = i - 1; j >= 0; --j)          // no early exit: independent loads
            if (equal(E.P[j], v.P)) c = j;
#else
        for (int j = 0; j < i; ++j)
            if (equal(E.P[j], v.P)) { c = j; break; }
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
    const vec3 n = normalize(cross(E.P[b] - pa, E.P[c] - pa));
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
    while (i < ne && (E.edge[i] >> 16) != want) ++i;
    if (i < ne) {
        for (int k = i; k < ne - 1; ++k) E.edge[k] = E.edge[k + 1];
        --ne;
        return;
    }
    if (ne >= kEpaMaxEdges) { ovf |= OVF_EPA_EDGES; return; }
    E.edge[ne++] = (uint32_t)a | ((uint32_t)b << 8) | (ca << 16) | (cb << 24);
}

// The polytope between two iterations of ResolveCollision: everything else in the arena is scratch of one iteration.
struct EpaProgress { int nv, nf, ci, it; float cur; };
constexpr int kEpaPaused = 2;     // epa_loop: 0 = no collision, 1 = collision (outputs filled), 2 = stopped at it_stop

// The iteration loop of ResolveCollision (code/nans.cpp:805-903) from the state g.  kCanPause: stop (state in g,
// arena consistent) once it_stop iterations are done, so that a long pair can be carried to another kernel.
template <bool AS, bool BS, bool kCanPause>
__device__ __forceinline__ int epa_loop(const NpShapes &S, EpaGenericArena &E, EpaProgress &g, int it_stop,
                                        vec3 &outPA, vec3 &outPB, vec3 &outN, int &ovf, int &max_faces)
{
    int nv = g.nv, nf = g.nf, ne = 0, ci = g.ci, it = g.it;
    float cur = g.cur;
    while (true) {
        if (kCanPause && it >= it_stop) { g.nv = nv; g.nf = nf; g.ci = ci; g.it = it; g.cur = cur; return kEpaPaused; }
        if (!(it++ <= 64)) break;   // MAX_EPA_ITERATIONS, code/nans.h:56: while (it++ <= 64)
        max_faces = max(max_faces, nf);
        const float4 cnd = E.fnd[ci];
        const vec3 N = face_normal_flipped(cnd);
        const GjkVertex<AS, BS> ns = calc_support<AS, BS>(S, N);
        if (fsub(dot(N, ns.P), cur) < 0.001f) {   // MAX_EPA_ERROR, code/nans.h:55
            const uint32_t f = E.fidx[ci];
            const int a = f & 255, b = (f >> 8) & 255, c = (f >> 16) & 255;
            // Barycentric, code/nans.cpp:772-785
            const vec3 Pp = N * cur;
            const vec3 A0 = E.P[a];
            const vec3 v0 = E.P[b] - A0, v1 = E.P[c] - A0, v2 = Pp - A0;
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
        int keep = 0, nvis = 0;
        float4 nd_next = E.fnd[0];
        uint32_t f_next = E.fidx[0];
        for (int i = 0; i < nf; ++i) {
            const float4 nd = nd_next;
            const uint32_t f = f_next;
            if (i + 1 < nf) { nd_next = E.fnd[i + 1]; f_next = E.fidx[i + 1]; }
            const vec3 tmp = ns.P - E.P[f & 255];
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
    }
    return 0;
}

// the start of ResolveCollision (:791-802): the simplex becomes the first four vertices and faces
template <bool AS, bool BS>
__device__ __forceinline__ void epa_start(const GjkVertex<AS, BS> (&s)[4], EpaGenericArena &E, EpaProgress &g)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) epa_store_vertex<AS, BS>(E, k, s[k]);
    int nf = 0, ci = 0;
    float cur = 0.f;
    epa_push_face(E, nf, 0, 1, 2, s[0].P, cur, ci);  // ABC
    epa_push_face(E, nf, 0, 2, 3, s[0].P, cur, ci);  // ACD
    epa_push_face(E, nf, 0, 3, 1, s[0].P, cur, ci);  // ADB
    epa_push_face(E, nf, 1, 3, 2, s[1].P, cur, ci);  // BDC
    g.nv = 4; g.nf = nf; g.ci = ci; g.it = 0; g.cur = cur;
}

// ResolveCollision, code/nans.cpp:788-904.  Returns the bool32 result; fills PointA/PointB/N.
template <bool AS, bool BS>
__device__ __forceinline__ int epa_resolve(const NpShapes &S, const GjkVertex<AS, BS> (&s)[4], EpaGenericArena &E,
                                           vec3 &outPA, vec3 &outPB, vec3 &outN, int &ovf, int &max_faces)
{
    EpaProgress g;
    epa_start<AS, BS>(s, E, g);
    return epa_loop<AS, BS, false>(S, E, g, 0, outPA, outPB, outN, ovf, max_faces);
}

// ---- carrying a paused polytope to another thread -------------------------------------------------------------
// A paused EPA (epa_loop with kCanPause) is nv vertices and nf faces; epa_save / epa_restore move exactly that
// between a per-thread arena and a global record of kCarryQuads float4, so that a long pair could be finished by
// another kernel next to pairs of its own length without repeating an iteration.  Host-checked (tests/test_np_host.py,
// "carry": pause, save, wipe the arena, restore, finish).  A kernel pair built on it for config C3 (pause after 6 / 8 /
// 10 iterations, paused polytopes to a pool, a second EPA kernel over them) measured 24.2 / 23.9 / 23.3 ms against
// 23.7 ms without and failed one GPU parity test when the GPU budget of the round ran out; it is not in the tree.
constexpr int kCarryVerts = 16, kCarryFaces = 32;
constexpr int kCarryQuads = 2 + 3 * kCarryVerts + kCarryFaces + kCarryFaces / 4;   // 90 float4 = 1440 B

__device__ __forceinline__ bool epa_can_carry(const EpaProgress &g) { return g.nv <= kCarryVerts && g.nf <= kCarryFaces; }

template <bool AS, bool BS>
__device__ __forceinline__ void epa_save(const EpaGenericArena &E, const EpaProgress &g, float4 *r)
{
    r[0] = make_float4(__int_as_float(g.nv), __int_as_float(g.nf), __int_as_float(g.ci), __int_as_float(g.it));
    r[1] = make_float4(g.cur, 0.f, 0.f, 0.f);
    for (int i = 0; i < g.nv; ++i) {
        uint32_t pk = E.cid[i];
        if constexpr (!AS) pk |= (uint32_t)E.ia[i] << 8;
        if constexpr (!BS) pk |= (uint32_t)E.ib[i] << 16;
        r[2 + 3 * i] = make_float4(E.P[i].x, E.P[i].y, E.P[i].z, __int_as_float((int)pk));
        if constexpr (AS) r[3 + 3 * i] = make_float4(E.SA[i].x, E.SA[i].y, E.SA[i].z, 0.f);
        if constexpr (BS) r[4 + 3 * i] = make_float4(E.SB[i].x, E.SB[i].y, E.SB[i].z, 0.f);
    }
    float4 *f = r + 2 + 3 * kCarryVerts;
    for (int i = 0; i < g.nf; ++i) f[i] = E.fnd[i];
    float4 *fi = f + kCarryFaces;
    for (int i = 0; i < g.nf; i += 4)
        fi[i >> 2] = make_float4(__int_as_float((int)E.fidx[i]), __int_as_float((int)(i + 1 < g.nf ? E.fidx[i + 1] : 0u)),
                                 __int_as_float((int)(i + 2 < g.nf ? E.fidx[i + 2] : 0u)),
                                 __int_as_float((int)(i + 3 < g.nf ? E.fidx[i + 3] : 0u)));
}

template <bool AS, bool BS>
__device__ __forceinline__ void epa_restore(EpaGenericArena &E, EpaProgress &g, const float4 *r)
{
    const float4 h = r[0];
    g.nv = __float_as_int(h.x); g.nf = __float_as_int(h.y); g.ci = __float_as_int(h.z); g.it = __float_as_int(h.w);
    g.cur = r[1].x;
    for (int i = 0; i < g.nv; ++i) {
        const float4 v = r[2 + 3 * i];
        const uint32_t pk = (uint32_t)__float_as_int(v.w);
        E.P[i] = V3(v);
        E.cid[i] = (uint8_t)(pk & 255u);
        if constexpr (!AS) E.ia[i] = (uint8_t)((pk >> 8) & 255u);
        if constexpr (!BS) E.ib[i] = (uint8_t)((pk >> 16) & 255u);
        if constexpr (AS) E.SA[i] = V3(r[3 + 3 * i]);
        if constexpr (BS) E.SB[i] = V3(r[4 + 3 * i]);
    }
    const float4 *f = r + 2 + 3 * kCarryVerts;
    for (int i = 0; i < g.nf; ++i) E.fnd[i] = f[i];
    const float4 *fi = f + kCarryFaces;
    for (int i = 0; i < g.nf; ++i) {
        const float4 q = fi[i >> 2];
        const float w = (i & 3) == 0 ? q.x : (i & 3) == 1 ? q.y : (i & 3) == 2 ? q.z : q.w;
        E.fidx[i] = (uint32_t)__float_as_int(w);
    }
}

// ---- EPA as a resumable state machine ------------------------------------------------------------
// The same algorithm cut at the iteration boundary, so that a lane can finish one pair and start the next
// while its neighbours are in the middle of theirs (epa_refill_kernel): a step = "create the faces the
// previous step (or the start) left pending, then one iteration up to the horizon edges".  The pending
// faces are vertex triples in E.edge: the start leaves the four tetrahedron faces ABC, ACD, ADB, BDC
// (:799-802), an iteration leaves (new vertex, e.A, e.B) per horizon edge (:894-901).  Face creation is ONE
// loop for both, so a lane that has just started shares its instructions with lanes in mid-flight.
struct EpaState {
    int nv, nf, ne, ci, it;
    int newv;      // >= 0: E.edge holds horizon edges of vertex newv; < 0: E.edge holds explicit triples
    float cur;
};
enum { kEpaContinue = 2 };   // epa_step: 0 = no collision, 1 = collision (outputs filled), 2 = call again

template <bool AS, bool BS>
__device__ __forceinline__ void epa_begin(const GjkVertex<AS, BS> (&s)[4], EpaGenericArena &E, EpaState &st)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) epa_store_vertex<AS, BS>(E, k, s[k]);
    E.edge[0] = 0u | (1u << 8) | (2u << 16);   // ABC
    E.edge[1] = 0u | (2u << 8) | (3u << 16);   // ACD
    E.edge[2] = 0u | (3u << 8) | (1u << 16);   // ADB
    E.edge[3] = 1u | (3u << 8) | (2u << 16);   // BDC
    st.nv = 4; st.nf = 0; st.ne = 4; st.ci = 0; st.it = 0; st.newv = -1; st.cur = 0.f;
}

template <bool AS, bool BS>
__device__ __forceinline__ int epa_step(const NpShapes &S, EpaGenericArena &E, EpaState &st,
                                        vec3 &outPA, vec3 &outPB, vec3 &outN, int &ovf, int &max_faces)
{
    int nf = st.nf, ci = st.ci;
    float cur = st.cur;
    // pending faces, in list order
    if (nf + st.ne > kEpaMaxFaces) { ovf |= OVF_EPA_FACES; return 0; }
    for (int i = 0; i < st.ne; ++i) {
        const uint32_t ed = E.edge[i];
        const bool tri = st.newv < 0;
        const int a = tri ? (int)(ed & 255) : st.newv;
        const int b = tri ? (int)((ed >> 8) & 255) : (int)(ed & 255);
        const int c = tri ? (int)((ed >> 16) & 255) : (int)((ed >> 8) & 255);
        epa_push_face(E, nf, a, b, c, E.P[a], cur, ci);
    }
    st.ne = 0;
    if (st.it++ > 64) { st.nf = nf; return 0; }   // MAX_EPA_ITERATIONS, code/nans.h:56: while (it++ <= 64)
    max_faces = max(max_faces, nf);
    const float4 cnd = E.fnd[ci];
    const vec3 N = face_normal_flipped(cnd);
    const GjkVertex<AS, BS> ns = calc_support<AS, BS>(S, N);
    if (fsub(dot(N, ns.P), cur) < 0.001f) {   // MAX_EPA_ERROR, code/nans.h:55
        const uint32_t f = E.fidx[ci];
        const int a = f & 255, b = (f >> 8) & 255, c = (f >> 16) & 255;
        // Barycentric, code/nans.cpp:772-785
        const vec3 Pp = N * cur;
        const vec3 A0 = E.P[a];
        const vec3 v0 = E.P[b] - A0, v1 = E.P[c] - A0, v2 = Pp - A0;
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
    if (st.nv >= kEpaMaxVerts) { ovf |= OVF_EPA_FACES; return 0; }
    epa_store_vertex<AS, BS>(E, st.nv, ns);
    // dissolve every face the new point can see (:869-891); survivors keep their order
    int keep = 0, nvis = 0, ne = 0;
    float4 nd_next = E.fnd[0];
    uint32_t f_next = E.fidx[0];
    for (int i = 0; i < nf; ++i) {
        const float4 nd = nd_next;
        const uint32_t f = f_next;
        if (i + 1 < nf) { nd_next = E.fnd[i + 1]; f_next = E.fidx[i + 1]; }
        const vec3 tmp = ns.P - E.P[f & 255];
        if (dot(face_normal_flipped(nd), tmp) > 0.0f) {
            E.vis[nvis++] = f;
        } else {
            if (keep != i) { E.fnd[keep] = nd; E.fidx[keep] = f; }
            epa_track_min(nd.w, keep, cur, ci);
            ++keep;
        }
    }
    for (int j = 0; j < nvis; ++j) {
        uint32_t f = E.vis[j];
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {           // AB, BC, CA
            epa_push_edge(E, ne, f & 255, (f >> 8) & 255, ovf);
            f = (f >> 8) | ((f & 255) << 16);
        }
    }
    st.nf = keep; st.ne = ne; st.ci = ci; st.cur = cur;
    st.newv = st.nv;
    ++st.nv;
    return kEpaContinue;
}

// ---- EPA, box-box specialisation ---------------------------------------------------------------
// Same algorithm and the same list orders as epa_resolve above; only where the data lives differs.
#ifndef NANS_NP_BOX_EPA
#define NANS_NP_BOX_EPA 0   // 1: DRAM traffic of the kernel halves (914 -> 444 MB), run time +7-10 % (more LDS/ALU work)
#endif

__device__ __forceinline__ vec3 box_corner_P(const NpShapes &S, uint32_t code)
{
    return S.vertex(0, code & 15) - S.vertex(1, (code >> 4) & 15);   // CalculateSupport's P = SupA - SupB
}

// appends the vertex code and returns the corner descriptor code | cid<<8 of polytope vertex i
__device__ __forceinline__ uint32_t epa_box_store_vertex(EpaBoxArena &E, const NpShapes &S, int i, vec3 P, uint32_t code)
{
    uint32_t c = (uint32_t)i;
    if (!equal(P, P)) {
        c = kCidNaN;
    } else {
        // lowest earlier vertex with an equal P (by value: different box-vertex pairs can give the same point);
        // x decides almost always, y and z are only formed when x matches
        for (int j0 = 0; j0 < i; j0 += 4) {
            uint32_t wv = E.vcw[j0 >> 2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = j0 + k;
                const uint32_t cj = wv & 255;
                wv >>= 8;
                if (j < i && c == (uint32_t)i) {
                    if (fsub(S.vertex_x(0, cj & 15), S.vertex_x(1, cj >> 4)) == P.x && equal(box_corner_P(S, cj), P))
                        c = (uint32_t)j;
                }
            }
        }
    }
    const int w = i >> 2, sh = 8 * (i & 3);
    const uint32_t old = sh ? E.vcw[w] : 0u;
    E.vcw[w] = old | (code << sh);
    return code | (c << 8);
}

__device__ __forceinline__ void epa_box_push_face(EpaBoxArena &E, int &nf, uint32_t c0, uint32_t c1, uint32_t c2,
                                                  vec3 pa, vec3 pb, vec3 pc, float &cur, int &ci)
{
    // PushTriangle, code/nans.cpp:293-322 (flip folded into the sign of d)
    const vec3 n = normalize(cross(pb - pa, pc - pa));
    const float d = dot(pa, n);
    // stored FLIPPED (PushTriangle's N, :316-320) with |d|: every later use wants exactly these two
    const vec3 nfl = d < 0.0f ? n * -1.0f : n;
    E.fnd[nf] = make_float4(nfl.x, nfl.y, nfl.z, fabsf(d));
    E.fcr[nf] = make_uint2((c0 & 255) | ((c1 & 255) << 8) | ((c2 & 255) << 16),
                           (c0 >> 8) | ((c1 >> 8) << 8) | ((c2 >> 8) << 16));
    epa_track_min(d, nf, cur, ci);
    ++nf;
}

// PushEdge, code/nans.cpp:233-266, on corner descriptors
__device__ __forceinline__ void epa_box_push_edge(EpaBoxArena &E, int &ne, uint32_t codeA, uint32_t codeB,
                                                  uint32_t ca, uint32_t cb, int &ovf)
{
    const uint32_t want = (cb == kCidNaN ? 255u : cb) | ((ca == kCidNaN ? 255u : ca) << 8);
    int i = 0;
    while (i < ne && (E.edge[i] >> 16) != want) ++i;
    if (i < ne) {
        for (int k = i; k < ne - 1; ++k) E.edge[k] = E.edge[k + 1];
        --ne;
        return;
    }
    if (ne >= kEpaMaxEdges) { ovf |= OVF_EPA_EDGES; return; }
    E.edge[ne++] = codeA | (codeB << 8) | (ca << 16) | (cb << 24);
}

// ResolveCollision, code/nans.cpp:788-904, for two boxes
__device__ __forceinline__ int epa_resolve_box(const NpShapes &S, const GjkVertex<false, false> (&s)[4], EpaBoxArena &E,
                                               vec3 &outPA, vec3 &outPB, vec3 &outN, int &ovf, int &max_faces)
{
    uint32_t c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        c[k] = epa_box_store_vertex(E, S, k, s[k].P, (uint32_t)s[k].a.idx | ((uint32_t)s[k].b.idx << 4));
    int nv = 4, nf = 0, ne = 0, ci = 0;
    float cur = 0.f;
    epa_box_push_face(E, nf, c[0], c[1], c[2], s[0].P, s[1].P, s[2].P, cur, ci);  // ABC
    epa_box_push_face(E, nf, c[0], c[2], c[3], s[0].P, s[2].P, s[3].P, cur, ci);  // ACD
    epa_box_push_face(E, nf, c[0], c[3], c[1], s[0].P, s[3].P, s[1].P, cur, ci);  // ADB
    epa_box_push_face(E, nf, c[1], c[3], c[2], s[1].P, s[3].P, s[2].P, cur, ci);  // BDC
    int it = 0;
    while (it++ <= 64) {            // MAX_EPA_ITERATIONS, code/nans.h:56
        max_faces = max(max_faces, nf);
        const float4 cnd = E.fnd[ci];
        const vec3 N = V3(cnd);
        const GjkVertex<false, false> ns = calc_support<false, false>(S, N);
        if (fsub(dot(N, ns.P), cur) < 0.001f) {   // MAX_EPA_ERROR, code/nans.h:55
            const uint32_t f = E.fcr[ci].x;
            const uint32_t k0 = f & 255, k1 = (f >> 8) & 255, k2 = (f >> 16) & 255;
            // Barycentric, code/nans.cpp:772-785
            const vec3 Pp = N * cur;
            const vec3 A0 = box_corner_P(S, k0);
            const vec3 v0 = box_corner_P(S, k1) - A0, v1 = box_corner_P(S, k2) - A0, v2 = Pp - A0;
            const float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1);
            const float d20 = dot(v2, v0), d21 = dot(v2, v1);
            const float denom = fsub(fmul(d00, d11), fmul(d01, d01));
            const float bv = fdiv(fsub(fmul(d11, d20), fmul(d01, d21)), denom);
            const float bw = fdiv(fsub(fmul(d00, d21), fmul(d01, d20)), denom);
            const float bu = fsub(fsub(1.0f, bv), bw);
            if (fabsf(bu) > 1.0f || fabsf(bv) > 1.0f || fabsf(bw) > 1.0f) return 0;
            if (!isfinite(bu) || !isfinite(bv) || !isfinite(bw)) return 0;   // IsValid, :4-17
            outPA = ((bu * S.vertex(0, k0 & 15)) + (bv * S.vertex(0, k1 & 15))) + (bw * S.vertex(0, k2 & 15));
            outN = -1.0f * N;
            outPB = ((bu * S.vertex(1, k0 >> 4)) + (bv * S.vertex(1, k1 >> 4))) + (bw * S.vertex(1, k2 >> 4));
            return 1;
        }
        if (nv >= kEpaMaxVerts) { ovf |= OVF_EPA_FACES; return 0; }
        const uint32_t cn = epa_box_store_vertex(E, S, nv, ns.P, (uint32_t)ns.a.idx | ((uint32_t)ns.b.idx << 4));
        // dissolve every face the new point can see (:869-891); survivors keep their order
        int keep = 0, nvis = 0;
        float4 nd_next = E.fnd[0];
        uint2 f_next = E.fcr[0];
        for (int i = 0; i < nf; ++i) {
            const float4 nd = nd_next;
            const uint2 f = f_next;
            if (i + 1 < nf) { nd_next = E.fnd[i + 1]; f_next = E.fcr[i + 1]; }
            const vec3 tmp = ns.P - box_corner_P(S, f.x & 255);
            if (dot(V3(nd), tmp) > 0.0f) {
                E.vis[nvis++] = f;
            } else {
                if (keep != i) { E.fnd[keep] = nd; E.fcr[keep] = f; }
                epa_track_min(nd.w, keep, cur, ci);
                ++keep;
            }
        }
        nf = keep;
        for (int j = 0; j < nvis; ++j) {
            uint2 f = E.vis[j];
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {           // AB, BC, CA
                epa_box_push_edge(E, ne, f.x & 255, (f.x >> 8) & 255, f.y & 255, (f.y >> 8) & 255, ovf);
                f.x = (f.x >> 8) | ((f.x & 255) << 16);
                f.y = (f.y >> 8) | ((f.y & 255) << 16);
            }
        }
        // one new face per horizon edge, in edge-list order (:894-901)
        if (nf + ne > kEpaMaxFaces) { ovf |= OVF_EPA_FACES; return 0; }
        for (int i = 0; i < ne; ++i) {
            const uint32_t ed = E.edge[i];
            const uint32_t ka = ed & 255, kb = (ed >> 8) & 255;
            epa_box_push_face(E, nf, cn, ka | ((ed >> 16) & 255) << 8, kb | (ed >> 24) << 8,
                              ns.P, box_corner_P(S, ka), box_corner_P(S, kb), cur, ci);
        }
        ne = 0;
        ++nv;
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
#ifdef NANS_NP_SKIP_EPA   // timing experiment only: GJK without EPA
    if (false) {
#else
    if (ev == kFoundIntersection) {
#endif
#if NANS_NP_BOX_EPA
        if constexpr (!AS && !BS) r.hit = epa_resolve_box(S, s, E.b, r.PA, r.PB, r.N, ovf, max_faces);
        else
#endif
#ifdef NANS_NP_STEPPED   // the resumable form run to completion (host check of epa_begin / epa_step)
        {
            EpaState st;
            epa_begin<AS, BS>(s, E.g, st);
            int rr;
            do rr = epa_step<AS, BS>(S, E.g, st, r.PA, r.PB, r.N, ovf, max_faces); while (rr == kEpaContinue);
            r.hit = rr;
        }
#elif defined(NANS_NP_CARRY)   // pause after NANS_NP_CARRY iterations, move the polytope through a record into a
        {                          // wiped arena and finish there (host check of epa_loop / epa_save / epa_restore)
            EpaProgress g;
            epa_start<AS, BS>(s, E.g, g);
            int rr = epa_loop<AS, BS, true>(S, E.g, g, NANS_NP_CARRY, r.PA, r.PB, r.N, ovf, max_faces);
            if (rr == kEpaPaused) {
                if (epa_can_carry(g)) {
                    static float4 rec[kCarryQuads];
                    epa_save<AS, BS>(E.g, g, rec);
                    memset(&E, 0xAB, sizeof(E));
                    g.nv = g.nf = g.ci = g.it = -1; g.cur = -1.f;
                    epa_restore<AS, BS>(E.g, g, rec);
                }
                rr = epa_loop<AS, BS, false>(S, E.g, g, 0, r.PA, r.PB, r.N, ovf, max_faces);
            }
            r.hit = rr;
        }
#else
            r.hit = epa_resolve<AS, BS>(S, s, E.g, r.PA, r.PB, r.N, ovf, max_faces);
#endif
    }
    return r;
}

}  // namespace nans
