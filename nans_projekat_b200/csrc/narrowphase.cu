// narrowphase.cu — kernels around narrowphase.cuh.
//
//   narrowphase_world_gjk_kernel, narrowphase_world_epa_kernel
//                              the world's candidate pairs of a cube-only world (the step): GJK alone (capped at 6
//                              evolutions) with the intersecting pairs appended to a list, then EPA over whole
//                              chunks of that list, the few GJK stragglers finished first.  One pair per thread;
//                              warps fetch 32-entry chunks from a global counter, grid = SM count x resident CTAs.
//   narrowphase_world_kernel   the same in ONE kernel (GJK + EPA per pair): worlds with spheres, where nearly every
//                              candidate intersects and the simplex does not reduce to eight vertex indices.
//   gjk_split_kernel, gjk_continue_kernel, epa_refill_kernel
//                              the same over stand-alone shape pairs (nans_check_collision_batch/_device, config
//                              C3), where half the pairs miss and GJK / EPA lengths are long-tailed: GJK capped,
//                              stragglers and intersecting pairs compacted into lists, EPA over the lists.
//
// Not HBM bound: 216 B in + 48 B out per pair against ~1-5 kflop of unfused fp32.  GJK is bound by
// issue slots, EPA by the L1 capacity left for the per-thread polytope arenas (narrowphase.cuh).
#include <stdlib.h>

#include "narrowphase.cuh"

#ifndef NANS_NP_MINBLOCKS
#define NANS_NP_MINBLOCKS 5   // 96 registers/thread: best of the sweep in profiles/ (4: 128 regs, 6: 80 regs + spills)
#endif

#ifndef NANS_GJK_MINBLOCKS
#define NANS_GJK_MINBLOCKS 6   // GJK-only kernel of the split path (80 registers)
#endif

#ifndef NANS_EPA_MINBLOCKS
#define NANS_EPA_MINBLOCKS 5   // EPA refill kernel of the split path
#endif

#ifndef NANS_EPAW_MINBLOCKS
#define NANS_EPAW_MINBLOCKS 5   // EPA kernel of the two-kernel world narrowphase
#endif

#ifndef NANS_NP_STREAM
#define NANS_NP_STREAM 1
#endif

namespace nans {

__device__ __forceinline__ void load_box(int side, const float4 *__restrict__ v6)
{
    float v[24];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
#if NANS_NP_STREAM
        const float4 t = __ldcs(v6 + q);     // evict-first: keep the L2 for the per-thread EPA arenas
#else
        const float4 t = __ldg(v6 + q);
#endif
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    NpShapes::store_box(side, v);
}

__device__ __forceinline__ NpResult dispatch(bool a_sphere, bool b_sphere, NpShapes &S, EpaArena &E,
                                             int &ovf, int &max_faces)
{
    if (!a_sphere && !b_sphere) return check_collision<false, false>(S, E, ovf, max_faces);
    if (!a_sphere && b_sphere) return check_collision<false, true>(S, E, ovf, max_faces);
    if (a_sphere && !b_sphere) return check_collision<true, false>(S, E, ovf, max_faces);
    return check_collision<true, true>(S, E, ovf, max_faces);
}

// (Round 2 measured a CTA-level regrouping of the pairs between EPA iterations -- one iteration per round, the pairs
// still iterating compacted onto the first lanes, polytopes in a slot-addressed global pool so that a pair can change
// lanes -- to attack the 13.5 of 32 active lanes of this kernel.  Bit-exact, and slower: 0.77 vs 0.50 ms on the
// 1 M-cube pile.  Lanes per instruction only rose to 15.6: the divergence is INSIDE an iteration (how many faces the
// new point sees, how many horizon edges survive), not in the iteration counts, and the barriers cost 33 % of the
// stall samples at 20 warps per SM.  The begin / iterate split of EPA it needed also cost the explicit-pairs EPA
// kernel 12 % (config C3: 23.8 -> 26.6 ms).  profiles/r2_np_regroup_*.  The one-thread-per-pair kernel below stays.)
__device__ __forceinline__ void np_world_pair(const DeviceWorld &w, int p, int ra, int rb, NpShapes &S, EpaArena &E,
                                              int &ovf, int &max_faces, int &found)
{
    const bool a_sphere = ra >= w.n_cubes;
    const bool b_sphere = rb >= w.n_cubes;   // statics are negative -> box
    S.posA = V3(w.pos[ra]);
    S.radA = 0.f;
    if (a_sphere) S.radA = w.scale[ra].w; else load_box(0, w.verts + 6 * (size_t)ra);
    S.radB = 0.f;
    if (rb < 0) {
        const int k = -rb - 1;
        S.posB = V3(w.st_pos[k]);
        load_box(1, w.st_verts + 6 * k);
    } else {
        S.posB = V3(w.pos[rb]);
        if (b_sphere) S.radB = w.scale[rb].w; else load_box(1, w.verts + 6 * (size_t)rb);
    }
    const NpResult r = dispatch(a_sphere, b_sphere, S, E, ovf, max_faces);
    found += (r.gjk == kFoundIntersection);
    w.pair_hit[p] = r.hit;
    if (r.hit) {
        float4 *o = w.pair_out + 3 * (size_t)p;
#if NANS_NP_STREAM
        __stcs(o, make_float4(r.PA.x, r.PA.y, r.PA.z, 0.f));
        __stcs(o + 1, make_float4(r.PB.x, r.PB.y, r.PB.z, 0.f));
        __stcs(o + 2, make_float4(r.N.x, r.N.y, r.N.z, 0.f));
#else
        o[0] = make_float4(r.PA.x, r.PA.y, r.PA.z, 0.f);
        o[1] = make_float4(r.PB.x, r.PB.y, r.PB.z, 0.f);
        o[2] = make_float4(r.N.x, r.N.y, r.N.z, 0.f);
#endif
    }
}

__global__ void __launch_bounds__(kNpThreads, NANS_NP_MINBLOCKS) narrowphase_world_kernel(DeviceWorld w, int *work_counter)
{
    EpaArena E;
    const int lane = threadIdx.x & 31;
    const int n_pairs = w.counters->n_pairs;
    int ovf = 0, max_faces = 0, found = 0;
    NpShapes S;
    constexpr int kPerTicket = 32;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(work_counter, kPerTicket);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_pairs) break;
        const int p = base + lane;
        if (p < n_pairs) np_world_pair(w, p, w.pair_a[p], w.pair_b[p], S, E, ovf, max_faces, found);
    }
    // per-warp stats
    found = __reduce_add_sync(0xffffffffu, found);
    ovf = __reduce_or_sync(0xffffffffu, ovf);
    max_faces = __reduce_max_sync(0xffffffffu, max_faces);
    if (lane == 0) {
        if (found) atomicAdd(&w.counters->n_gjk_found, found);
        if (ovf) atomicOr(&w.counters->overflow, ovf);
        atomicMax(&w.counters->max_epa_faces, max_faces);
    }
}

// ---- the world narrowphase in two kernels ---------------------------------------------------------------------------
// In the one-kernel form the lanes whose pair misses idle through the EPA of their 32-pair chunk (24 % of the lanes on
// the 100^3 pile).  Here a first kernel runs GJK alone over every candidate (box pairs; a pair with a sphere still runs
// GJK + EPA in place) and appends the intersecting box pairs to ONE list -- entry = {pair, the final simplex as eight
// 4-bit box-vertex indices}: P = SupA - SupB is re-formed from the same operands, so nothing else has to be saved --
// and a second kernel runs EPA over whole 32-entry chunks of that list.  Results are written per pair, so the list's
// (atomic) order is not observable.  The list lives in the solver's schedule buffers, which are idle until the
// contact list exists (inc: 8 B x 2 x max_contacts).
constexpr int kNpListCount = 4, kNpListTicket = 5, kNpDeferCount = 6;    // Counters::pad slots (zeroed with the block per detect)
#ifndef NANS_NP_GJK_CAP_WORLD
#define NANS_NP_GJK_CAP_WORLD 6
#endif
constexpr int kNpGjkCap = NANS_NP_GJK_CAP_WORLD;   // evolutions in the GJK kernel (a pile's pairs take 5; 1000 = no cap)

__device__ __forceinline__ void np_load_world_shapes(const DeviceWorld &w, int ra, int rb, bool a_sphere, bool b_sphere, NpShapes &S)
{
    S.posA = V3(w.pos[ra]);
    S.radA = 0.f;
    if (a_sphere) S.radA = w.scale[ra].w; else load_box(0, w.verts + 6 * (size_t)ra);
    S.radB = 0.f;
    if (rb < 0) {
        const int k = -rb - 1;
        S.posB = V3(w.st_pos[k]);
        load_box(1, w.st_verts + 6 * k);
    } else {
        S.posB = V3(w.pos[rb]);
        if (b_sphere) S.radB = w.scale[rb].w; else load_box(1, w.verts + 6 * (size_t)rb);
    }
}

__device__ __forceinline__ void np_store_hit(const DeviceWorld &w, int p, const NpResult &r)
{
    w.pair_hit[p] = r.hit;
    if (r.hit) {
        float4 *o = w.pair_out + 3 * (size_t)p;
        __stcs(o, make_float4(r.PA.x, r.PA.y, r.PA.z, 0.f));
        __stcs(o + 1, make_float4(r.PB.x, r.PB.y, r.PB.z, 0.f));
        __stcs(o + 2, make_float4(r.N.x, r.N.y, r.N.z, 0.f));
    }
}

template <bool HAS_SPHERES>
__global__ void __launch_bounds__(kNpThreads, HAS_SPHERES ? NANS_NP_MINBLOCKS : NANS_GJK_MINBLOCKS) narrowphase_world_gjk_kernel(DeviceWorld w, int *work_counter)
{
    const int lane = threadIdx.x & 31;
    const int n_pairs = w.counters->n_pairs;
    uint2 *list = reinterpret_cast<uint2 *>(w.inc);
    const int list_cap = w.max_contacts;
    int ovf = 0, max_faces = 0, found = 0;
    NpShapes S;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(work_counter, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_pairs) break;
        const int p = base + lane;
        bool want = false, defer = false;
        uint32_t rec = 0;
        int state = 0;
        if (p < n_pairs) {
            const int ra = w.pair_a[p], rb = w.pair_b[p];
            const bool a_sphere = ra >= w.n_cubes, b_sphere = rb >= w.n_cubes;
            np_load_world_shapes(w, ra, rb, a_sphere, b_sphere, S);
            if (HAS_SPHERES && (a_sphere || b_sphere)) {
                if constexpr (HAS_SPHERES) {
                    EpaArena E;
                    const NpResult r = dispatch(a_sphere, b_sphere, S, E, ovf, max_faces);
                    found += (r.gjk == kFoundIntersection);
                    np_store_hit(w, p, r);
                }
            } else {
                GjkVertex<false, false> s[4];
                int n = 0, iter = 0;
                const int ev = gjk_resume<false, false>(S, s, n, iter, kNpGjkCap);
                want = ev == kFoundIntersection;
                defer = ev == kStillEvolving && iter <= 64;
                if (want || defer) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < n) rec |= ((uint32_t)s[k].a.idx | ((uint32_t)s[k].b.idx << 4)) << (8 * k);
                    state = n | (iter << 8);
                } else {
                    w.pair_hit[p] = 0;
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, want);
        if (m) {
            int at = 0;
            if (lane == 0) at = atomicAdd(&w.counters->pad[kNpListCount], __popc(m));
            at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
            if (want) {
                if (at < list_cap) list[at] = make_uint2((uint32_t)p, rec);
                else { ovf |= OVF_CONTACTS; w.pair_hit[p] = 0; }
            }
            found += want;
        }
        // the few pairs still evolving after kNpGjkCap evolutions (misses that cycle up to the reference's limit of 65:
        // 0.5-2 % of a pile's candidates, but a chunk waits for its slowest lane) are finished by the EPA kernel
        const unsigned md = __ballot_sync(0xffffffffu, defer);
        if (md) {
            int at = 0;
            if (lane == 0) at = atomicAdd(&w.counters->pad[kNpDeferCount], __popc(md));
            at = __shfl_sync(0xffffffffu, at, 0) + __popc(md & ((1u << lane) - 1u));
            if (defer) {
                if (at < list_cap) { w.succ_a[at] = p; w.succ_b[at] = (int32_t)rec; w.run_flag[at] = state; }
                else { ovf |= OVF_CONTACTS; w.pair_hit[p] = 0; }
            }
        }
    }
    found = __reduce_add_sync(0xffffffffu, found);
    ovf = __reduce_or_sync(0xffffffffu, ovf);
    max_faces = __reduce_max_sync(0xffffffffu, max_faces);
    if (lane == 0) {
        if (found) atomicAdd(&w.counters->n_gjk_found, found);
        if (ovf) atomicOr(&w.counters->overflow, ovf);
        if (max_faces) atomicMax(&w.counters->max_epa_faces, max_faces);
    }
}

__global__ void __launch_bounds__(kNpThreads, NANS_EPAW_MINBLOCKS) narrowphase_world_epa_kernel(DeviceWorld w)
{
    EpaArena E;
    NpShapes S;
    const int lane = threadIdx.x & 31;
    const uint2 *list = reinterpret_cast<const uint2 *>(w.inc);
    const int count = min(w.counters->pad[kNpListCount], w.max_contacts);
    int ovf = 0, max_faces = 0, found = 0;
    // tickets [0, nd32): the deferred GJK pairs, 32 per ticket, taken FIRST (each runs up to ~60 more evolutions: started
    // early they end inside the kernel instead of forming its tail); tickets from nd32 on: the EPA list
    const int nd = min(w.counters->pad[kNpDeferCount], w.max_contacts);
    const int nd32 = (nd + 31) & ~31;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&w.counters->pad[kNpListTicket], 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= nd32 + count) break;
        if (base < nd32) {
            const int k = base + lane;
            if (k < nd) {
                const int p = w.succ_a[k];
                const uint32_t rec = (uint32_t)w.succ_b[k];
                int n = w.run_flag[k] & 255, iter = w.run_flag[k] >> 8;
                const int ra = w.pair_a[p], rb = w.pair_b[p];
                np_load_world_shapes(w, ra, rb, false, false, S);
                GjkVertex<false, false> s[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[j].a.idx = (int)((rec >> (8 * j)) & 15u);
                    s[j].b.idx = (int)((rec >> (8 * j + 4)) & 15u);
                    s[j].P = S.vertex(0, s[j].a.idx) - S.vertex(1, s[j].b.idx);
                }
                const int ev = gjk_resume<false, false>(S, s, n, iter, 1000);
                NpResult r;
                r.gjk = ev;
                r.hit = 0;
                r.PA = r.PB = r.N = V3(0.f, 0.f, 0.f);
                if (ev == kFoundIntersection) {
                    ++found;
                    r.hit = epa_resolve<false, false>(S, s, E, r.PA, r.PB, r.N, ovf, max_faces);
                }
                np_store_hit(w, p, r);
            }
            continue;
        }
        const int k = base - nd32 + lane;
        if (k < count) {
            const uint2 e = list[k];
            const int p = (int)e.x;
            const int ra = w.pair_a[p], rb = w.pair_b[p];
            load_box(0, w.verts + 6 * (size_t)ra);
            if (rb < 0) load_box(1, w.st_verts + 6 * (-rb - 1)); else load_box(1, w.verts + 6 * (size_t)rb);
            GjkVertex<false, false> s[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[j].a.idx = (int)((e.y >> (8 * j)) & 15u);
                s[j].b.idx = (int)((e.y >> (8 * j + 4)) & 15u);
                s[j].P = S.vertex(0, s[j].a.idx) - S.vertex(1, s[j].b.idx);   // CalculateSupport's P = SupA - SupB
            }
            NpResult r;
            r.gjk = kFoundIntersection;
            r.PA = r.PB = r.N = V3(0.f, 0.f, 0.f);
            r.hit = epa_resolve<false, false>(S, s, E, r.PA, r.PB, r.N, ovf, max_faces);
            np_store_hit(w, p, r);
        }
    }
    ovf = __reduce_or_sync(0xffffffffu, ovf);
    max_faces = __reduce_max_sync(0xffffffffu, max_faces);
    found = __reduce_add_sync(0xffffffffu, found);
    if (lane == 0) {
        if (found) atomicAdd(&w.counters->n_gjk_found, found);
        if (ovf) atomicOr(&w.counters->overflow, ovf);
        atomicMax(&w.counters->max_epa_faces, max_faces);
    }
}

// ---- split narrowphase: GJK over every pair, then EPA with lane refill over the intersecting ones ----------
// In the one-kernel form a lane that misses (or converges early) idles until the slowest pair of its 32-pair
// chunk is done.  Where hits are sparse or EPA lengths spread widely (random pairs, config C3: 48 % hits, EPA
// 1-65 iterations; EPA is 89 % of that run) most lanes idle most of the time.  Here a first kernel runs GJK
// only and appends the intersecting pairs to one list per shape-type class together with their final simplex
// (128 B record: per vertex the two support points, or the two box vertex indices); a second, persistent
// kernel runs EPA as the resumable state machine of narrowphase.cuh: a lane whose pair has finished takes the
// next pair of the list at once.  Results are written per pair, so the list order (atomics) is not observable.
struct SplitScratch {
    int32_t *head;        // [cls] pairs listed for EPA, [4 + cls] tickets taken; [8 + cls], [12 + cls]: the same for glist
    int32_t *list[4];     // intersecting pairs, class = 2 * a_sphere + b_sphere
    int32_t *glist[4];    // pairs whose GJK was still evolving after NANS_GJK_CAP evolutions
    float4 *rec;          // [pairs][8]: simplex of pair p (4 x support A, 4 x support B); .w of [0], [1]: n, iter
};

struct BatchSrc {
    int n;
    const int32_t *type;
    const float4 *posrad_a, *verts_a, *posrad_b, *verts_b;
    int32_t *hit, *gjk;
    float4 *out;
    __device__ __forceinline__ int n_pairs() const { return n; }
    __device__ __forceinline__ void classify(int p, bool &as, bool &bs) const
    {
        const int t = type[p];
        as = (t == NANS_SS || t == NANS_SF);
        bs = (t == NANS_CS || t == NANS_SS);
    }
    __device__ __forceinline__ void load(int p, bool as, bool bs, NpShapes &S) const
    {
        const float4 pa = posrad_a[p], pb = posrad_b[p];
        S.posA = V3(pa); S.radA = pa.w;
        S.posB = V3(pb); S.radB = pb.w;
        if (!as) load_box(0, verts_a + 6 * (size_t)p);
        if (!bs) load_box(1, verts_b + 6 * (size_t)p);
    }
    __device__ __forceinline__ void store_gjk(int p, int ev) const
    {
        hit[p] = 0;
        if (gjk) gjk[p] = ev;
        float4 *o = out + 3 * (size_t)p;
        o[0] = o[1] = o[2] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void store_hit(int p, vec3 PA, vec3 PB, vec3 N) const
    {
        hit[p] = 1;
        float4 *o = out + 3 * (size_t)p;
        o[0] = make_float4(PA.x, PA.y, PA.z, 0.f);
        o[1] = make_float4(PB.x, PB.y, PB.z, 0.f);
        o[2] = make_float4(N.x, N.y, N.z, 0.f);
    }
};

template <bool SPHERE> __device__ __forceinline__ float4 pack_sup(const SupRec<SPHERE> &r)
{
    if constexpr (SPHERE) return make_float4(r.v.x, r.v.y, r.v.z, 0.f);
    else return make_float4(__int_as_float(r.idx), 0.f, 0.f, 0.f);
}
template <bool SPHERE> __device__ __forceinline__ vec3 unpack_sup(const NpShapes &S, int side, float4 q, SupRec<SPHERE> &r)
{
    if constexpr (SPHERE) { r.v = V3(q); return r.v; }
    else { r.idx = __float_as_int(q.x); return S.vertex(side, r.idx); }
}

#ifndef NANS_GJK_CAP
#define NANS_GJK_CAP 6     // evolutions in the first GJK kernel; 0 = run every pair to the end there
                           // (config C3, 16 Mi pairs: cap 0 / 6 / 8 / 12 -> 28.0 / 23.8 / 24.0 / 24.4 ms)
#endif

// append pair p to list[cls] (warp-aggregated over the lanes converged here, which all belong to this class)
// and save its simplex
template <bool AS, bool BS>
__device__ __forceinline__ void list_pair(int32_t *head, int32_t *list, float4 *rec, int p, bool want,
                                          const GjkVertex<AS, BS> (&s)[4], int n, int iter)
{
    const unsigned act = __activemask();
    const unsigned m = __ballot_sync(act, want);
    if (!want) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(head, __popc(m));
    base = __shfl_sync(m, base, leader);
    list[base + __popc(m & ((1u << lane) - 1u))] = p;
    float4 *r = rec + 8 * (size_t)p;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;      // entries past n are not part of the simplex
        if (k < n) { a = pack_sup<AS>(s[k].a); b = pack_sup<BS>(s[k].b); }
        if (k == 0) a.w = __int_as_float(n);
        if (k == 1) a.w = __int_as_float(iter);
        r[k] = a; r[4 + k] = b;
    }
}

// the simplex a list entry saved (P = SupA - SupB re-formed: same operands, same bits)
template <bool AS, bool BS>
__device__ __forceinline__ void load_simplex(const NpShapes &S, const float4 *rec, int p, GjkVertex<AS, BS> (&s)[4], int &n, int &iter)
{
    const float4 *r = rec + 8 * (size_t)p;
    n = __float_as_int(r[0].w); iter = __float_as_int(r[1].w);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const vec3 sa = unpack_sup<AS>(S, 0, r[j], s[j].a);
        const vec3 sb = unpack_sup<BS>(S, 1, r[4 + j], s[j].b);
        s[j].P = sa - sb;
    }
}

// GJK of one pair, at most NANS_GJK_CAP evolutions: an intersecting pair goes to its class's EPA list, a pair
// that is still evolving to its class's GJK list (GJK lengths are long-tailed too: most pairs decide in 1 or
// 5 evolutions, a few cycle up to the limit of 65, and a chunk waits for its slowest lane)
template <bool AS, bool BS, typename Src>
__device__ __noinline__ int gjk_and_list(const Src &src, const SplitScratch &sc, NpShapes &S, int p)
{
    constexpr int cls = 2 * (int)AS + (int)BS;
    GjkVertex<AS, BS> s[4];
    int n = 0, iter = 0;
    const int ev = gjk_resume<AS, BS>(S, s, n, iter, NANS_GJK_CAP > 0 ? NANS_GJK_CAP : 1000);
    list_pair<AS, BS>(sc.head + cls, sc.list[cls], sc.rec, p, ev == kFoundIntersection, s, 4, iter);
    list_pair<AS, BS>(sc.head + 8 + cls, sc.glist[cls], sc.rec, p, ev == kStillEvolving && iter <= 64, s, n, iter);
    return ev;
}

// the rest of GJK for the pairs the first kernel left evolving, over whole chunks of their (compacted) lists
template <bool AS, bool BS, typename Src>
__device__ __noinline__ void gjk_continue_list(const Src &src, const SplitScratch &sc, NpShapes &S)
{
    constexpr int cls = 2 * (int)AS + (int)BS;
    const int lane = threadIdx.x & 31;
    const int count = sc.head[8 + cls];
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(sc.head + 12 + cls, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
        const int k = base + lane;
        if (k < count) {
            const int p = sc.glist[cls][k];
            src.load(p, AS, BS, S);
            GjkVertex<AS, BS> s[4];
            int n, iter;
            load_simplex<AS, BS>(S, sc.rec, p, s, n, iter);
            const int ev = gjk_resume<AS, BS>(S, s, n, iter, 1000);
            src.store_gjk(p, ev);
            list_pair<AS, BS>(sc.head + cls, sc.list[cls], sc.rec, p, ev == kFoundIntersection, s, 4, iter);
        }
    }
}

template <typename Src>
__global__ void __launch_bounds__(kNpThreads, NANS_GJK_MINBLOCKS) gjk_continue_kernel(Src src, SplitScratch sc)
{
    NpShapes S;
    gjk_continue_list<false, false>(src, sc, S);
    gjk_continue_list<false, true>(src, sc, S);
    gjk_continue_list<true, false>(src, sc, S);
    gjk_continue_list<true, true>(src, sc, S);
}

template <typename Src>
__global__ void __launch_bounds__(kNpThreads, NANS_GJK_MINBLOCKS) gjk_split_kernel(Src src, SplitScratch sc, int *work_counter, int32_t *n_found)
{
    const int lane = threadIdx.x & 31;
    const int n_pairs = src.n_pairs();
    int found = 0;
    NpShapes S;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(work_counter, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_pairs) break;
        const int p = base + lane;
        if (p < n_pairs) {
            bool as, bs;
            src.classify(p, as, bs);
            src.load(p, as, bs, S);
            int ev;
            if (!as && !bs) ev = gjk_and_list<false, false>(src, sc, S, p);
            else if (!as && bs) ev = gjk_and_list<false, true>(src, sc, S, p);
            else if (as && !bs) ev = gjk_and_list<true, false>(src, sc, S, p);
            else ev = gjk_and_list<true, true>(src, sc, S, p);
            src.store_gjk(p, ev);
            found += ev == kFoundIntersection;
        }
    }
    if (n_found) {
        found = __reduce_add_sync(0xffffffffu, found);
        if (lane == 0 && found) atomicAdd(n_found, found);
    }
}

// EPA over one class list, whole 32-pair chunks of the COMPACTED list.  Measured on config C3 (16 Mi pairs) in round 1:
// 27.2 ms, against 29.6-35.4 ms for a resumable per-iteration form with lane-level refill (lanes at different iteration
// numbers carry polytopes of very different sizes and in lock step every lane pays for the largest) and 40.9 ms for
// GJK + EPA in one kernel.  Resident CTAs per SM 2..5 make no difference.
//
// Round 2: in PASSES with growing iteration budgets.  A chunk runs as long as its slowest lane, and EPA lengths of
// random pairs are long-tailed (cube-cube: median 6, p99 12, max 23; cube-sphere: median 9, p99 24, max 64 iterations;
// each iteration costs more than the one before it, the polytope grows): 0.37 (CC) / 0.25 (CS) of the lane-time of a
// chunk is useful.  A pass gives every pair of its list `budget` iterations; a pair that needs more is appended to
// the next pass's list and RUN AGAIN from its saved simplex there, among pairs that are all long.  Redoing the first
// iterations of the long pairs costs less than the short pairs' idle lanes (oracle iteration counts, cost model
// k (5 + k): two budgets bring the lane-time to 0.69 (CC) / 0.57 (CS) of the single pass).  Results are per pair and
// the same code runs each time, so nothing observable changes.
struct EpaPass {
    const int32_t *list;     // pairs of this pass
    const int32_t *count;    // their number (device)
    int32_t *ticket;         // chunk tickets
    int32_t *next_list;      // pairs that ran out of budget (null in the last pass)
    int32_t *next_count;
    int budget[4];           // iterations per class (65 = no limit)
};

template <bool AS, bool BS, typename Src>
__device__ __noinline__ void epa_chunk_list(const Src &src, const SplitScratch &sc, const EpaPass (&pass)[4], NpShapes &S,
                                            EpaArena &E, int &ovf, int &max_faces)
{
    constexpr int cls = 2 * (int)AS + (int)BS;
    const int lane = threadIdx.x & 31;
    const EpaPass &ps = pass[cls];
    const int32_t *list = ps.list;
    const int count = *ps.count;
    const int budget = ps.budget[cls];
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(ps.ticket, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
        const int k = base + lane;
        bool again = false;
        int p = 0;
        if (k < count) {
            p = list[k];
            src.load(p, AS, BS, S);
            GjkVertex<AS, BS> s[4];
            int n_, iter_;
            load_simplex<AS, BS>(S, sc.rec, p, s, n_, iter_);
            vec3 PA, PB, N;
            const int hit = epa_resolve<AS, BS>(S, s, E, PA, PB, N, ovf, max_faces, budget);
            if (hit == 1) src.store_hit(p, PA, PB, N);
            again = hit == kEpaOutOfBudget && budget < 65;
        }
        const unsigned m = __ballot_sync(0xffffffffu, again);
        if (m && ps.next_list) {
            int at = 0;
            if (lane == 0) at = atomicAdd(ps.next_count, __popc(m));
            at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
            if (again) ps.next_list[at] = p;
        }
    }
}

template <typename Src>
__global__ void __launch_bounds__(kNpThreads, NANS_EPA_MINBLOCKS) epa_refill_kernel(Src src, SplitScratch sc, Counters *counters,
                                                                                     EpaPass p0, EpaPass p1, EpaPass p2, EpaPass p3)
{
    EpaArena E;
    NpShapes S;
    int ovf = 0, max_faces = 0;
    const EpaPass pass[4] = {p0, p1, p2, p3};
    epa_chunk_list<false, false>(src, sc, pass, S, E, ovf, max_faces);
    epa_chunk_list<false, true>(src, sc, pass, S, E, ovf, max_faces);
    epa_chunk_list<true, false>(src, sc, pass, S, E, ovf, max_faces);
    epa_chunk_list<true, true>(src, sc, pass, S, E, ovf, max_faces);
    ovf = __reduce_or_sync(0xffffffffu, ovf);
    max_faces = __reduce_max_sync(0xffffffffu, max_faces);
    if ((threadIdx.x & 31) == 0 && counters) {
        if (ovf) atomicOr(&counters->overflow, ovf);
        atomicMax(&counters->max_epa_faces, max_faces);
    }
}

static int np_grid(int blocks_needed)
{
    static int per_sm = 0;
    if (!per_sm) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, narrowphase_world_kernel, kNpThreads, 0);
        if (per_sm < 1) per_sm = 1;
    }
    const int cap = kNumSMs * per_sm;
    return blocks_needed < cap ? (blocks_needed < 1 ? 1 : blocks_needed) : cap;
}

int launch_narrowphase(World *w)
{
    DeviceWorld &d = w->d;
    if (d.nb == 0) return NANS_OK;
    // the work counter lives in the Counters pad (zeroed with the block at the start of detect)
    int *work = &d.counters->pad[0];
    // pair count is device-resident: size the grid for the capacity, CTAs beyond the work exit at once
    const int need = div_up(d.max_pairs, kNpThreads);
    // two kernels (GJK, then EPA over the list of intersecting pairs) for a cube-only world, one kernel for a world
    // with spheres (measured: 4096 x 64-body worlds with 16 spheres each, 0.257 ms in one kernel, 0.280 in two; capping
    // the one-kernel form's box-pair GJK and deferring the stragglers the same way: 0.318, its register budget breaks;
    // the box segments of such a world through the two kernels and only the sphere pairs through the one: 0.277 vs
    // 0.239 -- 99 % of C4's candidates intersect, there are no missing lanes to win back and no cycling misses);
    // NANS_NP_SPLIT=0/1 forces either form (A/B runs)
    static int split_env = -2;
    if (split_env == -2) { const char *e = getenv("NANS_NP_SPLIT"); split_env = e ? atoi(e) : -1; }
    const bool split = split_env < 0 ? d.n_spheres == 0 : split_env != 0;
    if (!split) {
        narrowphase_world_kernel<<<np_grid(need), kNpThreads, 0, w->stream>>>(d, work);
        NANS_LAUNCH_CHECK();
        return NANS_OK;
    }
    static int gjk_per_sm[2] = {0, 0}, epa_per_sm = 0;
    if (!epa_per_sm) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&gjk_per_sm[0], narrowphase_world_gjk_kernel<false>, kNpThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&gjk_per_sm[1], narrowphase_world_gjk_kernel<true>, kNpThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&epa_per_sm, narrowphase_world_epa_kernel, kNpThreads, 0);
        if (gjk_per_sm[0] < 1) gjk_per_sm[0] = 1;
        if (gjk_per_sm[1] < 1) gjk_per_sm[1] = 1;
        if (epa_per_sm < 1) epa_per_sm = 1;
        const char *e = getenv("NANS_EPAW_CTAS");       // experiment: fewer resident CTAs = more L1 per polytope arena
        if (e && atoi(e) > 0 && atoi(e) < epa_per_sm) epa_per_sm = atoi(e);
    }
    const int sph = d.n_spheres > 0;
    const int g1 = need < kNumSMs * gjk_per_sm[sph] ? need : kNumSMs * gjk_per_sm[sph];
    if (sph) narrowphase_world_gjk_kernel<true><<<g1, kNpThreads, 0, w->stream>>>(d, work);
    else narrowphase_world_gjk_kernel<false><<<g1, kNpThreads, 0, w->stream>>>(d, work);
    NANS_LAUNCH_CHECK();
    const int need2 = div_up(d.max_contacts, kNpThreads);
    const int g2 = need2 < kNumSMs * epa_per_sm ? need2 : kNumSMs * epa_per_sm;
    narrowphase_world_epa_kernel<<<g2, kNpThreads, 0, w->stream>>>(d);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

// device scratch of the split path, grown on demand and kept (one per device; the batch entry points are
// not re-entrant across streams of the same device)
static int split_scratch(int n, SplitScratch &sc)
{
    constexpr int kMaxDevices = 64;
    static char *blks[kMaxDevices] = {nullptr};
    static size_t caps[kMaxDevices] = {0};
    int dev = 0;
    NANS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    char *&blk = blks[dev];
    size_t &cap = caps[dev];
    const size_t need = 256 + 8 * sizeof(int32_t) * (size_t)n + 8 * sizeof(float4) * (size_t)n;
    if (need > cap) {
        if (blk) NANS_CUDA(cudaFree(blk));
        blk = nullptr; cap = 0;
        NANS_CUDA(cudaMalloc(&blk, need));
        cap = need;
    }
    sc.head = reinterpret_cast<int32_t *>(blk);
    sc.rec = reinterpret_cast<float4 *>(blk + 256);
    int32_t *l = reinterpret_cast<int32_t *>(blk + 256 + 8 * sizeof(float4) * (size_t)n);
    for (int k = 0; k < 4; ++k) { sc.list[k] = l + (size_t)k * n; sc.glist[k] = l + (size_t)(4 + k) * n; }
    return NANS_OK;
}

int launch_narrowphase_batch(int n, const int32_t *type, const float4 *posrad_a, const float4 *verts_a,
                             const float4 *posrad_b, const float4 *verts_b, int32_t *hit, int32_t *gjk,
                             float4 *out, int *work_counter, Counters *counters, cudaStream_t s)
{
    if (n <= 0) return NANS_OK;
    // at most kSplitChunk pairs per round (bounds the scratch: 144 B per pair)
    constexpr int kSplitChunk = 1 << 24;   // (scratch: 160 B per pair)
    static int gjk_per_sm = 0, epa_per_sm = 0;
    if (!gjk_per_sm) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&gjk_per_sm, gjk_split_kernel<BatchSrc>, kNpThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&epa_per_sm, epa_refill_kernel<BatchSrc>, kNpThreads, 0);
        if (gjk_per_sm < 1) gjk_per_sm = 1;
        if (epa_per_sm < 1) epa_per_sm = 1;
    }
    for (int o = 0; o < n; o += kSplitChunk) {
        const int m = n - o < kSplitChunk ? n - o : kSplitChunk;
        SplitScratch sc;
        const int rc = split_scratch(m, sc);
        if (rc) return rc;
        BatchSrc src;
        src.n = m; src.type = type + o;
        src.posrad_a = posrad_a + o; src.verts_a = verts_a ? verts_a + 6 * (size_t)o : nullptr;
        src.posrad_b = posrad_b + o; src.verts_b = verts_b ? verts_b + 6 * (size_t)o : nullptr;
        src.hit = hit + o; src.gjk = gjk ? gjk + o : nullptr; src.out = out + 3 * (size_t)o;
        NANS_CUDA(cudaMemsetAsync(sc.head, 0, 256, s));
        NANS_CUDA(cudaMemsetAsync(work_counter, 0, sizeof(int), s));
        const int need = div_up(m, kNpThreads);
        const int g1 = need < kNumSMs * gjk_per_sm ? need : kNumSMs * gjk_per_sm;
        gjk_split_kernel<BatchSrc><<<g1, kNpThreads, 0, s>>>(src, sc, work_counter, nullptr);
        NANS_LAUNCH_CHECK();
        if (NANS_GJK_CAP > 0) {
            gjk_continue_kernel<BatchSrc><<<g1, kNpThreads, 0, s>>>(src, sc);
            NANS_LAUNCH_CHECK();
        }
        const int g2 = need < kNumSMs * epa_per_sm ? need : kNumSMs * epa_per_sm;
        // EPA in up to three passes: list -> (pairs out of budget) glist -> list.  The GJK lists (glist) are idle once
        // gjk_continue_kernel has run; the original list is idle once pass 0 has.  head: [16 + cls] / [20 + cls] count and
        // tickets of pass 1, [24 + cls] / [28 + cls] of pass 2.
        static int budgets[2][4] = {{-1}};
        if (budgets[0][0] < 0) {
            // per class {CC, CS, SC, SS}; NANS_EPA_BUDGETS="cc0,cs0,cc1,cs1" overrides (65 = no limit; A/B runs)
            int b[4] = {7, 10, 12, 18};      // C3, 16 Mi pairs: 23.98 ms in one pass, 21.85 with these (6,8,10,14: 23.1; 8,12,-,-: 22.2)
            const char *e = getenv("NANS_EPA_BUDGETS");
            if (e) sscanf(e, "%d,%d,%d,%d", &b[0], &b[1], &b[2], &b[3]);
            for (int k = 0; k < 4; ++k) if (b[k] < 1 || b[k] > 65) b[k] = 65;
            budgets[0][0] = b[0]; budgets[0][1] = budgets[0][2] = b[1]; budgets[0][3] = 65;
            budgets[1][0] = b[2]; budgets[1][1] = budgets[1][2] = b[3]; budgets[1][3] = 65;
        }
        for (int pass = 0; pass < 3; ++pass) {
            EpaPass ps[4];
            bool any_budget = false;
            for (int c = 0; c < 4; ++c) {
                EpaPass &q = ps[c];
                const bool odd = pass & 1;
                q.list = odd ? sc.glist[c] : sc.list[c];
                q.count = pass == 0 ? sc.head + c : sc.head + 16 + 8 * (pass - 1) + c;
                q.ticket = pass == 0 ? sc.head + 4 + c : sc.head + 20 + 8 * (pass - 1) + c;
                q.next_list = pass == 2 ? nullptr : (odd ? sc.list[c] : sc.glist[c]);
                q.next_count = pass == 2 ? nullptr : sc.head + 16 + 8 * pass + c;
                for (int k = 0; k < 4; ++k) q.budget[k] = pass == 2 ? 65 : budgets[pass][k];
                if (q.budget[c] < 65) any_budget = true;
            }
            epa_refill_kernel<BatchSrc><<<g2, kNpThreads, 0, s>>>(src, sc, counters, ps[0], ps[1], ps[2], ps[3]);
            NANS_LAUNCH_CHECK();
            if (!any_budget) break;          // this pass ran to the reference's own limit: nothing was deferred
        }
    }
    return NANS_OK;
}

}  // namespace nans
