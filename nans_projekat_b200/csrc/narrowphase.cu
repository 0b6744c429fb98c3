// narrowphase.cu — kernels around narrowphase.cuh: GJK+EPA over the world's candidate pairs and
// over stand-alone shape pairs (config C3).  One pair per thread; warps fetch 32-pair chunks from
// a global counter (pairs differ by >10x in work: a GJK miss is one support call, an EPA hit is
// dozens), grid = SM count x resident CTAs.  FP32-pipe / divergence bound, not HBM bound:
// 216 B in + 48 B out per pair against ~1-5 kflop of unfused fp32.
#include "narrowphase.cuh"

#ifndef NANS_NP_MINBLOCKS
#define NANS_NP_MINBLOCKS 5   // 96 registers/thread: best of the sweep in profiles/ (4: 128 regs, 6: 80 regs + spills)
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

__global__ void __launch_bounds__(kNpThreads, NANS_NP_MINBLOCKS) narrowphase_world_kernel(DeviceWorld w, int *work_counter)
{
    EpaArena E;
    const int lane = threadIdx.x & 31;
    const int n_pairs = w.counters->n_pairs;
    int ovf = 0, max_faces = 0, found = 0;
    NpShapes S;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(work_counter, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_pairs) break;
        const int p = base + lane;
        if (p < n_pairs) {
            const int ra = w.pair_a[p], rb = w.pair_b[p];
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

// stand-alone pairs (config C3): type per pair, shapes given explicitly
__global__ void __launch_bounds__(kNpThreads, NANS_NP_MINBLOCKS) narrowphase_batch_kernel(
    int n, const int32_t *__restrict__ type, const float4 *__restrict__ posrad_a,
    const float4 *__restrict__ verts_a, const float4 *__restrict__ posrad_b,
    const float4 *__restrict__ verts_b, int32_t *__restrict__ hit, int32_t *__restrict__ gjk,
    float4 *__restrict__ out, int *work_counter, Counters *counters)
{
    EpaArena E;
    const int lane = threadIdx.x & 31;
    int ovf = 0, max_faces = 0;
    NpShapes S;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(work_counter, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const int p = base + lane;
        if (p < n) {
            const int t = type[p];
            const bool a_sphere = (t == NANS_SS || t == NANS_SF);
            const bool b_sphere = (t == NANS_CS || t == NANS_SS);
            const float4 pa = posrad_a[p], pb = posrad_b[p];
            S.posA = V3(pa); S.radA = pa.w;
            S.posB = V3(pb); S.radB = pb.w;
            if (!a_sphere) load_box(0, verts_a + 6 * (size_t)p);
            if (!b_sphere) load_box(1, verts_b + 6 * (size_t)p);
            const NpResult r = dispatch(a_sphere, b_sphere, S, E, ovf, max_faces);
            hit[p] = r.hit;
            if (gjk) gjk[p] = r.gjk;
            float4 *o = out + 3 * (size_t)p;
            o[0] = make_float4(r.PA.x, r.PA.y, r.PA.z, 0.f);
            o[1] = make_float4(r.PB.x, r.PB.y, r.PB.z, 0.f);
            o[2] = make_float4(r.N.x, r.N.y, r.N.z, 0.f);
        }
    }
    ovf = __reduce_or_sync(0xffffffffu, ovf);
    max_faces = __reduce_max_sync(0xffffffffu, max_faces);
    if (lane == 0 && counters) {
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
    const int grid = np_grid(div_up(d.max_pairs, kNpThreads));
    narrowphase_world_kernel<<<grid, kNpThreads, 0, w->stream>>>(d, work);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

int launch_narrowphase_batch(int n, const int32_t *type, const float4 *posrad_a, const float4 *verts_a,
                             const float4 *posrad_b, const float4 *verts_b, int32_t *hit, int32_t *gjk,
                             float4 *out, int *work_counter, Counters *counters, cudaStream_t s)
{
    if (n <= 0) return NANS_OK;
    NANS_CUDA(cudaMemsetAsync(work_counter, 0, sizeof(int), s));
    const int grid = np_grid(div_up(n, kNpThreads));
    narrowphase_batch_kernel<<<grid, kNpThreads, 0, s>>>(n, type, posrad_a, verts_a, posrad_b, verts_b, hit, gjk,
                                                         out, work_counter, counters);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

}  // namespace nans
