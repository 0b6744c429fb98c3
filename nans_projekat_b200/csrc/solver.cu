// solver.cu — sequential-impulse contact solver, exact reference order, parallel by dependency level.
//
// The reference (SolveConstraints code/nans.cpp:1539-1548 -> Constraint :1021-1329) makes ONE
// Gauss-Seidel pass over the contact list in list order; each Constraint reads the velocities the
// previous ones left.  Two contacts commute iff they share no dynamic body, so the sweep is a DAG:
// contact c depends on the previous contact touching its body A and the previous one touching its
// body B.  The colours here are the levels of that DAG (an ORDER-PRESERVING colouring): every
// level is a set of body-disjoint contacts, applied atomics-free, and the result is bit-identical
// to the sequential sweep (a free greedy colouring would reorder the sweep and change velocities
// by far more than 1e-4 wherever contacts share bodies).
//
//   incidence_count / scan / fill     per-body lists of incident contacts
//   schedule_kernel                   sort each list by contact id -> successor links + in-degrees
//   solve_levels_kernel (cooperative) frontier = contacts with in-degree 0; per level: apply the
//                                     frontier, decrement successors, grid.sync()
//
// HBM-bound in bytes (184 B/contact), latency-bound in practice: depth x (grid sync + one Constraint).
#include <cooperative_groups.h>

#include "nans_math.cuh"
#include "world.cuh"

namespace cg = cooperative_groups;

namespace nans {

__global__ void __launch_bounds__(256) incidence_count_kernel(DeviceWorld w)
{
    const int n = w.counters->n_contacts;
    const int stride = gridDim.x * blockDim.x;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
        atomicAdd(&w.deg[w.c_a[c]], 1u);
        const int b = w.c_b[c];
        if (b >= 0) atomicAdd(&w.deg[b], 1u);
        w.indeg[c] = 0;
        w.succ_a[c] = -1;
        w.succ_b[c] = -1;
    }
}

__global__ void __launch_bounds__(256) incidence_fill_kernel(DeviceWorld w)
{
    const int n = w.counters->n_contacts;
    const int stride = gridDim.x * blockDim.x;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
        const int a = w.c_a[c], b = w.c_b[c];
        w.inc[w.deg[a] + atomicAdd(&w.cursor[a], 1u)] = c;
        if (b >= 0) w.inc[w.deg[b] + atomicAdd(&w.cursor[b], 1u)] = c;
    }
}

// one thread per body: order its incident contacts by list position, link successors
__global__ void __launch_bounds__(256) schedule_kernel(DeviceWorld w)
{
    const int body = blockIdx.x * blockDim.x + threadIdx.x;
    if (body >= w.nb) return;
    const uint32_t beg = w.deg[body], end = w.deg[body + 1];
    int32_t *l = w.inc + beg;
    const int n = (int)(end - beg);
    for (int i = 1; i < n; ++i) {
        const int v = l[i];
        int j = i;
        while (j > 0 && l[j - 1] > v) { l[j] = l[j - 1]; --j; }
        l[j] = v;
    }
    for (int k = 0; k < n; ++k) {
        const int c = l[k];
        const int next = (k + 1 < n) ? l[k + 1] : -1;
        if (w.c_a[c] == body) w.succ_a[c] = next; else w.succ_b[c] = next;
        if (k > 0) atomicAdd(&w.indeg[c], 1);
    }
}

// Constraint, code/nans.cpp:1021-1329 — bug-compatible: minus sign on body B's
// angular JMJ term, cross(W, N) instead of cross(W, R), un-normalised T1, 70 iterations over
// constants of which only the last delta is applied, friction bound evaluated in fp64.
__device__ __forceinline__ void apply_constraint(const DeviceWorld &w, int c, float dt)
{
    const int ia = w.c_a[c], ib = w.c_b[c];
    const float4 pa4 = w.pos[ia];
    float4 va4 = __ldcg(&w.vel[ia]);      // w = 1/Mass
    float4 wa4 = __ldcg(&w.angvel[ia]);   // w = 1/MOI
    const vec3 posA = V3(pa4);
    const float invM1 = va4.w, invI1 = wa4.w;
    vec3 V1 = V3(va4), W1 = V3(wa4);
    vec3 posB, V2, W2;
    float invM2, invI2;
    float4 vb4 = make_float4(0, 0, 0, 0), wb4 = make_float4(0, 0, 0, 0);
    if (ib >= 0) {
        posB = V3(w.pos[ib]);
        vb4 = __ldcg(&w.vel[ib]);
        wb4 = __ldcg(&w.angvel[ib]);
        invM2 = vb4.w; invI2 = wb4.w;
        V2 = V3(vb4); W2 = V3(wb4);
    } else {
        const int k = -ib - 1;                 // the Floor: V = W = 0, never updated (:1278-1289)
        const float4 sp = w.st_pos[k], sa = w.st_ang[k];
        posB = V3(sp);
        invM2 = sp.w; invI2 = sa.w;
        V2 = V3(0.f, 0.f, 0.f); W2 = V3(0.f, 0.f, 0.f);
    }
    vec3 N = normalize(V3(w.c_n[c]));
    if (equal(N, V3(0.f, 0.f, 0.f))) N = normalize(posB - posA);   // :1115-1119
    const vec3 R1 = V3(w.c_pa[c]) - posA;
    const vec3 R2 = V3(w.c_pb[c]) - posB;
    vec3 T1;
    if (N.x >= 0.57735f) T1 = V3(N.y, -N.x, 0.0f); else T1 = V3(0.0f, N.z, -N.y);
    const vec3 T2 = cross(N, T1);                                    // T1 is NOT normalised (:1133)
    const float depth = dot((posA + R1) - (posB + R2), N);
    const vec3 RN1 = cross(R1, N), RN2 = cross(R2, N);
    float JMJn = fadd(invM1, invM2);
    JMJn = fadd(JMJn, fsub(fmul(invI1, dot(RN1, RN1)), fmul(invI2, dot(-RN2, -RN2))));
    JMJn = fdiv(1.0f, JMJn);
    const vec3 dVn = ((V1 + cross(W1, N)) - V2) - cross(W2, N);
    const float JdVn = dot(dVn, N);
    const float Beta = 0.3f, Cr = 0.1f;
    const float B = fadd(fmul(fdiv(-Beta, dt), depth), fmul(Cr, JdVn));
    const vec3 R1T1 = cross(R1, T1), R2T1 = cross(R2, T1), R1T2 = cross(R1, T2), R2T2 = cross(R2, T2);
    float JMJt1 = fadd(invM1, invM2);
    JMJt1 = fadd(JMJt1, fsub(fmul(invI1, dot(R1T1, R1T1)), fmul(invI2, dot(-R2T1, -R2T1))));
    JMJt1 = fdiv(1.0f, JMJt1);
    float JMJt2 = fadd(invM1, invM2);
    JMJt2 = fadd(JMJt2, fsub(fmul(invI1, dot(R1T2, R1T2)), fmul(invI2, dot(-R2T2, -R2T2))));
    JMJt2 = fdiv(1.0f, JMJt2);
    const vec3 dVt1 = ((V1 + cross(W1, T1)) - V2) - cross(W2, T1);
    const float JdVt1 = dot(dVt1, T1);
    const vec3 dVt2 = ((V1 + cross(W1, T2)) - V2) - cross(W2, T2);
    const float JdVt2 = dot(dVt2, T2);

    // :1176-1227 — the accumulators start at zero (SolveConstraints works on a copy, :1545)
    const float lambdaN = fmul(fadd(-JdVn, B), JMJn);
    const float lambdaT1 = fmul(-JdVt1, JMJt1);
    const float lambdaT2 = fmul(-JdVt2, JMJt2);
    const double kFric = 1.4142135623730951 * (double)0.1f;   // sqrt(2) * Cf, evaluated in fp64 (:1197)
    float DLN = 0.f, sumN = 0.f, DLT1 = 0.f, sumT1 = 0.f, DLT2 = 0.f, sumT2 = 0.f;
#pragma unroll 2
    for (int iter = 0; iter < 70; ++iter) {
        const float oldN = sumN;
        sumN = fadd(sumN, lambdaN);
        if (sumN < 0) sumN = 0.0f;
        DLN = fsub(sumN, oldN);
        const float maxT = __double2float_rn(__dmul_rn(kFric, (double)sumN));
        const float oldT1 = sumT1;
        sumT1 = fadd(sumT1, lambdaT1);
        if (sumT1 < -maxT) sumT1 = -maxT;
        if (sumT1 > maxT) sumT1 = maxT;
        DLT1 = fsub(sumT1, oldT1);
        const float oldT2 = sumT2;
        sumT2 = fadd(sumT2, lambdaT2);
        if (sumT2 < -maxT) sumT2 = -maxT;
        if (sumT2 > maxT) sumT2 = maxT;
        DLT2 = fsub(sumT2, oldT2);
    }
    const vec3 LI = N * DLN, LIT1 = T1 * DLT1, LIT2 = T2 * DLT2;
    const vec3 AI1 = RN1 * DLN, AI2 = RN2 * DLN;
    const vec3 AI1T1 = R1T1 * DLT1, AI2T1 = R2T1 * DLT1;
    const vec3 AI1T2 = R1T2 * DLT2, AI2T2 = R2T2 * DLT2;
    // :1229-1328 — normal, then T1, then T2; a != b so register accumulation equals the
    // reference's read-modify-write sequence
    V1 = V1 + invM1 * LI;     W1 = W1 + invI1 * AI1;
    V1 = V1 + invM1 * LIT1;   W1 = W1 + invI1 * AI1T1;
    V1 = V1 + invM1 * LIT2;   W1 = W1 + invI1 * AI1T2;
    __stcg(&w.vel[ia], make_float4(V1.x, V1.y, V1.z, va4.w));
    __stcg(&w.angvel[ia], make_float4(W1.x, W1.y, W1.z, wa4.w));
    if (ib >= 0) {
        V2 = V2 - invM2 * LI;     W2 = W2 - invI2 * AI2;
        V2 = V2 - invM2 * LIT1;   W2 = W2 - invI2 * AI2T1;
        V2 = V2 - invM2 * LIT2;   W2 = W2 - invI2 * AI2T2;
        __stcg(&w.vel[ib], make_float4(V2.x, V2.y, V2.z, vb4.w));
        __stcg(&w.angvel[ib], make_float4(W2.x, W2.y, W2.z, wb4.w));
    }
}

constexpr int kSolveThreads = 256;

__global__ void __launch_bounds__(kSolveThreads) solve_levels_kernel(DeviceWorld w, float dt)
{
    cg::grid_group grid = cg::this_grid();
    const int n = w.counters->n_contacts;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    volatile int32_t *fn = w.counters->frontier_n;
    // level-0 frontier: contacts with no predecessor
    for (int c = tid; c < n; c += nthreads)
        if (w.indeg[c] == 0) w.frontier[0][atomicAdd(&w.counters->frontier_n[0], 1)] = c;
    grid.sync();
    int level = 0;
    while (true) {
        const int cur = level % 3, nxt = (level + 1) % 3, clr = (level + 2) % 3;
        const int fcount = fn[cur];
        if (fcount == 0) break;
        if (tid == 0) fn[clr] = 0;
        const int32_t *fr = w.frontier[cur];
        for (int i = tid; i < fcount; i += nthreads) {
            const int c = __ldcg(&fr[i]);
            apply_constraint(w, c, dt);
            const int sa = w.succ_a[c], sb = w.succ_b[c];
            if (sa >= 0 && atomicSub(&w.indeg[sa], 1) == 1)
                w.frontier[nxt][atomicAdd(&w.counters->frontier_n[nxt], 1)] = sa;
            if (sb >= 0 && atomicSub(&w.indeg[sb], 1) == 1)
                w.frontier[nxt][atomicAdd(&w.counters->frontier_n[nxt], 1)] = sb;
        }
        ++level;
        grid.sync();
    }
    if (tid == 0) w.counters->solver_levels = level;
}

int launch_solver(World *w, float dt)
{
    DeviceWorld &d = w->d;
    if (d.nb == 0) return NANS_OK;
    cudaStream_t s = w->stream;
    NANS_CUDA(cudaMemsetAsync(d.deg, 0, sizeof(uint32_t) * ((size_t)d.nb + 1), s));
    NANS_CUDA(cudaMemsetAsync(d.cursor, 0, sizeof(uint32_t) * (size_t)d.nb, s));
    NANS_CUDA(cudaMemsetAsync(d.counters->frontier_n, 0, sizeof(int32_t) * 3, s));
    const int grid = min(div_up(d.max_contacts, 256), kNumSMs * 8);
    incidence_count_kernel<<<grid, 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    int rc = exclusive_scan_u32(d.deg, d.deg, d.nb + 1, d.scan_block, s);
    if (rc) return rc;
    incidence_fill_kernel<<<grid, 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    schedule_kernel<<<div_up(d.nb, 256), 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();

    if (!w->coop_blocks_per_sm) {
        int per_sm = 0;
        NANS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_levels_kernel, kSolveThreads, 0));
        w->coop_blocks_per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
    }
    cudaDeviceProp prop;
    static int sm_count = 0;
    if (!sm_count) { NANS_CUDA(cudaGetDeviceProperties(&prop, w->device)); sm_count = prop.multiProcessorCount; }
    void *args[] = {(void *)&d, (void *)&dt};
    NANS_CUDA(cudaLaunchCooperativeKernel((void *)solve_levels_kernel, dim3(sm_count * w->coop_blocks_per_sm),
                                          dim3(kSolveThreads), args, 0, s));
    ++g_launches;
    return NANS_OK;
}

}  // namespace nans
