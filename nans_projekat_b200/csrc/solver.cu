// solver.cu — sequential-impulse contact solver in the reference's exact order, run as a dataflow.
//
// The reference (SolveConstraints code/nans.cpp:1539-1548 -> Constraint :1021-1329) makes ONE
// Gauss-Seidel pass over the contact list in list order; each Constraint reads the velocities the
// previous ones left.  Two contacts commute iff they share no dynamic body, so the sweep is a DAG:
// contact c depends on the previous contact touching its body A and the previous one touching its
// body B.  Executing the DAG in any topological order is bit-identical to the sequential sweep (a
// free greedy colouring would reorder it and change velocities by far more than 1e-4 wherever
// contacts share bodies).  The "colours" here are therefore the DAG's own antichains, discovered
// on the fly: no atomics on body state, every body is touched by one contact at a time.
//
//   incidence_count / scan / fill   per-body lists of incident contacts
//   schedule_kernel                 sort each list by contact id -> successor links + in-degrees
//   seed_kernel                     contacts with in-degree 0 enter the ready queue
//   solve_dataflow_kernel           persistent warps take 32 queue tickets at a time, apply every
//                                   ticket whose contact has arrived (converged lanes), decrement
//                                   the successors' in-degrees and append the newly ready ones.
//
// Progress: queue slot t is filled once the contacts in slots < t that it depends on are done, and
// tickets are issued in order, so every ticket a warp waits on is owed by a warp that is already
// running (no co-residency requirement, no grid barrier).  A spin cap turns any violation into an
// error instead of a hang.  v1 of this file ran level-synchronously (cooperative grid.sync per DAG
// level: 63 % of stall samples sat at the barrier, profiles/r1_v1_ncu_full_summary.txt); that
// kernel is kept for A/B runs (NANS_SOLVER=levels).
//
// HBM-bound in bytes (184 B/contact), latency-bound in practice: critical path = DAG depth x
// (publish -> poll -> gather -> one Constraint).
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "nans_math.cuh"
namespace nans { __device__ unsigned int g_accum_fallbacks; }   // contacts whose accumulation fell back to the literal loop
#define NANS_ACCUM_ON_FALLBACK atomicAdd(&nans::g_accum_fallbacks, 1u)
#include "solver_accum.cuh"
#include "solver_constraint.cuh"
#include "world.cuh"

namespace cg = cooperative_groups;

namespace nans {

__global__ void __launch_bounds__(256) incidence_count_kernel(DeviceWorld w, int versioned)
{
    const int n = w.counters->n_contacts;
    const int stride = gridDim.x * blockDim.x;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
        atomicAdd(&w.deg[w.c_a[c]], 1u);
        const int b = w.c_b[c];
        if (b >= 0) atomicAdd(&w.deg[b], 1u);
        // dataflow: in-degree (filled by schedule_kernel); versioned: 1 where a run of contacts on the same body A starts
        w.indeg[c] = (versioned && (c == 0 || w.c_a[c - 1] != w.c_a[c])) ? 1 : 0;
        w.succ_a[c] = -1;
        w.succ_b[c] = -1;
        w.frontier[0][c] = -1;   // ready queue: empty slots
        w.frontier[1][c] = 0;    // DAG level of the contact (statistic)
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
// versioned != 0: succ_a / succ_b receive the contact's POSITION in the body's sequence instead (the
// number of earlier contacts touching that body = the row version the contact waits for)
__global__ void __launch_bounds__(256) schedule_kernel(DeviceWorld w, int versioned)
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
        const int next = versioned ? k : ((k + 1 < n) ? l[k + 1] : -1);
        if (w.c_a[c] == body) w.succ_a[c] = next; else w.succ_b[c] = next;
        if (k > 0 && !versioned) atomicAdd(&w.indeg[c], 1);
    }
}

// ready queue = frontier[0]; counters->frontier_n[0] = head (tickets issued), [1] = tail (slots filled)
__global__ void __launch_bounds__(256) seed_kernel(DeviceWorld w)
{
    const int n = w.counters->n_contacts;
    const int stride = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int c0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; c0 < n; c0 += stride) {
        const int c = c0 + lane;
        const bool root = c < n && w.indeg[c] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, root);
        if (m == 0) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(&w.counters->frontier_n[1], __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (root) {
            w.frontier[0][base + __popc(m & ((1u << lane) - 1u))] = c;
            w.frontier[1][c] = 1;
        }
    }
}

// dataflow / levels solvers: the records go to memory, with the successor links
__global__ void __launch_bounds__(256) contact_prep_kernel(DeviceWorld w, float dt)
{
    const int n = w.counters->n_contacts;
    const int stride = gridDim.x * blockDim.x;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
        float4 q[kRecQuads];
        int ia, ib;
        constraint_prepare(w, c, dt, q, ia, ib);
        float4 *r = w.crec + (size_t)kRecQuads * c;
#pragma unroll
        for (int k = 0; k < 9; ++k) r[k] = q[k];
        const int sa = w.succ_a[c], sb = w.succ_b[c];
        r[9] = make_float4(__int_as_float(ia), __int_as_float(ib), __int_as_float(sa), __int_as_float(sb));
        // body rows of the successors, so that whoever runs a successor can issue its record loads and
        // its velocity loads in ONE round trip instead of two
        const int none = 0x7fffffff;
        r[10] = make_float4(__int_as_float(sa >= 0 ? __float_as_int(w.c_pa[sa].w) : none),
                            __int_as_float(sa >= 0 ? __float_as_int(w.c_pb[sa].w) : none),
                            __int_as_float(sb >= 0 ? __float_as_int(w.c_pa[sb].w) : none),
                            __int_as_float(sb >= 0 ? __float_as_int(w.c_pb[sb].w) : none));
    }
}

// the velocity-dependent part.  rows = (ia, ib) if the caller already knows the body rows (chain
// following), else kRowsUnknown; returns the successor links and the successors' body rows.
constexpr int kRowsUnknown = 0x7fffffff;
struct NextRows { int sa, sb, sa_ia, sa_ib, sb_ia, sb_ib; };

__device__ __forceinline__ NextRows apply_prepared(const DeviceWorld &w, int c, int known_ia, int known_ib)
{
    const float4 *r = w.crec + (size_t)kRecQuads * c;
    float4 q[kRecQuads];
#pragma unroll
    for (int k = 0; k < kRecQuads; ++k) q[k] = __ldcg(r + k);
    int ia = known_ia, ib = known_ib;
    if (ia == kRowsUnknown) { ia = __float_as_int(q[9].x); ib = __float_as_int(q[9].y); }   // second round trip
    float4 va4 = __ldcg(&w.vel[ia]), wa4 = __ldcg(&w.angvel[ia]);
    float4 vb4 = make_float4(0, 0, 0, 0), wb4 = make_float4(0, 0, 0, 0);
    if (ib >= 0) { vb4 = __ldcg(&w.vel[ib]); wb4 = __ldcg(&w.angvel[ib]); }
    NextRows nx;
    nx.sa = __float_as_int(q[9].z); nx.sb = __float_as_int(q[9].w);
    nx.sa_ia = __float_as_int(q[10].x); nx.sa_ib = __float_as_int(q[10].y);
    nx.sb_ia = __float_as_int(q[10].z); nx.sb_ib = __float_as_int(q[10].w);
    // the successors' records will be wanted next: pull them towards L2 while this contact computes
    if (nx.sa >= 0) {
        const char *p = (const char *)(w.crec + (size_t)kRecQuads * nx.sa);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 128));
    }
    if (nx.sb >= 0) {
        const char *p = (const char *)(w.crec + (size_t)kRecQuads * nx.sb);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 128));
    }
    vec3 V1 = V3(va4), W1 = V3(wa4), V2 = V3(vb4), W2 = V3(wb4);   // the Floor: V = W = 0 (:1278-1289)
    constraint_apply(q, V1, W1, V2, W2, ib >= 0);
    __stcg(&w.vel[ia], make_float4(V1.x, V1.y, V1.z, va4.w));
    __stcg(&w.angvel[ia], make_float4(W1.x, W1.y, W1.z, wa4.w));
    if (ib >= 0) {
        __stcg(&w.vel[ib], make_float4(V2.x, V2.y, V2.z, vb4.w));
        __stcg(&w.angvel[ib], make_float4(W2.x, W2.y, W2.z, wb4.w));
    }
    return nx;
}

// ---- dataflow execution -------------------------------------------------------------------------
constexpr int kFlowThreads = 256;
constexpr int kMaxHops = 8;        // contacts a lane runs back to back before the warp polls again
constexpr int kSpinCap = 1 << 22;   // polls before declaring the schedule broken (seconds of wall time)

// gpu-scope acquire/release primitives (cheaper than the sequentially-consistent __threadfence())
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel()
{
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ int atom_add_relaxed(int *p, int v)
{
    int o;
    asm volatile("atom.relaxed.gpu.global.add.s32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
    return o;
}
__device__ __forceinline__ int atom_add_acq_rel(int *p, int v)
{
    int o;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
    return o;
}

// Memory ordering: a contact's velocity stores are released by the acq_rel decrement of each
// successor's in-degree; whoever performs the LAST decrement has thereby acquired both predecessors'
// stores (RMW chain on the same counter) and either runs the successor itself or hands it over
// through a release store to the queue slot, which the ticket holder reads with an acquire load.
__global__ void __launch_bounds__(kFlowThreads) solve_dataflow_kernel(DeviceWorld w, unsigned long long *trace,
                                                                      int max_hops, int atomic_mode, int sleep_ns, int track_levels)
{
    const int n = w.counters->n_contacts;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int *head = &w.counters->frontier_n[0];
    int *tail = &w.counters->frontier_n[1];
    int *finished = &w.counters->frontier_n[2];
    volatile int *abort_flag = &w.counters->pad[1];
    int32_t *queue = w.frontier[0];
    int32_t *level = w.frontier[1];
    int max_level = 0;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(head, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const int t = base + lane;
        bool done = t >= n;
        bool ticket_open = !done;     // queue slot t not consumed yet
        int c = -1;                   // contact in hand (from the queue, or followed along a chain)
        int rows_a = kRowsUnknown, rows_b = kRowsUnknown;   // its body rows when already known
        int spins = 0, processed = 0;
        while (!__all_sync(0xffffffffu, done)) {
            if (ticket_open && c < 0) {
                c = ld_acquire(queue + t);
                if (c >= 0) ticket_open = false;
            }
            const bool go = c >= 0;
            if (go) {
                // run the contact in hand, then keep following the chain it unlocks (bounded, so the
                // sibling lanes get back to polling their tickets)
                for (int hop = 0; hop < max_hops && c >= 0; ++hop) {
                    const int c_now = c;
                    if (trace) {   // debug: wall-clock (ns) at which each contact starts; 4 slots per contact
                        unsigned long long tns;
                        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
                        trace[4 * (size_t)c] = tns;
                        trace[4 * (size_t)c + 3] = (rows_a == kRowsUnknown) ? 0ull : 1ull;   // 0 = from the queue, 1 = chain
                    }
                    const int lv = track_levels ? __ldcg(&level[c]) : 0;
                    const NextRows nx = apply_prepared(w, c, rows_a, rows_b);
                    if (trace) {
                        unsigned long long tns;
                        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
                        trace[4 * (size_t)c_now + 1] = tns;
                    }
                    ++processed;
                    max_level = max(max_level, lv);
                    if (track_levels) {   // statistic only (DAG depth); two more RMWs to drain before the fence
                        if (nx.sa >= 0) atomicMax(&level[nx.sa], lv + 1);
                        if (nx.sb >= 0) atomicMax(&level[nx.sb], lv + 1);
                    }
                    int oa = 0, ob = 0;
                    if (atomic_mode == 0) {            // one release fence, two relaxed RMWs in flight together
                        fence_acq_rel();               // release: this contact's velocity stores
                        if (nx.sa >= 0) oa = atom_add_relaxed(&w.indeg[nx.sa], -1);
                        if (nx.sb >= 0) ob = atom_add_relaxed(&w.indeg[nx.sb], -1);
                        if (oa == 1 || ob == 1) fence_acq_rel();   // acquire: the other predecessors' stores
                    } else {                           // acq_rel RMWs
                        if (nx.sa >= 0) oa = atom_add_acq_rel(&w.indeg[nx.sa], -1);
                        if (nx.sb >= 0) ob = atom_add_acq_rel(&w.indeg[nx.sb], -1);
                    }
                    const bool ra = oa == 1, rb = ob == 1;
                    if (trace) {
                        unsigned long long tns;
                        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
                        trace[4 * (size_t)c_now + 2] = tns;
                    }
                    // a successor we completed is run by this lane straight away (no queue round
                    // trip); if both became ready the second one goes to the queue
                    const bool push_b = ra && rb;
                    const unsigned am = __activemask();
                    const unsigned pb = __ballot_sync(am, push_b);
                    if (pb) {
                        const int leader = __ffs(am) - 1;
                        int slot = 0;
                        if (lane == leader) slot = atomicAdd(tail, __popc(pb));
                        slot = __shfl_sync(am, slot, leader) + __popc(pb & lt);
                        if (push_b) st_release(queue + slot, nx.sb);
                    }
                    c = ra ? nx.sa : (rb ? nx.sb : -1);
                    rows_a = ra ? nx.sa_ia : nx.sb_ia;
                    rows_b = ra ? nx.sa_ib : nx.sb_ib;
                }
                if (c < 0) { rows_a = kRowsUnknown; rows_b = kRowsUnknown; if (!ticket_open) done = true; }
            }
            if (!__any_sync(0xffffffffu, go)) {
                // nothing arrived: publish this warp's progress, then check whether everything is
                // finished (chain following leaves tickets unfilled, so emptiness is decided by count)
                const int p = __reduce_add_sync(0xffffffffu, processed);
                if (p) {
                    if (lane == 0) atomicAdd(finished, p);
                    processed = 0;
                }
                if (*(volatile int *)finished >= n) { done = true; continue; }
                if (++spins > kSpinCap || *abort_flag) { *abort_flag = 1; return; }
                if (sleep_ns) __nanosleep(sleep_ns);
            }
        }
        processed = __reduce_add_sync(0xffffffffu, processed);
        if (lane == 0 && processed) atomicAdd(finished, processed);
    }
    max_level = __reduce_max_sync(0xffffffffu, max_level);
    if (lane == 0 && max_level) atomicMax(&w.counters->solver_levels, max_level);
}

// ---- v3: versioned body rows (default) ----------------------------------------------------------
// The dependency of a contact on its predecessors is carried by the DATA it needs: during the solve
// every body's velocity lives in a 16-byte row (V.xyz, version) and its angular velocity in a second
// row (W.xyz, version | level << 20), where version = number of contacts applied to that body so
// far.  Contact c waits until both of its bodies show the versions it was scheduled for (its
// position in each body's contact sequence), applies itself and stores the rows with version + 1.
// A row is one aligned 128-bit access, so value and version arrive together: no fences, no in-degree
// atomics, no ready queue; the hop from a contact to its successor is one L2 store -> poll.
//
// Warps take contacts in LIST order, 32 at a time.  Progress: every contact a lane waits for is
// earlier in the list, so its ticket is already held by a running warp; by induction the earliest
// unfinished contact is always runnable.  (A spin cap turns any violation into an error.)
#ifndef NANS_VER_THREADS
#define NANS_VER_THREADS 256
#endif
#ifndef NANS_VER_MINBLOCKS
#define NANS_VER_MINBLOCKS 2   // resident CTAs the register budget is sized for (sweep, solver stage: 256x2 0.374, 128x4 0.375,
#endif                         // 256x3 / 128x6 (80 registers, spills) 0.455, 256x4 (64 registers) 0.539 ms)
constexpr int kVerThreads = NANS_VER_THREADS;
constexpr int kVerMask = 0xfffff;   // version bits kept in the angular row (the rest carries the DAG level)

__device__ __forceinline__ float4 ld_row(const float4 *p)
{
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_row(float4 *p, vec3 v, int tag)
{
    asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(__int_as_float(tag)) : "memory");
}

// rows for the solve: (vel.xyz, 0) and (angvel.xyz, 0).  They borrow the AABB arrays, which are dead
// between detection and the next step's broadphase.
__global__ void __launch_bounds__(256) ver_seed_kernel(DeviceWorld w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.nb) return;
    float4 v = w.vel[i], a = w.angvel[i];
    v.w = 0.f; a.w = 0.f;
    w.aabb_lo[i] = v;
    w.aabb_hi[i] = a;
}
__global__ void __launch_bounds__(256) ver_finish_kernel(DeviceWorld w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.nb) return;
    const float4 v = w.aabb_lo[i], a = w.aabb_hi[i];
    if (__float_as_int(v.w) == 0) return;            // untouched by any contact
    float4 *pv = &w.vel[i], *pa = &w.angvel[i];
    pv->x = v.x; pv->y = v.y; pv->z = v.z;          // .w keeps 1/Mass, 1/MOI
    pa->x = a.x; pa->y = a.y; pa->z = a.z;
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// run r = contacts [run_start[r], run_start[r + 1]) = consecutive contacts with the same body A
__global__ void __launch_bounds__(256) run_scatter_kernel(DeviceWorld w)
{
    const int n = w.counters->n_contacts;
    const int stride = gridDim.x * blockDim.x;
    int32_t *run_start = w.frontier[0];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
        const uint32_t r = w.pair_hit_scan[c];
        if (w.indeg[c]) run_start[r] = c;
        if (c == n - 1) w.counters->frontier_n[1] = (int)r + w.indeg[c];   // number of runs
    }
}

// One lane per RUN: the lane keeps body A's velocity in registers along the run and only body B's
// rows go through memory, and the lanes of a warp fire their j-th contacts together (a lane per
// contact left ~8 of 32 lanes active per firing: the kernel was bound by issue slots).
// trace (debug, NANS_SOLVER_TRACE=1): per contact {fire ns, stored ns, ticket ns, polls}; frontier[1] = DAG level
__global__ void __launch_bounds__(kVerThreads, NANS_VER_MINBLOCKS) solve_versioned_kernel(DeviceWorld w, float dt, int sleep_ns, unsigned long long *trace)
{
    const int n = w.counters->n_contacts;
    const int n_runs = w.counters->frontier_n[1];
    const int lane = threadIdx.x & 31;
    int *head = &w.counters->frontier_n[0];
    volatile int *abort_flag = &w.counters->pad[1];
    const int32_t *run_start = w.frontier[0];
    float4 *sv = w.aabb_lo, *sw = w.aabb_hi;
    int max_level = 0;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(head, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_runs) break;
        const int r = base + lane;
        bool pending = r < n_runs;
        int c = 0, c_end = 0;
        if (pending) {
            c = run_start[r];
            c_end = (r + 1 < n_runs) ? run_start[r + 1] : n;
        }
        float4 q[kRecQuads];
        int ia = 0, ib = -1, ea = 0, eb = 0, lva = 0;
        vec3 V1 = V3(0.f, 0.f, 0.f), W1 = V3(0.f, 0.f, 0.f);
        bool loaded = false, have_a = false, have_b = false;
        float4 vb4 = make_float4(0, 0, 0, 0), wb4 = make_float4(0, 0, 0, 0);
        int spins = 0;
        unsigned long long t_ticket = 0;
        if (trace) t_ticket = global_ns();
        while (__any_sync(0xffffffffu, pending)) {
            if (pending) {
                if (!loaded) {
                    // the velocity-independent part of the contact, straight into registers (no record
                    // round trip through memory); runs while the predecessors are still at work
                    int ra, rb;
                    constraint_prepare(w, c, dt, q, ra, rb);
                    ib = rb; eb = w.succ_b[c];
                    if (!have_a) { ia = ra; ea = w.succ_a[c]; }
                    loaded = true;
                }
                // all outstanding rows in ONE round trip (the hop latency is what bounds the solve)
                const bool need_a = !have_a, need_b = !have_b && ib >= 0;
                float4 va, wa;
                if (need_a) { va = ld_row(sv + ia); wa = ld_row(sw + ia); }
                if (need_b) { vb4 = ld_row(sv + ib); wb4 = ld_row(sw + ib); }
                if (need_a && __float_as_int(va.w) == ea && (__float_as_int(wa.w) & kVerMask) == (ea & kVerMask)) {
                    V1 = V3(va); W1 = V3(wa);
                    lva = (int)(__float_as_uint(wa.w) >> 20);
                    have_a = true;
                }
                if (need_b) have_b = __float_as_int(vb4.w) == eb && (__float_as_int(wb4.w) & kVerMask) == (eb & kVerMask);
                if (ib < 0) have_b = true;
            }
            // every ready lane fires at once (holding ready lanes back to fire more of them together was
            // measured: each poll of patience costs ~50 us per step, the solve is bound by hop latency)
            const bool ready = pending && have_a && have_b;
            const bool go = __any_sync(0xffffffffu, ready);
            if (go && ready) {
                unsigned long long t_fire = 0;
                long long ck = 0;
                if (trace) { t_fire = global_ns(); ck = clock64(); }
                vec3 V2 = V3(vb4), W2 = V3(wb4);                     // the Floor: V = W = 0
                constraint_apply(q, V1, W1, V2, W2, ib >= 0);
                int lv = lva;
                if (ib >= 0) lv = max(lv, (int)(__float_as_uint(wb4.w) >> 20));
                lva = min(lv + 1, 4095);
                max_level = max(max_level, lv + 1);
                if (ib >= 0) {
                    st_row(sv + ib, V2, eb + 1);
                    st_row(sw + ib, W2, ((eb + 1) & kVerMask) | (lva << 20));
                }
                if (trace) {
                    trace[4 * (size_t)c] = t_fire;
                    trace[4 * (size_t)c + 1] = global_ns();
                    trace[4 * (size_t)c + 2] = t_ticket;
                    trace[4 * (size_t)c + 3] = (unsigned long long)(clock64() - ck);   // cycles: apply + row stores
                    w.frontier[1][c] = lv;
                }
                ++ea; ++c;
                loaded = false;
                have_b = false;
                vb4 = make_float4(0, 0, 0, 0); wb4 = make_float4(0, 0, 0, 0);
                if (c == c_end) {          // body A leaves the run: publish it
                    st_row(sv + ia, V1, ea);
                    st_row(sw + ia, W1, (ea & kVerMask) | (lva << 20));
                    pending = false;
                }
            }
            if (!go) {
                if (++spins > kSpinCap || *abort_flag) { *abort_flag = 1; return; }
                if (sleep_ns) __nanosleep(sleep_ns);
            }
        }
    }
    max_level = __reduce_max_sync(0xffffffffu, max_level);
    if (lane == 0 && max_level) atomicMax(&w.counters->solver_levels, max_level);
}

// ---- v1: level-synchronous execution (kept for A/B, NANS_SOLVER=levels) -----------------------
constexpr int kSolveThreads = 256;

__global__ void __launch_bounds__(kSolveThreads) solve_levels_kernel(DeviceWorld w)
{
    cg::grid_group grid = cg::this_grid();
    const int n = w.counters->n_contacts;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    volatile int32_t *fn = w.counters->frontier_n;
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
            const NextRows nx = apply_prepared(w, c, kRowsUnknown, kRowsUnknown);
            const int sa = nx.sa, sb = nx.sb;
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
    static int mode = -1, sm_count = 0;
    if (mode < 0) {
        const char *e = getenv("NANS_SOLVER");   // versioned (default) | flow | levels
        mode = (e && !strcmp(e, "levels")) ? 1 : (e && !strcmp(e, "flow")) ? 0 : 2;
        cudaDeviceProp prop;
        NANS_CUDA(cudaGetDeviceProperties(&prop, w->device));
        sm_count = prop.multiProcessorCount;
    }
    NANS_CUDA(cudaMemsetAsync(d.deg, 0, sizeof(uint32_t) * ((size_t)d.nb + 1), s));
    NANS_CUDA(cudaMemsetAsync(d.cursor, 0, sizeof(uint32_t) * (size_t)d.nb, s));
    NANS_CUDA(cudaMemsetAsync(d.counters->frontier_n, 0, sizeof(int32_t) * 3, s));
    {
        void *fb = nullptr;
        NANS_CUDA(cudaGetSymbolAddress(&fb, g_accum_fallbacks));
        NANS_CUDA(cudaMemsetAsync(fb, 0, sizeof(unsigned int), s));
    }
    const int grid = min(div_up(d.max_contacts, 256), kNumSMs * 8);
    incidence_count_kernel<<<grid, 256, 0, s>>>(d, mode == 2);
    NANS_LAUNCH_CHECK();
    int rc = exclusive_scan_u32(d.deg, d.deg, d.nb + 1, d.scan_block, s);
    if (rc) return rc;
    incidence_fill_kernel<<<grid, 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    schedule_kernel<<<div_up(d.nb, 256), 256, 0, s>>>(d, mode == 2);
    NANS_LAUNCH_CHECK();
    if (mode != 2) {
        contact_prep_kernel<<<grid, 256, 0, s>>>(d, dt);
        NANS_LAUNCH_CHECK();
    }

    if (mode == 1) {
        if (!w->coop_blocks_per_sm) {
            int per_sm = 0;
            NANS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_levels_kernel, kSolveThreads, 0));
            w->coop_blocks_per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
        }
        void *args[] = {(void *)&d};
        NANS_CUDA(cudaLaunchCooperativeKernel((void *)solve_levels_kernel, dim3(sm_count * w->coop_blocks_per_sm),
                                              dim3(kSolveThreads), args, 0, s));
        ++g_launches;
        return NANS_OK;
    }
    if (mode == 2) {
        static int ver_blocks = 0, ver_sleep = 0;
        if (!ver_blocks) {
            int per_sm = 0;
            NANS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_versioned_kernel, kVerThreads, 0));
            const char *e;
            int want = (e = getenv("NANS_VER_BLOCKS")) ? atoi(e) : 2;
            if (want < 1) want = 1;
            ver_blocks = sm_count * (want < per_sm ? want : per_sm);   // every CTA must be resident (spinning lanes)
            ver_sleep = (e = getenv("NANS_VER_SLEEP")) ? atoi(e) : 0;
        }
        rc = exclusive_scan_u32_dn((const uint32_t *)d.indeg, d.pair_hit_scan, d.max_contacts, &d.counters->n_contacts, 0,
                                   d.scan_block, s);
        if (rc) return rc;
        run_scatter_kernel<<<grid, 256, 0, s>>>(d);
        NANS_LAUNCH_CHECK();
        ver_seed_kernel<<<div_up(d.nb, 256), 256, 0, s>>>(d);
        NANS_LAUNCH_CHECK();
        static int ver_trace = -1;
        if (ver_trace < 0) ver_trace = getenv("NANS_SOLVER_TRACE") ? 1 : 0;
        solve_versioned_kernel<<<ver_blocks, kVerThreads, 0, s>>>(d, dt, ver_sleep, ver_trace ? (unsigned long long *)d.pair_out : nullptr);
        NANS_LAUNCH_CHECK();
        ver_finish_kernel<<<div_up(d.nb, 256), 256, 0, s>>>(d);
        NANS_LAUNCH_CHECK();
        return NANS_OK;
    }
    seed_kernel<<<grid, 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    static int flow_blocks = 0;
    if (!flow_blocks) {
        int per_sm = 0;
        NANS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_dataflow_kernel, kFlowThreads, 0));
        (void)per_sm;
        flow_blocks = sm_count;   // one CTA per SM: more polling warps only add interference (profiles/r1 sweep)
    }
    static int trace_on = -1;
    if (trace_on < 0) trace_on = getenv("NANS_SOLVER_TRACE") ? 1 : 0;
    // debug trace reuses the narrowphase output block (idle during the solve)
    static int hops = -1, amode = 0, sleep_ns = 32, track_levels = 1;
    if (hops < 0) {   // tuning knobs (defaults chosen from the sweeps in profiles/)
        const char *e;
        hops = (e = getenv("NANS_FLOW_HOPS")) ? atoi(e) : 1;
        amode = (e = getenv("NANS_FLOW_ATOMICS")) ? atoi(e) : 0;
        sleep_ns = (e = getenv("NANS_FLOW_SLEEP")) ? atoi(e) : 32;
        if ((e = getenv("NANS_FLOW_BLOCKS"))) flow_blocks = sm_count * atoi(e);
        track_levels = (e = getenv("NANS_SOLVER_LEVELS")) ? atoi(e) : 1;
        if (hops < 1) hops = 1;
    }
    solve_dataflow_kernel<<<flow_blocks, kFlowThreads, 0, s>>>(d, trace_on ? (unsigned long long *)d.pair_out : nullptr,
                                                               hops, amode, sleep_ns, track_levels | trace_on);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

}  // namespace nans

namespace nans {
int solver_accum_fallbacks(World *w, int32_t *out)
{
    unsigned int v = 0;
    NANS_CUDA(cudaStreamSynchronize(w->stream));
    NANS_CUDA(cudaMemcpyFromSymbol(&v, g_accum_fallbacks, sizeof(v)));
    *out = (int32_t)v;
    return NANS_OK;
}
}  // namespace nans
