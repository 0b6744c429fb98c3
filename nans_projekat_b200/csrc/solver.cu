// solver.cu — sequential-impulse contact solver, run as a dataflow over versioned body rows.
//
// The reference (SolveConstraints code/nans.cpp:1539-1548 -> Constraint :1021-1329) makes ONE
// Gauss-Seidel pass over the contact list in list order; each Constraint reads the velocities the
// previous ones left.  Two contacts commute iff they share no dynamic body, so the sweep is a DAG:
// contact c depends on the previous contact touching its body A and the previous one touching its
// body B, and ANY topological execution is bit-identical to the sequential sweep.  The DAG's
// antichains are the "colours" of a greedy graph colouring in sweep order; they are discovered on
// the fly, with no atomics on body state (a body is touched by one contact at a time).
//
//   incidence_count / scan / fill   per-body lists of incident contacts
//   schedule_kernel                 each list ordered by sweep position -> every contact's position in
//                                   its bodies' sequences (= the row versions it waits for)
//   run_scatter_kernel              runs of consecutive contacts on the same body A (exact order only)
//   ver_seed / solve_versioned / ver_finish
//                                   the solve (below)
//
// Two sweep orders (nans_world_set_solver):
//   NANS_SOLVER_EXACT    the reference's list order: results bit-identical to SolveConstraints.  The
//                        dependency depth of a pile is ~100 levels x one store -> poll -> apply hop.
//   NANS_SOLVER_SHUFFLED the same single Gauss-Seidel pass in a fixed pseudo-random order (a bijective
//                        hash of the contact index): a greedy colouring in random order has
//                        O(log n) colours instead of the list order's long vertical chains, so the
//                        same kernel runs ~10 levels deep.  Deterministic, atomics-free, but NOT the
//                        reference's order: velocities differ from the exact sweep wherever contacts
//                        share bodies (deviation measured by bench.py / tests).
//
// One world over several GPUs (slab.cu): a ghost body's rows are RELEASED to its owner rank by a
// peer store over NVLink when this rank applies its last contact on it; the owner's contacts poll
// their own rows as usual, so the cross-GPU dependency is carried by the same data flow.
#include <stdlib.h>
#include <string.h>

#include "nans_math.cuh"
namespace nans { __device__ unsigned int g_accum_fallbacks; }   // contacts whose accumulation fell back to the literal loop
#define NANS_ACCUM_ON_FALLBACK atomicAdd(&nans::g_accum_fallbacks, 1u)
#include "solver_accum.cuh"
#include "solver_constraint.cuh"
#include "world.cuh"

namespace nans {

// ---- sweep order -----------------------------------------------------------------------------
// shuffled order: sweep position of contact c = mix(c) over k bits (2^k >= n), contact at sweep
// position t = unmix(t).  mix is a bijection on k-bit integers (odd multiply, xor-shift, odd
// multiply), so every contact has exactly one position; positions >= n hold no contact.
constexpr uint32_t kMixA = 0x9E3779B1u, kMixB = 0x85EBCA6Bu;
constexpr uint32_t kMixAInv = 0x0E8B2F51u, kMixBInv = 0xA5CB9243u;   // inverses modulo 2^32 (hence modulo 2^k)
static_assert((uint32_t)(kMixA * kMixAInv) == 1u && (uint32_t)(kMixB * kMixBInv) == 1u, "modular inverses");

__host__ __device__ __forceinline__ uint32_t mix_bits(uint32_t x, int k)
{
    const uint32_t m = k >= 32 ? 0xffffffffu : ((1u << k) - 1u);
    const int s = (k + 1) / 2;
    x = (x * kMixA) & m;
    x ^= x >> s;
    x = (x * kMixB) & m;
    x ^= x >> s;
    return x;
}
__host__ __device__ __forceinline__ uint32_t unmix_bits(uint32_t x, int k)
{
    const uint32_t m = k >= 32 ? 0xffffffffu : ((1u << k) - 1u);
    const int s = (k + 1) / 2;
    x ^= x >> s;                       // s >= k/2: one application inverts the xor-shift
    x = (x * kMixBInv) & m;
    x ^= x >> s;
    x = (x * kMixAInv) & m;
    return x;
}
__device__ __forceinline__ int order_bits_for(int n)   // smallest k >= 1 with 2^k >= n
{
    return n <= 2 ? 1 : 32 - __clz(n - 1);
}

__global__ void __launch_bounds__(256) incidence_count_kernel(DeviceWorld w, int shuffled)
{
    const int n = w.counters->n_contacts;
    const int stride = gridDim.x * blockDim.x;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
        atomicAdd(&w.deg[w.c_a[c]], 1u);
        const int b = w.c_b[c];
        if (b >= 0) atomicAdd(&w.deg[b], 1u);
        // 1 where a run of contacts on the same body A starts (the shuffled sweep has no runs)
        w.run_flag[c] = (!shuffled && (c == 0 || w.c_a[c - 1] != w.c_a[c])) ? 1 : 0;
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

// one thread per body: order its incident contacts by sweep position; succ_a / succ_b receive the contact's
// POSITION in the body's sequence (the number of earlier contacts touching that body = the row version the
// contact waits for)
__global__ void __launch_bounds__(256) schedule_kernel(DeviceWorld w, int shuffled)
{
    const int body = blockIdx.x * blockDim.x + threadIdx.x;
    if (body >= w.nb) return;
    const uint32_t beg = w.deg[body], end = w.deg[body + 1];
    int32_t *l = w.inc + beg;
    const int n = (int)(end - beg);
    if (n == 0) return;
    const int k = order_bits_for(w.counters->n_contacts);
    auto key = [&](int c) -> uint32_t { return shuffled ? mix_bits((uint32_t)c, k) : (uint32_t)c; };
    for (int i = 1; i < n; ++i) {
        const int v = l[i];
        const uint32_t kv = key(v);
        int j = i;
        while (j > 0 && key(l[j - 1]) > kv) { l[j] = l[j - 1]; --j; }
        l[j] = v;
    }
    for (int q = 0; q < n; ++q) {
        const int c = l[q];
        if (w.c_a[c] == body) w.succ_a[c] = q; else w.succ_b[c] = q;
    }
}

// run r = contacts [run_start[r], run_start[r + 1]) = consecutive contacts with the same body A
__global__ void __launch_bounds__(256) run_scatter_kernel(DeviceWorld w)
{
    const int n = w.counters->n_contacts;
    const int stride = gridDim.x * blockDim.x;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
        const uint32_t r = w.pair_hit_scan[c];
        if (w.run_flag[c]) w.run_start[r] = c;
        if (c == n - 1) w.counters->frontier_n[1] = (int)r + w.run_flag[c];   // number of runs
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) w.counters->frontier_n[1] = 0;
}

// ---- versioned body rows -------------------------------------------------------------------------
// During the solve every body's velocity lives in a 16-byte row (V.xyz, version) and its angular velocity in a
// second row (W.xyz, version | level << 20), version = number of contacts applied to that body so far.  Contact
// c waits until both of its bodies show the versions it was scheduled for, applies itself and stores the rows
// with version + 1.  A row is ONE 128-bit access (ld/st.relaxed.{gpu,sys}.global.b128, single-copy atomic in
// the PTX memory model), so value and version arrive together: no fences, no in-degree atomics, no ready
// queue; the hop from a contact to its successor is one L2 store -> poll.
//
// Warps take sweep positions in order, 32 at a time.  Progress: every contact a lane waits for is earlier in
// the sweep, so its ticket is already held by a running warp (the grid is sized to be resident); by induction
// the earliest unfinished contact is always runnable.  A watchdog (no progress for seconds) turns any
// violation into a sticky error instead of a hang.
#ifndef NANS_VER_THREADS
#define NANS_VER_THREADS 256
#endif
#ifndef NANS_VER_MINBLOCKS
#define NANS_VER_MINBLOCKS 2   // resident CTAs the register budget is sized for (round-1 sweep: 256x2 0.374, 128x4 0.375,
#endif                         // 256x3 / 128x6 (80 registers, spills) 0.455, 256x4 (64 registers) 0.539 ms)
constexpr int kVerThreads = NANS_VER_THREADS;
constexpr int kVerMask = 0xfffff;    // version bits kept in the angular row (the rest carries the DAG level)
constexpr int kPendingTag = 0xffffe; // slab mode: the row is still owed by the lower neighbour rank
constexpr long long kWatchdogNs = 4000000000ll;   // no row arrived for 4 s of wall time: the schedule is broken

template <bool SYS>
__device__ __forceinline__ float4 ld_row(const float4 *p)
{
    unsigned long long lo, hi;
    if constexpr (SYS)
        asm volatile("{\n\t.reg .b128 r;\n\tld.relaxed.sys.global.b128 r, [%2];\n\tmov.b128 {%0, %1}, r;\n\t}"
                     : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
    else
        asm volatile("{\n\t.reg .b128 r;\n\tld.relaxed.gpu.global.b128 r, [%2];\n\tmov.b128 {%0, %1}, r;\n\t}"
                     : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
    return make_float4(__uint_as_float((unsigned)lo), __uint_as_float((unsigned)(lo >> 32)),
                       __uint_as_float((unsigned)hi), __uint_as_float((unsigned)(hi >> 32)));
}
template <bool SYS>
__device__ __forceinline__ void st_row(float4 *p, vec3 v, int tag)
{
    const unsigned long long lo = (unsigned long long)__float_as_uint(v.x) | ((unsigned long long)__float_as_uint(v.y) << 32);
    const unsigned long long hi = (unsigned long long)__float_as_uint(v.z) | ((unsigned long long)(unsigned)tag << 32);
    if constexpr (SYS)
        asm volatile("{\n\t.reg .b128 r;\n\tmov.b128 r, {%1, %2};\n\tst.relaxed.sys.global.b128 [%0], r;\n\t}"
                     :: "l"(p), "l"(lo), "l"(hi) : "memory");
    else
        asm volatile("{\n\t.reg .b128 r;\n\tmov.b128 r, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], r;\n\t}"
                     :: "l"(p), "l"(lo), "l"(hi) : "memory");
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// rows for the solve: (vel.xyz, 0) and (angvel.xyz, 0).  Slab mode: a row that was sent to the lower neighbour
// this step was marked PENDING before the halo left (slab.cu) and is released by that rank -- not seeded here;
// a ghost row no local contact touches is released to its owner at once.
__global__ void __launch_bounds__(256) ver_seed_kernel(DeviceWorld w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.nb) return;
    if (w.live && w.sent_mark[i]) return;
    float4 v = w.vel[i], a = w.angvel[i];
    v.w = 0.f; a.w = 0.f;
    w.row_v[i] = v;
    w.row_w[i] = a;
    if (w.peer_row_v && i >= w.n_owned && i < live_nb(w) && w.deg[i + 1] == w.deg[i]) {
        const int orow = w.ghost_owner_row[i];
        st_row<true>(w.peer_row_v + orow, V3(v), 0);
        st_row<true>(w.peer_row_w + orow, V3(a), 0);
    }
}
// back into vel / angvel.  Slab mode: a sent row may have no local contact at all, so nothing has waited for
// its release yet: wait here (the lower rank's solve is running concurrently on its own GPU).
__global__ void __launch_bounds__(256) ver_finish_kernel(DeviceWorld w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && w.counters->pad[1]) atomicOr(w.sticky, 1);   // the watchdog fired: reported by the next download / stats
    if (i >= w.nb) return;
    float4 v, a;
    if (w.live && w.sent_mark[i]) {
        const unsigned long long t0 = global_ns();
        while (true) {
            v = ld_row<true>(w.row_v + i); a = ld_row<true>(w.row_w + i);
            if (__float_as_int(v.w) != kPendingTag && (__float_as_int(a.w) & kVerMask) != kPendingTag) break;
            if ((long long)(global_ns() - t0) > kWatchdogNs || *(volatile int *)&w.counters->pad[1]) { w.counters->pad[1] = 1; atomicOr(w.sticky, 1); return; }
        }
    } else {
        v = w.row_v[i]; a = w.row_w[i];
        if (__float_as_int(v.w) == 0) return;        // untouched by any contact
    }
    float4 *pv = &w.vel[i], *pa = &w.angvel[i];
    pv->x = v.x; pv->y = v.y; pv->z = v.z;          // .w keeps 1/Mass, 1/MOI
    pa->x = a.x; pa->y = a.y; pa->z = a.z;
}

// One lane per RUN (exact order) or per contact (shuffled order): the lane keeps body A's velocity in registers
// along the run and only body B's rows go through memory, and the lanes of a warp fire together.
// trace (debug, NANS_SOLVER_TRACE=1): per contact {fire ns, stored ns, ticket ns, cycles}; trace_level = DAG level
template <bool SHUFFLED, bool SLAB>
__global__ void __launch_bounds__(kVerThreads, NANS_VER_MINBLOCKS) solve_versioned_kernel(DeviceWorld w, unsigned long long *trace)
{
    const float dt = *w.dt;
    const int n = w.counters->n_contacts;
    const int kbits = order_bits_for(n);
    const int n_tickets = SHUFFLED ? (n > 0 ? (int)min((unsigned long long)1 << kbits, (unsigned long long)0x7fffffff) : 0)
                                   : w.counters->frontier_n[1];
    const int lane = threadIdx.x & 31;
    int *head = &w.counters->frontier_n[0];
    volatile int *abort_flag = &w.counters->pad[1];
    const int32_t *run_start = w.run_start;
    float4 *sv = w.row_v, *sw = w.row_w;
    int max_level = 0;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(head, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_tickets) break;
        const int r = base + lane;
        bool pending = r < n_tickets;
        int c = 0, c_end = 0;
        if (pending) {
            if constexpr (SHUFFLED) {
                c = (int)unmix_bits((uint32_t)r, kbits);
                c_end = c + 1;
                pending = c < n;            // sweep positions without a contact
            } else {
                c = run_start[r];
                c_end = (r + 1 < n_tickets) ? run_start[r + 1] : n;
            }
        }
        float4 q[kRecQuads];
        int ia = 0, ib = -1, ea = 0, eb = 0, lva = 0;
        vec3 V1 = V3(0.f, 0.f, 0.f), W1 = V3(0.f, 0.f, 0.f);
        bool loaded = false, have_a = false, have_b = false;
        float4 vb4 = make_float4(0, 0, 0, 0), wb4 = make_float4(0, 0, 0, 0);
        int spins = 0;
        unsigned long long t_wait = 0, t_ticket = 0;
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
                if (need_a) { va = ld_row<SLAB>(sv + ia); wa = ld_row<SLAB>(sw + ia); }
                if (need_b) { vb4 = ld_row<SLAB>(sv + ib); wb4 = ld_row<SLAB>(sw + ib); }
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
                    bool remote = false;
                    if constexpr (SLAB) {
                        // body B is a ghost and this was this rank's LAST contact on it: its rows go to the owner
                        // rank (peer memory over NVLink) as version 0 of the owner's own sequence
                        if (ib >= w.n_owned && (uint32_t)(eb + 1) == w.deg[ib + 1] - w.deg[ib]) {
                            const int orow = w.ghost_owner_row[ib];
                            st_row<true>(w.peer_row_v + orow, V2, 0);
                            st_row<true>(w.peer_row_w + orow, W2, lva << 20);
                            remote = true;
                        }
                    }
                    if (!remote) {
                        st_row<SLAB>(sv + ib, V2, eb + 1);
                        st_row<SLAB>(sw + ib, W2, ((eb + 1) & kVerMask) | (lva << 20));
                    }
                }
                if (trace) {
                    trace[4 * (size_t)c] = t_fire;
                    trace[4 * (size_t)c + 1] = global_ns();
                    trace[4 * (size_t)c + 2] = t_ticket;
                    trace[4 * (size_t)c + 3] = (unsigned long long)(clock64() - ck);   // cycles: apply + row stores
                    w.trace_level[c] = lv;
                }
                ++ea; ++c;
                loaded = false;
                have_b = false;
                vb4 = make_float4(0, 0, 0, 0); wb4 = make_float4(0, 0, 0, 0);
                if (c == c_end) {          // body A leaves the run: publish it
                    st_row<SLAB>(sv + ia, V1, ea);
                    st_row<SLAB>(sw + ia, W1, (ea & kVerMask) | (lva << 20));
                    pending = false;
                }
            }
            if (go) {
                spins = 0; t_wait = 0;
            } else if ((++spins & 0x3ff) == 0) {
                // watchdog on PROGRESS, not on a poll count: a deep but healthy chain keeps firing somewhere in
                // this warp; only a warp that has seen nothing arrive for seconds of wall time gives up
                if (*abort_flag) return;
                const unsigned long long now = global_ns();
                if (t_wait == 0) t_wait = now;
                else if ((long long)(now - t_wait) > kWatchdogNs) { *abort_flag = 1; return; }
            }
        }
    }
    max_level = __reduce_max_sync(0xffffffffu, max_level);
    if (lane == 0 && max_level) atomicMax(&w.counters->solver_levels, max_level);
}

int launch_solver(World *w)
{
    DeviceWorld &d = w->d;
    if (d.nb == 0) return NANS_OK;
    cudaStream_t s = w->stream;
    static int sm_count = 0;
    if (!sm_count) {
        cudaDeviceProp prop;
        NANS_CUDA(cudaGetDeviceProperties(&prop, w->device));
        sm_count = prop.multiProcessorCount;
    }
    const bool shuffled = w->solver_mode == NANS_SOLVER_SHUFFLED;
    const bool slab = w->slab != nullptr;
    if (shuffled && slab) {
        snprintf(g_err, sizeof(g_err), "the shuffled sweep order is not available for a slab-partitioned world (exact order only)");
        return NANS_ERR_STATE;
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
    incidence_count_kernel<<<grid, 256, 0, s>>>(d, shuffled);
    NANS_LAUNCH_CHECK();
    int rc = exclusive_scan_u32(d.deg, d.deg, d.nb + 1, d.scan_block, s);
    if (rc) return rc;
    incidence_fill_kernel<<<grid, 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    schedule_kernel<<<div_up(d.nb, 256), 256, 0, s>>>(d, shuffled);
    NANS_LAUNCH_CHECK();

    static int ver_blocks[4] = {0, 0, 0, 0};
    const int variant = (shuffled ? 2 : 0) | (slab ? 1 : 0);
    if (!ver_blocks[variant]) {
        int per_sm = 0;
        if (variant == 0) NANS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_versioned_kernel<false, false>, kVerThreads, 0));
        else if (variant == 1) NANS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_versioned_kernel<false, true>, kVerThreads, 0));
        else NANS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_versioned_kernel<true, false>, kVerThreads, 0));
        const char *e;
        int want = (e = getenv("NANS_VER_BLOCKS")) ? atoi(e) : 2;
        if (want < 1) want = 1;
        ver_blocks[variant] = sm_count * (want < per_sm ? want : per_sm);   // every CTA must be resident (spinning lanes)
    }
    if (!shuffled) {
        rc = exclusive_scan_u32_dn((const uint32_t *)d.run_flag, d.pair_hit_scan, d.max_contacts, &d.counters->n_contacts, 0,
                                   d.scan_block, s);
        if (rc) return rc;
        run_scatter_kernel<<<grid, 256, 0, s>>>(d);
        NANS_LAUNCH_CHECK();
    }
    ver_seed_kernel<<<div_up(d.nb, 256), 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    static int ver_trace = -1;
    if (ver_trace < 0) ver_trace = getenv("NANS_SOLVER_TRACE") ? 1 : 0;
    unsigned long long *trace = ver_trace ? (unsigned long long *)d.pair_out : nullptr;
    if (variant == 0) solve_versioned_kernel<false, false><<<ver_blocks[0], kVerThreads, 0, s>>>(d, trace);
    else if (variant == 1) solve_versioned_kernel<false, true><<<ver_blocks[1], kVerThreads, 0, s>>>(d, trace);
    else solve_versioned_kernel<true, false><<<ver_blocks[2], kVerThreads, 0, s>>>(d, trace);
    NANS_LAUNCH_CHECK();
    ver_finish_kernel<<<div_up(d.nb, 256), 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

int solver_accum_fallbacks(World *w, int32_t *out)
{
    unsigned int v = 0;
    NANS_CUDA(cudaStreamSynchronize(w->stream));
    NANS_CUDA(cudaMemcpyFromSymbol(&v, g_accum_fallbacks, sizeof(v)));
    *out = (int32_t)v;
    return NANS_OK;
}

}  // namespace nans
