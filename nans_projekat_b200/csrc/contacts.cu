// contacts.cu — narrowphase hits -> the reference's contact list (sdl_state.Pairs, code/nans.h:385).
//
// Candidates were emitted in reference order, so an order-preserving compaction of the hit flags IS
// the reference list — except for one live quirk of DetectCollisions' "Exists" scans
// (code/nans.cpp:1479-1489): for CS pairs the second clause matches an earlier CS pair with the
// two indices swapped, so a hit (cube i, sphere j), j < i, is dropped when (cube j, sphere i) hit.
// (The same scans for CC/CF/SF/SS can never match: the list is cleared every frame, :1758.)
#include "world.cuh"

namespace nans {

constexpr int SEG_CS = 3;  // segment order CC, CF, SF, CS, SS (broadphase.cu)

__global__ void __launch_bounds__(256) cs_dedup_kernel(DeviceWorld w)
{
    const int n_pairs = w.counters->n_pairs;
    const int stride = gridDim.x * blockDim.x;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p <= n_pairs; p += stride) {
        if (p == n_pairs) break;
        const int a = w.pair_a[p], b = w.pair_b[p];
        if (!(a < w.n_cubes && b >= w.n_cubes) || !w.pair_hit[p]) continue;   // hit CS pairs only
        const int cube = a, sph = b - w.n_cubes;
        if (!(sph < cube) || sph >= w.n_cubes || cube >= w.n_spheres) continue;
        // candidate run of (cube = sph, *): offsets from the scanned (type, body) table
        const uint32_t beg = w.pair_count[(size_t)SEG_CS * w.nb + sph];
        const uint32_t end = w.pair_count[(size_t)SEG_CS * w.nb + sph + 1];
        const int want = w.n_cubes + cube;
        for (uint32_t q = beg; q < end && q < (uint32_t)n_pairs; ++q)
            if (w.pair_b[q] == want) {
                if (w.pair_hit[q]) w.pair_hit[p] = 0;   // (sph, cube) itself can never be dropped
                break;
            }
    }
}

__global__ void __launch_bounds__(256) contact_compact_kernel(DeviceWorld w)
{
    const int n_pairs = w.counters->n_pairs;
    const int stride = gridDim.x * blockDim.x;
    const uint32_t cap = (uint32_t)w.max_contacts;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += stride) {
        if (!w.pair_hit[p]) continue;
        const uint32_t c = w.pair_hit_scan[p];
        if (c >= cap) continue;
        w.c_a[c] = w.pair_a[p];
        w.c_b[c] = w.pair_b[p];
        const float4 *o = w.pair_out + 3 * (size_t)p;
        float4 qa = o[0], qb = o[1];
        qa.w = __int_as_float(w.pair_a[p]);   // the solver reads the body rows from the w lanes
        qb.w = __int_as_float(w.pair_b[p]);
        w.c_pa[c] = qa;
        w.c_pb[c] = qb;
        w.c_n[c] = o[2];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint32_t total = w.pair_hit_scan[n_pairs];
        w.counters->n_contacts = (int32_t)min(total, cap);
        if (total > cap) atomicOr(&w.counters->overflow, OVF_CONTACTS);
    }
}

int launch_contacts(World *w)
{
    DeviceWorld &d = w->d;
    if (d.nb == 0) return NANS_OK;
    cudaStream_t s = w->stream;
    const int grid = min(div_up(d.max_pairs + 1, 256), kNumSMs * 8);
    // a world without spheres has no CS pairs to drop; the sentinel behind the hit flags is written by pair_emit_kernel
    if (d.n_spheres > 0) {
        cs_dedup_kernel<<<grid, 256, 0, s>>>(d);
        NANS_LAUNCH_CHECK();
    }
    int rc = exclusive_scan_u32_dn((const uint32_t *)d.pair_hit, d.pair_hit_scan, d.max_pairs + 1,
                                   &d.counters->n_pairs, 1, d.scan_block, s);
    if (rc) return rc;
    contact_compact_kernel<<<grid, 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

}  // namespace nans
