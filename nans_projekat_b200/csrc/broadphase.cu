// broadphase.cu — candidate-pair generation.  The reference has NO broadphase: DetectCollisions
// (code/nans.cpp:1352-1536) runs GJK on every pair of the world, O(N^2).  This stage produces a
// superset of the pairs whose GJK can report a hit (AABB overlap, inclusive, inflated), emitted
// ALREADY IN THE REFERENCE'S LIST ORDER — CC (i<j lexicographic), CF (by cube), SF (by sphere),
// CS (cube-major), SS (i<j) — so that compaction of the narrowphase hits yields the reference's
// contact list with no sort of the contacts.
//
//   aabb_kernel          AABB per body (8 vertices / centre +- radius) + this step's largest extent; also clears the
//                        cell table, the cell sizes and the pair counts the later kernels accumulate into
//   key_kernel           30-bit Morton key of the cell (edge = largest extent) holding the AABB centre
//   cell_insert_kernel   counting sort, pass 1: the body's cell is found-or-inserted in the cell table (open
//                        addressing; slot = the Morton key itself while it fits the table, so slot order IS
//                        Morton order) and the body takes a rank inside the cell (one atomicAdd)
//   (exclusive scan)     cell sizes -> cell start offsets
//   cell_scatter_kernel  pass 2: AABBs into cell order as 32-byte records {lo.xyz,row}{hi.xyz,world} (neighbour
//                        scans read contiguous ranges); the table entry becomes {key, start, end}
//                        (round 1 sorted with a 3-4 pass LSD radix sort + a gather: 13 launches, ~190 us at 1 M
//                        bodies; the order INSIDE a cell is irrelevant -- the emit pass sorts every run)
//   pair_count_kernel    per body ONE half-neighbourhood scan: counts per (type, body) + partners parked in slots
//   (exclusive scan)     offsets in reference order
//   pair_emit_kernel     moves the parked partners to their offsets, sorts each short run ascending
//
// All kernels are HBM/L2-bound integer + compare work; see DESIGN.md §4 for bytes per body.
#include <float.h>

#include "world.cuh"

namespace nans {

constexpr uint32_t kEmptyKey = 0xffffffffu;
constexpr int kMaxKey = 3;   // Counters::pad slot: the step's largest Morton key

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t expand_bits10(uint32_t v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton30(uint32_t x, uint32_t y, uint32_t z)
{
    return (expand_bits10(x) << 2) | (expand_bits10(y) << 1) | expand_bits10(z);
}
__device__ __forceinline__ uint32_t compact_bits10(uint32_t v)
{
    v &= 0x09249249u;
    v = (v | (v >> 2)) & 0x030c30c3u;
    v = (v | (v >> 4)) & 0x0300f00fu;
    v = (v | (v >> 8)) & 0x030000ffu;
    v = (v | (v >> 16)) & 0x3ffu;
    return v;
}

// Conservative inflation: GJK accepts touching shapes (AddSupport uses >=, code/nans.cpp:528) and a
// sphere support Radius*normalize(d) can exceed Radius by rounding, so boxes are grown by an
// absolute + relative margin.  Only a superset is required; the final arbiter is GJK itself.
__device__ __forceinline__ void inflate(float &lo, float &hi)
{
    const float m = 1e-3f + 1e-5f * fmaxf(fabsf(lo), fabsf(hi));
    lo -= m;
    hi += m;
}

__device__ __forceinline__ void box_aabb(const float4 *v6, float lo[3], float hi[3])
{
    float f[24];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const float4 t = v6[q];
        f[4 * q] = t.x; f[4 * q + 1] = t.y; f[4 * q + 2] = t.z; f[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) { lo[r] = f[r]; hi[r] = f[r]; }
#pragma unroll
    for (int k = 1; k < 8; ++k)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            lo[r] = fminf(lo[r], f[3 * k + r]);
            hi[r] = fmaxf(hi[r], f[3 * k + r]);
        }
#pragma unroll
    for (int r = 0; r < 3; ++r) inflate(lo[r], hi[r]);
}

__device__ __forceinline__ int cell_coord(float c, float inv_cell)
{
    // floor(c / cell) biased to [0, 1023]; clamping is monotone, so AABB-overlapping bodies stay in
    // adjacent cells even outside the addressable volume (only performance degrades there)
    float f = floorf(c * inv_cell);
    f = fminf(fmaxf(f, -512.0f), 511.0f);
    return (int)f + 512;   // NaN -> (int)NaN = 0 -> 512
}

// monotone map float -> uint32 (total order of the finite floats)
__device__ __forceinline__ uint32_t order_bits(float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float order_float(uint32_t e)
{
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// AABB per body + the largest AABB extent of this step (block reduce -> one atomicMax per block).
__global__ void __launch_bounds__(256) aabb_kernel(DeviceWorld w, int clear)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (clear) {
        // what the later kernels of the stage accumulate into, cleared here: coalesced 16-byte / 4-byte stores spread
        // over the grid instead of three memset nodes between the kernels (each a launch gap on the step's critical
        // path): broadphase stage 0.287 -> 0.275 ms on the 1 M-cube pile (profiles/r4_ab_*.json)
        const size_t nthreads = (size_t)gridDim.x * blockDim.x;
        const size_t table = (size_t)w.cell_mask + 1;
        for (size_t s = (size_t)i; s < table; s += nthreads) {
            w.cell_tab[s] = make_uint4(kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey);
            w.cell_count[s] = 0u;
        }
        if (i == 0) w.cell_count[table] = 0u;
        const size_t npc = (size_t)w.n_seg * w.nb + 1;
        for (size_t s = (size_t)i; s < npc; s += nthreads) w.pair_count[s] = 0u;
    }
    if (i < w.n_statics) {   // statics: AABB only (tested against every body, never sorted)
        float lo[3], hi[3];
        box_aabb(w.st_verts + 6 * i, lo, hi);
        w.st_aabb[2 * i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        w.st_aabb[2 * i + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    float ext = 0.f;
    float ctr[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    if (i < live_nb(w)) {
        float lo[3], hi[3];
        if (i < w.n_cubes) {
            box_aabb(w.verts + 6 * (size_t)i, lo, hi);
        } else {
            const float4 p = w.pos[i];
            const float r = w.scale[i].w;
            lo[0] = p.x - r; lo[1] = p.y - r; lo[2] = p.z - r;
            hi[0] = p.x + r; hi[1] = p.y + r; hi[2] = p.z + r;
#pragma unroll
            for (int k = 0; k < 3; ++k) inflate(lo[k], hi[k]);
        }
        w.aabb_lo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        w.aabb_hi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
        ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);   // fmaxf drops NaN
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float c = 0.5f * (lo[k] + hi[k]);                           // the expression key_kernel uses
            if (isfinite(c)) ctr[k] = c;
        }
    }
    // non-negative floats order like their bit patterns
    ext = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(ext, 0.f))));
    // smallest AABB centre per axis: the cell grid is anchored there, so the Morton keys start at 0
    // and the sort can skip the passes above the world's real extent
    uint32_t cmin[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) cmin[k] = __reduce_max_sync(0xffffffffu, ~order_bits(ctr[k]));
    __shared__ float wmax[8];
    __shared__ uint32_t wmin[8][3];
    if ((threadIdx.x & 31) == 0) {
        wmax[threadIdx.x >> 5] = ext;
#pragma unroll
        for (int k = 0; k < 3; ++k) wmin[threadIdx.x >> 5][k] = cmin[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = wmax[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) m = fmaxf(m, wmax[k]);
        atomicMax((unsigned int *)&w.counters->pad[2], __float_as_uint(m));   // non-negative floats order as uints
    }
    if (threadIdx.x < 3) {
        uint32_t m = wmin[0][threadIdx.x];
#pragma unroll
        for (int k = 1; k < 8; ++k) m = max(m, wmin[k][threadIdx.x]);
        unsigned int *slot = (unsigned int *)&w.counters->pad[kMinCentre + threadIdx.x];
        if (m > *(volatile unsigned int *)slot) atomicMax(slot, m);       // most blocks do not improve it
    }
}

// Morton cell key of the AABB centre.  The cell edge is this step's largest AABB extent (so that
// AABB-overlapping bodies always sit in adjacent cells) — far tighter than the static bound (box
// diagonal) while the bodies are near axis-aligned; the static bound is the fallback if it is not finite.
__global__ void __launch_bounds__(256) key_kernel(DeviceWorld w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= live_nb(w)) return;
    float cell = __uint_as_float((unsigned int)w.counters->pad[2]) * 1.0001f + 1e-4f;
    if (!isfinite(cell)) cell = w.cell_size;             // overflowed vertices: static bound (box diagonal)
    cell = fmaxf(cell, 0.05f);
    const float inv_cell = 1.0f / cell;
    const float4 lo = w.aabb_lo[i], hi = w.aabb_hi[i];
    int cx = cell_coord(0.5f * (lo.x + hi.x), inv_cell);
    int cy = cell_coord(0.5f * (lo.y + hi.y), inv_cell);
    int cz = cell_coord(0.5f * (lo.z + hi.z), inv_cell);
    // anchor the grid at the smallest centre (cell_coord is monotone, so the differences stay >= 0)
    const int bx = cell_coord(order_float(~(uint32_t)w.counters->pad[kMinCentre + 0]), inv_cell);
    const int by = cell_coord(order_float(~(uint32_t)w.counters->pad[kMinCentre + 1]), inv_cell);
    const int bz = cell_coord(order_float(~(uint32_t)w.counters->pad[kMinCentre + 2]), inv_cell);
    cy = max(cy - by, 0);
    if (w.world_id) {
        // batched independent worlds: each world owns a 16x16 column of cells in x,z
        const int wid = w.world_id[i];
        cx = min(max(cx - 512 + 8, 0), 15) + 16 * (wid & 63);
        cz = min(max(cz - 512 + 8, 0), 15) + 16 * ((wid >> 6) & 63);
    } else {
        cx = max(cx - bx, 0);
        cz = max(cz - bz, 0);
    }
    const uint32_t key = morton30((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
    w.key[0][i] = key;
    // the largest key of the step selects the cell table's addressing mode (cell_probe)
    const uint32_t m = __reduce_max_sync(__activemask(), key);
    unsigned int *slot = (unsigned int *)&w.counters->pad[kMaxKey];
    if ((threadIdx.x & 31) == 0 && m > *(volatile unsigned int *)slot) atomicMax(slot, m);   // rarely improves it
}

// ---------------------------------------------------------------------------------------------
// Counting sort by cell.  The cell table is open-addressed, in one of two modes chosen per step from the largest
// key (kMaxKey, written by key_kernel):
//   direct   every key fits the table: a key IS its slot (keys are anchored at the world's smallest cell, so a
//            100^3-cube pile has 18-bit keys against a 2^21-slot table): no probing, slot order is Morton order;
//   blocked  larger keys (wide or batched worlds, up to 30 bits): blocks of 8 Morton-consecutive cells (2 x 2 x 2)
//            keep their inner position and the BLOCK index is hashed; a collision probes the next block (stride 8),
//            so the table behaves like linear probing over blocks at load <= 0.5 (~1.5 probes) while neighbouring
//            cells stay neighbours in memory.  A stride-8 walk only sees one slot in eight, which a lopsided world
//            (every cell at the same position of its block) can fill up: after kBlockProbes steps the walk goes on
//            with stride 1, which terminates because the table is at most half full.  (Mixing the two per key -- direct where the key fits, hashed
//            otherwise -- left the direct region locally 90 % full: 646 probes on average, 23 ms, on the
//            250 x 16 x 250-cell pile.)
constexpr uint32_t kBlockProbes = 16;
struct CellProbe { uint32_t slot, stride; };
__device__ __forceinline__ uint32_t probe_next(uint32_t slot, uint32_t stride, uint32_t &n, uint32_t mask)
{
    return (slot + (++n <= kBlockProbes ? stride : 1u)) & mask;
}
__device__ __forceinline__ CellProbe cell_probe(const DeviceWorld &w, uint32_t key)
{
    CellProbe p;
    if ((uint32_t)w.counters->pad[kMaxKey] <= w.cell_mask) { p.slot = key; p.stride = 1u; return p; }
    uint32_t h = (key >> 3) * 0x9E3779B1u;
    h ^= h >> 15;
    p.slot = ((h << 3) | (key & 7u)) & w.cell_mask;
    p.stride = 8u;
    return p;
}

__global__ void __launch_bounds__(256) cell_insert_kernel(DeviceWorld w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= live_nb(w)) return;
    const uint32_t key = w.key[0][i];
    const CellProbe pr = cell_probe(w, key);
    uint32_t slot = pr.slot, np_ = 0;
    while (true) {
        const uint32_t prev = atomicCAS(&w.cell_tab[slot].x, kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) break;
        slot = probe_next(slot, pr.stride, np_, w.cell_mask);
    }
    w.val[0][i] = slot;
    w.val[1][i] = atomicAdd(&w.cell_count[slot], 1u);     // rank inside the cell (arbitrary, see the emit pass)
}

__global__ void __launch_bounds__(256) cell_scatter_kernel(DeviceWorld w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= live_nb(w)) return;
    const uint32_t slot = w.val[0][i], rank = w.val[1][i];
    const uint32_t start = w.cell_count[slot];            // scanned: first sorted position of the cell
    const uint32_t t = start + rank;
    float4 lo = w.aabb_lo[i], hi = w.aabb_hi[i];
    lo.w = __int_as_float(i);
    hi.w = __int_as_float(w.world_id ? w.world_id[i] : 0);
    w.sbox[2 * (size_t)t] = lo;
    w.sbox[2 * (size_t)t + 1] = hi;
    w.key[1][t] = w.key[0][i];                            // keys in cell order
    w.pair_fill[t] = 0u;                                  // pair-slot cursors of the counting pass
    if (rank == 0) {
        w.cell_tab[slot].y = start;
        w.cell_tab[slot].z = w.cell_count[slot + 1];
    }
}

__device__ __forceinline__ bool cell_lookup(const DeviceWorld &w, uint32_t key, uint32_t &start, uint32_t &end)
{
    const CellProbe pr = cell_probe(w, key);
    uint32_t slot = pr.slot, np_ = 0;
    while (true) {
        const uint4 e = __ldg(&w.cell_tab[slot]);
        if (e.x == key) { start = e.y; end = e.z; return true; }
        if (e.x == kEmptyKey) return false;
        slot = probe_next(slot, pr.stride, np_, w.cell_mask);
    }
}

__device__ __forceinline__ bool overlap(const float4 &alo, const float4 &ahi, const float4 &blo, const float4 &bhi)
{
    return alo.x <= bhi.x && blo.x <= ahi.x && alo.y <= bhi.y && blo.y <= ahi.y && alo.z <= bhi.z && blo.z <= ahi.z;
}

// segment order = reference list order (code/nans.cpp:1355-1535)
enum { SEG_CC = 0, SEG_CF = 1, SEG_SF = 2, SEG_CS = 3, SEG_SS = 4 };

// Visit every dynamic partner of sorted entry t that the reference would list under body `row`.
// F(seg, partner_row)
template <typename F>
__device__ __forceinline__ void for_each_partner(const DeviceWorld &w, const uint32_t *__restrict__ keys, int t, F f)
{
    const float4 alo = w.sbox[2 * (size_t)t], ahi = w.sbox[2 * (size_t)t + 1];
    const int row = __float_as_int(alo.w);
    const int wid = __float_as_int(ahi.w);
    const bool a_cube = row < w.n_cubes;
    const uint32_t key = keys[t];
    const int cx = (int)compact_bits10(key >> 2), cy = (int)compact_bits10(key >> 1), cz = (int)compact_bits10(key);
    // the 27 neighbour keys are combinations of 3 x 3 pre-expanded coordinates (9 bit expansions, not 81)
    const uint32_t ex0 = expand_bits10((uint32_t)(cx - 1)) << 2, ex1 = expand_bits10((uint32_t)cx) << 2,
                   ex2 = expand_bits10((uint32_t)(cx + 1)) << 2;
    const uint32_t ey0 = expand_bits10((uint32_t)(cy - 1)) << 1, ey1 = expand_bits10((uint32_t)cy) << 1,
                   ey2 = expand_bits10((uint32_t)(cy + 1)) << 1;
    const uint32_t ez0 = expand_bits10((uint32_t)(cz - 1)), ez1 = expand_bits10((uint32_t)cz),
                   ez2 = expand_bits10((uint32_t)(cz + 1));
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int nx = cx + dx, ny = cy + dy, nz = cz + dz;
                if ((unsigned)nx > 1023u || (unsigned)ny > 1023u || (unsigned)nz > 1023u) continue;
                const uint32_t nkey = (dx < 0 ? ex0 : dx == 0 ? ex1 : ex2) | (dy < 0 ? ey0 : dy == 0 ? ey1 : ey2) |
                                      (dz < 0 ? ez0 : dz == 0 ? ez1 : ez2);
                uint32_t s, e;
                if (!cell_lookup(w, nkey, s, e)) continue;
                for (uint32_t u = s; u < e; ++u) {
                    const float4 blo = __ldg(&w.sbox[2 * (size_t)u]), bhi = __ldg(&w.sbox[2 * (size_t)u + 1]);
                    const int brow = __float_as_int(blo.w);
                    if (__float_as_int(bhi.w) != wid) continue;
                    const bool b_cube = brow < w.n_cubes;
                    int seg;
                    if (a_cube) {
                        if (b_cube) { if (brow <= row) continue; seg = SEG_CC; }
                        else seg = SEG_CS;                       // cube-major: listed under the cube
                    } else {
                        if (b_cube || brow <= row) continue;     // CS is emitted by the cube
                        seg = SEG_SS;
                    }
                    if (!overlap(alo, ahi, blo, bhi)) continue;
                    f(seg, brow);
                }
            }
}

// The counting pass yields counts per (type, body) AND the partners themselves, parked in a
// fixed-stride slot buffer (kPairSlots per body, tagged seg<<28 | row).  The emit pass only moves
// them to their scanned offsets; a body with more partners than slots (crowded cell) is rescanned
// there with the full 27-cell walk (for_each_partner).
constexpr int kPairSlots = 24;

#ifndef NANS_PC_PAIRWISE
#define NANS_PC_PAIRWISE 1   // extra candidate records in flight per thread
#endif
#ifndef NANS_PC_PIPELINE
#define NANS_PC_PIPELINE 1
#endif
// (Round 2 also measured SEVERAL threads per body, each walking every 2nd / 4th / 7th neighbour cell: broadphase stage
// 0.283 -> 0.298 / 0.330 / 0.358 ms on the 1 M-cube pile.  The kernel is bound by issued instructions and L2 requests,
// not by the length of one body's probe chain; one thread per body stays.)
#ifndef NANS_PC_MINBLOCKS
#define NANS_PC_MINBLOCKS 10   // 47 registers with two candidate records in flight (sweep, broadphase stage: w2/b10 0.415, w2/b9 0.413, w2/b12 0.445 (spills), w3/b10 0.448, w3/b8 0.437, w4/b8 0.437, w4/b6 0.434, w1/b12 0.452 ms)
#endif

// Every unordered pair of dynamic bodies is looked at ONCE: a body scans the rest of its own cell
// (sorted positions after its own) and the 13 neighbour cells that follow its cell in (z, y, x)
// order — half the hash probes and half the AABB loads of a full 27-cell scan.  A hit is credited
// to the body the reference lists it under (lower cube / the cube of a cube-sphere pair / lower
// sphere), whichever of the two found it: per-(type, body) counts and the slot cursor are atomics,
// the slot order is arbitrary (the emit pass sorts every run by partner).
__global__ void __launch_bounds__(128, NANS_PC_MINBLOCKS) pair_count_kernel(DeviceWorld w)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nb_live = live_nb(w);
    if (t >= nb_live) return;
    const uint32_t *__restrict__ keys = w.key[1];
    uint32_t *__restrict__ fill = w.pair_fill;
    const float4 alo = w.sbox[2 * (size_t)t], ahi = w.sbox[2 * (size_t)t + 1];
    const int row = __float_as_int(alo.w);
    const int wid = __float_as_int(ahi.w);
    const bool a_cube = row < w.n_cubes;
    const uint32_t key = keys[t];

    auto test = [&](uint32_t u, const float4 &blo, const float4 &bhi) {
        if (__float_as_int(bhi.w) != wid) return;
        if (!overlap(alo, ahi, blo, bhi)) return;
        const int brow = __float_as_int(blo.w);
        const bool b_cube = brow < w.n_cubes;
        // owner = the body the pair is listed under; cube-sphere pairs are cube-major
        const bool mine = (a_cube == b_cube) ? row < brow : a_cube;
        const int orow = mine ? row : brow, prow = mine ? brow : row;
        const uint32_t ot = mine ? (uint32_t)t : u;
        if (orow >= w.n_owned) return;          // slab mode: a ghost's pairs belong to its owner rank
        const int seg = (a_cube && b_cube) ? SEG_CC : (a_cube != b_cube) ? SEG_CS : SEG_SS;
        atomicAdd(&w.pair_count[(size_t)seg * w.nb + orow], 1u);
        const uint32_t slot = atomicAdd(&fill[ot], 1u);
        if (slot < (uint32_t)kPairSlots) w.pair_tmp[(size_t)ot * kPairSlots + slot] = ((uint32_t)seg << 28) | (uint32_t)prow;
    };
    auto visit = [&](uint32_t u) {
        const float4 blo = __ldg(&w.sbox[2 * (size_t)u]), bhi = __ldg(&w.sbox[2 * (size_t)u + 1]);
        test(u, blo, bhi);
    };
    // candidates [s, e): NANS_PC_PAIRWISE + 1 records in flight
    auto visit_range = [&](uint32_t s, uint32_t e) {
#if NANS_PC_PAIRWISE
        constexpr int W = NANS_PC_PAIRWISE + 1;       // records in flight
        for (uint32_t u = s; u < e; u += W) {
            float4 lo[W], hi[W];
#pragma unroll
            for (int j = 0; j < W; ++j)
                if (j == 0 || u + j < e) { lo[j] = __ldg(&w.sbox[2 * (size_t)(u + j)]); hi[j] = __ldg(&w.sbox[2 * (size_t)(u + j) + 1]); }
#pragma unroll
            for (int j = 0; j < W; ++j)
                if (j == 0 || u + j < e) test(u + j, lo[j], hi[j]);
        }
#else
        for (uint32_t u = s; u < e; ++u) visit(u);
#endif
    };

    // the rest of the own cell
    for (uint32_t u = (uint32_t)t + 1; u < (uint32_t)nb_live && keys[u] == key; ++u) visit(u);
    // the 13 cells after it in (z, y, x) order.  Neighbour keys by arithmetic on the interleaved key itself
    // (x lives in bits 2, 5, ..., y in 1, 4, ..., z in 0, 3, ...): +1 on a field = add its lowest bit with the
    // other fields' bits set so the carry runs through them, -1 = subtract with them cleared; a field of all
    // ones / all zeros is the grid border.  (Formerly compact_bits + expand_bits per axis and k % 3, k / 3 per
    // cell: 53 % of this kernel's issued instructions, profiles/README.md.)
    constexpr uint32_t kMx = 0x24924924u, kMy = 0x12492492u, kMz = 0x09249249u;
    const uint32_t x0 = key & kMx, y0 = key & kMy, z0 = key & kMz;
    const uint32_t xm = (x0 - 4u) & kMx, xp = ((key | ~kMx) + 4u) & kMx;
    const uint32_t ym = (y0 - 2u) & kMy, yp = ((key | ~kMy) + 2u) & kMy;
    const uint32_t zp = ((key | ~kMz) + 1u) & kMz;
    const bool xm_ok = x0 != 0u, xp_ok = x0 != kMx, ym_ok = y0 != 0u, yp_ok = y0 != kMy, zp_ok = z0 != kMz;
    // 4 bits per cell: (dx + 1) | (dy + 1) << 2; cells 0..3 have dz = 0 (+x; then the row y + 1), cells 4..12 dz = +1
    constexpr unsigned long long kCells = 0xa98654210a986ull;
    // the hash probe of cell k + 1 is issued before cell k's candidates are tested (NANS_PC_PIPELINE), so its
    // L2 round trip hides behind their loads
    auto cell_key = [&](int k, uint32_t &nkey) -> bool {
        const uint32_t c = (uint32_t)(kCells >> (4 * k)) & 15u;
        const uint32_t sx = c & 3u, sy = c >> 2;
        const bool up = k >= 4;
        nkey = (sx == 0u ? xm : sx == 1u ? x0 : xp) | (sy == 0u ? ym : sy == 1u ? y0 : yp) | (up ? zp : z0);
        return (sx == 0u ? xm_ok : sx == 2u ? xp_ok : true) && (sy == 0u ? ym_ok : sy == 2u ? yp_ok : true) && (!up || zp_ok);
    };
#if NANS_PC_PIPELINE
    const bool direct = (uint32_t)w.counters->pad[kMaxKey] <= w.cell_mask;    // cell_probe's mode, hoisted
    const uint32_t stride = direct ? 1u : 8u;
    auto slot_of = [&](uint32_t k_) -> uint32_t {
        if (direct) return k_;
        uint32_t h = (k_ >> 3) * 0x9E3779B1u;
        h ^= h >> 15;
        return ((h << 3) | (k_ & 7u)) & w.cell_mask;
    };
    uint32_t nkey, slot = 0;
    bool ok = cell_key(0, nkey);
    uint4 ent = make_uint4(kEmptyKey, 0u, 0u, 0u);
    if (ok) { slot = slot_of(nkey); ent = __ldg(&w.cell_tab[slot]); }
#pragma unroll 1
    for (int k = 0; k < 13; ++k) {
        uint32_t nkey2 = 0, slot2 = 0;
        bool ok2 = false;
        uint4 ent2 = make_uint4(kEmptyKey, 0u, 0u, 0u);
        if (k + 1 < 13) {
            ok2 = cell_key(k + 1, nkey2);
            if (ok2) { slot2 = slot_of(nkey2); ent2 = __ldg(&w.cell_tab[slot2]); }
        }
        if (ok) {
            uint32_t np_ = 0;
            while (ent.x != nkey && ent.x != kEmptyKey) {       // probing, as cell_lookup
                slot = probe_next(slot, stride, np_, w.cell_mask);
                ent = __ldg(&w.cell_tab[slot]);
            }
            if (ent.x == nkey) visit_range(ent.y, ent.z);
        }
        nkey = nkey2; slot = slot2; ok = ok2; ent = ent2;
    }
#else
#pragma unroll 1
    for (int k = 0; k < 13; ++k) {
        uint32_t nkey;
        if (!cell_key(k, nkey)) continue;
        uint32_t s, e;
        if (!cell_lookup(w, nkey, s, e)) continue;
        for (uint32_t u = s; u < e; ++u) visit(u);
    }
#endif
    // statics: CF for cubes, SF for spheres, static index ascending (recomputed in the emit pass);
    // only the body itself writes these two segments
    if (row < w.n_owned) {
        uint32_t ns = 0;
        for (int k = 0; k < w.n_statics; ++k)
            if (overlap(alo, ahi, w.st_aabb[2 * k], w.st_aabb[2 * k + 1])) ++ns;
        w.pair_count[(size_t)(a_cube ? SEG_CF : SEG_SF) * w.nb + row] = ns;
    }
}

__global__ void __launch_bounds__(128) pair_emit_kernel(DeviceWorld w)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= live_nb(w)) return;
    const uint32_t *__restrict__ keys = w.key[1];
    const float4 alo = w.sbox[2 * (size_t)t], ahi = w.sbox[2 * (size_t)t + 1];
    const int row = __float_as_int(alo.w);
    if (row >= w.n_owned) {
        if (t == 0) {
            const uint32_t total = w.pair_count[(size_t)w.n_seg * w.nb];
            w.counters->n_pairs = (int32_t)min(total, (uint32_t)w.max_pairs);
            w.pair_hit[min(total, (uint32_t)w.max_pairs)] = 0;    // sentinel: the scan of the hit flags yields the total
            if (total > (uint32_t)w.max_pairs) atomicOr(&w.counters->overflow, OVF_PAIRS);
        }
        return;
    }
    uint32_t off[5], fill[5] = {0, 0, 0, 0, 0};
    // scanned table: offsets; the entry after (s,row) is the next run's start, i.e. this run's end
#pragma unroll
    for (int s = 0; s < 5; ++s) off[s] = w.pair_count[(size_t)s * w.nb + row];
    const uint32_t n_dyn = (w.pair_count[(size_t)SEG_CC * w.nb + row + 1] - off[SEG_CC]) +
                           (w.pair_count[(size_t)SEG_CS * w.nb + row + 1] - off[SEG_CS]) +
                           (w.pair_count[(size_t)SEG_SS * w.nb + row + 1] - off[SEG_SS]);
    const uint32_t cap = (uint32_t)w.max_pairs;
    auto put = [&](int seg, int brow) {
        const uint32_t p = off[seg] + fill[seg]++;
        if (p < cap) { w.pair_a[p] = row; w.pair_b[p] = brow; }
    };
    if (n_dyn <= (uint32_t)kPairSlots) {
        const uint32_t *slots = w.pair_tmp + (size_t)t * kPairSlots;
        for (uint32_t k = 0; k < n_dyn; ++k) {
            const uint32_t e = slots[k];
            put((int)(e >> 28), (int)(e & 0x0fffffffu));
        }
    } else {
        for_each_partner(w, keys, t, put);
    }
    const int sseg = row < w.n_cubes ? SEG_CF : SEG_SF;
    for (int k = 0; k < w.n_statics; ++k)
        if (overlap(alo, ahi, w.st_aabb[2 * k], w.st_aabb[2 * k + 1])) put(sseg, -(k + 1));
    // each (body, type) run in ascending partner order = the reference's inner loop order
    // (insertion sort in place; runs are a handful of entries)
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        if (s == SEG_CF || s == SEG_SF) continue;
        const uint32_t b = off[s], n = fill[s];
        if (b + n > cap) continue;
        for (uint32_t i = 1; i < n; ++i) {
            const int v = w.pair_b[b + i];
            uint32_t j = i;
            while (j > 0 && w.pair_b[b + j - 1] > v) { w.pair_b[b + j] = w.pair_b[b + j - 1]; --j; }
            w.pair_b[b + j] = v;
        }
    }
    if (t == 0) {
        const uint32_t total = w.pair_count[(size_t)w.n_seg * w.nb];
        w.counters->n_pairs = (int32_t)min(total, cap);
        w.pair_hit[min(total, cap)] = 0;                          // sentinel: the scan of the hit flags yields the total
        if (total > cap) atomicOr(&w.counters->overflow, OVF_PAIRS);
    }
}

// AABBs only (slab mode needs them before the halo exchange)
int launch_aabb_only(World *w)
{
    DeviceWorld &d = w->d;
    if (d.nb == 0) return NANS_OK;
    aabb_kernel<<<div_up(max(d.nb, d.n_statics), 256), 256, 0, w->stream>>>(d, 0);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

// ---------------------------------------------------------------------------------------------
int launch_broadphase(World *w)
{
    DeviceWorld &d = w->d;
    cudaStream_t s = w->stream;
    const int nb = d.nb;
    NANS_CUDA(cudaMemsetAsync(d.counters, 0, sizeof(Counters), s));
    if (nb == 0) return NANS_OK;
    aabb_kernel<<<div_up(max(nb, d.n_statics), 256), 256, 0, s>>>(d, 1);
    NANS_LAUNCH_CHECK();
    key_kernel<<<div_up(nb, 256), 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();

    // counting sort by cell (table + cell sizes were cleared by aabb_kernel): insert, scan, scatter
    const size_t table = (size_t)d.cell_mask + 1;
    cell_insert_kernel<<<div_up(nb, 256), 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    int rc = exclusive_scan_u32(d.cell_count, d.cell_count, (int)table + 1, d.scan_block, s);
    if (rc) return rc;
    cell_scatter_kernel<<<div_up(nb, 256), 256, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    // counts per (type, body) are accumulated with atomics (cleared by aabb_kernel); a cube-only world only has the CC
    // and CF segments
    pair_count_kernel<<<div_up(nb, 128), 128, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    rc = exclusive_scan_u32(d.pair_count, d.pair_count, d.n_seg * nb + 1, d.scan_block, s);
    if (rc) return rc;
    pair_emit_kernel<<<div_up(nb, 128), 128, 0, s>>>(d);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

}  // namespace nans
