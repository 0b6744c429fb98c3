// scan.cu — exclusive prefix sum over uint32 (pair offsets, hit compaction, incidence offsets).
// Default: single pass with decoupled look-back (8 B/element, one launch).  The first version —
// three phases: per-tile reduce -> single-block scan of tile sums -> per-tile scan + offset
// (12 B/element, three launches) — is kept for A/B runs (NANS_SCAN_3PASS=1).
#include <stdlib.h>

#include "world.cuh"

namespace nans {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems; // 4096

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t block_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0u;
        uint32_t si = warp_incl_scan(s);
        warp_sums[lane] = si - s;
        if (lane == 31) block_total = si;
    }
    __syncthreads();
    uint32_t r = incl - v + warp_sums[wid];
    *total = block_total;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t *__restrict__ in, int n,
                                                                   uint32_t *__restrict__ tile_sums,
                                                                   const int32_t *__restrict__ d_n, int extra)
{
    if (d_n) n = min(n, *d_n + extra);   // device-resident length: tiles past it contribute zero
    const int base = blockIdx.x * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int i = base + k * kScanThreads + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(uint32_t *tile_sums, int n_tiles)
{
    // single block: serial over chunks of blockDim
    uint32_t carry = 0;
    for (int base = 0; base < n_tiles; base += kScanThreads) {
        int i = base + threadIdx.x;
        uint32_t v = i < n_tiles ? tile_sums[i] : 0u;
        uint32_t total;
        uint32_t ex = block_excl_scan(v, &total);
        if (i < n_tiles) tile_sums[i] = ex + carry;
        carry += total;
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t *__restrict__ in,
                                                                  uint32_t *__restrict__ out, int n,
                                                                  const uint32_t *__restrict__ tile_sums,
                                                                  const int32_t *__restrict__ d_n, int extra)
{
    if (d_n) n = min(n, *d_n + extra);
    if (blockIdx.x * kScanTile >= n) return;
    // each thread owns kScanItems CONSECUTIVE elements so the in-thread scan is sequential
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int i = base + k;
        v[k] = i < n ? in[i] : 0u;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_excl_scan(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int i = base + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
}

// ---- single pass: decoupled look-back ------------------------------------------------------------
// One kernel per scan, n read once and written once (8 B/element).  Tiles take their index from a
// ticket counter (so every tile a block waits on is already running), publish {epoch, status, value}
// as ONE 64-bit word (status 1 = tile aggregate, 2 = inclusive prefix) and look back over their
// predecessors 32 descriptors at a time.  The epoch lives in device memory and is advanced by the
// last block to finish, so stale descriptors never need clearing and the launch replays unchanged
// inside a CUDA graph.  scratch: [0] ticket, [1] epoch, [2] finished blocks, [4..] descriptors.
// (Round 2 measured a FLAT look-back for scans of up to 1024 tiles -- every block sums the aggregates of all its
// predecessors in one sweep instead of walking back 32 at a time: 0.943 vs 0.940 ms per step, no difference; the
// ~11 us of a 2 M-element scan is not the look-back chain.  Removed.)
#ifndef NANS_LB_THREADS
#define NANS_LB_THREADS 256
#endif
#ifndef NANS_LB_ITEMS
#define NANS_LB_ITEMS 16
#endif
constexpr int kLbThreads = NANS_LB_THREADS;
constexpr int kLbItems = NANS_LB_ITEMS;
constexpr int kLbTile = kLbThreads * kLbItems;   // 4096

__device__ __forceinline__ unsigned long long lb_pack(uint32_t epoch, uint32_t status, uint32_t value)
{
    return ((unsigned long long)((epoch << 2) | status) << 32) | value;
}

__global__ void __launch_bounds__(kLbThreads) scan_lookback_kernel(const uint32_t *in, uint32_t *out, int n,
                                                                   const int32_t *__restrict__ d_n, int extra,
                                                                   uint32_t *scratch)
{
    __shared__ uint32_t s_tile, s_epoch, s_prefix;
    volatile unsigned long long *desc = reinterpret_cast<volatile unsigned long long *>(scratch + 4);
    if (d_n) n = min(n, *d_n + extra);
    const int n_tiles = n > 0 ? (n + kLbTile - 1) / kLbTile : 0;
    // the grid is sized for the CAPACITY (the live length is device-resident): blocks beyond the live tiles
    // leave before taking a ticket -- with 24 M slots of pair capacity and 1.1 M live pairs that is 5600 of
    // 5860 blocks, each of which used to cost two atomics and a barrier
    const unsigned live_blocks = n_tiles > 0 ? (unsigned)n_tiles : 1u;
    if (blockIdx.x >= live_blocks) return;
    if (threadIdx.x == 0) {
        s_epoch = *(volatile uint32_t *)(scratch + 1);
        s_tile = atomicAdd(scratch, 1u);
    }
    __syncthreads();
    const uint32_t tile = s_tile, epoch = s_epoch;
    if ((int)tile < n_tiles) {
        const int base = (int)tile * kLbTile + threadIdx.x * kLbItems;
        uint32_t v[kLbItems];
        if (base + kLbItems <= n) {
            const uint4 *p = reinterpret_cast<const uint4 *>(in + base);
#pragma unroll
            for (int k = 0; k < kLbItems / 4; ++k) {
                const uint4 t = p[k];
                v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kLbItems; ++k) v[k] = base + k < n ? in[base + k] : 0u;
        }
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < kLbItems; ++k) s += v[k];
        uint32_t total;
        uint32_t ex = block_excl_scan(s, &total);
        if (tile == 0) {
            if (threadIdx.x == 0) { desc[0] = lb_pack(epoch, 2u, total); s_prefix = 0u; }
        } else {
            if (threadIdx.x == 0) desc[tile] = lb_pack(epoch, 1u, total);
            if (threadIdx.x < 32) {
                const int lane = threadIdx.x;
                int look = (int)tile - 1;
                uint32_t acc = 0;
                while (true) {
                    const int idx = look - lane;
                    unsigned long long d;
                    bool ok;
                    do {   // until the 32 descriptors in view are all of this epoch
                        d = idx >= 0 ? desc[idx] : lb_pack(epoch, 2u, 0u);
                        const uint32_t tag = (uint32_t)(d >> 32);
                        ok = (tag >> 2) == epoch && (tag & 3u) != 0u;
                    } while (!__all_sync(0xffffffffu, ok));
                    const uint32_t val = (uint32_t)d;
                    const unsigned pm = __ballot_sync(0xffffffffu, ((uint32_t)(d >> 32) & 3u) == 2u);
                    // lanes up to (and including) the nearest inclusive prefix contribute
                    const int stop = pm ? __ffs(pm) - 1 : 31;
                    acc += __reduce_add_sync(0xffffffffu, lane <= stop ? val : 0u);
                    if (pm) break;
                    look -= 32;
                }
                if (lane == 0) {
                    s_prefix = acc;
                    desc[tile] = lb_pack(epoch, 2u, acc + total);
                }
            }
        }
        __syncthreads();
        ex += s_prefix;
        if (base + kLbItems <= n) {
            uint4 *q = reinterpret_cast<uint4 *>(out + base);
#pragma unroll
            for (int k = 0; k < kLbItems / 4; ++k) {
                uint4 t;
                t.x = ex; ex += v[4 * k];
                t.y = ex; ex += v[4 * k + 1];
                t.z = ex; ex += v[4 * k + 2];
                t.w = ex; ex += v[4 * k + 3];
                q[k] = t;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kLbItems; ++k) {
                if (base + k < n) out[base + k] = ex;
                ex += v[k];
            }
        }
    }
    // the last block to finish re-arms the scratch for the next scan
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(scratch + 2, 1u) == live_blocks - 1u) {
            scratch[0] = 0u;
            scratch[2] = 0u;
            __threadfence();
            *(volatile uint32_t *)(scratch + 1) = (epoch + 1u) & 0x3fffffffu;
        }
    }
}

int scan_scratch_elems(int n) { return 2 * div_up(n, kLbTile) + 8; }

// out[i] = sum(in[0..i)) for i in [0, n).  in == out allowed.  To also get the grand total,
// scan n+1 elements with in[n] = 0.
int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int n, uint32_t *block_scratch, cudaStream_t s)
{
    return exclusive_scan_u32_dn(in, out, n, nullptr, 0, block_scratch, s);
}

// Same, but the length is min(cap_n, *d_n + extra) with d_n in device memory (no host sync):
// the grid is sized for cap_n, tiles past the live length exit immediately.
int exclusive_scan_u32_dn(const uint32_t *in, uint32_t *out, int cap_n, const int32_t *d_n, int extra,
                          uint32_t *block_scratch, cudaStream_t s)
{
    if (cap_n <= 0) return NANS_OK;
    static int three_pass = -1;
    if (three_pass < 0) three_pass = getenv("NANS_SCAN_3PASS") ? 1 : 0;
    const int tiles = div_up(cap_n, kScanTile);
    if (!three_pass) {
        scan_lookback_kernel<<<div_up(cap_n, kLbTile), kLbThreads, 0, s>>>(in, out, cap_n, d_n, extra, block_scratch);
        NANS_LAUNCH_CHECK();
        return NANS_OK;
    }
    scan_reduce_kernel<<<tiles, kScanThreads, 0, s>>>(in, cap_n, block_scratch, d_n, extra);
    NANS_LAUNCH_CHECK();
    scan_tiles_kernel<<<1, kScanThreads, 0, s>>>(block_scratch, tiles);
    NANS_LAUNCH_CHECK();
    scan_apply_kernel<<<tiles, kScanThreads, 0, s>>>(in, out, cap_n, block_scratch, d_n, extra);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

}  // namespace nans
