// scan.cu — exclusive prefix sum over uint32 (pair offsets, hit compaction, incidence offsets).
// Three phases: per-tile reduce -> single-block scan of tile sums -> per-tile scan + offset.
// HBM-bound: reads n twice, writes n once (12 B/element); tiles of 4096 keep loads 16 B wide.
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

int scan_scratch_elems(int n) { return div_up(n, kScanTile) + 1; }

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
    const int tiles = div_up(cap_n, kScanTile);
    scan_reduce_kernel<<<tiles, kScanThreads, 0, s>>>(in, cap_n, block_scratch, d_n, extra);
    NANS_LAUNCH_CHECK();
    scan_tiles_kernel<<<1, kScanThreads, 0, s>>>(block_scratch, tiles);
    NANS_LAUNCH_CHECK();
    scan_apply_kernel<<<tiles, kScanThreads, 0, s>>>(in, out, cap_n, block_scratch, d_n, extra);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

}  // namespace nans
