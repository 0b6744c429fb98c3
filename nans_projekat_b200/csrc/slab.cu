// slab.cu — one world split over several GPUs by contiguous body-index ranges ("slabs").
//
// The reference is a single process; this is the B200 extension of its step to one large world on
// an 8-GPU box (BASELINE.json config 5), kept EXACT: the multi-GPU result is
// bit-identical to the single-GPU (= sequential reference-order) result.
//
// Rank r owns body rows [lo_r, hi_r) of the global world.  Locally its world holds the owned rows
// first, then "ghost" copies of higher-rank bodies whose AABB reaches into the rank's bounding box.
// Local row order = global index order, so the locally emitted pair list is the global list
// restricted to pairs whose lower-index body is owned; every pair is tested and solved exactly once,
// by the owner of its lower-index body.  In the reference's sweep order every contact touching a
// body b that is processed on a lower rank precedes every contact touching b on b's own rank, so
// the solve is a pipeline: rank r receives the post-solve velocities of its boundary bodies from
// the lower ranks, solves, and hands the velocities of its ghosts on to their owners.
//
// This file holds the device side (bounds, order-preserving halo selection, pack/unpack); the
// exchange itself is NCCL point-to-point driven by the host (nans_projekat_b200/slab.py).
#include <float.h>

#include "world.cuh"

namespace nans {

constexpr int kBoundsThreads = 256;

__global__ void __launch_bounds__(kBoundsThreads) bounds_partial_kernel(DeviceWorld w, int n, float *partial)
{
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 a = w.aabb_lo[i], b = w.aabb_hi[i];
        lo[0] = fminf(lo[0], a.x); lo[1] = fminf(lo[1], a.y); lo[2] = fminf(lo[2], a.z);
        hi[0] = fmaxf(hi[0], b.x); hi[1] = fmaxf(hi[1], b.y); hi[2] = fmaxf(hi[2], b.z);
    }
    __shared__ float s[6][kBoundsThreads / 32];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float l = lo[k], h = hi[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((threadIdx.x & 31) == 0) { s[k][threadIdx.x >> 5] = l; s[3 + k][threadIdx.x >> 5] = h; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[threadIdx.x][0];
        for (int q = 1; q < kBoundsThreads / 32; ++q)
            v = threadIdx.x < 3 ? fminf(v, s[threadIdx.x][q]) : fmaxf(v, s[threadIdx.x][q]);
        partial[6 * blockIdx.x + threadIdx.x] = v;
    }
}

__global__ void bounds_final_kernel(const float *partial, int blocks, float *out6)
{
    const int k = threadIdx.x;
    if (k >= 6) return;
    float v = partial[k];
    for (int b = 1; b < blocks; ++b) v = k < 3 ? fminf(v, partial[6 * b + k]) : fmaxf(v, partial[6 * b + k]);
    out6[k] = v;
}

// flag owned rows whose AABB overlaps the box (inclusive, both are already inflated)
__global__ void __launch_bounds__(256) halo_flag_kernel(DeviceWorld w, int n_owned, float lx, float ly, float lz,
                                                        float hx, float hy, float hz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_owned) return;
    int f = 0;
    if (i < n_owned) {
        const float4 a = w.aabb_lo[i], b = w.aabb_hi[i];
        f = a.x <= hx && lx <= b.x && a.y <= hy && ly <= b.y && a.z <= hz && lz <= b.z;
    }
    w.pair_hit[i] = f;      // [n_owned] holds 0 so the exclusive scan yields the total there
}

// halo record: 10 x float4 = pos, vel, angvel, 6 x verts, (global id, 0, 0, 0)
constexpr int kHaloQuads = 10;

__global__ void __launch_bounds__(256) halo_pack_kernel(DeviceWorld w, int n_owned, int gid_base, float4 *out,
                                                        int32_t *sent_rows, int cap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_owned || !w.pair_hit[i]) return;
    const int k = (int)w.pair_hit_scan[i];     // order-preserving: ascending row = ascending global id
    if (k >= cap) return;
    float4 *o = out + (size_t)kHaloQuads * k;
    o[0] = w.pos[i]; o[1] = w.vel[i]; o[2] = w.angvel[i];
#pragma unroll
    for (int q = 0; q < 6; ++q) o[3 + q] = w.verts[6 * (size_t)i + q];
    o[9] = make_float4(__int_as_float(gid_base + i), 0.f, 0.f, 0.f);
    sent_rows[k] = i;
}

__global__ void __launch_bounds__(256) halo_unpack_kernel(DeviceWorld w, const float4 *in, int count, int row0,
                                                          int32_t *gid)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const float4 *r = in + (size_t)kHaloQuads * k;
    const int row = row0 + k;
    w.pos[row] = r[0]; w.vel[row] = r[1]; w.angvel[row] = r[2];
#pragma unroll
    for (int q = 0; q < 6; ++q) w.verts[6 * (size_t)row + q] = r[3 + q];
    w.force[row] = make_float4(0, 0, 0, 0);
    w.torque[row] = make_float4(0, 0, 0, 0);
    gid[row] = __float_as_int(r[9].x);
}

__global__ void __launch_bounds__(256) ghost_vel_pack_kernel(DeviceWorld w, int row0, int count, float4 *out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    out[2 * k] = w.vel[row0 + k];
    out[2 * k + 1] = w.angvel[row0 + k];
}

__global__ void __launch_bounds__(256) owned_vel_unpack_kernel(DeviceWorld w, const int32_t *rows, int count,
                                                               const float4 *in)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int row = rows[k];
    w.vel[row] = in[2 * k];
    w.angvel[row] = in[2 * k + 1];
}

// ---- host-side launchers (called from api.cu) -----------------------------------------------------
int slab_bounds(World *w, float *scratch, float *out6_dev)
{
    const int n = w->d.n_owned;
    if (n <= 0) return NANS_OK;
    const int blocks = min(div_up(n, kBoundsThreads), 1024);
    bounds_partial_kernel<<<blocks, kBoundsThreads, 0, w->stream>>>(w->d, n, scratch);
    NANS_LAUNCH_CHECK();
    bounds_final_kernel<<<1, 32, 0, w->stream>>>(scratch, blocks, out6_dev);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

int slab_pack_halo(World *w, const float box[6], int gid_base, float4 *out, int32_t *sent_rows, int cap,
                   int32_t *count_dev)
{
    DeviceWorld &d = w->d;
    const int n = d.n_owned;
    halo_flag_kernel<<<div_up(n + 1, 256), 256, 0, w->stream>>>(d, n, box[0], box[1], box[2], box[3], box[4], box[5]);
    NANS_LAUNCH_CHECK();
    int rc = exclusive_scan_u32((const uint32_t *)d.pair_hit, d.pair_hit_scan, n + 1, d.scan_block, w->stream);
    if (rc) return rc;
    if (n > 0) {
        halo_pack_kernel<<<div_up(n, 256), 256, 0, w->stream>>>(d, n, gid_base, out, sent_rows, cap);
        NANS_LAUNCH_CHECK();
    }
    NANS_CUDA(cudaMemcpyAsync(count_dev, d.pair_hit_scan + n, sizeof(int32_t), cudaMemcpyDeviceToDevice, w->stream));
    return NANS_OK;
}

int slab_unpack_halo(World *w, const float4 *in, int count, int row0, int32_t *gid)
{
    if (count <= 0) return NANS_OK;
    halo_unpack_kernel<<<div_up(count, 256), 256, 0, w->stream>>>(w->d, in, count, row0, gid);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

int slab_pack_ghost_vel(World *w, int row0, int count, float4 *out)
{
    if (count <= 0) return NANS_OK;
    ghost_vel_pack_kernel<<<div_up(count, 256), 256, 0, w->stream>>>(w->d, row0, count, out);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

int slab_unpack_owned_vel(World *w, const int32_t *rows, int count, const float4 *in)
{
    if (count <= 0) return NANS_OK;
    owned_vel_unpack_kernel<<<div_up(count, 256), 256, 0, w->stream>>>(w->d, rows, count, in);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

}  // namespace nans
