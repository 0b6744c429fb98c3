// api.cu — the extern "C" layer (include/nans_b200.h): device arena, state transfer, stage entries.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "world.cuh"

namespace nans {

thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

int launch_narrowphase_batch(int n, const int32_t *type, const float4 *posrad_a, const float4 *verts_a,
                             const float4 *posrad_b, const float4 *verts_b, int32_t *hit, int32_t *gjk,
                             float4 *out, int *work_counter, Counters *counters, cudaStream_t s);

static int fail(int code, const char *msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

// ---- arena: a bump allocator over one device block (the reference's memory_arena idea,
// code/utilities.cpp:25-49, which the reference never got to use: code/nans.cpp:1556-1582) -------
struct Bump {
    char *base;
    size_t used;
    template <typename T> T *take(size_t count)
    {
        used = (used + 255) & ~(size_t)255;
        T *p = base ? (T *)(base + used) : nullptr;
        used += sizeof(T) * count;
        return p;
    }
};

static uint32_t pow2_at_least(uint32_t v)
{
    uint32_t p = 1024;
    while (p < v) p <<= 1;
    return p;
}

struct Staging {            // per-field upload/download slots (device), carved from the same arena
    float *vec[7];          // pos, vel, force, ang, angvel, torque, scale : [nb][3]
    float *dvec[6];         // separate slots for the pipelined downloads (so H2D and D2H never share one)
    float *scal[3];         // mass, moi, radius : [nb]
    float *verts;           // [n_cubes][24]
    int32_t *wid;           // [nb]
};

struct Layout {
    DeviceWorld d;
    Staging st;
    float4 *snap;
    size_t bytes;
};

static Layout carve(const nans_world_desc &desc, char *base)
{
    Layout L;
    memset(&L, 0, sizeof(L));
    DeviceWorld &d = L.d;
    Bump b{base, 0};
    d.n_cubes = desc.n_cubes; d.n_spheres = desc.n_spheres; d.n_statics = desc.n_statics;
    d.nb = desc.n_cubes + desc.n_spheres;
    const size_t nb = (size_t)(d.nb > 0 ? d.nb : 1);
    const size_t nc = (size_t)(d.n_cubes > 0 ? d.n_cubes : 1);
    const size_t ns = (size_t)(d.n_statics > 0 ? d.n_statics : 1);
    d.max_pairs = desc.max_pairs > 0 ? desc.max_pairs : (int32_t)(24 * nb + 1024);
    d.max_contacts = desc.max_contacts > 0 ? desc.max_contacts : (int32_t)(8 * nb + 1024);
    if (d.max_pairs < d.max_contacts) d.max_pairs = d.max_contacts;   // contacts are a subset of the pairs; the solver borrows pair-sized scratch
    const size_t mp = (size_t)d.max_pairs, mc = (size_t)d.max_contacts;
    d.pos = b.take<float4>(nb); d.vel = b.take<float4>(nb); d.angvel = b.take<float4>(nb);
    d.ang = b.take<float4>(nb); d.force = b.take<float4>(nb); d.torque = b.take<float4>(nb);
    d.scale = b.take<float4>(nb);
    d.verts = b.take<float4>(6 * nc);
    d.world_id = b.take<int32_t>(nb);
    d.gid = b.take<int32_t>(nb);
    d.n_owned = d.nb;
    d.sent_mark = b.take<int32_t>(nb);
    d.ghost_owner_row = b.take<int32_t>(nb);
    d.st_pos = b.take<float4>(ns); d.st_ang = b.take<float4>(ns); d.st_scale = b.take<float4>(ns);
    d.st_verts = b.take<float4>(6 * ns); d.st_aabb = b.take<float4>(2 * ns);
    d.aabb_lo = b.take<float4>(nb); d.aabb_hi = b.take<float4>(nb);
    d.sbox = b.take<float4>(2 * nb);
    d.pair_tmp = b.take<uint32_t>(24 * nb);
    for (int k = 0; k < 2; ++k) { d.key[k] = b.take<uint32_t>(nb); d.val[k] = b.take<uint32_t>(nb); }
    const uint32_t table = pow2_at_least((uint32_t)(2 * nb));
    d.cell_mask = table - 1;
    d.cell_bits = 0;
    while ((1u << d.cell_bits) < table) ++d.cell_bits;
    d.n_seg = desc.n_spheres > 0 ? 5 : 2;
    d.cell_tab = b.take<uint4>(table);
    d.cell_count = b.take<uint32_t>((size_t)table + 1);
    d.pair_fill = b.take<uint32_t>(nb);
    d.pair_count = b.take<uint32_t>(5 * nb + 1);
    d.pair_a = b.take<int32_t>(mp); d.pair_b = b.take<int32_t>(mp);
    d.pair_hit = b.take<int32_t>(mp + 1); d.pair_hit_scan = b.take<uint32_t>(mp + 1);
    d.pair_out = b.take<float4>(3 * mp);
    d.c_a = b.take<int32_t>(mc); d.c_b = b.take<int32_t>(mc);
    d.c_pa = b.take<float4>(mc); d.c_pb = b.take<float4>(mc); d.c_n = b.take<float4>(mc);
    d.deg = b.take<uint32_t>(nb + 1); d.cursor = b.take<uint32_t>(nb);
    d.inc = b.take<int32_t>(2 * mc);
    d.succ_a = b.take<int32_t>(mc); d.succ_b = b.take<int32_t>(mc); d.run_flag = b.take<int32_t>(mc);
    d.run_start = b.take<int32_t>(mc); d.trace_level = b.take<int32_t>(mc);
    d.row_v = b.take<float4>(nb); d.row_w = b.take<float4>(nb);
    size_t scan_n = mp + 1;
    if (5 * nb + 1 > scan_n) scan_n = 5 * nb + 1;
    if ((size_t)table + 1 > scan_n) scan_n = (size_t)table + 1;
    d.scan_block = b.take<uint32_t>((size_t)scan_scratch_elems((int)scan_n) + 8);
    d.counters = b.take<Counters>(1);
    d.dt = b.take<float>(1);
    d.sticky = b.take<int32_t>(4);
    for (int k = 0; k < 7; ++k) L.st.vec[k] = b.take<float>(3 * nb);
    for (int k = 0; k < 6; ++k) L.st.dvec[k] = b.take<float>(3 * nb);
    for (int k = 0; k < 3; ++k) L.st.scal[k] = b.take<float>(nb);
    L.st.verts = b.take<float>(24 * nc);
    L.st.wid = b.take<int32_t>(nb);
    L.snap = b.take<float4>(6 * nb + 6 * nc);   // device-side snapshot of the dynamic state
    d.cell_size = 2.0f;
    L.bytes = (b.used + 255) & ~(size_t)255;
    return L;
}

struct IoPipe {               // pipelined host I/O: copies on their own streams, overlapping the step
    bool init;
    cudaStream_t up, down;
    cudaEvent_t ev_up, ev_unpack, ev_pack, ev_down[8];
    bool have_unpack, have_down;
    int next_ticket;
    int deferred;             // force/torque uploads in flight whose unpack waits for the step's detection phase
};

struct WorldImpl : World {
    IoPipe io;
    Staging st;
    int32_t cap_nb;          // capacity of the body arrays (slab mode varies nb below it)
    float4 *snap;
    bool has_world_id;
    int32_t *d_world_id_storage;
};

// ---- pack / unpack kernels (host [n][3] arrays <-> float4 SoA rows; w lanes preserved) ---------
__global__ void unpack_vec3_kernel(const float *__restrict__ src, float4 *__restrict__ dst, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = dst[i];
    v.x = src[3 * i]; v.y = src[3 * i + 1]; v.z = src[3 * i + 2];
    dst[i] = v;
}
__global__ void pack_vec3_kernel(const float4 *__restrict__ src, float *__restrict__ dst, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = src[i];
    dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z;
}
// which: 0 mass -> pos.w, vel.w = 1/m ; 1 moi -> ang.w, angvel.w = 1/moi ; 2 radius -> scale.w
__global__ void unpack_scalar_kernel(const float *__restrict__ src, DeviceWorld w, int which)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.nb) return;
    const float s = src[i];
    if (which == 0) { w.pos[i].w = s; w.vel[i].w = __fdiv_rn(1.0f, s); }
    else if (which == 1) { w.ang[i].w = s; w.angvel[i].w = __fdiv_rn(1.0f, s); }
    else w.scale[i].w = s;
}
__global__ void pack_scalar_kernel(float *__restrict__ dst, DeviceWorld w, int which)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.nb) return;
    dst[i] = which == 0 ? w.pos[i].w : which == 1 ? w.ang[i].w : w.scale[i].w;
}
__global__ void add_force_kernel(DeviceWorld w, int row, float3 f, float3 t)
{
    float4 a = w.force[row], b = w.torque[row];
    a.x = __fadd_rn(a.x, f.x); a.y = __fadd_rn(a.y, f.y); a.z = __fadd_rn(a.z, f.z);
    b.x = __fadd_rn(b.x, t.x); b.y = __fadd_rn(b.y, t.y); b.z = __fadd_rn(b.z, t.z);
    w.force[row] = a; w.torque[row] = b;
}
__global__ void set_body_kernel(DeviceWorld w, int row, float3 p, float3 v, float3 av, int mask)
{
    if (mask & 1) { float4 x = w.pos[row]; x.x = p.x; x.y = p.y; x.z = p.z; w.pos[row] = x; }
    if (mask & 2) { float4 x = w.vel[row]; x.x = v.x; x.y = v.y; x.z = v.z; w.vel[row] = x; }
    if (mask & 4) { float4 x = w.angvel[row]; x.x = av.x; x.y = av.y; x.z = av.z; w.angvel[row] = x; }
}

static inline WorldImpl *impl(nans_world *w) { return reinterpret_cast<WorldImpl *>(w); }

}  // namespace nans

using namespace nans;

extern "C" {

static void graph_invalidate(WorldImpl *w);
static int flush_deferred(WorldImpl *w);

const char *nans_last_error(void) { return g_err; }

int nans_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

uint64_t nans_kernel_launches(void) { return g_launches; }

int nans_debug_scan(const uint32_t *in, uint32_t *out, int32_t n)
{
    if (n < 0 || (n > 0 && (!in || !out))) return fail(NANS_ERR_ARG, "nans_debug_scan: bad arguments");
    if (n == 0) return NANS_OK;
    uint32_t *d = nullptr, *scratch = nullptr;
    const size_t sb = sizeof(uint32_t) * (size_t)scan_scratch_elems(n);
    NANS_CUDA(cudaMalloc(&d, sizeof(uint32_t) * (size_t)n));
    if (cudaMalloc(&scratch, sb) != cudaSuccess) { cudaFree(d); return fail(NANS_ERR_CUDA, "nans_debug_scan: out of device memory"); }
    int rc = NANS_OK;
    auto body = [&]() -> int {
        NANS_CUDA(cudaMemset(scratch, 0, sb));
        NANS_CUDA(cudaMemcpy(d, in, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
        for (int rep = 0; rep < 2; ++rep) {        // twice: the second run sees the re-armed scratch (epoch + 1) and, in place, the first one's sums
            if (rep == 1) NANS_CUDA(cudaMemcpy(d, in, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
            const int r = exclusive_scan_u32(d, d, n, scratch, 0);
            if (r) return r;
        }
        NANS_CUDA(cudaMemcpy(out, d, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
        return NANS_OK;
    };
    rc = body();
    cudaFree(d); cudaFree(scratch);
    return rc;
}

uint64_t nans_world_arena_bytes(const nans_world_desc *desc)
{
    if (!desc) return 0;
    return carve(*desc, nullptr).bytes;
}

int nans_world_create(const nans_world_desc *desc, nans_world **out)
{
    if (!desc || !out) return fail(NANS_ERR_ARG, "nans_world_create: null argument");
    if (desc->n_cubes < 0 || desc->n_spheres < 0 || desc->n_statics < 0 || desc->n_statics > kMaxStatics)
        return fail(NANS_ERR_ARG, "nans_world_create: bad body counts (statics <= 16)");
    int ndev = 0;
    NANS_CUDA(cudaGetDeviceCount(&ndev));
    if (desc->device < 0 || desc->device >= ndev) return fail(NANS_ERR_CUDA, "nans_world_create: no such CUDA device");
    NANS_CUDA(cudaSetDevice(desc->device));
    WorldImpl *w = new WorldImpl();
    memset(static_cast<World *>(w), 0, sizeof(World));
    memset(&w->io, 0, sizeof(w->io));
    w->desc = *desc;
    w->device = desc->device;
    const size_t need = carve(*desc, nullptr).bytes;
    if (desc->arena) {
        if (desc->arena_bytes < need || ((uintptr_t)desc->arena & 255)) {
            delete w;
            return fail(NANS_ERR_ARG, "nans_world_create: caller arena too small or not 256 B aligned");
        }
        w->arena = desc->arena;
        w->owns_arena = false;
    } else {
        cudaError_t e = cudaMalloc(&w->arena, need);   // the ONE allocation of the world's lifetime
        if (e != cudaSuccess) {
            delete w;
            snprintf(g_err, sizeof(g_err), "cudaMalloc(%zu): %s", need, cudaGetErrorString(e));
            return NANS_ERR_CUDA;
        }
        w->owns_arena = true;
    }
    w->arena_bytes = need;
    Layout L = carve(*desc, (char *)w->arena);
    w->d = L.d;
    w->st = L.st;
    w->cap_nb = L.d.nb;
    w->snap = L.snap;
    w->d_world_id_storage = L.d.world_id;
    w->d.world_id = nullptr;
    w->has_world_id = false;
    auto finish = [&]() -> int {
        if (desc->stream) { w->stream = (cudaStream_t)desc->stream; w->owns_stream = false; }
        else { NANS_CUDA(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking)); w->owns_stream = true; }
        NANS_CUDA(cudaMemsetAsync(w->arena, 0, need, w->stream));
        NANS_CUDA(cudaMallocHost((void **)&w->h_counters, sizeof(Counters) + 64));
        NANS_CUDA(cudaStreamSynchronize(w->stream));
        return NANS_OK;
    };
    const int frc = finish();
    if (frc) {                      // nothing leaks on a failed create (the arena, the stream, the host mirror)
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        nans_world_destroy(reinterpret_cast<nans_world *>(w));
        memcpy(g_err, keep, sizeof(keep));
        return frc;
    }
    *out = reinterpret_cast<nans_world *>(w);
    return NANS_OK;
}

void nans_world_destroy(nans_world *h)
{
    if (!h) return;
    WorldImpl *w = impl(h);
    cudaSetDevice(w->device);
    if (w->stream) cudaStreamSynchronize(w->stream);
    slab_destroy(w);
    if (w->graph_exec) cudaGraphExecDestroy(w->graph_exec);
    if (w->graph_exec_b) cudaGraphExecDestroy(w->graph_exec_b);
    if (w->io.init) {
        cudaStreamSynchronize(w->io.up); cudaStreamSynchronize(w->io.down);
        cudaStreamDestroy(w->io.up); cudaStreamDestroy(w->io.down);
        cudaEventDestroy(w->io.ev_up); cudaEventDestroy(w->io.ev_unpack); cudaEventDestroy(w->io.ev_pack);
        for (auto &e : w->io.ev_down) cudaEventDestroy(e);
    }
    if (w->owns_arena) cudaFree(w->arena);
    if (w->owns_stream && w->stream) cudaStreamDestroy(w->stream);
    if (w->h_counters) cudaFreeHost(w->h_counters);
    delete w;
}

int nans_synchronize(nans_world *h)
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    { const int frc = flush_deferred(impl(h)); if (frc) return frc; }
    NANS_CUDA(cudaStreamSynchronize(impl(h)->stream));
    return NANS_OK;
}

int nans_rebuild_vertices(nans_world *h);

int nans_world_upload(nans_world *h, const nans_scene_view *sc)
{
    if (!h || !sc) return fail(NANS_ERR_ARG, "nans_world_upload: null argument");
    WorldImpl *w = impl(h);
    DeviceWorld &d = w->d;
    NANS_CUDA(cudaSetDevice(w->device));
    { const int frc = flush_deferred(w); if (frc) return frc; }
    cudaStream_t s = w->stream;
    const int nb = d.nb;
    const int grid = div_up(nb > 0 ? nb : 1, 256);
    const float *vsrc[7] = {sc->pos, sc->vel, sc->force, sc->ang, sc->angvel, sc->torque, sc->scale};
    float4 *vdst[7] = {d.pos, d.vel, d.force, d.ang, d.angvel, d.torque, d.scale};
    if (w->io.init) { NANS_CUDA(cudaStreamSynchronize(w->io.up)); NANS_CUDA(cudaStreamSynchronize(w->io.down)); }
    if (nb > 0) {
        for (int k = 0; k < 7; ++k) {
            if (!vsrc[k]) continue;
            NANS_CUDA(cudaMemcpyAsync(w->st.vec[k], vsrc[k], sizeof(float) * 3 * (size_t)nb, cudaMemcpyHostToDevice, s));
            unpack_vec3_kernel<<<grid, 256, 0, s>>>(w->st.vec[k], vdst[k], nb);
            NANS_LAUNCH_CHECK();
        }
        const float *ssrc[3] = {sc->mass, sc->moi, sc->radius};
        for (int k = 0; k < 3; ++k) {
            if (!ssrc[k]) continue;
            NANS_CUDA(cudaMemcpyAsync(w->st.scal[k], ssrc[k], sizeof(float) * (size_t)nb, cudaMemcpyHostToDevice, s));
            unpack_scalar_kernel<<<grid, 256, 0, s>>>(w->st.scal[k], d, k);
            NANS_LAUNCH_CHECK();
        }
        if (sc->verts && d.n_cubes > 0)   // 24 floats per cube == 6 float4: same bytes, direct copy
            NANS_CUDA(cudaMemcpyAsync(d.verts, sc->verts, sizeof(float) * 24 * (size_t)d.n_cubes,
                                      cudaMemcpyHostToDevice, s));
        if (sc->world_id) {
            NANS_CUDA(cudaMemcpyAsync(w->d_world_id_storage, sc->world_id, sizeof(int32_t) * (size_t)nb,
                                      cudaMemcpyHostToDevice, s));
            d.world_id = w->d_world_id_storage;
            w->has_world_id = true;
        }
    }
    // statics are few: converted on the host
    const int ns = d.n_statics;
    if (ns > 0) {
        std::vector<float4> tmp(ns);
        auto up3 = [&](const float *src, float4 *dst, const float *wsrc) -> int {
            // read back current rows so the w lanes survive a partial update
            NANS_CUDA(cudaMemcpyAsync(tmp.data(), dst, sizeof(float4) * ns, cudaMemcpyDeviceToHost, s));
            NANS_CUDA(cudaStreamSynchronize(s));
            for (int k = 0; k < ns; ++k) {
                if (src) { tmp[k].x = src[3 * k]; tmp[k].y = src[3 * k + 1]; tmp[k].z = src[3 * k + 2]; }
                if (wsrc) tmp[k].w = 1.0f / wsrc[k];
            }
            NANS_CUDA(cudaMemcpyAsync(dst, tmp.data(), sizeof(float4) * ns, cudaMemcpyHostToDevice, s));
            NANS_CUDA(cudaStreamSynchronize(s));
            return NANS_OK;
        };
        int rc;
        if (sc->st_pos || sc->st_mass) { rc = up3(sc->st_pos, d.st_pos, sc->st_mass); if (rc) return rc; }
        if (sc->st_ang || sc->st_moi) { rc = up3(sc->st_ang, d.st_ang, sc->st_moi); if (rc) return rc; }
        if (sc->st_scale) { rc = up3(sc->st_scale, d.st_scale, nullptr); if (rc) return rc; }
        if (sc->st_verts)
            NANS_CUDA(cudaMemcpyAsync(d.st_verts, sc->st_verts, sizeof(float) * 24 * (size_t)ns, cudaMemcpyHostToDevice, s));
    }
    // broadphase cell >= the largest inflated AABB extent any body can have (box diagonal / diameter)
    if ((sc->scale || sc->radius) && nb > 0) {
        float ext = 0.0f;
        for (int i = 0; i < nb; ++i) {
            if (i < d.n_cubes && sc->scale) {
                const float *q = sc->scale + 3 * i;
                ext = fmaxf(ext, sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]));
            } else if (i >= d.n_cubes && sc->radius) {
                ext = fmaxf(ext, 2.0f * sc->radius[i]);
            }
        }
        const float cell = ext * 1.02f + 0.05f;
        if (cell > d.cell_size || d.cell_size == 2.0f) d.cell_size = fmaxf(cell, 0.25f);
    }
    w->have_contacts = false;
    if (sc->scale || sc->radius || sc->world_id) graph_invalidate(w);   // launch parameters changed
    NANS_CUDA(cudaStreamSynchronize(s));
    return NANS_OK;
}

int nans_world_download(nans_world *h, nans_scene_view *sc)
{
    if (!h || !sc) return fail(NANS_ERR_ARG, "nans_world_download: null argument");
    WorldImpl *w = impl(h);
    DeviceWorld &d = w->d;
    NANS_CUDA(cudaSetDevice(w->device));
    { const int frc = flush_deferred(w); if (frc) return frc; }
    cudaStream_t s = w->stream;
    const int nb = d.nb;
    const int grid = div_up(nb > 0 ? nb : 1, 256);
    float *vdst[7] = {sc->pos, sc->vel, sc->force, sc->ang, sc->angvel, sc->torque, sc->scale};
    const float4 *vsrc[7] = {d.pos, d.vel, d.force, d.ang, d.angvel, d.torque, d.scale};
    if (w->io.init) { NANS_CUDA(cudaStreamSynchronize(w->io.up)); NANS_CUDA(cudaStreamSynchronize(w->io.down)); }
    // (writing the packed rows straight into the caller's pinned, device-mapped buffers from one fused kernel -- no
    // staging slot, no memcpy -- was measured in round 2: 12-byte-stride stores over PCIe, end-to-end throughput
    // fell from 5.4e8 to 1.8e8 body-steps/s; the staged copy below stays)
    if (nb > 0) {
        for (int k = 0; k < 7; ++k) {
            if (!vdst[k]) continue;
            pack_vec3_kernel<<<grid, 256, 0, s>>>(vsrc[k], w->st.vec[k], nb);
            NANS_LAUNCH_CHECK();
            NANS_CUDA(cudaMemcpyAsync(vdst[k], w->st.vec[k], sizeof(float) * 3 * (size_t)nb, cudaMemcpyDeviceToHost, s));
        }
        float *sdst[3] = {sc->mass, sc->moi, sc->radius};
        for (int k = 0; k < 3; ++k) {
            if (!sdst[k]) continue;
            pack_scalar_kernel<<<grid, 256, 0, s>>>(w->st.scal[k], d, k);
            NANS_LAUNCH_CHECK();
            NANS_CUDA(cudaMemcpyAsync(sdst[k], w->st.scal[k], sizeof(float) * (size_t)nb, cudaMemcpyDeviceToHost, s));
        }
        if (sc->verts && d.n_cubes > 0)
            NANS_CUDA(cudaMemcpyAsync(sc->verts, d.verts, sizeof(float) * 24 * (size_t)d.n_cubes, cudaMemcpyDeviceToHost, s));
    }
    if (sc->st_verts && d.n_statics > 0)
        NANS_CUDA(cudaMemcpyAsync(sc->st_verts, d.st_verts, sizeof(float) * 24 * (size_t)d.n_statics,
                                  cudaMemcpyDeviceToHost, s));
    // sticky errors ride along: a stalled solve (watchdog) must not look like a finished step
    int32_t *sticky = w->h_counters + sizeof(Counters) / sizeof(int32_t);
    NANS_CUDA(cudaMemcpyAsync(sticky, d.sticky, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    NANS_CUDA(cudaStreamSynchronize(s));
    if (*sticky & 1) return fail(NANS_ERR_STATE, "solver: dependency schedule stalled in an earlier step (watchdog); velocities are partially solved");
    return NANS_OK;
}

// ---- pipelined I/O ------------------------------------------------------------------------------
// The reference's frame is synchronous (host writes forces, steps, reads poses).  For hosts that can
// take the poses one frame late (a renderer), the copies can ride on separate streams and overlap the
// step: H2D of frame k+1's inputs and D2H of frame k's poses both run while frame k+1 computes.
static int io_init(WorldImpl *w)
{
    if (w->io.init) return NANS_OK;
    NANS_CUDA(cudaStreamCreateWithFlags(&w->io.up, cudaStreamNonBlocking));
    NANS_CUDA(cudaStreamCreateWithFlags(&w->io.down, cudaStreamNonBlocking));
    NANS_CUDA(cudaEventCreateWithFlags(&w->io.ev_up, cudaEventDisableTiming));
    NANS_CUDA(cudaEventCreateWithFlags(&w->io.ev_unpack, cudaEventDisableTiming));
    NANS_CUDA(cudaEventCreateWithFlags(&w->io.ev_pack, cudaEventDisableTiming));
    for (auto &e : w->io.ev_down) NANS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    w->io.have_unpack = w->io.have_down = false;
    w->io.next_ticket = 0;
    w->io.deferred = 0;
    w->io.init = true;
    return NANS_OK;
}

// Forces and torques are read by integrate-forces only, and collision detection reads neither them nor
// the velocities, so an upload that carries nothing else does not have to finish before the step starts:
// its unpack is DEFERRED until nans_step has queued the detection phase (flush_deferred), and the
// host-to-device copy overlaps broadphase + narrowphase.  Any other entry point flushes first.
static int flush_deferred(WorldImpl *w)
{
    if (!w->io.init || !w->io.deferred) return NANS_OK;
    DeviceWorld &d = w->d;
    float4 *vdst[6] = {d.pos, d.vel, d.force, d.ang, d.angvel, d.torque};
    NANS_CUDA(cudaStreamWaitEvent(w->stream, w->io.ev_up, 0));
    const int grid = div_up(d.nb > 0 ? d.nb : 1, 256);
    for (int k = 0; k < 6; ++k)
        if ((w->io.deferred >> k) & 1) {
            unpack_vec3_kernel<<<grid, 256, 0, w->stream>>>(w->st.vec[k], vdst[k], d.nb);
            NANS_LAUNCH_CHECK();
        }
    NANS_CUDA(cudaEventRecord(w->io.ev_unpack, w->stream));
    w->io.have_unpack = true;
    w->io.deferred = 0;
    return NANS_OK;
}

int nans_world_upload_async(nans_world *h, const nans_scene_view *sc)
{
    if (!h || !sc) return fail(NANS_ERR_ARG, "null argument");
    WorldImpl *w = impl(h);
    DeviceWorld &d = w->d;
    NANS_CUDA(cudaSetDevice(w->device));
    int rc = io_init(w);
    if (rc) return rc;
    const int nb = d.nb;
    if (nb == 0) return NANS_OK;
    const float *vsrc[6] = {sc->pos, sc->vel, sc->force, sc->ang, sc->angvel, sc->torque};
    float4 *vdst[6] = {d.pos, d.vel, d.force, d.ang, d.angvel, d.torque};
    int mask = 0;
    for (int k = 0; k < 6; ++k) mask |= vsrc[k] ? 1 << k : 0;
    const int kForceTorque = (1 << 2) | (1 << 5);
    const bool defer = mask && !(mask & ~kForceTorque);
    // a pending deferred upload is unpacked first (keeps the order of writes to the same field)
    if (w->io.deferred && (!defer || (w->io.deferred & mask))) { rc = flush_deferred(w); if (rc) return rc; }
    // the upload slots are free once the previous unpack has run (downloads use their own slots)
    if (w->io.have_unpack) NANS_CUDA(cudaStreamWaitEvent(w->io.up, w->io.ev_unpack, 0));
    for (int k = 0; k < 6; ++k)
        if (vsrc[k])
            NANS_CUDA(cudaMemcpyAsync(w->st.vec[k], vsrc[k], sizeof(float) * 3 * (size_t)nb, cudaMemcpyHostToDevice, w->io.up));
    NANS_CUDA(cudaEventRecord(w->io.ev_up, w->io.up));
    if (defer) { w->io.deferred |= mask; return NANS_OK; }
    NANS_CUDA(cudaStreamWaitEvent(w->stream, w->io.ev_up, 0));
    const int grid = div_up(nb, 256);
    for (int k = 0; k < 6; ++k)
        if (vsrc[k]) {
            unpack_vec3_kernel<<<grid, 256, 0, w->stream>>>(w->st.vec[k], vdst[k], nb);
            NANS_LAUNCH_CHECK();
        }
    NANS_CUDA(cudaEventRecord(w->io.ev_unpack, w->stream));
    w->io.have_unpack = true;
    return NANS_OK;
}

int nans_world_download_async(nans_world *h, nans_scene_view *sc, int32_t *ticket)
{
    if (!h || !sc || !ticket) return fail(NANS_ERR_ARG, "null argument");
    WorldImpl *w = impl(h);
    DeviceWorld &d = w->d;
    NANS_CUDA(cudaSetDevice(w->device));
    int rc = io_init(w);
    if (rc) return rc;
    { const int frc = flush_deferred(w); if (frc) return frc; }
    const int nb = d.nb;
    float *vdst[6] = {sc->pos, sc->vel, sc->force, sc->ang, sc->angvel, sc->torque};
    const float4 *vsrc[6] = {d.pos, d.vel, d.force, d.ang, d.angvel, d.torque};
    const int t = w->io.next_ticket;
    // the previous D2H must have drained the staging slots before they are packed again
    if (w->io.have_down) NANS_CUDA(cudaStreamWaitEvent(w->stream, w->io.ev_down[(t + 7) % 8], 0));
    const int grid = div_up(nb > 0 ? nb : 1, 256);
    for (int k = 0; k < 6 && nb > 0; ++k)
        if (vdst[k]) {
            pack_vec3_kernel<<<grid, 256, 0, w->stream>>>(vsrc[k], w->st.dvec[k], nb);
            NANS_LAUNCH_CHECK();
        }
    NANS_CUDA(cudaEventRecord(w->io.ev_pack, w->stream));
    NANS_CUDA(cudaStreamWaitEvent(w->io.down, w->io.ev_pack, 0));
    for (int k = 0; k < 6 && nb > 0; ++k)
        if (vdst[k])
            NANS_CUDA(cudaMemcpyAsync(vdst[k], w->st.dvec[k], sizeof(float) * 3 * (size_t)nb, cudaMemcpyDeviceToHost, w->io.down));
    NANS_CUDA(cudaEventRecord(w->io.ev_down[t % 8], w->io.down));
    w->io.have_down = true;
    w->io.next_ticket = t + 1;
    *ticket = t;
    return NANS_OK;
}

int nans_world_wait(nans_world *h, int32_t ticket)
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    WorldImpl *w = impl(h);
    NANS_CUDA(cudaSetDevice(w->device));
    if (ticket < 0) { const int frc = flush_deferred(w); if (frc) return frc; }
    if (!w->io.init || ticket < 0) {
        if (w->io.init) { NANS_CUDA(cudaStreamSynchronize(w->io.up)); NANS_CUDA(cudaStreamSynchronize(w->io.down)); }
        NANS_CUDA(cudaStreamSynchronize(w->stream));
        return NANS_OK;
    }
    if (ticket >= w->io.next_ticket || ticket < w->io.next_ticket - 8) return fail(NANS_ERR_ARG, "nans_world_wait: stale or unknown ticket");
    NANS_CUDA(cudaEventSynchronize(w->io.ev_down[ticket % 8]));
    return NANS_OK;
}

int nans_world_add_force(nans_world *h, int32_t row, const float f[3], const float t[3])
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    WorldImpl *w = impl(h);
    if (row < 0 || row >= w->d.nb) return fail(NANS_ERR_ARG, "nans_world_add_force: body row out of range");
    NANS_CUDA(cudaSetDevice(w->device));
    { const int frc = flush_deferred(w); if (frc) return frc; }
    const float3 ff = f ? make_float3(f[0], f[1], f[2]) : make_float3(0, 0, 0);
    const float3 tt = t ? make_float3(t[0], t[1], t[2]) : make_float3(0, 0, 0);
    add_force_kernel<<<1, 1, 0, w->stream>>>(w->d, row, ff, tt);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

int nans_world_set_body(nans_world *h, int32_t row, const float p[3], const float v[3], const float av[3])
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    WorldImpl *w = impl(h);
    if (row < 0 || row >= w->d.nb) return fail(NANS_ERR_ARG, "nans_world_set_body: body row out of range");
    NANS_CUDA(cudaSetDevice(w->device));
    const int mask = (p ? 1 : 0) | (v ? 2 : 0) | (av ? 4 : 0);
    const float3 z = make_float3(0, 0, 0);
    set_body_kernel<<<1, 1, 0, w->stream>>>(w->d, row, p ? make_float3(p[0], p[1], p[2]) : z,
                                            v ? make_float3(v[0], v[1], v[2]) : z,
                                            av ? make_float3(av[0], av[1], av[2]) : z, mask);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

// Device-to-device snapshot / restore of the dynamic state (pose, velocities, forces, vertices):
// the hot-reload host keeps its whole world in one block so it can be checkpointed by copying it
// (code/sdl_nans.cpp:541-555); this is the device-side equivalent (episode reset, bench windows).
static int snapshot_copy(WorldImpl *w, bool restore)
{
    DeviceWorld &d = w->d;
    NANS_CUDA(cudaSetDevice(w->device));
    { const int frc = flush_deferred(w); if (frc) return frc; }
    float4 *rows[6] = {d.pos, d.vel, d.angvel, d.ang, d.force, d.torque};
    // fixed layout (capacity strides): slab mode varies nb / n_cubes between snapshot and restore
    const size_t nb = (size_t)w->cap_nb, nc = (size_t)(d.n_spheres ? d.n_cubes : w->cap_nb);
    for (int k = 0; k < 6 && nb; ++k) {
        float4 *snap = w->snap + k * nb;
        NANS_CUDA(cudaMemcpyAsync(restore ? rows[k] : snap, restore ? snap : rows[k], sizeof(float4) * nb,
                                  cudaMemcpyDeviceToDevice, w->stream));
    }
    if (nc) {
        float4 *snap = w->snap + 6 * nb;
        NANS_CUDA(cudaMemcpyAsync(restore ? d.verts : snap, restore ? snap : d.verts, sizeof(float4) * 6 * nc,
                                  cudaMemcpyDeviceToDevice, w->stream));
    }
    return NANS_OK;
}
int nans_world_snapshot(nans_world *h) { return h ? snapshot_copy(impl(h), false) : fail(NANS_ERR_ARG, "null world"); }
int nans_world_restore(nans_world *h) { return h ? snapshot_copy(impl(h), true) : fail(NANS_ERR_ARG, "null world"); }

// ---- stages ------------------------------------------------------------------------------------
int nans_integrate_forces(nans_world *h, float dt)
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    NANS_CUDA(cudaSetDevice(impl(h)->device));
    { const int frc = flush_deferred(impl(h)); if (frc) return frc; }
    { const int rc = launch_set_dt(impl(h), dt); if (rc) return rc; }
    return launch_integrate_forces(impl(h));
}

int nans_detect_collisions(nans_world *h)
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    WorldImpl *w = impl(h);
    NANS_CUDA(cudaSetDevice(w->device));
    int rc = launch_broadphase(w);
    if (rc) return rc;
    rc = launch_narrowphase(w);
    if (rc) return rc;
    rc = launch_contacts(w);
    if (rc) return rc;
    w->have_contacts = true;
    return NANS_OK;
}

int nans_solve_constraints(nans_world *h, float dt)
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    WorldImpl *w = impl(h);
    if (!w->have_contacts) return fail(NANS_ERR_STATE, "nans_solve_constraints: no contact list (call detect or set_contacts)");
    NANS_CUDA(cudaSetDevice(w->device));
    { const int rc = launch_set_dt(w, dt); if (rc) return rc; }
    return launch_solver(w);
}

int nans_world_set_solver(nans_world *h, int32_t mode)
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    if (mode != NANS_SOLVER_EXACT && mode != NANS_SOLVER_SHUFFLED) return fail(NANS_ERR_ARG, "nans_world_set_solver: unknown mode");
    WorldImpl *w = impl(h);
    if (mode == NANS_SOLVER_SHUFFLED && w->slab) return fail(NANS_ERR_STATE, "a slab-partitioned world runs the exact order only");
    if (w->solver_mode != mode) graph_invalidate(w);
    w->solver_mode = mode;
    return NANS_OK;
}

int nans_integrate_velocities(nans_world *h, float dt)
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    NANS_CUDA(cudaSetDevice(impl(h)->device));
    { const int rc = launch_set_dt(impl(h), dt); if (rc) return rc; }
    return launch_integrate_velocities(impl(h));
}

int nans_rebuild_vertices(nans_world *h)
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    WorldImpl *w = impl(h);
    NANS_CUDA(cudaSetDevice(w->device));
    { const int rc = launch_set_dt(w, 0.0f); if (rc) return rc; }
    return launch_integrate_velocities(w);   // Position += 0*V is exact for finite V; rebuilds every Model
}

// Draw data: Model matrices of every body, built on the device (code/nans.cpp:1870-1881, 1913-1941, 1971-1990).
// d_out (device, e.g. a CUDA-GL interop buffer) and/or h_out (host): [(nb + n_statics)][16] floats, column-major.
int nans_world_models(nans_world *h, void *d_out, float *h_out)
{
    if (!h || (!d_out && !h_out)) return fail(NANS_ERR_ARG, "nans_world_models: no output buffer");
    WorldImpl *w = impl(h);
    NANS_CUDA(cudaSetDevice(w->device));
    { const int frc = flush_deferred(w); if (frc) return frc; }
    const size_t n = (size_t)w->d.nb + (size_t)w->d.n_statics;
    // without a device buffer of the caller's the matrices are staged in the narrowphase output block (idle here)
    if (!d_out && 4 * n > 3 * (size_t)w->d.max_pairs) return fail(NANS_ERR_CAPACITY, "nans_world_models: pass a device buffer for a world this size");
    float4 *dst = d_out ? (float4 *)d_out : w->d.pair_out;
    const int rc = launch_models(w, dst);
    if (rc) return rc;
    if (h_out) {
        NANS_CUDA(cudaMemcpyAsync(h_out, dst, sizeof(float) * 16 * n, cudaMemcpyDeviceToHost, w->stream));
        NANS_CUDA(cudaStreamSynchronize(w->stream));
    }
    return NANS_OK;
}

// One step = IntegrateForces, DetectCollisions, SolveConstraints, IntegrateVelocities (code/nans.cpp:1758-1762).
// Detection reads positions and vertices only and integrate-forces writes velocities and clears forces only, so
// the two commute bit for bit; the step runs detection FIRST, so that a force/torque upload still in flight
// (nans_world_upload_async) overlaps broadphase + narrowphase and is unpacked just before integrate-forces.
static int step_phase_a(nans_world *h) { return nans_detect_collisions(h); }
static int step_phase_b(nans_world *h)       // dt is already in device memory (launch_set_dt)
{
    // (forking integrate-forces onto a side stream next to the solver's schedule kernels was measured:
    // 1.813 vs 1.808 ms per step, the fork/join costs what the overlap saves)
    // (seeding the solver's velocity rows from integrate-forces and consuming them in integrate-velocities --
    // two launches and 130 MB of traffic fewer -- was measured too: 1.7936 vs 1.7944 ms, not kept)
    WorldImpl *w = impl(h);
    int rc = launch_integrate_forces(w);
    if (rc) return rc;
    if (!w->have_contacts) return fail(NANS_ERR_STATE, "no contact list");
    rc = launch_solver(w);
    if (rc) return rc;
    return launch_integrate_velocities(w);
}
static int step_eager(nans_world *h)
{
    int rc = step_phase_a(h);
    if (rc) return rc;
    rc = flush_deferred(impl(h));
    if (rc) return rc;
    return step_phase_b(h);
}

static void graph_invalidate(WorldImpl *w)
{
    if (w->graph_exec) { cudaGraphExecDestroy(w->graph_exec); w->graph_exec = nullptr; }
    if (w->graph_exec_b) { cudaGraphExecDestroy(w->graph_exec_b); w->graph_exec_b = nullptr; }
    if (w->graph_state > 0) w->graph_state = 0;
}

// The step is ~35 launches whose parameters do not change from frame to frame -- every size that varies, and dt
// itself, lives in device memory -- so it is captured once into two CUDA graphs and replayed.  The first step runs
// eagerly (it also warms the static launch-configuration caches), the second is captured; a dt that changes every
// frame (the reference host passes the measured frame time, code/sdl_nans.cpp:999) does not invalidate the graphs.
// NANS_GRAPH=0 disables.
int nans_step(nans_world *h, float dt)
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    WorldImpl *w = impl(h);
    if (w->slab) return fail(NANS_ERR_STATE, "this world is one rank's slab of a larger world: use nans_slab_step");
    static int enabled = -1;
    if (enabled < 0) { const char *e = getenv("NANS_GRAPH"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
    NANS_CUDA(cudaSetDevice(w->device));
    { const int rc = launch_set_dt(w, dt); if (rc) return rc; }
    if (!enabled || w->graph_state < 0 || w->d.nb == 0) return step_eager(h);
    auto replay = [&]() -> int {
        NANS_CUDA(cudaGraphLaunch(w->graph_exec, w->stream));
        const int rc = flush_deferred(w);            // waits for the upload stream, unpacks forces/torques
        if (rc) return rc;
        NANS_CUDA(cudaGraphLaunch(w->graph_exec_b, w->stream));
        g_launches += w->graph_launches;
        w->have_contacts = true;
        return NANS_OK;
    };
    if (w->graph_state == 2) return replay();
    if (w->graph_state == 1) {
        // two graphs: detection, and everything that follows the (uncaptured) deferred-input unpack
        const unsigned long long before = g_launches;
        auto capture = [&](cudaGraphExec_t *exec, bool phase_b) -> bool {
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return false;
            if (phase_b) w->have_contacts = true;
            const int rc = phase_b ? step_phase_b(h) : step_phase_a(h);
            const cudaError_t ee = cudaStreamEndCapture(w->stream, &graph);
            const bool ok = !rc && ee == cudaSuccess && graph && cudaGraphInstantiate(exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (!ok) *exec = nullptr;
            return ok;
        };
        const bool ok = capture(&w->graph_exec, false) && capture(&w->graph_exec_b, true);
        w->graph_launches = (unsigned)(g_launches - before);
        g_launches = before;
        if (!ok) {
            cudaGetLastError();
            graph_invalidate(w);
            w->graph_state = -1;            // capture not possible here: stay eager
            return step_eager(h);
        }
        w->graph_state = 2;
        return replay();
    }
    w->graph_state = 1;
    return step_eager(h);
}

// One step with a CUDA event between every stage (on the world's stream); stage_ms[8]:
// 0 integrate_forces, 1 broadphase, 2 narrowphase, 3 contact compaction, 4 solver (schedule + levels),
// 5 integrate_velocities + vertex rebuild, 6 whole step, 7 unused.  Synchronises.
int nans_step_profiled(nans_world *h, float dt, float *stage_ms)
{
    if (!h || !stage_ms) return fail(NANS_ERR_ARG, "null argument");
    WorldImpl *w = impl(h);
    NANS_CUDA(cudaSetDevice(w->device));
    static cudaEvent_t ev[7];
    static bool have = false;
    if (!have) { for (auto &e : ev) NANS_CUDA(cudaEventCreate(&e)); have = true; }
    cudaStream_t s = w->stream;
    int rc;
    if ((rc = flush_deferred(w))) return rc;
    if ((rc = launch_set_dt(w, dt))) return rc;
    NANS_CUDA(cudaEventRecord(ev[0], s));
    if ((rc = launch_integrate_forces(w))) return rc;
    NANS_CUDA(cudaEventRecord(ev[1], s));
    if ((rc = launch_broadphase(w))) return rc;
    NANS_CUDA(cudaEventRecord(ev[2], s));
    if ((rc = launch_narrowphase(w))) return rc;
    NANS_CUDA(cudaEventRecord(ev[3], s));
    if ((rc = launch_contacts(w))) return rc;
    w->have_contacts = true;
    NANS_CUDA(cudaEventRecord(ev[4], s));
    if ((rc = launch_solver(w))) return rc;
    NANS_CUDA(cudaEventRecord(ev[5], s));
    if ((rc = launch_integrate_velocities(w))) return rc;
    NANS_CUDA(cudaEventRecord(ev[6], s));
    NANS_CUDA(cudaEventSynchronize(ev[6]));
    for (int k = 0; k < 6; ++k) NANS_CUDA(cudaEventElapsedTime(&stage_ms[k], ev[k], ev[k + 1]));
    NANS_CUDA(cudaEventElapsedTime(&stage_ms[6], ev[0], ev[6]));
    stage_ms[7] = 0.f;
    return NANS_OK;
}

int nans_get_stats(nans_world *h, nans_step_stats *out)
{
    if (!h || !out) return fail(NANS_ERR_ARG, "null argument");
    WorldImpl *w = impl(h);
    NANS_CUDA(cudaSetDevice(w->device));
    NANS_CUDA(cudaMemcpyAsync(w->h_counters, w->d.counters, sizeof(Counters), cudaMemcpyDeviceToHost, w->stream));
    NANS_CUDA(cudaStreamSynchronize(w->stream));
    const Counters *c = reinterpret_cast<const Counters *>(w->h_counters);
    memset(out, 0, sizeof(*out));
    out->n_pairs = c->n_pairs; out->n_contacts = c->n_contacts; out->n_gjk_found = c->n_gjk_found;
    out->solver_levels = c->solver_levels; out->overflow = c->overflow; out->max_epa_faces = c->max_epa_faces;
    { const int frc = solver_accum_fallbacks(w, &out->accum_fallbacks); if (frc) return frc; }
    if (c->pad[1]) return fail(NANS_ERR_STATE, "solver: dependency schedule stalled (watchdog: no row arrived for seconds)");
    if (c->overflow) {
        snprintf(g_err, sizeof(g_err), "capacity exceeded (overflow bits 0x%x: 1 pairs, 2 contacts, 4 EPA faces, 8 EPA edges)",
                 c->overflow);
        return NANS_ERR_CAPACITY;
    }
    return NANS_OK;
}

int nans_get_contacts(nans_world *h, nans_contact *out, int32_t cap, int32_t *count)
{
    if (!h || !count) return fail(NANS_ERR_ARG, "null argument");
    WorldImpl *w = impl(h);
    DeviceWorld &d = w->d;
    NANS_CUDA(cudaSetDevice(w->device));
    nans_step_stats st;
    int rc = nans_get_stats(h, &st);
    if (rc && rc != NANS_ERR_CAPACITY) return rc;
    const int n = st.n_contacts;
    *count = n;
    if (!out || n == 0) return rc;
    const int m = n < cap ? n : cap;
    std::vector<int32_t> a(m), b(m);
    std::vector<float4> pa(m), pb(m), nn(m);
    cudaStream_t s = w->stream;
    NANS_CUDA(cudaMemcpyAsync(a.data(), d.c_a, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s));
    NANS_CUDA(cudaMemcpyAsync(b.data(), d.c_b, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s));
    NANS_CUDA(cudaMemcpyAsync(pa.data(), d.c_pa, sizeof(float4) * m, cudaMemcpyDeviceToHost, s));
    NANS_CUDA(cudaMemcpyAsync(pb.data(), d.c_pb, sizeof(float4) * m, cudaMemcpyDeviceToHost, s));
    NANS_CUDA(cudaMemcpyAsync(nn.data(), d.c_n, sizeof(float4) * m, cudaMemcpyDeviceToHost, s));
    NANS_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < m; ++i) {
        nans_contact &c = out[i];
        const int ra = a[i], rb = b[i];
        const bool a_sph = ra >= d.n_cubes;
        c.a = a_sph ? ra - d.n_cubes : ra;
        if (rb < 0) { c.b = -rb - 1; c.type = a_sph ? NANS_SF : NANS_CF; }
        else {
            const bool b_sph = rb >= d.n_cubes;
            c.b = b_sph ? rb - d.n_cubes : rb;
            c.type = a_sph ? NANS_SS : (b_sph ? NANS_CS : NANS_CC);
        }
        c.point_a[0] = pa[i].x; c.point_a[1] = pa[i].y; c.point_a[2] = pa[i].z;
        c.point_b[0] = pb[i].x; c.point_b[1] = pb[i].y; c.point_b[2] = pb[i].z;
        c.n[0] = nn[i].x; c.n[1] = nn[i].y; c.n[2] = nn[i].z;
    }
    return rc;
}

int nans_set_contacts(nans_world *h, const nans_contact *in, int32_t count)
{
    if (!h || (count > 0 && !in)) return fail(NANS_ERR_ARG, "null argument");
    WorldImpl *w = impl(h);
    DeviceWorld &d = w->d;
    if (count < 0 || count > d.max_contacts) return fail(NANS_ERR_CAPACITY, "nans_set_contacts: over capacity");
    NANS_CUDA(cudaSetDevice(w->device));
    std::vector<int32_t> a(count), b(count);
    std::vector<float4> pa(count), pb(count), nn(count);
    for (int i = 0; i < count; ++i) {
        const nans_contact &c = in[i];
        const bool a_sph = (c.type == NANS_SS || c.type == NANS_SF);
        a[i] = a_sph ? d.n_cubes + c.a : c.a;
        if (c.type == NANS_CF || c.type == NANS_SF) b[i] = -(c.b + 1);
        else b[i] = (c.type == NANS_CS || c.type == NANS_SS) ? d.n_cubes + c.b : c.b;
        pa[i] = make_float4(c.point_a[0], c.point_a[1], c.point_a[2], 0.f);
        pb[i] = make_float4(c.point_b[0], c.point_b[1], c.point_b[2], 0.f);
        memcpy(&pa[i].w, &a[i], 4);   // body rows ride in the w lanes (solver.cu)
        memcpy(&pb[i].w, &b[i], 4);
        nn[i] = make_float4(c.n[0], c.n[1], c.n[2], 0.f);
    }
    cudaStream_t s = w->stream;
    NANS_CUDA(cudaMemsetAsync(d.counters, 0, sizeof(Counters), s));
    if (count > 0) {
        NANS_CUDA(cudaMemcpyAsync(d.c_a, a.data(), sizeof(int32_t) * count, cudaMemcpyHostToDevice, s));
        NANS_CUDA(cudaMemcpyAsync(d.c_b, b.data(), sizeof(int32_t) * count, cudaMemcpyHostToDevice, s));
        NANS_CUDA(cudaMemcpyAsync(d.c_pa, pa.data(), sizeof(float4) * count, cudaMemcpyHostToDevice, s));
        NANS_CUDA(cudaMemcpyAsync(d.c_pb, pb.data(), sizeof(float4) * count, cudaMemcpyHostToDevice, s));
        NANS_CUDA(cudaMemcpyAsync(d.c_n, nn.data(), sizeof(float4) * count, cudaMemcpyHostToDevice, s));
    }
    NANS_CUDA(cudaMemcpyAsync(&d.counters->n_contacts, &count, sizeof(int32_t), cudaMemcpyHostToDevice, s));
    NANS_CUDA(cudaStreamSynchronize(s));
    w->have_contacts = true;
    return NANS_OK;
}

// debug: per-contact start time (ns, %globaltimer) and DAG level of the last solve; needs NANS_SOLVER_TRACE=1
int nans_debug_solver_trace(nans_world *h, uint64_t *times, int32_t *levels, int32_t cap)
{
    if (!h) return fail(NANS_ERR_ARG, "null world");
    WorldImpl *w = impl(h);
    NANS_CUDA(cudaSetDevice(w->device));
    NANS_CUDA(cudaStreamSynchronize(w->stream));
    if (times) NANS_CUDA(cudaMemcpy(times, w->d.pair_out, sizeof(uint64_t) * 4 * cap, cudaMemcpyDeviceToHost));   // 4 slots per contact
    if (levels) NANS_CUDA(cudaMemcpy(levels, w->d.trace_level, sizeof(int32_t) * cap, cudaMemcpyDeviceToHost));
    return NANS_OK;
}

int nans_get_pairs(nans_world *h, int32_t *pa, int32_t *pb, int32_t cap, int32_t *count)
{
    if (!h || !count) return fail(NANS_ERR_ARG, "null argument");
    WorldImpl *w = impl(h);
    nans_step_stats st;
    int rc = nans_get_stats(h, &st);
    if (rc && rc != NANS_ERR_CAPACITY) return rc;
    *count = st.n_pairs;
    const int m = st.n_pairs < cap ? st.n_pairs : cap;
    if (m > 0 && pa && pb) {
        NANS_CUDA(cudaMemcpyAsync(pa, w->d.pair_a, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, w->stream));
        NANS_CUDA(cudaMemcpyAsync(pb, w->d.pair_b, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, w->stream));
        NANS_CUDA(cudaStreamSynchronize(w->stream));
    }
    return rc;
}

// ---- stand-alone narrowphase ---------------------------------------------------------------------
static char *g_np_dev_block[64] = {nullptr};   // per device: work counter + Counters of the device-resident batch entry
int nans_check_collision_device(int32_t n, const int32_t *d_type, const float *d_posrad_a, const float *d_verts_a,
                                const float *d_posrad_b, const float *d_verts_b, int32_t *d_hit, float *d_out,
                                void *stream)
{
    int dev = 0;
    NANS_CUDA(cudaGetDevice(&dev));
    char *&blk = g_np_dev_block[(dev >= 0 && dev < 64) ? dev : 0];   // work counter at +0, Counters at +256
    if (!blk) { NANS_CUDA(cudaMalloc(&blk, 512)); NANS_CUDA(cudaMemsetAsync(blk, 0, 512, (cudaStream_t)stream)); }
    return launch_narrowphase_batch(n, d_type, (const float4 *)d_posrad_a, (const float4 *)d_verts_a,
                                    (const float4 *)d_posrad_b, (const float4 *)d_verts_b, d_hit, nullptr,
                                    (float4 *)d_out, (int *)blk, (Counters *)(blk + 256), (cudaStream_t)stream);
}

// EPA-arena overflow bits accumulated by nans_check_collision_device on the current device since the last call
// (an overflowing pair is otherwise indistinguishable from a miss); synchronises the device.
int nans_check_collision_device_status(int32_t *overflow_bits)
{
    if (!overflow_bits) return fail(NANS_ERR_ARG, "null argument");
    int dev = 0;
    NANS_CUDA(cudaGetDevice(&dev));
    char *blk = g_np_dev_block[(dev >= 0 && dev < 64) ? dev : 0];
    *overflow_bits = 0;
    if (!blk) return NANS_OK;
    Counters c;
    NANS_CUDA(cudaDeviceSynchronize());
    NANS_CUDA(cudaMemcpy(&c, blk + 256, sizeof(c), cudaMemcpyDeviceToHost));
    NANS_CUDA(cudaMemset(blk + 256, 0, sizeof(c)));
    *overflow_bits = c.overflow;
    if (c.overflow) {
        snprintf(g_err, sizeof(g_err), "EPA arena capacity exceeded (bits 0x%x)", c.overflow);
        return NANS_ERR_CAPACITY;
    }
    return NANS_OK;
}

int nans_check_collision_batch(int32_t n, const int32_t *type, const float *pos_a, const float *verts_a,
                               const float *rad_a, const float *pos_b, const float *verts_b, const float *rad_b,
                               int32_t *hit, int32_t *gjk, float *out_n, float *out_pa, float *out_pb,
                               int32_t device)
{
    if (n < 0 || (n > 0 && (!type || !pos_a || !pos_b || !hit))) return fail(NANS_ERR_ARG, "null argument");
    if (n == 0) return NANS_OK;
    int ndev = 0;
    NANS_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(NANS_ERR_CUDA, "no such CUDA device");
    NANS_CUDA(cudaSetDevice(device));
    const size_t N = (size_t)n;
    std::vector<float4> pra(N), prb(N);
    for (size_t i = 0; i < N; ++i) {
        pra[i] = make_float4(pos_a[3 * i], pos_a[3 * i + 1], pos_a[3 * i + 2], rad_a ? rad_a[i] : 0.f);
        prb[i] = make_float4(pos_b[3 * i], pos_b[3 * i + 1], pos_b[3 * i + 2], rad_b ? rad_b[i] : 0.f);
    }
    char *blk = nullptr;
    Bump b{nullptr, 0};
    auto plan = [&](Bump &q, int32_t *&dt, float4 *&dpa, float4 *&dva, float4 *&dpb, float4 *&dvb, int32_t *&dh,
                    int32_t *&dg, float4 *&dout, int *&dwork, Counters *&dc) {
        dt = q.take<int32_t>(N); dpa = q.take<float4>(N); dva = q.take<float4>(6 * N); dpb = q.take<float4>(N);
        dvb = q.take<float4>(6 * N); dh = q.take<int32_t>(N); dg = q.take<int32_t>(N); dout = q.take<float4>(3 * N);
        dwork = q.take<int>(64); dc = q.take<Counters>(1);
    };
    int32_t *dt, *dh, *dg; float4 *dpa, *dva, *dpb, *dvb, *dout; int *dwork; Counters *dc;
    plan(b, dt, dpa, dva, dpb, dvb, dh, dg, dout, dwork, dc);
    NANS_CUDA(cudaMalloc(&blk, b.used + 256));
    Bump q{blk, 0};
    plan(q, dt, dpa, dva, dpb, dvb, dh, dg, dout, dwork, dc);
    cudaStream_t s = 0;
    int rc = NANS_OK;
    auto body = [&]() -> int {
        NANS_CUDA(cudaMemsetAsync(blk, 0, b.used, s));
        NANS_CUDA(cudaMemcpyAsync(dt, type, sizeof(int32_t) * N, cudaMemcpyHostToDevice, s));
        NANS_CUDA(cudaMemcpyAsync(dpa, pra.data(), sizeof(float4) * N, cudaMemcpyHostToDevice, s));
        NANS_CUDA(cudaMemcpyAsync(dpb, prb.data(), sizeof(float4) * N, cudaMemcpyHostToDevice, s));
        if (verts_a) NANS_CUDA(cudaMemcpyAsync(dva, verts_a, sizeof(float) * 24 * N, cudaMemcpyHostToDevice, s));
        if (verts_b) NANS_CUDA(cudaMemcpyAsync(dvb, verts_b, sizeof(float) * 24 * N, cudaMemcpyHostToDevice, s));
        int r = launch_narrowphase_batch(n, dt, dpa, dva, dpb, dvb, dh, dg, dout, dwork, dc, s);
        if (r) return r;
        std::vector<float4> ho(3 * N);
        Counters hc;
        NANS_CUDA(cudaMemcpyAsync(hit, dh, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, s));
        if (gjk) NANS_CUDA(cudaMemcpyAsync(gjk, dg, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, s));
        NANS_CUDA(cudaMemcpyAsync(ho.data(), dout, sizeof(float4) * 3 * N, cudaMemcpyDeviceToHost, s));
        NANS_CUDA(cudaMemcpyAsync(&hc, dc, sizeof(Counters), cudaMemcpyDeviceToHost, s));
        NANS_CUDA(cudaStreamSynchronize(s));
        for (size_t i = 0; i < N; ++i) {
            const float4 a = ho[3 * i], bb = ho[3 * i + 1], c = ho[3 * i + 2];
            if (out_pa) { out_pa[3 * i] = a.x; out_pa[3 * i + 1] = a.y; out_pa[3 * i + 2] = a.z; }
            if (out_pb) { out_pb[3 * i] = bb.x; out_pb[3 * i + 1] = bb.y; out_pb[3 * i + 2] = bb.z; }
            if (out_n) { out_n[3 * i] = c.x; out_n[3 * i + 1] = c.y; out_n[3 * i + 2] = c.z; }
        }
        if (hc.overflow) {
            snprintf(g_err, sizeof(g_err), "EPA arena capacity exceeded (bits 0x%x)", hc.overflow);
            return NANS_ERR_CAPACITY;
        }
        return NANS_OK;
    };
    rc = body();
    cudaFree(blk);
    return rc;
}

}  // extern "C"
