// slab.cu — ONE world over several GPUs: spatial slabs, NCCL halo exchange, solver coupling over NVLink peer memory.
//
// The reference is a single process with a 16-cube cap (code/nans.h:52-53) and an all-pairs loop
// (code/nans.cpp:1357-1364); this is the B200 extension of its step to one world too large or too slow for one
// GPU (BASELINE.json config 5), kept EXACT: the decomposed world is bit-identical to the same
// world stepped on one GPU, i.e. to the reference's sequential sweep.
//
// Decomposition.  The world's bodies are numbered slab-major: rank r owns the contiguous global index range
// [gid_base_r, gid_base_r + n_owned_r), and the scene builder lays those ranges out as slabs along x
// (scenes.cube_pile_slabs), so index ranges ARE spatial x-slabs.  Locally a rank's world holds its owned rows
// first and, behind them, GHOST rows: copies of the upper neighbour's bodies whose AABB reaches into this rank's
// bounding box.  Local row order = global index order, so the locally emitted pair list is the global list
// restricted to pairs whose lower-index body is owned: every pair of the world is tested and solved exactly
// once, by the owner of its lower-index body.
//
// One step, all of it queued on the world's stream — no host round trip, no device-to-host copy:
//   1. integrate forces (owned rows);  AABBs + bounding box of the owned rows
//   2. ncclAllGather of the boxes (8 floats per rank), staying in device memory
//   3. halo selection against the LOWER neighbour's box read from device memory (flag, scan, pack): a
//      fixed-capacity message {count, records[cap]} of 160 B records; the selected rows' solver rows are marked
//      PENDING (they will be released by the lower rank, see 6)
//   4. grouped ncclSend (to rank-1) / ncclRecv (from rank+1) of the fixed-size message
//   5. unpack into the ghost rows; the live row count (owned + ghosts) stays on the device and bounds every
//      kernel of the detection phase (DeviceWorld::live)
//   6. detection and the exact-order solve as on one GPU.  The solve is ONE dataflow across all GPUs: when a
//      rank applies its last contact on a ghost body (or finds it has none), it stores the body's velocity rows
//      straight into the owner's memory (st.relaxed.sys.b128 over NVLink, version 0 of the owner's sequence);
//      the owner's contacts poll their own rows as they always do.  In the reference's sweep order every
//      contact a lower rank applies to body b precedes every contact b's owner applies, so this IS the
//      sequential order; no rank ever waits for a higher rank, so the flow cannot deadlock.
//   7. integrate velocities + vertex rebuild (owned rows)
//
// Requirements, checked on the device every step and reported as sticky errors (nans_slab_status): an owned
// body may reach only into the box of the rank directly below (slabs thicker than a body; an exploded world
// fails loudly instead of silently dropping an exchange), and the halo must fit the message capacity.
// Bodies do not migrate between ranks (ownership is by index).
#include <dlfcn.h>
#include <float.h>
#include <string.h>

#include "world.cuh"

#if __has_include(<nccl.h>)
#include <nccl.h>
#else
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclFloat32 = 7, ncclFloat = 7 } ncclDataType_t;
#endif

namespace nans {

// ---- NCCL through dlopen (no link-time dependency: single-GPU users never load it; inside a torch process the
// already-loaded libnccl.so.2 is picked up, so there is one NCCL per process) -------------------------------
struct NcclApi {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl = {};

static int nccl_load()
{
    if (g_nccl.lib) return NANS_OK;
    const char *names[] = {getenv("NANS_NCCL_LIB"), "libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { snprintf(g_err, sizeof(g_err), "slab mode needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror()); return NANS_ERR_STATE; }
#define NANS_SYM(field, name)                                                                         \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                                        \
    if (!g_nccl.field) { snprintf(g_err, sizeof(g_err), "NCCL symbol %s missing", name); return NANS_ERR_STATE; }
    NANS_SYM(GetUniqueId, "ncclGetUniqueId");
    NANS_SYM(CommInitRank, "ncclCommInitRank");
    NANS_SYM(CommDestroy, "ncclCommDestroy");
    NANS_SYM(AllGather, "ncclAllGather");
    NANS_SYM(Send, "ncclSend");
    NANS_SYM(Recv, "ncclRecv");
    NANS_SYM(GroupStart, "ncclGroupStart");
    NANS_SYM(GroupEnd, "ncclGroupEnd");
    NANS_SYM(GetErrorString, "ncclGetErrorString");
#undef NANS_SYM
    g_nccl.lib = h;
    return NANS_OK;
}

#define NANS_NCCL(expr)                                                                                  \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != ncclSuccess) {                                                                         \
            snprintf(g_err, sizeof(g_err), "%s:%d %s: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
            return NANS_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

constexpr int kHaloQuads = 10;        // halo record: pos, vel, angvel, 6 x verts, (global id, owner row, 0, 0) = 160 B
constexpr int kBoundsThreads = 256;
constexpr int kPendingTag = 0xffffe;  // solver.cu: the row is owed by the lower neighbour
constexpr int kMaxRanks = 64;

struct SlabState {
    int rank, nranks, n_owned, ghost_cap, gid_base;
    ncclComm_t comm;
    char *blk;                 // one device block for everything below
    float4 *halo_send, *halo_recv;   // [1 + kHaloQuads * ghost_cap]: header {count, 0, 0, 0} then the records
    float *bounds_own;         // [8]   lo.xyz, -, hi.xyz, -
    float *bounds_all;         // [nranks][8]
    float *bounds_scratch;     // [6 * 1024]
    int32_t *live;             // device: owned + this step's ghosts
    int32_t *err;              // device: sticky SLAB_ERR_* bits
    int32_t *h_status;         // pinned: {err, live, sent count}
    void *peer_base;           // the upper neighbour's arena, opened through CUDA IPC
    size_t halo_floats;        // floats per message
    cudaGraphExec_t graph;     // the whole step (NCCL calls included), captured on the second call
    int graph_state;           // 0 first call pending, 1 eager step done, 2 captured, -1 capture not possible
    unsigned graph_launches;
};

struct IpcBlob {               // what nans_slab_ipc_handle hands out (128 bytes)
    cudaIpcMemHandle_t handle; // 64 bytes
    uint64_t arena_bytes;
    int32_t nb, rank;
    char pad[128 - 64 - 16];
};
static_assert(sizeof(IpcBlob) == 128, "IpcBlob is 128 bytes");

// ---- kernels ----------------------------------------------------------------------------------------------
__global__ void slab_begin_kernel(DeviceWorld w, int32_t *live)
{
    *live = w.n_owned;
}

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
__global__ void bounds_final_kernel(const float *partial, int blocks, float *out8)
{
    const int k = threadIdx.x;
    if (k >= 6) return;
    float v = blocks > 0 ? partial[k] : (k < 3 ? FLT_MAX : -FLT_MAX);   // no owned bodies: an empty box
    for (int b = 1; b < blocks; ++b) v = k < 3 ? fminf(v, partial[6 * b + k]) : fmaxf(v, partial[6 * b + k]);
    out8[k < 3 ? k : k + 1] = v;          // lo at [0..2], hi at [4..6]
}

__device__ __forceinline__ bool box_overlap(const float4 &a, const float4 &b, const float *box)
{
    return a.x <= box[4] && box[0] <= b.x && a.y <= box[5] && box[1] <= b.y && a.z <= box[6] && box[2] <= b.z;
}

// flag owned rows whose AABB overlaps the lower neighbour's box (inclusive, both already inflated); a row that
// reaches into the box of any rank further down breaks the exchange pattern: sticky error
__global__ void __launch_bounds__(256) halo_flag_kernel(DeviceWorld w, const float *bounds_all, int rank, int32_t *err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > w.n_owned) return;
    int f = 0;
    if (i < w.n_owned) {
        const float4 a = w.aabb_lo[i], b = w.aabb_hi[i];
        f = rank > 0 && box_overlap(a, b, bounds_all + 8 * (rank - 1));
        for (int q = 0; q < rank - 1; ++q)
            if (box_overlap(a, b, bounds_all + 8 * q)) { atomicOr(err, SLAB_ERR_NOT_ADJACENT); break; }
    }
    w.pair_hit[i] = f;      // [n_owned] holds 0 so the exclusive scan yields the total there
}

template <bool SYS>
__device__ __forceinline__ void st_row_tag(float4 *p, float4 v, int tag)
{
    const unsigned long long lo = (unsigned long long)__float_as_uint(v.x) | ((unsigned long long)__float_as_uint(v.y) << 32);
    const unsigned long long hi = (unsigned long long)__float_as_uint(v.z) | ((unsigned long long)(unsigned)tag << 32);
    asm volatile("{\n\t.reg .b128 r;\n\tmov.b128 r, {%1, %2};\n\tst.relaxed.sys.global.b128 [%0], r;\n\t}"
                 :: "l"(p), "l"(lo), "l"(hi) : "memory");
}

// order-preserving pack (ascending row = ascending global id) + PENDING marks on the rows that leave
__global__ void __launch_bounds__(256) halo_pack_kernel(DeviceWorld w, int gid_base, float4 *msg, int cap, int32_t *err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        const int total = (int)w.pair_hit_scan[w.n_owned];
        if (total > cap) atomicOr(err, SLAB_ERR_HALO_CAP);
        msg[0] = make_float4(__int_as_float(min(total, cap)), 0.f, 0.f, 0.f);
    }
    if (i >= w.n_owned) return;
    const int k = (int)w.pair_hit_scan[i];
    const bool sent = w.pair_hit[i] && k < cap;
    w.sent_mark[i] = sent ? 1 : 0;
    if (!sent) return;
    float4 *o = msg + 1 + (size_t)kHaloQuads * k;
    const float4 v = w.vel[i], a = w.angvel[i];
    o[0] = w.pos[i]; o[1] = v; o[2] = a;
#pragma unroll
    for (int q = 0; q < 6; ++q) o[3 + q] = w.verts[6 * (size_t)i + q];
    o[9] = make_float4(__int_as_float(gid_base + i), __int_as_float(i), 0.f, 0.f);
    // the lower rank will release these rows (solver.cu); they must read PENDING before the message leaves
    st_row_tag<true>(w.row_v + i, v, kPendingTag);
    st_row_tag<true>(w.row_w + i, a, kPendingTag);
}

__global__ void __launch_bounds__(256) halo_unpack_kernel(DeviceWorld w, const float4 *msg, int cap, int32_t *live)
{
    const int count = min(__float_as_int(msg[0].x), cap);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) *live = w.n_owned + count;
    if (k >= count) return;
    const float4 *r = msg + 1 + (size_t)kHaloQuads * k;
    const int row = w.n_owned + k;
    w.pos[row] = r[0]; w.vel[row] = r[1]; w.angvel[row] = r[2];
#pragma unroll
    for (int q = 0; q < 6; ++q) w.verts[6 * (size_t)row + q] = r[3 + q];
    w.force[row] = make_float4(0, 0, 0, 0);
    w.torque[row] = make_float4(0, 0, 0, 0);
    w.gid[row] = __float_as_int(r[9].x);
    w.ghost_owner_row[row] = __float_as_int(r[9].y);
}

static SlabState *state(World *w) { return reinterpret_cast<SlabState *>(w->slab); }

void slab_destroy(World *w)
{
    SlabState *S = state(w);
    if (!S) return;
    if (S->graph) cudaGraphExecDestroy(S->graph);
    if (S->peer_base) cudaIpcCloseMemHandle(S->peer_base);
    if (S->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(S->comm);
    if (S->blk) cudaFree(S->blk);
    if (S->h_status) cudaFreeHost(S->h_status);
    delete S;
    w->slab = nullptr;
}

}  // namespace nans

using namespace nans;

extern "C" {

int nans_slab_unique_id(void *out128)
{
    if (!out128) { snprintf(g_err, sizeof(g_err), "null argument"); return NANS_ERR_ARG; }
    int rc = nccl_load();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NANS_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return NANS_OK;
}

int nans_slab_init(nans_world *h, int32_t rank, int32_t nranks, const void *unique_id128, int32_t n_owned,
                   int32_t gid_base, int32_t halo_cap)
{
    World *w = reinterpret_cast<World *>(h);
    if (!w || !unique_id128) { snprintf(g_err, sizeof(g_err), "null argument"); return NANS_ERR_ARG; }
    DeviceWorld &d = w->d;
    if (d.n_spheres != 0 || d.world_id) { snprintf(g_err, sizeof(g_err), "slab mode supports cube-only single worlds"); return NANS_ERR_ARG; }
    if (rank < 0 || nranks < 1 || rank >= nranks || nranks > kMaxRanks || n_owned < 0 || halo_cap < 0 ||
        (long long)n_owned + halo_cap > d.nb) {
        snprintf(g_err, sizeof(g_err), "nans_slab_init: bad rank, or owned + halo capacity exceed the world's %d rows", d.nb);
        return NANS_ERR_ARG;
    }
    if (w->slab) { snprintf(g_err, sizeof(g_err), "nans_slab_init: already initialised"); return NANS_ERR_STATE; }
    int rc = nccl_load();
    if (rc) return rc;
    NANS_CUDA(cudaSetDevice(w->device));
    SlabState *S = new SlabState();
    memset(S, 0, sizeof(*S));
    S->rank = rank; S->nranks = nranks; S->n_owned = n_owned; S->gid_base = gid_base;
    S->ghost_cap = halo_cap;             // the same on every rank: the messages have a fixed size
    S->halo_floats = 4 * (1 + (size_t)kHaloQuads * (size_t)(S->ghost_cap > 0 ? S->ghost_cap : 1));
    // one block: two messages, bounds, scratch, counters (allocated once here, never while stepping)
    const size_t msg_bytes = (sizeof(float) * S->halo_floats + 255) & ~(size_t)255;
    const size_t bytes = 2 * msg_bytes + 256 + 32 * kMaxRanks + 4 * 6 * 1024 + 512;
    if (cudaMalloc(&S->blk, bytes) != cudaSuccess) { delete S; snprintf(g_err, sizeof(g_err), "cudaMalloc(%zu) for the halo buffers failed", bytes); return NANS_ERR_CUDA; }
    cudaMemset(S->blk, 0, bytes);
    char *p = S->blk;
    S->halo_send = (float4 *)p; p += msg_bytes;
    S->halo_recv = (float4 *)p; p += msg_bytes;
    S->bounds_own = (float *)p; p += 256;
    S->bounds_all = (float *)p; p += 32 * kMaxRanks;
    S->bounds_scratch = (float *)p; p += 4 * 6 * 1024;
    S->live = (int32_t *)p; p += 256;
    S->err = (int32_t *)p;
    w->slab = S;
    NANS_CUDA(cudaMallocHost((void **)&S->h_status, 64));
    ncclUniqueId id;
    memcpy(&id, unique_id128, 128);
    NANS_NCCL(g_nccl.CommInitRank(&S->comm, nranks, id, rank));
    d.n_owned = n_owned;
    d.live = S->live;
    // sent_mark / ghost_owner_row / gid / row_v / row_w live in the arena (carved for every world)
    NANS_CUDA(cudaMemsetAsync(d.sent_mark, 0, sizeof(int32_t) * (size_t)d.nb, w->stream));
    slab_begin_kernel<<<1, 1, 0, w->stream>>>(d, S->live);
    NANS_LAUNCH_CHECK();
    NANS_CUDA(cudaStreamSynchronize(w->stream));
    return NANS_OK;
}

int nans_slab_ipc_handle(nans_world *h, void *out128)
{
    World *w = reinterpret_cast<World *>(h);
    if (!w || !out128 || !w->slab) { snprintf(g_err, sizeof(g_err), "nans_slab_ipc_handle: not a slab world"); return NANS_ERR_ARG; }
    if (!w->owns_arena) { snprintf(g_err, sizeof(g_err), "slab mode needs a library-owned arena (cudaMalloc) to export it over CUDA IPC"); return NANS_ERR_ARG; }
    NANS_CUDA(cudaSetDevice(w->device));
    IpcBlob b;
    memset(&b, 0, sizeof(b));
    NANS_CUDA(cudaIpcGetMemHandle(&b.handle, w->arena));
    b.arena_bytes = w->arena_bytes; b.nb = w->d.nb; b.rank = state(w)->rank;
    memcpy(out128, &b, sizeof(b));
    return NANS_OK;
}

// handles: [nranks][128], rank-major.  Opens the UPPER neighbour's arena (the only peer this rank writes to).
int nans_slab_connect(nans_world *h, const void *handles)
{
    World *w = reinterpret_cast<World *>(h);
    if (!w || !handles || !w->slab) { snprintf(g_err, sizeof(g_err), "nans_slab_connect: not a slab world"); return NANS_ERR_ARG; }
    SlabState *S = state(w);
    NANS_CUDA(cudaSetDevice(w->device));
    if (S->rank + 1 >= S->nranks) return NANS_OK;       // the top rank has no ghosts
    IpcBlob b;
    memcpy(&b, (const char *)handles + 128 * (size_t)(S->rank + 1), sizeof(b));
    if (b.rank != S->rank + 1 || b.arena_bytes != w->arena_bytes || b.nb != w->d.nb) {
        snprintf(g_err, sizeof(g_err), "nans_slab_connect: rank %d's world was created with other capacities "
                 "(every rank must use the same n_cubes / max_pairs / max_contacts so that the arena layouts match)", S->rank + 1);
        return NANS_ERR_ARG;
    }
    NANS_CUDA(cudaIpcOpenMemHandle(&S->peer_base, b.handle, cudaIpcMemLazyEnablePeerAccess));
    const ptrdiff_t off_v = (const char *)w->d.row_v - (const char *)w->arena;
    const ptrdiff_t off_w = (const char *)w->d.row_w - (const char *)w->arena;
    w->d.peer_row_v = (float4 *)((char *)S->peer_base + off_v);
    w->d.peer_row_w = (float4 *)((char *)S->peer_base + off_w);
    return NANS_OK;
}

// everything of one step except dt (which travels through device memory, so the captured graph does not depend on it)
static int slab_step_body(World *w)
{
    SlabState *S = state(w);
    DeviceWorld &d = w->d;
    cudaStream_t s = w->stream;
    int rc;
    slab_begin_kernel<<<1, 1, 0, s>>>(d, S->live);                      // live rows = owned rows
    NANS_LAUNCH_CHECK();
    if ((rc = launch_integrate_forces(w))) return rc;
    if ((rc = launch_aabb_only(w))) return rc;                          // owned rows (live = n_owned)
    const int n = d.n_owned;
    const int blocks = n > 0 ? min(div_up(n, kBoundsThreads), 1024) : 0;
    if (blocks) { bounds_partial_kernel<<<blocks, kBoundsThreads, 0, s>>>(d, n, S->bounds_scratch); NANS_LAUNCH_CHECK(); }
    bounds_final_kernel<<<1, 32, 0, s>>>(S->bounds_scratch, blocks, S->bounds_own);
    NANS_LAUNCH_CHECK();
    NANS_NCCL(g_nccl.AllGather(S->bounds_own, S->bounds_all, 8, ncclFloat, S->comm, s));
    halo_flag_kernel<<<div_up(n + 1, 256), 256, 0, s>>>(d, S->bounds_all, S->rank, S->err);
    NANS_LAUNCH_CHECK();
    if ((rc = exclusive_scan_u32((const uint32_t *)d.pair_hit, d.pair_hit_scan, n + 1, d.scan_block, s))) return rc;
    halo_pack_kernel<<<div_up(n > 0 ? n : 1, 256), 256, 0, s>>>(d, S->gid_base, S->halo_send, S->ghost_cap, S->err);
    NANS_LAUNCH_CHECK();
    NANS_NCCL(g_nccl.GroupStart());
    if (S->rank > 0) NANS_NCCL(g_nccl.Send(S->halo_send, S->halo_floats, ncclFloat, S->rank - 1, S->comm, s));
    if (S->rank + 1 < S->nranks) NANS_NCCL(g_nccl.Recv(S->halo_recv, S->halo_floats, ncclFloat, S->rank + 1, S->comm, s));
    NANS_NCCL(g_nccl.GroupEnd());
    if (S->rank + 1 < S->nranks && S->ghost_cap > 0) {
        halo_unpack_kernel<<<div_up(S->ghost_cap, 256), 256, 0, s>>>(d, S->halo_recv, S->ghost_cap, S->live);
        NANS_LAUNCH_CHECK();
    }
    if ((rc = launch_broadphase(w))) return rc;
    if ((rc = launch_narrowphase(w))) return rc;
    if ((rc = launch_contacts(w))) return rc;
    w->have_contacts = true;
    if ((rc = launch_solver(w))) return rc;
    return launch_integrate_velocities(w);
}

// One step of this rank's share of the world, exchanges included; asynchronous on the world's stream.  Like
// nans_step, the step is ~45 launches (two of them NCCL) with frame-independent parameters: the first call runs
// eagerly, the second is captured into ONE CUDA graph (NCCL supports capture; every rank captures the same
// sequence), later calls replay it.  NANS_GRAPH=0 disables.
int nans_slab_step(nans_world *h, float dt)
{
    World *w = reinterpret_cast<World *>(h);
    if (!w || !w->slab) { snprintf(g_err, sizeof(g_err), "nans_slab_step: not a slab world"); return NANS_ERR_ARG; }
    SlabState *S = state(w);
    NANS_CUDA(cudaSetDevice(w->device));
    int rc;
    if ((rc = launch_set_dt(w, dt))) return rc;
    static int enabled = -1;
    if (enabled < 0) { const char *e = getenv("NANS_GRAPH"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
    if (!enabled || S->graph_state < 0) return slab_step_body(w);
    if (S->graph_state == 2) {
        NANS_CUDA(cudaGraphLaunch(S->graph, w->stream));
        g_launches += S->graph_launches;
        w->have_contacts = true;
        return NANS_OK;
    }
    if (S->graph_state == 1) {
        const unsigned long long before = g_launches;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            rc = slab_step_body(w);
            const cudaError_t ee = cudaStreamEndCapture(w->stream, &graph);
            ok = !rc && ee == cudaSuccess && graph && cudaGraphInstantiate(&S->graph, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
        }
        S->graph_launches = (unsigned)(g_launches - before);
        g_launches = before;
        if (!ok) {
            cudaGetLastError();
            S->graph = nullptr;
            S->graph_state = -1;                 // stay eager (every rank takes the same decision or NCCL would hang:
            return slab_step_body(w);            // capture support is a property of the NCCL / driver pair, equal on all ranks)
        }
        S->graph_state = 2;
        NANS_CUDA(cudaGraphLaunch(S->graph, w->stream));
        g_launches += S->graph_launches;
        w->have_contacts = true;
        return NANS_OK;
    }
    S->graph_state = 1;
    return slab_step_body(w);
}

// synchronises; err_bits: SLAB_ERR_* (sticky), live_rows: owned + ghosts of the last step
int nans_slab_status(nans_world *h, int32_t *err_bits, int32_t *live_rows, int64_t *halo_bytes_per_message)
{
    World *w = reinterpret_cast<World *>(h);
    if (!w || !w->slab) { snprintf(g_err, sizeof(g_err), "nans_slab_status: not a slab world"); return NANS_ERR_ARG; }
    SlabState *S = state(w);
    NANS_CUDA(cudaSetDevice(w->device));
    NANS_CUDA(cudaMemcpyAsync(S->h_status, S->err, 4, cudaMemcpyDeviceToHost, w->stream));
    NANS_CUDA(cudaMemcpyAsync(S->h_status + 1, S->live, 4, cudaMemcpyDeviceToHost, w->stream));
    NANS_CUDA(cudaStreamSynchronize(w->stream));
    if (err_bits) *err_bits = S->h_status[0];
    if (live_rows) *live_rows = S->h_status[1];
    if (halo_bytes_per_message) *halo_bytes_per_message = (int64_t)(sizeof(float) * S->halo_floats);
    if (S->h_status[0]) {
        snprintf(g_err, sizeof(g_err), "slab exchange violated (bits 0x%x: 1 = halo larger than the message capacity, "
                 "2 = a body reaches into a rank other than the lower neighbour): the decomposed world is no longer exact",
                 S->h_status[0]);
        return NANS_ERR_STATE;
    }
    return NANS_OK;
}

// global ids of the live rows (owned: gid_base + row; ghosts: as received)
int nans_slab_row_gids(nans_world *h, int32_t *out, int32_t cap, int32_t *count)
{
    World *w = reinterpret_cast<World *>(h);
    if (!w || !w->slab || !count) { snprintf(g_err, sizeof(g_err), "nans_slab_row_gids: bad argument"); return NANS_ERR_ARG; }
    SlabState *S = state(w);
    NANS_CUDA(cudaSetDevice(w->device));
    int32_t live = 0;
    NANS_CUDA(cudaMemcpyAsync(&live, S->live, 4, cudaMemcpyDeviceToHost, w->stream));
    NANS_CUDA(cudaStreamSynchronize(w->stream));
    *count = live;
    if (!out) return NANS_OK;
    const int m = live < cap ? live : cap;
    for (int i = 0; i < m && i < S->n_owned; ++i) out[i] = S->gid_base + i;
    if (m > S->n_owned)
        NANS_CUDA(cudaMemcpy(out + S->n_owned, w->d.gid + S->n_owned, sizeof(int32_t) * (size_t)(m - S->n_owned), cudaMemcpyDeviceToHost));
    return NANS_OK;
}

}  // extern "C"
