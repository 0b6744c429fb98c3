// world.cuh — device-resident world: one arena, SoA body / shape / pair / contact regions.
//
// Replaces `sdl_state` (reference code/nans.h:374-386) and every std::vector on the step
// (Pairs code/nans.h:385, Simplex/Edge/Triangle code/nans.cpp:791-792,1370).  See DESIGN.md §3.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nans_b200.h"

namespace nans {

constexpr int kNumSMs = 148;            // B200
constexpr int kEpaMaxVerts = 72;        // 4 + 65 iterations (MAX_EPA_ITERATIONS, code/nans.h:56), padded
constexpr int kEpaMaxFaces = 200;       // manifold bound is 4 + 2*65 = 134; slack for degenerate horizons
constexpr int kEpaMaxEdges = 96;
constexpr int kMaxStatics = 16;

// overflow bits (nans_step_stats.overflow)
enum { OVF_PAIRS = 1, OVF_CONTACTS = 2, OVF_EPA_FACES = 4, OVF_EPA_EDGES = 8, OVF_CELL = 16 };

// Device counters block (one per world, zeroed per detect).
struct Counters {
    int32_t n_pairs;
    int32_t n_contacts;
    int32_t n_gjk_found;
    int32_t solver_levels;
    int32_t overflow;
    int32_t max_epa_faces;
    int32_t frontier_n[3];   // solver: [0] sweep tickets issued, [1] number of runs
    int32_t pad[11];         // [0] narrowphase work counter, [1] solver abort flag (watchdog), [2] max AABB extent bits,
                             // [3] largest Morton key of the step, [4] narrowphase hit-list length, [5] its chunk tickets,
                             // [6] deferred-GJK list length (narrowphase.cu),
                             // [7..9] smallest AABB centre per axis (order-encoded, complemented: 0 = none yet),
                             // [10] slab error bits (SLAB_ERR_*)
};

constexpr int kMinCentre = 7;                              // Counters::pad slots used by the broadphase
constexpr int kSlabErr = 10;                               // Counters::pad slot of the slab error bits
enum { SLAB_ERR_HALO_CAP = 1,      // more halo bodies than the fixed-capacity message holds
       SLAB_ERR_NOT_ADJACENT = 2 };// an owned body reaches into the box of a rank other than the lower neighbour

// All pointers are DEVICE pointers into the arena.
struct DeviceWorld {
    int32_t n_cubes, n_spheres, n_statics, nb;
    int32_t n_owned;   // slab mode: rows >= n_owned are ghosts of higher ranks (== nb otherwise)
    int32_t max_pairs, max_contacts;
    // --- bodies (SoA, float4): rows cubes then spheres
    float4 *pos;      // xyz = Position,  w = Mass
    float4 *vel;      // xyz = V,         w = 1/Mass   (1.0f / Mass, the value Constraint recomputes)
    float4 *angvel;   // xyz = W,         w = 1/MOI
    float4 *ang;      // xyz = Angles,    w = MOI
    float4 *force;    // xyz = Forces
    float4 *torque;   // xyz = Torque
    float4 *scale;    // xyz = model scale, w = Radius (spheres, else 0)
    float4 *verts;    // [n_cubes][6] float4 = 8 packed vec3 (96 B per cube, reference vertex order)
    int32_t *world_id; // [nb] or nullptr
    int32_t *gid;      // [nb] global body id (slab mode)
    // --- statics
    float4 *st_pos;    // xyz = Position, w = 1/Mass
    float4 *st_ang;    // xyz = Angles,   w = 1/MOI
    float4 *st_scale;
    float4 *st_verts;  // [n_statics][6]
    float4 *st_aabb;   // [n_statics][2]
    // --- broadphase
    float4 *aabb_lo, *aabb_hi;          // [nb] by body row
    uint32_t *key[2];                   // Morton cell keys: [0] by body row, [1] in cell order
    uint32_t *val[2];                   // counting sort: [0] the body's cell slot, [1] its rank inside the cell
    uint32_t *cell_count;               // [table + 1] bodies per cell slot -> (scanned) first sorted position of the cell
    uint32_t *pair_fill;                // [nb] pair-slot cursors of the counting pass (by sorted position)
    uint4 *cell_tab;                    // open-addressed table of cells: {key, first sorted position, one past last, -}
    float4 *sbox;                       // [2*nb] Morton-ordered AABB records {lo.xyz,row}{hi.xyz,world}
    uint32_t *pair_tmp;                 // [nb * 24] partner slots filled by the counting pass
    uint32_t cell_mask;                 // table size - 1
    int32_t cell_bits;                  // log2(table size)
    int32_t n_seg;                      // pair segments in use: 2 (CC, CF) for a cube-only world, else 5
    uint32_t *pair_count;               // [5 * nb + 1] per (type, body) candidate counts -> offsets
    int32_t *pair_a, *pair_b;           // [max_pairs] body rows; static k encoded as -(k+1)
    float cell_size;                    // >= largest body AABB extent (fixed at upload)
    // --- narrowphase results per candidate
    int32_t *pair_hit;                  // [max_pairs] 0/1
    uint32_t *pair_hit_scan;            // [max_pairs + 1]
    float4 *pair_out;                   // [max_pairs][3]: PointA, PointB, N
    // --- contacts (reference order)
    int32_t *c_a, *c_b;                 // body rows / -(k+1)
    float4 *c_pa, *c_pb, *c_n;          // [max_contacts]
    // --- solver schedule
    uint32_t *deg;                      // [nb + 1] incidence counts -> offsets
    uint32_t *cursor;                   // [nb]
    int32_t *inc;                       // [2 * max_contacts] contact ids grouped by body
    int32_t *succ_a, *succ_b;           // [max_contacts] the contact's position in body A's / B's contact sequence
    int32_t *run_flag;                  // [max_contacts] 1 where a run of contacts on the same body A starts
    int32_t *run_start;                 // [max_contacts] first contact of every run
    int32_t *trace_level;               // [max_contacts] DAG level per contact (NANS_SOLVER_TRACE only)
    float4 *row_v, *row_w;              // [nb] versioned rows of the solve: (V.xyz, version), (W.xyz, version | level << 20)
    // --- one world over several GPUs (slab.cu); all null / unused otherwise
    const int32_t *live;                // device-side count of live body rows (owned + this step's ghosts); null = nb
    int32_t *sent_mark;                 // [nb] 1 = the row went to the lower neighbour this step (released remotely)
    int32_t *ghost_owner_row;           // [nb] a ghost row's row index on its owner rank
    float4 *peer_row_v, *peer_row_w;    // the UPPER neighbour's row_v / row_w (peer memory over NVLink)
    // --- scan scratch
    uint32_t *scan_block;               // block sums
    Counters *counters;
    float *dt;                          // this step's dt, in device memory: the captured step graph does not depend on it
    int32_t *sticky;                    // sticky error bits (survive the per-step counter reset): 1 = solver schedule stalled
};

// host-side bookkeeping
struct World {
    DeviceWorld d;
    nans_world_desc desc;
    void *arena;
    size_t arena_bytes;
    bool owns_arena;
    cudaStream_t stream;
    bool owns_stream;
    bool have_contacts;
    int solver_mode;            // NANS_SOLVER_EXACT / NANS_SOLVER_SHUFFLED
    void *slab;                 // SlabState (slab.cu) when the world is one rank's share of a larger world
    // whole-step CUDA graph (captured on the 2nd step with an unchanged dt; any change of the launch
    // parameters -- body counts, cell size, world ids -- invalidates it)
    cudaGraphExec_t graph_exec;     // detection (broadphase, narrowphase, contact list)
    cudaGraphExec_t graph_exec_b;   // integrate forces, solve, integrate velocities
    float graph_dt;
    int graph_state;            // 0 none, 1 one eager step seen with graph_dt, 2 captured, -1 disabled
    unsigned graph_launches;    // kernels inside the captured step
    int device;
    int coop_blocks_per_sm;
    int32_t *h_counters;   // pinned mirror of Counters
};

extern thread_local char g_err[512];
extern unsigned long long g_launches;

#define NANS_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            snprintf(nans::g_err, sizeof(nans::g_err), "%s:%d %s: %s", __FILE__, __LINE__,   \
                     #expr, cudaGetErrorString(_e));                                         \
            return NANS_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

#define NANS_LAUNCH_CHECK()                                                                  \
    do { ++nans::g_launches; NANS_CUDA(cudaGetLastError()); } while (0)

// live body rows: everything a kernel of the detection phase may touch (slab mode: owned + received ghosts)
__device__ __forceinline__ int live_nb(const DeviceWorld &w) { return w.live ? *w.live : w.nb; }

inline int div_up(int a, int b) { return (a + b - 1) / b; }
inline int div_up_sz(size_t a, size_t b) { return (int)((a + b - 1) / b); }

// stage launchers (each file owns its kernels)
int launch_set_dt(World *w, float dt);          // every stage reads dt from device memory (DeviceWorld::dt)
int launch_integrate_forces(World *w);
int launch_integrate_velocities(World *w);
int launch_rebuild_statics(World *w);
int launch_models(World *w, float4 *d_out);
int launch_broadphase(World *w);
int launch_narrowphase(World *w);
int launch_contacts(World *w);
int launch_solver(World *w);
int solver_accum_fallbacks(World *w, int32_t *out);
int launch_aabb_only(World *w);
void slab_destroy(World *w);
int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int n, uint32_t *block_scratch, cudaStream_t s);
int exclusive_scan_u32_dn(const uint32_t *in, uint32_t *out, int cap_n, const int32_t *d_n, int extra,
                          uint32_t *block_scratch, cudaStream_t s);
int scan_scratch_elems(int n);

}  // namespace nans
