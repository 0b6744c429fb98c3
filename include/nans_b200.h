/* nans_b200.h — C ABI of libnans_b200.so: the B200 (sm_100a) rigid-body step.
 *
 * This is the thin extern "C" CUDA layer behind the reference's game-layer plugin entry
 *     extern "C" void SimUpdateAndRender(memory*, sdl_input*, sdl_render*, real32 dt)
 * (reference: code/nans.h:396-397, code/nans.cpp:1719).  Each entry point below replaces one
 * reference function or data structure on that path (cited per declaration).  Signatures are
 * plain C: opaque handle, plain pointers and sizes; no torch / C++ types.
 *
 * Memory model: the reference places its whole world (`sdl_state`, code/nans.h:374-386) at
 * offset 0 of one host-owned mmap block (code/sdl_nans.cpp:541-555) and keeps contact lists and
 * GJK/EPA scratch in std::vector.  Here the world lives in ONE device arena (a single cudaMalloc,
 * or caller-provided device memory) carved once into SoA body / shape / pair / contact / scratch
 * regions; nothing is allocated during stepping.
 *
 * Error behaviour: the reference path returns void and reports nothing.  Every call here returns
 * 0 on success or a negative nans_status; nans_last_error() gives the text.  There is NO CPU
 * fallback: without a usable CUDA device every call fails with NANS_ERR_CUDA.
 */
#ifndef NANS_B200_H
#define NANS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    NANS_OK = 0,
    NANS_ERR_CUDA = -1,      /* CUDA runtime error / no device */
    NANS_ERR_ARG = -2,       /* bad argument */
    NANS_ERR_CAPACITY = -3,  /* pair / contact / EPA-arena capacity exceeded (results truncated) */
    NANS_ERR_STATE = -4      /* call order (e.g. solve before detect) */
} nans_status;

/* collision_type, code/nans.h:71-87 */
enum { NANS_CC = 0, NANS_CS = 1, NANS_CF = 2, NANS_SS = 3, NANS_SF = 4 };

typedef struct nans_world nans_world; /* opaque; replaces sdl_state (code/nans.h:374-386) */

/* World capacities; replaces MAX_CUBE_COUNT / MAX_SPHERE_COUNT (code/nans.h:52-53). */
typedef struct {
    int32_t n_cubes;        /* dynamic boxes, body rows [0, n_cubes) */
    int32_t n_spheres;      /* dynamic spheres, body rows [n_cubes, n_cubes+n_spheres) */
    int32_t n_statics;      /* floor-type slabs (reference: exactly one, the Floor) */
    int32_t max_pairs;      /* broadphase candidate capacity (0 = default 24 * bodies) */
    int32_t max_contacts;   /* contact capacity            (0 = default 8 * bodies) */
    int32_t device;         /* CUDA device ordinal */
    void *arena;            /* optional caller-provided device arena (e.g. a torch tensor) */
    uint64_t arena_bytes;   /* its size; ignored when arena == NULL */
    void *stream;           /* cudaStream_t to run on (NULL = library-owned stream) */
} nans_world_desc;

/* Host-side view of the world state: the same fields as struct cube / sphere
 * (code/nans.h:303-336), as arrays.  Any pointer may be NULL (field skipped).
 * Rows: cubes first, then spheres ("nb" = n_cubes + n_spheres). */
typedef struct {
    float *pos, *vel, *force;       /* [nb][3]  Position, V, Forces */
    float *ang, *angvel, *torque;   /* [nb][3]  Angles, W, Torque */
    float *mass, *moi;              /* [nb]     Mass, MOI */
    float *scale;                   /* [nb][3]  model scale (Size, or (.5,1,.5)*Size for the debug box) */
    float *radius;                  /* [nb]     Radius (spheres) */
    float *verts;                   /* [n_cubes][8][3] Vertices (world space, reference order) */
    float *st_pos, *st_ang, *st_scale; /* [n_statics][3] */
    float *st_mass, *st_moi;        /* [n_statics] */
    float *st_verts;                /* [n_statics][8][3] */
    int32_t *world_id;              /* [nb] optional independent-world id (batched worlds); NULL = one world */
} nans_scene_view;

/* The meaningful fields of contact_pair (code/nans.h:339-372), reference index conventions:
 * a/b index cubes for CC; cube,sphere for CS; cube,static for CF; spheres for SS; sphere,static
 * for SF. */
typedef struct {
    int32_t type, a, b;
    float point_a[3], point_b[3], n[3];
} nans_contact;

/* Per-step counters (device -> host on request). */
typedef struct {
    int32_t n_pairs;          /* broadphase candidates */
    int32_t n_contacts;       /* contacts after GJK+EPA (reference-order list length) */
    int32_t n_gjk_found;      /* pairs whose GJK reported FoundIntersection */
    int32_t solver_levels;    /* dependency levels the exact-order solver ran */
    int32_t overflow;         /* bit0 pairs, bit1 contacts, bit2 EPA faces, bit3 EPA edges, bit4 cell list */
    int32_t max_epa_faces;    /* high-water mark of the per-thread polytope arena */
    int32_t accum_fallbacks;  /* contacts of the last solve whose 70-iteration accumulation took the literal loop */
    int32_t reserved[1];
} nans_step_stats;

/* ---- lifecycle ------------------------------------------------------------------------- */
/* replaces the host's mmap + `sdl_state at PermanentStorage+0` (code/sdl_nans.cpp:541-555,
 * code/nans.cpp:1721); desc->arena NULL => one cudaMalloc. */
int nans_world_create(const nans_world_desc *desc, nans_world **out);
void nans_world_destroy(nans_world *w);
uint64_t nans_world_arena_bytes(const nans_world_desc *desc); /* bytes create() will carve */
const char *nans_last_error(void);
int nans_device_count(void);

/* ---- state transfer (host buffers; copies are on the world's stream, synchronous on return) */
/* replaces Init's field-by-field scene set-up (code/nans.cpp:1598-1678) */
int nans_world_upload(nans_world *w, const nans_scene_view *scene);
int nans_world_download(nans_world *w, nans_scene_view *scene);
/* CubeAddForce/CubeAddTorque/SphereAddForce (code/nans.cpp:81-110): += on one body row */
int nans_world_add_force(nans_world *w, int32_t body_row, const float force[3], const float torque[3]);
/* overwrite a body's pose/velocities (ShootSphere teleport, code/nans.cpp:113-120; debug pokes :149-186) */
int nans_world_set_body(nans_world *w, int32_t body_row, const float pos[3], const float vel[3],
                        const float angvel[3]);

/* Pipelined host I/O for hosts that can take the poses one frame late (a renderer): the copies run on
 * their own streams and overlap the step.  Host buffers should be pinned and must stay valid until the
 * copy completes (upload: until the next nans_step has been issued and nans_world_wait(w, -1) or a later
 * download ticket returned; download: until nans_world_wait on its ticket).  Fields: pos, vel, force, ang,
 * angvel, torque.  The reference's frame is synchronous (code/sdl_nans.cpp:986); use the calls above for that.
 *
 * An upload that carries ONLY force and/or torque is additionally deferred: nans_step runs collision detection
 * first (it reads neither forces nor velocities, so it commutes with integrate-forces bit for bit) and unpacks
 * the uploaded rows just before integrate-forces, so the copy overlaps broadphase + narrowphase even when the
 * same frame's poses are then read back synchronously (upload_async -> step -> download: the loop bench.py
 * times as e2e).  Every other entry point sees the upload as already applied. */
int nans_world_upload_async(nans_world *w, const nans_scene_view *scene);
int nans_world_download_async(nans_world *w, nans_scene_view *scene, int32_t *ticket);
int nans_world_wait(nans_world *w, int32_t ticket);   /* ticket < 0: everything outstanding */

/* device-to-device snapshot / restore of the dynamic state (pose, velocities, forces, vertices);
 * asynchronous on the world's stream.  The reference checkpoints implicitly: its whole world is one
 * host block that survives plugin reloads (code/sdl_nans.cpp:541-555). */
int nans_world_snapshot(nans_world *w);
int nans_world_restore(nans_world *w);

/* ---- ONE world over several GPUs of a box (config C5): spatial slabs, one process (or thread) per GPU.
 * No reference counterpart (the reference is one process capped at 16 cubes, code/nans.h:52-53); the decomposed
 * world is bit-identical to the same world on one GPU.  Bodies are numbered slab-major: rank r owns the global index
 * range [gid_base, gid_base + n_owned), laid out by the scene as a slab along x; rows [n_owned, n_owned + halo_cap)
 * of its world receive, every step, copies of the upper neighbour's bodies that reach into its bounding box.
 * Exchange: NCCL (all-gather of the slab boxes, fixed-capacity halo send/recv, all on the world's stream, no host
 * round trip), and the solve runs as ONE dataflow across the GPUs through peer stores over NVLink (csrc/slab.cu,
 * csrc/solver.cu).  Every rank must create its world with the SAME capacities (cube-only, n_cubes >= n_owned +
 * halo_cap, library-owned arena).  Set-up, with the two blobs carried between the ranks by the host
 * (torch.distributed, MPI, a file: INTEGRATION.md):
 *   rank 0: nans_slab_unique_id(id)            -> broadcast id (128 bytes)
 *   all:    nans_slab_init(w, rank, n, id, ...) ; nans_slab_ipc_handle(w, blob) -> all-gather the blobs (128 bytes each)
 *   all:    nans_slab_connect(w, blobs)        ; then nans_slab_step(w, dt) per frame. */
int nans_slab_unique_id(void *out128);
int nans_slab_init(nans_world *w, int32_t rank, int32_t nranks, const void *unique_id128, int32_t n_owned,
                   int32_t gid_base, int32_t halo_cap);
int nans_slab_ipc_handle(nans_world *w, void *out128);
int nans_slab_connect(nans_world *w, const void *blobs /* [nranks][128] */);
int nans_slab_step(nans_world *w, float dt);   /* the step of code/nans.cpp:1758-1762 on this rank's share; asynchronous */
/* synchronises; err_bits (sticky): 1 = halo larger than halo_cap, 2 = a body reaches into a rank other than the lower
 * neighbour (either makes the call return NANS_ERR_STATE: the world is no longer exact); live_rows = owned + ghosts */
int nans_slab_status(nans_world *w, int32_t *err_bits, int32_t *live_rows, int64_t *halo_bytes_per_message);
int nans_slab_row_gids(nans_world *w, int32_t *out, int32_t cap, int32_t *count);   /* global ids of the live rows */

/* Sweep order of SolveConstraints (code/nans.cpp:1539-1548).  EXACT (default): the reference's list order, results
 * bit-identical to it.  SHUFFLED: the same single Gauss-Seidel pass in a fixed pseudo-random order, whose dependency
 * graph is ~10 levels deep instead of ~100 (a greedy graph colouring in random order): faster, deterministic,
 * atomics-free, but NOT the reference's results wherever contacts share bodies (DESIGN.md states the measured
 * deviation).  Not available for slab-partitioned worlds. */
enum { NANS_SOLVER_EXACT = 0, NANS_SOLVER_SHUFFLED = 1 };
int nans_world_set_solver(nans_world *w, int32_t mode);

/* ---- the four stages, one entry per reference function (asynchronous on the world's stream) */
int nans_integrate_forces(nans_world *w, float dt);     /* IntegrateForces     code/nans.cpp:975  */
int nans_detect_collisions(nans_world *w);              /* DetectCollisions    code/nans.cpp:1352 */
int nans_solve_constraints(nans_world *w, float dt);    /* SolveConstraints    code/nans.cpp:1539 */
int nans_integrate_velocities(nans_world *w, float dt); /* IntegrateVelocities code/nans.cpp:1332
                                                           + model/vertex rebuild :1870-1881,1913-1941 */
/* the draw section's model rebuild on its own (code/nans.cpp:1870-1881,1913-1941 + UpdateVertices
 * :395-407): Model = T*Rx*Ry*Rz*S -> 8 world vertices for every cube and static, from the pose */
int nans_rebuild_vertices(nans_world *w);
/* the draw section's per-body "Model" uniform (code/nans.cpp:1870-1881 floor, :1913-1941 cubes, :1971-1990 spheres)
 * for every body at once, as instanced draw data: [(n_cubes + n_spheres + n_statics)][16] floats, column-major like
 * glm::mat4 (bit-identical to the reference's matrices).  d_out: a device buffer (e.g. a CUDA-GL interop VBO bound
 * as a per-instance mat4 attribute), h_out: a host buffer; either may be NULL. */
int nans_world_models(nans_world *w, void *d_out, float *h_out);
/* the step as SimUpdateAndRender runs it (code/nans.cpp:1758-1762).  Queued as DetectCollisions, then
 * IntegrateForces, SolveConstraints, IntegrateVelocities: detection reads positions/vertices only and
 * IntegrateForces writes velocities and clears forces only, so the result is bit-identical to the reference's
 * order and a pending force/torque upload overlaps detection (see nans_world_upload_async). */
int nans_step(nans_world *w, float dt);
int nans_synchronize(nans_world *w);
/* nans_step with a CUDA event between the stages, recorded on the world's stream (measurement only):
 * stage_ms[8] = integrate_forces, broadphase, narrowphase, contact compaction, solver,
 * integrate_velocities+rebuild, whole step, 0.  Synchronises. */
int nans_step_profiled(nans_world *w, float dt, float *stage_ms);

/* results of the last detect/step */
int nans_get_stats(nans_world *w, nans_step_stats *out);                 /* synchronises */
int nans_get_contacts(nans_world *w, nans_contact *out, int32_t cap, int32_t *count); /* Pairs vector, code/nans.h:385 */
int nans_get_pairs(nans_world *w, int32_t *pair_a, int32_t *pair_b, int32_t cap, int32_t *count); /* body rows; statics as -(k+1) */
/* replace the contact list (tests: solve from the oracle's contacts) */
int nans_set_contacts(nans_world *w, const nans_contact *in, int32_t count);

/* ---- stand-alone narrowphase, CheckCollision (code/nans.cpp:907-966) over n pairs (config C3)
 * Inputs are world-space shapes: box = 8 vertices, sphere = centre + radius.  Host buffers;
 * outputs: hit (bool32 result), gjk (evolve_result), N / PointA / PointB. */
int nans_check_collision_batch(int32_t n, const int32_t *type,
                               const float *pos_a, const float *verts_a, const float *rad_a,
                               const float *pos_b, const float *verts_b, const float *rad_b,
                               int32_t *hit, int32_t *gjk, float *out_n, float *out_pa, float *out_pb,
                               int32_t device);

/* device-resident variant for throughput measurement: buffers are DEVICE pointers
 * (pos/rad: [n][4] float4 = xyz+radius; verts: [n][8][3]; out_contact: [n][3] float4) */
int nans_check_collision_device(int32_t n, const int32_t *d_type,
                                const float *d_posrad_a, const float *d_verts_a,
                                const float *d_posrad_b, const float *d_verts_b,
                                int32_t *d_hit, float *d_out, void *stream);

/* EPA-arena overflow bits of the nans_check_collision_device calls on the current device since the last query
 * (NANS_ERR_CAPACITY if any: an overflowing pair is reported as a miss); synchronises the device */
int nans_check_collision_device_status(int32_t *overflow_bits);

/* debug aid: per-contact start time (ns) and dependency level of the last solve (NANS_SOLVER_TRACE=1) */
int nans_debug_solver_trace(nans_world *w, uint64_t *times, int32_t *levels, int32_t cap);

/* debug aid: the library's exclusive prefix sum over n uint32 (csrc/scan.cu; every offset table of the step comes out of
 * it), host buffers in and out, on the current device: out[i] = in[0] + ... + in[i-1].  Lets the tests drive
 * it at any length. */
int nans_debug_scan(const uint32_t *in, uint32_t *out, int32_t n);

/* kernel-launch counter (all launches issued by this library in this process) */
uint64_t nans_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif
