// nans_plugin.cpp — the hot-reloadable game-layer plugin: a new `nans.so` exporting
// SimUpdateAndRender with the reference's ABI (include/nans_plugin.h), whose physics runs on the
// GPU through the thin extern "C" layer of libnans_b200.so (include/nans_b200.h).
//
// What it mirrors from the reference's game layer (code/nans.cpp):
//   Init :1551-1717 (camera + the demo scene), HandleInput :123-190, UpdateCamera :20-38,
//   ShootSphere :113-120, the frame order of SimUpdateAndRender :1719-1775, and the per-frame
//   model rebuild of the draw section (:1870-1881,1913-1941 — done on the device, fused into the
//   last kernel of the step).  The GL drawing (:1779-2025) is out of scope (renderer optional).
//
// Hot reload (code/sdl_nans.cpp:922-930): nothing that must survive dlclose/dlopen lives in plugin
// statics.  All state — camera, host mirror of the bodies, and the handle of the device world —
// is inside Memory->PermanentStorage; libnans_b200.so (which owns the CUDA context and the device
// arena) is pinned with RTLD_NODELETE so unloading this plugin never tears CUDA down.
//
// Host-side camera math follows glm's operation order in plain fp32 (build with
// -ffp-contract=off): Camera.Front feeds ShootSphere's force, i.e. the physics.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/nans_b200.h"
#include "../../include/nans_plugin.h"

namespace {

struct v3 { float x, y, z; };
inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
inline v3 operator+(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline v3 operator-(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline v3 operator*(float s, v3 a) { return V3(s * a.x, s * a.y, s * a.z); }
inline v3 operator*(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline float dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline v3 cross(v3 x, v3 y) { return V3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
inline v3 normalize(v3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }

// sdl_camera, code/nans.h:279-293
struct camera {
    float FOV, Pitch, Yaw, Speed;
    v3 Position, Target, Direction, Up, Front, Right;
};

const uint64_t kMagic = 0x4e414e5342323030ull;  // "NANSB200"

struct plugin_state {
    uint64_t magic;
    nans_world *world;          // device world (lives in libnans_b200.so, survives reloads)
    int32_t n_cubes, n_spheres, n_statics;
    int32_t frame, n_contacts;
    camera cam;
    size_t arena_used;          // bump allocator over PermanentStorage (cf. code/utilities.cpp:25-49)
    float *pos, *ang, *vel, *angvel;   // host mirror, [nb][3] each, inside PermanentStorage
};

void *push(memory *m, plugin_state *s, size_t bytes)
{
    s->arena_used = (s->arena_used + 63) & ~(size_t)63;
    if (s->arena_used + bytes > m->PermanentStorageSize) {
        fprintf(stderr, "nans plugin: PermanentStorage exhausted\n");
        abort();
    }
    void *p = (char *)m->PermanentStorage + s->arena_used;
    s->arena_used += bytes;
    return p;
}

void die(const char *what, int rc)
{
    fprintf(stderr, "nans plugin: %s failed (%d): %s\n", what, rc, nans_last_error());
    abort();   // the reference's boundary has no error channel; there is no CPU fallback
}
#define CK(call) do { int _rc = (call); if (_rc) die(#call, _rc); } while (0)

// ---- scenes ------------------------------------------------------------------------------
struct scene_buf {
    int nc, ns, nst;
    float *pos, *ang, *scale, *mass, *moi, *radius, *st_pos, *st_scale, *st_mass, *st_moi;
};

float cube_moi(float mass, float size) { return (mass / 12.0f) * (2.0f * size * size); }

uint64_t g_rng = 0x9E3779B97F4A7C15ull;
float urand() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return (float)((g_rng >> 40) * (1.0 / 16777216.0)); }

// Init's scene, code/nans.cpp:1598-1678
void demo_scene(scene_buf &b)
{
    const float cube_pos[4][3] = {{2.0f, 3.5f, 2.0f}, {2.0f, 1.0f, 2.0f}, {2.0f, 4.5f, 2.0f}, {1.0f, 1.0f, 1.0f}};
    for (int i = 0; i < 4; ++i) {
        memcpy(b.pos + 3 * i, cube_pos[i], 12);
        b.mass[i] = 1.0f;
        b.moi[i] = cube_moi(1.0f, 1.0f);
        b.scale[3 * i] = b.scale[3 * i + 1] = b.scale[3 * i + 2] = 1.0f;
    }
    b.scale[9] = 0.5f; b.scale[10] = 1.0f; b.scale[11] = 0.5f;   // Cubes[3]: the debug box (:1930-1941)
    const float sp[3] = {0.1f, 1.1f, 1.1f};
    memcpy(b.pos + 12, sp, 12);
    b.radius[4] = 0.25f; b.mass[4] = 2.0f;
    b.moi[4] = ((2.0f / 5.0f) * 2.0f) * (0.25f * 0.25f);
    b.scale[12] = b.scale[13] = b.scale[14] = 0.25f;
    const float fp[3] = {1.2f, -0.5f, 1.0f}, fs[3] = {100.0f, 1.0f, 100.0f};
    memcpy(b.st_pos, fp, 12); memcpy(b.st_scale, fs, 12);
    b.st_mass[0] = 100000.0f;
    b.st_moi[0] = (100000.0f / 12.0f) * (2.0f * 100.0f * 100.0f);
}

// jittered lattice of unit cubes over one floor (configs C2/C5 through the plugin boundary)
void pile_scene(scene_buf &b, int side, float spacing, float jitter)
{
    for (int i = 0; i < b.nc; ++i) {
        const int y = i / (side * side), rem = i % (side * side), z = rem / side, x = rem % side;
        b.pos[3 * i] = 0.5f + x * spacing + (2 * urand() - 1) * jitter;
        b.pos[3 * i + 1] = 0.52f + y * spacing + (2 * urand() - 1) * jitter;
        b.pos[3 * i + 2] = 0.5f + z * spacing + (2 * urand() - 1) * jitter;
        b.mass[i] = 1.0f; b.moi[i] = cube_moi(1.0f, 1.0f);
        b.scale[3 * i] = b.scale[3 * i + 1] = b.scale[3 * i + 2] = 1.0f;
    }
    const float ext = side * spacing;
    float size = 64.0f;
    while (size < ext + 16.0f) size *= 2.0f;
    b.st_pos[0] = ext / 2 + 0.13f; b.st_pos[1] = -0.5f; b.st_pos[2] = ext / 2 + 0.07f;
    b.st_scale[0] = size; b.st_scale[1] = 1.0f; b.st_scale[2] = size;
    b.st_mass[0] = 100000.0f;
    b.st_moi[0] = (100000.0f / 12.0f) * (2.0f * size * size);
}

// Init, code/nans.cpp:1551-1717
void Init(memory *Memory, plugin_state *S)
{
    memset(S, 0, sizeof(*S));
    S->magic = kMagic;
    S->arena_used = sizeof(plugin_state);
    // the CUDA layer must outlive this plugin image (hot reload): pin it
    if (!dlopen("libnans_b200.so", RTLD_NOW | RTLD_GLOBAL | RTLD_NODELETE))
        fprintf(stderr, "nans plugin: note: could not pin libnans_b200.so (%s)\n", dlerror());

    camera &c = S->cam;
    c.FOV = 45.0f; c.Pitch = 0.0f; c.Yaw = -90.0f; c.Speed = 0.05f;
    c.Position = V3(0.0f, 0.0f, 3.0f);
    const v3 Up = V3(0.0f, 1.0f, 0.0f);
    c.Target = V3(0.0f, 0.0f, 0.0f);
    c.Direction = normalize(c.Position - c.Target);
    c.Front = V3(0.0f, 0.0f, -1.0f);
    c.Right = normalize(cross(Up, c.Direction));
    c.Up = cross(c.Direction, c.Right);

    // scene selection: the reference has exactly one (Init's); larger ones reuse the same boundary
    const char *sel = getenv("NANS_SCENE");
    int side = 0, layers = 0;
    if (sel && sscanf(sel, "pile:%d:%d", &side, &layers) == 2 && side > 0 && layers > 0) {
        S->n_cubes = side * side * layers; S->n_spheres = 0; S->n_statics = 1;
    } else {
        S->n_cubes = 4; S->n_spheres = 1; S->n_statics = 1; side = 0;
    }
    const int nb = S->n_cubes + S->n_spheres;
    S->pos = (float *)push(Memory, S, sizeof(float) * 3 * nb);
    S->ang = (float *)push(Memory, S, sizeof(float) * 3 * nb);
    S->vel = (float *)push(Memory, S, sizeof(float) * 3 * nb);
    S->angvel = (float *)push(Memory, S, sizeof(float) * 3 * nb);

    // scene scratch comes from TransientStorage (host-owned, zero-filled mmap)
    float *t = (float *)Memory->TransientStorage;
    scene_buf b;
    b.nc = S->n_cubes; b.ns = S->n_spheres; b.nst = S->n_statics;
    b.pos = t; t += 3 * nb; b.ang = t; t += 3 * nb; b.scale = t; t += 3 * nb;
    b.mass = t; t += nb; b.moi = t; t += nb; b.radius = t; t += nb;
    b.st_pos = t; t += 3; b.st_scale = t; t += 3; b.st_mass = t; t += 1; b.st_moi = t; t += 1;
    float *zeros = t; t += 3 * nb;
    float *verts = t; t += 24 * (size_t)S->n_cubes;
    float *st_verts = t; t += 24;
    float *st_ang = t; t += 3;
    if ((uint64_t)((char *)t - (char *)Memory->TransientStorage) > Memory->TransientStorageSize)
    {
        fprintf(stderr, "nans plugin: scene scratch does not fit the host's TransientStorage block "
                        "(reduce NANS_SCENE or grow the block)\n");
        abort();
    }
    memset(Memory->TransientStorage, 0, (char *)t - (char *)Memory->TransientStorage);
    if (side) pile_scene(b, side, 1.02f, 0.005f); else demo_scene(b);

    nans_world_desc d;
    memset(&d, 0, sizeof(d));
    d.n_cubes = S->n_cubes; d.n_spheres = S->n_spheres; d.n_statics = S->n_statics;
    const char *dev = getenv("NANS_DEVICE");
    d.device = dev ? atoi(dev) : 0;
    CK(nans_world_create(&d, &S->world));
    // sweep order of SolveConstraints: the reference's (default) or the faster shuffled one (NOT the reference's
    // results; include/nans_b200.h).  Hot reload is the reference's config mechanism; an env var is ours.
    const char *sv = getenv("NANS_SOLVER");
    if (sv && !strcmp(sv, "shuffled")) CK(nans_world_set_solver(S->world, NANS_SOLVER_SHUFFLED));

    nans_scene_view v;
    memset(&v, 0, sizeof(v));
    v.pos = b.pos; v.ang = b.ang; v.scale = b.scale; v.mass = b.mass; v.moi = b.moi; v.radius = b.radius;
    v.vel = zeros; v.angvel = zeros; v.force = zeros; v.torque = zeros;
    v.st_pos = b.st_pos; v.st_ang = st_ang; v.st_scale = b.st_scale; v.st_mass = b.st_mass; v.st_moi = b.st_moi;
    if (!side) {
        // frame-0 state of the reference: every Model is identity, so every cube's and the floor's
        // collision vertices are the unit cube at the origin (code/nans.cpp:1599-1600,1667-1668)
        static const float corner[24] = {.5f, .5f, .5f, .5f, .5f, -.5f, -.5f, .5f, .5f, -.5f, .5f, -.5f,
                                         .5f, -.5f, .5f, .5f, -.5f, -.5f, -.5f, -.5f, .5f, -.5f, -.5f, -.5f};
        for (int i = 0; i < S->n_cubes; ++i) memcpy(verts + 24 * i, corner, 96);
        memcpy(st_verts, corner, 96);
        v.verts = verts; v.st_verts = st_verts;
    }
    CK(nans_world_upload(S->world, &v));
    if (side) CK(nans_rebuild_vertices(S->world));   // large scenes start with consistent vertices
    memcpy(S->pos, b.pos, sizeof(float) * 3 * nb);
}

// ShootSphere, code/nans.cpp:113-120
void ShootSphere(plugin_state *S)
{
    if (S->n_spheres < 1) return;
    const int row = S->n_cubes;     // Spheres[0]
    const float zero[3] = {0, 0, 0};
    const camera &c = S->cam;
    const float p[3] = {c.Position.x, c.Position.y, c.Position.z};
    CK(nans_world_set_body(S->world, row, p, zero, zero));
    const v3 f = 4001.0f * c.Front;    // SHOOT_FORCE, code/nans.h:48 (forces are zero here: cleared every frame)
    const float ff[3] = {f.x, f.y, f.z};
    CK(nans_world_add_force(S->world, row, ff, zero));
    memcpy(S->pos + 3 * row, p, 12);
}

// HandleInput, code/nans.cpp:123-190
void HandleInput(plugin_state *S, sdl_input *In, float dt)
{
    camera &c = S->cam;
    auto &K = In->KeyboardController;
    if (K.MoveForward.EndedDown) c.Position = c.Position + c.Speed * c.Front;
    if (K.MoveLeft.EndedDown) c.Position = c.Position - normalize(cross(c.Front, c.Up)) * c.Speed;
    if (K.MoveBack.EndedDown) c.Position = c.Position - c.Speed * c.Front;
    if (K.MoveRight.EndedDown) c.Position = c.Position + normalize(cross(c.Front, c.Up)) * c.Speed;
    if (K.ShootAction.EndedDown) ShootSphere(S);
    if (S->n_cubes > 3) {   // debug pokes on Cubes[3] (:149-172)
        v3 p = V3(S->pos[9], S->pos[10], S->pos[11]);
        bool moved = false;
        auto poke = [&](bool down, v3 d) { if (down) { p = p + dt * d; moved = true; } };
        poke(K.DebugLeft.EndedDown, V3(1.0f, 0.0f, 0.0f));
        poke(K.DebugRight.EndedDown, V3(-1.0f, 0.0f, 0.0f));
        poke(K.DebugUp.EndedDown, V3(0.0f, 1.0f, 0.0f));
        poke(K.DebugDown.EndedDown, V3(0.0f, -1.0f, 0.0f));
        poke(K.DebugForward.EndedDown, V3(0.0f, 0.0f, 1.0f));
        poke(K.DebugBack.EndedDown, V3(0.0f, 0.0f, -1.0f));
        if (moved) {
            const float q[3] = {p.x, p.y, p.z};
            CK(nans_world_set_body(S->world, 3, q, nullptr, nullptr));
            memcpy(S->pos + 9, q, 12);
        }
        if (K.DebugReset.EndedDown) {   // :174-189
            const float rp[4][3] = {{2.0f, 3.5f, 2.0f}, {2.0f, 1.0f, 2.0f}, {2.0f, 4.5f, 2.0f}, {0.0f, 0.0f, 0.0f}};
            const float zero[3] = {0, 0, 0};
            for (int i = 0; i < 4; ++i) {
                CK(nans_world_set_body(S->world, i, rp[i], zero, zero));
                memcpy(S->pos + 3 * i, rp[i], 12);
            }
        }
    }
}

// UpdateCamera, code/nans.cpp:20-38
void UpdateCamera(plugin_state *S, sdl_input *In)
{
    camera &c = S->cam;
    c.Yaw += In->MouseController.XRel * In->MouseController.Sensitivity;
    c.Pitch += -In->MouseController.YRel * In->MouseController.Sensitivity;
    In->MouseController.XRel = 0;
    In->MouseController.YRel = 0;
    if (c.Pitch > 89.0f) c.Pitch = 89.0f;
    if (c.Pitch < -89.0f) c.Pitch = -89.0f;
    if (c.Position.y < 0.5) c.Position.y = 0.5;
}

// glm::perspective (RH, -1..1) and glm::lookAt (RH), column-major
void perspective(float *m, float fovy, float aspect, float zn, float zf)
{
    const float t = tanf(fovy / 2.0f);
    memset(m, 0, 64);
    m[0] = 1.0f / (aspect * t);
    m[5] = 1.0f / t;
    m[10] = -(zf + zn) / (zf - zn);
    m[11] = -1.0f;
    m[14] = -(2.0f * zf * zn) / (zf - zn);
}
void look_at(float *m, v3 eye, v3 center, v3 up)
{
    const v3 f = normalize(center - eye);
    const v3 s = normalize(cross(f, up));
    const v3 u = cross(s, f);
    memset(m, 0, 64);
    m[0] = s.x; m[4] = s.y; m[8] = s.z;
    m[1] = u.x; m[5] = u.y; m[9] = u.z;
    m[2] = -f.x; m[6] = -f.y; m[10] = -f.z;
    m[12] = -dot(s, eye); m[13] = -dot(u, eye); m[14] = dot(f, eye);
    m[15] = 1.0f;
}

}  // namespace

extern "C" SIM_UPDATE_AND_RENDER(SimUpdateAndRender)
{
    plugin_state *S = (plugin_state *)Memory->PermanentStorage;
    if (!Memory->IsInitialized) {      // lazy init keyed on the host's flag (code/nans.cpp:1725-1729)
        Init(Memory, S);
        Memory->IsInitialized = 1;
    }
    camera &c = S->cam;
    // coordinate systems (:1735-1744)
    perspective(Render->Projection, radians(45.0f), 1920.0f / 1080.0f, 0.1f, 100.0f);
    look_at(Render->View, c.Position, c.Position + c.Front, c.Up);
    // input (:1751-1752)
    HandleInput(S, Input, dt);
    UpdateCamera(S, Input);
    // physics (:1758-1762) — IntegrateForces, DetectCollisions, SolveConstraints, IntegrateVelocities
    // (+ the draw section's model/vertex rebuild), all on the device
    CK(nans_step(S->world, dt));
    // poses back for the renderer / game layer
    nans_scene_view v;
    memset(&v, 0, sizeof(v));
    v.pos = S->pos; v.ang = S->ang;
    if (!getenv("NANS_NO_VELOCITY_READBACK")) { v.vel = S->vel; v.angvel = S->angvel; }
    CK(nans_world_download(S->world, &v));
    nans_step_stats st;
    int rc = nans_get_stats(S->world, &st);
    if (rc) die("nans_get_stats", rc);
    S->n_contacts = st.n_contacts;
    // camera front (:1768-1775)
    v3 Front;
    Front.x = cosf(radians(c.Yaw)) * cosf(radians(c.Pitch));
    Front.y = sinf(radians(c.Pitch));
    Front.z = sinf(radians(c.Yaw)) * cosf(radians(c.Pitch));
    c.Front = normalize(Front);
    look_at(Render->View, c.Position, c.Position + c.Front, c.Up);
    S->frame++;
}

extern "C" int NansPluginPeek(const memory *Memory, nans_plugin_view *out)
{
    const plugin_state *S = (const plugin_state *)Memory->PermanentStorage;
    if (!Memory->IsInitialized || S->magic != kMagic) return -1;
    out->n_cubes = S->n_cubes; out->n_spheres = S->n_spheres; out->n_contacts = S->n_contacts;
    out->frame = S->frame;
    out->pos = S->pos; out->ang = S->ang; out->vel = S->vel; out->angvel = S->angvel;
    const camera &c = S->cam;
    out->cam_pos[0] = c.Position.x; out->cam_pos[1] = c.Position.y; out->cam_pos[2] = c.Position.z;
    out->cam_front[0] = c.Front.x; out->cam_front[1] = c.Front.y; out->cam_front[2] = c.Front.z;
    out->cam_yaw = c.Yaw; out->cam_pitch = c.Pitch;
    return S->n_cubes + S->n_spheres;
}
