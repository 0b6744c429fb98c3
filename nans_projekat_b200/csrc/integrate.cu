// integrate.cu — RK4 velocity integration and position/vertex update.
//
//   integrate_forces_kernel      <- IntegrateForces      (reference code/nans.cpp:975-1018, RK4 :51-78)
//   integrate_velocities_kernel  <- IntegrateVelocities  (code/nans.cpp:1332-1349) fused with the model
//                                   rebuild of the draw section (:1870-1881,1913-1941) and
//                                   UpdateVertices (:395-407): Model = T*Rx*Ry*Rz*S -> 8 vertices.
//
// Both are one thread per body over float4 SoA rows: fully coalesced 16 B loads/stores, no reuse,
// HBM-bound (algorithmic bytes: 104 B/body and 168 B/body).  In the reference the
// vertices used by collision detection in frame k are those of the pose at the end of frame k-1
// (SURVEY.md Appendix B); here the vertex array is persistent state rebuilt once, at the end of
// the step, which is the same thing.
#include "integrate.cuh"
#include "world.cuh"

namespace nans {

__global__ void __launch_bounds__(256) integrate_forces_kernel(DeviceWorld w)
{
    const float dt = *w.dt;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n_owned) return;      // == nb unless the world is one rank's slab (ghost rows belong to their owner)
    const float4 p = w.pos[i];      // w = Mass
    float4 v = w.vel[i];            // w = 1/Mass
    float4 a = w.angvel[i];         // w = 1/MOI
    const float4 f = w.force[i];
    const float4 t = w.torque[i];
    const vec3 nv = rk4<true>(dt, V3(v), V3(f), p.w, v.w);
    const vec3 na = rk4<false>(dt, V3(a), V3(t), 0.0f, a.w);
    w.vel[i] = make_float4(nv.x, nv.y, nv.z, v.w);
    w.angvel[i] = make_float4(na.x, na.y, na.z, a.w);
    // CubeClearForces / SphereClearForces, code/nans.cpp:95-110
    w.force[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    w.torque[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ void store_verts(float4 *dst, const float v[24])
{
#pragma unroll
    for (int q = 0; q < 6; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// The 8 vertices of a cube are 96 contiguous bytes (AoS: the narrowphase gathers whole boxes).  A thread storing its
// own six float4 makes every store instruction of the warp touch 32 different sectors, half used (ncu: drain /
// lg_throttle / mio_throttle 40 % of the stall samples, 48 % of the HBM peak).  The block's 128 boxes are one contiguous
// 12 KB range, so they go through shared memory (stride 25: conflict-free writes) and leave as 768 consecutive float4.
constexpr int kIvThreads = 128, kIvStride = 25;

__global__ void __launch_bounds__(kIvThreads) integrate_velocities_kernel(DeviceWorld w)
{
    __shared__ float sv[kIvThreads * kIvStride];
    const float dt = *w.dt;
    const int b0 = blockIdx.x * kIvThreads;
    const int i = b0 + threadIdx.x;
    if (i < w.n_owned) {
        float4 p = w.pos[i];
        float4 a = w.ang[i];
        const float4 v = w.vel[i];
        const float4 av = w.angvel[i];
        const vec3 np = V3(p) + dt * V3(v);     // Position += dt * V
        const vec3 na = V3(a) + dt * V3(av);    // Angles   += dt * W
        w.pos[i] = make_float4(np.x, np.y, np.z, p.w);
        w.ang[i] = make_float4(na.x, na.y, na.z, a.w);
        if (i < w.n_cubes) {
            float verts[24];
            model_vertices(np, na, V3(w.scale[i]), verts);
#pragma unroll
            for (int q = 0; q < 24; ++q) sv[threadIdx.x * kIvStride + q] = verts[q];
        }
    }
    __syncthreads();
    const int n_box = min(min(w.n_owned, w.n_cubes) - b0, kIvThreads);      // boxes this block rebuilt (<= 0: none)
    float4 *dst = w.verts + 6 * (size_t)b0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int q4 = threadIdx.x + kIvThreads * k;       // float4 number q4 of the block's range: box q4 / 6, part q4 % 6
        const int j = q4 / 6, part = q4 - 6 * j;
        if (j < n_box) {
            const float *src = sv + j * kIvStride + 4 * part;
            dst[q4] = make_float4(src[0], src[1], src[2], src[3]);
        }
    }
}

// statics (the Floor): rebuilt once at upload — FloorUpdateVertices, code/nans.cpp:380-392,1870-1881
__global__ void rebuild_statics_kernel(DeviceWorld w)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= w.n_statics) return;
    float verts[24];
    model_vertices(V3(w.st_pos[k]), V3(w.st_ang[k]), V3(w.st_scale[k]), verts);
    store_verts(w.st_verts + 6 * k, verts);
}

// Instanced draw data: the Model matrix the reference's draw section builds per cube / sphere / floor
// and uploads as the "Model" uniform (code/nans.cpp:1870-1881, 1913-1941, 1971-1990), here for EVERY body in one
// pass, column-major like glm::mat4, so a renderer can bind `out` as a per-instance mat4 attribute buffer.
// rows: cubes, spheres (scale = Radius), then the statics.
__global__ void __launch_bounds__(128) models_kernel(DeviceWorld w, float4 *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.nb + w.n_statics) return;
    vec3 pos, ang, scale;
    if (i < w.nb) {
        pos = V3(w.pos[i]); ang = V3(w.ang[i]);
        const float4 sc = w.scale[i];
        scale = i < w.n_cubes ? V3(sc) : V3(sc.w, sc.w, sc.w);
    } else {
        const int k = i - w.nb;
        pos = V3(w.st_pos[k]); ang = V3(w.st_ang[k]); scale = V3(w.st_scale[k]);
    }
    mat4 m;
    model_matrix(pos, ang, scale, m);
#pragma unroll
    for (int c = 0; c < 4; ++c) out[4 * (size_t)i + c] = make_float4(m.c[c][0], m.c[c][1], m.c[c][2], m.c[c][3]);
}

int launch_models(World *w, float4 *d_out)
{
    const int n = w->d.nb + w->d.n_statics;
    if (n == 0) return NANS_OK;
    models_kernel<<<div_up(n, 128), 128, 0, w->stream>>>(w->d, d_out);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

__global__ void set_dt_kernel(float *dst, float dt) { *dst = dt; }

// dt lives in device memory so that the captured step graph is independent of it: the reference host passes the
// MEASURED frame time, which changes every frame (code/sdl_nans.cpp:986,999)
int launch_set_dt(World *w, float dt)
{
    set_dt_kernel<<<1, 1, 0, w->stream>>>(w->d.dt, dt);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

int launch_integrate_forces(World *w)
{
    if (w->d.nb == 0) return NANS_OK;
    integrate_forces_kernel<<<div_up(w->d.nb, 256), 256, 0, w->stream>>>(w->d);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

int launch_integrate_velocities(World *w)
{
    if (w->d.nb == 0) return NANS_OK;
    integrate_velocities_kernel<<<div_up(w->d.nb, kIvThreads), kIvThreads, 0, w->stream>>>(w->d);
    NANS_LAUNCH_CHECK();
    // the draw section also refreshes the Floor's Model and vertices every frame (:1870-1881)
    return launch_rebuild_statics(w);
}

int launch_rebuild_statics(World *w)
{
    if (w->d.n_statics == 0) return NANS_OK;
    rebuild_statics_kernel<<<1, 32, 0, w->stream>>>(w->d);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

}  // namespace nans
