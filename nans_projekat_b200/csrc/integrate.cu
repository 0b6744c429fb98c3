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
#include "glibc_sincosf.cuh"
#include "nans_math.cuh"
#include "world.cuh"

namespace nans {

// MovementFunction, code/nans.cpp:64-70
__device__ __forceinline__ vec3 movement_fn(vec3 v, vec3 forces, float mass, float inv_mass)
{
    const float g = fmul(mass, 9.81f);
    const vec3 grav = V3(fmul(g, 0.0f), fmul(g, -1.0f), fmul(g, 0.0f));
    return inv_mass * ((forces + grav) - 1.5f * v);
}
// RotationFunction, code/nans.cpp:73-78
__device__ __forceinline__ vec3 rotation_fn(vec3 w, vec3 torque, float inv_moi)
{
    return inv_moi * (torque - 1.5f * w);
}

// RK4, code/nans.cpp:51-61 (only velocities are integrated; x is not part of the state)
template <bool kLinear>
__device__ __forceinline__ vec3 rk4(float dt, vec3 y0, vec3 sum, float m, float inv_m)
{
    auto F = [&](vec3 y) { return kLinear ? movement_fn(y, sum, m, inv_m) : rotation_fn(y, sum, inv_m); };
    const vec3 k1 = dt * F(y0);
    const vec3 k2 = dt * F(y0 + (k1 / 2.0f));
    const vec3 k3 = dt * F(y0 + (k2 / 2.0f));
    const vec3 k4 = dt * F(y0 + k3);
    return y0 + 0.16666667f * (((k1 + 2.0f * k2) + 2.0f * k3) + k4);
}

__global__ void __launch_bounds__(256) integrate_forces_kernel(DeviceWorld w, float dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.nb) return;
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

// ---- glm::translate / rotate / scale in the reference's operation order -------
struct mat4 { float c[4][4]; };  // c[col][row]

__device__ __forceinline__ void glm_rotate(mat4 &m, float angle, vec3 axis_in)
{
    const float c = nans_glibc::cosf_glibc(angle);
    const float s = nans_glibc::sinf_glibc(angle);
    const vec3 axis = normalize(axis_in);
    const vec3 temp = fsub(1.0f, c) * axis;
    float R[3][3];
    R[0][0] = fadd(c, fmul(temp.x, axis.x));
    R[0][1] = fadd(fmul(temp.x, axis.y), fmul(s, axis.z));
    R[0][2] = fsub(fmul(temp.x, axis.z), fmul(s, axis.y));
    R[1][0] = fsub(fmul(temp.y, axis.x), fmul(s, axis.z));
    R[1][1] = fadd(c, fmul(temp.y, axis.y));
    R[1][2] = fadd(fmul(temp.y, axis.z), fmul(s, axis.x));
    R[2][0] = fadd(fmul(temp.z, axis.x), fmul(s, axis.y));
    R[2][1] = fsub(fmul(temp.z, axis.y), fmul(s, axis.x));
    R[2][2] = fadd(c, fmul(temp.z, axis.z));
    mat4 o;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r)
            o.c[i][r] = fadd(fadd(fmul(m.c[0][r], R[i][0]), fmul(m.c[1][r], R[i][1])), fmul(m.c[2][r], R[i][2]));
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) m.c[i][r] = o.c[i][r];
}

// Model = T(pos) * Rx(radians(ang.x)) * Ry(..y) * Rz(..z) * S(scale); out = 8 world vertices
__device__ __forceinline__ void model_vertices(vec3 pos, vec3 ang, vec3 scale, float out[24])
{
    mat4 m;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) m.c[c][r] = (c == r) ? 1.0f : 0.0f;
#pragma unroll
    for (int r = 0; r < 4; ++r)
        m.c[3][r] = fadd(fadd(fadd(fmul(m.c[0][r], pos.x), fmul(m.c[1][r], pos.y)), fmul(m.c[2][r], pos.z)), m.c[3][r]);
    const float rad = 0.01745329251994329576923690768489f;  // glm::radians — Angles are fed as degrees
    glm_rotate(m, fmul(ang.x, rad), V3(1.0f, 0.0f, 0.0f));
    glm_rotate(m, fmul(ang.y, rad), V3(0.0f, 1.0f, 0.0f));
    glm_rotate(m, fmul(ang.z, rad), V3(0.0f, 0.0f, 1.0f));
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        m.c[0][r] = fmul(m.c[0][r], scale.x);
        m.c[1][r] = fmul(m.c[1][r], scale.y);
        m.c[2][r] = fmul(m.c[2][r], scale.z);
    }
    // UpdateVertices: vec3(Model * vec4(+-.5, +-.5, +-.5, 1)) = (m0*x + m1*y) + (m2*z + m3*w)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float cx = (k & 2) ? -0.5f : 0.5f;
        const float cy = (k & 4) ? -0.5f : 0.5f;
        const float cz = (k & 1) ? -0.5f : 0.5f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
            out[3 * k + r] = fadd(fadd(fmul(m.c[0][r], cx), fmul(m.c[1][r], cy)),
                                  fadd(fmul(m.c[2][r], cz), fmul(m.c[3][r], 1.0f)));
    }
}

__device__ __forceinline__ void store_verts(float4 *dst, const float v[24])
{
#pragma unroll
    for (int q = 0; q < 6; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

__global__ void __launch_bounds__(128) integrate_velocities_kernel(DeviceWorld w, float dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.nb) return;
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
        store_verts(w.verts + 6 * (size_t)i, verts);
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

int launch_integrate_forces(World *w, float dt)
{
    if (w->d.nb == 0) return NANS_OK;
    integrate_forces_kernel<<<div_up(w->d.nb, 256), 256, 0, w->stream>>>(w->d, dt);
    NANS_LAUNCH_CHECK();
    return NANS_OK;
}

int launch_integrate_velocities(World *w, float dt)
{
    if (w->d.nb == 0) return NANS_OK;
    integrate_velocities_kernel<<<div_up(w->d.nb, 128), 128, 0, w->stream>>>(w->d, dt);
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
