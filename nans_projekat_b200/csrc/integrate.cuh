// integrate.cuh — the arithmetic of the two integrators: RK4 and its derivative functions (code/nans.cpp:51-78)
// and the draw section's Model = T*Rx*Ry*Rz*S -> 8 world vertices (:1870-1881,1913-1941, UpdateVertices :395-407).
// Pure device functions; tests/test_integrate_host.py compiles this header for the host and compares it with the
// oracle bit for bit.
#pragma once
#include "glibc_sincosf.cuh"
#include "nans_math.cuh"

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
    // k / 2.0f as k * 0.5f: halving is exact (or rounds the same real number the same way in the subnormal range), so the
    // product has the quotient's bits without the six IEEE divisions per body
    const vec3 k2 = dt * F(y0 + (k1 * 0.5f));
    const vec3 k3 = dt * F(y0 + (k2 * 0.5f));
    const vec3 k4 = dt * F(y0 + k3);
    return y0 + 0.16666667f * (((k1 + 2.0f * k2) + 2.0f * k3) + k4);
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

// Model = T(pos) * Rx(radians(ang.x)) * Ry(..y) * Rz(..z) * S(scale): the draw section's rebuild
// (code/nans.cpp:1870-1881 floor, :1913-1941 cubes, :1971-1984 spheres), glm's operation order
__device__ __forceinline__ void model_matrix(vec3 pos, vec3 ang, vec3 scale, mat4 &m)
{
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
}

// out = the 8 world vertices of the unit cube under Model (UpdateVertices, code/nans.cpp:395-407)
__device__ __forceinline__ void model_vertices(vec3 pos, vec3 ang, vec3 scale, float out[24])
{
    mat4 m;
    model_matrix(pos, ang, scale, m);
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

}  // namespace nans
