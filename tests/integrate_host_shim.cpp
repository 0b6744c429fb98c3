// integrate_host_shim.cpp — TEST INFRASTRUCTURE.  Compiles csrc/integrate.cuh (RK4 + the Model/vertex rebuild, the
// device code of integrate_forces_kernel / integrate_velocities_kernel) as host C++ so it can be compared with the
// oracle without a GPU.  The per-body statements below are those of the two kernels (csrc/integrate.cu).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }

#include "../nans_projekat_b200/csrc/integrate.cuh"

using namespace nans;

// IntegrateForces (code/nans.cpp:975-1018) over [n][3] arrays, in place
extern "C" void integrate_forces_host(int n, float *vel, float *angvel, float *force, float *torque, const float *mass,
                                      const float *moi, float dt)
{
    for (int i = 0; i < n; ++i) {
        const vec3 nv = rk4<true>(dt, V3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]),
                                  V3(force[3 * i], force[3 * i + 1], force[3 * i + 2]), mass[i], 1.0f / mass[i]);
        const vec3 na = rk4<false>(dt, V3(angvel[3 * i], angvel[3 * i + 1], angvel[3 * i + 2]),
                                   V3(torque[3 * i], torque[3 * i + 1], torque[3 * i + 2]), 0.0f, 1.0f / moi[i]);
        vel[3 * i] = nv.x; vel[3 * i + 1] = nv.y; vel[3 * i + 2] = nv.z;
        angvel[3 * i] = na.x; angvel[3 * i + 1] = na.y; angvel[3 * i + 2] = na.z;
        for (int k = 0; k < 3; ++k) force[3 * i + k] = torque[3 * i + k] = 0.0f;
    }
}

// IntegrateVelocities (:1332-1349) + Model rebuild + UpdateVertices: pos / ang in place, verts [n_cubes][24] out
extern "C" void integrate_velocities_host(int n, int n_cubes, float *pos, float *ang, const float *vel, const float *angvel,
                                          const float *scale, float *verts, float dt)
{
    for (int i = 0; i < n; ++i) {
        const vec3 np = V3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]) + dt * V3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        const vec3 na = V3(ang[3 * i], ang[3 * i + 1], ang[3 * i + 2]) + dt * V3(angvel[3 * i], angvel[3 * i + 1], angvel[3 * i + 2]);
        pos[3 * i] = np.x; pos[3 * i + 1] = np.y; pos[3 * i + 2] = np.z;
        ang[3 * i] = na.x; ang[3 * i + 1] = na.y; ang[3 * i + 2] = na.z;
        if (i < n_cubes) model_vertices(np, na, V3(scale[3 * i], scale[3 * i + 1], scale[3 * i + 2]), verts + 24 * i);
    }
}
