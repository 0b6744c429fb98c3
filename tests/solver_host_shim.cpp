// solver_host_shim.cpp — TEST INFRASTRUCTURE.  Compiles csrc/solver_constraint.cuh (the device code of one
// Constraint: constraint_prepare + constraint_apply with the pipelined accumulation the kernel runs) as host C++
// and runs the reference's sequential sweep (SolveConstraints, code/nans.cpp:1539-1548) with it, so the solver's
// arithmetic can be compared with the oracle without a GPU.  The GPU tests cover the parallel schedule.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float cuda_fmaxf(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return signbit(a) ? b : a;
    return a > b ? a : b;
}
static inline float cuda_fminf(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return signbit(a) ? a : b;
    return a < b ? a : b;
}
#define fmaxf cuda_fmaxf
#define fminf cuda_fminf

#include "../nans_projekat_b200/csrc/solver_constraint.cuh"

using namespace nans;

// contacts: the 48-byte records of include/nans_b200.h (type, a, b, point_a, point_b, n).  Body arrays are
// [nb][3] (cubes first, then spheres), statics [ns][3]; vel / angvel are updated in place.
extern "C" void solver_host_sweep(int n_cubes, int n_spheres, int n_statics, const float *pos, float *vel, float *angvel,
                                  const float *mass, const float *moi, const float *st_pos, const float *st_mass,
                                  const float *st_moi, const nans_contact *contacts, int n_contacts, float dt)
{
    const int nb = n_cubes + n_spheres;
    std::vector<float4> P(nb), V(nb), W(nb), SP(n_statics > 0 ? n_statics : 1), SA(n_statics > 0 ? n_statics : 1);
    for (int i = 0; i < nb; ++i) {
        P[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], mass[i]);
        V[i] = make_float4(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2], 1.0f / mass[i]);
        W[i] = make_float4(angvel[3 * i], angvel[3 * i + 1], angvel[3 * i + 2], 1.0f / moi[i]);
    }
    for (int k = 0; k < n_statics; ++k) {
        SP[k] = make_float4(st_pos[3 * k], st_pos[3 * k + 1], st_pos[3 * k + 2], 1.0f / st_mass[k]);
        SA[k] = make_float4(0.f, 0.f, 0.f, 1.0f / st_moi[k]);
    }
    std::vector<float4> cpa(n_contacts > 0 ? n_contacts : 1), cpb(cpa.size()), cn(cpa.size());
    for (int i = 0; i < n_contacts; ++i) {          // the row encoding of nans_set_contacts (api.cu)
        const nans_contact &c = contacts[i];
        const bool a_sph = (c.type == NANS_SS || c.type == NANS_SF);
        const int a = a_sph ? n_cubes + c.a : c.a;
        int b;
        if (c.type == NANS_CF || c.type == NANS_SF) b = -(c.b + 1);
        else b = (c.type == NANS_CS || c.type == NANS_SS) ? n_cubes + c.b : c.b;
        cpa[i] = make_float4(c.point_a[0], c.point_a[1], c.point_a[2], __int_as_float(a));
        cpb[i] = make_float4(c.point_b[0], c.point_b[1], c.point_b[2], __int_as_float(b));
        cn[i] = make_float4(c.n[0], c.n[1], c.n[2], 0.f);
    }
    DeviceWorld w;
    memset(&w, 0, sizeof(w));
    w.n_cubes = n_cubes; w.n_spheres = n_spheres; w.n_statics = n_statics; w.nb = nb; w.n_owned = nb;
    w.pos = P.data(); w.vel = V.data(); w.angvel = W.data();
    w.st_pos = SP.data(); w.st_ang = SA.data();
    w.c_pa = cpa.data(); w.c_pb = cpb.data(); w.c_n = cn.data();
    for (int c = 0; c < n_contacts; ++c) {          // list order: each Constraint reads what the previous ones wrote
        float4 q[kRecQuads];
        int ia, ib;
        constraint_prepare(w, c, dt, q, ia, ib);
        vec3 V1 = V3(V[ia]), W1 = V3(W[ia]);
        vec3 V2 = V3(0.f, 0.f, 0.f), W2 = V3(0.f, 0.f, 0.f);          // the Floor: V = W = 0, never written
        if (ib >= 0) { V2 = V3(V[ib]); W2 = V3(W[ib]); }
        constraint_apply(q, V1, W1, V2, W2, ib >= 0);
        V[ia].x = V1.x; V[ia].y = V1.y; V[ia].z = V1.z; W[ia].x = W1.x; W[ia].y = W1.y; W[ia].z = W1.z;
        if (ib >= 0) { V[ib].x = V2.x; V[ib].y = V2.y; V[ib].z = V2.z; W[ib].x = W2.x; W[ib].y = W2.y; W[ib].z = W2.z; }
    }
    for (int i = 0; i < nb; ++i) {
        vel[3 * i] = V[i].x; vel[3 * i + 1] = V[i].y; vel[3 * i + 2] = V[i].z;
        angvel[3 * i] = W[i].x; angvel[3 * i + 1] = W[i].y; angvel[3 * i + 2] = W[i].z;
    }
}
