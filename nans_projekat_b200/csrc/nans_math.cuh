// nans_math.cuh — device fp32 vector math with the reference's exact operation order.
//
// The reference binary is x86-64 SSE2 at -O0: plain IEEE fp32, never fused, glm 0.9.9 scalar
// code paths.  Bit-exact GJK flags need the same roundings on the GPU, so
// every operation here is an explicit round-to-nearest intrinsic (__fadd_rn/__fmul_rn/__fdiv_rn/
// __fsqrt_rn are never contracted into FMA by nvcc, whatever -fmad says), denormals are kept
// (default -ftz=false), and division/sqrt are the IEEE-correct ones (no rcp/rsqrt approximations).
#pragma once
#include <cuda_runtime.h>
#include <float.h>

namespace nans {

struct vec3 { float x, y, z; };

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
// 1.0f / a: the correctly rounded reciprocal IS the correctly rounded quotient (same real number, same rounding), at
// about half the instructions of the general division
__device__ __forceinline__ float frcp(float a) { return __frcp_rn(a); }

__device__ __forceinline__ vec3 V3(float x, float y, float z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ vec3 V3(float4 v) { return V3(v.x, v.y, v.z); }
__device__ __forceinline__ vec3 operator+(vec3 a, vec3 b) { return V3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
__device__ __forceinline__ vec3 operator-(vec3 a, vec3 b) { return V3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
__device__ __forceinline__ vec3 operator*(vec3 a, float s) { return V3(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
__device__ __forceinline__ vec3 operator*(float s, vec3 a) { return V3(fmul(s, a.x), fmul(s, a.y), fmul(s, a.z)); }
__device__ __forceinline__ vec3 operator/(vec3 a, float s) { return V3(fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)); }
__device__ __forceinline__ vec3 operator-(vec3 a) { return V3(-a.x, -a.y, -a.z); }
// glm::dot(vec3): (x*x' + y*y') + z*z'
__device__ __forceinline__ float dot(vec3 a, vec3 b)
{
    return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z));
}
// glm::cross: (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
__device__ __forceinline__ vec3 cross(vec3 x, vec3 y)
{
    return V3(fsub(fmul(x.y, y.z), fmul(y.y, x.z)),
              fsub(fmul(x.z, y.x), fmul(y.z, x.x)),
              fsub(fmul(x.x, y.y), fmul(y.x, x.y)));
}
__device__ __forceinline__ float length(vec3 a) { return fsqrt(dot(a, a)); }
// glm::normalize: v * inversesqrt(dot(v,v)), inversesqrt(x) = 1.0f / sqrt(x)  (normalize(0) = NaN)
__device__ __forceinline__ vec3 normalize(vec3 a) { return a * frcp(fsqrt(dot(a, a))); }
__device__ __forceinline__ bool equal(vec3 a, vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

}  // namespace nans
