/* TEST INFRASTRUCTURE -- stand-in for CUDA-samples' helper_math.h, which the reference's spc_math.h:26 includes but
 * does not vendor.  Only the float3 operators the two compiled reference files use (spc_raytrace_cuda_kernel.cu:121-122:
 * float * float3; spc_raytrace_cuda.cpp:88-90: float3 - float3, normalize).  Semantics are those of the CUDA samples
 * header (component-wise ops; normalize = v * rsqrtf(dot(v, v)) there -- on the host path used here, 1/sqrtf).  The
 * ray-generation helper that calls normalize() is not part of any parity test. */
#pragma once
#include <cuda_runtime.h>
#include <math.h>
inline __host__ __device__ float3 operator*(float a, float3 b) { return make_float3(a * b.x, a * b.y, a * b.z); }
inline __host__ __device__ float3 operator*(float3 b, float a) { return make_float3(b.x * a, b.y * a, b.z * a); }
inline __host__ __device__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline __host__ __device__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline __host__ __device__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline __host__ __device__ float3 normalize(float3 v) { const float r = 1.0f / sqrtf(dot(v, v)); return v * r; }
