// Shared device/host helpers for the coma_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/coma_b200.h"

namespace coma {

void set_error(const char* fmt, ...);

#define COMA_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      coma::set_error(__VA_ARGS__);    \
      return COMA_ERR_INVALID;         \
    }                                  \
  } while (0)

#define COMA_CHECK_LAUNCH(name)                                                  \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      coma::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return COMA_ERR_CUDA;                                                      \
    }                                                                            \
  } while (0)

template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// Load 8 consecutive elements (16-byte aligned for bf16, 32-byte for f32) as floats.
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
// streaming variants: read-once data, do not allocate in L1
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// raw (still packed) 8-element group: issue the loads of several groups first, unpack at the point of use
template <typename T> struct Raw8;
template <> struct Raw8<float> { uint4 q[2]; };
template <> struct Raw8<__nv_bfloat16> { uint4 q[1]; };
__device__ __forceinline__ void ldraw_stream(const float* p, Raw8<float>& r) { r.q[0] = ldg_stream16(p); r.q[1] = ldg_stream16(p + 4); }
__device__ __forceinline__ void ldraw_stream(const __nv_bfloat16* p, Raw8<__nv_bfloat16>& r) { r.q[0] = ldg_stream16(p); }
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float (&v)[8]) {
  v[0] = __uint_as_float(r.q[0].x); v[1] = __uint_as_float(r.q[0].y); v[2] = __uint_as_float(r.q[0].z); v[3] = __uint_as_float(r.q[0].w);
  v[4] = __uint_as_float(r.q[1].x); v[5] = __uint_as_float(r.q[1].y); v[6] = __uint_as_float(r.q[1].z); v[7] = __uint_as_float(r.q[1].w);
}
__device__ __forceinline__ void unpack8(const Raw8<__nv_bfloat16>& r, float (&v)[8]) {
  const uint32_t w[4] = {r.q[0].x, r.q[0].y, r.q[0].z, r.q[0].w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void load8_stream(const float* p, float (&v)[8]) {
  const uint4 a = ldg_stream16(p), b = ldg_stream16(p + 4);
  v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
  v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
}
__device__ __forceinline__ void load8_stream(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 r = ldg_stream16(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  r.x = pack_bf16x2(v[0], v[1]); r.y = pack_bf16x2(v[2], v[3]);
  r.z = pack_bf16x2(v[4], v[5]); r.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = r;
}

// activation on the pre-activation u; slope is the PReLU / LeakyReLU negative slope
__device__ __forceinline__ float act_fwd(int act, float u, float slope) {
  switch (act) {
    case COMA_ACT_RELU: return u > 0.f ? u : 0.f;
    case COMA_ACT_LEAKY: return u > 0.f ? u : slope * u;
    case COMA_ACT_SIGMOID: return 1.f / (1.f + __expf(-u));
    case COMA_ACT_LEAKY_RELU: return u > 0.f ? u : fmaxf(slope * u, 0.f);
    default: return u;
  }
}
// d act / d u
__device__ __forceinline__ float act_grad(int act, float u, float slope) {
  switch (act) {
    case COMA_ACT_RELU: return u > 0.f ? 1.f : 0.f;
    case COMA_ACT_LEAKY: return u > 0.f ? 1.f : slope;
    case COMA_ACT_SIGMOID: { float s = 1.f / (1.f + __expf(-u)); return s * (1.f - s); }
    case COMA_ACT_LEAKY_RELU: return u > 0.f ? 1.f : (slope < 0.f ? slope : 0.f);
    default: return 1.f;
  }
}
// d act / d slope (per unit of upstream gradient)
__device__ __forceinline__ float act_slope_grad(int act, float u, float slope) {
  if (act == COMA_ACT_LEAKY) return u < 0.f ? u : 0.f;
  if (act == COMA_ACT_LEAKY_RELU) return (u < 0.f && slope < 0.f) ? u : 0.f;
  return 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace coma
