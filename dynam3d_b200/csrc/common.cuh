// Shared helpers for libdynam3d_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dynam3d_b200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns 0 or a negative D3D_E* code; message via d3d_last_error()
// ---------------------------------------------------------------------------------------------
void d3d_set_error(const char* fmt, ...);

#define D3D_CHECK_CUDA(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      d3d_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));        \
      return D3D_ECUDA;                                                                           \
    }                                                                                             \
  } while (0)

#define D3D_REQUIRE(cond, msg)                                                                    \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      d3d_set_error("%s:%d: requirement failed: %s (%s)", __FILE__, __LINE__, #cond, msg);        \
      return D3D_EINVAL;                                                                          \
    }                                                                                             \
  } while (0)

// every kernel launch site is followed by this macro: it also counts the launch (d3d_launch_count(), bench.py `gpu_launches`)
extern unsigned long long g_d3d_launches;
#define D3D_CHECK_LAUNCH()                  \
  do {                                      \
    ++g_d3d_launches;                       \
    D3D_CHECK_CUDA(cudaGetLastError());     \
  } while (0)

#define D3D_TRY(expr)                                                                             \
  do {                                                                                            \
    int _r = (expr);                                                                              \
    if (_r != 0) return _r;                                                                       \
  } while (0)

static inline int d3d_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
int d3d_num_sms();

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

// storage type selected at run time: 0 = fp16, 1 = bf16, 2 = fp32 (matches D3D_F16 / D3D_BF16 / D3D_OUT_F32);
// the fp32 kind is what the "precise" (split fp16x2 operand) pipeline stores between kernels
__device__ __forceinline__ float ld16(const void* p, size_t i, int kind) {
  if (kind == D3D_OUT_F32) return ((const float*)p)[i];
  return kind == D3D_BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : __half2float(((const __half*)p)[i]);
}
__device__ __forceinline__ void st16(void* p, size_t i, float v, int kind) {
  if (kind == D3D_OUT_F32) ((float*)p)[i] = v;
  else if (kind == D3D_BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
  else ((__half*)p)[i] = __float2half_rn(v);
}
__device__ __forceinline__ uint32_t pack16x2(float a, float b, int kind) {
  if (kind == D3D_BF16) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
  }
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack16x2(uint32_t v, int kind) {
  if (kind == D3D_BF16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}
