// Shared helpers for the sm_100a kernels behind include/bnn_b200.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bnn_b200.h"

namespace bnn {

void set_error(const char* fmt, ...);
// conv_tc.cu: the exit-head classifier as a tcgen05 GEMM (fp32 logits); used by kernels_head.cu
int conv_tc_head_gemm(const void* a, const void* w3, const float* bias_pad, float* logits, int dtype, int M, int Fa, int Cpad,
                      int kblocks, int head_c, void* stream);
int check_device();  // BNN_OK iff current device is sm_100
int sm_count();

#define BNN_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      bnn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return BNN_E_CUDA;                                                                    \
    }                                                                                       \
  } while (0)

#define BNN_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      bnn::set_error(__VA_ARGS__);    \
      return BNN_E_ARG;               \
    }                                 \
  } while (0)

#define BNN_LAUNCH_OK()                                                           \
  do {                                                                            \
    cudaError_t _e = cudaGetLastError();                                          \
    if (_e != cudaSuccess) {                                                      \
      bnn::set_error("%s:%d: launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return BNN_E_CUDA;                                                          \
    }                                                                             \
  } while (0)

// ---- storage <-> float --------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <>
__device__ __forceinline__ float to_f32<uint8_t>(uint8_t v) { return (float)v; }   // 8-bit activations: the integer itself

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ uint8_t from_f32<uint8_t>(float v) {   // round half to even, saturate to [0, 255]
  return (uint8_t)min(max(__float2int_rn(v), 0), 255);
}
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) {
  // saturate instead of producing inf: fp16 activations must stay finite
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// pack two floats into one 32-bit word of 16-bit storage
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <>
__device__ __forceinline__ uint32_t pack2<int8_t>(float, float) { return 0u; }   // 8-bit kernels never take this path
template <typename T>
__device__ __forceinline__ float2 unpack2(uint32_t v);
template <>
__device__ __forceinline__ float2 unpack2<__half>(uint32_t v) {
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}
template <>
__device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
template <>
__device__ __forceinline__ float2 unpack2<int8_t>(uint32_t) { return make_float2(0.f, 0.f); }

}  // namespace bnn
