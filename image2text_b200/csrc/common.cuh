// Shared device/host helpers for libi2t (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/i2t.h"

namespace i2t {

// ---- error plumbing (thread-local message, never throws across the ABI) -----------------------------
char* err_buf();
int fail(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define I2T_REQUIRE(cond, ...)                              \
  do {                                                      \
    if (!(cond)) return ::i2t::fail(I2T_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define I2T_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      (void)cudaGetLastError(); /* do not leave a sticky error for unrelated later launches */ \
      return ::i2t::fail(I2T_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    }                                                                                        \
  } while (0)

// call after every kernel launch
#define I2T_LAUNCHED()                                                                        \
  do {                                                                                        \
    ::i2t::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess)                                                                   \
      return ::i2t::fail(I2T_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// ---- programmatic dependent launch (PDL): a kernel launched through launch_pdl may start while its predecessor in the
// stream is still running; it must call pdl_wait() before it touches anything the predecessor wrote (or overwrites anything
// the predecessor reads), and calls pdl_launch_dependents() early so that ITS successor can do the same.  Weights never
// depend on a predecessor: the decode step's kernels fetch them before pdl_wait().  i2t_set_pdl(0) switches the attribute off.
extern std::atomic<int> g_pdl;
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl.load() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool valid_dtype(int d) { return d == I2T_F32 || d == I2T_BF16; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
int num_sms();

// ---- device helpers ---------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// load 4 consecutive elements as fp32 (pointer must be 16 B aligned for float, 8 B for bf16)
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}

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

// GELU flavours.  tanhf/erff (not the fast intrinsics): the fp32 path is the 1e-4 parity anchor.
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.0f + tanhf(k0 * (x + k1 * x * x * x)));
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f)); }
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == I2T_ACT_GELU_TANH) return gelu_tanh_f(x);
  if (act == I2T_ACT_GELU_ERF) return gelu_erf_f(x);
  return x;
}
// derivative of the activation w.r.t. its pre-activation input
__device__ __forceinline__ float act_grad(float x, int act) {
  if (act == I2T_ACT_GELU_TANH) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float u = k0 * (x + k1 * x * x * x);
    float t = tanhf(u);
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * x * x);
  }
  if (act == I2T_ACT_GELU_ERF) {
    return 0.5f * (1.0f + erff(x * 0.7071067811865476f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
  }
  return 1.0f;
}

}  // namespace i2t
