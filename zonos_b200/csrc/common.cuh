// Internal helpers shared by the kernels of libzonos_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/zonos_b200.h"

#define ZB_NUM_SMS 148
#define ZB_NUM_COUNTERS (1u << 20)

typedef __nv_bfloat16 bf16;

struct zb_ctx {
  int device = 0;
  int num_sms = ZB_NUM_SMS;
  std::string err;
  int64_t launches = 0;
  // scratch arenas (device), grown on demand.  `scratch` serves the backbone (its pointers are baked into the
  // CUDA graph of a live generate session, so it must not move while `scratch_pins` > 0); `dac_scratch`
  // serves zb_dac_decode, which may run between the steps of a session (streaming decode).
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  int scratch_pins = 0;
  void* dac_scratch = nullptr;
  size_t dac_scratch_bytes = 0;
  int32_t* counters = nullptr;  // zeroed int32 words for last-CTA-done patterns
  cudaStream_t capture_stream = nullptr;   // graph capture happens here (the caller's stream may be the legacy one)
};

extern thread_local std::string g_zb_create_error;

zb_status zb_fail(zb_ctx* ctx, zb_status code, const char* fmt, ...);

#define ZB_CUDA(ctx, expr)                                                                      \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return zb_fail((ctx), ZB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                       \
  } while (0)

#define ZB_CHECK_LAUNCH(ctx)                                                                 \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return zb_fail((ctx), ZB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                  \
                     cudaGetErrorString(_e), __FILE__, __LINE__);                            \
    (ctx)->launches++;                                                                       \
  } while (0)

#define ZB_REQUIRE(ctx, cond, ...)                                    \
  do {                                                                \
    if (!(cond)) return zb_fail((ctx), ZB_ERR_INVALID, __VA_ARGS__);  \
  } while (0)

zb_status zb_scratch_reserve(zb_ctx* ctx, size_t bytes);
zb_status zb_dac_scratch_reserve(zb_ctx* ctx, size_t bytes);

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
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

// bf16 pair packed in a 32-bit word -> two floats (exact)
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16_rn(v); }
// round a float to bf16 precision and come back (the reference rounds after every op)
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) |
         ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
}

// streaming 16-byte load: weights and KV are read once per step, keep them out of L1
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Programmatic dependent launch (PDL): wait for the producer grid / let the consumer grid start early
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
