// Internal helpers shared by the kernels of libzonos_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/zonos_b200.h"

#define ZB_NUM_SMS 148
#define ZB_NUM_COUNTERS (1u << 20)

typedef __nv_bfloat16 bf16;

// Device + pinned memory of one generate session.  cudaMalloc / cudaFree / cudaFreeHost per session cost up to hundreds
// of milliseconds on a busy box (measured: zb_generate_end 0.6 - 700 ms), so finished sessions hand their slab back to
// the context and the next session reuses it.
struct zb_gen_slab {
  void* dev = nullptr; size_t dev_bytes = 0; int32_t* host = nullptr; int32_t* host_dev = nullptr; bool in_use = false;
  std::vector<unsigned char> layers; size_t layers_off = 0;   // host copy of the layer table last uploaded into this slab
  // pinned staging of that table: the upload is asynchronous; `stage_ev` marks the last copy OUT of the staging buffer
  void* stage = nullptr; size_t stage_bytes = 0; cudaEvent_t stage_ev = nullptr;
};

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
  void* tc_ws = nullptr;          // split-K partial sums of the tcgen05 GEMM
  size_t tc_ws_bytes = 0;
  int32_t* counters = nullptr;  // zeroed int32 words for last-CTA-done patterns
  cudaStream_t capture_stream = nullptr;   // graph capture happens here (the caller's stream may be the legacy one)
  std::vector<zb_gen_slab> gen_slabs;      // session memory, reused across generate calls
  // cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE setting: high-water mark per kernel for this context's device
  std::unordered_map<const void*, size_t> smem_attr;
};

// Raise the dynamic shared-memory limit of `kernel` on the context's device if it is below `smem`.
template <typename K>
inline cudaError_t zb_ensure_smem(zb_ctx* ctx, K kernel, size_t smem) {
  size_t& cur = ctx->smem_attr[reinterpret_cast<const void*>(kernel)];
  if (cur >= smem) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) cur = smem;
  return e;
}

// Every extern "C" entry point runs on the context's device, whatever device the calling thread has current
// (the reference wraps generate in `with torch.device(device)`), and restores the caller's device on return.
struct zb_device_guard {
  int prev = -1;
  explicit zb_device_guard(const zb_ctx* ctx) {
    if (!ctx) return;
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != ctx->device && cudaSetDevice(ctx->device) == cudaSuccess) prev = cur;
  }
  ~zb_device_guard() { if (prev >= 0) cudaSetDevice(prev); }
  zb_device_guard(const zb_device_guard&) = delete;
  zb_device_guard& operator=(const zb_device_guard&) = delete;
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
zb_status zb_tc_workspace_reserve(zb_ctx* ctx, size_t bytes);

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

// L2 prefetch of `bytes` (multiple of 16) starting at p: pulls weights HBM -> L2 ahead of the dependency wait
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Launch with the programmatic-stream-serialization attribute: the kernel may start while its predecessor in the
// stream is still running and must call pdl_wait() before touching anything the predecessor writes.
template <typename... KArgs, typename... Args>
inline cudaError_t zb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ---- mbarrier + bulk-copy (TMA engine, UBLKCP) helpers ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a lost transaction must end in a trap (an error the host sees), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spins = 0;; ++spins) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spins > (1u << 24)) asm volatile("trap;");
  }
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy (bytes % 16 == 0, both sides 16-byte aligned), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// same without an L2 cache hint (measured on B200: the evict_first hint costs 6-8 % of streaming bandwidth)
__device__ __forceinline__ void bulk_g2s_nohint(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Reduce V (power of two) per-lane values across the warp with V-1 + (5 - log2 V) shuffles instead of 5*V:
// every halving step trades half of the values with the xor-partner.  On return v[0] of lane L holds the full sum of
// value index multi_reduce_index<V>(L) (all lanes sharing that index hold the same number).
template <int V>
__device__ __forceinline__ void warp_reduce_multi(float (&v)[V]) {
  constexpr int LOGV = (V == 1) ? 0 : (V == 2) ? 1 : (V == 4) ? 2 : (V == 8) ? 3 : 4;
  static_assert((1 << LOGV) == V && V <= 16, "V must be a power of two <= 16");
  const int lane = threadIdx.x & 31;
  int n = V;
#pragma unroll
  for (int s = 0; s < LOGV; ++s) {
    const int o = 16 >> s;
    const bool upper = (lane & o) != 0;
    n >>= 1;
#pragma unroll
    for (int j = 0; j < V / 2; ++j) {
      if (j < n) {
        const float keep = upper ? v[j + n] : v[j];
        const float send = upper ? v[j] : v[j + n];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
  }
#pragma unroll
  for (int o = 16 >> LOGV; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}
template <int V>
__device__ __forceinline__ int multi_reduce_index(int lane) {
  constexpr int LOGV = (V == 1) ? 0 : (V == 2) ? 1 : (V == 4) ? 2 : (V == 8) ? 3 : 4;
  int idx = 0;
#pragma unroll
  for (int s = 0; s < LOGV; ++s) idx += ((lane >> (4 - s)) & 1) * (V >> (s + 1));
  return idx;
}

// ---- packed fp32x2 FMA (sm_100: FFMA2) ------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ float sum_f32x2(unsigned long long v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return a + b;
}
// raw 32-bit shared-address forms (no generic->shared conversion inside hot loops)
__device__ __forceinline__ void mbar_wait_u32(uint32_t addr, uint32_t parity) {
  for (uint32_t spins = 0;; ++spins) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spins > (1u << 24)) asm volatile("trap;");
  }
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
