// Shared helpers for the AutoMoE sm_100a kernels (host + device).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdlib.h>

#include <atomic>
#include <utility>

#include "../../include/automoe_b200.h"

struct amoe_ctx {
  int device;
  int sm_count;
  std::atomic<int64_t> launches;
  int walk_reverse;   // tcgen05 convolutions walk their tiles back to front (amoe_set_walk_reverse)
  // cuTensorMapEncodeTiled resolved through the runtime (no -lcuda needed)
  CUresult (*encode_tiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                           const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
};

void amoe_set_error(const char* fmt, ...);

#define AMOE_CHECK_CUDA(expr)                                                          \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      amoe_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -1;                                                                       \
    }                                                                                  \
  } while (0)

#define AMOE_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      amoe_set_error(__VA_ARGS__);   \
      return -1;                     \
    }                                \
  } while (0)

// launch-error check without synchronising
#define AMOE_LAUNCH_OK(ctx)                                  \
  do {                                                       \
    (ctx)->launches.fetch_add(1, std::memory_order_relaxed); \
    AMOE_CHECK_CUDA(cudaPeekAtLastError());                  \
  } while (0)

// Every entry point that launches work runs on the context's device, whatever the calling thread's
// current device is (a model on cuda:1 driven from a thread whose current device is 0), and restores it.
struct amoe_device_scope {
  int prev = -1;
  explicit amoe_device_scope(const amoe_ctx* ctx) {
    if (ctx == nullptr) return;
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != ctx->device && cudaSetDevice(ctx->device) == cudaSuccess) prev = cur;
  }
  ~amoe_device_scope() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  amoe_device_scope(const amoe_device_scope&) = delete;
  amoe_device_scope& operator=(const amoe_device_scope&) = delete;
};
#define AMOE_ENTER(ctx) amoe_device_scope _amoe_dev_scope(ctx)

// AMOE_PDL=0 turns programmatic dependent launch of the tensor-core convolution chain off (A/B switch)
static inline bool amoe_pdl_enabled() {
  static const int on = [] {
    const char* e = getenv("AMOE_PDL");
    return (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }();
  return on != 0;
}

// Launch `kern` so that it may start before the previous kernel of the stream has drained (see tc_common.cuh:
// griddep_wait).  Only for kernels that call griddepcontrol.wait before their first dependent global access.
template <typename... KArgs, typename... Args>
static inline cudaError_t amoe_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = amoe_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// Same, as thread-block clusters of cluster_x CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
static inline cudaError_t amoe_launch_pdl_cluster(void (*kern)(KArgs...), int cluster_x, dim3 grid, dim3 block, size_t smem,
                                                  cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = amoe_pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- dtype-generic scalar load/store (device) ----
template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void st_from_float(T* p, float v);
template <>
__device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
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

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
