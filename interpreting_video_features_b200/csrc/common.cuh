// Internal helpers shared by the libivf.so translation units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ivf.h"

struct ivf_handle {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  int64_t launches = 0;
  std::mutex mu;
  // cache of encoded TMA tensor maps keyed by a byte string of everything that shapes them
  std::map<std::string, CUtensorMap> tmaps;
  bool tc_attr_set[4] = {false, false, false, false};
  bool slab_attr_set[16] = {};  // per kernel instantiation: dynamic shared memory opt-in done ([15]: grouped)
  // small per-handle scratch (partial logits of the head kernels); allocated once in ivf_create so
  // that no call allocates (CUDA-graph capture safe), freed in ivf_destroy
  float* scratch = nullptr;
  size_t scratch_bytes = 0;
};

void ivf_set_error(const char* fmt, ...);

// Every entry point that takes a handle issues its work on the handle's device, whatever device is current
// in the calling thread (an engine built for cuda:1 may be driven while cuda:0 is current): switch for the
// duration of the call, restore on return.  cudaGetDevice is a thread-local read; no switch, no cost.
struct ivf_device_guard {
  int prev = -1;
  bool switched = false;
  explicit ivf_device_guard(const ivf_handle* h) {
    if (h && cudaGetDevice(&prev) == cudaSuccess && prev != h->device)
      switched = cudaSetDevice(h->device) == cudaSuccess;
  }
  ~ivf_device_guard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define IVF_ON_DEVICE(h) ivf_device_guard ivf_guard__(h)

#define IVF_FAIL(code, ...)       \
  do {                            \
    ivf_set_error(__VA_ARGS__);   \
    return (code);                \
  } while (0)

#define IVF_REQUIRE(cond, ...)                       \
  do {                                               \
    if (!(cond)) IVF_FAIL(IVF_EINVAL, __VA_ARGS__);  \
  } while (0)

#define IVF_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      IVF_FAIL(IVF_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),       \
               __FILE__, __LINE__);                                                      \
  } while (0)

// every kernel launch is followed by this: counts the launch and surfaces launch errors
#define IVF_LAUNCHED(h)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess)                                                              \
      IVF_FAIL(IVF_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),   \
               __FILE__, __LINE__);                                                      \
    (h)->launches++;                                                                     \
  } while (0)

static inline int ivf_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// A kernel that calls ivf_pdl_wait() before its first access to memory written by earlier kernels may be
// launched with the programmatic-stream-serialisation attribute: its CTAs become resident and run their set-up
// (barrier init, TMEM allocation, tensor-map prefetch, constant per-channel vectors) while the previous kernel of
// the stream drains, and block at the wait until that kernel has completed and flushed.  ivf_pdl_trigger() at the
// top of a kernel lets ITS dependents start the same way.  Both are no-ops for a normal launch.  Under stream
// capture the attribute becomes a programmatic edge of the CUDA graph.  IVF_PDL=0 launches everything normally.
__device__ __forceinline__ void ivf_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void ivf_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool ivf_pdl_enabled();

// launch `kernel` (which must contain ivf_pdl_wait) with or without the PDL attribute, optionally as CTA pairs
template <typename... P, typename... A>
inline cudaError_t ivf_launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                              A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (ivf_pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

// ---- dtype helpers (device) -------------------------------------------------------
__device__ __forceinline__ float ivf_to_float(float v) { return v; }
__device__ __forceinline__ float ivf_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T ivf_from_float(float v);
template <>
__device__ __forceinline__ float ivf_from_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 ivf_from_float<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

__device__ __forceinline__ float ivf_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ivf_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float ivf_warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// entry points implemented per translation unit
int ivf_conv3d_f32_launch(ivf_handle* h, const ivf_conv_desc* d, const float* in, const float* w,
                          const float* scale, const float* shift, const float* acc_in,
                          const float* mask_y, const float* mask_scale, float* out,
                          cudaStream_t st);
int ivf_conv3d_tc_launch(ivf_handle* h, const ivf_conv_desc* d, const void* in, const void* w,
                         const float* scale, const float* shift, const float* acc_in,
                         const void* mask_y, const float* mask_scale, void* out, cudaStream_t st,
                         const ivf_conv_split* sp = nullptr, const void* in2 = nullptr, void* out2 = nullptr);
// halo-slab tcgen05 kernel (conv_slab.cu) for the wide stride-1 'same' layers; the launcher of the im2col
// kernel above serves everything else the bf16 path supports
bool ivf_conv3d_slab_eligible(const ivf_handle* h, const ivf_conv_desc* d);
// two independent slab-kernel convolutions as ONE launch; IVF_EUNSUPPORTED (and nothing launched) when the pair
// cannot be grouped - the caller then issues them one after the other
struct ivf_conv_operands {
  const void* in;
  const void* w;
  const float* scale;
  const float* shift;
  const float* acc_in;
  const void* mask_y;
  const float* mask_scale;
  void* out;
};
int ivf_conv3d_slab_launch_pair(ivf_handle* h, const ivf_conv_desc* const d[2], const ivf_conv_operands op[2],
                                cudaStream_t st);
int ivf_conv3d_slab_launch(ivf_handle* h, const ivf_conv_desc* d, const void* in, const void* w,
                           const float* scale, const float* shift, const float* acc_in,
                           const void* mask_y, const float* mask_scale, void* out, cudaStream_t st,
                           const float* lstm_c_prev = nullptr, float* lstm_c_next = nullptr,
                           void* lstm_h_next = nullptr);
