// Classifier head, forward and backward, one block per clip.
// Reference: pt/models/I3D_doubled.py:360-371 — AvgPool3d over the whole Mixed_5c map,
// dropout (identity in eval), 1x1x1 conv with bias, squeeze, optional softmax(dim=1) — and its
// autograd.  With p == 1 it is the Linear(+Softmax) of pt/models/CLSTM_4.py:78-83.
// Kept in fp32: the class gradient w.r.t. the mask is ~1e-9 with random weights.
#include "common.cuh"

namespace {

constexpr int HEAD_THREADS = 256;

__device__ float block_reduce(float v, float* red, bool is_max) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = is_max ? ivf_warp_max(v) : ivf_warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  int nw = blockDim.x >> 5;
  float r = threadIdx.x < nw ? red[threadIdx.x] : (is_max ? -INFINITY : 0.f);
  if (w == 0) {
    r = is_max ? ivf_warp_max(r) : ivf_warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  __syncthreads();
  return r;
}

constexpr int HEAD_CHUNK = 64;  // channels per block: 32 lanes x 2 adjacent channels
constexpr int HEAD_WARPS = HEAD_THREADS / 32;
constexpr int HEAD_MAX_PP = 16;  // positions per warp kept in flight at once

// two adjacent channels k, k+1 of one position (vector load when the row is even-aligned)
template <typename T, bool VEC2>
__device__ __forceinline__ float2 head_load2(const T* p, bool has2) {
  if (VEC2) {
    if (sizeof(T) == 2) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(p);
      return make_float2(__low2float(v), __high2float(v));
    }
    return *reinterpret_cast<const float2*>(p);
  }
  return make_float2(ivf_to_float(p[0]), has2 ? ivf_to_float(p[1]) : 0.f);
}

// Forward, stage 1: grid (clip, channel chunk).  The head is a few MFLOP on a few MB: pure latency, so
// the kernels are shaped to have every load of a phase in flight at once.  Lane = channel pair,
// warp = position group (pool) then class group (logits); partial logits -> scratch[clip][chunk][class].
template <typename T, bool VEC2>
__global__ void __launch_bounds__(HEAD_THREADS)
head_fwd_partial_kernel(const T* __restrict__ feat, int p, int c, int ld, const float* __restrict__ w,
                        int ncls, float* __restrict__ partial) {
  __shared__ float2 wsum[HEAD_WARPS][32];
  __shared__ float2 avg[32];
  const int n = blockIdx.x, chunk = blockIdx.y, nchunks = gridDim.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = chunk * HEAD_CHUNK + 2 * lane;
  const bool has1 = k < c, has2 = k + 1 < c;
  const T* f = feat + (size_t)n * p * ld + k;
  // fixed summation order: warp g sums positions g, g+8, ... ; warps are combined in order below
  float2 s = make_float2(0.f, 0.f);
  if (has1) {
    for (int i0 = warp; i0 < p; i0 += HEAD_WARPS * HEAD_MAX_PP) {
      float2 v[HEAD_MAX_PP];
#pragma unroll
      for (int u = 0; u < HEAD_MAX_PP; ++u) {
        const int i = i0 + u * HEAD_WARPS;
        v[u] = i < p ? head_load2<T, VEC2>(f + (size_t)i * ld, has2) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < HEAD_MAX_PP; ++u) {
        s.x += v[u].x;
        s.y += v[u].y;
      }
    }
  }
  wsum[warp][lane] = s;
  __syncthreads();
  if (warp == 0) {
    float2 a = make_float2(0.f, 0.f);
#pragma unroll
    for (int g = 0; g < HEAD_WARPS; ++g) {
      a.x += wsum[g][lane].x;
      a.y += wsum[g][lane].y;
    }
    avg[lane] = make_float2(a.x / (float)p, a.y / (float)p);
  }
  __syncthreads();
  const float2 av = avg[lane];
  float* pout = partial + ((size_t)n * nchunks + chunk) * ncls;
  for (int j0 = warp; j0 < ncls; j0 += HEAD_WARPS * 8) {  // eight classes per pass, loads overlapped
    float a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u * HEAD_WARPS;
      float2 wv = make_float2(0.f, 0.f);
      if (j < ncls && has1) wv = head_load2<float, VEC2>(w + (size_t)j * c + k, has2);
      a[u] = fmaf(wv.x, av.x, wv.y * av.y);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float r = ivf_warp_sum(a[u]);
      const int j = j0 + u * HEAD_WARPS;
      if (lane == 0 && j < ncls) pout[j] = r;
    }
  }
}

// Forward, stage 2: grid (clip): sum the chunk partials in a fixed order (deterministic), bias, softmax.
__global__ void __launch_bounds__(HEAD_THREADS)
head_fwd_finish_kernel(const float* __restrict__ partial, int nchunks, const float* __restrict__ b, int ncls,
                       int softmax, float* __restrict__ logits, float* __restrict__ out) {
  extern __shared__ float lg[];
  __shared__ float red[32];
  const int n = blockIdx.x;
  for (int j = threadIdx.x; j < ncls; j += blockDim.x) {
    float s = b ? b[j] : 0.f;
    for (int ch = 0; ch < nchunks; ++ch) s += partial[((size_t)n * nchunks + ch) * ncls + j];
    lg[j] = s;
    if (logits) logits[(size_t)n * ncls + j] = s;
  }
  __syncthreads();
  if (!softmax) {
    for (int j = threadIdx.x; j < ncls; j += blockDim.x) out[(size_t)n * ncls + j] = lg[j];
    return;
  }
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < ncls; j += blockDim.x) mx = fmaxf(mx, lg[j]);
  mx = block_reduce(mx, red, true);
  float se = 0.f;
  for (int j = threadIdx.x; j < ncls; j += blockDim.x) se += expf(lg[j] - mx);
  se = block_reduce(se, red, false);
  for (int j = threadIdx.x; j < ncls; j += blockDim.x)
    out[(size_t)n * ncls + j] = expf(lg[j] - mx) / se;
}

// Backward: grid (clip, channel chunk of HEAD_CHUNK).  Every block rebuilds dlogits (ncls values); then
// lane = channel pair, warp = class group: each thread has all its W loads in flight at once (the block
// reads just its [ncls][HEAD_CHUNK] slice of W); davg is combined over the warps in a fixed order and
// the block writes its channels of all p positions, warp = position group.
constexpr int HEAD_MAX_CLS = 24;  // classes per thread per pass

template <typename T, bool VEC2>
__global__ void __launch_bounds__(HEAD_THREADS)
head_bwd_kernel(int p, int c, int ld, const float* __restrict__ w, int ncls, int softmax,
                const float* __restrict__ out, const float* __restrict__ dout, int flags,
                const T* __restrict__ mask_y, int mask_ld, int mask_coff,
                const float* __restrict__ mask_scale, void* __restrict__ dfeat) {
  extern __shared__ float sm[];  // dlogit[ncls]
  __shared__ float red[32];
  __shared__ float2 part[HEAD_WARPS][32];
  __shared__ float2 davg_s[32];
  float* dl = sm;
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.y * HEAD_CHUNK + 2 * lane;
  const bool has1 = k < c, has2 = k + 1 < c;
  const float* o = out + (size_t)n * ncls;
  const float* g = dout + (size_t)n * ncls;
  if (softmax) {
    float dot = 0.f;
    for (int j = threadIdx.x; j < ncls; j += blockDim.x) dot += g[j] * o[j];
    dot = block_reduce(dot, red, false);
    for (int j = threadIdx.x; j < ncls; j += blockDim.x) dl[j] = o[j] * (g[j] - dot);
  } else {
    for (int j = threadIdx.x; j < ncls; j += blockDim.x) dl[j] = g[j];
  }
  __syncthreads();
  {
    float2 a = make_float2(0.f, 0.f);
    if (has1) {
      for (int j0 = warp; j0 < ncls; j0 += HEAD_WARPS * HEAD_MAX_CLS) {
        float2 wv[HEAD_MAX_CLS];
#pragma unroll
        for (int u = 0; u < HEAD_MAX_CLS; ++u) {
          const int j = j0 + u * HEAD_WARPS;
          wv[u] = j < ncls ? head_load2<float, VEC2>(w + (size_t)j * c + k, has2) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < HEAD_MAX_CLS; ++u) {
          const int j = j0 + u * HEAD_WARPS;
          const float d = j < ncls ? dl[j] : 0.f;
          a.x = fmaf(wv[u].x, d, a.x);
          a.y = fmaf(wv[u].y, d, a.y);
        }
      }
    }
    part[warp][lane] = a;
  }
  __syncthreads();
  if (warp == 0) {
    float2 a = make_float2(0.f, 0.f);
#pragma unroll
    for (int gq = 0; gq < HEAD_WARPS; ++gq) {
      a.x += part[gq][lane].x;
      a.y += part[gq][lane].y;
    }
    davg_s[lane] = make_float2(a.x / (float)p, a.y / (float)p);
  }
  __syncthreads();
  if (!has1) return;
  const float2 dv = davg_s[lane];
  float2 ms = make_float2(1.f, 1.f);
  if (flags & IVF_EP_MASK) ms = make_float2(mask_scale[k], has2 ? mask_scale[k + 1] : 1.f);
  const size_t base = (size_t)n * p;
  for (int i0 = warp; i0 < p; i0 += HEAD_WARPS * HEAD_MAX_PP) {
    float2 yv[HEAD_MAX_PP];
    if (flags & IVF_EP_MASK) {
#pragma unroll
      for (int u = 0; u < HEAD_MAX_PP; ++u) {
        const int i = i0 + u * HEAD_WARPS;
        yv[u] = i < p ? head_load2<T, VEC2>(mask_y + (base + i) * mask_ld + mask_coff + k, has2)
                      : make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < HEAD_MAX_PP; ++u) {
      const int i = i0 + u * HEAD_WARPS;
      if (i >= p) continue;
      float2 v = dv;
      if (flags & IVF_EP_MASK) {
        v.x = yv[u].x > 0.f ? v.x * ms.x : 0.f;
        v.y = yv[u].y > 0.f ? v.y * ms.y : 0.f;
      }
      const size_t idx = (base + i) * ld + k;
      if (flags & IVF_EP_OUT_F32) {
        float* d = reinterpret_cast<float*>(dfeat) + idx;
        if (VEC2) {
          *reinterpret_cast<float2*>(d) = v;
        } else {
          d[0] = v.x;
          if (has2) d[1] = v.y;
        }
      } else {
        T* d = reinterpret_cast<T*>(dfeat) + idx;
        if (VEC2 && sizeof(T) == 2) {
          *reinterpret_cast<__nv_bfloat162*>(d) = __floats2bfloat162_rn(v.x, v.y);
        } else if (VEC2) {
          *reinterpret_cast<float2*>(d) = v;
        } else {
          d[0] = ivf_from_float<T>(v.x);
          if (has2) d[1] = ivf_from_float<T>(v.y);
        }
      }
    }
  }
}

}  // namespace

extern "C" size_t ivf_i3d_head_workspace_bytes(int n, int c, int ncls) {
  if (n <= 0 || c <= 0 || ncls <= 0) return 0;
  return (size_t)n * ((c + HEAD_CHUNK - 1) / HEAD_CHUNK) * ncls * sizeof(float);
}

extern "C" int ivf_i3d_head_fwd(ivf_handle* h, int dtype, const void* feat, int n, int p, int c,
                                int ld, const float* w, const float* b, int ncls, int softmax,
                                float* logits, float* out, float* workspace, size_t workspace_bytes,
                                void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && feat && w && out && workspace, "ivf_i3d_head_fwd: null argument");
  IVF_REQUIRE(n > 0 && p > 0 && c > 0 && ncls > 0 && ld >= c, "ivf_i3d_head_fwd: bad extent");
  const int nchunks = (c + HEAD_CHUNK - 1) / HEAD_CHUNK;
  size_t need = (size_t)n * nchunks * ncls * sizeof(float);
  if (need > workspace_bytes)
    IVF_FAIL(IVF_EWORKSPACE, "ivf_i3d_head_fwd: workspace of %zu bytes, %zu needed", workspace_bytes, need);
  IVF_REQUIRE((size_t)ncls * sizeof(float) <= 48 * 1024, "ivf_i3d_head_fwd: ncls too large");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(n, nchunks);
  IVF_REQUIRE(dtype == IVF_F32 || dtype == IVF_BF16, "ivf_i3d_head_fwd: unknown dtype %d", dtype);
  const size_t pair_bytes = dtype == IVF_F32 ? 8 : 4;
  const bool vec2 = c % 2 == 0 && ld % 2 == 0 && (uintptr_t)feat % pair_bytes == 0 && (uintptr_t)w % 8 == 0;
#define IVF_HEAD_FWD(T, V) \
  head_fwd_partial_kernel<T, V><<<grid, HEAD_THREADS, 0, st>>>((const T*)feat, p, c, ld, w, ncls, workspace)
  if (dtype == IVF_F32) {
    if (vec2) IVF_HEAD_FWD(float, true); else IVF_HEAD_FWD(float, false);
  } else {
    if (vec2) IVF_HEAD_FWD(__nv_bfloat16, true); else IVF_HEAD_FWD(__nv_bfloat16, false);
  }
#undef IVF_HEAD_FWD
  IVF_LAUNCHED(h);
  head_fwd_finish_kernel<<<n, HEAD_THREADS, ncls * sizeof(float), st>>>(workspace, nchunks, b, ncls, softmax, logits,
                                                                        out);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_i3d_head_bwd(ivf_handle* h, int dtype, int n, int p, int c, int ld,
                                const float* w, int ncls, int softmax, const float* out,
                                const float* dout, int flags, const void* mask_y, int mask_ld,
                                int mask_coff, const float* mask_scale, void* dfeat, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && w && out && dout && dfeat, "ivf_i3d_head_bwd: null argument");
  IVF_REQUIRE(n > 0 && p > 0 && c > 0 && ncls > 0 && ld >= c, "ivf_i3d_head_bwd: bad extent");
  if (flags & IVF_EP_MASK) IVF_REQUIRE(mask_y && mask_scale, "ivf_i3d_head_bwd: MASK needs mask_y/mask_scale");
  size_t smem = (size_t)ncls * sizeof(float);
  IVF_REQUIRE(smem <= 40 * 1024, "ivf_i3d_head_bwd: ncls too large (%d)", ncls);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(n, (c + HEAD_CHUNK - 1) / HEAD_CHUNK);
  IVF_REQUIRE(dtype == IVF_F32 || dtype == IVF_BF16, "ivf_i3d_head_bwd: unknown dtype %d", dtype);
  const size_t pair_bytes = dtype == IVF_F32 ? 8 : 4;
  const size_t out_pair = (flags & IVF_EP_OUT_F32) ? 8 : pair_bytes;
  bool vec2 = c % 2 == 0 && ld % 2 == 0 && (uintptr_t)dfeat % out_pair == 0 && (uintptr_t)w % 8 == 0;
  if (flags & IVF_EP_MASK)
    vec2 = vec2 && mask_ld % 2 == 0 && mask_coff % 2 == 0 && (uintptr_t)mask_y % pair_bytes == 0;
#define IVF_HEAD_BWD(T, V)                                                                           \
  head_bwd_kernel<T, V><<<grid, HEAD_THREADS, smem, st>>>(p, c, ld, w, ncls, softmax, out, dout, flags, \
                                                          (const T*)mask_y, mask_ld, mask_coff, mask_scale, dfeat)
  if (dtype == IVF_F32) {
    if (vec2) IVF_HEAD_BWD(float, true); else IVF_HEAD_BWD(float, false);
  } else {
    if (vec2) IVF_HEAD_BWD(__nv_bfloat16, true); else IVF_HEAD_BWD(__nv_bfloat16, false);
  }
#undef IVF_HEAD_BWD
  IVF_LAUNCHED(h);
  return IVF_OK;
}
