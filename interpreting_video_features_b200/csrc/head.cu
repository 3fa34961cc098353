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

template <typename T>
__global__ void __launch_bounds__(HEAD_THREADS)
head_fwd_kernel(const T* __restrict__ feat, int p, int c, int ld, const float* __restrict__ w,
                const float* __restrict__ b, int ncls, int softmax, float* __restrict__ logits,
                float* __restrict__ out) {
  extern __shared__ float sm[];  // avg[c] | logit[ncls]
  __shared__ float red[32];
  float* avg = sm;
  float* lg = sm + c;
  const int n = blockIdx.x;
  const T* f = feat + (size_t)n * p * ld;
  const float inv = 1.f / (float)p;
  for (int k = threadIdx.x; k < c; k += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < p; ++i) s += ivf_to_float(f[(size_t)i * ld + k]);
    avg[k] = s * inv;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < ncls; j += nw) {
    const float* wr = w + (size_t)j * c;
    float s = 0.f;
    for (int k = lane; k < c; k += 32) s = fmaf(__ldg(wr + k), avg[k], s);
    s = ivf_warp_sum(s);
    if (lane == 0) lg[j] = s + (b ? b[j] : 0.f);
  }
  __syncthreads();
  if (logits)
    for (int j = threadIdx.x; j < ncls; j += blockDim.x) logits[(size_t)n * ncls + j] = lg[j];
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

template <typename T>
__global__ void __launch_bounds__(HEAD_THREADS)
head_bwd_kernel(int p, int c, int ld, const float* __restrict__ w, int ncls, int softmax,
                const float* __restrict__ out, const float* __restrict__ dout, int flags,
                const T* __restrict__ mask_y, int mask_ld, int mask_coff,
                const float* __restrict__ mask_scale, void* __restrict__ dfeat) {
  extern __shared__ float sm[];  // dlogit[ncls] | davg[c]
  __shared__ float red[32];
  float* dl = sm;
  float* davg = sm + ncls;
  const int n = blockIdx.x;
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
  const float inv = 1.f / (float)p;
  for (int k = threadIdx.x; k < c; k += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < ncls; ++j) s = fmaf(__ldg(w + (size_t)j * c + k), dl[j], s);
    davg[k] = s * inv;
  }
  __syncthreads();
  const size_t base = (size_t)n * p;
  for (int e = threadIdx.x; e < p * c; e += blockDim.x) {
    int i = e / c, k = e - i * c;
    float v = davg[k];
    if (flags & IVF_EP_MASK) {
      float y = ivf_to_float(mask_y[(base + i) * mask_ld + mask_coff + k]);
      v = y > 0.f ? v * mask_scale[k] : 0.f;
    }
    size_t idx = (base + i) * ld + k;
    if (flags & IVF_EP_OUT_F32)
      reinterpret_cast<float*>(dfeat)[idx] = v;
    else
      reinterpret_cast<T*>(dfeat)[idx] = ivf_from_float<T>(v);
  }
}

}  // namespace

extern "C" int ivf_i3d_head_fwd(ivf_handle* h, int dtype, const void* feat, int n, int p, int c,
                                int ld, const float* w, const float* b, int ncls, int softmax,
                                float* logits, float* out, void* stream) {
  IVF_REQUIRE(h && feat && w && out, "ivf_i3d_head_fwd: null argument");
  IVF_REQUIRE(n > 0 && p > 0 && c > 0 && ncls > 0 && ld >= c, "ivf_i3d_head_fwd: bad extent");
  size_t smem = (size_t)(c + ncls) * sizeof(float);
  IVF_REQUIRE(smem <= 48 * 1024, "ivf_i3d_head_fwd: c + ncls too large (%d + %d)", c, ncls);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IVF_F32)
    head_fwd_kernel<float><<<n, HEAD_THREADS, smem, st>>>((const float*)feat, p, c, ld, w, b, ncls,
                                                          softmax, logits, out);
  else if (dtype == IVF_BF16)
    head_fwd_kernel<__nv_bfloat16><<<n, HEAD_THREADS, smem, st>>>((const __nv_bfloat16*)feat, p, c,
                                                                  ld, w, b, ncls, softmax, logits, out);
  else
    IVF_FAIL(IVF_EINVAL, "ivf_i3d_head_fwd: unknown dtype %d", dtype);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_i3d_head_bwd(ivf_handle* h, int dtype, int n, int p, int c, int ld,
                                const float* w, int ncls, int softmax, const float* out,
                                const float* dout, int flags, const void* mask_y, int mask_ld,
                                int mask_coff, const float* mask_scale, void* dfeat, void* stream) {
  IVF_REQUIRE(h && w && out && dout && dfeat, "ivf_i3d_head_bwd: null argument");
  IVF_REQUIRE(n > 0 && p > 0 && c > 0 && ncls > 0 && ld >= c, "ivf_i3d_head_bwd: bad extent");
  if (flags & IVF_EP_MASK) IVF_REQUIRE(mask_y && mask_scale, "ivf_i3d_head_bwd: MASK needs mask_y/mask_scale");
  size_t smem = (size_t)(c + ncls) * sizeof(float);
  IVF_REQUIRE(smem <= 48 * 1024, "ivf_i3d_head_bwd: c + ncls too large (%d + %d)", c, ncls);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IVF_F32)
    head_bwd_kernel<float><<<n, HEAD_THREADS, smem, st>>>(p, c, ld, w, ncls, softmax, out, dout, flags,
                                                          (const float*)mask_y, mask_ld, mask_coff,
                                                          mask_scale, dfeat);
  else if (dtype == IVF_BF16)
    head_bwd_kernel<__nv_bfloat16><<<n, HEAD_THREADS, smem, st>>>(
        p, c, ld, w, ncls, softmax, out, dout, flags, (const __nv_bfloat16*)mask_y, mask_ld,
        mask_coff, mask_scale, dfeat);
  else
    IVF_FAIL(IVF_EINVAL, "ivf_i3d_head_bwd: unknown dtype %d", dtype);
  IVF_LAUNCHED(h);
  return IVF_OK;
}
