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

constexpr int HEAD_CHUNK = 128;  // channels per forward block

// Forward, stage 1: grid (clip, channel chunk).  256 threads = 128 channels x 2 position halves;
// average-pool the chunk, then its partial logits -> scratch[clip][chunk][class].
template <typename T>
__global__ void __launch_bounds__(HEAD_THREADS)
head_fwd_partial_kernel(const T* __restrict__ feat, int p, int c, int ld, const float* __restrict__ w,
                        int ncls, float* __restrict__ partial) {
  __shared__ float half_sum[2][HEAD_CHUNK];
  __shared__ float avg[HEAD_CHUNK];
  const int n = blockIdx.x, chunk = blockIdx.y, nchunks = gridDim.y;
  const int k0 = chunk * HEAD_CHUNK;
  const int kl = threadIdx.x % HEAD_CHUNK, half = threadIdx.x / HEAD_CHUNK;
  const T* f = feat + (size_t)n * p * ld;
  float s = 0.f;
  if (k0 + kl < c)
    for (int i = half; i < p; i += 2) s += ivf_to_float(f[(size_t)i * ld + k0 + kl]);
  half_sum[half][kl] = s;
  __syncthreads();
  if (threadIdx.x < HEAD_CHUNK) avg[kl] = (half_sum[0][kl] + half_sum[1][kl]) / (float)p;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < ncls; j += nw) {
    const float* wr = w + (size_t)j * c + k0;
    float a = 0.f;
    for (int k = lane; k < HEAD_CHUNK && k0 + k < c; k += 32) a = fmaf(__ldg(wr + k), avg[k], a);
    a = ivf_warp_sum(a);
    if (lane == 0) partial[((size_t)n * nchunks + chunk) * ncls + j] = a;
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

// Backward: grid (clip, position chunk).  Every block rebuilds dlogits (ncls values) and the channel
// gradient of its HEAD_CHUNK..c range is recomputed per block (ncls*c MACs, cheap) so that the p*c
// output elements are written by many blocks instead of one.
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
  const int pchunks = gridDim.y;
  const int per = (p + pchunks - 1) / pchunks;
  const int i0 = blockIdx.y * per, i1 = min(p, i0 + per);
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
    float s0 = 0.f, s1 = 0.f;
    int j = 0;
    for (; j + 1 < ncls; j += 2) {
      s0 = fmaf(__ldg(w + (size_t)j * c + k), dl[j], s0);
      s1 = fmaf(__ldg(w + (size_t)(j + 1) * c + k), dl[j + 1], s1);
    }
    if (j < ncls) s0 = fmaf(__ldg(w + (size_t)j * c + k), dl[j], s0);
    davg[k] = (s0 + s1) * inv;
  }
  __syncthreads();
  const size_t base = (size_t)n * p;
  for (int e = threadIdx.x; e < (i1 - i0) * c; e += blockDim.x) {
    int i = i0 + e / c, k = e % c;
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
  const int nchunks = (c + HEAD_CHUNK - 1) / HEAD_CHUNK;
  size_t need = (size_t)n * nchunks * ncls * sizeof(float);
  IVF_REQUIRE(need <= h->scratch_bytes, "ivf_i3d_head_fwd: n*c*ncls too large for the handle scratch (%zu B)", need);
  IVF_REQUIRE((size_t)ncls * sizeof(float) <= 48 * 1024, "ivf_i3d_head_fwd: ncls too large");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(n, nchunks);
  if (dtype == IVF_F32)
    head_fwd_partial_kernel<float><<<grid, HEAD_THREADS, 0, st>>>((const float*)feat, p, c, ld, w, ncls, h->scratch);
  else if (dtype == IVF_BF16)
    head_fwd_partial_kernel<__nv_bfloat16><<<grid, HEAD_THREADS, 0, st>>>((const __nv_bfloat16*)feat, p, c, ld, w,
                                                                          ncls, h->scratch);
  else
    IVF_FAIL(IVF_EINVAL, "ivf_i3d_head_fwd: unknown dtype %d", dtype);
  IVF_LAUNCHED(h);
  head_fwd_finish_kernel<<<n, HEAD_THREADS, ncls * sizeof(float), st>>>(h->scratch, nchunks, b, ncls, softmax, logits,
                                                                        out);
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
  int pchunks = p >= 14 ? (p + 13) / 14 : 1;
  dim3 grid(n, pchunks);
  if (dtype == IVF_F32)
    head_bwd_kernel<float><<<grid, HEAD_THREADS, smem, st>>>(p, c, ld, w, ncls, softmax, out, dout, flags,
                                                          (const float*)mask_y, mask_ld, mask_coff,
                                                          mask_scale, dfeat);
  else if (dtype == IVF_BF16)
    head_bwd_kernel<__nv_bfloat16><<<grid, HEAD_THREADS, smem, st>>>(
        p, c, ld, w, ncls, softmax, out, dout, flags, (const __nv_bfloat16*)mask_y, mask_ld,
        mask_coff, mask_scale, dfeat);
  else
    IVF_FAIL(IVF_EINVAL, "ivf_i3d_head_bwd: unknown dtype %d", dtype);
  IVF_LAUNCHED(h);
  return IVF_OK;
}
