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
  if (k0 + kl < c) {
    // fixed summation order (positions half, half+2, ...), four independent partial sums so the
    // loads are in flight together
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const T* fp = f + k0 + kl;
    int i = half;
    for (; i + 6 < p; i += 8) {
      s0 += ivf_to_float(fp[(size_t)i * ld]);
      s1 += ivf_to_float(fp[(size_t)(i + 2) * ld]);
      s2 += ivf_to_float(fp[(size_t)(i + 4) * ld]);
      s3 += ivf_to_float(fp[(size_t)(i + 6) * ld]);
    }
    for (; i < p; i += 2) s0 += ivf_to_float(fp[(size_t)i * ld]);
    s = (s0 + s1) + (s2 + s3);
  }
  half_sum[half][kl] = s;
  __syncthreads();
  if (threadIdx.x < HEAD_CHUNK) avg[kl] = (half_sum[0][kl] + half_sum[1][kl]) / (float)p;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < ncls; j += 2 * nw) {  // two classes per pass: their loads overlap
    const int j2 = j + nw;
    const float* wr = w + (size_t)j * c + k0;
    const float* wr2 = w + (size_t)(j2 < ncls ? j2 : j) * c + k0;
    float a = 0.f, a2 = 0.f;
    for (int k = lane; k < HEAD_CHUNK && k0 + k < c; k += 32) {
      a = fmaf(__ldg(wr + k), avg[k], a);
      a2 = fmaf(__ldg(wr2 + k), avg[k], a2);
    }
    a = ivf_warp_sum(a);
    a2 = ivf_warp_sum(a2);
    if (lane == 0) {
      partial[((size_t)n * nchunks + chunk) * ncls + j] = a;
      if (j2 < ncls) partial[((size_t)n * nchunks + chunk) * ncls + j2] = a2;
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

// Backward: grid (clip, channel chunk of HEAD_CHUNK).  Every block rebuilds dlogits (ncls values), then
// the 256 threads = HEAD_CHUNK channels x 2 class halves form davg for the block's channels only (so a
// block reads just its [ncls][HEAD_CHUNK] slice of W, eight independent loads in flight per thread) and
// write those channels of all p positions.
template <typename T>
__global__ void __launch_bounds__(HEAD_THREADS)
head_bwd_kernel(int p, int c, int ld, const float* __restrict__ w, int ncls, int softmax,
                const float* __restrict__ out, const float* __restrict__ dout, int flags,
                const T* __restrict__ mask_y, int mask_ld, int mask_coff,
                const float* __restrict__ mask_scale, void* __restrict__ dfeat) {
  extern __shared__ float sm[];  // dlogit[ncls]
  __shared__ float red[32];
  __shared__ float part[2][HEAD_CHUNK];
  __shared__ float davg[HEAD_CHUNK], mscale[HEAD_CHUNK];
  float* dl = sm;
  const int n = blockIdx.x;
  const int k0 = blockIdx.y * HEAD_CHUNK;
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
  const int kl = threadIdx.x % HEAD_CHUNK, half = threadIdx.x / HEAD_CHUNK;
  const int k = k0 + kl;
  {
    const int jh = (ncls + 1) / 2;
    const int j0 = half * jh, j1 = min(ncls, j0 + jh);
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    if (k < c) {
      const float* wk = w + k;
      int j = j0;
      for (; j + 8 <= j1; j += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fmaf(__ldg(wk + (size_t)(j + u) * c), dl[j + u], acc[u]);
      }
      for (; j < j1; ++j) acc[0] = fmaf(__ldg(wk + (size_t)j * c), dl[j], acc[0]);
    }
    part[half][kl] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  }
  __syncthreads();
  if (threadIdx.x < HEAD_CHUNK) {
    davg[kl] = (part[0][kl] + part[1][kl]) / (float)p;
    mscale[kl] = ((flags & IVF_EP_MASK) && k < c) ? mask_scale[k] : 1.f;
  }
  __syncthreads();
  const size_t base = (size_t)n * p;
  const int kc = min(HEAD_CHUNK, c - k0);
  for (int e = threadIdx.x; e < p * HEAD_CHUNK; e += blockDim.x) {
    const int i = e / HEAD_CHUNK, kk = e % HEAD_CHUNK;
    if (kk >= kc) continue;
    float v = davg[kk];
    if (flags & IVF_EP_MASK) {
      float y = ivf_to_float(mask_y[(base + i) * mask_ld + mask_coff + k0 + kk]);
      v = y > 0.f ? v * mscale[kk] : 0.f;
    }
    size_t idx = (base + i) * ld + k0 + kk;
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
  size_t smem = (size_t)ncls * sizeof(float);
  IVF_REQUIRE(smem <= 40 * 1024, "ivf_i3d_head_bwd: ncls too large (%d)", ncls);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(n, (c + HEAD_CHUNK - 1) / HEAD_CHUNK);
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
