// Fused Grad-CAM tail: channel-pooled gradient weights -> weighted activation sum -> ReLU ->
// bilinear upsample (cv2.resize INTER_LINEAR, half-pixel centres, replicated border) ->
// temporal repeat -> min/max normalisation.  One kernel; the reference does this on the host
// in numpy/cv2 after three device->host copies (pt/grad_cam_videos.py:85-140: weights = mean of
// the gradient over (t,h,w) :98, 1024-iteration Python channel loop :101-108, np.maximum :110,
// cv2.resize + np.repeat per feature-time slice :116-125, per-slice or global normalisation
// :129-135 — an all-zero slice gives 0/0 = NaN there and here).
// Grid = (clip, slice) when normalising per slice, (clip, 1) when normalising per clip.
#include "common.cuh"

namespace {

constexpr int GC_THREADS = 256;

__device__ __forceinline__ float bilinear(const float* __restrict__ lr, int hp, int wp, int oy, int ox,
                                          double sy, double sx) {
  // cv2 resize (float path): fx = (float)((dx + 0.5) * scale - 0.5); sx = floor(fx); fx -= sx;
  // clamp: sx < 0 -> sx = 0, fx = 0; sx >= src-1 -> sx = src-1, fx = 0 (second tap = same pixel)
  float fy = (float)((oy + 0.5) * sy - 0.5);
  int y0 = (int)floorf(fy);
  fy -= (float)y0;
  if (y0 < 0) {
    y0 = 0;
    fy = 0.f;
  }
  if (y0 >= hp - 1) {
    y0 = hp - 1;
    fy = 0.f;
  }
  int y1 = min(y0 + 1, hp - 1);
  float fx = (float)((ox + 0.5) * sx - 0.5);
  int x0 = (int)floorf(fx);
  fx -= (float)x0;
  if (x0 < 0) {
    x0 = 0;
    fx = 0.f;
  }
  if (x0 >= wp - 1) {
    x0 = wp - 1;
    fx = 0.f;
  }
  int x1 = min(x0 + 1, wp - 1);
  // horizontal pass on both rows, then vertical (cv2's HResize then VResize order)
  float r0 = lr[y0 * wp + x0] * (1.f - fx) + lr[y0 * wp + x1] * fx;
  float r1 = lr[y1 * wp + x0] * (1.f - fx) + lr[y1 * wp + x1] * fx;
  return r0 * (1.f - fy) + r1 * fy;
}

template <typename T, typename TG>
__global__ void __launch_bounds__(GC_THREADS)
gradcam_kernel(const T* __restrict__ act, const TG* __restrict__ grad, int tp, int hp, int wp, int c,
               int ld, int step, int hout, int wout, int per_frame, float* __restrict__ cam,
               float* __restrict__ cam_lowres) {
  extern __shared__ float sm[];  // w[c] | lowres[slices*hp*wp]
  __shared__ float red_min[32], red_max[32];
  float* wk = sm;
  float* lr = sm + c;
  const int n = blockIdx.x;
  const int s0 = per_frame ? blockIdx.y : 0;
  const int s1 = per_frame ? blockIdx.y + 1 : tp;
  const int pp = hp * wp;
  const int pall = tp * pp;
  const TG* g = grad + (size_t)n * pall * ld;
  const T* a = act + (size_t)n * pall * ld;
  // 1. channel weights: mean of the gradient over every position (time AND space)
  const float inv = 1.f / (float)pall;
  for (int k = threadIdx.x; k < c; k += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < pall; ++i) s += ivf_to_float(g[(size_t)i * ld + k]);
    wk[k] = s * inv;
  }
  __syncthreads();
  // 2. low-resolution CAM of this block's slices: relu(sum_k w_k A_k), one warp per position
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int pos = s0 * pp + warp; pos < s1 * pp; pos += nw) {
    const T* ap = a + (size_t)pos * ld;
    float s = 0.f;
    for (int k = lane; k < c; k += 32) s = fmaf(wk[k], ivf_to_float(ap[k]), s);
    s = ivf_warp_sum(s);
    if (lane == 0) {
      s = fmaxf(s, 0.f);
      lr[pos - s0 * pp] = s;
      if (cam_lowres) cam_lowres[(size_t)n * pall + pos] = s;
    }
  }
  __syncthreads();
  if (!cam) return;  // low-resolution map only
  // 3. min / max of the upsampled maps (all of this block's slices)
  const double sy = (double)hp / (double)hout, sx = (double)wp / (double)wout;
  const int npix = hout * wout;
  float mn = INFINITY, mx = -INFINITY;
  for (int sl = s0; sl < s1; ++sl) {
    const float* l = lr + (sl - s0) * pp;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
      float v = bilinear(l, hp, wp, i / wout, i % wout, sy, sx);
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
  }
  mn = ivf_warp_min(mn);
  mx = ivf_warp_max(mx);
  if (lane == 0) {
    red_min[warp] = mn;
    red_max[warp] = mx;
  }
  __syncthreads();
  mn = INFINITY;
  mx = -INFINITY;
  for (int i = 0; i < nw; ++i) {
    mn = fminf(mn, red_min[i]);
    mx = fmaxf(mx, red_max[i]);
  }
  const float range = mx - mn;  // reference: x -= min; x /= max(x)
  // 4. write normalised, temporally repeated output
  for (int sl = s0; sl < s1; ++sl) {
    const float* l = lr + (sl - s0) * pp;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
      float v = (bilinear(l, hp, wp, i / wout, i % wout, sy, sx) - mn) / range;
      for (int r = 0; r < step; ++r)
        cam[((size_t)n * tp * step + (size_t)sl * step + r) * npix + i] = v;
    }
  }
}

}  // namespace

extern "C" int ivf_gradcam(ivf_handle* h, int act_dtype, int grad_dtype, const void* act,
                           const void* grad, int n, int tp, int hp, int wp, int c, int ld, int step,
                           int hout, int wout, int per_frame, float* cam, float* cam_lowres,
                           void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && act && grad && (cam || cam_lowres), "ivf_gradcam: null argument");
  IVF_REQUIRE(n > 0 && tp > 0 && hp > 0 && wp > 0 && c > 0 && ld >= c && step > 0 && hout > 0 && wout > 0,
              "ivf_gradcam: bad extent");
  int slices = per_frame ? 1 : tp;
  size_t smem = ((size_t)c + (size_t)slices * hp * wp) * sizeof(float);
  IVF_REQUIRE(smem <= 48 * 1024, "ivf_gradcam: c + map too large for shared memory");
  dim3 grid(n, per_frame ? tp : 1);
  cudaStream_t st = (cudaStream_t)stream;
#define IVF_GC_LAUNCH(TA, TG)                                                                     \
  gradcam_kernel<TA, TG><<<grid, GC_THREADS, smem, st>>>((const TA*)act, (const TG*)grad, tp, hp, wp, \
                                                         c, ld, step, hout, wout, per_frame, cam,    \
                                                         cam_lowres)
  if (act_dtype == IVF_F32 && grad_dtype == IVF_F32)
    IVF_GC_LAUNCH(float, float);
  else if (act_dtype == IVF_BF16 && grad_dtype == IVF_F32)
    IVF_GC_LAUNCH(__nv_bfloat16, float);
  else if (act_dtype == IVF_BF16 && grad_dtype == IVF_BF16)
    IVF_GC_LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else
    IVF_FAIL(IVF_EINVAL, "ivf_gradcam: unsupported dtypes act %d grad %d", act_dtype, grad_dtype);
#undef IVF_GC_LAUNCH
  IVF_LAUNCHED(h);
  return IVF_OK;
}
