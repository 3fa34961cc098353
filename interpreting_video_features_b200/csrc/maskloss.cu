// Mask objective and optimiser step, one launch for all clips.
// Reference: pt/FindMasksComparison_I3D_smth.py:191-214 —
//   mask_clip = sigmoid(time_mask); l1 = lam1*sum|mask_clip|; tv = lam2*calc_tv_norm(mask_clip,3,3);
//   loss = l1 + tv + class_loss; Adam([time_mask], lr=0.2).step()
// and pt/mask.py:88-100 calc_tv_norm: val = sum_{u=1}^{T-2} |m[u-1]-m[u]|^p + |m[u+1]-m[u]|^p,
// then (val^(1/p))^q.  d/dval of the pow chain is (q/p) val^(q/p-1), evaluated the way autograd
// does (two pow backward nodes), so a constant mask (val == 0) yields NaN exactly as the
// reference does (pt/mask.py:163-165 works around it).
// torch.optim.Adam semantics (no amsgrad, no weight decay): step_size = lr/(1-b1^t),
// denom = sqrt(v)/sqrt(1-b2^t) + eps.
#include "common.cuh"

namespace {

constexpr int MAX_T = 1024;

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = ivf_warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < nw; ++i) r += red[i];
  __syncthreads();
  return r;
}

// TV value and d(TV)/d(s[u]) for general p, q from s[] in shared memory.
// Pair (i,i+1) appears twice in the reference's double sum except the first and last pair.
__device__ __forceinline__ float pair_weight(int i, int t) { return (i == 0 || i == t - 2) ? 1.f : 2.f; }

__global__ void mask_loss_adam_kernel(float* __restrict__ m, float* __restrict__ exp_avg,
                                      float* __restrict__ exp_avg_sq,
                                      const float* __restrict__ dclass, int t, int step_arg,
                                      int* __restrict__ step_dev, float lam1,
                                      float lam2, float lr, float beta1, float beta2, float eps,
                                      float* __restrict__ losses, float* __restrict__ sig_out) {
  __shared__ float s[MAX_T];
  __shared__ float red[32];
  const int clip = blockIdx.x;
  float* mm = m + (size_t)clip * t;
  for (int u = threadIdx.x; u < t; u += blockDim.x) s[u] = sigmoidf_(mm[u]);
  __syncthreads();
  float l1 = 0.f, tv = 0.f;
  for (int u = threadIdx.x; u < t; u += blockDim.x) {
    l1 += fabsf(s[u]);
    if (u + 1 < t && t >= 3) {
      float dlt = fabsf(s[u + 1] - s[u]);
      tv += pair_weight(u, t) * dlt * dlt * dlt;
    }
  }
  l1 = block_sum(l1, red);
  tv = block_sum(tv, red);
  // autograd of (val^(1/3))^3: d = 3*(val^(1/3))^2 * (1/3)*val^(1/3-1)
  const float r = powf(tv, 1.f / 3.f);
  const float chain = (3.f * r * r) * ((1.f / 3.f) * powf(tv, 1.f / 3.f - 1.f));
  const float tv_val = r * r * r;
  if (threadIdx.x == 0 && losses) {
    losses[clip * 3 + 0] = lam1 * l1;
    losses[clip * 3 + 1] = lam2 * tv_val;
    losses[clip * 3 + 2] = lam1 * l1 + lam2 * tv_val;
  }
  // step counter: per-clip device counter (CUDA-graph replay safe) or the host argument
  const int step = step_dev ? step_dev[clip] + 1 : step_arg;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  for (int u = threadIdx.x; u < t; u += blockDim.x) {
    float su = s[u];
    float dtv = 0.f;
    if (t >= 3) {
      if (u >= 1) {
        float dlt = su - s[u - 1];
        dtv += pair_weight(u - 1, t) * 3.f * dlt * fabsf(dlt);
      }
      if (u + 1 < t) {
        float dlt = s[u + 1] - su;
        dtv -= pair_weight(u, t) * 3.f * dlt * fabsf(dlt);
      }
    }
    float sgn = su > 0.f ? 1.f : (su < 0.f ? -1.f : 0.f);
    float gs = lam1 * sgn + lam2 * chain * dtv + (dclass ? dclass[(size_t)clip * t + u] : 0.f);
    float g = gs * su * (1.f - su);
    size_t i = (size_t)clip * t + u;
    float ea = beta1 * exp_avg[i] + (1.f - beta1) * g;
    float es = beta2 * exp_avg_sq[i] + (1.f - beta2) * g * g;
    exp_avg[i] = ea;
    exp_avg_sq[i] = es;
    float denom = sqrtf(es) / sqrtf(bc2) + eps;
    float nm = mm[u] - (lr / bc1) * (ea / denom);
    mm[u] = nm;
    if (sig_out) sig_out[i] = sigmoidf_(nm);
  }
  __syncthreads();
  if (step_dev && threadIdx.x == 0) step_dev[clip] = step;
}

__global__ void sigmoid_kernel(const float* __restrict__ m, float* __restrict__ out, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = sigmoidf_(m[i]);
}

__global__ void tv_norm_kernel(const float* __restrict__ mask, int t, float p, float q,
                               float* __restrict__ val, float* __restrict__ dmask) {
  __shared__ float s[MAX_T];
  __shared__ float red[32];
  for (int u = threadIdx.x; u < t; u += blockDim.x) s[u] = mask[u];
  __syncthreads();
  float tv = 0.f;
  for (int u = threadIdx.x; u + 1 < t; u += blockDim.x)
    if (t >= 3) tv += pair_weight(u, t) * powf(fabsf(s[u + 1] - s[u]), p);
  tv = block_sum(tv, red);
  const float r = powf(tv, 1.f / p);
  const float out = powf(r, q);
  const float chain = (q * powf(r, q - 1.f)) * ((1.f / p) * powf(tv, 1.f / p - 1.f));
  if (threadIdx.x == 0) val[0] = out;
  if (dmask) {
    for (int u = threadIdx.x; u < t; u += blockDim.x) {
      float dtv = 0.f;
      if (t >= 3) {
        if (u >= 1) {
          float dlt = s[u] - s[u - 1];
          float sg = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
          dtv += pair_weight(u - 1, t) * p * powf(fabsf(dlt), p - 1.f) * sg;
        }
        if (u + 1 < t) {
          float dlt = s[u + 1] - s[u];
          float sg = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
          dtv -= pair_weight(u, t) * p * powf(fabsf(dlt), p - 1.f) * sg;
        }
      }
      dmask[u] = chain * dtv;
    }
  }
}

// scores[row][clip] = probs[clip][targets[clip]]: the class score the drivers read after every forward
// (pt/mask.py:128-129,140-143 `model(...)[batch_index, target[batch_index]]`), kept on the device
__global__ void select_scores_kernel(const float* __restrict__ probs, const int* __restrict__ targets, int n, int ncls,
                                     float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = probs[(size_t)i * ncls + targets[i]];
}

// pt/mask.py:121-154 'central' initialisation for n clips at once, from the scores of all candidates:
// scores[0] = unperturbed clip, scores[1] = fully frozen clip, scores[1 + i] (i = 1 .. T/2-1) = centred window
// with i frames off at both ends.  The reference shrinks the window while (orig - central)/(orig - frozen) >=
// threshold and keeps the first candidate that fails (or the last tried); 0 -> -5, 1 -> +5.  A NaN ratio
// (orig == frozen) never compares below the threshold, as in the reference.
__global__ void init_mask_select_kernel(const float* __restrict__ scores, int n, int t, float threshold,
                                        float* __restrict__ raw, int* __restrict__ chosen_out) {
  const int b = blockIdx.x;
  __shared__ int chosen;
  if (threadIdx.x == 0) {
    const float orig = scores[b], frozen = scores[n + b];
    int c = 0;  // 0 = no candidate tried: the mask stays all ones (T < 4)
    for (int i = 1; i < t / 2; ++i) {
      c = i;
      const float ratio = (orig - scores[(size_t)(1 + i) * n + b]) / (orig - frozen);
      if (ratio < threshold) break;
    }
    chosen = c;
    if (chosen_out) chosen_out[b] = c;
  }
  __syncthreads();
  const int c = chosen;
  for (int u = threadIdx.x; u < t; u += blockDim.x) raw[(size_t)b * t + u] = (u < c || u >= t - c) ? -5.f : 5.f;
}

}  // namespace

extern "C" int ivf_select_scores(ivf_handle* h, const float* probs, const int* targets, int n, int ncls, float* out,
                                 void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && probs && targets && out && n > 0 && ncls > 0, "ivf_select_scores: bad argument");
  select_scores_kernel<<<ivf_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(probs, targets, n, ncls, out);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_init_mask_select(ivf_handle* h, const float* scores, int n, int t, float threshold, float* raw,
                                    int* chosen, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && scores && raw && n > 0 && t > 0 && t <= MAX_T, "ivf_init_mask_select: bad argument");
  init_mask_select_kernel<<<n, 32, 0, (cudaStream_t)stream>>>(scores, n, t, threshold, raw, chosen);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_mask_loss_adam(ivf_handle* h, float* m, float* exp_avg, float* exp_avg_sq,
                                  const float* dclass, int nclip, int t, int step, int* step_dev,
                                  float lam1,
                                  float lam2, float lr, float beta1, float beta2, float eps,
                                  float* losses, float* sig_out, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && m && exp_avg && exp_avg_sq, "ivf_mask_loss_adam: null argument");
  IVF_REQUIRE(nclip > 0 && t > 0 && t <= MAX_T && (step >= 1 || step_dev),
              "ivf_mask_loss_adam: bad nclip/t/step");
  int threads = t <= 32 ? 32 : (t <= 64 ? 64 : 128);
  mask_loss_adam_kernel<<<nclip, threads, 0, (cudaStream_t)stream>>>(
      m, exp_avg, exp_avg_sq, dclass, t, step, step_dev, lam1, lam2, lr, beta1, beta2, eps, losses,
      sig_out);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_sigmoid(ivf_handle* h, const float* m, float* out, int count, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && m && out && count > 0, "ivf_sigmoid: bad argument");
  sigmoid_kernel<<<ivf_cdiv(count, 256), 256, 0, (cudaStream_t)stream>>>(m, out, count);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_tv_norm(ivf_handle* h, const float* mask, int t, float p, float q, float* val,
                           float* dmask, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && mask && val, "ivf_tv_norm: null argument");
  IVF_REQUIRE(t > 0 && t <= MAX_T, "ivf_tv_norm: bad t");
  tv_norm_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(mask, t, p, q, val, dmask);
  IVF_LAUNCHED(h);
  return IVF_OK;
}
