// Temporal perturbation of a clip under a mask, forward and backward-to-the-mask.
// Reference: pt/mask.py:4-56 perturb_sequence —
//   'freeze'  P[0] = x[0];  P[u] = (1-m[u]) x[u] + m[u] P[u-1]          (:13-22, T-1 tensor clones)
//   'reverse' inside every maximal run of m > 0.1 (find_submasks_from_mask, :60-85) the u-th and
//             u-th-last frames are blended with each other using the FRONT frame's mask value
//             on both sides; the middle frame of an odd run and everything else is copied (:24-56)
// and the autograd of both w.r.t. the mask:
//   freeze   dm[u] = sum_px G[u] (P[u-1] - x[u]),  G[u] = g[u] + m[u+1] G[u+1]
//   reverse  dm[i] = sum_px (g[i] - g[j]) (x[j] - x[i])   for a pair (i front, j back)
// One fused scan kernel each way (the reference touches the clip O(T) times per call);
// bandwidth-bound, coalesced along W, warp-shuffle + shared reduction to dm[T].
//
// Output formats (ivf.h IVF_PFMT_*): NCDHW fp32 (drop-in result), NDHWC fp32 (fp32 conv path),
// space-to-depth bf16 [b][t/2][h/2][w/2][32] (operand of the stem convolution, whose stride 2
// becomes a stride-1 4x4x4 convolution over 8*c channels).
#include "common.cuh"

namespace {

constexpr int MAX_T = 64;
constexpr int MAX_C = 4;

struct MaskInfo {
  float m[MAX_T];      // mask values of this clip
  float coef[MAX_T];   // reverse: blend factor of frame u (0 = copied)
  int partner[MAX_T];  // reverse: the frame u is blended with (u itself if copied)
  int front[MAX_T];    // reverse: 1 if u is the front frame of its pair
};

// thread 0..: build MaskInfo for one clip in shared memory
__device__ void build_mask_info(MaskInfo* mi, const float* __restrict__ mask, int t, int mode) {
  for (int u = threadIdx.x; u < t; u += blockDim.x) {
    mi->m[u] = mask[u];
    mi->coef[u] = 0.f;
    mi->partner[u] = u;
    mi->front[u] = 0;
  }
  __syncthreads();
  if (mode == 1 && threadIdx.x == 0) {
    // pt/mask.py:60-85: runs of mask > 0.1 (strict), pairs (run[k], run[len-1-k]) for k < len/2
    int u = 0;
    while (u < t) {
      if (mi->m[u] > 0.1f) {
        int e = u;
        while (e + 1 < t && mi->m[e + 1] > 0.1f) ++e;
        int len = e - u + 1;
        for (int k = 0; k < len / 2; ++k) {
          int i = u + k, j = e - k;
          float a = mi->m[i];
          mi->coef[i] = a;
          mi->coef[j] = a;
          mi->partner[i] = j;
          mi->partner[j] = i;
          mi->front[i] = 1;
        }
        u = e + 1;
      } else {
        ++u;
      }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------ forward, planar outputs
// one thread per (pixel hw, channel), grid.y = clip
template <int FMT>
__global__ void perturb_fwd_planar_kernel(int mode, const float* __restrict__ x,
                                          const float* __restrict__ mask, int mask_bstride, int c,
                                          int t, int hw, float* __restrict__ out) {
  __shared__ MaskInfo mi;
  const int b = blockIdx.y;
  build_mask_info(&mi, mask + (size_t)b * mask_bstride, t, mode);
  const int total = c * hw;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ch = idx / hw, p = idx - ch * hw;
    const float* xs = x + ((size_t)(b * c + ch) * t) * hw + p;
    float P = 0.f;
    for (int u = 0; u < t; ++u) {
      float xv = xs[(size_t)u * hw];
      float v;
      if (mode == 0) {
        v = (u == 0) ? xv : (1.f - mi.m[u]) * xv + mi.m[u] * P;
        P = v;
      } else {
        int q = mi.partner[u];
        v = (q == u) ? xv : (1.f - mi.coef[u]) * xv + mi.coef[u] * xs[(size_t)q * hw];
      }
      if (FMT == IVF_PFMT_NCDHW_F32)
        out[((size_t)(b * c + ch) * t + u) * hw + p] = v;
      else if (FMT == IVF_PFMT_NDHWC_F32)
        out[(((size_t)b * t + u) * hw + p) * c + ch] = v;
      else  // IVF_PFMT_TBHWC_F32: time-major frames for the ConvLSTM
        out[(((size_t)u * gridDim.y + b) * hw + p) * c + ch] = v;
    }
  }
}

// ------------------------------------------------------------------ forward, space-to-depth bf16
// one thread per macro pixel (h/2, w/2): 2x2 pixels x c channels.
// S2D3 = true : I3D stem operand, two frames per 64-byte record [b][t/2][h/2][w/2][32]
// S2D3 = false: ConvLSTM x-conv operand, one frame per 32-byte record, time-major [t][b][h/2][w/2][16]
template <bool S2D3>
__global__ void perturb_fwd_s2d_kernel(int mode, const float* __restrict__ x,
                                       const float* __restrict__ mask, int mask_bstride, int c,
                                       int t, int hh, int ww, __nv_bfloat16* __restrict__ out) {
  __shared__ MaskInfo mi;
  const int b = blockIdx.y;
  const int nb = gridDim.y;
  build_mask_info(&mi, mask + (size_t)b * mask_bstride, t, mode);
  const int h2 = hh / 2, w2 = ww / 2, t2 = t / 2;
  const size_t hw = (size_t)hh * ww;
  constexpr int FR = S2D3 ? 2 : 1;    // frames per record
  constexpr int REC = S2D3 ? 32 : 16;  // channels per record
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < h2 * w2; idx += gridDim.x * blockDim.x) {
    const int y = idx / w2, xq = idx - y * w2;
    float P[MAX_C][2][2];
    for (int r = 0; r < t / FR; ++r) {
      __align__(16) __nv_bfloat16 rec[REC];
#pragma unroll
      for (int i = 0; i < REC; ++i) rec[i] = __float2bfloat16_rn(0.f);
#pragma unroll
      for (int dt = 0; dt < FR; ++dt) {
        const int u = FR * r + dt;
        const float mu = mi.m[u], cf = mi.coef[u];
        const int q = mi.partner[u];
#pragma unroll
        for (int ch = 0; ch < MAX_C; ++ch) {
          if (ch >= c) break;
          const float* xf = x + ((size_t)(b * c + ch) * t) * hw;
#pragma unroll
          for (int dh = 0; dh < 2; ++dh) {
            const size_t off = (size_t)(2 * y + dh) * ww + 2 * xq;
            float2 xv = *reinterpret_cast<const float2*>(xf + (size_t)u * hw + off);
            float v0, v1;
            if (mode == 0) {
              if (u == 0) {
                v0 = xv.x;
                v1 = xv.y;
              } else {
                v0 = (1.f - mu) * xv.x + mu * P[ch][dh][0];
                v1 = (1.f - mu) * xv.y + mu * P[ch][dh][1];
              }
              P[ch][dh][0] = v0;
              P[ch][dh][1] = v1;
            } else if (q == u) {
              v0 = xv.x;
              v1 = xv.y;
            } else {
              float2 xp = *reinterpret_cast<const float2*>(xf + (size_t)q * hw + off);
              v0 = (1.f - cf) * xv.x + cf * xp.x;
              v1 = (1.f - cf) * xv.y + cf * xp.y;
            }
            rec[((dt * 2 + dh) * 2 + 0) * c + ch] = __float2bfloat16_rn(v0);
            rec[((dt * 2 + dh) * 2 + 1) * c + ch] = __float2bfloat16_rn(v1);
          }
        }
      }
      size_t rec_index = S2D3 ? ((((size_t)b * t2 + r) * h2 + y) * w2 + xq)
                              : ((((size_t)r * nb + b) * h2 + y) * w2 + xq);
      uint4* dst = reinterpret_cast<uint4*>(out + rec_index * REC);
      const uint4* src = reinterpret_cast<const uint4*>(rec);
#pragma unroll
      for (int i = 0; i < REC / 8; ++i) dst[i] = src[i];
    }
  }
}

// ------------------------------------------------------------------ backward
// Per (pixel, channel) item: accumulate this item's contribution to dm[0..T) into acc[].
// XL(u) -> x value of frame u, GL(u) -> upstream gradient of frame u.
template <int TT, typename XL, typename GL>
__device__ __forceinline__ void item_bwd(int mode, int t, const MaskInfo& mi, XL xl, GL gl,
                                         float (&acc)[TT]) {
  if (mode == 0) {
    float D[TT];  // D[u] = P[u-1] - x[u]
    float P = 0.f;
#pragma unroll
    for (int u = 0; u < TT; ++u) {
      if (u < t) {
        float xv = xl(u);
        if (u == 0) {
          D[u] = 0.f;
          P = xv;
        } else {
          D[u] = P - xv;
          P = (1.f - mi.m[u]) * xv + mi.m[u] * P;
        }
      }
    }
    float G = 0.f;
#pragma unroll
    for (int u = TT - 1; u >= 1; --u) {
      if (u < t) {
        G = gl(u) + ((u + 1 < t) ? mi.m[u + 1] * G : 0.f);
        acc[u] = fmaf(G, D[u], acc[u]);
      }
    }
  } else {
#pragma unroll
    for (int u = 0; u < TT; ++u) {
      if (u < t && mi.front[u]) {
        int q = mi.partner[u];
        acc[u] = fmaf(gl(u) - gl(q), xl(q) - xl(u), acc[u]);
      }
    }
  }
}

template <int TT>
__device__ __forceinline__ void block_reduce_dm(float (&acc)[TT], int t, float* red /*[warps][TT]*/,
                                                float* __restrict__ dm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int u = 0; u < TT; ++u) {
    float v = ivf_warp_sum(acc[u]);
    if (lane == 0) red[warp * TT + u] = v;
  }
  __syncthreads();
  for (int u = threadIdx.x; u < t; u += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[w * TT + u];
    atomicAdd(dm + u, s);
  }
}

// planar gout (fp32, NCDHW or NDHWC); one thread per (pixel, channel); grid.y = clip
template <int TT, int FMT>
__global__ void __launch_bounds__(256)
perturb_bwd_planar_kernel(int mode, const float* __restrict__ x, const float* __restrict__ mask,
                          int mask_bstride, int c, int t, int hw, const float* __restrict__ gout,
                          float* __restrict__ dmask) {
  __shared__ MaskInfo mi;
  __shared__ float red[8 * TT];
  const int b = blockIdx.y;
  build_mask_info(&mi, mask + (size_t)b * mask_bstride, t, mode);
  float acc[TT];
#pragma unroll
  for (int u = 0; u < TT; ++u) acc[u] = 0.f;
  const int total = c * hw;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ch = idx / hw, p = idx - ch * hw;
    const float* xs = x + ((size_t)(b * c + ch) * t) * hw + p;
    auto xl = [&](int u) { return xs[(size_t)u * hw]; };
    if (FMT == IVF_PFMT_NCDHW_F32) {
      const float* gs = gout + ((size_t)(b * c + ch) * t) * hw + p;
      auto gl = [&](int u) { return gs[(size_t)u * hw]; };
      item_bwd<TT>(mode, t, mi, xl, gl, acc);
    } else if (FMT == IVF_PFMT_NDHWC_F32) {
      const float* gs = gout + ((size_t)b * t * hw + p) * c + ch;
      auto gl = [&](int u) { return gs[(size_t)u * hw * c]; };
      item_bwd<TT>(mode, t, mi, xl, gl, acc);
    } else {  // time-major frames
      const float* gs = gout + ((size_t)b * hw + p) * c + ch;
      const size_t fstride = (size_t)gridDim.y * hw * c;
      auto gl = [&](int u) { return gs[(size_t)u * fstride]; };
      item_bwd<TT>(mode, t, mi, xl, gl, acc);
    }
  }
  block_reduce_dm<TT>(acc, t, red, dmask + (size_t)b * t);
}

// space-to-depth gout (bf16 or fp32 records): one block per (macro row, clip); the row's records for
// all frames are staged in shared memory with 16-byte loads, then each thread walks full-resolution
// pixels of the two rows (coalesced x reads along W).  S2D3: 32-channel two-frame records (I3D);
// otherwise 16-channel one-frame time-major records (ConvLSTM).
template <int TT, typename GT, bool S2D3>
__global__ void __launch_bounds__(256)
perturb_bwd_s2d_kernel(int mode, const float* __restrict__ x, const float* __restrict__ mask,
                       int mask_bstride, int c, int t, int hh, int ww, const GT* __restrict__ gout,
                       float* __restrict__ dmask) {
  extern __shared__ __align__(16) uint8_t gsm_raw[];
  GT* gsm = reinterpret_cast<GT*>(gsm_raw);  // [records][w2][REC]
  __shared__ MaskInfo mi;
  __shared__ float red[8 * TT];
  const int b = blockIdx.y, y = blockIdx.x, nb = gridDim.y;
  build_mask_info(&mi, mask + (size_t)b * mask_bstride, t, mode);
  constexpr int FR = S2D3 ? 2 : 1;
  constexpr int REC = S2D3 ? 32 : 16;
  const int h2 = hh / 2, w2 = ww / 2, nrec = t / FR;
  constexpr int PER16 = 16 / sizeof(GT);
  const int vec_per_frame = w2 * REC / PER16;
  for (int i = threadIdx.x; i < nrec * vec_per_frame; i += blockDim.x) {
    int r = i / vec_per_frame, v = i - r * vec_per_frame;
    size_t row = S2D3 ? (((size_t)b * nrec + r) * h2 + y) : (((size_t)r * nb + b) * h2 + y);
    const uint4* src = reinterpret_cast<const uint4*>(gout + row * w2 * REC) + v;
    reinterpret_cast<uint4*>(gsm + (size_t)r * w2 * REC)[v] = *src;
  }
  __syncthreads();
  float acc[TT];
#pragma unroll
  for (int u = 0; u < TT; ++u) acc[u] = 0.f;
  const size_t hw = (size_t)hh * ww;
  const int items = c * 2 * ww;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int wx = it % ww;
    const int r = it / ww;
    const int dh = r & 1, ch = r >> 1;
    const float* xs = x + ((size_t)(b * c + ch) * t) * hw + (size_t)(2 * y + dh) * ww + wx;
    const GT* gs = gsm + (size_t)(wx >> 1) * REC + (dh * 2 + (wx & 1)) * c + ch;
    auto xl = [&](int u) { return xs[(size_t)u * hw]; };
    auto gl = [&](int u) {
      return S2D3 ? ivf_to_float(gs[(size_t)(u >> 1) * w2 * REC + (u & 1) * 4 * c])
                  : ivf_to_float(gs[(size_t)u * w2 * REC]);
    };
    item_bwd<TT>(mode, t, mi, xl, gl, acc);
  }
  block_reduce_dm<TT>(acc, t, red, dmask + (size_t)b * t);
}

template <int TT, typename GT, bool S2D3>
int launch_bwd_s2d(ivf_handle* h, int mode, const float* x, const float* mask, int mask_bstride, int b,
                   int c, int t, int hh, int ww, const void* gout, float* dmask, cudaStream_t st) {
  constexpr int FR = S2D3 ? 2 : 1;
  constexpr int REC = S2D3 ? 32 : 16;
  size_t smem = (size_t)(t / FR) * (ww / 2) * REC * sizeof(GT);
  IVF_REQUIRE(smem <= 200 * 1024, "perturb_bwd(s2d): row tile of %zu bytes exceeds shared memory", smem);
  // opt in to large dynamic shared memory once per instantiation and device (not per launch, so a
  // captured iteration contains launches only)
  static bool attr_done[16] = {};
  int dev = h->device & 15;
  if (!attr_done[dev]) {
    IVF_CUDA(cudaFuncSetAttribute(perturb_bwd_s2d_kernel<TT, GT, S2D3>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done[dev] = true;
  }
  dim3 grid(hh / 2, b);
  perturb_bwd_s2d_kernel<TT, GT, S2D3><<<grid, 256, smem, st>>>(mode, x, mask, mask_bstride, c, t, hh, ww,
                                                                (const GT*)gout, dmask);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

template <int TT>
int launch_bwd(ivf_handle* h, int mode, const float* x, const float* mask, int mask_bstride, int b,
               int c, int t, int hh, int ww, int out_fmt, int gout_dtype, const void* gout,
               float* dmask, cudaStream_t st) {
  const int hw = hh * ww;
  if (out_fmt == IVF_PFMT_S2D_BF16) {
    if (gout_dtype == IVF_F32)
      return launch_bwd_s2d<TT, float, true>(h, mode, x, mask, mask_bstride, b, c, t, hh, ww, gout, dmask, st);
    return launch_bwd_s2d<TT, __nv_bfloat16, true>(h, mode, x, mask, mask_bstride, b, c, t, hh, ww, gout, dmask, st);
  }
  if (out_fmt == IVF_PFMT_S2D2_BF16) {
    if (gout_dtype == IVF_F32)
      return launch_bwd_s2d<TT, float, false>(h, mode, x, mask, mask_bstride, b, c, t, hh, ww, gout, dmask, st);
    return launch_bwd_s2d<TT, __nv_bfloat16, false>(h, mode, x, mask, mask_bstride, b, c, t, hh, ww, gout, dmask, st);
  }
  int bx = std::min(ivf_cdiv((long long)c * hw, 256), 8 * h->sm_count / std::max(b, 1) + 1);
  dim3 grid(bx, b);
  if (out_fmt == IVF_PFMT_NCDHW_F32)
    perturb_bwd_planar_kernel<TT, IVF_PFMT_NCDHW_F32><<<grid, 256, 0, st>>>(
        mode, x, mask, mask_bstride, c, t, hw, (const float*)gout, dmask);
  else if (out_fmt == IVF_PFMT_NDHWC_F32)
    perturb_bwd_planar_kernel<TT, IVF_PFMT_NDHWC_F32><<<grid, 256, 0, st>>>(
        mode, x, mask, mask_bstride, c, t, hw, (const float*)gout, dmask);
  else
    perturb_bwd_planar_kernel<TT, IVF_PFMT_TBHWC_F32><<<grid, 256, 0, st>>>(
        mode, x, mask, mask_bstride, c, t, hw, (const float*)gout, dmask);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

int check_common(int mode, int b, int c, int t, int hh, int ww, int out_fmt) {
  IVF_REQUIRE(mode == 0 || mode == 1, "perturb: mode must be 0 (freeze) or 1 (reverse)");
  IVF_REQUIRE(b > 0 && c > 0 && t > 0 && hh > 0 && ww > 0, "perturb: non-positive extent");
  IVF_REQUIRE(t <= MAX_T, "perturb: t = %d exceeds %d frames", t, MAX_T);
  IVF_REQUIRE(out_fmt >= 0 && out_fmt <= 4, "perturb: unknown out_fmt %d", out_fmt);
  if (out_fmt == IVF_PFMT_S2D_BF16)
    IVF_REQUIRE(c <= MAX_C && t % 2 == 0 && hh % 2 == 0 && ww % 2 == 0,
                "perturb(s2d): needs c <= 4 and even t/h/w (got c%d t%d h%d w%d)", c, t, hh, ww);
  if (out_fmt == IVF_PFMT_S2D2_BF16)
    IVF_REQUIRE(c <= MAX_C && hh % 2 == 0 && ww % 2 == 0,
                "perturb(s2d2): needs c <= 4 and even h/w (got c%d h%d w%d)", c, hh, ww);
  return IVF_OK;
}

}  // namespace

extern "C" int ivf_perturb_fwd(ivf_handle* h, int mode, const float* x, const float* mask,
                               int mask_bstride, int b, int c, int t, int hh, int ww, int out_fmt,
                               void* out, void* stream) {
  IVF_REQUIRE(h && x && mask && out, "ivf_perturb_fwd: null argument");
  int rc = check_common(mode, b, c, t, hh, ww, out_fmt);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int hw = hh * ww;
  if (out_fmt == IVF_PFMT_S2D_BF16 || out_fmt == IVF_PFMT_S2D2_BF16) {
    int items = (hh / 2) * (ww / 2);
    dim3 grid(ivf_cdiv(items, 128), b);
    if (out_fmt == IVF_PFMT_S2D_BF16)
      perturb_fwd_s2d_kernel<true><<<grid, 128, 0, st>>>(mode, x, mask, mask_bstride, c, t, hh, ww,
                                                         (__nv_bfloat16*)out);
    else
      perturb_fwd_s2d_kernel<false><<<grid, 128, 0, st>>>(mode, x, mask, mask_bstride, c, t, hh, ww,
                                                          (__nv_bfloat16*)out);
  } else {
    dim3 grid(ivf_cdiv((long long)c * hw, 256), b);
    if (out_fmt == IVF_PFMT_NCDHW_F32)
      perturb_fwd_planar_kernel<IVF_PFMT_NCDHW_F32><<<grid, 256, 0, st>>>(mode, x, mask, mask_bstride,
                                                                          c, t, hw, (float*)out);
    else if (out_fmt == IVF_PFMT_NDHWC_F32)
      perturb_fwd_planar_kernel<IVF_PFMT_NDHWC_F32><<<grid, 256, 0, st>>>(mode, x, mask, mask_bstride,
                                                                          c, t, hw, (float*)out);
    else
      perturb_fwd_planar_kernel<IVF_PFMT_TBHWC_F32><<<grid, 256, 0, st>>>(mode, x, mask, mask_bstride,
                                                                          c, t, hw, (float*)out);
  }
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_perturb_bwd(ivf_handle* h, int mode, const float* x, const float* mask,
                               int mask_bstride, int b, int c, int t, int hh, int ww, int out_fmt,
                               int gout_dtype, const void* gout, float* dmask, void* stream) {
  IVF_REQUIRE(h && x && mask && gout && dmask, "ivf_perturb_bwd: null argument");
  int rc = check_common(mode, b, c, t, hh, ww, out_fmt);
  if (rc) return rc;
  if (out_fmt != IVF_PFMT_S2D_BF16 && out_fmt != IVF_PFMT_S2D2_BF16)
    IVF_REQUIRE(gout_dtype == IVF_F32, "perturb_bwd: planar gout must be fp32");
  cudaStream_t st = (cudaStream_t)stream;
  IVF_CUDA(cudaMemsetAsync(dmask, 0, (size_t)b * t * sizeof(float), st));
  if (t <= 16)
    return launch_bwd<16>(h, mode, x, mask, mask_bstride, b, c, t, hh, ww, out_fmt, gout_dtype, gout, dmask, st);
  if (t <= 32)
    return launch_bwd<32>(h, mode, x, mask, mask_bstride, b, c, t, hh, ww, out_fmt, gout_dtype, gout, dmask, st);
  return launch_bwd<64>(h, mode, x, mask, mask_bstride, b, c, t, hh, ww, out_fmt, gout_dtype, gout, dmask, st);
}
