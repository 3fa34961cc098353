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
// S2D3 = true : I3D stem operand, two frames per 64-byte record [b][t/2][h/2][w/2][32]
// S2D3 = false: ConvLSTM x-conv operand, one frame per 32-byte record, time-major [t][b][h/2][w/2][16]
// Block = 128 macro pixels (h/2, w/2) x C channels, one thread per (macro pixel, channel) so the x
// reads of a warp are 256 contiguous bytes and the occupancy is not capped by a fat per-thread state.
// The record of a macro pixel interleaves channels, so the threads assemble records in shared memory
// (odd word pitch: conflict-free 2-byte scatter) and the block writes them out as one contiguous run
// of 128 records per frame (pair).  Shared memory is double-buffered: one barrier per record.
constexpr int S2D_PX = 128;

template <bool S2D3, int C>
__global__ void __launch_bounds__(S2D_PX * C)
perturb_fwd_s2d_kernel(int mode, const float* __restrict__ x, const float* __restrict__ mask,
                       int mask_bstride, int t, int hh, int ww, __nv_bfloat16* __restrict__ out) {
  constexpr int FR = S2D3 ? 2 : 1;        // frames per record
  constexpr int REC = S2D3 ? 32 : 16;     // bf16 channels per record
  constexpr int WORDS = REC / 2, PITCH = WORDS + 1;
  __shared__ MaskInfo mi;
  __shared__ uint32_t recs[2][S2D_PX * PITCH];
  const int b = blockIdx.y, nb = gridDim.y;
  build_mask_info(&mi, mask + (size_t)b * mask_bstride, t, mode);
  const int h2 = hh / 2, w2 = ww / 2, nrec = t / FR, npx = h2 * w2;
  const size_t hw = (size_t)hh * ww;
  const int px = threadIdx.x % S2D_PX, ch = threadIdx.x / S2D_PX;
  const int idx0 = blockIdx.x * S2D_PX;
  const int idx = idx0 + px;
  const bool live = idx < npx;
  const int y = live ? idx / w2 : 0, xq = live ? idx - y * w2 : 0;
  // channels 4*FR*C.. of a record are padding: zero them once in both buffers
  for (int i = threadIdx.x; i < 2 * S2D_PX * PITCH; i += blockDim.x) (&recs[0][0])[i] = 0u;
  __syncthreads();
  const float* xs = x + ((size_t)(b * C + ch) * t) * hw + (size_t)(2 * y) * ww + 2 * xq;
  float2 cur[FR][2], nxt[FR][2];
  float P[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  if (live) {
#pragma unroll
    for (int dt = 0; dt < FR; ++dt)
#pragma unroll
      for (int dh = 0; dh < 2; ++dh)
        nxt[dt][dh] = *reinterpret_cast<const float2*>(xs + (size_t)dt * hw + (size_t)dh * ww);
  }
  const int block_px = min(S2D_PX, npx - idx0);
  for (int r = 0; r < nrec; ++r) {
    __nv_bfloat16* mine = reinterpret_cast<__nv_bfloat16*>(&recs[r & 1][px * PITCH]);
    if (live) {
#pragma unroll
      for (int dt = 0; dt < FR; ++dt)
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) cur[dt][dh] = nxt[dt][dh];
      if (r + 1 < nrec) {
#pragma unroll
        for (int dt = 0; dt < FR; ++dt)
#pragma unroll
          for (int dh = 0; dh < 2; ++dh)
            nxt[dt][dh] = *reinterpret_cast<const float2*>(xs + (size_t)(FR * (r + 1) + dt) * hw +
                                                           (size_t)dh * ww);
      }
#pragma unroll
      for (int dt = 0; dt < FR; ++dt) {
        const int u = FR * r + dt;
        const float mu = mi.m[u], cf = mi.coef[u];
        const int q = mi.partner[u];
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) {
          const float2 xv = cur[dt][dh];
          float v0, v1;
          if (mode == 0) {
            if (u == 0) {
              v0 = xv.x;
              v1 = xv.y;
            } else {
              v0 = (1.f - mu) * xv.x + mu * P[dh][0];
              v1 = (1.f - mu) * xv.y + mu * P[dh][1];
            }
            P[dh][0] = v0;
            P[dh][1] = v1;
          } else if (q == u) {
            v0 = xv.x;
            v1 = xv.y;
          } else {
            const float2 xp = *reinterpret_cast<const float2*>(xs + (size_t)q * hw + (size_t)dh * ww);
            v0 = (1.f - cf) * xv.x + cf * xp.x;
            v1 = (1.f - cf) * xv.y + cf * xp.y;
          }
          mine[((dt * 2 + dh) * 2 + 0) * C + ch] = __float2bfloat16_rn(v0);
          mine[((dt * 2 + dh) * 2 + 1) * C + ch] = __float2bfloat16_rn(v1);
        }
      }
    }
    __syncthreads();
    // records of this block for record row r are contiguous in the output
    const size_t rec0 = S2D3 ? (((size_t)b * nrec + r) * npx + idx0) : (((size_t)r * nb + b) * npx + idx0);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + rec0 * REC);
    const uint32_t* src = recs[r & 1];
    for (int i = threadIdx.x; i < block_px * WORDS; i += blockDim.x)
      dst[i] = src[(i / WORDS) * PITCH + (i % WORDS)];
    // no second barrier: the next record goes to the other buffer, and the barrier of iteration r+1
    // orders these reads before the writes of iteration r+2
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

// Two horizontally adjacent pixels of one channel at once (float2 x loads, two g values of one record).
template <int TT, typename XL, typename GL>
__device__ __forceinline__ void item_bwd2(int mode, int t, const MaskInfo& mi, XL xl, GL gl,
                                          float (&acc)[TT]) {
  if (mode == 0) {
    float2 D[TT];
    float2 P = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < TT; ++u) {
      if (u < t) {
        float2 xv = xl(u);
        if (u == 0) {
          D[u] = make_float2(0.f, 0.f);
          P = xv;
        } else {
          const float mu = mi.m[u];
          D[u] = make_float2(P.x - xv.x, P.y - xv.y);
          P.x = (1.f - mu) * xv.x + mu * P.x;
          P.y = (1.f - mu) * xv.y + mu * P.y;
        }
      }
    }
    float2 G = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = TT - 1; u >= 1; --u) {
      if (u < t) {
        const float mn = (u + 1 < t) ? mi.m[u + 1] : 0.f;
        const float2 g = gl(u);
        G.x = fmaf(mn, G.x, g.x);
        G.y = fmaf(mn, G.y, g.y);
        acc[u] = fmaf(G.x, D[u].x, fmaf(G.y, D[u].y, acc[u]));
      }
    }
  } else {
#pragma unroll
    for (int u = 0; u < TT; ++u) {
      if (u < t && mi.front[u]) {
        const int q = mi.partner[u];
        const float2 gu = gl(u), gq = gl(q), xu = xl(u), xq = xl(q);
        acc[u] = fmaf(gu.x - gq.x, xq.x - xu.x, fmaf(gu.y - gq.y, xq.y - xu.y, acc[u]));
      }
    }
  }
}

// space-to-depth gout (bf16 or fp32 records): one block per (macro row, clip).  The real channels of
// the row's records for all frames are staged in shared memory (16-byte global loads, records re-pitched
// to an odd number of words so that threads reading the same element of neighbouring records hit
// different banks), then each thread owns (row parity, macro column) and walks the channels: float2 x
// reads, 256 contiguous bytes per warp.  S2D3: 32-channel two-frame records (I3D); otherwise
// 16-channel one-frame time-major records (ConvLSTM).
template <int TT, typename GT, bool S2D3>
__global__ void __launch_bounds__(256, (TT <= 16 && sizeof(GT) == 2) ? 3 : 1)
perturb_bwd_s2d_kernel(int mode, const float* __restrict__ x, const float* __restrict__ mask,
                       int mask_bstride, int c, int t, int hh, int ww, const GT* __restrict__ gout,
                       float* __restrict__ dmask) {
  extern __shared__ __align__(16) uint32_t gsm_words[];  // [records][w2][pitch]
  __shared__ MaskInfo mi;
  __shared__ float red[8 * TT];
  const int b = blockIdx.y, y = blockIdx.x, nb = gridDim.y;
  build_mask_info(&mi, mask + (size_t)b * mask_bstride, t, mode);
  constexpr int FR = S2D3 ? 2 : 1;
  constexpr int REC = S2D3 ? 32 : 16;
  constexpr int PER16 = 16 / sizeof(GT);
  constexpr int PERW = 4 / sizeof(GT);
  const int h2 = hh / 2, w2 = ww / 2, nrec = t / FR;
  const int real_words = FR * 4 * c / PERW;  // even
  const int pitch = real_words + 1;
  const int nvec = (real_words + 3) / 4;
  const int per_row = w2 * nvec;
  // the loads of a batch are all issued before the first shared-memory store waits on one of them
  constexpr int STAGE_BATCH = 6;
  const int stage_total = nrec * per_row;
  for (int i0 = threadIdx.x; i0 < stage_total; i0 += STAGE_BATCH * blockDim.x) {
    uint4 val[STAGE_BATCH];
    int dsto[STAGE_BATCH], left[STAGE_BATCH];
#pragma unroll
    for (int k = 0; k < STAGE_BATCH; ++k) {
      const int i = i0 + k * blockDim.x;
      left[k] = 0;
      if (i < stage_total) {
        const int r = i / per_row, rem = i - r * per_row;
        const int px = rem / nvec, v = rem - px * nvec;
        const size_t row = S2D3 ? (((size_t)b * nrec + r) * h2 + y) : (((size_t)r * nb + b) * h2 + y);
        val[k] = *reinterpret_cast<const uint4*>(gout + (row * w2 + px) * REC + v * PER16);
        dsto[k] = (r * w2 + px) * pitch + v * 4;
        left[k] = real_words - v * 4;
      }
    }
#pragma unroll
    for (int k = 0; k < STAGE_BATCH; ++k) {
      if (left[k] > 0) {
        uint32_t* dst = gsm_words + dsto[k];
        dst[0] = val[k].x;
        if (left[k] > 1) dst[1] = val[k].y;
        if (left[k] > 2) dst[2] = val[k].z;
        if (left[k] > 3) dst[3] = val[k].w;
      }
    }
  }
  __syncthreads();
  const GT* gsm = reinterpret_cast<const GT*>(gsm_words);
  const int gpitch = pitch * PERW;  // record pitch in GT elements
  float acc[TT];
#pragma unroll
  for (int u = 0; u < TT; ++u) acc[u] = 0.f;
  const size_t hw = (size_t)hh * ww;
  const int items = c * 2 * w2;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int xq = it % w2;
    const int r = it / w2;
    const int dh = r & 1, ch = r >> 1;
    const float* xs = x + ((size_t)(b * c + ch) * t) * hw + (size_t)(2 * y + dh) * ww + 2 * xq;
    const GT* gs = gsm + (size_t)xq * gpitch + (dh * 2) * c + ch;
    auto xl = [&](int u) { return *reinterpret_cast<const float2*>(xs + (size_t)u * hw); };
    auto gl = [&](int u) {
      const GT* g = S2D3 ? gs + (size_t)(u >> 1) * w2 * gpitch + (u & 1) * 4 * c : gs + (size_t)u * w2 * gpitch;
      return make_float2(ivf_to_float(g[0]), ivf_to_float(g[c]));
    };
    item_bwd2<TT>(mode, t, mi, xl, gl, acc);
  }
  block_reduce_dm<TT>(acc, t, red, dmask + (size_t)b * t);
}

template <int TT, typename GT, bool S2D3>
int launch_bwd_s2d(ivf_handle* h, int mode, const float* x, const float* mask, int mask_bstride, int b,
                   int c, int t, int hh, int ww, const void* gout, float* dmask, cudaStream_t st) {
  constexpr int FR = S2D3 ? 2 : 1;
  const int real_words = FR * 4 * c / (4 / (int)sizeof(GT));
  size_t smem = (size_t)(t / FR) * (ww / 2) * (real_words + 1) * 4;
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
  // one thread per (row parity, macro column) when the row fits a block: the channel loop then has no
  // ragged last pass
  const int threads = std::min(256, std::max(64, (ww + 31) / 32 * 32));
  perturb_bwd_s2d_kernel<TT, GT, S2D3><<<grid, threads, smem, st>>>(mode, x, mask, mask_bstride, c, t, hh, ww,
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
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && x && mask && out, "ivf_perturb_fwd: null argument");
  int rc = check_common(mode, b, c, t, hh, ww, out_fmt);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int hw = hh * ww;
  if (out_fmt == IVF_PFMT_S2D_BF16 || out_fmt == IVF_PFMT_S2D2_BF16) {
    int items = (hh / 2) * (ww / 2);
    dim3 grid(ivf_cdiv(items, S2D_PX), b);
    const bool s3 = out_fmt == IVF_PFMT_S2D_BF16;
#define IVF_PFWD(S, CC)                                                                              \
  perturb_fwd_s2d_kernel<S, CC><<<grid, S2D_PX * CC, 0, st>>>(mode, x, mask, mask_bstride, t, hh, ww, \
                                                              (__nv_bfloat16*)out)
    switch (c) {
      case 1: if (s3) IVF_PFWD(true, 1); else IVF_PFWD(false, 1); break;
      case 2: if (s3) IVF_PFWD(true, 2); else IVF_PFWD(false, 2); break;
      case 3: if (s3) IVF_PFWD(true, 3); else IVF_PFWD(false, 3); break;
      default: if (s3) IVF_PFWD(true, 4); else IVF_PFWD(false, 4); break;
    }
#undef IVF_PFWD
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
  IVF_ON_DEVICE(h);
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
