// Channels-last max-pool with TF-'same' ZERO padding, forward and backward.
// Reference: pt/models/I3D_doubled.py:8-40 (MaxPool3dSamePadding = F.pad with zeros, then
// nn.MaxPool3d without padding) — the padded zeros take part in the max, and ATen's
// max_pool3d_with_indices keeps the FIRST maximum in (d,h,w) window-scan order; backward sends
// the gradient to that element only (a padded winner drops it).  Also used as nn.MaxPool2d
// (pt/models/convolution_lstm.py:79) with kd=1, pad 0.
// Bandwidth-bound: one thread per (pixel, 16-byte channel vector), coalesced along C.
#include "common.cuh"

namespace {

template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T* p, float (&f)[VEC]) {
  if constexpr (VEC == 1) {
    f[0] = ivf_to_float(p[0]);
  } else {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < VEC; ++i) f[i] = ivf_to_float(e[i]);
  }
}
template <typename T, int VEC>
__device__ __forceinline__ void store_vec(T* p, const float (&f)[VEC]) {
  if constexpr (VEC == 1) {
    p[0] = ivf_from_float<T>(f[0]);
  } else {
    uint4 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int i = 0; i < VEC; ++i) e[i] = ivf_from_float<T>(f[i]);
    *reinterpret_cast<uint4*>(p) = raw;
  }
}

// Compile-time window geometry (KD..SW > 0) removes every runtime division from the tap loops and lets the
// compiler unroll them; GEN = true keeps the fully general runtime-geometry path.
template <int KD, int KH, int KW, int SD, int SH, int SW>
struct PoolGeo {
  __device__ static int kd(const ivf_pool_desc&) { return KD; }
  __device__ static int kh(const ivf_pool_desc&) { return KH; }
  __device__ static int kw(const ivf_pool_desc&) { return KW; }
  __device__ static int sd(const ivf_pool_desc&) { return SD; }
  __device__ static int sh(const ivf_pool_desc&) { return SH; }
  __device__ static int sw(const ivf_pool_desc&) { return SW; }
};
struct PoolGeoDyn {
  __device__ static int kd(const ivf_pool_desc& d) { return d.kd; }
  __device__ static int kh(const ivf_pool_desc& d) { return d.kh; }
  __device__ static int kw(const ivf_pool_desc& d) { return d.kw; }
  __device__ static int sd(const ivf_pool_desc& d) { return d.sd; }
  __device__ static int sh(const ivf_pool_desc& d) { return d.sh; }
  __device__ static int sw(const ivf_pool_desc& d) { return d.sw; }
};

// bf16 x 8 channels, packed arithmetic, one block row per output row.  The generic kernels below spend
// ~5 instructions per channel per tap (convert, compare, two selects) plus 64-bit index arithmetic per
// tap and are INSTRUCTION bound, not memory bound (ncu: 134 us for the 48 MB Mixed_3c pool; an
// L1-friendlier thread mapping changed nothing).  Here
//   * blockIdx.y/z enumerate (clip, depth, row): everything but the column is block-uniform, so the tap
//     offsets and the depth/row bounds tests live in uniform registers, the thread adds one offset;
//   * a tap costs per channel PAIR one bf16x2 compare-to-mask (+ one for the NaN rule), one bf16x2 max and
//     two LOP3 on the 2 x 16-bit running index — ATen's "first maximum wins, NaN wins" rule; values are
//     never converted.
// Host guarantees every element offset fits 31 bits.
template <typename G>
__global__ void __launch_bounds__(256)
maxpool_fwd_bf16x8_kernel(ivf_pool_desc d, const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                          uint8_t* __restrict__ argmax, int rows) {
  const int row = blockIdx.z * gridDim.y + blockIdx.y;
  if (row >= rows) return;
  const int cv = d.c >> 3;
  const int el = blockIdx.x * 256 + threadIdx.x;
  if (el >= d.ow * cv) return;
  const int KD = G::kd(d), KH = G::kh(d), KW = G::kw(d), SD = G::sd(d), SH = G::sh(d), SW = G::sw(d);
  const int ow = el / cv, c = (el - ow * cv) << 3;
  const int oh = row % d.oh;
  const int t = row / d.oh;
  const int od = t % d.od, n = t / d.od;
  const int zd0 = od * SD - d.pd, zh0 = oh * SH - d.ph, zw0 = ow * SW - d.pw;
  const int pix0 = ((n * d.id + zd0) * d.ih + zh0) * d.iw + zw0;  // window origin (may lie in the padding)
  const __nv_bfloat16* p0 = in + (long long)pix0 * d.in_ld + d.in_coff + c;
  __nv_bfloat162 best[4];
  uint32_t bidx[4];
  const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    best[i] = ninf;
    bidx[i] = 0u;
  }
#pragma unroll
  for (int a = 0; a < KD; ++a) {
    const bool dok = (unsigned)(zd0 + a) < (unsigned)d.id;
#pragma unroll
    for (int b = 0; b < KH; ++b) {
      const bool hok = dok && (unsigned)(zh0 + b) < (unsigned)d.ih;
      const int delta_row = ((a * d.ih + b) * d.iw) * d.in_ld;
#pragma unroll
      for (int e = 0; e < KW; ++e) {
        const uint32_t tap2 = (uint32_t)((a * KH + b) * KW + e) * 0x00010001u;
        uint4 raw = make_uint4(0u, 0u, 0u, 0u);  // explicit zero padding
        if (hok && (unsigned)(zw0 + e) < (unsigned)d.iw)
          raw = *reinterpret_cast<const uint4*>(p0 + delta_row + e * d.in_ld);
        const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t m = __hgt2_mask(v[i], best[i]) | __hneu2_mask(v[i], v[i]);
          best[i] = __hmax2_nan(best[i], v[i]);
          bidx[i] = (bidx[i] & ~m) | (tap2 & m);
        }
      }
    }
  }
  const int opix = row * d.ow + ow;
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&best[0]);
  o.y = *reinterpret_cast<uint32_t*>(&best[1]);
  o.z = *reinterpret_cast<uint32_t*>(&best[2]);
  o.w = *reinterpret_cast<uint32_t*>(&best[3]);
  *reinterpret_cast<uint4*>(out + (long long)opix * d.out_ld + d.out_coff + c) = o;
  if (argmax) {
    uint2 pk;  // 2 x 16-bit indices per word -> bytes
    pk.x = (bidx[0] & 0xffu) | ((bidx[0] >> 8) & 0xff00u) | ((bidx[1] & 0xffu) << 16) | ((bidx[1] & 0xff0000u) << 8);
    pk.y = (bidx[2] & 0xffu) | ((bidx[2] >> 8) & 0xff00u) | ((bidx[3] & 0xffu) << 16) | ((bidx[3] & 0xff0000u) << 8);
    *reinterpret_cast<uint2*>(argmax + (long long)opix * d.c + c) = pk;
  }
}

// Backward of the same shape: one block row per INPUT row; a thread owns 8 channels of one input pixel and
// visits the windows that cover it.  Depth/row window indices are block-uniform; one SIMD byte compare per
// four channels decides whether the 16-byte gradient load is needed at all.
template <typename G>
__global__ void __launch_bounds__(256)
maxpool_bwd_bf16x8_kernel(ivf_pool_desc d, const __nv_bfloat16* __restrict__ dy,
                          const uint8_t* __restrict__ argmax, const float* __restrict__ acc_in,
                          const __nv_bfloat16* __restrict__ mask_y, const float* __restrict__ mask_scale,
                          void* __restrict__ dx, int rows) {
  const int row = blockIdx.z * gridDim.y + blockIdx.y;
  if (row >= rows) return;
  const int cv = d.c >> 3;
  const int el = blockIdx.x * 256 + threadIdx.x;
  if (el >= d.iw * cv) return;
  const int KD = G::kd(d), KH = G::kh(d), KW = G::kw(d), SD = G::sd(d), SH = G::sh(d), SW = G::sw(d);
  const int iw = el / cv, c = (el - iw * cv) << 3;
  const int ih = row % d.ih;
  const int t = row / d.ih;
  const int idd = t % d.id, n = t / d.id;
  float g[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) g[i] = 0.f;
  // The kernel is issue bound (ncu: 72-78 % issue-slot utilisation, DRAM at 20 %), so only the windows that
  // can contain this element are visited: along a dimension of stride S these are the taps congruent to
  // (x + pad) mod S - ceil(K/S) of them instead of K (4 of 9 for the 1x3x3 stride-2 pools, 8 of 27 for 3x3x3
  // stride 2, 1 of 8 for 2x2x2 stride 2), with no divisibility test left inside the loop.
  const int pd0 = (idd + d.pd) % SD, ph0 = (ih + d.ph) % SH, pw0 = (iw + d.pw) % SW;
  const int qd = (idd + d.pd) / SD, qh = (ih + d.ph) / SH, qw = (iw + d.pw) / SW;  // window of tap == remainder
#pragma unroll
  for (int ja = 0; ja * SD < KD; ++ja) {
    const int a = pd0 + ja * SD, od = qd - ja;
    if (a >= KD || od < 0 || od >= d.od) continue;
#pragma unroll
    for (int jb = 0; jb * SH < KH; ++jb) {
      const int b = ph0 + jb * SH, oh = qh - jb;
      if (b >= KH || oh < 0 || oh >= d.oh) continue;
      const int orow = ((n * d.od + od) * d.oh + oh) * d.ow;
#pragma unroll
      for (int je = 0; je * SW < KW; ++je) {
        const int e = pw0 + je * SW, ow = qw - je;
        if (e >= KW || ow < 0 || ow >= d.ow) continue;
        const int opix = orow + ow;
        const uint2 pk = *reinterpret_cast<const uint2*>(argmax + (long long)opix * d.c + c);
        const uint32_t tap4 = (uint32_t)((a * KH + b) * KW + e) * 0x01010101u;
        const uint32_t e0 = __vcmpeq4(pk.x, tap4), e1 = __vcmpeq4(pk.y, tap4);
        if ((e0 | e1) == 0u) continue;
        const uint4 raw = *reinterpret_cast<const uint4*>(dy + (long long)opix * d.out_ld + d.out_coff + c);
        const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (((i < 4 ? e0 : e1) >> (8 * (i & 3))) & 1u) g[i] += __bfloat162float(v[i]);
      }
    }
  }
  const int ipix = row * d.iw + iw;
  const long long o = (long long)ipix * d.in_ld + d.in_coff + c;
  if (d.flags & IVF_EP_ACCUM) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float4 a4 = *reinterpret_cast<const float4*>(acc_in + o + 4 * i);
      g[4 * i] += a4.x;
      g[4 * i + 1] += a4.y;
      g[4 * i + 2] += a4.z;
      g[4 * i + 3] += a4.w;
    }
  }
  if (d.flags & IVF_EP_MASK) {
    const uint4 raw = *reinterpret_cast<const uint4*>(mask_y + (long long)ipix * d.mask_ld + d.mask_coff + c);
    const __nv_bfloat16* y = reinterpret_cast<const __nv_bfloat16*>(&raw);
    const float4 s0 = *reinterpret_cast<const float4*>(mask_scale + c);
    const float4 s1 = *reinterpret_cast<const float4*>(mask_scale + c + 4);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = __bfloat162float(y[i]) > 0.f ? g[i] * sc[i] : 0.f;
  }
  if (d.flags & IVF_EP_OUT_F32) {
    float* p = reinterpret_cast<float*>(dx) + o;
    reinterpret_cast<float4*>(p)[0] = make_float4(g[0], g[1], g[2], g[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(g[4], g[5], g[6], g[7]);
  } else {
    store_vec<__nv_bfloat16, 8>(reinterpret_cast<__nv_bfloat16*>(dx) + o, g);
  }
}

template <typename T, int VEC, typename G>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(ivf_pool_desc d, const T* __restrict__ in, T* __restrict__ out,
                   uint8_t* __restrict__ argmax, long long total) {
  const int cv = d.c / VEC;
  const int KD = G::kd(d), KH = G::kh(d), KW = G::kw(d), SD = G::sd(d), SH = G::sh(d), SW = G::sw(d);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % cv) * VEC;
    long long opix = idx / cv;
    int ow = (int)(opix % d.ow);
    long long t = opix / d.ow;
    int oh = (int)(t % d.oh);
    t /= d.oh;
    int od = (int)(t % d.od);
    int n = (int)(t / d.od);
    float best[VEC];
    int bidx[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      best[i] = -INFINITY;
      bidx[i] = 0;
    }
    const T* base = in + d.in_coff + c;
#pragma unroll
    for (int a = 0; a < KD; ++a) {
      int zd = od * SD - d.pd + a;
#pragma unroll
      for (int b = 0; b < KH; ++b) {
        int zh = oh * SH - d.ph + b;
#pragma unroll
        for (int e = 0; e < KW; ++e) {
          int zw = ow * SW - d.pw + e;
          const int tap = (a * KH + b) * KW + e;
          float v[VEC];
          if ((unsigned)zd < (unsigned)d.id && (unsigned)zh < (unsigned)d.ih && (unsigned)zw < (unsigned)d.iw) {
            size_t pix = (((size_t)n * d.id + zd) * d.ih + zh) * d.iw + zw;
            load_vec<T, VEC>(base + pix * d.in_ld, v);
          } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[i] = 0.f;  // explicit zero padding
          }
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            if (v[i] > best[i] || v[i] != v[i]) {
              best[i] = v[i];
              bidx[i] = tap;
            }
          }
        }
      }
    }
    store_vec<T, VEC>(out + (size_t)opix * d.out_ld + d.out_coff + c, best);
    if (argmax) {
      uint8_t* am = argmax + (size_t)opix * d.c + c;
      if constexpr (VEC == 8) {
        uint2 pk;
        pk.x = bidx[0] | (bidx[1] << 8) | (bidx[2] << 16) | (bidx[3] << 24);
        pk.y = bidx[4] | (bidx[5] << 8) | (bidx[6] << 16) | (bidx[7] << 24);
        *reinterpret_cast<uint2*>(am) = pk;
      } else if constexpr (VEC == 4) {
        *reinterpret_cast<uint32_t*>(am) = bidx[0] | (bidx[1] << 8) | (bidx[2] << 16) | (bidx[3] << 24);
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) am[i] = (uint8_t)bidx[i];
      }
    }
  }
}

template <typename T, int VEC, typename G>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(ivf_pool_desc d, const T* __restrict__ dy, const uint8_t* __restrict__ argmax,
                   const float* __restrict__ acc_in, const T* __restrict__ mask_y,
                   const float* __restrict__ mask_scale, void* __restrict__ dx, long long total) {
  const int cv = d.c / VEC;
  const int KD = G::kd(d), KH = G::kh(d), KW = G::kw(d), SD = G::sd(d), SH = G::sh(d), SW = G::sw(d);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % cv) * VEC;
    long long ipix = idx / cv;
    int iw = (int)(ipix % d.iw);
    long long t = ipix / d.iw;
    int ih = (int)(t % d.ih);
    t /= d.ih;
    int idd = (int)(t % d.id);
    int n = (int)(t / d.id);
    float g[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[i] = 0.f;
#pragma unroll
    for (int a = 0; a < KD; ++a) {
      int nd = idd + d.pd - a;
      if (nd < 0 || (SD > 1 && nd % SD)) continue;
      int od = SD > 1 ? nd / SD : nd;
      if (od >= d.od) continue;
#pragma unroll
      for (int b = 0; b < KH; ++b) {
        int nh = ih + d.ph - b;
        if (nh < 0 || (SH > 1 && nh % SH)) continue;
        int oh = SH > 1 ? nh / SH : nh;
        if (oh >= d.oh) continue;
#pragma unroll
        for (int e = 0; e < KW; ++e) {
          int nw = iw + d.pw - e;
          if (nw < 0 || (SW > 1 && nw % SW)) continue;
          int ow = SW > 1 ? nw / SW : nw;
          if (ow >= d.ow) continue;
          const int tap = (a * KH + b) * KW + e;
          size_t opix = (((size_t)n * d.od + od) * d.oh + oh) * d.ow + ow;
          const uint8_t* am = argmax + opix * d.c + c;
          uint32_t a0, a1 = 0;
          if constexpr (VEC == 8) {
            uint2 pk = *reinterpret_cast<const uint2*>(am);
            a0 = pk.x;
            a1 = pk.y;
          } else if constexpr (VEC == 4) {
            a0 = *reinterpret_cast<const uint32_t*>(am);
          } else {
            a0 = am[0];
          }
          // skip the 16-byte gradient load when no channel of this vector selected this tap: one SIMD byte
          // compare per four channels (0xff where argmax == tap)
          const uint32_t tap4 = (uint32_t)tap * 0x01010101u;
          const uint32_t e0 = __vcmpeq4(a0, tap4), e1 = VEC > 4 ? __vcmpeq4(a1, tap4) : 0u;
          if (VEC >= 4 ? ((e0 | e1) == 0u) : ((e0 & 0xffu) == 0u)) continue;
          float v[VEC];
          load_vec<T, VEC>(dy + opix * d.out_ld + d.out_coff + c, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i)
            if (((i < 4 ? e0 : e1) >> (8 * (i & 3))) & 1u) g[i] += v[i];
        }
      }
    }
    size_t o = (size_t)ipix * d.in_ld + d.in_coff + c;
    if (d.flags & IVF_EP_ACCUM) {
      if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i) {
          float4 a4 = *reinterpret_cast<const float4*>(acc_in + o + 4 * i);
          g[4 * i] += a4.x;
          g[4 * i + 1] += a4.y;
          g[4 * i + 2] += a4.z;
          g[4 * i + 3] += a4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) g[i] += acc_in[o + i];
      }
    }
    if (d.flags & IVF_EP_MASK) {
      float y[VEC];
      load_vec<T, VEC>(mask_y + (size_t)ipix * d.mask_ld + d.mask_coff + c, y);
#pragma unroll
      for (int i = 0; i < VEC; ++i) g[i] = y[i] > 0.f ? g[i] * mask_scale[c + i] : 0.f;
    }
    if (d.flags & IVF_EP_OUT_F32) {
      float* p = reinterpret_cast<float*>(dx) + o;
#pragma unroll
      for (int i = 0; i < VEC; ++i) p[i] = g[i];
    } else {
      store_vec<T, VEC>(reinterpret_cast<T*>(dx) + o, g);
    }
  }
}

template <typename T>
constexpr int full_vec() {
  return 16 / sizeof(T);
}

bool vec_ok(const ivf_pool_desc* d, int vec, const void* a, const void* b, const void* c) {
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return d->c % vec == 0 && d->in_ld % vec == 0 && d->in_coff % vec == 0 && d->out_ld % vec == 0 &&
         d->out_coff % vec == 0 && d->mask_ld % vec == 0 && d->mask_coff % vec == 0 && al(a) &&
         al(b) && al(c);
}

int check_pool(const ivf_pool_desc* d) {
  IVF_REQUIRE(d->n > 0 && d->id > 0 && d->ih > 0 && d->iw > 0 && d->c > 0 && d->od > 0 && d->oh > 0 &&
                  d->ow > 0,
              "maxpool: non-positive extent");
  IVF_REQUIRE(d->kd * d->kh * d->kw <= 255, "maxpool: window too large for uint8 argmax");
  IVF_REQUIRE(d->in_ld >= d->in_coff + d->c && d->out_ld >= d->out_coff + d->c,
              "maxpool: channel slice exceeds ld");
  IVF_REQUIRE(d->dtype == IVF_F32 || d->dtype == IVF_BF16, "maxpool: unknown dtype");
  return IVF_OK;
}

// every element offset the row-block kernels form fits a signed 32-bit int
bool fits31(const ivf_pool_desc* d) {
  const long long lim = (1ll << 31) - 1;
  const long long ip = (long long)d->n * d->id * d->ih * d->iw, op = (long long)d->n * d->od * d->oh * d->ow;
  const long long ld = d->in_ld > d->mask_ld ? d->in_ld : d->mask_ld;
  return (ip + (long long)d->iw * d->ih * 4) * ld < lim && op * (d->out_ld > d->c ? d->out_ld : d->c) < lim;
}
dim3 row_grid(int rows, int elems_per_row) {
  const int gy = rows < 65535 ? rows : 65535;
  return dim3((elems_per_row + 255) / 256, gy, (rows + gy - 1) / gy);
}

int pool_blocks(ivf_handle* h, long long total) {
  long long b = (total + 255) / 256;
  long long cap = (long long)h->sm_count * 64;
  return (int)(b < cap ? b : cap);
}

#define IVF_POOL_GEO_DISPATCH(CALL)                                                              \
  do {                                                                                           \
    const int kd = d->kd, kh = d->kh, kw = d->kw, sd = d->sd, sh = d->sh, sw = d->sw;            \
    if (kd == 3 && kh == 3 && kw == 3 && sd == 1 && sh == 1 && sw == 1) { using G = PoolGeo<3, 3, 3, 1, 1, 1>; CALL; } \
    else if (kd == 1 && kh == 3 && kw == 3 && sd == 1 && sh == 2 && sw == 2) { using G = PoolGeo<1, 3, 3, 1, 2, 2>; CALL; } \
    else if (kd == 3 && kh == 3 && kw == 3 && sd == 2 && sh == 2 && sw == 2) { using G = PoolGeo<3, 3, 3, 2, 2, 2>; CALL; } \
    else if (kd == 2 && kh == 2 && kw == 2 && sd == 2 && sh == 2 && sw == 2) { using G = PoolGeo<2, 2, 2, 2, 2, 2>; CALL; } \
    else if (kd == 1 && kh == 1 && kw == 1 && sd == 1 && sh == 1 && sw == 1) { using G = PoolGeo<1, 1, 1, 1, 1, 1>; CALL; } \
    else { using G = PoolGeoDyn; CALL; }                                                         \
  } while (0)

template <typename T>
int fwd_t(ivf_handle* h, const ivf_pool_desc* d, const void* in, void* out, uint8_t* argmax,
          cudaStream_t st) {
  constexpr int V = full_vec<T>();
  long long opix = (long long)d->n * d->od * d->oh * d->ow;
  const int threads = 256;
  if (vec_ok(d, V, in, out, argmax)) {
    long long total = opix * (d->c / V);
    if constexpr (sizeof(T) == 2) {
      if (fits31(d)) {  // row-block packed kernel
        const int rows = d->n * d->od * d->oh;
        IVF_POOL_GEO_DISPATCH((maxpool_fwd_bf16x8_kernel<G><<<row_grid(rows, d->ow * (d->c / V)), threads, 0, st>>>(
            *d, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, argmax, rows)));
        IVF_LAUNCHED(h);
        return IVF_OK;
      }
    }
    IVF_POOL_GEO_DISPATCH((maxpool_fwd_kernel<T, V, G><<<pool_blocks(h, total), threads, 0, st>>>(
        *d, (const T*)in, (T*)out, argmax, total)));
  } else {
    long long total = opix * d->c;
    maxpool_fwd_kernel<T, 1, PoolGeoDyn><<<pool_blocks(h, total), threads, 0, st>>>(*d, (const T*)in, (T*)out,
                                                                                  argmax, total);
  }
  IVF_LAUNCHED(h);
  return IVF_OK;
}

template <typename T>
int bwd_t(ivf_handle* h, const ivf_pool_desc* d, const void* dy, const uint8_t* argmax,
          const float* acc_in, const void* mask_y, const float* mask_scale, void* dx,
          cudaStream_t st) {
  constexpr int V = full_vec<T>();
  long long ipix = (long long)d->n * d->id * d->ih * d->iw;
  const int threads = 256;
  // the fp32-out path stores scalars, so only the typed loads need 16-byte alignment
  bool v_ok = vec_ok(d, V, dy, mask_y, (d->flags & IVF_EP_OUT_F32) ? nullptr : dx) &&
              (reinterpret_cast<uintptr_t>(argmax) & 15) == 0 &&
              (acc_in == nullptr || (reinterpret_cast<uintptr_t>(acc_in) & 15) == 0);
  if (v_ok) {
    long long total = ipix * (d->c / V);
    if constexpr (sizeof(T) == 2) {
      const bool f32_ok = (!(d->flags & IVF_EP_OUT_F32) || (reinterpret_cast<uintptr_t>(dx) & 15) == 0) &&
                          (reinterpret_cast<uintptr_t>(mask_scale) & 15) == 0;
      if (fits31(d) && f32_ok && d->in_ld % 4 == 0) {
        const int rows = d->n * d->id * d->ih;
        IVF_POOL_GEO_DISPATCH((maxpool_bwd_bf16x8_kernel<G><<<row_grid(rows, d->iw * (d->c / V)), threads, 0, st>>>(
            *d, (const __nv_bfloat16*)dy, argmax, acc_in, (const __nv_bfloat16*)mask_y, mask_scale, dx, rows)));
        IVF_LAUNCHED(h);
        return IVF_OK;
      }
    }
    IVF_POOL_GEO_DISPATCH((maxpool_bwd_kernel<T, V, G><<<pool_blocks(h, total), threads, 0, st>>>(
        *d, (const T*)dy, argmax, acc_in, (const T*)mask_y, mask_scale, dx, total)));
  } else {
    long long total = ipix * d->c;
    maxpool_bwd_kernel<T, 1, PoolGeoDyn><<<pool_blocks(h, total), threads, 0, st>>>(
        *d, (const T*)dy, argmax, acc_in, (const T*)mask_y, mask_scale, dx, total);
  }
  IVF_LAUNCHED(h);
  return IVF_OK;
}

}  // namespace

extern "C" int ivf_maxpool3d_fwd(ivf_handle* h, const ivf_pool_desc* d, const void* in, void* out,
                                 uint8_t* argmax, void* stream) {
  IVF_REQUIRE(h && d && in && out, "ivf_maxpool3d_fwd: null argument");
  int rc = check_pool(d);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  return d->dtype == IVF_F32 ? fwd_t<float>(h, d, in, out, argmax, st)
                             : fwd_t<__nv_bfloat16>(h, d, in, out, argmax, st);
}

extern "C" int ivf_maxpool3d_bwd(ivf_handle* h, const ivf_pool_desc* d, const void* dy,
                                 const uint8_t* argmax, const float* acc_in, const void* mask_y,
                                 const float* mask_scale, void* dx, void* stream) {
  IVF_REQUIRE(h && d && dy && argmax && dx, "ivf_maxpool3d_bwd: null argument");
  int rc = check_pool(d);
  if (rc) return rc;
  if (d->flags & IVF_EP_ACCUM) IVF_REQUIRE(acc_in, "ivf_maxpool3d_bwd: ACCUM needs acc_in");
  if (d->flags & IVF_EP_MASK) IVF_REQUIRE(mask_y && mask_scale, "ivf_maxpool3d_bwd: MASK needs mask_y/mask_scale");
  cudaStream_t st = (cudaStream_t)stream;
  return d->dtype == IVF_F32
             ? bwd_t<float>(h, d, dy, argmax, acc_in, mask_y, mask_scale, dx, st)
             : bwd_t<__nv_bfloat16>(h, d, dy, argmax, acc_in, mask_y, mask_scale, dx, st);
}
