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
          // skip the 16-byte gradient load when no channel of this vector selected this tap
          bool any = false;
#pragma unroll
          for (int i = 0; i < VEC; ++i) any |= (((i < 4 ? a0 : a1) >> (8 * (i & 3))) & 0xffu) == (unsigned)tap;
          if (!any) continue;
          float v[VEC];
          load_vec<T, VEC>(dy + opix * d.out_ld + d.out_coff + c, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i)
            if ((((i < 4 ? a0 : a1) >> (8 * (i & 3))) & 0xffu) == (unsigned)tap) g[i] += v[i];
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
