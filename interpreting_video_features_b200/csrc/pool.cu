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
struct Vec {
  T v[VEC];
};

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

template <typename T, int VEC>
__global__ void maxpool_fwd_kernel(ivf_pool_desc d, const T* __restrict__ in, T* __restrict__ out,
                                   uint8_t* __restrict__ argmax, long long total) {
  const int cv = d.c / VEC;
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
    int tap = 0;
    for (int a = 0; a < d.kd; ++a) {
      int zd = od * d.sd - d.pd + a;
      for (int b = 0; b < d.kh; ++b) {
        int zh = oh * d.sh - d.ph + b;
        for (int e = 0; e < d.kw; ++e, ++tap) {
          int zw = ow * d.sw - d.pw + e;
          float v[VEC];
          if (zd >= 0 && zd < d.id && zh >= 0 && zh < d.ih && zw >= 0 && zw < d.iw) {
            size_t pix = (((size_t)n * d.id + zd) * d.ih + zh) * d.iw + zw;
            load_vec<T, VEC>(in + pix * d.in_ld + d.in_coff + c, v);
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
#pragma unroll
      for (int i = 0; i < VEC; ++i) argmax[(size_t)opix * d.c + c + i] = (uint8_t)bidx[i];
    }
  }
}

template <typename T, int VEC>
__global__ void maxpool_bwd_kernel(ivf_pool_desc d, const T* __restrict__ dy,
                                   const uint8_t* __restrict__ argmax,
                                   const float* __restrict__ acc_in, const T* __restrict__ mask_y,
                                   const float* __restrict__ mask_scale, void* __restrict__ dx,
                                   long long total) {
  const int cv = d.c / VEC;
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
    int tap = 0;
    for (int a = 0; a < d.kd; ++a) {
      int nd = idd + d.pd - a;
      for (int b = 0; b < d.kh; ++b) {
        int nh = ih + d.ph - b;
        for (int e = 0; e < d.kw; ++e, ++tap) {
          int nw = iw + d.pw - e;
          if (nd < 0 || nh < 0 || nw < 0) continue;
          if (nd % d.sd || nh % d.sh || nw % d.sw) continue;
          int od = nd / d.sd, oh = nh / d.sh, ow = nw / d.sw;
          if (od >= d.od || oh >= d.oh || ow >= d.ow) continue;
          size_t opix = (((size_t)n * d.od + od) * d.oh + oh) * d.ow + ow;
          float v[VEC];
          load_vec<T, VEC>(dy + opix * d.out_ld + d.out_coff + c, v);
          const uint8_t* am = argmax + opix * d.c + c;
#pragma unroll
          for (int i = 0; i < VEC; ++i)
            if (am[i] == tap) g[i] += v[i];
        }
      }
    }
    size_t o = (size_t)ipix * d.in_ld + d.in_coff + c;
    if (d.flags & IVF_EP_ACCUM) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) g[i] += acc_in[o + i];
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

template <typename T>
int fwd_t(ivf_handle* h, const ivf_pool_desc* d, const void* in, void* out, uint8_t* argmax,
          cudaStream_t st) {
  constexpr int V = full_vec<T>();
  long long opix = (long long)d->n * d->od * d->oh * d->ow;
  const int threads = 256;
  if (vec_ok(d, V, in, out, nullptr)) {
    long long total = opix * (d->c / V);
    int blocks = (int)std::min<long long>((total + threads - 1) / threads, (long long)h->sm_count * 32);
    maxpool_fwd_kernel<T, V><<<blocks, threads, 0, st>>>(*d, (const T*)in, (T*)out, argmax, total);
  } else {
    long long total = opix * d->c;
    int blocks = (int)std::min<long long>((total + threads - 1) / threads, (long long)h->sm_count * 32);
    maxpool_fwd_kernel<T, 1><<<blocks, threads, 0, st>>>(*d, (const T*)in, (T*)out, argmax, total);
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
  bool v_ok = vec_ok(d, V, dy, mask_y, (d->flags & IVF_EP_OUT_F32) ? nullptr : dx);
  if (v_ok) {
    long long total = ipix * (d->c / V);
    int blocks = (int)std::min<long long>((total + threads - 1) / threads, (long long)h->sm_count * 32);
    maxpool_bwd_kernel<T, V><<<blocks, threads, 0, st>>>(*d, (const T*)dy, argmax, acc_in,
                                                         (const T*)mask_y, mask_scale, dx, total);
  } else {
    long long total = ipix * d->c;
    int blocks = (int)std::min<long long>((total + threads - 1) / threads, (long long)h->sm_count * 32);
    maxpool_bwd_kernel<T, 1><<<blocks, threads, 0, st>>>(*d, (const T*)dy, argmax, acc_in,
                                                         (const T*)mask_y, mask_scale, dx, total);
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
