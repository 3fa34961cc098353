// fp32 CUDA-core implicit-GEMM convolution: the "fp32 mode" of the hot path (1e-4 parity)
// and the on-device cross-check of the tcgen05 path.  Handles every geometry ivf_conv_desc can
// describe, including the transposed (strided data-gradient) gather.
//
// Reference semantics: pt/models/I3D_doubled.py:83-118 (Unit3D: explicit asymmetric zero pad,
// conv3d without bias, eval BatchNorm3d, ReLU) and autograd's convolution_backward for the
// data gradient.
#include "common.cuh"

namespace {

struct F32ConvParams {
  ivf_conv_desc d;
  int M;  // n*od*oh*ow
  int K;  // taps*cin
};

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_igemm_f32_kernel(F32ConvParams p, const float* __restrict__ in, const float* __restrict__ w,
                      const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ acc_in, const float* __restrict__ mask_y,
                      const float* __restrict__ mask_scale, float* __restrict__ out) {
  constexpr int NT = (BM / TM) * (BN / TN);
  static_assert(NT % BK == 0, "thread count must be a multiple of BK");
  constexpr int ROWS_PER_PASS = NT / BK;
  constexpr int A_PASSES = (BM + ROWS_PER_PASS - 1) / ROWS_PER_PASS;

  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int rowN[BM], rowD[BM], rowH[BM], rowW[BM];

  const ivf_conv_desc& d = p.d;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // decode the output pixel of every tile row once
  for (int r = tid; r < BM; r += NT) {
    int m = m0 + r;
    if (m < p.M) {
      int ow = m % d.ow;
      int t = m / d.ow;
      int oh = t % d.oh;
      t /= d.oh;
      int od = t % d.od;
      int n = t / d.od;
      rowN[r] = n;
      rowD[r] = od;
      rowH[r] = oh;
      rowW[r] = ow;
    } else {
      rowN[r] = -1;
      rowD[r] = rowH[r] = rowW[r] = 0;
    }
  }
  __syncthreads();

  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);

  // Blocked summation: products are accumulated in `part` over 64 consecutive k and then folded
  // into `acc`, so the rounding error grows like 64 + K/64 instead of K (K is up to 3 456 here; the
  // class gradient of the mask search is a cancellation-heavy 1e-9 quantity).
  float acc[TM][TN], part[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = part[i][j] = 0.f;

  const int a_k = tid % BK;
  const int a_r = tid / BK;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    // ---- A tile: gather ----
    {
      int k = k0 + a_k;
      bool kvalid = k < p.K;
      int tap = kvalid ? k / d.cin : 0;
      int c = kvalid ? k - tap * d.cin : 0;
      int kw_i = tap % d.kw;
      int t2 = tap / d.kw;
      int kh_i = t2 % d.kh;
      int kd_i = t2 / d.kh;
#pragma unroll
      for (int ps = 0; ps < A_PASSES; ++ps) {
        int r = a_r + ps * ROWS_PER_PASS;
        if (r < BM) {
          float v = 0.f;
          int n = rowN[r];
          if (kvalid && n >= 0) {
            int zd, zh, zw;
            bool ok = true;
            if (!d.transposed) {
              zd = rowD[r] * d.sd - d.pd + kd_i;
              zh = rowH[r] * d.sh - d.ph + kh_i;
              zw = rowW[r] * d.sw - d.pw + kw_i;
            } else {
              int nd = rowD[r] + d.pd - kd_i;
              int nh = rowH[r] + d.ph - kh_i;
              int nw = rowW[r] + d.pw - kw_i;
              ok = nd >= 0 && nh >= 0 && nw >= 0 && (nd % d.sd == 0) && (nh % d.sh == 0) &&
                   (nw % d.sw == 0);
              zd = nd / d.sd;
              zh = nh / d.sh;
              zw = nw / d.sw;
            }
            ok = ok && zd >= 0 && zd < d.id && zh >= 0 && zh < d.ih && zw >= 0 && zw < d.iw;
            if (ok) {
              size_t pix = (((size_t)n * d.id + zd) * d.ih + zh) * d.iw + zw;
              v = __ldg(in + pix * d.in_ld + d.in_coff + c);
            }
          }
          As[a_k][r] = v;
        }
      }
    }
    // ---- B tile ----
    for (int e = tid; e < BK * BN; e += NT) {
      int kk = e / BN, nn = e % BN;
      int k = k0 + kk, n = n0 + nn;
      Bs[kk][nn] = (k < p.K && n < d.cout) ? __ldg(w + (size_t)k * d.cout + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
    if (((k0 / BK) & 3) == 3 || k0 + BK >= p.K) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[i][j] += part[i][j];
          part[i][j] = 0.f;
        }
    }
    __syncthreads();
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= d.cout) continue;
      float v = acc[i][j];
      size_t o = (size_t)m * d.out_ld + d.out_coff + n;
      if (d.flags & IVF_EP_ACCUM) v += acc_in[o];
      if (d.flags & IVF_EP_AFFINE) v = fmaf(v, scale[n], shift[n]);
      if (d.flags & IVF_EP_RELU) v = fmaxf(v, 0.f);
      if (d.flags & IVF_EP_MASK) {
        float y = mask_y[(size_t)m * d.mask_ld + d.mask_coff + n];
        v = y > 0.f ? v * mask_scale[n] : 0.f;
      }
      out[o] = v;
    }
  }
}

}  // namespace

int ivf_conv3d_f32_launch(ivf_handle* h, const ivf_conv_desc* d, const float* in, const float* w,
                          const float* scale, const float* shift, const float* acc_in,
                          const float* mask_y, const float* mask_scale, float* out,
                          cudaStream_t st) {
  F32ConvParams p;
  p.d = *d;
  long long M = (long long)d->n * d->od * d->oh * d->ow;
  long long K = (long long)d->kd * d->kh * d->kw * d->cin;
  IVF_REQUIRE(M < (1ll << 31) && K < (1ll << 31), "ivf_conv3d(f32): problem too large");
  p.M = (int)M;
  p.K = (int)K;
  if (d->cout <= 8) {
    constexpr int BM = 256, BN = 8, TM = 4, TN = 2;
    dim3 grid(ivf_cdiv(M, BM), ivf_cdiv(d->cout, BN));
    conv_igemm_f32_kernel<BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(
        p, in, w, scale, shift, acc_in, mask_y, mask_scale, out);
  } else if (d->cout <= 32) {
    constexpr int BM = 128, BN = 32, TM = 4, TN = 4;
    dim3 grid(ivf_cdiv(M, BM), ivf_cdiv(d->cout, BN));
    conv_igemm_f32_kernel<BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(
        p, in, w, scale, shift, acc_in, mask_y, mask_scale, out);
  } else {
    constexpr int BM = 64, BN = 64, TM = 4, TN = 4;
    dim3 grid(ivf_cdiv(M, BM), ivf_cdiv(d->cout, BN));
    conv_igemm_f32_kernel<BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(
        p, in, w, scale, shift, acc_in, mask_y, mask_scale, out);
  }
  IVF_LAUNCHED(h);
  return IVF_OK;
}
