// Model-load kernels: weight packing (OIDHW fp32 -> the operand layouts of the convolution kernels),
// eval-BatchNorm folding, buffer fills and the one-hot class selector.  They run once per engine, not per
// iteration; they are kernels (not host loops over torch ops) so that building an engine is a few dozen
// launches instead of >1 000 tiny elementwise ones, and so that every device operation of the path is libivf's.
//
// Reference call sites: nn.Conv3d / nn.BatchNorm3d parameters of pt/models/I3D_doubled.py:66-75 (read by ATen's
// cuDNN conv and batch_norm there); nn.Conv2d gate weights of pt/models/convolution_lstm.py:25-32; one_hot of
// pt/grad_cam_videos.py:73-79 and the class_loss selector of pt/FindMasksComparison_I3D_smth.py:205.
#include "common.cuh"

namespace {

struct PackParams {
  int co, ci, kd, kh, kw;   // source OIDHW
  int cis;                  // operand channels per parity block (>= ci)
  int fd, fh, fw;           // space-to-depth factors (1 or 2)
  int wd, wh, ww;           // taps of the (space-to-depth) operand
  int ceff;                 // fd*fh*fw*ci
  int swap, flip, layout;
  int n_pad, k_pad, n_off, k_off, n_stride, k_stride;
  int taps;                 // wd*wh*ww
};

// one thread per (n_local, tap, k_local) of the source block
template <typename OutT>
__global__ void pack_weights_kernel(PackParams p, const float* __restrict__ src, OutT* __restrict__ dst) {
  const int nsrc = p.swap ? p.ceff : p.co;
  const int ksrc = p.swap ? p.co : p.ceff;
  const long long total = (long long)nsrc * p.taps * ksrc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % ksrc);
    const int tap = (int)((i / ksrc) % p.taps);
    const int n = (int)(i / ((long long)ksrc * p.taps));
    const int co = p.swap ? k : n;
    const int ce = p.swap ? n : k;
    int dw = tap % p.ww, dh = (tap / p.ww) % p.wh, dt = tap / (p.ww * p.wh);
    if (p.flip) {  // data gradient of a stride-1 convolution: taps flipped
      dw = p.ww - 1 - dw;
      dh = p.wh - 1 - dh;
      dt = p.wd - 1 - dt;
    }
    // operand channel -> (parity a,b,c ; source channel)
    const int ch = ce % p.cis;
    const int par = ce / p.cis;
    const int c_ = par % p.fw, b_ = (par / p.fw) % p.fh, a_ = par / (p.fw * p.fh);
    const int kt = p.fd * dt + a_, khh = p.fh * dh + b_, kww = p.fw * dw + c_;
    float v = 0.f;
    if (ch < p.ci && kt < p.kd && khh < p.kh && kww < p.kw)
      v = src[((((long long)co * p.ci + ch) * p.kd + kt) * p.kh + khh) * p.kw + kww];
    const long long nn = p.n_off + (long long)n * p.n_stride, kk = p.k_off + (long long)k * p.k_stride;
    long long o;
    if (p.layout == IVF_PACK_KMAJOR) o = (nn * p.taps + tap) * p.k_pad + kk;
    else o = ((long long)tap * p.k_pad + kk) * p.n_pad + nn;
    dst[o] = ivf_from_float<OutT>(v);
  }
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               const float* __restrict__ conv_bias, int c, float* __restrict__ scale,
                               float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float s = 1.f, sh = 0.f;
  if (gamma) {
    s = gamma[i] / sqrtf(var[i] + eps);
    sh = beta[i] - mean[i] * s;
  }
  if (conv_bias) sh += s * conv_bias[i];
  scale[i] = s;
  shift[i] = sh;
}

__global__ void fill_u32_kernel(uint32_t* __restrict__ p, size_t words, uint32_t v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
}

// uint8 frames (what the loaders decode, pt/data_loader_jpg.py:27-37 before .float()) -> fp32 0..255, 16 per thread
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t count) {
  const size_t nvec = count / 16;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(src)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float4* o = reinterpret_cast<float4*>(dst) + i * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = make_float4((float)(w[j] & 0xff), (float)((w[j] >> 8) & 0xff), (float)((w[j] >> 16) & 0xff),
                         (float)(w[j] >> 24));
  }
  for (size_t i = nvec * 16 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}

__global__ void one_hot_kernel(const int* __restrict__ targets, int n, int ncls, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * ncls) return;
  out[i] = (i % ncls) == targets[i / ncls] ? 1.f : 0.f;
}

// argmax over classes per row -> int32 targets (pt/grad_cam_videos.py:70-71 np.argmax of the output)
__global__ void argmax_rows_kernel(const float* __restrict__ x, int n, int ncls, int* __restrict__ out) {
  const int row = blockIdx.x;
  const int lane = threadIdx.x;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane; j < ncls; j += 32) {
    const float v = x[(size_t)row * ncls + j];
    if (v > best || (v == best && j < bi) || (v != v && !(best != best))) {  // first maximum; NaN wins like np.argmax
      best = v;
      bi = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const bool onan = ov != ov, bnan = best != best;
    bool take;
    if (onan || bnan) take = onan && (!bnan || oi < bi);
    else take = ov > best || (ov == best && oi < bi);
    if (take) {
      best = ov;
      bi = oi;
    }
  }
  if (lane == 0) out[row] = bi == 0x7fffffff ? 0 : bi;
}

}  // namespace

extern "C" int ivf_pack_weights(ivf_handle* h, const ivf_pack_desc* d, const float* src, void* dst, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && d && src && dst, "ivf_pack_weights: null argument");
  IVF_REQUIRE(d->co > 0 && d->ci > 0 && d->kd > 0 && d->kh > 0 && d->kw > 0, "ivf_pack_weights: bad source shape");
  IVF_REQUIRE((d->s2d_d == 1 || d->s2d_d == 2) && (d->s2d_h == 1 || d->s2d_h == 2) && (d->s2d_w == 1 || d->s2d_w == 2),
              "ivf_pack_weights: space-to-depth factors must be 1 or 2");
  IVF_REQUIRE(d->layout == IVF_PACK_KMAJOR || d->layout == IVF_PACK_TAPMAJOR, "ivf_pack_weights: unknown layout");
  IVF_REQUIRE(d->dtype == IVF_F32 || d->dtype == IVF_BF16, "ivf_pack_weights: unknown dtype");
  PackParams p;
  p.co = d->co; p.ci = d->ci; p.kd = d->kd; p.kh = d->kh; p.kw = d->kw;
  p.fd = d->s2d_d; p.fh = d->s2d_h; p.fw = d->s2d_w;
  p.wd = (d->kd + p.fd - 1) / p.fd; p.wh = (d->kh + p.fh - 1) / p.fh; p.ww = (d->kw + p.fw - 1) / p.fw;
  p.cis = d->ci_stride > 0 ? d->ci_stride : d->ci;
  IVF_REQUIRE(p.cis >= d->ci, "ivf_pack_weights: ci_stride below ci");
  p.ceff = p.fd * p.fh * p.fw * p.cis;
  IVF_REQUIRE(d->dgrad >= 0 && d->dgrad <= 2, "ivf_pack_weights: dgrad must be 0, 1 or 2");
  p.swap = d->dgrad != 0; p.flip = d->dgrad == 1; p.layout = d->layout;
  p.n_pad = d->n_pad; p.k_pad = d->k_pad; p.n_off = d->n_off; p.k_off = d->k_off;
  p.n_stride = d->n_stride > 0 ? d->n_stride : 1; p.k_stride = d->k_stride > 0 ? d->k_stride : 1;
  p.taps = p.wd * p.wh * p.ww;
  const int nsrc = p.swap ? p.ceff : p.co, ksrc = p.swap ? p.co : p.ceff;
  IVF_REQUIRE(d->n_off >= 0 && d->k_off >= 0 && d->n_off + (nsrc - 1) * p.n_stride < d->n_pad &&
                  d->k_off + (ksrc - 1) * p.k_stride < d->k_pad,
              "ivf_pack_weights: block %dx%d at (%d,%d) exceeds the packed matrix %dx%d", nsrc, ksrc, d->n_off,
              d->k_off, d->n_pad, d->k_pad);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t elem = d->dtype == IVF_BF16 ? 2 : 4;
  const size_t dst_bytes = (size_t)d->n_pad * p.taps * d->k_pad * elem;
  if (d->zero_first) {
    IVF_REQUIRE(dst_bytes % 4 == 0 && (uintptr_t)dst % 4 == 0, "ivf_pack_weights: destination not 4-byte sized");
    fill_u32_kernel<<<ivf_cdiv((long long)(dst_bytes / 4), 256 * 8) < 1184 ? ivf_cdiv((long long)(dst_bytes / 4), 256 * 8) : 1184,
                      256, 0, st>>>((uint32_t*)dst, dst_bytes / 4, 0u);
    IVF_LAUNCHED(h);
  }
  const long long total = (long long)nsrc * p.taps * ksrc;
  int blocks = ivf_cdiv(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (d->dtype == IVF_BF16)
    pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p, src, (__nv_bfloat16*)dst);
  else
    pack_weights_kernel<float><<<blocks, 256, 0, st>>>(p, src, (float*)dst);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_bn_fold(ivf_handle* h, const float* gamma, const float* beta, const float* mean, const float* var,
                           float eps, const float* conv_bias, int c, float* scale, float* shift, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && scale && shift && c > 0, "ivf_bn_fold: bad argument");
  if (gamma) IVF_REQUIRE(beta && mean && var, "ivf_bn_fold: gamma without beta/mean/var");
  bn_fold_kernel<<<ivf_cdiv(c, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, eps, conv_bias, c, scale,
                                                                     shift);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_fill_u32(ivf_handle* h, void* dst, size_t bytes, uint32_t pattern, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && dst, "ivf_fill_u32: null argument");
  IVF_REQUIRE(bytes % 4 == 0 && (uintptr_t)dst % 4 == 0, "ivf_fill_u32: 4-byte granularity");
  if (bytes == 0) return IVF_OK;
  const size_t words = bytes / 4;
  long long blocks = (long long)((words + 256 * 8 - 1) / (256 * 8));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  fill_u32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((uint32_t*)dst, words, pattern);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_u8_to_f32(ivf_handle* h, const uint8_t* src, float* dst, size_t count, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && src && dst, "ivf_u8_to_f32: null argument");
  IVF_REQUIRE((uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0, "ivf_u8_to_f32: 16-byte aligned buffers");
  if (count == 0) return IVF_OK;
  long long blocks = (long long)((count / 16 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  u8_to_f32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, count);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_one_hot(ivf_handle* h, const int* targets, int n, int ncls, float* out, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && targets && out && n > 0 && ncls > 0, "ivf_one_hot: bad argument");
  one_hot_kernel<<<ivf_cdiv((long long)n * ncls, 256), 256, 0, (cudaStream_t)stream>>>(targets, n, ncls, out);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_argmax_rows(ivf_handle* h, const float* x, int n, int ncls, int* out, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && x && out && n > 0 && ncls > 0, "ivf_argmax_rows: bad argument");
  argmax_rows_kernel<<<n, 32, 0, (cudaStream_t)stream>>>(x, n, ncls, out);
  IVF_LAUNCHED(h);
  return IVF_OK;
}
