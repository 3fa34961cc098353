// ConvLSTM element-wise stages: fused gate non-linearities + state update (forward and BPTT
// backward) and the per-step BatchNorm2d(eval) + MaxPool2d(2) (forward and backward).
// Reference: pt/models/convolution_lstm.py:38-48 —
//   i = sig(Wxi x + Whi h + c*Wci), f = sig(Wxf x + Whf h + c*Wcf),
//   c' = f*c + i*tanh(Wxc x + Whc h), o = sig(Wxo x + Who h + c'*Wco), h' = o*tanh(c')
// with Wci/Wcf/Wco constant zero tensors (:50-54), and :120-124 dropout(eval) -> shared bn -> mp.
// The 8 gate convolutions arrive here already summed: `pre` = x-conv (all steps batched, bias in
// the epilogue) + h-conv (accumulated in the conv epilogue), gate order [i | f | c | o] along
// the channel axis.  Bandwidth-bound, coalesced along channels.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigm(float v) { return 1.f / (1.f + expf(-v)); }

template <typename T>
__global__ void gates_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ c_prev,
                                 long long m, int hid, float* __restrict__ c_next,
                                 T* __restrict__ h_next, float* __restrict__ gate_act) {
  const long long total = m * hid;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / hid;
    int k = (int)(idx - r * hid);
    const float* p = pre + r * 4 * hid;
    float gi = sigm(p[k]);
    float gf = sigm(p[hid + k]);
    float gg = tanhf(p[2 * hid + k]);
    float go = sigm(p[3 * hid + k]);
    float cp = c_prev ? c_prev[idx] : 0.f;
    float cn = gf * cp + gi * gg;
    c_next[idx] = cn;
    h_next[idx] = ivf_from_float<T>(go * tanhf(cn));
    if (gate_act) {
      float* a = gate_act + r * 4 * hid;
      a[k] = gi;
      a[hid + k] = gf;
      a[2 * hid + k] = gg;
      a[3 * hid + k] = go;
    }
  }
}

template <typename T>
__global__ void gates_bwd_kernel(const float* __restrict__ gate_act, const float* __restrict__ c_prev,
                                 const float* __restrict__ c_next, const float* __restrict__ dh,
                                 float* __restrict__ dc_io, long long m, int hid,
                                 T* __restrict__ dgates) {
  const long long total = m * hid;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / hid;
    int k = (int)(idx - r * hid);
    const float* a = gate_act + r * 4 * hid;
    float gi = a[k], gf = a[hid + k], gg = a[2 * hid + k], go = a[3 * hid + k];
    float cp = c_prev ? c_prev[idx] : 0.f;
    float tc = tanhf(c_next[idx]);
    float dhv = dh[idx];
    float dcn = dhv * go * (1.f - tc * tc) + dc_io[idx];
    float d_o = dhv * tc;
    float d_i = dcn * gg, d_g = dcn * gi, d_f = dcn * cp;
    dc_io[idx] = dcn * gf;
    T* d = dgates + r * 4 * hid;
    d[k] = ivf_from_float<T>(d_i * gi * (1.f - gi));
    d[hid + k] = ivf_from_float<T>(d_f * gf * (1.f - gf));
    d[2 * hid + k] = ivf_from_float<T>(d_g * (1.f - gg * gg));
    d[3 * hid + k] = ivf_from_float<T>(d_o * go * (1.f - go));
  }
}

template <typename T>
__global__ void bn_pool2d_fwd_kernel(const T* __restrict__ x, int hh, int ww, int c,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     T* __restrict__ y, uint8_t* __restrict__ argmax, int s2d,
                                     long long total) {
  const int ho = hh / 2, wo = ww / 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int k = (int)(idx % c);
    long long t = idx / c;
    int ox = (int)(t % wo);
    t /= wo;
    int oy = (int)(t % ho);
    long long n = t / ho;
    float s = scale ? scale[k] : 1.f, b = shift ? shift[k] : 0.f;
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float v = ivf_to_float(x[((n * hh + 2 * oy + a) * ww + 2 * ox + e) * c + k]);
        v = fmaf(v, s, b);
        if (v > best || v != v) {
          best = v;
          bi = a * 2 + e;
        }
      }
    // s2d: [n][ho/2][wo/2][4c], channel = ((oy&1)*2 + (ox&1))*c + k — the operand layout of the next
    // layer's stride-2 x-convolution presented as a stride-1 3x3 convolution
    long long yo = s2d ? ((((n * (ho / 2) + (oy >> 1)) * (wo / 2) + (ox >> 1)) * 4 + ((oy & 1) * 2 + (ox & 1))) * c + k)
                       : idx;
    y[yo] = ivf_from_float<T>(best);
    argmax[idx] = (uint8_t)bi;
  }
}

template <typename T>
__global__ void bn_pool2d_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ argmax, int hh,
                                     int ww, int c, const float* __restrict__ scale,
                                     const float* __restrict__ acc_in, float* __restrict__ dx, int s2d,
                                     long long total) {
  const int ho = hh / 2, wo = ww / 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int k = (int)(idx % c);
    long long t = idx / c;
    int ix = (int)(t % ww);
    t /= ww;
    int iy = (int)(t % hh);
    long long n = t / hh;
    float g = 0.f;
    int oy = iy / 2, ox = ix / 2;
    if (oy < ho && ox < wo) {
      long long o = ((n * ho + oy) * wo + ox) * c + k;
      long long yo = s2d ? ((((n * (ho / 2) + (oy >> 1)) * (wo / 2) + (ox >> 1)) * 4 + ((oy & 1) * 2 + (ox & 1))) * c + k)
                         : o;
      if (argmax[o] == (iy & 1) * 2 + (ix & 1)) g = ivf_to_float(dy[yo]) * (scale ? scale[k] : 1.f);
    }
    if (acc_in) g += acc_in[idx];
    dx[idx] = g;
  }
}

int grid_for(ivf_handle* h, long long total) {
  long long b = (total + 255) / 256;
  long long cap = (long long)h->sm_count * 32;
  return (int)(b < cap ? b : cap);
}

}  // namespace

extern "C" int ivf_clstm_gates_fwd(ivf_handle* h, int dtype, const float* pre, const float* c_prev,
                                   int m, int hid, float* c_next, void* h_next, float* gate_act,
                                   void* stream) {
  IVF_REQUIRE(h && pre && c_next && h_next, "ivf_clstm_gates_fwd: null argument");
  IVF_REQUIRE(m > 0 && hid > 0, "ivf_clstm_gates_fwd: bad extent");
  long long total = (long long)m * hid;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IVF_F32)
    gates_fwd_kernel<float><<<grid_for(h, total), 256, 0, st>>>(pre, c_prev, m, hid, c_next,
                                                                (float*)h_next, gate_act);
  else if (dtype == IVF_BF16)
    gates_fwd_kernel<__nv_bfloat16><<<grid_for(h, total), 256, 0, st>>>(
        pre, c_prev, m, hid, c_next, (__nv_bfloat16*)h_next, gate_act);
  else
    IVF_FAIL(IVF_EINVAL, "ivf_clstm_gates_fwd: unknown dtype");
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_clstm_gates_bwd(ivf_handle* h, int dtype, const float* gate_act,
                                   const float* c_prev, const float* c_next, const float* dh,
                                   float* dc_io, int m, int hid, void* dgates, void* stream) {
  IVF_REQUIRE(h && gate_act && c_next && dh && dc_io && dgates, "ivf_clstm_gates_bwd: null argument");
  IVF_REQUIRE(m > 0 && hid > 0, "ivf_clstm_gates_bwd: bad extent");
  long long total = (long long)m * hid;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IVF_F32)
    gates_bwd_kernel<float><<<grid_for(h, total), 256, 0, st>>>(gate_act, c_prev, c_next, dh, dc_io, m,
                                                                hid, (float*)dgates);
  else if (dtype == IVF_BF16)
    gates_bwd_kernel<__nv_bfloat16><<<grid_for(h, total), 256, 0, st>>>(
        gate_act, c_prev, c_next, dh, dc_io, m, hid, (__nv_bfloat16*)dgates);
  else
    IVF_FAIL(IVF_EINVAL, "ivf_clstm_gates_bwd: unknown dtype");
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_bn_pool2d_fwd(ivf_handle* h, int dtype, const void* x, int n, int hh, int ww, int c,
                                 const float* scale, const float* shift, void* y, uint8_t* argmax,
                                 int s2d, void* stream) {
  IVF_REQUIRE(h && x && y && argmax, "ivf_bn_pool2d_fwd: null argument");
  IVF_REQUIRE(n > 0 && hh >= 2 && ww >= 2 && c > 0, "ivf_bn_pool2d_fwd: bad extent");
  if (s2d) IVF_REQUIRE((hh / 2) % 2 == 0 && (ww / 2) % 2 == 0, "ivf_bn_pool2d_fwd: s2d needs an even pooled map");
  long long total = (long long)n * (hh / 2) * (ww / 2) * c;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IVF_F32)
    bn_pool2d_fwd_kernel<float><<<grid_for(h, total), 256, 0, st>>>((const float*)x, hh, ww, c, scale,
                                                                    shift, (float*)y, argmax, s2d, total);
  else if (dtype == IVF_BF16)
    bn_pool2d_fwd_kernel<__nv_bfloat16><<<grid_for(h, total), 256, 0, st>>>(
        (const __nv_bfloat16*)x, hh, ww, c, scale, shift, (__nv_bfloat16*)y, argmax, s2d, total);
  else
    IVF_FAIL(IVF_EINVAL, "ivf_bn_pool2d_fwd: unknown dtype");
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_bn_pool2d_bwd(ivf_handle* h, int dtype, const void* dy, const uint8_t* argmax, int n,
                                 int hh, int ww, int c, const float* scale, const float* acc_in,
                                 float* dx, int s2d, void* stream) {
  IVF_REQUIRE(h && dy && argmax && dx, "ivf_bn_pool2d_bwd: null argument");
  IVF_REQUIRE(n > 0 && hh >= 2 && ww >= 2 && c > 0, "ivf_bn_pool2d_bwd: bad extent");
  long long total = (long long)n * hh * ww * c;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IVF_F32)
    bn_pool2d_bwd_kernel<float><<<grid_for(h, total), 256, 0, st>>>((const float*)dy, argmax, hh, ww, c,
                                                                    scale, acc_in, dx, s2d, total);
  else if (dtype == IVF_BF16)
    bn_pool2d_bwd_kernel<__nv_bfloat16><<<grid_for(h, total), 256, 0, st>>>(
        (const __nv_bfloat16*)dy, argmax, hh, ww, c, scale, acc_in, dx, s2d, total);
  else
    IVF_FAIL(IVF_EINVAL, "ivf_bn_pool2d_bwd: unknown dtype");
  IVF_LAUNCHED(h);
  return IVF_OK;
}
