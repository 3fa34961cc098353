// Training step of the I3D classifier (pt/train_i3d_smth.py:192-250: model.train(), forward,
// CrossEntropyLoss, loss.backward(), optimizer.step()): what the interpretation path does not need and the
// training path adds - BatchNorm3d with batch statistics (pt/models/I3D_doubled.py:75, eps 1e-3, momentum 0.01)
// forward and backward, the convolution WEIGHT gradient, the classifier head with dropout and the loss, and the
// SGD / Adam update.  Forward convolutions, data gradients and max-pools are the kernels of the interpretation
// path (ivf_conv3d, ivf_maxpool3d_*).  All tensors are channels-last [pixels][ld] with a channel offset, like every
// activation of the engine; element type fp32 or bf16, statistics and gradients of parameters fp32.
//
// These are first-correct CUDA-core kernels (the weight gradient is a tiled fp32 outer-product GEMM with atomics
// across pixel splits, not a tcgen05 kernel): parity against torch autograd first, see DESIGN.md "Training step".
#include <limits.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ float ldf(const T* p, long long i) { return ivf_to_float(p[i]); }

// ---------------------------------------------------------------------------------------------------------
// BatchNorm (training): blocks of 32 channels x 8 row lanes; grid.x = channel groups, grid.y = row splits
constexpr int BN_ROWS = 8;

// Per-channel sums are accumulated in DOUBLE from the first addend on, like ATen's CPU BatchNorm (acc_type<float> =
// double): the backward sums cancel almost completely on this network (a BatchNorm bias that feeds another
// convolution + BatchNorm: sum g is ~1e-4 of sum |g|), and fp32 partial sums per thread left dbeta of such a tensor
// 2 % off.  The kernels are bound by the memory traffic, not by the fp64 adds.
__device__ __forceinline__ void bn_block_reduce2(double a, double b, double* dst_a, double* dst_b, int c, bool cok) {
  __shared__ double sa[BN_ROWS][33], sb[BN_ROWS][33];
  sa[threadIdx.y][threadIdx.x] = a;
  sb[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && cok) {
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int r = 0; r < BN_ROWS; ++r) {
      ta += sa[r][threadIdx.x];
      tb += sb[r][threadIdx.x];
    }
    atomicAdd(dst_a + c, ta);
    if (dst_b) atomicAdd(dst_b + c, tb);
  }
}

// ONE pass over z: ws[c] += sum (z - z0), ws[C + c] += sum (z - z0)^2 with z0 = the channel's first element as the
// shift.  In double both sums are exact to ~1e-16 of their size, so the centred variance E[(z-z0)^2] - E[z-z0]^2
// keeps ~12 digits even where |mean| >> sigma (the two-pass form read z a second time for the same digits).
template <typename T>
__global__ void __launch_bounds__(32 * BN_ROWS)
bn_stats_kernel(const T* __restrict__ z, int ld, int coff, long long m, int C, double* __restrict__ ws) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool cok = c < C;
  double s1 = 0.0, s2 = 0.0;
  if (cok) {
    const double z0 = (double)ldf(z, (long long)coff + c);
    for (long long r = (long long)blockIdx.y * BN_ROWS + threadIdx.y; r < m; r += (long long)gridDim.y * BN_ROWS) {
      const double v = (double)ldf(z, r * ld + coff + c) - z0;
      s1 += v;
      s2 += v * v;
    }
  }
  bn_block_reduce2(s1, s2, ws, ws + C, c, cok);
}

template <typename T>
__global__ void __launch_bounds__(32 * BN_ROWS)
bn_apply_kernel(const T* __restrict__ z, int z_ld, int z_coff, long long m, int C, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                float* __restrict__ running_var, float* __restrict__ save_mean, float* __restrict__ save_rstd,
                const double* __restrict__ ws, T* __restrict__ y, int y_ld, int y_coff, int relu) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (c >= C) return;
  const double d1 = ws[c] / (double)m;
  const double var_d = fmax(ws[C + c] / (double)m - d1 * d1, 0.0);  // biased: what normalises the batch
  const float mean = (float)((double)ldf(z, (long long)z_coff + c) + d1);
  const float var = (float)var_d;
  const float rstd = rsqrtf(var + eps);
  const float g = gamma[c] * rstd, b = beta[c] - mean * g;
  if (blockIdx.y == 0 && threadIdx.y == 0) {
    save_mean[c] = mean;
    save_rstd[c] = rstd;
    if (running_mean) {  // nn.BatchNorm3d: running_var takes the UNBIASED batch variance
      const float unb = m > 1 ? (float)(var_d * (double)m / (double)(m - 1)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
    }
  }
  for (long long r = (long long)blockIdx.y * BN_ROWS + threadIdx.y; r < m; r += (long long)gridDim.y * BN_ROWS) {
    float v = fmaf(ldf(z, r * z_ld + z_coff + c), g, b);
    if (relu) v = fmaxf(v, 0.f);
    y[r * y_ld + y_coff + c] = ivf_from_float<T>(v);
  }
}

// backward pass 1: ws[c] += sum g, ws[C + c] += sum g * xhat, g = dy * [y > 0]
template <typename T, typename TG>
__global__ void __launch_bounds__(32 * BN_ROWS)
bn_bwd_stats_kernel(const TG* __restrict__ dy, int dy_ld, int dy_coff, const T* __restrict__ y, int y_ld, int y_coff,
                    const T* __restrict__ z, int z_ld, int z_coff, long long m, int C,
                    const float* __restrict__ save_mean, const float* __restrict__ save_rstd, double* __restrict__ ws) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool cok = c < C;
  double sg = 0.0, sgx = 0.0;
  if (cok) {
    const float mean = save_mean[c], rstd = save_rstd[c];
    for (long long r = (long long)blockIdx.y * BN_ROWS + threadIdx.y; r < m; r += (long long)gridDim.y * BN_ROWS) {
      float g = ldf(dy, r * dy_ld + dy_coff + c);
      if (y && !(ldf(y, r * y_ld + y_coff + c) > 0.f)) g = 0.f;
      sg += (double)g;
      sgx += (double)g * (double)((ldf(z, r * z_ld + z_coff + c) - mean) * rstd);
    }
  }
  bn_block_reduce2(sg, sgx, ws, ws + C, c, cok);
}

// backward pass 2: dz = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat))
template <typename T, typename TG>
__global__ void __launch_bounds__(32 * BN_ROWS)
bn_bwd_apply_kernel(const TG* __restrict__ dy, int dy_ld, int dy_coff, const T* __restrict__ y, int y_ld, int y_coff,
                    const T* __restrict__ z, int z_ld, int z_coff, long long m, int C, const float* __restrict__ gamma,
                    const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                    const double* __restrict__ ws, T* __restrict__ dz, int dz_ld, int dz_coff,
                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (c >= C) return;
  const float mean = save_mean[c], rstd = save_rstd[c];
  const float sg = (float)ws[c], sgx = (float)ws[C + c];
  if (blockIdx.y == 0 && threadIdx.y == 0) {
    dbeta[c] = sg;
    dgamma[c] = sgx;
  }
  const float k = gamma[c] * rstd, mg = sg / (float)m, mgx = sgx / (float)m;
  for (long long r = (long long)blockIdx.y * BN_ROWS + threadIdx.y; r < m; r += (long long)gridDim.y * BN_ROWS) {
    float g = ldf(dy, r * dy_ld + dy_coff + c);
    if (y && !(ldf(y, r * y_ld + y_coff + c) > 0.f)) g = 0.f;
    const float xh = (ldf(z, r * z_ld + z_coff + c) - mean) * rstd;
    dz[r * dz_ld + dz_coff + c] = ivf_from_float<T>(k * (g - mg - xh * mgx));
  }
}

dim3 bn_grid(const ivf_handle* h, long long m, int C) {
  const int gx = (C + 31) / 32;
  long long gy = (m + BN_ROWS - 1) / BN_ROWS;
  const long long want = (long long)(h->sm_count * 8 + gx - 1) / gx;  // ~8 blocks per SM over all channel groups
  if (gy > want) gy = want;
  if (gy < 1) gy = 1;
  return dim3(gx, (unsigned)gy);
}

// ---------------------------------------------------------------------------------------------------------
// Convolution weight gradient as a GEMM over the pixels: dw[co][j] = sum_p dz[p][co] * col[p][j], where the columns
// j = tap * cin + ci run over the flattened (tap, input channel) axis - the im2col row of output pixel p, gathered
// on the fly (zero outside the tensor: the 'same' padding).  A block owns a 64 x 64 tile of (co, j) for a range of
// pixels and adds it to dw (fp32 OIDHW) with atomics; 256 threads = 16 x 16, each 4 x 4 outputs.  Tiling the
// flattened axis keeps the tiles full whatever cin is: the stem (cin 3, 343 taps) is 17 column tiles of 64 instead
// of 343 tiles holding three channels each, and a tile that spans taps still reads runs of (kw, ci) that are
// contiguous in the channels-last input.
constexpr int WG_PT = 16;  // pixels per shared-memory stage

template <typename TX, typename T>
__global__ void __launch_bounds__(256)
wgrad_kernel(ivf_conv_desc d, const TX* __restrict__ x, const T* __restrict__ dz, float* __restrict__ dw,
             int col_tiles, long long pix_per_block) {
  constexpr int TCO = 64, TCJ = 64;
  __shared__ float sdz[WG_PT][TCO + 4];
  __shared__ float sx[WG_PT][TCJ + 4];
  __shared__ int col_off[TCJ];              // element offset of the column's (tap, ci) from the pixel's window origin
  __shared__ int col_zyx[TCJ];              // kd | kh << 8 | kw << 16, -1: past the last column
  __shared__ long long pix_base[WG_PT];     // element offset of the window origin (may lie in the padding)
  __shared__ int pix_zyx[WG_PT][3];         // window origin coordinates; z = INT_MIN/2: no pixel
  __shared__ long long zoff[WG_PT];
  const int taps = d.kd * d.kh * d.kw, ncols = taps * d.cin;
  const int cot = blockIdx.x / col_tiles, colt = blockIdx.x - cot * col_tiles;
  const int co0 = cot * TCO, j0 = colt * TCJ;
  const long long P = (long long)d.n * d.od * d.oh * d.ow;
  const long long p_lo = (long long)blockIdx.y * pix_per_block;
  const long long p_hi = p_lo + pix_per_block < P ? p_lo + pix_per_block : P;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (threadIdx.x < TCJ) {
    const int j = j0 + threadIdx.x;
    if (j < ncols) {
      const int tap = j / d.cin, ci = j - tap * d.cin;
      const int kw_i = tap % d.kw, kh_i = (tap / d.kw) % d.kh, kd_i = tap / (d.kw * d.kh);
      col_off[threadIdx.x] = ((kd_i * d.ih + kh_i) * d.iw + kw_i) * d.in_ld + ci;
      col_zyx[threadIdx.x] = kd_i | (kh_i << 8) | (kw_i << 16);
    } else {
      col_off[threadIdx.x] = 0;
      col_zyx[threadIdx.x] = -1;
    }
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long p0 = p_lo; p0 < p_hi; p0 += WG_PT) {
    __syncthreads();  // the previous stage has been consumed (and the column table is complete)
    if (threadIdx.x < WG_PT) {
      const long long p = p0 + threadIdx.x;
      if (p < p_hi) {
        zoff[threadIdx.x] = p * d.out_ld + d.out_coff;
        const int ow = (int)(p % d.ow);
        long long t = p / d.ow;
        const int oh = (int)(t % d.oh);
        t /= d.oh;
        const int od = (int)(t % d.od);
        const int n = (int)(t / d.od);
        const int iz = od * d.sd - d.pd, iy = oh * d.sh - d.ph, ix = ow * d.sw - d.pw;
        pix_zyx[threadIdx.x][0] = iz;
        pix_zyx[threadIdx.x][1] = iy;
        pix_zyx[threadIdx.x][2] = ix;
        pix_base[threadIdx.x] = ((((long long)n * d.id + iz) * d.ih + iy) * d.iw + ix) * d.in_ld + d.in_coff;
      } else {
        zoff[threadIdx.x] = -1;
        pix_zyx[threadIdx.x][0] = INT_MIN / 2;
        pix_zyx[threadIdx.x][1] = pix_zyx[threadIdx.x][2] = 0;
        pix_base[threadIdx.x] = 0;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < WG_PT * TCO; i += 256) {
      const int pp = i / TCO, cc = i - pp * TCO;
      const long long zo = zoff[pp];
      sdz[pp][cc] = (zo >= 0 && co0 + cc < d.cout) ? ldf(dz, zo + co0 + cc) : 0.f;
    }
    for (int i = threadIdx.x; i < WG_PT * TCJ; i += 256) {
      const int pp = i / TCJ, cc = i - pp * TCJ;
      const int k = col_zyx[cc];
      float v = 0.f;
      if (k >= 0) {
        const int iz = pix_zyx[pp][0] + (k & 255), iy = pix_zyx[pp][1] + ((k >> 8) & 255), ix = pix_zyx[pp][2] + (k >> 16);
        if ((unsigned)iz < (unsigned)d.id && (unsigned)iy < (unsigned)d.ih && (unsigned)ix < (unsigned)d.iw)
          v = ldf(x, pix_base[pp] + col_off[cc]);
      }
      sx[pp][cc] = v;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < WG_PT; ++pp) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sdz[pp][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sx[pp][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= d.cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = j0 + tx * 4 + j;
      if (col >= ncols || acc[i][j] == 0.f) continue;
      const int tap = col / d.cin, ci = col - tap * d.cin;
      atomicAdd(dw + ((long long)co * d.cin + ci) * taps + tap, acc[i][j]);
    }
  }
}

// The same GEMM on the tensor cores for bf16 gradients (mixed-precision step): operands staged as bf16 in shared
// memory ([pixel][co] and [pixel][column], the layouts they have in global memory), fragments fetched with
// ldmatrix.trans (both operands are "K-rows" here: the reduction index, the pixel, is the slow one), fp32
// accumulators in registers (mma.sync.m16n8k16; a tcgen05 kernel needs both operands MN-major and is the next step).
// Block = 8 warps on a 64 x 64 tile: warp (wm, wn) owns rows 16*wm.. and columns 32*wn..; 32 pixels per stage.
// Rows are padded to 72 elements (144 bytes) so that the eight 16-byte rows of an ldmatrix tile fall into
// different bank groups.
constexpr int WM_PT = 64, WM_LD = 72;  // pixels per stage (two 32-pixel halves per thread), padded row length

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ld8_bf16(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ uint4 ld8_bf16(const float* p) {  // fp32 source (the stem's clip): rounded while staging
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  __nv_bfloat162 q0 = __floats2bfloat162_rn(a.x, a.y), q1 = __floats2bfloat162_rn(a.z, a.w);
  __nv_bfloat162 q2 = __floats2bfloat162_rn(b.x, b.y), q3 = __floats2bfloat162_rn(b.z, b.w);
  return make_uint4(*reinterpret_cast<uint32_t*>(&q0), *reinterpret_cast<uint32_t*>(&q1),
                    *reinterpret_cast<uint32_t*>(&q2), *reinterpret_cast<uint32_t*>(&q3));
}

// VEC: cin % 8 == 0 and every channel offset % 8 == 0 - eight consecutive columns are eight consecutive channels of
// one tap, fetched as one 16-byte vector; otherwise element by element (the stem: cin 3).
// S2D: d describes the stride-1 convolution over the 2x2x2 space-to-depth record of a stride-2 layer (the stem's
// tensor-core form: operand channel = ((a*2 + b)*2 + c)*ci + ch, operand tap t per axis, source tap = 2*t + parity,
// as ivf_pack_weights lays the weights out); the result is written in the ORIGINAL [cout][ci][kd][kh][kw] layout,
// taps past the kernel (the eighth of a 7-tap axis) dropped.  og = {ci, kd, kh, kw} of the original layer.
template <typename TX, bool VEC, bool S2D = false>
__global__ void __launch_bounds__(256)
wgrad_mma_kernel(ivf_conv_desc d, const TX* __restrict__ x, const __nv_bfloat16* __restrict__ dz,
                 float* __restrict__ dw, int col_tiles, long long pix_per_block, int4 og = make_int4(0, 0, 0, 0)) {
  __shared__ __align__(16) __nv_bfloat16 sdz[WM_PT][WM_LD];
  __shared__ __align__(16) __nv_bfloat16 sx[WM_PT][WM_LD];
  __shared__ int col_off[64];
  __shared__ int col_zyx[64];
  const int taps = d.kd * d.kh * d.kw, ncols = taps * d.cin;
  const int cot = blockIdx.x / col_tiles, colt = blockIdx.x - cot * col_tiles;
  const int co0 = cot * 64, j0 = colt * 64;
  const long long P = (long long)d.n * d.od * d.oh * d.ow;
  const long long p_lo = (long long)blockIdx.y * pix_per_block;
  const long long p_hi = p_lo + pix_per_block < P ? p_lo + pix_per_block : P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 1, wn = warp & 1;
  if (threadIdx.x < 64) {
    const int j = j0 + threadIdx.x;
    if (j < ncols) {
      const int tap = j / d.cin, ci = j - tap * d.cin;
      const int kw_i = tap % d.kw, kh_i = (tap / d.kw) % d.kh, kd_i = tap / (d.kw * d.kh);
      col_off[threadIdx.x] = ((kd_i * d.ih + kh_i) * d.iw + kw_i) * d.in_ld + ci;
      col_zyx[threadIdx.x] = kd_i | (kh_i << 8) | (kw_i << 16);
    } else {
      col_off[threadIdx.x] = 0;
      col_zyx[threadIdx.x] = -1;
    }
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // ldmatrix row addresses of this lane.  A (dz^T): matrices (k 0-7, m 0-7), (k 0-7, m 8-15), (k 8-15, m 0-7),
  // (k 8-15, m 8-15) -> a0..a3.  B: two n8 tiles per x4: (k 0-7, n 0-7), (k 8-15, n 0-7), (k 0-7, n 8-15), (k 8-15, n 8-15).
  const int lr = lane & 7, lm = lane >> 3;
  const uint32_t a_addr = (uint32_t)__cvta_generic_to_shared(&sdz[(lm >> 1) * 8 + lr][wm * 16 + (lm & 1) * 8]);
  const uint32_t b_addr = (uint32_t)__cvta_generic_to_shared(&sx[(lm & 1) * 8 + lr][wn * 32 + (lm >> 1) * 8]);
  __syncthreads();  // the column table is complete
  // This thread's piece of a stage: pixel pp, the 8-element group c8 of both operand tiles.  The pixel is decoded by
  // the thread itself (eight threads share one) so that the loads of stage i + 1 can be issued BEFORE the MMAs of
  // stage i: one barrier pair per stage, global latency behind the tensor work.
  const int pp = threadIdx.x >> 3, c8 = (threadIdx.x & 7) * 8;
  int ck[8], co_[8];
  if (VEC) {
    ck[0] = col_zyx[c8];
    co_[0] = col_off[c8];
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ck[e] = col_zyx[c8 + e];
      co_[e] = col_off[c8 + e];
    }
  }
  auto fetch = [&](long long p0, int half, uint4& vz, uint4& vx) {
    vz = make_uint4(0u, 0u, 0u, 0u);
    vx = make_uint4(0u, 0u, 0u, 0u);
    const long long p = p0 + pp + 32 * half;
    if (p >= p_hi) return;
    const long long zo = p * d.out_ld + d.out_coff;
    const int ow = (int)(p % d.ow);
    long long t = p / d.ow;
    const int oh = (int)(t % d.oh);
    t /= d.oh;
    const int od = (int)(t % d.od);
    const int n = (int)(t / d.od);
    const int z0 = od * d.sd - d.pd, y0 = oh * d.sh - d.ph, x0 = ow * d.sw - d.pw;
    const long long base = ((((long long)n * d.id + z0) * d.ih + y0) * d.iw + x0) * d.in_ld + d.in_coff;
    if (VEC) {
      if (co0 + c8 < d.cout) vz = ld8_bf16(dz + zo + co0 + c8);  // cout % 8 == 0: whole groups
      const int k = ck[0];
      if (k >= 0) {
        const int iz = z0 + (k & 255), iy = y0 + ((k >> 8) & 255), ix = x0 + (k >> 16);
        if ((unsigned)iz < (unsigned)d.id && (unsigned)iy < (unsigned)d.ih && (unsigned)ix < (unsigned)d.iw)
          vx = ld8_bf16(x + base + co_[0]);
      }
    } else {
      __nv_bfloat16 ez[8], ex[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        ez[e] = co0 + c8 + e < d.cout ? dz[zo + co0 + c8 + e] : __float2bfloat16_rn(0.f);
        float v = 0.f;
        const int k = ck[e];
        if (k >= 0) {
          const int iz = z0 + (k & 255), iy = y0 + ((k >> 8) & 255), ix = x0 + (k >> 16);
          if ((unsigned)iz < (unsigned)d.id && (unsigned)iy < (unsigned)d.ih && (unsigned)ix < (unsigned)d.iw)
            v = ldf(x, base + co_[e]);
        }
        ex[e] = __float2bfloat16_rn(v);
      }
      vz = *reinterpret_cast<uint4*>(ez);
      vx = *reinterpret_cast<uint4*>(ex);
    }
  };
  uint4 vz[2], vx[2];
  fetch(p_lo, 0, vz[0], vx[0]);
  fetch(p_lo, 1, vz[1], vx[1]);
  for (long long p0 = p_lo; p0 < p_hi; p0 += WM_PT) {
    __syncthreads();  // the previous stage has been consumed
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      *reinterpret_cast<uint4*>(&sdz[pp + 32 * hf][c8]) = vz[hf];
      *reinterpret_cast<uint4*>(&sx[pp + 32 * hf][c8]) = vx[hf];
    }
    __syncthreads();
    if (p0 + WM_PT < p_hi) {  // in flight while this stage multiplies
      fetch(p0 + WM_PT, 0, vz[0], vx[0]);
      fetch(p0 + WM_PT, 1, vz[1], vx[1]);
    }
#pragma unroll
    for (int ks = 0; ks < WM_PT / 16; ++ks) {
      uint32_t a[4], b01[4], b23[4];
      const uint32_t koff = (uint32_t)(ks * 16 * WM_LD * 2);
      ldsm_x4_t(a_addr + koff, a);
      ldsm_x4_t(b_addr + koff, b01);        // n tiles 0, 1 of this warp's 32 columns
      ldsm_x4_t(b_addr + koff + 32u, b23);  // n tiles 2, 3 (16 columns = 32 bytes further)
      mma_bf16_16816(acc[0], a, b01[0], b01[1]);
      mma_bf16_16816(acc[1], a, b01[2], b01[3]);
      mma_bf16_16816(acc[2], a, b23[0], b23[1]);
      mma_bf16_16816(acc[3], a, b23[2], b23[3]);
    }
  }
  // accumulator fragment: c0, c1 = (row g, columns 2t, 2t+1), c2, c3 = (row g + 8, same columns)
  const int g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int co = co0 + wm * 16 + g + (e >> 1) * 8;
      const int col = j0 + wn * 32 + nt * 8 + t2 + (e & 1);
      const float v = acc[nt][e];
      if (co >= d.cout || col >= ncols || v == 0.f) continue;
      const int tap = col / d.cin, ci = col - tap * d.cin;
      if (S2D) {
        const int ch = ci % og.x, par = ci / og.x;
        const int kt = 2 * (tap / (d.kw * d.kh)) + (par >> 2), kh_ = 2 * ((tap / d.kw) % d.kh) + ((par >> 1) & 1);
        const int kw_ = 2 * (tap % d.kw) + (par & 1);
        if (kt >= og.y || kh_ >= og.z || kw_ >= og.w) continue;
        atomicAdd(dw + ((((long long)co * og.x + ch) * og.y + kt) * og.z + kh_) * og.w + kw_, v);
      } else {
        atomicAdd(dw + ((long long)co * d.cin + ci) * taps + tap, v);
      }
    }
}

template <typename TX, typename T>
int wgrad_launch(ivf_handle* h, const ivf_conv_desc* d, const void* x, const void* dz, float* dw, cudaStream_t st) {
  const int taps = d->kd * d->kh * d->kw;
  const long long P = (long long)d->n * d->od * d->oh * d->ow;
  IVF_REQUIRE(d->kd < 256 && d->kh < 256 && d->kw < 256, "ivf_conv3d_wgrad: kernel extent above 255");
  IVF_REQUIRE((long long)d->n * d->id * d->ih * d->iw * d->in_ld < (1ll << 62), "ivf_conv3d_wgrad: tensor too large");
  IVF_REQUIRE((long long)d->kd * d->ih * d->iw * d->in_ld < (1ll << 31), "ivf_conv3d_wgrad: window span above 2^31");
  IVF_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->cout * d->cin * taps, st));
  const int co_tiles = (d->cout + 63) / 64, col_tiles = (taps * d->cin + 63) / 64;
  const long long bx = (long long)co_tiles * col_tiles;
  IVF_REQUIRE(bx < (1ll << 31), "ivf_conv3d_wgrad: too many tiles");
  long long splits = ((long long)h->sm_count * 8 + bx - 1) / bx;  // ~8 blocks per SM in total
  // at least 256 pixels per block on the CUDA-core kernel, 128 pixels (two stages) on the tensor-core kernel (the 7x7 and
  // 14x14 layers have 784 / 6 272 pixels; one stage per block made the fp32 atomics of the epilogue the bound: ncu,
  // 60 % issue slots, 27 us for a 42-tile layer)
  const long long max_splits = std::is_same<T, __nv_bfloat16>::value ? (P + 2 * WM_PT - 1) / (2 * WM_PT) : (P + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  long long ppb = (P + splits - 1) / splits;
  ppb = (ppb + WG_PT - 1) / WG_PT * WG_PT;
  splits = (P + ppb - 1) / ppb;
  const dim3 grid((unsigned)bx, (unsigned)splits);
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    static const bool use_mma = !(getenv("IVF_WGRAD_MMA") && atoi(getenv("IVF_WGRAD_MMA")) == 0);
    if (use_mma) {
      const bool vec = d->cin % 8 == 0 && d->in_ld % 8 == 0 && d->in_coff % 8 == 0 && d->cout % 8 == 0 &&
                       d->out_ld % 8 == 0 && d->out_coff % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)dz % 16) == 0;
      ppb = (ppb + WM_PT - 1) / WM_PT * WM_PT;
      const dim3 grid2((unsigned)bx, (unsigned)((P + ppb - 1) / ppb));
      if (vec) wgrad_mma_kernel<TX, true><<<grid2, 256, 0, st>>>(*d, (const TX*)x, (const __nv_bfloat16*)dz, dw, col_tiles, ppb);
      else wgrad_mma_kernel<TX, false><<<grid2, 256, 0, st>>>(*d, (const TX*)x, (const __nv_bfloat16*)dz, dw, col_tiles, ppb);
      IVF_LAUNCHED(h);
      return IVF_OK;
    }
  }
  wgrad_kernel<TX, T><<<grid, 256, 0, st>>>(*d, (const TX*)x, (const T*)dz, dw, col_tiles, ppb);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Classifier head in training mode (pt/models/I3D_doubled.py:360-371 with the dropout active;
// pt/train_i3d_smth.py:124-127 CrossEntropyLoss, mean over the batch).  The feature map must be exactly the
// average pool's window (one pooled position per clip), as in the engine.
template <typename T>
__global__ void head_pool_kernel(const T* __restrict__ feat, int ld, int coff, int pix, int C,
                                 const float* __restrict__ drop, float* __restrict__ pooled) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int p = 0; p < pix; ++p) s += ldf(feat, ((long long)b * pix + p) * ld + coff + c);
  s /= (float)pix;
  if (drop) s *= drop[(long long)b * C + c];
  pooled[(long long)b * C + c] = s;
}

// logits = pooled . W^T + bias: one warp per (clip, class), eight classes per block (one block per clip took 250 us
// for 8 x 174 x 1024: a serial loop over the classes)
__global__ void __launch_bounds__(256)
head_logits_kernel(const float* __restrict__ pooled, const float* __restrict__ w, const float* __restrict__ bias, int C,
                   int K, float* __restrict__ logits) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + warp;
  if (k >= K) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(pooled[(long long)b * C + c], w[(long long)k * C + c], s);
  s = ivf_warp_sum(s);
  if (lane == 0) logits[(long long)b * K + k] = s + bias[k];
}

// one block per clip: softmax of its logits, loss_b = -log p[target], dlogits
__global__ void __launch_bounds__(256)
head_ce_kernel(const float* __restrict__ logits, const int* __restrict__ target, int B, int K,
               float* __restrict__ dlogits, float* __restrict__ loss) {
  __shared__ float red[8];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* sl = logits + (long long)b * K;
  float mx = -INFINITY;
  for (int k = threadIdx.x; k < K; k += 256) mx = fmaxf(mx, sl[k]);
  mx = ivf_warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float se = 0.f;
  for (int k = threadIdx.x; k < K; k += 256) se += expf(sl[k] - mx);
  se = ivf_warp_sum(se);
  if (lane == 0) red[warp] = se;
  __syncthreads();
  se = 0.f;
  for (int i = 0; i < 8; ++i) se += red[i];
  const int t = target[b];
  const float lse = mx + logf(se);
  for (int k = threadIdx.x; k < K; k += 256)
    dlogits[(long long)b * K + k] = (expf(sl[k] - lse) - (k == t ? 1.f : 0.f)) / (float)B;
  if (threadIdx.x == 0) atomicAdd(loss, (lse - sl[t]) / (float)B);
}

// dW[k][c] = sum_b dlogits[b][k] pooled[b][c] ; db[k] = sum_b dlogits[b][k]
__global__ void head_wgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ pooled, int B, int C, int K,
                                  float* __restrict__ dw, float* __restrict__ db) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)K * C) {
    const int k = (int)(i / C), c = (int)(i - (long long)k * C);
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dlogits[(long long)b * K + k], pooled[(long long)b * C + c], s);
    dw[i] = s;
  }
  if (i < K) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dlogits[(long long)b * K + i];
    db[i] = s;
  }
}

// d feat[b][p][c] = (sum_k dlogits[b][k] W[k][c]) * drop[b][c] / pix, the same for every pooled pixel
template <typename T>
__global__ void head_dfeat_kernel(const float* __restrict__ dlogits, const float* __restrict__ w,
                                  const float* __restrict__ drop, int pix, int C, int K, T* __restrict__ dfeat, int ld,
                                  int coff) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int k = 0; k < K; ++k) s = fmaf(dlogits[(long long)b * K + k], w[(long long)k * C + c], s);
  if (drop) s *= drop[(long long)b * C + c];
  s /= (float)pix;
  for (int p = 0; p < pix; ++p) dfeat[((long long)b * pix + p) * ld + coff + c] = ivf_from_float<T>(s);
}

// ---------------------------------------------------------------------------------------------------------
// torch.optim.SGD (momentum, weight decay, no dampening / Nesterov) and torch.optim.Adam (L2 weight decay)
__global__ void optim_kernel(int kind, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ s1,
                             float* __restrict__ s2, long long n, float lr, float b1, float b2, float eps, float wd,
                             int step) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i] + wd * p[i];
  if (kind == 0) {
    if (b1 != 0.f) {
      const float buf = step == 1 ? gi : b1 * s1[i] + gi;  // first step: the buffer IS the gradient
      s1[i] = buf;
      gi = buf;
    }
    p[i] -= lr * gi;
  } else {
    const float m = b1 * s1[i] + (1.f - b1) * gi;
    const float v = b2 * s2[i] + (1.f - b2) * gi * gi;
    s1[i] = m;
    s2[i] = v;
    const float c1 = 1.f - powf(b1, (float)step), c2 = 1.f - powf(b2, (float)step);
    p[i] -= lr / c1 * m / (sqrtf(v) / sqrtf(c2) + eps);
  }
}

// the same update for MANY tensors in one launch: table rows {p, g, s1, s2, n} (five 64-bit words), one block per
// row, rows of at most a few thousand elements (the host cuts long tensors into chunks); grad_scale multiplies the
// gradient first (1 / world size after the all-reduce of a data-parallel step)
__global__ void __launch_bounds__(256)
optim_multi_kernel(const long long* __restrict__ table, int kind, float lr, float b1, float b2, float eps, float wd,
                   int step, float grad_scale) {
  const long long* row = table + 5ll * blockIdx.x;
  float* __restrict__ p = reinterpret_cast<float*>(row[0]);
  const float* __restrict__ g = reinterpret_cast<const float*>(row[1]);
  float* __restrict__ s1 = reinterpret_cast<float*>(row[2]);
  float* __restrict__ s2 = reinterpret_cast<float*>(row[3]);
  const int n = (int)row[4];
  const float c1 = 1.f - powf(b1, (float)step), c2 = 1.f - powf(b2, (float)step);
  for (int i = threadIdx.x; i < n; i += 256) {
    float gi = g[i] * grad_scale + wd * p[i];
    if (kind == 0) {
      if (b1 != 0.f) {
        const float buf = step == 1 ? gi : b1 * s1[i] + gi;
        s1[i] = buf;
        gi = buf;
      }
      p[i] -= lr * gi;
    } else {
      const float m = b1 * s1[i] + (1.f - b1) * gi;
      const float v = b2 * s2[i] + (1.f - b2) * gi * gi;
      s1[i] = m;
      s2[i] = v;
      p[i] -= lr / c1 * m / (sqrtf(v) / sqrtf(c2) + eps);
    }
  }
}

// dropout mask, already scaled: 0 with probability p, else 1/(1-p); one counter-based hash per element
// (the reference draws from torch's generator, pt/models/I3D_doubled.py:319 - no stream can match it bit for bit)
__global__ void dropout_mask_kernel(float* __restrict__ out, long long n, float p, unsigned long long seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);  // splitmix64
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  const float u = (float)(x >> 40) * (1.0f / 16777216.0f);
  out[i] = u < p ? 0.f : 1.f / (1.f - p);
}

}  // namespace

extern "C" int ivf_dropout_mask(ivf_handle* h, float* out, long long n, float p, unsigned long long seed,
                                void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && out && n > 0 && p >= 0.f && p < 1.f, "ivf_dropout_mask: null argument or p outside [0, 1)");
  dropout_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n, p, seed);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int ivf_bn_train_fwd(ivf_handle* h, int dtype, const void* z, int z_ld, int z_coff, long long m, int c,
                                const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                                float* running_var, float* save_mean, float* save_rstd, double* ws, void* y, int y_ld,
                                int y_coff, int relu, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && z && gamma && beta && save_mean && save_rstd && ws && y && m > 0 && c > 0,
              "ivf_bn_train_fwd: null argument or empty tensor");
  IVF_REQUIRE(dtype == IVF_F32 || dtype == IVF_BF16, "ivf_bn_train_fwd: unknown dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  IVF_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * (size_t)c, st));
  const dim3 grid = bn_grid(h, m, c), block(32, BN_ROWS);
#define IVF_BN_FWD(T)                                                                                              \
  do {                                                                                                             \
    bn_stats_kernel<T><<<grid, block, 0, st>>>((const T*)z, z_ld, z_coff, m, c, ws);                               \
    IVF_LAUNCHED(h);                                                                                               \
    bn_apply_kernel<T><<<grid, block, 0, st>>>((const T*)z, z_ld, z_coff, m, c, gamma, beta, eps, momentum,       \
                                               running_mean, running_var, save_mean, save_rstd, ws, (T*)y, y_ld,  \
                                               y_coff, relu);                                                      \
    IVF_LAUNCHED(h);                                                                                               \
  } while (0)
  if (dtype == IVF_F32) IVF_BN_FWD(float);
  else IVF_BN_FWD(__nv_bfloat16);
#undef IVF_BN_FWD
  return IVF_OK;
}

extern "C" int ivf_bn_train_bwd(ivf_handle* h, int dtype, int dy_dtype, const void* dy, int dy_ld, int dy_coff,
                                const void* y, int y_ld, int y_coff, const void* z, int z_ld, int z_coff, long long m,
                                int c, const float* gamma, const float* save_mean, const float* save_rstd, double* ws,
                                void* dz, int dz_ld, int dz_coff, float* dgamma, float* dbeta, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && dy && z && gamma && save_mean && save_rstd && ws && dz && dgamma && dbeta && m > 0 && c > 0,
              "ivf_bn_train_bwd: null argument or empty tensor");
  IVF_REQUIRE((dtype == IVF_F32 || dtype == IVF_BF16) && (dy_dtype == IVF_F32 || dy_dtype == dtype),
              "ivf_bn_train_bwd: dtype %d / dy_dtype %d (dy is fp32 or of the activation type)", dtype, dy_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  IVF_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * (size_t)c, st));
  const dim3 grid = bn_grid(h, m, c), block(32, BN_ROWS);
#define IVF_BN_BWD(T, TG)                                                                                          \
  do {                                                                                                             \
    bn_bwd_stats_kernel<T, TG><<<grid, block, 0, st>>>((const TG*)dy, dy_ld, dy_coff, (const T*)y, y_ld, y_coff,  \
                                                       (const T*)z, z_ld, z_coff, m, c, save_mean, save_rstd, ws); \
    IVF_LAUNCHED(h);                                                                                               \
    bn_bwd_apply_kernel<T, TG><<<grid, block, 0, st>>>((const TG*)dy, dy_ld, dy_coff, (const T*)y, y_ld, y_coff,  \
                                                       (const T*)z, z_ld, z_coff, m, c, gamma, save_mean,          \
                                                       save_rstd, ws, (T*)dz, dz_ld, dz_coff, dgamma, dbeta);      \
    IVF_LAUNCHED(h);                                                                                               \
  } while (0)
  if (dtype == IVF_F32) IVF_BN_BWD(float, float);
  else if (dy_dtype == IVF_F32) IVF_BN_BWD(__nv_bfloat16, float);
  else IVF_BN_BWD(__nv_bfloat16, __nv_bfloat16);
#undef IVF_BN_BWD
  return IVF_OK;
}

extern "C" int ivf_conv3d_wgrad(ivf_handle* h, const ivf_conv_desc* d, int x_dtype, const void* x, const void* dz,
                                float* dw, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && d && x && dz && dw, "ivf_conv3d_wgrad: null argument");
  IVF_REQUIRE(d->n > 0 && d->id > 0 && d->ih > 0 && d->iw > 0 && d->od > 0 && d->oh > 0 && d->ow > 0 && d->cin > 0 &&
                  d->cout > 0 && d->kd > 0 && d->kh > 0 && d->kw > 0 && d->sd > 0 && d->sh > 0 && d->sw > 0 &&
                  !d->transposed && d->in_ld >= d->in_coff + d->cin && d->out_ld >= d->out_coff + d->cout,
              "ivf_conv3d_wgrad: bad descriptor");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == IVF_F32 && x_dtype == IVF_F32) return wgrad_launch<float, float>(h, d, x, dz, dw, st);
  if (d->dtype == IVF_BF16 && x_dtype == IVF_BF16) return wgrad_launch<__nv_bfloat16, __nv_bfloat16>(h, d, x, dz, dw, st);
  if (d->dtype == IVF_BF16 && x_dtype == IVF_F32) return wgrad_launch<float, __nv_bfloat16>(h, d, x, dz, dw, st);
  IVF_FAIL(IVF_EINVAL, "ivf_conv3d_wgrad: dtype %d (dz) / %d (x): fp32/fp32, bf16/bf16 or bf16 dz with fp32 x", d->dtype,
           x_dtype);
}

extern "C" int ivf_head_train_fwd(ivf_handle* h, int dtype, const void* feat, int ld, int coff, int batch, int pix,
                                  int c, const float* drop, const float* w, const float* bias, const int* target,
                                  int classes, float* pooled, float* logits, float* dlogits, float* loss,
                                  void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && feat && w && bias && target && pooled && logits && dlogits && loss && batch > 0 && pix > 0 &&
                  c > 0 && classes > 0,
              "ivf_head_train_fwd: null argument or bad size");
  IVF_REQUIRE(dtype == IVF_F32 || dtype == IVF_BF16, "ivf_head_train_fwd: unknown dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  IVF_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  const dim3 grid((c + 127) / 128, batch);
  if (dtype == IVF_F32)
    head_pool_kernel<float><<<grid, 128, 0, st>>>((const float*)feat, ld, coff, pix, c, drop, pooled);
  else
    head_pool_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)feat, ld, coff, pix, c, drop, pooled);
  IVF_LAUNCHED(h);
  head_logits_kernel<<<dim3((classes + 7) / 8, batch), 256, 0, st>>>(pooled, w, bias, c, classes, logits);
  IVF_LAUNCHED(h);
  head_ce_kernel<<<batch, 256, 0, st>>>(logits, target, batch, classes, dlogits, loss);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_head_train_bwd(ivf_handle* h, int dtype, const float* dlogits, const float* pooled,
                                  const float* drop, const float* w, int batch, int pix, int c, int classes, float* dw,
                                  float* db, void* dfeat, int ld, int coff, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && dlogits && pooled && w && dw && db && dfeat && batch > 0 && pix > 0 && c > 0 && classes > 0,
              "ivf_head_train_bwd: null argument or bad size");
  IVF_REQUIRE(dtype == IVF_F32 || dtype == IVF_BF16, "ivf_head_train_bwd: unknown dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)classes * c;
  head_wgrad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dlogits, pooled, batch, c, classes, dw, db);
  IVF_LAUNCHED(h);
  const dim3 grid((c + 127) / 128, batch);
  if (dtype == IVF_F32)
    head_dfeat_kernel<float><<<grid, 128, 0, st>>>(dlogits, w, drop, pix, c, classes, (float*)dfeat, ld, coff);
  else
    head_dfeat_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(dlogits, w, drop, pix, c, classes, (__nv_bfloat16*)dfeat, ld,
                                                           coff);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_optim_step(ivf_handle* h, int kind, float* p, const float* g, float* s1, float* s2, long long n,
                              float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                              void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && p && g && n > 0 && step >= 1, "ivf_optim_step: null argument, empty tensor or step < 1");
  IVF_REQUIRE(kind == 0 || kind == 1, "ivf_optim_step: kind must be 0 (SGD) or 1 (Adam)");
  IVF_REQUIRE(kind == 0 ? (beta1 == 0.f || s1) : (s1 && s2), "ivf_optim_step: optimizer state buffers missing");
  optim_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, p, g, s1, s2, n, lr, beta1, beta2,
                                                                             eps, weight_decay, step);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_optim_step_multi(ivf_handle* h, int kind, const void* table, int rows, float lr, float beta1,
                                    float beta2, float eps, float weight_decay, int step, float grad_scale,
                                    void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && table && rows > 0 && step >= 1, "ivf_optim_step_multi: null argument, empty table or step < 1");
  IVF_REQUIRE(kind == 0 || kind == 1, "ivf_optim_step_multi: kind must be 0 (SGD) or 1 (Adam)");
  optim_multi_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>((const long long*)table, kind, lr, beta1, beta2, eps,
                                                             weight_decay, step, grad_scale);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

extern "C" int ivf_conv3d_wgrad_s2d(ivf_handle* h, const ivf_conv_desc* d, const void* x, const void* dz, float* dw,
                                    int ci, int kd, int kh, int kw, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && d && x && dz && dw, "ivf_conv3d_wgrad_s2d: null argument");
  IVF_REQUIRE(d->dtype == IVF_BF16 && !d->transposed && d->sd == 1 && d->sh == 1 && d->sw == 1 && ci > 0 &&
                  d->cin == 8 * ci && d->kd == (kd + 1) / 2 && d->kh == (kh + 1) / 2 && d->kw == (kw + 1) / 2 &&
                  d->in_ld >= d->in_coff + d->cin && d->out_ld >= d->out_coff + d->cout,
              "ivf_conv3d_wgrad_s2d: d must describe the stride-1 bf16 convolution over the 2x2x2 space-to-depth record "
              "(cin = 8 * ci, kernel = ceil(k / 2))");
  IVF_REQUIRE(d->cin % 8 == 0 && d->in_ld % 8 == 0 && d->in_coff % 8 == 0 && d->cout % 8 == 0 && d->out_ld % 8 == 0 &&
                  d->out_coff % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)dz % 16) == 0,
              "ivf_conv3d_wgrad_s2d: channel counts / offsets must be multiples of 8 and the buffers 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long P = (long long)d->n * d->od * d->oh * d->ow;
  IVF_REQUIRE((long long)d->kd * d->ih * d->iw * d->in_ld < (1ll << 31), "ivf_conv3d_wgrad_s2d: window span above 2^31");
  IVF_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->cout * ci * kd * kh * kw, st));
  const int taps = d->kd * d->kh * d->kw;
  const int co_tiles = (d->cout + 63) / 64, col_tiles = (taps * d->cin + 63) / 64;
  const long long bx = (long long)co_tiles * col_tiles;
  long long splits = ((long long)h->sm_count * 8 + bx - 1) / bx;
  const long long max_splits = (P + WM_PT - 1) / WM_PT;
  if (splits > max_splits) splits = max_splits;
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  long long ppb = (P + splits - 1) / splits;
  ppb = (ppb + WM_PT - 1) / WM_PT * WM_PT;
  const dim3 grid((unsigned)bx, (unsigned)((P + ppb - 1) / ppb));
  wgrad_mma_kernel<__nv_bfloat16, true, true><<<grid, 256, 0, st>>>(
      *d, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dz, dw, col_tiles, ppb, make_int4(ci, kd, kh, kw));
  IVF_LAUNCHED(h);
  return IVF_OK;
}
