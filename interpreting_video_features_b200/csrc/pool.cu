// Channels-last max-pool with TF-'same' ZERO padding, forward and backward.
// Reference: pt/models/I3D_doubled.py:8-40 (MaxPool3dSamePadding = F.pad with zeros, then
// nn.MaxPool3d without padding) — the padded zeros take part in the max, and ATen's
// max_pool3d_with_indices keeps the FIRST maximum in (d,h,w) window-scan order; backward sends
// the gradient to that element only (a padded winner drops it).  Also used as nn.MaxPool2d
// (pt/models/convolution_lstm.py:79) with kd=1, pad 0.
// Bandwidth-bound: one thread per (pixel, 16-byte channel vector), coalesced along C.
#include "common.cuh"

#include <climits>

namespace {

#define COMMA ,

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
  enum : int { KD_ = KD, KH_ = KH, KW_ = KW, SD_ = SD, SH_ = SH, SW_ = SW };
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

// 8 bf16 -> one byte, bit i = (element i > 0).  A positive bf16 is a positive 16-bit integer (a positive NaN
// counts as positive here, unlike the float compare; a ReLU output is never NaN unless its input was).
__device__ __forceinline__ uint8_t relu_byte(const uint4& raw) {
  const uint32_t m0 = __vcmpgts2(raw.x, 0u), m1 = __vcmpgts2(raw.y, 0u), m2 = __vcmpgts2(raw.z, 0u),
                 m3 = __vcmpgts2(raw.w, 0u);
  return (uint8_t)((m0 & 1u) | ((m0 >> 15) & 2u) | ((m1 & 1u) << 2) | ((m1 >> 13) & 8u) | ((m2 & 1u) << 4) |
                   ((m2 >> 11) & 32u) | ((m3 & 1u) << 6) | ((m3 >> 9) & 128u));
}

// --- integer keys.  "First maximum in scan order wins" is one integer max per element when the bf16 value and
// the tap share a 32-bit key: [bf16 bits made monotone as a signed integer : 16][255 - tap : 16].  Positive floats
// already order like their bit patterns; negative ones get their 15 magnitude bits flipped.  A later tap carries a
// smaller low half, so it only takes over when its value is strictly greater.  A (positive) NaN has the largest
// pattern and wins like in ATen (with several NaNs in one window the first is kept, ATen keeps the last); -0
// orders below +0 (ATen: equal) - neither occurs in a ReLU network.  ~4 instructions per element and tap
// (PRMT, SHF, LOP3, IMNMX) against ~6 of the compare-to-mask form, which also needed two 16-bit lanes per word.
// NONNEG (IVF_POOL_NONNEG: the caller vouches that no input is negative - ReLU outputs) skips the flip: 2
// instructions per element and tap.
// 16-byte read-only load the compiler may not move (asm volatile keeps program order among themselves)
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
template <bool NONNEG>
__device__ __forceinline__ int pool_key_flip(int k) {
  return NONNEG ? k : (k ^ ((k >> 31) & 0x7fff0000));
}
// keys of the two bf16 of a word; tapc = 255 - tap
template <bool NONNEG>
__device__ __forceinline__ void pool_keys_max(const uint4& raw, uint32_t tapc, int (&best)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int lo = pool_key_flip<NONNEG>((int)__byte_perm(w[i], tapc, 0x1054));
    const int hi = pool_key_flip<NONNEG>((int)__byte_perm(w[i], tapc, 0x3254));
    best[2 * i] = max(best[2 * i], lo);
    best[2 * i + 1] = max(best[2 * i + 1], hi);
  }
}
// eight keys -> eight bf16 values and eight arg-max bytes
template <bool NONNEG>
__device__ __forceinline__ void pool_keys_unpack(const int (&k)[8], uint4& val, uint2& idx) {
  uint32_t v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    v[i] = __byte_perm((uint32_t)pool_key_flip<NONNEG>(k[2 * i]), (uint32_t)pool_key_flip<NONNEG>(k[2 * i + 1]), 0x7632);
  val = make_uint4(v[0], v[1], v[2], v[3]);
  const uint32_t a = __byte_perm((uint32_t)k[0], (uint32_t)k[1], 0x0040), b = __byte_perm((uint32_t)k[2], (uint32_t)k[3], 0x0040);
  const uint32_t c = __byte_perm((uint32_t)k[4], (uint32_t)k[5], 0x0040), e = __byte_perm((uint32_t)k[6], (uint32_t)k[7], 0x0040);
  idx.x = ~__byte_perm(a, b, 0x5410);
  idx.y = ~__byte_perm(c, e, 0x5410);
}

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
// MODE 0: any input; 1: any input + ReLU' bit mask of the input (relu_bits); 2: non-negative input.
template <typename G, int MODE>
__global__ void __launch_bounds__(256)
maxpool_fwd_bf16x8_kernel(ivf_pool_desc d, const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                          uint8_t* __restrict__ argmax, uint8_t* __restrict__ relu_bits, int rows) {
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
  int best[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) best[i] = INT_MIN;  // below every key: the first tap is always taken
#pragma unroll
  for (int a = 0; a < KD; ++a) {
    const bool dok = (unsigned)(zd0 + a) < (unsigned)d.id;
#pragma unroll
    for (int b = 0; b < KH; ++b) {
      const bool hok = dok && (unsigned)(zh0 + b) < (unsigned)d.ih;
      const int delta_row = ((a * d.ih + b) * d.iw) * d.in_ld;
#pragma unroll
      for (int e = 0; e < KW; ++e) {
        const uint32_t tapc = 255u - (uint32_t)((a * KH + b) * KW + e);
        uint4 raw = make_uint4(0u, 0u, 0u, 0u);  // explicit zero padding
        if (hok && (unsigned)(zw0 + e) < (unsigned)d.iw) {
          raw = *reinterpret_cast<const uint4*>(p0 + delta_row + e * d.in_ld);
          // ReLU' bit mask of the INPUT for the backward pass (one bit per element instead of re-reading the
          // producer's bf16 output there): every input element belongs to exactly one window origin - the taps
          // below the stride - so that output's thread, which holds the element anyway, writes its byte
          if (MODE == 1 && a < SD && b < SH && e < SW) {
            const int ipix = pix0 + (a * d.ih + b) * d.iw + e;
            relu_bits[(long long)ipix * cv + (c >> 3)] = relu_byte(raw);
          }
        }
        pool_keys_max<MODE == 2>(raw, tapc, best);
      }
    }
  }
  const int opix = row * d.ow + ow;
  uint4 o;
  uint2 pk;
  pool_keys_unpack<MODE == 2>(best, o, pk);
  *reinterpret_cast<uint4*>(out + (long long)opix * d.out_ld + d.out_coff + c) = o;
  if (argmax) *reinterpret_cast<uint2*>(argmax + (long long)opix * d.c + c) = pk;
}

// Forward of the stride-2 stage pools: like the kernel above, but a thread keeps its (column, 8 channels) for a
// SEGMENT of consecutive output rows.  The one-row-per-block form executed ~450 instructions per output vector of
// which ~250 were index arithmetic and bounds set-up (ncu: issue bound, 66 % issue slots), and its blocks re-read each
// other's halo rows from the other die's L2 (dram__bytes_read = 2x the tensor); here the set-up is paid once per
// segment and consecutive rows of a slice stay on one SM.
template <typename G, bool NONNEG>
__global__ void __launch_bounds__(256)
maxpool_fwd_rows_kernel(ivf_pool_desc d, const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                        uint8_t* __restrict__ argmax, int items, int rseg, int nseg) {
  const int item = blockIdx.z * gridDim.y + blockIdx.y;  // (clip, output depth, row segment)
  if (item >= items) return;
  const int cv = d.c >> 3;
  const int el = blockIdx.x * 256 + threadIdx.x;
  if (el >= d.ow * cv) return;
  constexpr int KD = G::KD_, KH = G::KH_, KW = G::KW_, SD = G::SD_, SH = G::SH_, SW = G::SW_;
  const int ow = el / cv, c = (el - ow * cv) << 3;
  const int seg = item % nseg;
  const int t = item / nseg;
  const int od = t % d.od, n = t / d.od;
  const int oh0 = seg * rseg, oh1 = min(d.oh, oh0 + rseg);
  const int zd0 = od * SD - d.pd, zw0 = ow * SW - d.pw;
  bool wok[KW], dok[KD];
#pragma unroll
  for (int e = 0; e < KW; ++e) wok[e] = (unsigned)(zw0 + e) < (unsigned)d.iw;
#pragma unroll
  for (int a = 0; a < KD; ++a) dok[a] = (unsigned)(zd0 + a) < (unsigned)d.id;
  const int row_elems = d.iw * d.in_ld;
  const int plane_elems = d.ih * row_elems;
  // column of this thread in plane zd0, row 0 (only dereferenced where the tap is inside the tensor)
  const __nv_bfloat16* pcol = in + ((long long)(n * d.id + zd0) * d.ih * d.iw + zw0) * d.in_ld + d.in_coff + c;
  const int out_row0 = ((n * d.od + od) * d.oh) * d.ow + ow;
#pragma unroll 1
  for (int oh = oh0; oh < oh1; ++oh) {
    const int zh0 = oh * SH - d.ph;
    int best[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) best[i] = INT_MIN;
#pragma unroll
    for (int a = 0; a < KD; ++a) {
      // all loads of a plane are issued before the first compare waits on one of them (the compiler otherwise
      // interleaves load and use: seven exposed memory latencies per output)
      uint4 raw[KH * KW];
#pragma unroll
      for (int b = 0; b < KH; ++b) {
        const bool rok = dok[a] && (unsigned)(zh0 + b) < (unsigned)d.ih;
        const __nv_bfloat16* prow = pcol + (long long)a * plane_elems + (long long)(zh0 + b) * row_elems;
#pragma unroll
        for (int e = 0; e < KW; ++e) {
          raw[b * KW + e] = make_uint4(0u, 0u, 0u, 0u);  // explicit zero padding
          if (rok && wok[e]) raw[b * KW + e] = ld_nc_v4(prow + e * d.in_ld);
        }
      }
#pragma unroll
      for (int k = 0; k < KH * KW; ++k) pool_keys_max<NONNEG>(raw[k], 255u - (uint32_t)(a * KH * KW + k), best);
    }
    const int opix = out_row0 + oh * d.ow;
    uint4 o;
    uint2 pk;
    pool_keys_unpack<NONNEG>(best, o, pk);
    *reinterpret_cast<uint4*>(out + (long long)opix * d.out_ld + d.out_coff + c) = o;
    if (argmax) *reinterpret_cast<uint2*>(argmax + (long long)opix * d.c + c) = pk;
  }
}

// Forward of the 3x3x3 stride-1 branch pools: a thread owns one (row, column, 8 channels) and walks the depth.
// The 27-tap window of consecutive depths shares two of its three planes, so the thread reduces each input
// plane ONCE to its 3x3 maximum (value + first-maximum tap, 9 taps) and combines three plane results per
// output (2 more compares): 11 compare blocks per output instead of 27, and 9 loads instead of 27.  Scan
// order is (depth, row, column), so "first maximum wins" composes: within a plane the first tap wins, across
// planes a later plane must be strictly greater (or NaN) to take over - the same rule the flat scan applies.
struct PlaneMax {
  int k[8];  // integer keys (pool_keys), low half = 255 - tap within the plane
};

template <bool NONNEG>
__global__ void __launch_bounds__(256)
maxpool_fwd_s1col_kernel(ivf_pool_desc d, const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                         uint8_t* __restrict__ argmax, int rows, int dseg) {
  const int row = blockIdx.z * gridDim.y + blockIdx.y;  // (clip, depth segment, output row)
  if (row >= rows) return;
  const int cv = d.c >> 3;
  const int el = blockIdx.x * 256 + threadIdx.x;
  if (el >= d.ow * cv) return;
  const int ow = el / cv, c = (el - ow * cv) << 3;
  const int oh = row % d.oh;
  const int t = row / d.oh;
  const int nseg = (d.od + dseg - 1) / dseg;
  const int seg = t % nseg, n = t / nseg;
  const int od0 = seg * dseg, od1 = min(d.od, od0 + dseg);
  const int zh0 = oh - d.ph, zw0 = ow - d.pw;
  // tap validity in the plane (block-uniform rows, per-thread columns)
  bool hok[3], wok[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    hok[k] = (unsigned)(zh0 + k) < (unsigned)d.ih;
    wok[k] = (unsigned)(zw0 + k) < (unsigned)d.iw;
  }
  const int plane_stride = d.ih * d.iw * d.in_ld;
  const __nv_bfloat16* p00 = in + (long long)(((n * d.id) * d.ih + zh0) * d.iw + zw0) * d.in_ld + d.in_coff + c;
  auto plane_max = [&](int zd, PlaneMax& pm) {
    if ((unsigned)zd >= (unsigned)d.id) {  // a plane of padding: all zeros, first tap wins
#pragma unroll
      for (int i = 0; i < 8; ++i) pm.k[i] = 255;
      return;
    }
    const __nv_bfloat16* pz = p00 + (long long)zd * plane_stride;
    uint4 raw[9];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        raw[b * 3 + e] = make_uint4(0u, 0u, 0u, 0u);  // explicit zero padding
        if (hok[b] && wok[e]) raw[b * 3 + e] = *reinterpret_cast<const uint4*>(pz + (b * d.iw + e) * d.in_ld);
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) pm.k[i] = INT_MIN;
#pragma unroll
    for (int k = 0; k < 9; ++k) pool_keys_max<NONNEG>(raw[k], 255u - (uint32_t)k, pm.k);
  };
  PlaneMax pa, pb, pc;
  plane_max(od0 - d.pd, pa);
  plane_max(od0 - d.pd + 1, pb);
  for (int od = od0; od < od1; ++od) {
    plane_max(od - d.pd + 2, pc);
    int best[8];
    // taps of the second / third plane are 9 / 18 later in scan order: their low halves drop by as much (no
    // borrow: 255 - 8 - 18 > 0), so an earlier plane keeps a tie
#pragma unroll
    for (int i = 0; i < 8; ++i) best[i] = max(max(pa.k[i], pb.k[i] - 9), pc.k[i] - 18);
    const int opix = ((n * d.od + od) * d.oh + oh) * d.ow + ow;
    uint4 o;
    uint2 pk;
    pool_keys_unpack<NONNEG>(best, o, pk);
    *reinterpret_cast<uint4*>(out + (long long)opix * d.out_ld + d.out_coff + c) = o;
    if (argmax) *reinterpret_cast<uint2*>(argmax + (long long)opix * d.c + c) = pk;
    pa = pb;
    pb = pc;
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

// Backward of the stride-2 (in H and W) pools: a thread owns the 2x2 input patch that shares one window
// origin, for 8 channels.  The patch sits in at most 2x2 windows per depth candidate (one for 2x2 windows);
// their argmax and gradient vectors are loaded ONCE and serve all four cells - nine (cell, window) tests for
// a 3x3 window instead of the sixteen loads and tests four independent cells make (the kernel above, which
// measured issue bound at ~435 instructions per cell).  A test is one SIMD byte compare per four channels and
// per channel a byte-replicate (the 0xff/0x00 compare result becomes a 32-bit mask), an AND and an add.
template <typename G>
__global__ void __launch_bounds__(256)
maxpool_bwd_s2patch_kernel(ivf_pool_desc d, const __nv_bfloat16* __restrict__ dy,
                           const uint8_t* __restrict__ argmax, const float* __restrict__ acc_in,
                           const __nv_bfloat16* __restrict__ mask_y, const uint8_t* __restrict__ relu_bits,
                           const float* __restrict__ mask_scale, void* __restrict__ dx, int rows, int qhn, int qwn) {
  const int row = blockIdx.z * gridDim.y + blockIdx.y;
  if (row >= rows) return;
  const int cv = d.c >> 3;
  const int el = blockIdx.x * 256 + threadIdx.x;
  if (el >= qwn * cv) return;
  const int KD = G::kd(d), KH = G::kh(d), KW = G::kw(d), SD = G::sd(d);
  const int qw = el / cv, c = (el - qw * cv) << 3;
  const int qh = row % qhn;
  const int t = row / qhn;
  const int idd = t % d.id, n = t / d.id;
  float g[2][2][8];
#pragma unroll
  for (int i = 0; i < 32; ++i) (&g[0][0][0])[i] = 0.f;
  const int pd0 = (idd + d.pd) % SD, qd = (idd + d.pd) / SD;
#pragma unroll
  for (int ja = 0; ja * SD < KD; ++ja) {
    const int a = pd0 + ja * SD, od = qd - ja;
    if (a >= KD || od < 0 || od >= d.od) continue;
#pragma unroll
    for (int jh = 0; jh * 2 < KH; ++jh) {
      const int oh = qh - jh;
      if (oh < 0 || oh >= d.oh) continue;
#pragma unroll
      for (int jw = 0; jw * 2 < KW; ++jw) {
        const int ow = qw - jw;
        if (ow < 0 || ow >= d.ow) continue;
        const int opix = ((n * d.od + od) * d.oh + oh) * d.ow + ow;
        const uint2 pk = *reinterpret_cast<const uint2*>(argmax + (long long)opix * d.c + c);
        const uint4 raw = *reinterpret_cast<const uint4*>(dy + (long long)opix * d.out_ld + d.out_coff + c);
        // fp32 images of the eight bf16 gradients
        const uint32_t f[8] = {raw.x << 16, raw.x & 0xffff0000u, raw.y << 16, raw.y & 0xffff0000u,
                               raw.z << 16, raw.z & 0xffff0000u, raw.w << 16, raw.w & 0xffff0000u};
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          const int b = rh + 2 * jh;
          if (b >= KH) continue;
#pragma unroll
          for (int rw = 0; rw < 2; ++rw) {
            const int e = rw + 2 * jw;
            if (e >= KW) continue;
            const uint32_t tap4 = (uint32_t)((a * KH + b) * KW + e) * 0x01010101u;
            const uint32_t e0 = __vcmpeq4(pk.x, tap4), e1 = __vcmpeq4(pk.y, tap4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t m = __byte_perm(i < 4 ? e0 : e1, 0u, 0x1111u * (i & 3));
              g[rh][rw][i] += __uint_as_float(f[i] & m);
            }
          }
        }
      }
    }
  }
  const float4 s0 = (d.flags & IVF_EP_MASK) ? *reinterpret_cast<const float4*>(mask_scale + c) : make_float4(1, 1, 1, 1);
  const float4 s1 = (d.flags & IVF_EP_MASK) ? *reinterpret_cast<const float4*>(mask_scale + c + 4) : make_float4(1, 1, 1, 1);
  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
  for (int rh = 0; rh < 2; ++rh) {
    const int ih = 2 * qh + rh - d.ph;
    if (ih < 0 || ih >= d.ih) continue;
#pragma unroll
    for (int rw = 0; rw < 2; ++rw) {
      const int iw = 2 * qw + rw - d.pw;
      if (iw < 0 || iw >= d.iw) continue;
      float* gg = g[rh][rw];
      const int ipix = ((n * d.id + idd) * d.ih + ih) * d.iw + iw;
      const long long o = (long long)ipix * d.in_ld + d.in_coff + c;
      if (d.flags & IVF_EP_ACCUM) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float4 a4 = *reinterpret_cast<const float4*>(acc_in + o + 4 * i);
          gg[4 * i] += a4.x;
          gg[4 * i + 1] += a4.y;
          gg[4 * i + 2] += a4.z;
          gg[4 * i + 3] += a4.w;
        }
      }
      if ((d.flags & IVF_EP_MASK) && relu_bits) {  // the forward pass left one bit per element
        const uint32_t by = relu_bits[(long long)ipix * cv + (c >> 3)];
#pragma unroll
        for (int i = 0; i < 8; ++i) gg[i] = ((by >> i) & 1u) ? gg[i] * sc[i] : 0.f;
      } else if (d.flags & IVF_EP_MASK) {
        const uint4 rawy = *reinterpret_cast<const uint4*>(mask_y + (long long)ipix * d.mask_ld + d.mask_coff + c);
        const uint32_t yw[4] = {rawy.x, rawy.y, rawy.z, rawy.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float y = __uint_as_float((i & 1) ? (yw[i >> 1] & 0xffff0000u) : (yw[i >> 1] << 16));
          gg[i] = y > 0.f ? gg[i] * sc[i] : 0.f;
        }
      }
      if (d.flags & IVF_EP_OUT_F32) {
        float* q = reinterpret_cast<float*>(dx) + o;
        reinterpret_cast<float4*>(q)[0] = make_float4(gg[0], gg[1], gg[2], gg[3]);
        reinterpret_cast<float4*>(q)[1] = make_float4(gg[4], gg[5], gg[6], gg[7]);
      } else {
        uint4 pkd;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(gg[0], gg[1]), h1 = __floats2bfloat162_rn(gg[2], gg[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(gg[4], gg[5]), h3 = __floats2bfloat162_rn(gg[6], gg[7]);
        pkd.x = *reinterpret_cast<uint32_t*>(&h0);
        pkd.y = *reinterpret_cast<uint32_t*>(&h1);
        pkd.z = *reinterpret_cast<uint32_t*>(&h2);
        pkd.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dx) + o) = pkd;
      }
    }
  }
}

// Pure-routing backward of the stride-2 stage pools (bf16 in and out, no epilogue: the engine applies the
// producer's ReLU'/BN' one stage earlier, see engine._emit_stage_bwd).  Same patch ownership as the kernel above -
// a thread owns the 2x2 input cells that share a window origin, for 8 channels - but it keeps its column for a
// SEGMENT of patch rows: the windows of output row qh serve patch row qh (their rows 0/1) and patch row qh + 1
// (their row 2), so each window's arg-max / gradient vectors are loaded once and carried in registers to the next
// iteration, and the index arithmetic is paid once per segment (the one-row form executed ~660 instructions per
// patch for 9 x 26 of routing work; ncu: issue bound).
template <typename G>
__global__ void __launch_bounds__(256)
maxpool_bwd_s2route_kernel(ivf_pool_desc d, const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ argmax,
                           __nv_bfloat16* __restrict__ dx, int items, int qseg, int nseg, int qhn, int qwn) {
  const int item = blockIdx.z * gridDim.y + blockIdx.y;  // (clip, input depth, patch-row segment)
  if (item >= items) return;
  const int cv = d.c >> 3;
  const int el = blockIdx.x * 256 + threadIdx.x;
  if (el >= qwn * cv) return;
  constexpr int KD = G::KD_, KH = G::KH_, KW = G::KW_, SD = G::SD_;
  constexpr int JA = (KD + SD - 1) / SD, JH = (KH + 1) / 2, JW = (KW + 1) / 2;
  const int qw = el / cv, c = (el - qw * cv) << 3;
  const int seg = item % nseg;
  const int t = item / nseg;
  const int idd = t % d.id, n = t / d.id;
  const int qh0 = seg * qseg, qh1 = min(qhn, qh0 + qseg);
  const int pd0 = (idd + d.pd) % SD, qd = (idd + d.pd) / SD;
  // per window column / depth candidate: validity, tap base, offset of (od, row 0, ow) in output pixels
  bool wv[JA][JW];
  int obase[JA][JW];
  uint32_t tapa[JA];
#pragma unroll
  for (int ja = 0; ja < JA; ++ja) {
    const int a = pd0 + ja * SD, od = qd - ja;
    const bool dv = a < KD && od >= 0 && od < d.od;
    tapa[ja] = (uint32_t)(a * KH * KW) * 0x01010101u;
#pragma unroll
    for (int jw = 0; jw < JW; ++jw) {
      const int ow = qw - jw;
      wv[ja][jw] = dv && ow >= 0 && ow < d.ow;
      obase[ja][jw] = ((n * d.od + od) * d.oh) * d.ow + ow;
    }
  }
  auto load_win = [&](int ja, int jw, int oh, uint2& pk, uint4& raw) {
    pk = make_uint2(0xffffffffu, 0xffffffffu);  // tap 255: matches nothing
    raw = make_uint4(0u, 0u, 0u, 0u);
    if (wv[ja][jw] && oh >= 0 && oh < d.oh) {
      const int opix = obase[ja][jw] + oh * d.ow;
      pk = *reinterpret_cast<const uint2*>(argmax + (long long)opix * d.c + c);
      raw = *reinterpret_cast<const uint4*>(dy + (long long)opix * d.out_ld + d.out_coff + c);
    }
  };
  // routes the window's gradients whose arg-max is (depth tap of ja, row b, column rw + 2 jw) into g[rw]
  auto route_row = [&](const uint2& pk, const uint4& raw, uint32_t tap_row, int jw, float (&g)[2][8]) {
    const uint32_t f[8] = {raw.x << 16, raw.x & 0xffff0000u, raw.y << 16, raw.y & 0xffff0000u,
                           raw.z << 16, raw.z & 0xffff0000u, raw.w << 16, raw.w & 0xffff0000u};
#pragma unroll
    for (int rw = 0; rw < 2; ++rw) {
      const int e = rw + 2 * jw;
      if (e >= KW) continue;
      const uint32_t tap4 = tap_row + (uint32_t)e * 0x01010101u;
      const uint32_t e0 = __vcmpeq4(pk.x, tap4), e1 = __vcmpeq4(pk.y, tap4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t m = __byte_perm(i < 4 ? e0 : e1, 0u, 0x1111u * (i & 3));
        g[rw][i] += __uint_as_float(f[i] & m);
      }
    }
  };
  uint2 pkT[JA][JW];
  uint4 rawT[JA][JW];
  if (JH == 2) {
#pragma unroll
    for (int ja = 0; ja < JA; ++ja)
#pragma unroll
      for (int jw = 0; jw < JW; ++jw) load_win(ja, jw, qh0 - 1, pkT[ja][jw], rawT[ja][jw]);
  }
  const int iw0 = 2 * qw - d.pw;
  const int in_plane = (n * d.id + idd) * d.ih;
#pragma unroll 1
  for (int qh = qh0; qh < qh1; ++qh) {
    float g[2][2][8];
#pragma unroll
    for (int i = 0; i < 32; ++i) (&g[0][0][0])[i] = 0.f;
#pragma unroll
    for (int ja = 0; ja < JA; ++ja) {
#pragma unroll
      for (int jw = 0; jw < JW; ++jw) {
        uint2 pk;
        uint4 raw;
        load_win(ja, jw, qh, pk, raw);
        if (JH == 2)  // the windows one output row up reach patch row qh with their last row
          route_row(pkT[ja][jw], rawT[ja][jw], tapa[ja] + (uint32_t)(2 * KW) * 0x01010101u, jw, g[0]);
        route_row(pk, raw, tapa[ja], jw, g[0]);
        if (KH >= 2) route_row(pk, raw, tapa[ja] + (uint32_t)KW * 0x01010101u, jw, g[1]);
        if (JH == 2) {
          pkT[ja][jw] = pk;
          rawT[ja][jw] = raw;
        }
      }
    }
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int ih = 2 * qh + rh - d.ph;
      if (ih < 0 || ih >= d.ih) continue;
#pragma unroll
      for (int rw = 0; rw < 2; ++rw) {
        const int iw = iw0 + rw;
        if (iw < 0 || iw >= d.iw) continue;
        const float* gg = g[rh][rw];
        const int ipix = (in_plane + ih) * d.iw + iw;
        uint4 pkd;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(gg[0], gg[1]), h1 = __floats2bfloat162_rn(gg[2], gg[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(gg[4], gg[5]), h3 = __floats2bfloat162_rn(gg[6], gg[7]);
        pkd.x = *reinterpret_cast<uint32_t*>(&h0);
        pkd.y = *reinterpret_cast<uint32_t*>(&h1);
        pkd.z = *reinterpret_cast<uint32_t*>(&h2);
        pkd.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dx + (long long)ipix * d.in_ld + d.in_coff + c) = pkd;
      }
    }
  }
}

// Backward as a SCATTER into a shared-memory tile.  The gather kernel above tests every window that covers an
// input element (27 for the 3x3x3 stride-1 branch pools) and is issue bound at ~1000 instructions per 8
// channels; but every output element routes its gradient to exactly ONE input element, so a block that owns a
// tile of the input (all depths x TH rows x all columns x CB channels, in shared memory) can walk the outputs
// whose windows touch the tile once each and add dy at the decoded argmax position with a shared-memory
// atomic: ~10 instructions per channel per OUTPUT.  Depth and width carry their zero padding in the tile
// (winners that lie in the padding land in cells nobody reads - ATen drops them the same way); rows are
// bounds-checked against the tile, which also discards contributions that belong to the neighbouring tile
// (its own block picks them up: halo output rows are read by both).
//
// Shared-memory fp32 atomicAdd is a compare-and-swap loop on this hardware (SASS: ATOMS.CAST.SPIN; measured
// no faster than the gather), 32-bit integer ATOMS.ADD is native.  The tile is therefore block floating
// point: a first pass finds the largest |dy| the block will read, the addends are scaled by a power of two so
// that (windows per cell) x max fits 31 bits - 25 fraction bits below the largest addend's leading bit for
// the 27-window pools, 28 for the 3x3 stride-2 ones - and added as integers.  bf16 addends within 2^-17 of
// the tile maximum are exact; smaller ones are rounded at 2^-26 of the maximum, below the fp32 rounding of a
// sum of that size.  Integer adds commute, so the result does not depend on the order of the atomics.
// A non-finite dy poisons its tile (every cell of the tile becomes NaN).
struct ScatterPlan {
  int td, th, htiles, cb, cgs_shift, tdp, iwp, dslices, fbits;
  uint32_t m_orow, m_irow, m_th;  // ceil(2^32 / divisor) for ow*cgs, iw*cgs, th
};

__device__ __forceinline__ uint32_t fastdiv(uint32_t n, uint32_t magic, uint32_t div) {
  return div == 1u ? n : __umulhi(n, magic);
}
inline uint32_t fastdiv_magic(uint32_t div) { return (uint32_t)(((1ull << 32) + div - 1) / div); }

constexpr int SCATTER_THREADS = 512;

template <typename G>
__global__ void __launch_bounds__(SCATTER_THREADS)
maxpool_bwd_scatter_kernel(ivf_pool_desc d, ScatterPlan p, const __nv_bfloat16* __restrict__ dy,
                           const uint8_t* __restrict__ argmax, const float* __restrict__ acc_in,
                           const __nv_bfloat16* __restrict__ mask_y, const float* __restrict__ mask_scale,
                           void* __restrict__ dx) {
  extern __shared__ __align__(16) int tile[];  // [tdp][th][iwp][cb] fixed point
  __shared__ int lut[64];                      // tap -> tile offset | kh index << 24
  __shared__ uint32_t m_noh_s;
  __shared__ uint32_t amax_s[SCATTER_THREADS / 32];
  const int KD = G::kd(d), KH = G::kh(d), KW = G::kw(d), SD = G::sd(d), SH = G::sh(d), SW = G::sw(d);
  const int CB = p.cb, TH = p.th, IWP = p.iwp;
  const int cgs = 1 << p.cgs_shift;
  const int h0 = blockIdx.x * TH;
  const int c0 = blockIdx.y * CB;
  const int n = blockIdx.z / p.dslices, d0 = blockIdx.z - n * p.dslices;  // d0 = 0 when the tile holds all depths
  // windows whose rows intersect [h0, h0 + TH)
  const int lo_num = h0 + d.ph - (KH - 1);
  const int oh_lo = lo_num <= 0 ? 0 : (lo_num + SH - 1) / SH;
  const int oh_hi = min(d.oh - 1, (h0 + TH - 1 + d.ph) / SH);
  const int noh = max(oh_hi - oh_lo + 1, 0);
  const int nod = p.td == 1 ? 1 : d.od;
  if (threadIdx.x < KD * KH * KW) {
    const int t = threadIdx.x;
    const int a = t / (KH * KW), r = t - a * KH * KW, b = r / KW, e = r - b * KW;
    lut[t] = (((a * TH + b) * IWP + e) * CB) | (b << 24);
  }
  if (threadIdx.x == 64) m_noh_s = noh > 1 ? (uint32_t)(((1ull << 32) + noh - 1) / noh) : 0u;
  const int tile_words = p.tdp * TH * IWP * CB;
  for (int i = threadIdx.x; i < tile_words / 4; i += SCATTER_THREADS) reinterpret_cast<int4*>(tile)[i] = make_int4(0, 0, 0, 0);
  __syncthreads();
  const uint32_t m_noh = m_noh_s;
  const int per_orow = d.ow << p.cgs_shift;
  const int oitems = nod * noh * per_orow;
  auto out_item = [&](int idx, int& od, int& oh, int& ow, int& cg) {
    const int row_id = fastdiv(idx, p.m_orow, per_orow);
    const int rem = idx - row_id * per_orow;
    ow = rem >> p.cgs_shift;
    cg = rem & (cgs - 1);
    const int od_rel = fastdiv(row_id, m_noh, noh);
    oh = oh_lo + row_id - od_rel * noh;
    od = p.td == 1 ? d0 : od_rel;
  };
  // pass 1: largest |dy| (bf16 magnitudes order like their bit patterns)
  uint32_t amax = 0u;
  for (int idx = threadIdx.x; idx < oitems; idx += SCATTER_THREADS) {
    int od, oh, ow, cg;
    out_item(idx, od, oh, ow, cg);
    const int opix = ((n * d.od + od) * d.oh + oh) * d.ow + ow;
    const uint4 raw = *reinterpret_cast<const uint4*>(dy + (long long)opix * d.out_ld + d.out_coff + c0 + (cg << 3));
    amax = __vmaxu2(amax, __vmaxu2(__vmaxu2(raw.x & 0x7fff7fffu, raw.y & 0x7fff7fffu),
                                   __vmaxu2(raw.z & 0x7fff7fffu, raw.w & 0x7fff7fffu)));
  }
  amax = max(amax & 0xffffu, amax >> 16);
  amax = __reduce_max_sync(0xffffffffu, amax);
  if ((threadIdx.x & 31) == 0) amax_s[threadIdx.x >> 5] = amax;
  __syncthreads();
  amax = 0u;
#pragma unroll
  for (int w = 0; w < SCATTER_THREADS / 32; ++w) amax = max(amax, amax_s[w]);
  const bool poisoned = amax >= 0x7f80u;
  const int emax = max((int)(amax >> 7), 64);              // biased exponent of the largest addend
  const float to_fixed = __uint_as_float((uint32_t)(p.fbits + 254 - emax) << 23);   // 2^(fbits + 127 - emax)
  const float from_fixed = __uint_as_float((uint32_t)(emax - p.fbits) << 23);       // its inverse
  if (amax != 0u && !poisoned) {
    for (int idx = threadIdx.x; idx < oitems; idx += SCATTER_THREADS) {
      int od, oh, ow, cg;
      out_item(idx, od, oh, ow, cg);
      const int opix = ((n * d.od + od) * d.oh + oh) * d.ow + ow;
      const int c = c0 + (cg << 3);
      const uint2 pk = *reinterpret_cast<const uint2*>(argmax + (long long)opix * d.c + c);
      const uint4 raw = *reinterpret_cast<const uint4*>(dy + (long long)opix * d.out_ld + d.out_coff + c);
      const int rbase = oh * SH - d.ph - h0;
      const int z0 = p.td == 1 ? 0 : od * SD;
      int* base = tile + ((z0 * TH + rbase) * IWP + ow * SW) * CB + (cg << 3);
      const uint32_t words[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t tap = ((i < 4 ? pk.x : pk.y) >> (8 * (i & 3))) & 0xffu;
        const int L = lut[tap];
        const int r = rbase + (L >> 24);
        const float v = __uint_as_float((i & 1) ? (words[i >> 1] & 0xffff0000u) : (words[i >> 1] << 16));
        const int q = __float2int_rn(v * to_fixed);
        if ((unsigned)r < (unsigned)TH && q != 0) atomicAdd(base + (L & 0xffffff) + i, q);
      }
    }
  }
  __syncthreads();
  // tile -> dx with the epilogue (consumer-sum accumulate, ReLU'/BN' mask)
  {
    const int per_row = d.iw << p.cgs_shift;
    const int items = p.td * TH * per_row;
    const int zpad = p.td == 1 ? 0 : d.pd;
    const float poison = poisoned ? __uint_as_float(0x7fc00000u) : 0.f;
    for (int idx = threadIdx.x; idx < items; idx += SCATTER_THREADS) {
      const int irow = fastdiv(idx, p.m_irow, per_row);
      const int rem = idx - irow * per_row;
      const int iw = rem >> p.cgs_shift, cg = rem & (cgs - 1);
      const int z_rel = fastdiv(irow, p.m_th, TH);
      const int r = irow - z_rel * TH;
      const int ih = h0 + r;
      if (ih >= d.ih) continue;
      const int* src = tile + (((z_rel + zpad) * TH + r) * IWP + iw + d.pw) * CB + (cg << 3);
      const int4 q0 = reinterpret_cast<const int4*>(src)[0], q1 = reinterpret_cast<const int4*>(src)[1];
      const int q[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      float g[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = (float)q[i] * from_fixed + poison;
      const int c = c0 + (cg << 3);
      const int ipix = ((n * d.id + d0 + z_rel) * d.ih + ih) * d.iw + iw;
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
        const uint4 rawy = *reinterpret_cast<const uint4*>(mask_y + (long long)ipix * d.mask_ld + d.mask_coff + c);
        const __nv_bfloat16* y = reinterpret_cast<const __nv_bfloat16*>(&rawy);
        const float4 s0 = *reinterpret_cast<const float4*>(mask_scale + c);
        const float4 s1 = *reinterpret_cast<const float4*>(mask_scale + c + 4);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = __bfloat162float(y[i]) > 0.f ? g[i] * sc[i] : 0.f;
      }
      if (d.flags & IVF_EP_OUT_F32) {
        float* qd = reinterpret_cast<float*>(dx) + o;
        reinterpret_cast<float4*>(qd)[0] = make_float4(g[0], g[1], g[2], g[3]);
        reinterpret_cast<float4*>(qd)[1] = make_float4(g[4], g[5], g[6], g[7]);
      } else {
        store_vec<__nv_bfloat16, 8>(reinterpret_cast<__nv_bfloat16*>(dx) + o, g);
      }
    }
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
          uint8_t* relu_bits, cudaStream_t st) {
  constexpr int V = full_vec<T>();
  long long opix = (long long)d->n * d->od * d->oh * d->ow;
  const int threads = 256;
  if (vec_ok(d, V, in, out, argmax)) {
    long long total = opix * (d->c / V);
    if constexpr (sizeof(T) == 2) {
      const bool nonneg = (d->flags & IVF_POOL_NONNEG) != 0;
      if (fits31(d)) {  // row-block packed kernel
        static const bool col_on = [] {
          const char* e = getenv("IVF_POOL_S1COL");
          return !e || atoi(e) != 0;
        }();
        if (col_on && !relu_bits && d->kd == 3 && d->kh == 3 && d->kw == 3 && d->sd == 1 && d->sh == 1 && d->sw == 1 &&
            d->od == d->id && d->oh == d->ih && d->ow == d->iw && d->od >= 2) {
          // depth segments: enough threads to fill the machine, at least 4 outputs per thread when split
          long long threads_total = (long long)d->n * d->oh * d->ow * (d->c / V);
          int segs = 1;
          while (threads_total * segs < (long long)h->sm_count * 2048 && d->od / (segs * 2) >= 4) segs *= 2;
          const int dseg = (d->od + segs - 1) / segs;
          const int nseg = (d->od + dseg - 1) / dseg;
          const int rows = d->n * nseg * d->oh;
          if (nonneg)
            maxpool_fwd_s1col_kernel<true><<<row_grid(rows, d->ow * (d->c / V)), threads, 0, st>>>(
                *d, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, argmax, rows, dseg);
          else
            maxpool_fwd_s1col_kernel<false><<<row_grid(rows, d->ow * (d->c / V)), threads, 0, st>>>(
                *d, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, argmax, rows, dseg);
          IVF_LAUNCHED(h);
          return IVF_OK;
        }
        static const bool rows_on = [] {
          const char* e = getenv("IVF_POOL_ROWS");
          return !e || atoi(e) != 0;
        }();
        // measured (B200, tools/pool_bench.py): equal to the one-row kernel at 8 clips, 4 % faster at 32 for the
        // 1x3x3 pools; slower for the 27- and 8-tap pools (4a, 5a), which keep the one-row kernel
        if (rows_on && !relu_bits && d->kd == 1 && d->sh == 2 && d->sw == 2 && d->oh >= 4) {
          // row segments: enough threads for ~1.5 full-occupancy waves, at least two rows per thread
          const long long per_seg = (long long)d->n * d->od * d->ow * (d->c / V);
          static const int waves = [] {
            const char* e = getenv("IVF_POOL_ROWS_WAVES");  // full-occupancy waves to aim for
            return e ? std::max(1, atoi(e)) : 4;
          }();
          long long want = ((long long)h->sm_count * 2048 * waves + per_seg - 1) / per_seg;
          int nseg = (int)std::max<long long>(1, std::min<long long>(want, d->oh / 2));
          const int rseg = (d->oh + nseg - 1) / nseg;
          nseg = (d->oh + rseg - 1) / rseg;
          const int items = d->n * d->od * nseg;
          const dim3 rgrid = row_grid(items, d->ow * (d->c / V));
          bool done = true;
#define IVF_FWD_ROWS(GEO)                                                                                          \
  do {                                                                                                             \
    if (nonneg)                                                                                                    \
      maxpool_fwd_rows_kernel<GEO, true><<<rgrid, threads, 0, st>>>(*d, (const __nv_bfloat16*)in,                  \
                                                                    (__nv_bfloat16*)out, argmax, items, rseg, nseg); \
    else                                                                                                           \
      maxpool_fwd_rows_kernel<GEO, false><<<rgrid, threads, 0, st>>>(*d, (const __nv_bfloat16*)in,                 \
                                                                     (__nv_bfloat16*)out, argmax, items, rseg, nseg); \
  } while (0)
          if (d->kd == 1 && d->kh == 3 && d->kw == 3 && d->sd == 1) IVF_FWD_ROWS(PoolGeo<1 COMMA 3 COMMA 3 COMMA 1 COMMA 2 COMMA 2>);
          else if (d->kd == 3 && d->kh == 3 && d->kw == 3 && d->sd == 2) IVF_FWD_ROWS(PoolGeo<3 COMMA 3 COMMA 3 COMMA 2 COMMA 2 COMMA 2>);
          else if (d->kd == 2 && d->kh == 2 && d->kw == 2 && d->sd == 2) IVF_FWD_ROWS(PoolGeo<2 COMMA 2 COMMA 2 COMMA 2 COMMA 2 COMMA 2>);
          else done = false;
#undef IVF_FWD_ROWS
          if (done) {
            IVF_LAUNCHED(h);
            return IVF_OK;
          }
        }
        const int rows = d->n * d->od * d->oh;
        const dim3 fgrid = row_grid(rows, d->ow * (d->c / V));
        if (relu_bits)
          IVF_POOL_GEO_DISPATCH((maxpool_fwd_bf16x8_kernel<G, 1><<<fgrid, threads, 0, st>>>(
              *d, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, argmax, relu_bits, rows)));
        else if (nonneg)
          IVF_POOL_GEO_DISPATCH((maxpool_fwd_bf16x8_kernel<G, 2><<<fgrid, threads, 0, st>>>(
              *d, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, argmax, nullptr, rows)));
        else
          IVF_POOL_GEO_DISPATCH((maxpool_fwd_bf16x8_kernel<G, 0><<<fgrid, threads, 0, st>>>(
              *d, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, argmax, nullptr, rows)));
        IVF_LAUNCHED(h);
        return IVF_OK;
      }
    }
    IVF_REQUIRE(!relu_bits, "maxpool(fwd): the ReLU bit mask needs the packed bf16 kernel (bf16, 31-bit offsets)");
    IVF_POOL_GEO_DISPATCH((maxpool_fwd_kernel<T, V, G><<<pool_blocks(h, total), threads, 0, st>>>(
        *d, (const T*)in, (T*)out, argmax, total)));
  } else {
    IVF_REQUIRE(!relu_bits, "maxpool(fwd): the ReLU bit mask needs 8-channel aligned bf16 tensors");
    long long total = opix * d->c;
    maxpool_fwd_kernel<T, 1, PoolGeoDyn><<<pool_blocks(h, total), threads, 0, st>>>(*d, (const T*)in, (T*)out,
                                                                                  argmax, total);
  }
  IVF_LAUNCHED(h);
  return IVF_OK;
}

// Tile plan of the scatter backward; false if the shape does not suit it (then the gather kernel runs).
bool scatter_plan(const ivf_pool_desc* d, ScatterPlan* p) {
  static const int limit_kb = [] {
    const char* e = getenv("IVF_POOL_SCATTER_KB");  // 0 disables the scatter kernel
    return e ? atoi(e) : 64;
  }();
  if (limit_kb <= 0 || d->kd * d->kh * d->kw > 64 || d->c % 8) return false;
  if (d->kd == 1 && !(d->sd == 1 && d->pd == 0 && d->od == d->id)) return false;
  p->td = d->kd == 1 ? 1 : d->id;
  p->dslices = d->kd == 1 ? d->id : 1;
  p->tdp = d->kd == 1 ? 1 : (d->od - 1) * d->sd + d->kd;
  p->iwp = (d->ow - 1) * d->sw + d->kw;
  if (p->tdp < d->id + d->pd || p->iwp < d->iw + d->pw) {  // inputs beyond the last window: pad up
    p->tdp = std::max(p->tdp, d->kd == 1 ? 1 : d->id + d->pd);
    p->iwp = std::max(p->iwp, d->iw + d->pw);
  }
  p->cb = d->c % 16 == 0 ? 16 : 8;
  p->cgs_shift = p->cb == 16 ? 1 : 0;
  const long long row_floats = (long long)p->tdp * p->iwp * p->cb;
  const long long limit = (long long)limit_kb * 256;
  if (row_floats > limit) return false;
  int th = (int)std::min<long long>(d->ih, limit / row_floats);
  p->htiles = (d->ih + th - 1) / th;
  p->th = (d->ih + p->htiles - 1) / p->htiles;
  if ((long long)p->tdp * p->th * p->iwp * p->cb >= (1 << 24)) return false;
  int cover = 1;  // windows that can contain one element
  cover *= (d->kd + d->sd - 1) / d->sd;
  cover *= (d->kh + d->sh - 1) / d->sh;
  cover *= (d->kw + d->sw - 1) / d->sw;
  // measured (B200): 1.5x faster than the gather for the 27-window stride-1 pools, slower for the stride-2
  // pools whose elements sit in at most 8 windows
  static const int min_cover = [] {
    const char* e = getenv("IVF_POOL_SCATTER_MIN_COVER");
    return e ? atoi(e) : 9;
  }();
  if (cover < min_cover) return false;
  int hb = 0;
  while ((1 << hb) < cover) ++hb;
  p->fbits = 30 - hb;  // |addend| < 2^(fbits+1), cover of them < 2^31
  p->m_orow = fastdiv_magic((uint32_t)(d->ow << p->cgs_shift));
  p->m_irow = fastdiv_magic((uint32_t)(d->iw << p->cgs_shift));
  p->m_th = fastdiv_magic((uint32_t)p->th);
  return (long long)d->n * p->dslices < 65536 && d->c / p->cb < 65536;
}

template <typename G>
int launch_bwd_scatter(ivf_handle* h, const ivf_pool_desc* d, const ScatterPlan& p, const void* dy,
                       const uint8_t* argmax, const float* acc_in, const void* mask_y, const float* mask_scale,
                       void* dx, cudaStream_t st) {
  const size_t smem = (size_t)p.tdp * p.th * p.iwp * p.cb * sizeof(int);
  static bool attr_done[16] = {};
  const int dev = h->device & 15;
  if (!attr_done[dev]) {
    IVF_CUDA(cudaFuncSetAttribute(maxpool_bwd_scatter_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  200 * 1024));
    attr_done[dev] = true;
  }
  dim3 grid(p.htiles, d->c / p.cb, d->n * p.dslices);
  maxpool_bwd_scatter_kernel<G><<<grid, SCATTER_THREADS, smem, st>>>(*d, p, (const __nv_bfloat16*)dy, argmax, acc_in,
                                                         (const __nv_bfloat16*)mask_y, mask_scale, dx);
  IVF_LAUNCHED(h);
  return IVF_OK;
}

template <typename T>
int bwd_t(ivf_handle* h, const ivf_pool_desc* d, const void* dy, const uint8_t* argmax,
          const float* acc_in, const void* mask_y, const uint8_t* relu_bits, const float* mask_scale, void* dx,
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
        ScatterPlan sp;
        if (scatter_plan(d, &sp)) {
          int rc = IVF_OK;
          IVF_POOL_GEO_DISPATCH((rc = launch_bwd_scatter<G>(h, d, sp, dy, argmax, acc_in, mask_y, mask_scale, dx, st)));
          return rc;
        }
        static const bool patch_on = [] {
          const char* e = getenv("IVF_POOL_S2PATCH");
          return !e || atoi(e) != 0;
        }();
        if (patch_on && d->sh == 2 && d->sw == 2 && d->kh >= 2 && d->kh <= 3 && d->kw >= 2 && d->kw <= 3 &&
            d->kd <= 3 && d->sd <= 2 && d->kd >= d->sd) {
          const int qhn = (d->ih + d->ph + 1) / 2, qwn = (d->iw + d->pw + 1) / 2;
          const int prow = d->n * d->id * qhn;
          bool done = true;
          using GeoStem = PoolGeo<1, 3, 3, 1, 2, 2>;
          using Geo3s2 = PoolGeo<3, 3, 3, 2, 2, 2>;
          using Geo2s2 = PoolGeo<2, 2, 2, 2, 2, 2>;
          static const bool route_on = [] {
            const char* e = getenv("IVF_POOL_S2ROUTE");
            return !e || atoi(e) != 0;
          }();
          // pure routing: row-segment kernel.  Measured (tools/pool_bench.py, 8 / 32 clips): 2a 51 -> 41 / 182 -> 129 us,
          // 3a 40 -> 33 / 137 -> 100 us; the two-depth-candidate 4a pool (100 registers) and 5a are not faster
          if (route_on && d->flags == 0 && !relu_bits && d->kd == 1 && qhn >= 4) {
            const long long per_seg = (long long)d->n * d->id * qwn * (d->c / V);
            static const int waves = [] {
              const char* e = getenv("IVF_POOL_ROUTE_WAVES");
              return e ? std::max(1, atoi(e)) : 3;
            }();
            long long want = ((long long)h->sm_count * 2048 * waves + per_seg - 1) / per_seg;
            int nseg = (int)std::max<long long>(1, std::min<long long>(want, qhn / 2));
            const int qseg = (qhn + nseg - 1) / nseg;
            nseg = (qhn + qseg - 1) / qseg;
            const int items = d->n * d->id * nseg;
            const dim3 rgrid = row_grid(items, qwn * (d->c / V));
#define IVF_S2ROUTE(GEO)                                                                                  \
  maxpool_bwd_s2route_kernel<GEO><<<rgrid, threads, 0, st>>>(*d, (const __nv_bfloat16*)dy, argmax,        \
                                                             (__nv_bfloat16*)dx, items, qseg, nseg, qhn, qwn)
            bool routed = true;
            if (d->kd == 1 && d->kh == 3 && d->kw == 3 && d->sd == 1) IVF_S2ROUTE(GeoStem);
            else if (d->kd == 3 && d->kh == 3 && d->kw == 3 && d->sd == 2) IVF_S2ROUTE(Geo3s2);
            else if (d->kd == 2 && d->kh == 2 && d->kw == 2 && d->sd == 2) IVF_S2ROUTE(Geo2s2);
            else routed = false;
#undef IVF_S2ROUTE
            if (routed) {
              IVF_LAUNCHED(h);
              return IVF_OK;
            }
          }
          const dim3 pgrid = row_grid(prow, qwn * (d->c / V));
#define IVF_S2PATCH(GEO)                                                                                   \
  maxpool_bwd_s2patch_kernel<GEO><<<pgrid, threads, 0, st>>>(*d, (const __nv_bfloat16*)dy, argmax, acc_in, \
                                                             (const __nv_bfloat16*)mask_y, relu_bits, mask_scale, dx, prow, qhn, qwn)
          if (d->kd == 1 && d->kh == 3 && d->kw == 3 && d->sd == 1) IVF_S2PATCH(GeoStem);
          else if (d->kd == 3 && d->kh == 3 && d->kw == 3 && d->sd == 2) IVF_S2PATCH(Geo3s2);
          else if (d->kd == 2 && d->kh == 2 && d->kw == 2 && d->sd == 2) IVF_S2PATCH(Geo2s2);
          else done = false;
#undef IVF_S2PATCH
          if (done) {
            IVF_LAUNCHED(h);
            return IVF_OK;
          }
        }
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

extern "C" int ivf_maxpool3d_fwd_bits(ivf_handle* h, const ivf_pool_desc* d, const void* in, void* out,
                                      uint8_t* argmax, uint8_t* relu_bits, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && d && in && out, "ivf_maxpool3d_fwd: null argument");
  int rc = check_pool(d);
  if (rc) return rc;
  if (relu_bits) {
    // every input element must lie in the patch of some window origin, or its bits would never be written
    IVF_REQUIRE(d->dtype == IVF_BF16 && d->c % 8 == 0, "ivf_maxpool3d_fwd_bits: bf16, channels a multiple of 8");
    IVF_REQUIRE((d->id - 1 + d->pd) / d->sd < d->od && (d->ih - 1 + d->ph) / d->sh < d->oh &&
                    (d->iw - 1 + d->pw) / d->sw < d->ow && d->kd >= d->sd && d->kh >= d->sh && d->kw >= d->sw,
                "ivf_maxpool3d_fwd_bits: windows do not cover the input");
  }
  cudaStream_t st = (cudaStream_t)stream;
  return d->dtype == IVF_F32 ? fwd_t<float>(h, d, in, out, argmax, relu_bits, st)
                             : fwd_t<__nv_bfloat16>(h, d, in, out, argmax, relu_bits, st);
}

extern "C" int ivf_maxpool3d_fwd(ivf_handle* h, const ivf_pool_desc* d, const void* in, void* out,
                                 uint8_t* argmax, void* stream) {
  IVF_ON_DEVICE(h);
  return ivf_maxpool3d_fwd_bits(h, d, in, out, argmax, nullptr, stream);
}

extern "C" int ivf_maxpool3d_bwd_bits(ivf_handle* h, const ivf_pool_desc* d, const void* dy,
                                      const uint8_t* argmax, const float* acc_in, const void* mask_y,
                                      const uint8_t* relu_bits, const float* mask_scale, void* dx, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && d && dy && argmax && dx, "ivf_maxpool3d_bwd: null argument");
  int rc = check_pool(d);
  if (rc) return rc;
  if (d->flags & IVF_EP_ACCUM) IVF_REQUIRE(acc_in, "ivf_maxpool3d_bwd: ACCUM needs acc_in");
  if (d->flags & IVF_EP_MASK) IVF_REQUIRE(mask_y && mask_scale, "ivf_maxpool3d_bwd: MASK needs mask_y/mask_scale");
  cudaStream_t st = (cudaStream_t)stream;
  return d->dtype == IVF_F32
             ? bwd_t<float>(h, d, dy, argmax, acc_in, mask_y, relu_bits, mask_scale, dx, st)
             : bwd_t<__nv_bfloat16>(h, d, dy, argmax, acc_in, mask_y, relu_bits, mask_scale, dx, st);
}

extern "C" int ivf_maxpool3d_bwd(ivf_handle* h, const ivf_pool_desc* d, const void* dy,
                                 const uint8_t* argmax, const float* acc_in, const void* mask_y,
                                 const float* mask_scale, void* dx, void* stream) {
  IVF_ON_DEVICE(h);
  return ivf_maxpool3d_bwd_bits(h, d, dy, argmax, acc_in, mask_y, nullptr, mask_scale, dx, stream);
}
