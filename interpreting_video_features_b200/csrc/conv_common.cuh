// Device helpers shared by the two tcgen05 convolution kernels (conv_tc.cu: TMA-im2col operand;
// conv_slab.cu: shared-memory resident halo slab operand): mbarrier / TMA / UMMA / TMEM wrappers and the
// fused epilogue (eval-BatchNorm affine, ReLU, fp32 consumer sum, ReLU'/BN' mask, bf16 or fp32 store).
#pragma once
#include "common.cuh"

namespace ivf_tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done;
  // try_wait suspends in hardware for a bounded time per call; a pipeline bug must not hang the
  // GPU, so after ~2^24 failed probes (seconds) the CTA traps and the launch reports an error.
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 24)) {
      printf("libivf: mbarrier wait timed out (block %d,%d thread %d parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, parity);
      __trap();
    }
  } while (!done);
}

// start fetching a tensor map (a __grid_constant__ kernel parameter) before the first TMA that uses it
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// tiled 5-D box (c, w, h, d, n): out-of-range coordinates are zero filled = the 'same' padding
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c,
                                            int w, int h, int d, int n) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n)
      : "memory");
}

__device__ __forceinline__ void tma_load_im2col_5d(uint32_t dst, const CUtensorMap* map,
                                                   uint64_t* bar, int c, int w, int h, int d, int n,
                                                   uint16_t ow, uint16_t oh, uint16_t od) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(d),
      "r"(n), "h"(ow), "h"(oh), "h"(od)
      : "memory");
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout, version 1):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [49,52) base offset (0: the swizzle is
// applied on absolute shared-memory address bits, so a start address off the 8-row atom needs none),
// [61,64) swizzle type.  The issuing thread keeps the constant high word and adds to the low word: one 32-bit
// add per operand per MMA (rebuilding 64-bit descriptors in a divergent single-thread region bounded the
// first slab kernel at ~200 cycles per MMA).  lo = (smem_addr >> 4) | (1 << 16) [LBO = 1, unused for swizzled
// K-major];  hi = SBO>>4 | version<<14 | layout<<29.
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t smem_desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
__device__ __forceinline__ void umma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- cta_group::2 (CTA pair on one TPC: M = 256 over two SMs, each SM holds its own 128 activation rows
// and HALF of the weight rows, so the per-SM operand fetch per MMA drops from 4 KB + 32N to 4 KB + 16N bytes)
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER CTA's barrier at this offset (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
// TMA loads whose completion bytes are counted on the leader CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c,
                                                 int w, int h, int d, int n) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c), "r"(w), "r"(h),
      "r"(d), "r"(n)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_lo_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All K steps of one (tap, accumulator) in ONE asm block: the accumulate predicate is set once, the descriptor
// low words advance by 2 (32 bytes) per step inside the block.  The issue loop is a chain of dependent
// uniform-datapath instructions run by a single thread; per-MMA set-up instructions are what bounds the layers
// with few output channels (tools/mma_bench.cu), so every instruction taken out of it counts.
template <int KS, int NCTA>
__device__ __forceinline__ void umma_bf16_ksteps(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                 uint32_t idesc, uint32_t accumulate) {
  static_assert(KS >= 1 && KS <= 4, "K steps per stage: up to 4 (128-byte rows)");
#define IVF_MMA1(CG, PRED) "mov.b64 da, {al, %3};\n" "mov.b64 db, {bl, %3};\n" \
                           "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, " PRED ";\n"
#define IVF_MMA_NEXT(CG) "add.u32 al, al, 2;\n" "add.u32 bl, bl, 2;\n" IVF_MMA1(CG, "pt")
#define IVF_MMA_HEAD "{\n" ".reg .pred p, pt;\n" ".reg .b64 da, db;\n" ".reg .b32 al, bl;\n" \
                     "setp.ne.b32 p, %5, 0;\n" "setp.eq.b32 pt, %4, %4;\n" "mov.b32 al, %1;\n" "mov.b32 bl, %2;\n"
#define IVF_MMA_ARGS ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory"
  if constexpr (NCTA == 2) {
    if constexpr (KS == 1) asm volatile(IVF_MMA_HEAD IVF_MMA1("2", "p") "}\n" IVF_MMA_ARGS);
    if constexpr (KS == 2) asm volatile(IVF_MMA_HEAD IVF_MMA1("2", "p") IVF_MMA_NEXT("2") "}\n" IVF_MMA_ARGS);
    if constexpr (KS == 3)
      asm volatile(IVF_MMA_HEAD IVF_MMA1("2", "p") IVF_MMA_NEXT("2") IVF_MMA_NEXT("2") "}\n" IVF_MMA_ARGS);
    if constexpr (KS == 4)
      asm volatile(IVF_MMA_HEAD IVF_MMA1("2", "p") IVF_MMA_NEXT("2") IVF_MMA_NEXT("2") IVF_MMA_NEXT("2") "}\n" IVF_MMA_ARGS);
  } else {
    if constexpr (KS == 1) asm volatile(IVF_MMA_HEAD IVF_MMA1("1", "p") "}\n" IVF_MMA_ARGS);
    if constexpr (KS == 2) asm volatile(IVF_MMA_HEAD IVF_MMA1("1", "p") IVF_MMA_NEXT("1") "}\n" IVF_MMA_ARGS);
    if constexpr (KS == 3)
      asm volatile(IVF_MMA_HEAD IVF_MMA1("1", "p") IVF_MMA_NEXT("1") IVF_MMA_NEXT("1") "}\n" IVF_MMA_ARGS);
    if constexpr (KS == 4)
      asm volatile(IVF_MMA_HEAD IVF_MMA1("1", "p") IVF_MMA_NEXT("1") IVF_MMA_NEXT("1") IVF_MMA_NEXT("1") "}\n" IVF_MMA_ARGS);
  }
#undef IVF_MMA1
#undef IVF_MMA_NEXT
#undef IVF_MMA_HEAD
#undef IVF_MMA_ARGS
}

// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// one lane of a converged warp (elect.sync): role loops run warp-uniform and predicate the single-thread
// instructions on this, which keeps descriptors in uniform registers (no per-MMA election loop in SASS)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait: issue several, then tmem_ld_wait() once
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Epilogue of 16 consecutive produced channels [nb, nb+16) of one output pixel.
//   r          : the fp32 accumulators as loaded from TMEM
//   sc/sh/ms   : per-channel scale / shift / mask-scale of these 16 channels (shared memory)
//   out_row    : element offset of the pixel's channel 0 in `out` (and acc_in); mask_row likewise
struct EpilogueArgs {
  int cout, flags;
  const float* acc_in;
  const __nv_bfloat16* mask_y;
  void* out;
  // != 0: shared-memory address of this lane's 32-byte row in a staging box; the 16 bf16 results of a chunk
  // go there and the warp sends the box with one TMA store (rows and channels past the tensor are clipped by
  // the tensor map, so no validity tests apply)
  uint32_t stage_smem = 0;
  // IVF_EP_LSTM (ivf_conv3d_lstm): cell state in / out (fp32 [pix][hid]) and hidden state out (bf16 [pix][hid]);
  // `out` then is the fp32 gate-activation buffer and acc_in the x-convolution's pre-activations, both
  // [pix][4*hid] with UNIT-MAJOR channels (4*k + gate), so out_row / 4 is the pixel's offset in the hid-wide rows
  const float* lstm_c_prev = nullptr;
  float* lstm_c_next = nullptr;
  __nv_bfloat16* lstm_h_next = nullptr;
};

__device__ __forceinline__ float ivf_sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }

// TMA store of a staged box: [tensor map, {channel, row}] <- shared memory; bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(src_smem)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Global operands of one full 16-channel chunk's epilogue (fp32 consumer sum, bf16 ReLU mask), fetched one
// chunk AHEAD of their use so that their latency (~1 us) overlaps the TMEM loads and math of the previous
// chunk instead of serialising every chunk (the data-gradient epilogues were bounded by exactly that).
struct EpiPre {
  float4 ac[4];
  uint4 mk[2];
};
// FLAGS >= 0: the launch's epilogue flags as a compile-time constant (the kernels dispatch once per launch on the
// common combinations): the flag tests fold away and the chunk becomes straight-line code - with run-time flags a
// chunk executed ~250 instructions for ~60 of arithmetic, and three epilogue warps per scheduler made that the
// bound of the 1x1x1 GEMMs (trace: 0.42 us of "math" per 16-channel chunk).
template <int FLAGS = -1>
__device__ __forceinline__ void epilogue_prefetch(const EpilogueArgs& e_, int nb, size_t out_row, size_t mask_row,
                                                  bool active, EpiPre& pre) {
  struct { int cout, flags; const float* acc_in; const __nv_bfloat16* mask_y; } e = {
      e_.cout, FLAGS >= 0 ? FLAGS : e_.flags, e_.acc_in, e_.mask_y};
  if (!active || nb + 16 > e.cout) return;  // partial chunks take the scalar path at use
  if (e.flags & IVF_EP_ACCUM) {
    const float4* a4 = reinterpret_cast<const float4*>(e.acc_in + out_row + nb);
#pragma unroll
    for (int j = 0; j < 4; ++j) pre.ac[j] = a4[j];
  }
  if (e.flags & IVF_EP_MASK) {
    const uint4* m4 = reinterpret_cast<const uint4*>(e.mask_y + mask_row + nb);
    pre.mk[0] = m4[0];
    pre.mk[1] = m4[1];
  }
}

// 16 consecutive floats of a 16-byte aligned shared-memory vector (four LDS.128 instead of sixteen LDS.32)
__device__ __forceinline__ void lds16(const float* p, float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = reinterpret_cast<const float4*>(p)[j];
    v[4 * j] = t.x;
    v[4 * j + 1] = t.y;
    v[4 * j + 2] = t.z;
    v[4 * j + 3] = t.w;
  }
}

// sc / sh / ms must be 16-byte aligned (the kernels' per-channel vectors are, and chunks start at multiples of 16)
template <bool LSTM = false, int FLAGS = -1>
__device__ __forceinline__ void epilogue_chunk16(const EpilogueArgs& e_, const uint32_t (&r)[16], int nb,
                                                 const float* sc, const float* sh, const float* ms,
                                                 size_t out_row, size_t mask_row, const EpiPre& pre) {
  EpilogueArgs e = e_;
  if (FLAGS >= 0) e.flags = FLAGS;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  const bool full = nb + 16 <= e.cout;
  if (e.flags & IVF_EP_ACCUM) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[4 * j + 0] += pre.ac[j].x;
        v[4 * j + 1] += pre.ac[j].y;
        v[4 * j + 2] += pre.ac[j].z;
        v[4 * j + 3] += pre.ac[j].w;
      }
    } else {
      for (int j = 0; j < 16; ++j)
        if (nb + j < e.cout) v[j] += e.acc_in[out_row + nb + j];
    }
  }
  if constexpr (LSTM) {  // a separate kernel instantiation: the gate math costs the other epilogues no registers
    // ConvLSTM gates on four whole hidden units (pt/models/convolution_lstm.py:38-48, zero peepholes :50-54):
    // v = x-conv pre-activation (acc_in) + h-conv accumulator, channels [i f c o] per unit.  Accurate expf/tanhf:
    // tanh.approx perturbs the hidden state enough to flip 2x2 max-pool decisions downstream (clstm.cu).
    const size_t hrow = out_row / 4 + (size_t)(nb >> 2);
    float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.lstm_c_prev) cp = *reinterpret_cast<const float4*>(e.lstm_c_prev + hrow);
    const float cpv[4] = {cp.x, cp.y, cp.z, cp.w};
    float cn[4], hn[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float gi = ivf_sigmoid_f(v[4 * u]), gf = ivf_sigmoid_f(v[4 * u + 1]);
      const float gg = tanhf(v[4 * u + 2]), go = ivf_sigmoid_f(v[4 * u + 3]);
      cn[u] = fmaf(gf, cpv[u], gi * gg);
      hn[u] = go * tanhf(cn[u]);
      v[4 * u] = gi;
      v[4 * u + 1] = gf;
      v[4 * u + 2] = gg;
      v[4 * u + 3] = go;
    }
    *reinterpret_cast<float4*>(e.lstm_c_next + hrow) = make_float4(cn[0], cn[1], cn[2], cn[3]);
    __nv_bfloat162 h01 = __floats2bfloat162_rn(hn[0], hn[1]), h23 = __floats2bfloat162_rn(hn[2], hn[3]);
    uint2 hp;
    hp.x = *reinterpret_cast<uint32_t*>(&h01);
    hp.y = *reinterpret_cast<uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(e.lstm_h_next + hrow) = hp;
    float* o = reinterpret_cast<float*>(e.out) + out_row + nb;  // activated gates for the backward pass
#pragma unroll
    for (int j = 0; j < 4; ++j)
      reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    return;
  }
  if (e.flags & IVF_EP_AFFINE) {
    float scv[16], shv[16];
    lds16(sc, scv);
    lds16(sh, shv);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], scv[j], shv[j]);
  }
  if (e.flags & IVF_EP_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (e.flags & IVF_EP_MASK) {
    if (full) {
      float msv[16];
      lds16(ms, msv);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&pre.mk[hh]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int jj = hh * 8 + j;
          v[jj] = __bfloat162float(mb[j]) > 0.f ? v[jj] * msv[jj] : 0.f;
        }
      }
    } else {
      for (int j = 0; j < 16; ++j)
        if (nb + j < e.cout)
          v[j] = __bfloat162float(e.mask_y[mask_row + nb + j]) > 0.f ? v[j] * ms[j] : 0.f;
    }
  }
  if (e.flags & IVF_EP_OUT_F32) {
    float* o = reinterpret_cast<float*>(e.out) + out_row + nb;
    if (full) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      for (int j = 0; j < 16; ++j)
        if (nb + j < e.cout) o[j] = v[j];
    }
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + out_row + nb;
    if (e.stage_smem) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&b2);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(e.stage_smem), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                   "r"(pk[3])
                   : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(e.stage_smem + 16u), "r"(pk[4]), "r"(pk[5]),
                   "r"(pk[6]), "r"(pk[7])
                   : "memory");
    } else if (full) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&b2);
      }
      reinterpret_cast<uint4*>(o)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      reinterpret_cast<uint4*>(o)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    } else {
      for (int j = 0; j < 16; ++j)
        if (nb + j < e.cout) o[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

}  // namespace ivf_tc

// ---- host side shared by both kernels: driver entry points for tensor-map encoding -------------
typedef CUresult (*IvfEncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                      const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                      cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*IvfEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
extern IvfEncodeIm2colFn ivf_encode_im2col;
extern IvfEncodeTiledFn ivf_encode_tiled;
int ivf_load_driver_entry_points();
