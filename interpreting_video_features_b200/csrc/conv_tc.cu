// bf16 implicit-GEMM 3-D convolution on the 5th-gen tensor cores (tcgen05 + TMEM), operands
// staged by TMA: the activation tile with the im2col tensor-map mode (the hardware walks
// 128 consecutive output pixels across W/H/D/N, adds the filter-tap offset and zero-fills
// the 'same' padding), the K-major packed weights with a tiled map.  fp32 accumulation in
// TMEM; the epilogue (tcgen05.ld) fuses eval-BatchNorm affine + ReLU for the forward pass and
// sum-over-consumers + ReLU'/BN' masking for the data-gradient pass, and writes straight into
// the channel slice of the concat buffer.
//
// Stands in for aten::conv3d / convolution_backward(data) + native_batch_norm + relu +
// constant_pad_nd + cat at pt/models/I3D_doubled.py:96-118,146 and for the gate convolutions of
// pt/models/convolution_lstm.py:25-32.
//
// GEMM view: D[M=128 pixels][N=cout tile] += A[128][kchunk] * B[N][kchunk]^T per (tap, channel
// chunk).  Warp roles: warp 0 = TMA producer (+TMEM alloc), warp 1 = MMA issuer (one lane),
// warps 2-5 = epilogue (TMEM lane quarter = warp_idx % 4).
#include "conv_common.cuh"

#include <cstring>
#include <type_traits>

IvfEncodeIm2colFn ivf_encode_im2col = nullptr;
IvfEncodeTiledFn ivf_encode_tiled = nullptr;

namespace {

using namespace ivf_tc;

constexpr int TILE_M = 128;
// warp 0 = TMA producer, warp 1 = MMA issuer, then EPI_WARPS epilogue warps: four TMEM lane quarters (warp % 4)
// x EPI_GROUPS column groups that take the 16-channel chunks round-robin.  The epilogue is a latency chain per
// warp (TMEM load -> a few dependent ALU ops -> store); with one warp per scheduler it ran at ~0.2 IPC and was
// the whole kernel for the 1x1x1 layers, so it gets four warps per scheduler.
constexpr int EPI_GROUPS = 3;
constexpr int EPI_WARPS = 4 * EPI_GROUPS;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int MAX_STAGES = 12;
constexpr int EPI_STAGING_BYTES = EPI_WARPS * 2 * 3072;  // per warp two buffers of [32 x 16 fp32 | 32 x 16 bf16]

struct TcParams {
  int M;                 // n*od*oh*ow
  int od, oh, ow;
  int cout;              // real produced channels
  int bn;                // UMMA N of this launch (multiple of 16, <= 256)
  int kh, kw;            // tap decode
  int ntaps, cchunks;    // K iterations = ntaps*cchunks
  int cin, cin_pad;      // real channels (last chunk issues only its valid K steps); weight K pitch per tap
  int sd, sh, sw;
  int pd, ph, pw;
  int out_ld, out_coff;
  int mask_ld, mask_coff;
  int flags;
  int stages;
  int tmem_cols;
  uint32_t a_stage_bytes, b_stage_bytes;  // smem pitch of the A / B part of a stage
  uint32_t tx_bytes;                      // bytes the two TMA boxes of a stage deliver
  // split GEMMs (ivf_conv3d_split): produced channels >= split_cout go to a second tensor; channel chunks
  // >= ksplit are gathered from a second tensor (its chunk index restarts at 0)
  int split_cout, out2_ld, out2_coff;
  int ksplit;
  int mtiles, grid_x;        // 128-pixel tiles; a CTA walks blockIdx.x, blockIdx.x + gridDim.x, ...
  int acc_stages, acc_cols;  // TMEM accumulators (1|2) and the column pitch between them
  int tma_store;             // bf16 results staged in shared memory and written by TMA stores
  int tma_epi;               // data-gradient epilogue operands (fp32 consumer sum, bf16 ReLU mask) loaded by TMA
  uint32_t epi_off;          // their staging area, byte offset from the aligned dynamic shared-memory base
  long long* trace;          // IVF_TC_TRACE=1: globaltimer stamps of CTA (0, 0) (diagnostic), else null
};

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_TRACE(slot)                                                                   \
  do {                                                                                   \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) p.trace[slot] = gtime(); \
  } while (0)

// KCH = channels per K stage: 64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B, 16 -> SWIZZLE_32B
template <int KCH>
struct KTraits {
  static constexpr uint32_t row_bytes = KCH * 2;
  static constexpr uint32_t sbo = 8 * row_bytes;  // 8-row core-matrix group pitch
  static constexpr uint32_t layout = KCH == 64 ? 2u : (KCH == 32 ? 4u : 6u);
  static constexpr int ksteps = KCH / 16;
};

// Persistent over M tiles: CTA (x, ntile) walks the 128-pixel tiles x, x + gridDim.x, ... of its N tile.  The
// accumulator is double-buffered in TMEM whenever two buffers fit (p.acc_stages), so the epilogue of one tile -
// for the 1x1x1 layers, a handful of K steps per tile, the epilogue IS most of the work - runs while the
// producer and the MMA issuer are already on the next tile; the shared-memory ring keeps running across tiles.
template <int KCH>
__global__ void __launch_bounds__(NUM_THREADS)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmO,
               const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmM, const TcParams p, const float* __restrict__ scale,
               const float* __restrict__ shift, const float* __restrict__ acc_in,
               const __nv_bfloat16* __restrict__ mask_y, const float* __restrict__ mask_scale,
               void* __restrict__ out, void* __restrict__ out2) {
  using KT = KTraits<KCH>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(8) uint64_t epi_bar[EPI_WARPS][4];  // epilogue-operand boxes of a warp have landed
  __shared__ __align__(16) float s_scale[256], s_shift[256], s_mscale[256];
  // bf16 results leave through TMA stores: per epilogue warp two boxes of 32 rows x 16 channels
  __shared__ __align__(128) uint8_t stage_buf[EPI_WARPS][2][32 * 32];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntile = blockIdx.y;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  const int kiters = p.ntaps * p.cchunks;

  ivf_pdl_trigger();
  if (threadIdx.x == 0) {
    tma_prefetch_map(&tmA);  // descriptor fetch overlaps the set-up (first operands used to land ~2 us in)
    tma_prefetch_map(&tmB);
    if (p.ksplit > 0) tma_prefetch_map(&tmA2);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], EPI_WARPS);  // one arrival per epilogue warp
    }
    for (int w = 0; w < EPI_WARPS; ++w)
      for (int a = 0; a < 4; ++a) mbar_init(&epi_bar[w][a], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 64 && p.tma_store) {
    tma_prefetch_map(&tmO);
    if (p.split_cout > 0) tma_prefetch_map(&tmO2);
    if (p.tma_epi) {
      if (p.flags & IVF_EP_ACCUM) tma_prefetch_map(&tmC);
      if (p.flags & IVF_EP_MASK) tma_prefetch_map(&tmM);
    }
  }
  if (warp >= 2) {
    // per-channel epilogue vectors of this N tile
    for (int i = threadIdx.x - 64; i < p.bn; i += 32 * EPI_WARPS) {
      int n = ntile * p.bn + i;
      bool ok = n < p.cout;
      s_scale[i] = (ok && (p.flags & IVF_EP_AFFINE)) ? scale[n] : 1.f;
      s_shift[i] = (ok && (p.flags & IVF_EP_AFFINE)) ? shift[n] : 0.f;
      s_mscale[i] = (ok && (p.flags & IVF_EP_MASK)) ? mask_scale[n] : 0.f;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = tmem_base_slot;
  ivf_pdl_wait();  // set-up above touched constants only
  if (warp == 1) TC_TRACE(2);  // set-up done

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int tcount = 0;
    TC_TRACE(0);
    for (int mt = blockIdx.x; mt < p.mtiles; mt += gridDim.x, ++tcount) {
      if (tcount < 4) TC_TRACE(8 + tcount * 8 + 0);  // producer starts tile
      const int m0 = mt * TILE_M;
      int ow0 = m0 % p.ow;
      int t = m0 / p.ow;
      int oh0 = t % p.oh;
      t /= p.oh;
      int od0 = t % p.od;
      int n0 = t / p.od;
      const int cw = ow0 * p.sw - p.pw, ch = oh0 * p.sh - p.ph, cd = od0 * p.sd - p.pd;
      int cc = 0, kw_i = 0, kh_i = 0, kd_i = 0, tap = 0;
      for (int it = 0; it < kiters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        const uint32_t a_dst = smem_base + stage * stage_bytes;
        const uint32_t b_dst = a_dst + p.a_stage_bytes;
        if (leader) {
          mbar_expect_tx(&full_bar[stage], p.tx_bytes);
          const bool second = p.ksplit > 0 && cc >= p.ksplit;
          tma_load_im2col_5d(a_dst, second ? &tmA2 : &tmA, &full_bar[stage], (second ? cc - p.ksplit : cc) * KCH,
                             cw, ch, cd, n0, (uint16_t)kw_i, (uint16_t)kh_i, (uint16_t)kd_i);
          tma_load_2d(b_dst, &tmB, &full_bar[stage], tap * p.cin_pad + cc * KCH, ntile * p.bn);
        }
        __syncwarp();
        if (++cc == p.cchunks) {
          cc = 0;
          ++tap;
          if (++kw_i == p.kw) {
            kw_i = 0;
            if (++kh_i == p.kh) {
              kh_i = 0;
              ++kd_i;
            }
          }
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(TILE_M, p.bn);
    const uint32_t desc_hi = smem_desc_hi(KT::sbo, KT::layout);
    const int tail_ksteps = (p.cin - (p.cchunks - 1) * KCH + 15) / 16;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int tcount = 0;
    for (int mt = blockIdx.x; mt < p.mtiles; mt += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);  // the epilogue has drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_acc + (uint32_t)(acc * p.acc_cols);
      int cc = 0;
      if (tcount < 4) TC_TRACE(8 + tcount * 8 + 1);  // MMA warp has the accumulator
      for (int it = 0; it < kiters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        if (tcount < 4 && it == 0) TC_TRACE(8 + tcount * 8 + 2);  // first operands landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_lo = smem_desc_lo(smem_base + stage * stage_bytes);
        const uint32_t b_lo = smem_desc_lo(smem_base + stage * stage_bytes + p.a_stage_bytes);
        if (leader) {
          if (cc + 1 < p.cchunks || tail_ksteps == KT::ksteps) {
#pragma unroll
            for (int k = 0; k < KT::ksteps; ++k)
              umma_bf16_lo(d_tmem, a_lo + 2u * k, b_lo + 2u * k, desc_hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
          } else {  // last channel chunk of a tap: only the K steps that hold real channels
            for (int k = 0; k < tail_ksteps; ++k)
              umma_bf16_lo(d_tmem, a_lo + 2u * k, b_lo + 2u * k, desc_hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        }
        __syncwarp();
        if (++cc == p.cchunks) cc = 0;
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (leader) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
      __syncwarp();
      if (tcount < 4) TC_TRACE(8 + tcount * 8 + 3);  // all MMAs issued
      ++tcount;
      if (++acc == p.acc_stages) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int cgrp = (warp - 2) >> 2;  // column group: chunks cgrp, cgrp + EPI_GROUPS, ...
    const int row = q * 32 + lane;
    EpilogueArgs ea;
    ea.cout = p.cout;
    ea.flags = p.flags;
    ea.acc_in = acc_in;
    ea.mask_y = mask_y;
    ea.out = out;
    // second destination: its pointer is shifted so that the produced-channel index addresses it directly
    EpilogueArgs ea2 = ea;
    if (p.split_cout > 0)
      ea2.out = (p.flags & IVF_EP_OUT_F32) ? (void*)(reinterpret_cast<float*>(out2) - p.split_cout)
                                           : (void*)(reinterpret_cast<__nv_bfloat16*>(out2) - p.split_cout);
    // the tile loop, instantiated per epilogue-flag combination (FL >= 0: compile-time flags, see epilogue_chunk16)
    auto run_epilogue = [&](auto fl_tag) {
      constexpr int FL = decltype(fl_tag)::value;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tcount = 0;
      int rot = 0;  // tile count modulo EPI_GROUPS
      int sbuf = 0;
      // Data-gradient operands by TMA (p.tma_epi).  A lane owns an output ROW, so its 64 bytes of consumer sum and
      // 32 bytes of ReLU mask per chunk are a row-scattered access: 32 different lines per warp instruction, six
      // instructions per chunk - the L1 tag stage, not the memory, bounded these epilogues (trace: ~2 us per chunk
      // against 0.5 us of a forward chunk).  Instead each warp fetches the [32 rows x 16 channels] boxes of its NEXT
      // chunks (also across tiles, before that tile's MMAs finish) into its own ring of staging buffers and reads
      // its row from shared memory (64-/32-byte swizzle: conflict-free per quarter warp).
      constexpr bool OPS = FL >= 0 && (FL & (IVF_EP_ACCUM | IVF_EP_MASK)) != 0;
      const bool te = OPS && p.tma_epi;
      // per warp 6 KB of staging: two buffers of [2 KB sum | 1 KB mask], or four 1 KB mask buffers when there is no
      // sum - the boxes of the chunk NB - 1 ahead are requested while the current one is consumed
      constexpr bool HAS_ACC = FL >= 0 && (FL & IVF_EP_ACCUM) != 0;
      constexpr int NB = HAS_ACC ? 2 : 4;
      constexpr uint32_t BUF_BYTES = HAS_ACC ? 3072u : 1024u, MASK_OFF = HAS_ACC ? 2048u : 0u;
      const uint32_t epi_stage = smem_base + p.epi_off + (uint32_t)(warp - 2) * 6144u;
      const uint8_t* epi_gen = smem_raw + (smem_base - smem_u32(smem_raw)) + p.epi_off + (size_t)(warp - 2) * 6144u;
      auto chunk_start = [&](int r) { return 16 * ((cgrp + EPI_GROUPS - r) % EPI_GROUPS); };
      // prefetch iterator over this warp's chunks in consumption order: (tile, rotation, column)
      int pmt = blockIdx.x, prot = 0, pc0 = chunk_start(0);
      auto pf_norm = [&]() {
        while (pmt < p.mtiles && pc0 >= p.bn) {
          pmt += gridDim.x;
          prot = prot + 1 == EPI_GROUPS ? 0 : prot + 1;
          pc0 = chunk_start(prot);
        }
      };
      int issued = 0, cons = 0;
      // boxes of the iterator's chunk -> staging buffer issued % NB (whole warp calls; one lane issues)
      auto issue_next = [&]() {
        if (pmt >= p.mtiles) return;
        const int b = issued % NB;
        if (lane == 0) {
          constexpr uint32_t bytes = (HAS_ACC ? 2048u : 0u) + ((FL & IVF_EP_MASK) ? 1024u : 0u);
          mbar_expect_tx(&epi_bar[warp - 2][b], bytes);
          const int nb_ = ntile * p.bn + pc0, r0_ = pmt * TILE_M + q * 32;
          if (HAS_ACC) tma_load_2d(epi_stage + b * BUF_BYTES, &tmC, &epi_bar[warp - 2][b], nb_, r0_);
          if (FL & IVF_EP_MASK) tma_load_2d(epi_stage + b * BUF_BYTES + MASK_OFF, &tmM, &epi_bar[warp - 2][b], nb_, r0_);
        }
        ++issued;
        pc0 += 16 * EPI_GROUPS;
        pf_norm();
      };
      if (te) {
        pf_norm();
        for (int i = 0; i < NB - 1; ++i) issue_next();
      }
      EpilogueArgs eaT = ea;
      eaT.cout = 1 << 30;  // TMA zero-fills past the tensor and the TMA store clips: every chunk is a full chunk
      for (int mt = blockIdx.x; mt < p.mtiles; mt += gridDim.x) {
        const int m = mt * TILE_M + row;
        const bool row_ok = m < p.M;
        const uint32_t taddr_row = tmem_acc + (uint32_t)(acc * p.acc_cols) + ((uint32_t)(q * 32) << 16);
        const size_t out_row = (size_t)m * p.out_ld + p.out_coff;
        const size_t out_row2 = (size_t)m * p.out2_ld + p.out2_coff;
        const size_t mask_row = (size_t)m * p.mask_ld + p.mask_coff;
        // the first chunk's global operands are fetched while the MMAs still run
        EpiPre cur;
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == 2 && tcount < 4) TC_TRACE(8 + tcount * 8 + 4);  // accumulator complete
        // chunk c of tile t goes to column group (c + t) % EPI_GROUPS: with a chunk count that is not a multiple
        // of the group count (4 chunks of a 64-channel layer over 3 groups) the extra chunk rotates over the
        // groups from tile to tile instead of making one group the bottleneck of every tile
        for (int c0 = 16 * ((cgrp + EPI_GROUPS - rot) % EPI_GROUPS); c0 < p.bn; c0 += 16 * EPI_GROUPS) {
          const int nb = ntile * p.bn + c0;  // first produced channel of this chunk
          // global operands of the chunk are requested before the TMEM load; the other warps of the scheduler
          // cover the latency
          if (te) {
            // the buffer the next request goes to was last read in the previous iteration of this loop
            __syncwarp();
            issue_next();
            const int b = cons % NB;
            mbar_wait(&epi_bar[warp - 2][b], (uint32_t)(cons / NB) & 1u);
            ++cons;
            const uint8_t* st = epi_gen + b * BUF_BYTES;
            if (HAS_ACC) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                cur.ac[j] = *reinterpret_cast<const float4*>(st + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
            }
            if (FL & IVF_EP_MASK) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                cur.mk[j] = *reinterpret_cast<const uint4*>(st + MASK_OFF + lane * 32 + ((j ^ ((lane >> 2) & 1)) << 4));
            }
          } else {
            epilogue_prefetch<FL>(ea, nb, out_row, mask_row, row_ok, cur);
          }
          uint32_t r[16];
          if (warp == 2 && tcount == 1 && c0 < 64 * EPI_GROUPS) TC_TRACE(40 + (c0 / (16 * EPI_GROUPS)) * 3 + 0);
          tmem_ld16(taddr_row + c0, r);
          if (warp == 2 && tcount == 1 && c0 < 64 * EPI_GROUPS) TC_TRACE(40 + (c0 / (16 * EPI_GROUPS)) * 3 + 1);
          if (p.tma_store) {
            if (nb < p.cout) {  // warp-uniform
              const uint32_t box = smem_u32(&stage_buf[warp - 2][sbuf][0]);
              const bool tr = warp == 2 && tcount == 1 && c0 < 16 * EPI_GROUPS;  // this warp's first chunk of tile 1
              if (tr) TC_TRACE(52);
              if (lane == 0) tma_store_wait_read<1>();  // the store that last read this box has finished reading
              __syncwarp();
              if (tr) TC_TRACE(53);
              const bool second = p.split_cout > 0 && nb >= p.split_cout;
              if (row_ok) {
                ea.stage_smem = ea2.stage_smem = eaT.stage_smem = box + (uint32_t)lane * 32u;
                epilogue_chunk16<false, FL>(te ? eaT : (second ? ea2 : ea), r, nb, s_scale + c0, s_shift + c0, s_mscale + c0,
                                            out_row, mask_row, cur);
              }
              if (tr) TC_TRACE(54);
              fence_proxy_async_smem();
              __syncwarp();
              if (tr) TC_TRACE(55);
              if (lane == 0) {
                tma_store_2d(second ? &tmO2 : &tmO, box, second ? nb - p.split_cout : nb, mt * TILE_M + q * 32);
                tma_store_commit();
              }
              if (tr) TC_TRACE(56);
              sbuf ^= 1;
            }
          } else if (row_ok && nb < p.cout) {
            if (p.split_cout > 0 && nb >= p.split_cout)
              epilogue_chunk16<false, FL>(ea2, r, nb, s_scale + c0, s_shift + c0, s_mscale + c0, out_row2, mask_row, cur);
            else
              epilogue_chunk16<false, FL>(ea, r, nb, s_scale + c0, s_shift + c0, s_mscale + c0, out_row, mask_row, cur);
          }
          if (warp == 2 && tcount == 1 && c0 < 64 * EPI_GROUPS) TC_TRACE(40 + (c0 / (16 * EPI_GROUPS)) * 3 + 2);
        }
        // this warp's TMEM reads are complete (tmem_ld16 waits): hand the accumulator back
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        if (warp == 2 && tcount < 4) TC_TRACE(8 + tcount * 8 + 5);  // epilogue of the tile done
        ++tcount;
        if (++rot == EPI_GROUPS) rot = 0;
        if (++acc == p.acc_stages) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    };
    switch (p.flags) {
      case IVF_EP_AFFINE | IVF_EP_RELU: run_epilogue(std::integral_constant<int, IVF_EP_AFFINE | IVF_EP_RELU>{}); break;
      case IVF_EP_MASK: run_epilogue(std::integral_constant<int, IVF_EP_MASK>{}); break;
      case IVF_EP_ACCUM | IVF_EP_MASK: run_epilogue(std::integral_constant<int, IVF_EP_ACCUM | IVF_EP_MASK>{}); break;
      case 0: run_epilogue(std::integral_constant<int, 0>{}); break;
        // the ConvLSTM's recurrent convolutions (fp32 pre-activations: x-conv + bias, h-conv accumulated onto them)
        case IVF_EP_ACCUM | IVF_EP_OUT_F32: run_epilogue(std::integral_constant<int, IVF_EP_ACCUM | IVF_EP_OUT_F32>{}); break;
        case IVF_EP_AFFINE | IVF_EP_OUT_F32: run_epilogue(std::integral_constant<int, IVF_EP_AFFINE | IVF_EP_OUT_F32>{}); break;
      default: run_epilogue(std::integral_constant<int, -1>{}); break;
    }
    if (p.tma_store && lane == 0) tma_store_wait_read<0>();  // the boxes are read until the stores complete
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) TC_TRACE(1);
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ---- bring-up probe: one im2col TMA tile -> global, de-swizzled -------------------------
template <int KCH>
__global__ void __launch_bounds__(128)
probe_im2col_kernel(const __grid_constant__ CUtensorMap tmA, int cw, int ch, int cd, int n0, int c0,
                    int kw_i, int kh_i, int kd_i, __nv_bfloat16* __restrict__ tile_out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tile = smem_raw + (smem_base - smem_u32(smem_raw));
  // poison so that rows the TMA does not write are visible in the dump
  for (int i = threadIdx.x; i < TILE_M * KCH; i += blockDim.x)
    reinterpret_cast<__nv_bfloat16*>(tile)[i] = __float2bfloat16_rn(-777.f);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, TILE_M * KCH * 2);
    tma_load_im2col_5d(smem_base, &tmA, &bar, c0, cw, ch, cd, n0, (uint16_t)kw_i, (uint16_t)kh_i,
                       (uint16_t)kd_i);
  }
  mbar_wait(&bar, 0);
  __syncthreads();
  constexpr int chunks = KCH * 2 / 16;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < TILE_M * chunks; i += blockDim.x) {
    int r = i / chunks, j = i % chunks;
    int sw = KCH == 64 ? (r & 7) : (KCH == 32 ? ((r >> 1) & 3) : ((r >> 2) & 1));
    const uint4* src = reinterpret_cast<const uint4*>(tile + (size_t)r * KCH * 2 + ((j ^ sw) * 16));
    reinterpret_cast<uint4*>(tile_out + (size_t)r * KCH)[j] = *src;
  }
}

// ---- host side ---------------------------------------------------------------------------

}  // namespace

int ivf_load_driver_entry_points() {
  static std::once_flag once;
  static int status = IVF_OK;
  std::call_once(once, []() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      status = IVF_ECUDA;
      return;
    }
    ivf_encode_im2col = (IvfEncodeIm2colFn)fn;
    fn = nullptr;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      status = IVF_ECUDA;
      return;
    }
    ivf_encode_tiled = (IvfEncodeTiledFn)fn;
  });
  if (status != IVF_OK) ivf_set_error("cudaGetDriverEntryPoint(cuTensorMapEncode*) failed");
  return status;
}

namespace {

CUtensorMapSwizzle swizzle_for(int kch) {
  return kch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                   : (kch == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

struct MapKeyA {
  const void* base;
  int n, id, ih, iw, cin, ld, coff, kd, kh, kw, sd, sh, sw, pd, ph, pw, od, oh, ow, kch;
};
struct MapKeyB {
  const void* base;
  int ktot, cout_pad, kch, bn;
};

template <typename K>
std::string key_bytes(char tag, const K& k) {
  std::string s(1, tag);
  s.append(reinterpret_cast<const char*>(&k), sizeof(K));
  return s;
}

int get_map_a(ivf_handle* h, const ivf_conv_desc* d, const void* in, int kch, CUtensorMap* out) {
  MapKeyA key;
  memset(&key, 0, sizeof(key));
  key.base = in;
  key.n = d->n; key.id = d->id; key.ih = d->ih; key.iw = d->iw; key.cin = d->cin;
  key.ld = d->in_ld; key.coff = d->in_coff;
  key.kd = d->kd; key.kh = d->kh; key.kw = d->kw; key.sd = d->sd; key.sh = d->sh; key.sw = d->sw;
  key.pd = d->pd; key.ph = d->ph; key.pw = d->pw; key.od = d->od; key.oh = d->oh; key.ow = d->ow;
  key.kch = kch;
  std::string kb = key_bytes('A', key);
  {
    std::lock_guard<std::mutex> g(h->mu);
    auto it = h->tmaps.find(kb);
    if (it != h->tmaps.end()) {
      *out = it->second;
      return IVF_OK;
    }
  }
  const char* base = reinterpret_cast<const char*>(in) + (size_t)d->in_coff * 2;
  IVF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "conv(bf16): input slice not 16-B aligned");
  cuuint64_t dims[5] = {(cuuint64_t)d->cin, (cuuint64_t)d->iw, (cuuint64_t)d->ih, (cuuint64_t)d->id,
                        (cuuint64_t)d->n};
  cuuint64_t pix = (cuuint64_t)d->in_ld * 2;
  cuuint64_t strides[4] = {pix, pix * d->iw, pix * d->iw * d->ih, pix * d->iw * d->ih * d->id};
  // bounding box of the window origin: lower = -front pad; upper = back pad - (k-1)
  int pbw = (d->ow - 1) * d->sw + d->kw - d->iw - d->pw;
  int pbh = (d->oh - 1) * d->sh + d->kh - d->ih - d->ph;
  int pbd = (d->od - 1) * d->sd + d->kd - d->id - d->pd;
  int lower[3] = {-d->pw, -d->ph, -d->pd};
  int upper[3] = {pbw - (d->kw - 1), pbh - (d->kh - 1), pbd - (d->kd - 1)};
  for (int i = 0; i < 3; ++i)
    IVF_REQUIRE(lower[i] >= -16 && lower[i] <= 15 && upper[i] >= -16 && upper[i] <= 15,
                "conv(bf16): padding/kernel outside the im2col corner range");
  cuuint32_t estr[5] = {1, (cuuint32_t)d->sw, (cuuint32_t)d->sh, (cuuint32_t)d->sd, 1};
  CUtensorMap m;
  CUresult r = ivf_encode_im2col(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)base, dims, strides,
                               lower, upper, (cuuint32_t)kch, (cuuint32_t)TILE_M, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kch),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    IVF_FAIL(IVF_ECUDA, "cuTensorMapEncodeIm2col failed (%d): dims c%d w%d h%d d%d n%d ld%d k%d%d%d",
             (int)r, d->cin, d->iw, d->ih, d->id, d->n, d->in_ld, d->kd, d->kh, d->kw);
  // CUTLASS (cute/atom/copy_traits_sm90_im2col.hpp) clears bit 21 of the second descriptor word
  // for tensors smaller than 128 KiB on drivers <= 13.1; same workaround here.
  int drv = 0;
  cudaDriverGetVersion(&drv);
  unsigned long long span = strides[3] * (cuuint64_t)d->n;
  if (drv <= 13010 && span < 131072ull) reinterpret_cast<uint64_t*>(&m)[1] &= ~(1ull << 21);
  {
    std::lock_guard<std::mutex> g(h->mu);
    h->tmaps[kb] = m;
  }
  *out = m;
  return IVF_OK;
}

int get_map_b(ivf_handle* h, const void* w, int ktot, int cout_pad, int kch, int bn, CUtensorMap* out) {
  MapKeyB key;
  memset(&key, 0, sizeof(key));
  key.base = w; key.ktot = ktot; key.cout_pad = cout_pad; key.kch = kch; key.bn = bn;
  std::string kb = key_bytes('B', key);
  {
    std::lock_guard<std::mutex> g(h->mu);
    auto it = h->tmaps.find(kb);
    if (it != h->tmaps.end()) {
      *out = it->second;
      return IVF_OK;
    }
  }
  IVF_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "conv(bf16): weights not 16-B aligned");
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)cout_pad};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)kch, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = ivf_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w, dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kch),
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    IVF_FAIL(IVF_ECUDA, "cuTensorMapEncodeTiled(weights) failed (%d): ktot %d cout_pad %d", (int)r,
             ktot, cout_pad);
  {
    std::lock_guard<std::mutex> g(h->mu);
    h->tmaps[kb] = m;
  }
  *out = m;
  return IVF_OK;
}

// Output tensor map for the TMA-store epilogue: [rows = pixels][channels of the destination slice], boxes of
// 32 rows x 16 channels (one epilogue warp's chunk); writes past either extent are clipped.
int get_map_out(ivf_handle* h, const void* base, int coff, int ld, int channels, long long rows, CUtensorMap* out) {
  struct {
    const void* base;
    int coff, ld, channels;
    long long rows;
  } key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.coff = coff; key.ld = ld; key.channels = channels; key.rows = rows;
  std::string kb = key_bytes('O', key);
  {
    std::lock_guard<std::mutex> g(h->mu);
    auto it = h->tmaps.find(kb);
    if (it != h->tmaps.end()) {
      *out = it->second;
      return IVF_OK;
    }
  }
  const char* p0 = reinterpret_cast<const char*>(base) + (size_t)coff * 2;
  IVF_REQUIRE((reinterpret_cast<uintptr_t>(p0) & 15) == 0 && ld % 8 == 0, "conv(bf16): output slice not 16-B aligned");
  cuuint64_t dims[2] = {(cuuint64_t)channels, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {16, 32};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = ivf_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)p0, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    IVF_FAIL(IVF_ECUDA, "cuTensorMapEncodeTiled(output) failed (%d): channels %d ld %d rows %lld", (int)r, channels,
             ld, rows);
  {
    std::lock_guard<std::mutex> g(h->mu);
    h->tmaps[kb] = m;
  }
  *out = m;
  return IVF_OK;
}

// [rows = pixels][channels] tensor read in boxes of 32 rows x 16 channels with the swizzle that matches the box row
// (64 bytes of fp32, 32 bytes of bf16): the TMA-loaded epilogue operands.  Boxes past either extent are zero filled.
int get_map_rows(ivf_handle* h, const void* base, int elem_bytes, long long coff, int ld, int channels, long long rows,
                 CUtensorMap* out) {
  struct {
    const void* base;
    int elem_bytes, ld, channels;
    long long coff, rows;
  } key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.elem_bytes = elem_bytes; key.coff = coff; key.ld = ld; key.channels = channels; key.rows = rows;
  std::string kb = key_bytes('R', key);
  {
    std::lock_guard<std::mutex> g(h->mu);
    auto it = h->tmaps.find(kb);
    if (it != h->tmaps.end()) {
      *out = it->second;
      return IVF_OK;
    }
  }
  const char* p0 = reinterpret_cast<const char*>(base) + (size_t)coff * elem_bytes;
  cuuint64_t dims[2] = {(cuuint64_t)channels, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {16, 32};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = ivf_encode_tiled(&m, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                              2, (void*)p0, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              elem_bytes == 4 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    IVF_FAIL(IVF_ECUDA, "cuTensorMapEncodeTiled(epilogue operand) failed (%d): channels %d ld %d rows %lld", (int)r,
             channels, ld, rows);
  {
    std::lock_guard<std::mutex> g(h->mu);
    h->tmaps[kb] = m;
  }
  *out = m;
  return IVF_OK;
}

int check_bf16_desc(const ivf_conv_desc* d) {
  IVF_REQUIRE(!d->transposed,
              "conv(bf16): transposed gather is fp32-only; present strided layers space-to-depth");
  IVF_REQUIRE(d->cin % 8 == 0 && d->in_ld % 8 == 0 && d->in_coff % 8 == 0,
              "conv(bf16): cin/in_ld/in_coff must be multiples of 8 (got %d/%d/%d)", d->cin,
              d->in_ld, d->in_coff);
  IVF_REQUIRE(d->out_ld % 8 == 0 && d->out_coff % 8 == 0,
              "conv(bf16): out_ld/out_coff must be multiples of 8");
  if (d->flags & IVF_EP_MASK)
    IVF_REQUIRE(d->mask_ld % 8 == 0 && d->mask_coff % 8 == 0,
                "conv(bf16): mask_ld/mask_coff must be multiples of 8");
  IVF_REQUIRE(d->sd <= 8 && d->sh <= 8 && d->sw <= 8, "conv(bf16): stride > 8");
  return IVF_OK;
}

template <int KCH>
int launch_tc(ivf_handle* h, const ivf_conv_desc* d, const TcParams& p, const CUtensorMap& ma,
              const CUtensorMap& mb, const CUtensorMap& ma2, const CUtensorMap& mo, const CUtensorMap& mo2,
              const CUtensorMap& mc, const CUtensorMap& mm, int ntiles, const float* scale, const float* shift,
              const float* acc_in, const void* mask_y, const float* mask_scale, void* out, void* out2,
              cudaStream_t st) {
  const int max_smem = 196 * 1024;  // 227 KB minus the static part (barriers, epilogue vectors, TMA-store boxes)
  const int slot = KCH == 64 ? 0 : (KCH == 32 ? 1 : 2);
  if (!h->tc_attr_set[slot]) {
    IVF_CUDA(cudaFuncSetAttribute(conv_tc_kernel<KCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  max_smem));
    h->tc_attr_set[slot] = true;
  }
  size_t smem = (size_t)p.stages * (p.a_stage_bytes + p.b_stage_bytes) + 1024;
  if (p.tma_epi) smem = (size_t)p.epi_off + EPI_STAGING_BYTES + 1024;
  dim3 grid(p.grid_x, ntiles);
  IVF_CUDA(ivf_launch(conv_tc_kernel<KCH>, grid, dim3(NUM_THREADS), smem, st, 1, ma, mb, ma2, mo, mo2, mc, mm, p, scale, shift,
                      acc_in, (const __nv_bfloat16*)mask_y, mask_scale, out, out2));
  IVF_LAUNCHED(h);
  return IVF_OK;
}

}  // namespace

// Channels per K stage: one 128-byte swizzled row (64 channels) unless the operand is narrower.  A wider
// operand whose channel count is not a multiple of 64 (96, 112, 144, 160 ...) still moves 64-channel boxes
// (TMA zero-fills past the last channel) and the last chunk issues only its real K steps: 128-byte TMA rows
// and half as many pipeline stages as the former 32-channel chunking of those layers.
extern "C" int ivf_conv_bf16_kchunk(int cin) { return cin <= 16 ? 16 : (cin <= 32 ? 32 : 64); }
// K pitch of one filter tap in the packed weights
extern "C" int ivf_conv_bf16_cin_pad(int cin) {
  if (cin <= 16) return 16;
  if (cin <= 32) return 32;
  return (cin + 15) / 16 * 16;
}
extern "C" int ivf_conv_bf16_ntile(int cout) {
  int tiles = (cout + 255) / 256;
  int per = (cout + tiles - 1) / tiles;
  return (per + 15) / 16 * 16;
}
extern "C" int ivf_conv_bf16_cout_pad(int cout) {
  int tiles = (cout + 255) / 256;
  return ivf_conv_bf16_ntile(cout) * tiles;
}

int ivf_conv3d_tc_launch(ivf_handle* h, const ivf_conv_desc* d, const void* in, const void* w,
                         const float* scale, const float* shift, const float* acc_in,
                         const void* mask_y, const float* mask_scale, void* out, cudaStream_t st,
                         const ivf_conv_split* sp, const void* in2, void* out2) {
  int rc = check_bf16_desc(d);
  if (rc) return rc;
  rc = ivf_load_driver_entry_points();
  if (rc) return rc;
  const int split_cin = sp ? sp->split_cin : 0, split_cout = sp ? sp->split_cout : 0;
  const int kch = ivf_conv_bf16_kchunk(split_cin > 0 ? 64 : d->cin);
  // two sources: the first one's channels are padded up to whole K stages (TMA zero-fills past its extent,
  // the packed weights hold zeros there), the second follows
  const int ksplit = split_cin > 0 ? (split_cin + kch - 1) / kch : 0;
  const int cin_k = split_cin > 0 ? ksplit * kch + (d->cin - split_cin) : d->cin;  // K extent per tap
  const int cin_pad = split_cin > 0 ? (cin_k + 15) / 16 * 16 : ivf_conv_bf16_cin_pad(d->cin);
  int bn = ivf_conv_bf16_ntile(d->cout);
  const int cout_pad = ivf_conv_bf16_cout_pad(d->cout);
  int ntiles = cout_pad / bn;
  const int ntaps = d->kd * d->kh * d->kw;
  long long M = (long long)d->n * d->od * d->oh * d->ow;
  IVF_REQUIRE(M < (1ll << 31), "conv(bf16): too many output pixels");
  // Few 128-pixel tiles (the 14x14 and 7x7 stages): split the output channels over more CTAs so the
  // grid covers the SMs; the activation tile is re-read per N tile from L2, which is cheap there.
  // (Rows of the weight box beyond cout_pad are TMA zero fill; the epilogue stores only real channels.)
  const int mtiles = ivf_cdiv(M, TILE_M);
  // IVF_TC_SPREAD = the share of the SMs (per cent) one launch tries to cover this way (default 100)
  static const int spread = [] { const char* e = getenv("IVF_TC_SPREAD"); return e ? atoi(e) : 100; }();
  const int sm_target = std::max(1, h->sm_count * spread / 100);
  if (mtiles * ntiles < sm_target) {
    int want = std::max(1, sm_target / mtiles);  // one CTA per SM: stay within one wave
    int bn2 = (ivf_cdiv(d->cout, want) + 15) / 16 * 16;
    if (bn2 < 32) bn2 = 32;
    if (bn2 < bn) {
      bn = bn2;
      ntiles = ivf_cdiv(d->cout, bn);
    }
  }

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M;
  p.od = d->od; p.oh = d->oh; p.ow = d->ow;
  p.cout = d->cout;
  p.bn = bn;
  p.kh = d->kh; p.kw = d->kw;
  p.ntaps = ntaps;
  p.cchunks = (cin_k + kch - 1) / kch;
  p.cin = cin_k;
  p.cin_pad = cin_pad;
  p.ksplit = ksplit;
  p.split_cout = split_cout;
  p.out2_ld = sp ? sp->out2_ld : 0;
  p.out2_coff = sp ? sp->out2_coff : 0;
  p.sd = d->sd; p.sh = d->sh; p.sw = d->sw;
  p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
  p.out_ld = d->out_ld; p.out_coff = d->out_coff;
  p.mask_ld = d->mask_ld; p.mask_coff = d->mask_coff;
  p.flags = d->flags;
  p.a_stage_bytes = TILE_M * kch * 2;
  uint32_t bbytes = (uint32_t)bn * kch * 2;
  p.b_stage_bytes = (bbytes + 1023u) & ~1023u;
  p.tx_bytes = p.a_stage_bytes + bbytes;
  const uint32_t stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  int kiters = ntaps * p.cchunks;
  // TMEM and residency: an accumulator takes the next power of two >= bn columns.  Up to 128 columns two
  // CTAs share an SM with two accumulators each (4 x 128 = all 512 columns); wider tiles get one CTA per SM
  // with two accumulators and the whole shared memory as pipeline.  A CTA walks ceil(mtiles / capacity) tiles.
  int cols = 32;
  while (cols < bn) cols <<= 1;
  static const int acc_env = [] { const char* e = getenv("IVF_TC_ACC"); return e ? atoi(e) : 2; }();
  static const int cta_env = [] { const char* e = getenv("IVF_TC_CTAS"); return e ? atoi(e) : 0; }();
  int ctas_per_sm = 1;  // 18 warps and the whole shared memory per CTA
  if (cta_env > 0) ctas_per_sm = cta_env;
  int acc_stages = acc_env >= 2 ? 2 : 1;
  while (acc_stages > 1 && acc_stages * cols * ctas_per_sm > 512) --acc_stages;
  while (ctas_per_sm > 1 && acc_stages * cols * ctas_per_sm > 512) --ctas_per_sm;
  int capacity = ctas_per_sm * h->sm_count / ntiles;
  if (capacity < 1) capacity = 1;
  int grid_x = mtiles;
  if (mtiles > capacity) {
    const int per_cta = ivf_cdiv(mtiles, capacity);
    grid_x = ivf_cdiv(mtiles, per_cta);
  } else {
    acc_stages = 1;  // one tile per CTA: nothing to overlap
  }
  p.mtiles = mtiles;
  p.acc_stages = acc_stages;
  p.acc_cols = cols;
  p.tmem_cols = acc_stages * cols;
  // epilogue operands by TMA: bf16 results through the TMA store, operands 16-byte aligned with 16-byte row pitches
  static const bool tma_store_env = [] { const char* e = getenv("IVF_TC_TMA_STORE"); return !e || atoi(e) != 0; }();
  static const bool tma_epi_env = [] { const char* e = getenv("IVF_TC_TMA_EPI"); return !e || atoi(e) != 0; }();
  const bool want_ops = (d->flags == IVF_EP_MASK || d->flags == (IVF_EP_MASK | IVF_EP_ACCUM));
  bool te = tma_store_env && tma_epi_env && want_ops && ctas_per_sm == 1;
  if (te && (d->flags & IVF_EP_ACCUM))
    te = (reinterpret_cast<uintptr_t>(acc_in + d->out_coff) & 15) == 0 && d->out_ld % 4 == 0;
  if (te && (d->flags & IVF_EP_MASK))
    te = (reinterpret_cast<uintptr_t>(reinterpret_cast<const __nv_bfloat16*>(mask_y) + d->mask_coff) & 15) == 0 &&
         d->mask_ld % 8 == 0;
  const uint32_t pipe_kb = te ? 190u - EPI_STAGING_BYTES / 1024u : 190u;
  const uint32_t budget = (ctas_per_sm >= 2 && (long long)grid_x * ntiles > h->sm_count ? 100u : pipe_kb) * 1024u;
  int stages = (int)(budget / stage_bytes);
  if (stages < 2) stages = 2;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  while ((size_t)stages * stage_bytes > pipe_kb * 1024u && stages > 1) --stages;
  if ((size_t)stages * stage_bytes > pipe_kb * 1024u) te = false;  // a single stage does not leave the room
  if (stages > kiters * ivf_cdiv(mtiles, grid_x)) stages = kiters * ivf_cdiv(mtiles, grid_x);
  if (stages < 1) stages = 1;
  p.stages = stages;
  p.grid_x = grid_x;
  p.tma_epi = te ? 1 : 0;
  p.epi_off = (uint32_t)(((size_t)stages * stage_bytes + 1023) & ~(size_t)1023);
  static const bool trace_env = [] { const char* e = getenv("IVF_TC_TRACE"); return e && atoi(e) != 0; }();
  p.trace = trace_env ? reinterpret_cast<long long*>(h->scratch) : nullptr;

  CUtensorMap ma, mb, ma2;
  if (split_cin > 0) {
    ivf_conv_desc d1 = *d, d2 = *d;
    d1.cin = split_cin;
    d2.cin = d->cin - split_cin;
    d2.in_ld = sp->in2_ld;
    d2.in_coff = sp->in2_coff;
    rc = get_map_a(h, &d1, in, kch, &ma);
    if (rc) return rc;
    rc = get_map_a(h, &d2, in2, kch, &ma2);
    if (rc) return rc;
  } else {
    rc = get_map_a(h, d, in, kch, &ma);
    if (rc) return rc;
    ma2 = ma;
  }
  rc = get_map_b(h, w, ntaps * cin_pad, cout_pad, kch, bn, &mb);
  if (rc) return rc;
  // bf16 results go out through TMA stores (a thread holds one output ROW, and row-scattered 16-byte stores
  // measured ~0.2 us per warp instruction - the epilogue, not the MMAs, bounded the 1x1x1 layers)
  CUtensorMap mo = ma, mo2 = ma, mc = ma, mm = ma;
  p.tma_store = (tma_store_env && !(d->flags & IVF_EP_OUT_F32)) ? 1 : 0;
  if (p.tma_store) {
    rc = get_map_out(h, out, d->out_coff, d->out_ld, split_cout > 0 ? split_cout : d->cout, M, &mo);
    if (rc) return rc;
    mo2 = mo;
    if (split_cout > 0) {
      rc = get_map_out(h, out2, sp->out2_coff, sp->out2_ld, d->cout - split_cout, M, &mo2);
      if (rc) return rc;
    }
  }
  if (p.tma_epi) {
    if (d->flags & IVF_EP_ACCUM) {
      rc = get_map_rows(h, acc_in, 4, d->out_coff, d->out_ld, d->cout, M, &mc);
      if (rc) return rc;
    }
    if (d->flags & IVF_EP_MASK) {
      rc = get_map_rows(h, mask_y, 2, d->mask_coff, d->mask_ld, d->cout, M, &mm);
      if (rc) return rc;
    }
  }
  if (kch == 64)
    return launch_tc<64>(h, d, p, ma, mb, ma2, mo, mo2, mc, mm, ntiles, scale, shift, acc_in, mask_y, mask_scale, out, out2, st);
  if (kch == 32)
    return launch_tc<32>(h, d, p, ma, mb, ma2, mo, mo2, mc, mm, ntiles, scale, shift, acc_in, mask_y, mask_scale, out, out2, st);
  return launch_tc<16>(h, d, p, ma, mb, ma2, mo, mo2, mc, mm, ntiles, scale, shift, acc_in, mask_y, mask_scale, out, out2, st);
}

extern "C" int ivf_probe_im2col(ivf_handle* h, const ivf_conv_desc* d, const void* in, int m0,
                                int tap, int c0, void* tile_out, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && d && in && tile_out, "ivf_probe_im2col: null argument");
  int rc = check_bf16_desc(d);
  if (rc) return rc;
  rc = ivf_load_driver_entry_points();
  if (rc) return rc;
  const int kch = ivf_conv_bf16_kchunk(d->cin);
  CUtensorMap ma;
  rc = get_map_a(h, d, in, kch, &ma);
  if (rc) return rc;
  int ow0 = m0 % d->ow;
  int t = m0 / d->ow;
  int oh0 = t % d->oh;
  t /= d->oh;
  int od0 = t % d->od;
  int n0 = t / d->od;
  int cw = ow0 * d->sw - d->pw, ch = oh0 * d->sh - d->ph, cd = od0 * d->sd - d->pd;
  int kw_i = tap % d->kw, t2 = tap / d->kw, kh_i = t2 % d->kh, kd_i = t2 / d->kh;
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = (size_t)TILE_M * kch * 2 + 1024;
  if (kch == 64)
    probe_im2col_kernel<64><<<1, 128, smem, st>>>(ma, cw, ch, cd, n0, c0, kw_i, kh_i, kd_i,
                                                  (__nv_bfloat16*)tile_out);
  else if (kch == 32)
    probe_im2col_kernel<32><<<1, 128, smem, st>>>(ma, cw, ch, cd, n0, c0, kw_i, kh_i, kd_i,
                                                  (__nv_bfloat16*)tile_out);
  else
    probe_im2col_kernel<16><<<1, 128, smem, st>>>(ma, cw, ch, cd, n0, c0, kw_i, kh_i, kd_i,
                                                  (__nv_bfloat16*)tile_out);
  IVF_LAUNCHED(h);
  return IVF_OK;
}
