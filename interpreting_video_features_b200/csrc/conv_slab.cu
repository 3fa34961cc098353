// Stride-1 'same' 3-D convolution as a tcgen05 implicit GEMM whose activation operand is a
// shared-memory resident HALO SLAB instead of a per-tap im2col tile.
//
// Why: the im2col kernel (conv_tc.cu) re-fetches the 128-pixel activation tile from L2 once per
// filter tap (27x for 3x3x3, 64x for the space-to-depth stem) and the weight tile once per 128
// pixels; on the 56x56 / 28x28 / 112x112 stages that makes it L2->SM bandwidth bound (measured: 8.8
// TB/s of L2 traffic at 41 % tensor utilisation on Conv3d_2c).  Here one CTA tile is TH output rows
// of one (clip, depth) slice.  For every depth tap kd and 64-channel chunk ONE tiled TMA box
// {64 ch, W+kw-1, TH+kh-1} lands the zero-padded input rows in shared memory ('same' padding = TMA
// out-of-bounds zero fill, pt/models/I3D_doubled.py:96-106).  Pixels of the slab are 128-byte rows
// (SWIZZLE_128B, 8-row atoms, SBO 1024), so in "padded-width" pixel numbering v = r*(W+kw-1)+c the
// operand of tap (kh,kw) is the SAME slab read from a start address shifted by (kh*(W+kw-1)+kw) rows:
// kh*kw MMAs groups reuse one slab, and the kw-1 junk columns per row are computed and never stored.
// A tile holds MT (1..4) 128-row accumulators so one weight tile feeds MT MMAs.
//
// kw-merge (kwm > 1): an MMA's cost is bounded by fetching its 128 x 16 activation sub-tile from shared
// memory (32 cycles) unless N >= 128, and the layers with few output channels (stem 64, data gradients
// 24..96) would run at N/64 of the tensor rate.  For those the kwm filter taps of a row are stacked along
// N: P_g[u] = sum_k A[u + kh*wp][k] * W[kh, kw=g][k], one MMA of N = kwm*bn for all g, accumulated over
// (kd, kh, channel chunk) in TMEM, and the epilogue forms out[v] = sum_g P_g[v + g]: a row shift of g
// between column blocks, done with warp shuffles plus a shared-memory exchange of the first kwm-1 rows of
// the next 32-row segment (published by every warp for its columns, one block barrier per tile).
//
// depth stacking (ds = 2): the other way to a wide N for the layers with few output channels, without the shift-
// and-add epilogue of the kw-merge.  The slab of input depth z feeds output depth d through depth tap kd = z - d + pd,
// so a tile that owns TWO consecutive output depths d0, d0+1 reads each of its kd+1 slabs once and multiplies it by
// the weight rows [W[kd(d0)] ; W[kd(d0+1)]] stacked along N: one MMA of N = 2*bn writes both depths' accumulators,
// which sit side by side in TMEM (the first / last slab reach only one of the two depths: N = bn into that half).
// Same FLOPs, (kd+1)/(2 kd) of the MMA instructions and slab loads - and from N = 128 up an MMA runs at the math
// rate instead of the issue rate (tools/mma_bench.cu).  The slab of offset 0 reaches both depths and is issued first
// (it needs pd >= 1), so one accumulate flag serves the whole N.
//
// NCTA = 2 (cta_group::2): a CTA pair on one TPC works on two row-adjacent tiles in lockstep.  Each SM loads
// its own slab and HALF of the weight rows; the leader CTA issues M = 256 MMAs that read both SMs' shared
// memory and write each SM's TMEM.  Per SM an MMA then fetches 4 KB + 16N bytes instead of 4 KB + 32N (the
// operand fetch is what bounds this kernel), and the weights cross L2->SM once per pair.  TMA completions of
// both CTAs are counted on the leader's "full" barriers; MMA commits are multicast to both CTAs' "empty" and
// "accumulator full" barriers; both CTAs' epilogues arrive on the leader's "accumulator empty" barrier.
//
// Persistent, warp specialised: warp 0 slab TMA producer, warp 1 MMA issuer, warp 2 weight TMA
// producer (+ TMEM alloc), warps 3-6 epilogue; TMEM accumulators double buffered so the epilogue of
// tile i overlaps the MMAs of tile i+1.
#include "conv_common.cuh"

#include <type_traits>

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace {

using namespace ivf_tc;

// 3 role warps + SLAB_EPI_GROUPS epilogue warps per TMEM lane quarter, each group a contiguous share of the tile's
// 16-column chunks.  The epilogue is a dependent chain per warp (TMEM load, shuffles of the kw-merge, a few ALU
// ops, stores); it needs several warps per scheduler to run at speed (measured on the 1x1x1 kernel: one warp
// per scheduler ran at ~0.2 IPC), and the kw-merged plans are bounded by it.
#ifndef SLAB_EPI_GROUPS
#define SLAB_EPI_GROUPS 3
#endif
constexpr int EPI_THREADS = 128 * SLAB_EPI_GROUPS;
// Role warps: 0 slab (A) producer, 1 and 3 MMA issuers, 2 weight (B) producer + TMEM allocation.  ONE issuing
// thread sustains one tcgen05.mma per ~70 cycles whatever N (tools/mma_bench.cu on the B200: 75 cycles per
// M128 x N<=128 x K16 from one warp, 40-64 from two, i.e. the math rate N/2 from N = 128 up); the layers with few
// output channels (stem N = 64, data gradients N = 32..96) were bound by exactly that, so tiles with two or
// more accumulators are issued by two warps, even / odd accumulators each.
constexpr int SLAB_ROLE_WARPS = 4;
constexpr int SLAB_THREADS = 32 * SLAB_ROLE_WARPS + EPI_THREADS;
constexpr int MAX_A_STAGES = 4;
constexpr int MAX_B_STAGES = 8;
constexpr uint32_t SLAB_SMEM_BUDGET = 216u * 1024u;  // + 1 KB alignment slack + ~9.5 KB static = 227 KB
constexpr int SLAB_MAX_COUT = 1024;  // per-channel epilogue vectors live in shared memory
constexpr uint32_t SLAB_EST_BYTES = (uint32_t)(EPI_THREADS / 32) * 32u * 16u * 4u;  // transposition staging, 2 KB per warp

struct SlabParams {
  int n, dd, hh, ww;  // spatial extent (output == gathered tensor: stride 1, 'same')
  int kd, kh, kw, pd, ph, pw;
  int wp;      // padded row length W + kw - 1
  int th;      // output rows per tile
  int htiles;  // ceil(H / th)
  int mt;      // 128-row accumulators per tile = ceil(th*wp / 128)
  int cin, cchunks, cin_pad;
  int cout, bn, ntiles, slot;  // slot = TMEM columns of one 128-row accumulator (>= kwm*bn)
  int kwm;                     // kw taps merged into the MMA's N (1 = none)
  int ds;                      // output DEPTHS per tile stacked along the MMA's N (1 = none, 2), see below
  int dgroups;                 // dd / ds
  int num_tiles;
  int out_ld, out_coff, mask_ld, mask_coff, flags;
  int a_stages, b_stages, tmem_cols;
  int acc_stages;  // 2: epilogue of tile i overlaps the MMAs of tile i+1; 1: all TMEM columns for one tile
  int kch;
  int ncta;  // 1, or 2 = CTA pairs (cta_group::2)
  uint32_t xch_off;  // kw-merge: byte offset (from the aligned dynamic shared memory base) of the boundary-row
  int xch_seg;       // exchange [2 tile parities][4*mt segments][xch_seg floats], xch_seg = (kwm-1)^2 * bn
  uint32_t est_off;  // != 0: byte offset of the epilogue's transposition staging (SLAB_EST_BYTES, one 32 x 16 fp32
                     // block per epilogue warp): the warps then touch global memory row-contiguously, see the epilogue
  int diag;  // IVF_SLAB_DIAG (timing experiments, results are garbage): bit 0 / 1 = after the ring has filled
             // once, the slab / weight producer signals "full" without loading; bit 2 = the epilogue reads TMEM but
             // neither loads its global operands nor stores anything
  uint32_t a_stage_bytes, b_stage_bytes, b_tap_bytes, a_tx, b_tx;  // a B stage holds the kw taps of one row
  // IVF_EP_LSTM: the recurrent step's state buffers (see EpilogueArgs)
  const float* lstm_c_prev;
  float* lstm_c_next;
  __nv_bfloat16* lstm_h_next;
};

struct TileCoord {
  int nt, h0, dz, nn;
};
// hpairs = row tiles per slice in units of one work item: htiles (single CTA) or ceil(htiles/2) (CTA pair:
// rank r takes row tile 2*j + r; a tile past the last row is all padding and stores nothing)
__device__ __forceinline__ TileCoord decode_tile(const SlabParams& p, int tile, int ncta, int rank) {
  TileCoord t;
  t.nt = tile % p.ntiles;
  int r = tile / p.ntiles;
  const int hp = ncta == 2 ? (p.htiles + 1) / 2 : p.htiles;
  t.h0 = ((r % hp) * ncta + rank) * p.th;
  r /= hp;
  t.dz = (r % p.dgroups) * p.ds;  // first output depth of the tile
  t.nn = r / p.dgroups;
  return t;
}

// Slabs of a tile in issue order: offset of the idx-th slab's input depth from the tile's first output depth.
// ds == 1: ascending depth taps (kd = idx).  ds == 2: offset 0 first (it reaches both depths), then the rest.
__device__ __forceinline__ int slab_offset(const SlabParams& p, int idx) {
  if (p.ds == 1) return idx - p.pd;
  if (idx == 0) return 0;
  return idx <= p.pd ? -idx : idx - p.pd;
}

// KCH = channels per slab row: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B, for
// operands of <= 32 channels: the space-to-depth stem's 24-of-32 and the 16/32-channel bottlenecks)
// GROUPED: one launch serves TWO independent convolutions (the two 3x3x3 branches of an Inception module, forward
// or data gradient): CTAs [0, split) work on the first, the rest on the second, each walking its own tiles.  At 8
// clips those branch convolutions are 10-25 us launches that are set-up and pipeline latency end to end; as one launch
// they share it and run side by side on disjoint SMs (measured upper bound with the thin branch removed: -5 % step).
struct SlabOperands {
  const float* scale;
  const float* shift;
  const float* acc_in;
  const __nv_bfloat16* mask_y;
  const float* mask_scale;
  void* out;
};

template <int KCH, int NCTA, bool LSTM = false, bool GROUPED = false>
__global__ void __launch_bounds__(SLAB_THREADS, 1)
conv_slab_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ SlabParams p1, const SlabOperands o1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ SlabParams p2, const SlabOperands o2, const int split) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ SlabParams s_params;  // GROUPED: this CTA's problem
  const bool second = GROUPED && blockIdx.x >= (unsigned)split;
  if (GROUPED) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(second ? &p2 : &p1);
    for (int i = threadIdx.x; i < (int)(sizeof(SlabParams) / 4); i += blockDim.x)
      reinterpret_cast<uint32_t*>(&s_params)[i] = src[i];
    __syncthreads();
  }
  const SlabParams& p = GROUPED ? s_params : p1;
  const CUtensorMap* const tmA = second ? &tmA2 : &tmA1;
  const CUtensorMap* const tmB = second ? &tmB2 : &tmB1;
  const float* __restrict__ scale = second ? o2.scale : o1.scale;
  const float* __restrict__ shift = second ? o2.shift : o1.shift;
  const float* __restrict__ acc_in = second ? o2.acc_in : o1.acc_in;
  const __nv_bfloat16* __restrict__ mask_y = second ? o2.mask_y : o1.mask_y;
  const float* __restrict__ mask_scale = second ? o2.mask_scale : o1.mask_scale;
  void* __restrict__ out = second ? o2.out : o1.out;
  __shared__ __align__(8) uint64_t a_full[MAX_A_STAGES], a_empty[MAX_A_STAGES];
  __shared__ __align__(8) uint64_t b_full[MAX_B_STAGES], b_empty[MAX_B_STAGES];
  __shared__ __align__(8) uint64_t t_full[2], t_empty[2];
  __shared__ uint32_t tmem_base_slot;
  // per-channel epilogue vectors: s_scale = BN scale (AFFINE) or the producer's BN' mask scale (MASK; the two
  // are never combined on this path, the host checks), s_shift = BN shift
  __shared__ __align__(16) float s_scale[SLAB_MAX_COUT], s_shift[SLAB_MAX_COUT];

  constexpr uint32_t ROWB = KCH * 2;             // bytes per slab pixel / weight row
  constexpr uint32_t ROW16 = ROWB / 16;          // the same in descriptor (16-byte) units
  constexpr uint32_t LAYOUT = KCH == 64 ? 2u : 4u;
  constexpr int KSTEPS = KCH / 16;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = NCTA == 2 ? (int)cluster_ctarank() : 0;  // 0 = leader of the pair
  // GROUPED (single CTAs only): this problem's CTAs are [0, split) or [split, gridDim.x)
  const int item0 = NCTA == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x - (second ? split : 0);
  const int item_step = NCTA == 2 ? (int)(gridDim.x >> 1)
                                  : (GROUPED ? (second ? (int)gridDim.x - split : split) : (int)gridDim.x);
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + (uint32_t)p.a_stages * p.a_stage_bytes;

  ivf_pdl_trigger();  // the next kernel of the stream may begin its own set-up
  if (threadIdx.x == 0) {
    tma_prefetch_map(tmA);  // descriptor fetch overlaps the set-up
    tma_prefetch_map(tmB);
    const uint32_t nissue = p.mt >= 2 ? 2u : 1u;  // MMA-issuing warps: each commits once per slot / tile
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], nissue);
    }
    for (int s = 0; s < p.b_stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], nissue);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&t_full[s], nissue);
      mbar_init(&t_empty[s], 4 * SLAB_EPI_GROUPS * NCTA);  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    __syncwarp();
    if constexpr (NCTA == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(&tmem_base_slot)),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(&tmem_base_slot)),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (warp >= SLAB_ROLE_WARPS) {
    for (int i = threadIdx.x - 32 * SLAB_ROLE_WARPS; i < SLAB_MAX_COUT; i += EPI_THREADS) {
      bool ok = i < p.cout;
      float sc = 1.f;
      if (ok && (p.flags & IVF_EP_AFFINE)) sc = scale[i];
      if (ok && (p.flags & IVF_EP_MASK)) sc = mask_scale[i];
      s_scale[i] = sc;
      s_shift[i] = (ok && (p.flags & IVF_EP_AFFINE)) ? shift[i] : 0.f;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();  // the peer's barriers exist before anything signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = tmem_base_slot;
  ivf_pdl_wait();  // everything above touched constants only; activations / gradients of earlier kernels from here

  if (warp == 0) {
    // ===================== slab (A) TMA producer =====================
    {
      const bool leader = elect_one();
      int a_loads = 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = item0; tile < p.num_tiles; tile += item_step) {
        const TileCoord t = decode_tile(p, tile, NCTA, rank);
        for (int si = 0; si < p.kd + p.ds - 1; ++si) {
          const int zd = t.dz + slab_offset(p, si);
          if (zd < 0 || zd >= p.dd) continue;  // an all-padding depth tap contributes nothing
          for (int cc = 0; cc < p.cchunks; ++cc) {
            mbar_wait(&a_empty[stage], phase ^ 1u);
            if ((p.diag & 1) && a_loads >= p.a_stages) {
              if (leader && rank == 0) mbar_arrive(&a_full[stage]);
            } else if (leader) {
              if constexpr (NCTA == 2) {
                if (rank == 0) mbar_expect_tx(&a_full[stage], 2u * p.a_tx);  // both CTAs' slabs
                tma_load_5d_pair(a_base + stage * p.a_stage_bytes, tmA, &a_full[stage], cc * KCH, -p.pw,
                                 t.h0 - p.ph, zd, t.nn);
              } else {
                mbar_expect_tx(&a_full[stage], p.a_tx);
                tma_load_5d(a_base + stage * p.a_stage_bytes, tmA, &a_full[stage], cc * KCH, -p.pw,
                            t.h0 - p.ph, zd, t.nn);
              }
            }
            __syncwarp();
            ++a_loads;
            if (++stage == p.a_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== weight (B) TMA producer =====================
    {
      const bool leader = elect_one();
      int b_loads = 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = item0; tile < p.num_tiles; tile += item_step) {
        const TileCoord t = decode_tile(p, tile, NCTA, rank);
        for (int si = 0; si < p.kd + p.ds - 1; ++si) {
          const int off = slab_offset(p, si);
          const int zd = t.dz + off;
          if (zd < 0 || zd >= p.dd) continue;
          // output depths (of the tile's ds) this slab reaches: s_lo .. s_lo + ns - 1, through depth tap off - s + pd
          const int s_lo = max(0, off + p.pd - (p.kd - 1)), ns = min(p.ds - 1, off + p.pd) - s_lo + 1;
          const int kd_i = off - s_lo + p.pd;
          for (int cc = 0; cc < p.cchunks; ++cc) {
            for (int kh_i = 0; kh_i < p.kh; ++kh_i) {
              mbar_wait(&b_empty[stage], phase ^ 1u);
              const int tap0 = (kd_i * p.kh + kh_i) * p.kw;
              if ((p.diag & 2) && b_loads >= p.b_stages) {
                if (leader && rank == 0) mbar_arrive(&b_full[stage]);
              } else if (leader && p.ds == 2) {
                // stacked weight rows: block b of the MMA's N = depth s_lo + b = depth tap kd_i - b
                const uint32_t dst0 = b_base + stage * p.b_stage_bytes;
                const uint32_t blk = (uint32_t)p.bn * ROWB;  // bytes of one depth's rows of a tap
                if constexpr (NCTA == 2) {
                  // the pair splits N in halves: with two depths CTA r holds depth s_lo + r whole (two boxes of
                  // bn/2 rows), with one depth half of its rows as in the unstacked case
                  const uint32_t per_cta = (uint32_t)p.kw * (ns == 2 ? blk : blk / 2);
                  if (rank == 0) mbar_expect_tx(&b_full[stage], 2u * per_cta);
                  for (int kw_i = 0; kw_i < p.kw; ++kw_i) {
                    if (ns == 2) {
                      const int tap = ((kd_i - rank) * p.kh + kh_i) * p.kw + kw_i;
                      tma_load_2d_pair(dst0 + kw_i * p.b_tap_bytes, tmB, &b_full[stage], tap * p.cin_pad + cc * KCH,
                                       t.nt * p.bn);
                      tma_load_2d_pair(dst0 + kw_i * p.b_tap_bytes + blk / 2, tmB, &b_full[stage],
                                       tap * p.cin_pad + cc * KCH, t.nt * p.bn + p.bn / 2);
                    } else {
                      tma_load_2d_pair(dst0 + kw_i * p.b_tap_bytes, tmB, &b_full[stage],
                                       (tap0 + kw_i) * p.cin_pad + cc * KCH, t.nt * p.bn + rank * (p.bn / 2));
                    }
                  }
                } else {
                  mbar_expect_tx(&b_full[stage], (uint32_t)(p.kw * ns) * blk);
                  for (int kw_i = 0; kw_i < p.kw; ++kw_i)
                    for (int b = 0; b < ns; ++b)
                      tma_load_2d(dst0 + kw_i * p.b_tap_bytes + b * blk, tmB, &b_full[stage],
                                  (((kd_i - b) * p.kh + kh_i) * p.kw + kw_i) * p.cin_pad + cc * KCH, t.nt * p.bn);
                }
              } else if (leader) {
                if constexpr (NCTA == 2) {
                  // The N rows of an MMA (kwm stacked taps of bn rows) are split in halves over the pair:
                  // kwm == 1: this CTA holds rows [rank*bn/2, +bn/2) of every tap; kwm == 2 / 4: it holds
                  // the taps rank*kwm/2 .. of every merged group.  b_tap_bytes is this CTA's share of a tap.
                  if (rank == 0) mbar_expect_tx(&b_full[stage], 2u * p.b_tx);
                  const uint32_t dst0 = b_base + stage * p.b_stage_bytes;
                  if (p.kwm == 1) {
                    for (int kw_i = 0; kw_i < p.kw; ++kw_i)
                      tma_load_2d_pair(dst0 + kw_i * p.b_tap_bytes, tmB, &b_full[stage],
                                       (tap0 + kw_i) * p.cin_pad + cc * KCH, t.nt * p.bn + rank * (p.bn / 2));
                  } else {
                    const int half = p.kwm / 2;
                    int slot_i = 0;
                    for (int g0 = 0; g0 < p.kw; g0 += p.kwm)
                      for (int j = 0; j < half; ++j, ++slot_i)
                        tma_load_2d_pair(dst0 + slot_i * p.b_tap_bytes, tmB, &b_full[stage],
                                         (tap0 + g0 + rank * half + j) * p.cin_pad + cc * KCH, t.nt * p.bn);
                  }
                } else {
                  mbar_expect_tx(&b_full[stage], p.b_tx);
                  for (int kw_i = 0; kw_i < p.kw; ++kw_i)
                    tma_load_2d(b_base + stage * p.b_stage_bytes + kw_i * p.b_tap_bytes, tmB, &b_full[stage],
                                (tap0 + kw_i) * p.cin_pad + cc * KCH, t.nt * p.bn);
                }
              }
              __syncwarp();
              ++b_loads;
              if (++stage == p.b_stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers (warp-uniform loop, one elected lane issues) =====================
    // warp 1 takes the even accumulators of a tile, warp 3 the odd ones (it idles on single-accumulator tiles)
    const int mw = warp == 1 ? 0 : 1, mstep = p.mt >= 2 ? 2 : 1;
    if (rank == 0 && (mw == 0 || p.mt >= 2)) {  // of a pair only the leader CTA issues
      const bool leader = elect_one();
      const uint32_t idesc1 = make_idesc_bf16(128 * NCTA, p.bn * p.kwm);
      const uint32_t idesc2 = make_idesc_bf16(128 * NCTA, p.bn * 2);  // two stacked depths (ds == 2: kwm == 1)
      const int ds = p.ds;
      const uint32_t desc_hi = smem_desc_hi(8 * ROWB, LAYOUT);  // 8-row swizzle atoms back to back
      // everything the loop needs, in registers (not re-read from the parameter bank per MMA)
      const int kd_n = p.kd, kh_n = p.kh, kw_n = p.kw, pd = p.pd, dd = p.dd, cin = p.cin, cchunks = p.cchunks;
      const int mt = p.mt, a_stages = p.a_stages, b_stages = p.b_stages;
      const uint32_t kwm = (uint32_t)p.kwm;
      const uint32_t slot = (uint32_t)p.slot;
      const uint32_t wp8 = (uint32_t)p.wp * ROW16;           // one padded row of pixels, in 16-byte units
      // 16-byte units between the weight operands of consecutive merged groups (per CTA: its half of N)
      const uint32_t b_grp16 = (NCTA == 2 && p.kwm > 1 ? (uint32_t)(p.kwm / 2) : (uint32_t)p.kwm) * (p.b_tap_bytes >> 4);
      const int nm = (mt - mw + mstep - 1) / mstep;  // accumulators of a tile this warp issues (1..4)
      const int ngroups = kw_n / (int)kwm;            // merged tap groups per filter row (<= 5: checked on the host)
      int as = 0, bs = 0;
      uint32_t aphase = 0, bphase = 0;
      int it = 0;
      for (int tile = item0; tile < p.num_tiles; tile += item_step, ++it) {
        const TileCoord t = decode_tile(p, tile, NCTA, rank);
        const int acc = p.acc_stages == 2 ? (it & 1) : 0;
        const uint32_t tphase = p.acc_stages == 2 ? (((uint32_t)it >> 1) & 1u) : ((uint32_t)it & 1u);
        mbar_wait(&t_empty[acc], tphase ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_acc + (uint32_t)acc * (uint32_t)(mt * ds) * slot;
        uint32_t accum = 0;  // 0 for the first MMA group of the tile (overwrite), 1 afterwards
        for (int si = 0; si < kd_n + ds - 1; ++si) {
          const int off = slab_offset(p, si);
          const int zd = t.dz + off;
          if (zd < 0 || zd >= dd) continue;
          // the tile's output depths this slab reaches (ds == 1: the one) and the MMA shape / TMEM columns for them
          const int s_lo = max(0, off + pd - (kd_n - 1)), ns = min(ds - 1, off + pd) - s_lo + 1;
          const uint32_t idesc = ns == 2 ? idesc2 : idesc1;
          const uint32_t d_slab = d_tmem + (uint32_t)s_lo * slot;
          for (int cc = 0; cc < cchunks; ++cc) {
            const int crem = cin - cc * KCH;
            const int ksteps = crem >= KCH ? KSTEPS : (crem + 15) / 16;
            mbar_wait(&a_full[as], aphase);
            const uint32_t slab_lo = smem_desc_lo(a_base + as * p.a_stage_bytes);
            uint32_t row_lo = slab_lo;  // + kh * padded row
            for (int kh_i = 0; kh_i < kh_n; ++kh_i, row_lo += wp8) {
              mbar_wait(&b_full[bs], bphase);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              uint32_t b_lo = smem_desc_lo(b_base + bs * p.b_stage_bytes);
              // kw taps in merged groups of kwm (one MMA of N = kwm*bn each): group g reads the slab g*kwm pixels
              // further and the next group of weight rows.  Unrolled over the (at most 5) groups so that every
              // operand address is base + constant * g: no loop-carried adds between the MMAs.
              const uint32_t a_row = row_lo + (uint32_t)mw * 128u * ROW16, d0 = d_slab + (uint32_t)(mw * ds) * slot;
              const uint32_t a_inc = (uint32_t)mstep * 128u * ROW16, d_inc = (uint32_t)(mstep * ds) * slot;
              const uint32_t a_grp = kwm * ROW16;
              if (leader) {
#pragma unroll
                for (int g = 0; g < 5; ++g) {
                  if (g < ngroups) {
                    const uint32_t a_lo = a_row + (uint32_t)g * a_grp, bg = b_lo + (uint32_t)g * b_grp16;
                    const uint32_t acc_g = g == 0 ? accum : 1u;
                    // the K steps that hold real channels (the last chunk of a tap may have fewer)
#define IVF_ISSUE_BLOCKS(KS)                                                                                       \
  do {                                                                                                             \
    umma_bf16_ksteps<KS, NCTA>(d0, a_lo, bg, desc_hi, idesc, acc_g);                                               \
    if (nm > 1) umma_bf16_ksteps<KS, NCTA>(d0 + d_inc, a_lo + a_inc, bg, desc_hi, idesc, acc_g);                   \
    if (nm > 2) umma_bf16_ksteps<KS, NCTA>(d0 + 2u * d_inc, a_lo + 2u * a_inc, bg, desc_hi, idesc, acc_g);         \
    if (nm > 3) umma_bf16_ksteps<KS, NCTA>(d0 + 3u * d_inc, a_lo + 3u * a_inc, bg, desc_hi, idesc, acc_g);         \
  } while (0)
                    if (ksteps == KSTEPS) {
                      IVF_ISSUE_BLOCKS(KSTEPS);
                    } else if (ksteps == 1) {
                      IVF_ISSUE_BLOCKS(1);
                    } else if (ksteps == 2) {
                      IVF_ISSUE_BLOCKS(2);
                    } else {
                      IVF_ISSUE_BLOCKS(3);
                    }
#undef IVF_ISSUE_BLOCKS
                  }
                }
              }
              accum = 1u;
              __syncwarp();
              if (leader) {  // weight slot free once these MMAs have read it
                if constexpr (NCTA == 2) umma_commit_pair(&b_empty[bs]); else umma_commit(&b_empty[bs]);
              }
              if (++bs == b_stages) {
                bs = 0;
                bphase ^= 1u;
              }
            }
            if (leader) {  // slab slot free
              if constexpr (NCTA == 2) umma_commit_pair(&a_empty[as]); else umma_commit(&a_empty[as]);
            }
            if (++as == a_stages) {
              as = 0;
              aphase ^= 1u;
            }
          }
        }
        if (leader) {  // accumulators of this tile complete
          if constexpr (NCTA == 2) umma_commit_pair(&t_full[acc]); else umma_commit(&t_full[acc]);
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int cgrp = (warp - SLAB_ROLE_WARPS) >> 2;  // which share of the tile's 16-column chunks this warp handles
    EpilogueArgs ea;
    ea.cout = p.cout;
    ea.flags = p.flags;
    ea.acc_in = acc_in;
    ea.mask_y = mask_y;
    ea.out = out;
    ea.lstm_c_prev = p.lstm_c_prev;
    ea.lstm_c_next = p.lstm_c_next;
    ea.lstm_h_next = p.lstm_h_next;
    // the tile loop, instantiated per epilogue-flag combination (FL >= 0: compile-time flags, see epilogue_chunk16)
    auto run_epilogue = [&](auto fl_tag) {
      constexpr int FL = decltype(fl_tag)::value;
      int it = 0;
      for (int tile = item0; tile < p.num_tiles; tile += item_step, ++it) {
        const TileCoord t = decode_tile(p, tile, NCTA, rank);
        const int acc = p.acc_stages == 2 ? (it & 1) : 0;
        mbar_wait(&t_full[acc], p.acc_stages == 2 ? (((uint32_t)it >> 1) & 1u) : ((uint32_t)it & 1u));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // this warp's chunks: [c_lo, c_hi)
        const int nchunks = p.bn >> 4;
        const int c_lo = 16 * (nchunks * cgrp / SLAB_EPI_GROUPS), c_hi = 16 * (nchunks * (cgrp + 1) / SLAB_EPI_GROUPS);
        float* xbase = nullptr;
        if (p.kwm > 1) {
          // kw-merge, out[v] = sum_g P_g[v + g]: block g of the accumulator, g rows further down.  The rows v+g
          // that fall into the NEXT 32-row segment (next TMEM lane quarter, or quarter 0 of the next accumulator)
          // come through shared memory.  Every warp first publishes the first kwm-1 rows of each of its segments
          // for its columns, ONE barrier per tile makes them visible, then the tile is finished without further
          // synchronisation (the exchange is double buffered by tile parity; the barrier of the next tile orders
          // this tile's reads before the buffer is written again two tiles later).
          const int kwm = p.kwm, bn = p.bn, seg = p.xch_seg;
          xbase = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.xch_off) +
                  (size_t)(it & 1) * 4 * p.mt * seg;
          for (int m = 0; m < p.mt; ++m) {
            const int sidx = m * 4 + q;
            if (sidx == 0) continue;  // nobody looks below the first segment
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * p.mt + m) * p.slot);
            float* xw = xbase + (size_t)sidx * seg;
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
              for (int g = 1; g < kwm; ++g) {
                uint32_t bt[16];
                tmem_ld16(taddr + g * bn + c0, bt);
                if (lane < kwm - 1) {
                  float* dst = xw + ((g - 1) * (kwm - 1) + lane) * bn + c0;
  #pragma unroll
                  for (int j = 0; j < 16; ++j) dst[j] = __uint_as_float(bt[j]);
                }
              }
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");  // boundary rows of the tile visible
        }
        for (int sm = 0; sm < p.ds * p.mt; ++sm) {
          const int sdep = p.ds == 1 ? 0 : sm / p.mt, m = p.ds == 1 ? sm : sm - sdep * p.mt;  // stacked depth, accumulator
          const int v = m * 128 + q * 32 + lane;  // padded-width pixel number inside the tile
          const int r = v / p.wp;
          const int c = v - r * p.wp;
          const int hrow = t.h0 + r;
          const bool ok = r < p.th && hrow < p.hh && c < p.ww;
          const size_t pix = (((size_t)t.nn * p.dd + t.dz + sdep) * p.hh + hrow) * p.ww + c;
          const size_t out_row = pix * p.out_ld + p.out_coff;
          const size_t mask_row = pix * p.mask_ld + p.mask_coff;
          const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) +
                                 (uint32_t)(((acc * p.mt + m) * p.ds + sdep) * p.slot);
          EpiPre cur;  // global operands of a chunk are requested right before its TMEM loads; the other warps of
                       // the scheduler cover the latency
          if (FL >= 0 && !LSTM && p.kwm == 1 && p.est_off != 0) {
            // Row-contiguous epilogue.  tcgen05.ld hands every lane ONE pixel's 16 channels, and pixels are out_ld
            // elements apart: a 16-byte access per lane touches 32 different lines per instruction, which is what
            // the L1 pipeline charges for (32 lines per request, ~8 K cycles per 128 x 128 fp32 tile - the ConvLSTM's
            // recurrent convolutions were bound by exactly this, 31 us for 10 us of MMAs).  The chunk is therefore
            // transposed through a 2 KB block of shared memory (XOR-swizzled float4 columns: conflict-free both
            // ways), after which four lanes cover one pixel's 16 channels and an instruction touches 8 lines; every
            // per-element operand (consumer sum, ReLU mask, BN vectors) is read in that mapping too, the global
            // ones before the TMEM load so that their latency overlaps it.
            constexpr int F = FL >= 0 ? FL : 0;
            float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.est_off) +
                         (size_t)(warp - SLAB_ROLE_WARPS) * 512;
            const int pixi = (ok && !(p.diag & 4)) ? (int)pix : -1;
            int prow[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) prow[it] = __shfl_sync(0xffffffffu, pixi, it * 8 + (lane >> 2));
            const int j4 = lane & 3, swz = (lane >> 1) & 3;
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
              const int nb = t.nt * p.bn + c0;
              if (nb + 16 > p.cout) {  // partial chunk: the per-lane form (scalar tail)
                epilogue_prefetch<FL>(ea, nb, out_row, mask_row, ok && !(p.diag & 4), cur);
                uint32_t rr[16];
                tmem_ld16(taddr + c0, rr);
                if (ok && nb < p.cout && !(p.diag & 4))
                  epilogue_chunk16<LSTM, FL>(ea, rr, nb, s_scale + nb, s_shift + nb, s_scale + nb, out_row, mask_row, cur);
                continue;
              }
              const int cb = nb + 4 * j4;
              float4 ac[4];
              uint2 mk[4];
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                if (prow[it] < 0) continue;
                if (F & IVF_EP_ACCUM)
                  ac[it] = *reinterpret_cast<const float4*>(acc_in + (size_t)prow[it] * p.out_ld + p.out_coff + cb);
                if (F & IVF_EP_MASK)
                  mk[it] = *reinterpret_cast<const uint2*>(mask_y + (size_t)prow[it] * p.mask_ld + p.mask_coff + cb);
              }
              uint32_t rr[16];
              tmem_ld16(taddr + c0, rr);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(stg + lane * 16 + ((j ^ swz) << 2)) =
                    make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]),
                                __uint_as_float(rr[4 * j + 3]));
              __syncwarp();
              float4 scv = make_float4(1.f, 1.f, 1.f, 1.f), shv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (F & (IVF_EP_AFFINE | IVF_EP_MASK)) scv = *reinterpret_cast<const float4*>(s_scale + cb);
              if (F & IVF_EP_AFFINE) shv = *reinterpret_cast<const float4*>(s_shift + cb);
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int row = it * 8 + (lane >> 2);
                float4 x = *reinterpret_cast<const float4*>(stg + row * 16 + ((j4 ^ ((row >> 1) & 3)) << 2));
                if (prow[it] < 0) continue;
                if (F & IVF_EP_ACCUM) {
                  x.x += ac[it].x; x.y += ac[it].y; x.z += ac[it].z; x.w += ac[it].w;
                }
                if (F & IVF_EP_AFFINE) {
                  x.x = fmaf(x.x, scv.x, shv.x); x.y = fmaf(x.y, scv.y, shv.y);
                  x.z = fmaf(x.z, scv.z, shv.z); x.w = fmaf(x.w, scv.w, shv.w);
                }
                if (F & IVF_EP_RELU) {
                  x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
                }
                if (F & IVF_EP_MASK) {  // ReLU'(y) * BN scale of the producer: y > 0 as the sign test of its bf16 bits
                  x.x = __uint_as_float(mk[it].x << 16) > 0.f ? x.x * scv.x : 0.f;
                  x.y = __uint_as_float(mk[it].x & 0xffff0000u) > 0.f ? x.y * scv.y : 0.f;
                  x.z = __uint_as_float(mk[it].y << 16) > 0.f ? x.z * scv.z : 0.f;
                  x.w = __uint_as_float(mk[it].y & 0xffff0000u) > 0.f ? x.w * scv.w : 0.f;
                }
                const size_t o = (size_t)prow[it] * p.out_ld + p.out_coff + cb;
                if (F & IVF_EP_OUT_F32) {
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = x;
                } else {
                  const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
                  uint2 pk;
                  pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                  pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + o) = pk;
                }
              }
              __syncwarp();  // the block is rewritten by the next chunk
            }
          } else if (LSTM && p.kwm == 1 && p.est_off != 0) {
            // The fused recurrent step, row-contiguous like the branch above: with unit-major channels the four
            // lanes of a pixel each hold ONE hidden unit's [i f c o] pre-activations, so the gate math
            // (pt/models/convolution_lstm.py:38-48, zero peepholes :50-54; the same expressions as epilogue_chunk16's
            // LSTM branch) runs one unit per lane and pass, and every state access is 16 bytes per pixel (c, h: 4 / 2
            // bytes per lane) next to its neighbours' instead of one pixel per lane.
            float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.est_off) +
                         (size_t)(warp - SLAB_ROLE_WARPS) * 512;
            const int pixi = (ok && !(p.diag & 4)) ? (int)pix : -1;
            int prow[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) prow[it] = __shfl_sync(0xffffffffu, pixi, it * 8 + (lane >> 2));
            const int j4 = lane & 3, swz = (lane >> 1) & 3;
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
              const int nb = t.nt * p.bn + c0;
              if (nb + 16 > p.cout) {  // partial chunk (hidden sizes that are no multiple of 4 units): per-lane form
                epilogue_prefetch<FL>(ea, nb, out_row, mask_row, ok && !(p.diag & 4), cur);
                uint32_t rr[16];
                tmem_ld16(taddr + c0, rr);
                if (ok && nb < p.cout && !(p.diag & 4))
                  epilogue_chunk16<LSTM, FL>(ea, rr, nb, s_scale + nb, s_shift + nb, s_scale + nb, out_row, mask_row, cur);
                continue;
              }
              const int cb = nb + 4 * j4;
              float4 ac[4];
              float cpv[4];
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                cpv[it] = 0.f;
                if (prow[it] < 0) continue;
                const size_t o = (size_t)prow[it] * p.out_ld + p.out_coff + cb;
                ac[it] = *reinterpret_cast<const float4*>(acc_in + o);
                if (p.lstm_c_prev) cpv[it] = p.lstm_c_prev[o >> 2];
              }
              uint32_t rr[16];
              tmem_ld16(taddr + c0, rr);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(stg + lane * 16 + ((j ^ swz) << 2)) =
                    make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]),
                                __uint_as_float(rr[4 * j + 3]));
              __syncwarp();
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int row = it * 8 + (lane >> 2);
                const float4 x = *reinterpret_cast<const float4*>(stg + row * 16 + ((j4 ^ ((row >> 1) & 3)) << 2));
                if (prow[it] < 0) continue;
                const size_t o = (size_t)prow[it] * p.out_ld + p.out_coff + cb;
                const float gi = ivf_sigmoid_f(x.x + ac[it].x), gf = ivf_sigmoid_f(x.y + ac[it].y);
                const float gg = tanhf(x.z + ac[it].z), go = ivf_sigmoid_f(x.w + ac[it].w);
                const float cn = fmaf(gf, cpv[it], gi * gg);
                const float hn = go * tanhf(cn);
                p.lstm_c_next[o >> 2] = cn;
                p.lstm_h_next[o >> 2] = __float2bfloat16_rn(hn);
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = make_float4(gi, gf, gg, go);
              }
              __syncwarp();
            }
          } else if (p.kwm == 1) {
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
              const int nb = t.nt * p.bn + c0;
              epilogue_prefetch<FL>(ea, nb, out_row, mask_row, ok && !(p.diag & 4), cur);
              uint32_t rr[16];
              tmem_ld16(taddr + c0, rr);
              if (ok && nb < p.cout && !(p.diag & 4))
                epilogue_chunk16<LSTM, FL>(ea, rr, nb, s_scale + nb, s_shift + nb, s_scale + nb, out_row, mask_row, cur);
            }
          } else {
            const int kwm = p.kwm, bn = p.bn;
            // the segment after this one (rows past the last segment of the tile are padding nobody stores)
            const int snext = min(m * 4 + q + 1, 4 * p.mt - 1);
            const float* xr = xbase + (size_t)snext * p.xch_seg;
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
              epilogue_prefetch<FL>(ea, t.nt * bn + c0, out_row, mask_row, ok, cur);
              uint32_t tg[16];
              tmem_ld16(taddr + c0, tg);
              float accv[16];
  #pragma unroll
              for (int j = 0; j < 16; ++j) accv[j] = __uint_as_float(tg[j]);
              for (int g = 1; g < kwm; ++g) {
                tmem_ld16(taddr + g * bn + c0, tg);
                const int src = lane + g - 32;  // >= 0: the row lives in the next segment
                const float* xs = xr + ((g - 1) * (kwm - 1) + (src >= 0 ? src : 0)) * bn + c0;
  #pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float sv = __shfl_down_sync(0xffffffffu, __uint_as_float(tg[j]), g);
                  if (src >= 0) sv = xs[j];
                  accv[j] += sv;
                }
              }
              const int nb = t.nt * bn + c0;
              if (ok && nb < p.cout) {
                uint32_t rr[16];
  #pragma unroll
                for (int j = 0; j < 16; ++j) rr[j] = __float_as_uint(accv[j]);
                epilogue_chunk16<LSTM, FL>(ea, rr, nb, s_scale + nb, s_shift + nb, s_scale + nb, out_row, mask_row, cur);
              }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if constexpr (NCTA == 2) mbar_arrive_leader(&t_empty[acc]); else mbar_arrive(&t_empty[acc]);
        }
      }
    };
    if (LSTM) {
      run_epilogue(std::integral_constant<int, -1>{});
    } else {
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
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();  // the peer may still be reading this SM's operands / signalling
  if (warp == 2) {
    __syncwarp();
    if constexpr (NCTA == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------

struct SlabKeyA {
  const void* base;
  int n, id, ih, iw, cin, ld, coff, wp, rows, kch;
};
struct SlabKeyB {
  const void* base;
  int ktot, cout_pad, bn, kch;
};

template <typename K>
std::string slab_key(char tag, const K& k) {
  std::string s(1, tag);
  s.append(reinterpret_cast<const char*>(&k), sizeof(K));
  return s;
}

int slab_map_a(ivf_handle* h, const ivf_conv_desc* d, const void* in, int wp, int rows, int kch,
               CUtensorMap* out) {
  SlabKeyA key;
  memset(&key, 0, sizeof(key));
  key.base = in;
  key.n = d->n; key.id = d->id; key.ih = d->ih; key.iw = d->iw; key.cin = d->cin;
  key.ld = d->in_ld; key.coff = d->in_coff; key.wp = wp; key.rows = rows; key.kch = kch;
  std::string kb = slab_key('S', key);
  {
    std::lock_guard<std::mutex> g(h->mu);
    auto it = h->tmaps.find(kb);
    if (it != h->tmaps.end()) {
      *out = it->second;
      return IVF_OK;
    }
  }
  const char* base = reinterpret_cast<const char*>(in) + (size_t)d->in_coff * 2;
  IVF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "conv(slab): input slice not 16-B aligned");
  cuuint64_t dims[5] = {(cuuint64_t)d->cin, (cuuint64_t)d->iw, (cuuint64_t)d->ih, (cuuint64_t)d->id,
                        (cuuint64_t)d->n};
  cuuint64_t pix = (cuuint64_t)d->in_ld * 2;
  cuuint64_t strides[4] = {pix, pix * d->iw, pix * d->iw * d->ih, pix * d->iw * d->ih * d->id};
  cuuint32_t box[5] = {(cuuint32_t)kch, (cuuint32_t)wp, (cuuint32_t)rows, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = ivf_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)base, dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                kch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    IVF_FAIL(IVF_ECUDA, "cuTensorMapEncodeTiled(slab) failed (%d): c%d w%d h%d d%d n%d ld%d box %dx%dx%d",
             (int)r, d->cin, d->iw, d->ih, d->id, d->n, d->in_ld, kch, wp, rows);
  {
    std::lock_guard<std::mutex> g(h->mu);
    h->tmaps[kb] = m;
  }
  *out = m;
  return IVF_OK;
}

int slab_map_b(ivf_handle* h, const void* w, int ktot, int cout_pad, int bn, int kch, CUtensorMap* out) {
  // bn here = rows of one TMA box (the N tile, or half of it per CTA of a pair)
  SlabKeyB key;
  memset(&key, 0, sizeof(key));
  key.base = w; key.ktot = ktot; key.cout_pad = cout_pad; key.bn = bn; key.kch = kch;
  std::string kb = slab_key('T', key);
  {
    std::lock_guard<std::mutex> g(h->mu);
    auto it = h->tmaps.find(kb);
    if (it != h->tmaps.end()) {
      *out = it->second;
      return IVF_OK;
    }
  }
  IVF_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "conv(slab): weights not 16-B aligned");
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)cout_pad};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)kch, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = ivf_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE,
                                kch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    IVF_FAIL(IVF_ECUDA, "cuTensorMapEncodeTiled(slab weights) failed (%d): ktot %d cout_pad %d bn %d", (int)r,
             ktot, cout_pad, bn);
  {
    std::lock_guard<std::mutex> g(h->mu);
    h->tmaps[kb] = m;
  }
  *out = m;
  return IVF_OK;
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

// Tile configuration by a small cost model.  Candidates: N tiles (bn), kw taps merged into N (kwm),
// accumulators per tile (mt), single or double buffered TMEM.  What bounds this kernel is the tensor core's
// operand fetch from shared memory, not L2 or DRAM (ncu --set full: profiles/r01_ncu_full_conv_slab.md):
// measured over all the configurations of round 1 an M128 x N x K16 MMA (cta_group::1, both operands in
// shared memory) takes ~ 64 + N/2 cycles = (4 KB activations + 32N B weights) at 64 B/clk, against N/2 cycles
// of math: N = 32 -> 80 clk, 64 -> 96, 96 -> 112, 128 -> 128.  So efficiency is N/(128+N): as large an N as
// TMEM allows (kwm*bn <= 256), and the rest of the model only arbitrates ties:
//   mma   = #MMA x (64 + N/2) + half of the TMA shared-memory writes ((slab + weight bytes) / 128)
//   l2    = (slab bytes read + weight bytes) / 40      (B/clk/SM with all 148 SMs pulling = the LTS cap; a slab
//           pixel costs at least 128 B: the 48-byte rows of the 24-channel stem operand move at that rate)
// the epilogue is exposed only when TMEM is single buffered.  Returns false when nothing fits.
thread_local int g_force_kch = 0;  // grouped launches: 128-byte slab rows whatever the channel count
bool slab_config_impl(const ivf_conv_desc* d, int sm_count, SlabParams* best);
bool slab_config(const ivf_conv_desc* d, int sm_count, SlabParams* best) {
  // depth stacking is a request (plan_ds from the host-side tuner, IVF_SLAB_DS for diagnostics): a layer that
  // cannot stack (odd depth, no front padding, N too wide) silently runs unstacked
  ivf_conv_desc dd_ = *d;
  if (!dd_.plan_ds) dd_.plan_ds = env_int("IVF_SLAB_DS", 1);
  if (dd_.plan_ds == 2) {
    if (slab_config_impl(&dd_, sm_count, best)) return true;
    dd_.plan_ds = 1;
  }
  d = &dd_;
  if (slab_config_impl(d, sm_count, best)) return true;
  if (d->plan_kwm | d->plan_mt | d->plan_acc | d->plan_ncta | d->plan_ntiles | d->plan_ds) {  // unsatisfiable request
    ivf_conv_desc auto_d = *d;
    auto_d.plan_kwm = auto_d.plan_mt = auto_d.plan_acc = auto_d.plan_ncta = auto_d.plan_ntiles = 0;
    auto_d.plan_ds = 1;
    return slab_config_impl(&auto_d, sm_count, best);
  }
  return false;
}
bool slab_config_impl(const ivf_conv_desc* d, int sm_count, SlabParams* best) {
  memset(best, 0, sizeof(*best));
  const int cout = d->cout, cin = d->cin;
  // 64-byte rows (SWIZZLE_64B) cost more per MMA than 128-byte rows (tools/mma_bench.cu); IVF_SLAB_KCH64=1 gives the
  // narrow layers 128-byte rows too (the upper half is TMA zero fill and is never multiplied)
  const int kch = (cin <= 32 && !env_int("IVF_SLAB_KCH64", 0) && g_force_kch != 64) ? 32 : 64;
  const int rowb = kch * 2;
  const int wp = d->iw + d->kw - 1;
  const int cchunks = (cin + kch - 1) / kch;
  const int taps = d->kd * d->kh * d->kw;
  int ksteps_total = 0;  // MMAs of K=16 per tap over all channel chunks
  for (int cc = 0; cc < cchunks; ++cc) {
    int crem = cin - cc * kch;
    ksteps_total += crem >= kch ? kch / 16 : (crem + 15) / 16;
  }
  // requests: the descriptor's plan fields (host-side tuner), else the environment (diagnostics), else none
  const int forced_mt = d->plan_mt ? d->plan_mt : env_int("IVF_SLAB_MT", 0);
  const int forced_nt = d->plan_ntiles ? d->plan_ntiles : env_int("IVF_SLAB_NT", 0);
  const int forced_acc = d->plan_acc ? d->plan_acc : env_int("IVF_SLAB_ACC", 0);
  const int forced_kwm = d->plan_kwm ? d->plan_kwm : env_int("IVF_SLAB_KWM", 0);
  const int forced_ncta = d->plan_ncta;
  // depth stacking: 1 or 2, resolved by slab_config
  const int forced_ds = d->plan_ds ? d->plan_ds : 1;
  const bool ds2_ok = d->id % 2 == 0 && d->pd >= 1 && d->kd >= 2 && !(d->flags & IVF_EP_LSTM);
  const bool allow_pair = env_int("IVF_SLAB_2CTA", 1) != 0 && sm_count % 2 == 0;
  double best_cost = 1e30;
  bool found = false;
  for (int ntiles = 1; ntiles <= 16; ++ntiles) {
    if (forced_nt && ntiles != forced_nt) continue;
    const int bn = ((cout + ntiles - 1) / ntiles + 15) / 16 * 16;
    if (bn > 256 || ntiles * bn > SLAB_MAX_COUT + 15) continue;
    if (ntiles > 1 && bn < 32) continue;
    for (int ncta = 1; ncta <= 2; ++ncta) {
    if (ncta == 2 && !allow_pair) continue;
    if (forced_ncta && ncta != forced_ncta) continue;
    for (int kwm = 1; kwm <= 4 && kwm <= d->kw; ++kwm) {
    if (d->kw % kwm || kwm * bn > 256) continue;
    if (ncta == 2 && kwm == 3) continue;  // the pair splits the stacked N rows in halves: whole taps or half a tap
    if (forced_kwm && kwm != forced_kwm && !(forced_kwm > 1 && (d->kw % forced_kwm || forced_kwm * bn > 256))) continue;
    for (int ds = 1; ds <= 2; ++ds) {
    if (forced_ds && ds != forced_ds) continue;
    // two stacked depths: their accumulators are adjacent TMEM column blocks of exactly bn columns
    if (ds == 2 && (!ds2_ok || kwm != 1 || bn % 32 || 2 * bn > 256)) continue;
    const int slot = (kwm * bn + 31) / 32 * 32;
    // one CTA's share of a tap's weight rows: all bn, or (pair) bn/2 when kwm == 1; with kwm 2/4 a pair CTA holds
    // kwm/2 whole taps per merged group.  Taps are dense (bn % 16 == 0 keeps them on swizzle-atom boundaries).
    // Two stacked depths double the rows of a tap (a pair CTA then holds one depth's bn rows).
    const uint32_t b_tap = (uint32_t)((ncta == 2 && kwm == 1) ? bn / 2 : bn) * rowb * ds;
    const uint32_t b_stage = (uint32_t)bn * rowb * (uint32_t)d->kw / ncta * ds;
    if (b_tap % (8u * rowb)) continue;
    if (ds == 2 && ncta == 2 && (bn / 2) % 8) continue;
    for (int acc_stages = 2; acc_stages >= 1; --acc_stages) {
      if (forced_acc && acc_stages != forced_acc) continue;
      for (int mt = 4; mt >= 1; --mt) {
        if (forced_mt && mt != forced_mt) continue;
        int th = (mt * 128) / wp;
        if (th < 1) continue;
        if (th > d->ih) th = d->ih;
        const int htiles = (d->ih + th - 1) / th;
        th = (d->ih + htiles - 1) / htiles;  // balance the rows over the tiles
        const int mt_eff = (th * wp + 127) / 128;
        if (acc_stages * mt_eff * ds * slot > 512) continue;
        const int rows = th + d->kh - 1;
        if (rows > 256) continue;
        const uint32_t a_stage =
            (((uint32_t)(mt_eff * 128 + (d->kh - 1) * wp + d->kw) * rowb) + 1023u) & ~1023u;
        // kw-merge: boundary-row exchange of the epilogue, [2 tile parities][4*mt segments][(kwm-1)^2 * bn] floats
        const int xch_seg = kwm > 1 ? (kwm - 1) * (kwm - 1) * bn : 0;
        const uint32_t xch_bytes = (uint32_t)(2 * 4 * mt_eff * xch_seg) * 4u;
        const uint32_t budget = SLAB_SMEM_BUDGET - ((xch_bytes + 1023u) & ~1023u);
        if (2 * a_stage + 2 * b_stage > budget) continue;
        int a_stages = 2;
        if (3 * a_stage + 3 * b_stage <= budget) a_stages = 3;
        int b_stages = (int)((budget - (uint32_t)a_stages * a_stage) / b_stage);
        if (b_stages > MAX_B_STAGES) b_stages = MAX_B_STAGES;
        // ---- cost
        const double tiles = (double)d->n * (d->id / ds) * ((htiles + ncta - 1) / ncta) * ntiles;  // work items
        const double waves = ceil(tiles / (sm_count / ncta));
        // per tile: kd + ds - 1 slabs; with two depths the first and the last run at N = bn, the others at 2 bn -
        // modelled as (kd + 1) slabs' MMAs at the wide N (slightly pessimistic)
        const double n_mma = (double)(d->kd + ds - 1) * d->kh * d->kw / kwm * ksteps_total * mt_eff;
        const int n_eff = kwm * bn * ds;
        const int cin_real_bytes = (cin < kch ? cin : kch) * 2;
        const double slab_smem = (double)(d->kd + ds - 1) * cchunks * rows * wp * rowb;
        const double slab_l2 = (double)(d->kd + ds - 1) * cchunks * rows * wp * (cin_real_bytes < 128 ? 128 : cin_real_bytes);
        const double w_bytes = (double)taps * ds * cchunks * bn * rowb / ncta;  // per SM
        // per SM: 4 KB of activations + its share of the weight rows at ~64 B/clk, never below the math
        double per_mma = 64.0 + n_eff / (2.0 * ncta);
        if (per_mma < n_eff / 2.0) per_mma = n_eff / 2.0;
        const double mma_clk = n_mma * per_mma + 0.5 * (slab_smem + w_bytes) / 128.0;
        const double l2_clk = (slab_l2 + w_bytes) / 40.0;
        const double epi_clk = (double)mt_eff * ds * (bn / 16) *
                               (220.0 + 200.0 * (kwm - 1) + ((d->flags & (IVF_EP_MASK | IVF_EP_ACCUM)) ? 150.0 : 0.0)) +
                               ((d->flags & (IVF_EP_MASK | IVF_EP_ACCUM)) ? 1200.0 * mt_eff * ds : 0.0);
        double tile_clk = mma_clk > l2_clk ? mma_clk : l2_clk;
        if (acc_stages == 1) tile_clk += epi_clk;
        else if (epi_clk > tile_clk) tile_clk = epi_clk;
        const double cost = waves * tile_clk + 4000.0;
        if (cost < best_cost) {
          best_cost = cost;
          found = true;
          SlabParams* p = best;
          p->kch = kch;
          p->wp = wp;
          p->th = th;
          p->htiles = htiles;
          p->mt = mt_eff;
          p->bn = bn;
          p->ntiles = ntiles;
          p->slot = slot;
          p->kwm = kwm;
          p->ds = ds;
          p->dgroups = d->id / ds;
          p->ncta = ncta;
          p->acc_stages = acc_stages;
          p->a_stages = a_stages;
          p->b_stages = b_stages;
          p->xch_seg = xch_seg;
          // transposition staging of the epilogue (kwm == 1 only).  fp32 outputs get it (a weight stage is given up
          // if need be, never below two): ConvLSTM step 3.90 -> 3.49 ms at 8 clips, 10.15 -> 8.54 ms at 32.  bf16
          // outputs do not by default - their 32-byte row pieces are a quarter of the lines per element, the
          // epilogue hides behind the MMAs and the I3D step measured 0.5 % slower with it.  IVF_SLAB_XPOSE: 0 = never,
          // 1 = fp32 outputs (default), 2 = bf16 too where the plan leaves the room, 3 = bf16 too, shrinking
          p->est_off = 0;
          const int xmode = env_int("IVF_SLAB_XPOSE", 1);
          if (kwm == 1 && xmode && ((d->flags & IVF_EP_OUT_F32) || xmode >= 2) &&
              (long long)d->n * d->id * d->ih * d->iw < (1ll << 31) && d->out_ld % 4 == 0 && d->out_coff % 4 == 0) {
            const bool may_shrink = (d->flags & IVF_EP_OUT_F32) || xmode == 3;
            auto used = [&](int bst) { return (uint32_t)a_stages * a_stage + (((uint32_t)bst * b_stage + 1023u) & ~1023u); };
            int bst = b_stages;
            while (used(bst) + SLAB_EST_BYTES > SLAB_SMEM_BUDGET && may_shrink && bst > 2) --bst;
            if (used(bst) + SLAB_EST_BYTES <= SLAB_SMEM_BUDGET) {
              b_stages = bst;
              p->est_off = used(bst);
            }
          }
          p->b_stages = b_stages;
          p->xch_off = (uint32_t)a_stages * a_stage + (((uint32_t)b_stages * b_stage + 1023u) & ~1023u);
          p->a_stage_bytes = a_stage;
          p->b_stage_bytes = b_stage;
          p->b_tap_bytes = b_tap;
          p->a_tx = (uint32_t)rows * wp * rowb;
          p->b_tx = b_stage;
          int cols = 32;
          while (cols < acc_stages * mt_eff * ds * slot) cols <<= 1;
          p->tmem_cols = cols;
        }
      }
    }
    }
    }
    }
  }
  return found;
}

template <int KCH, int NCTA, bool LSTM>
int slab_launch_t(ivf_handle* h, const SlabParams& p, const CUtensorMap& ma, const CUtensorMap& mb,
                  const float* scale, const float* shift, const float* acc_in, const void* mask_y,
                  const float* mask_scale, void* out, cudaStream_t st) {
  const int slot = (KCH == 64 ? 0 : 1) + 2 * (NCTA - 1) + (LSTM ? 4 : 0);
  if (!h->slab_attr_set[slot]) {
    IVF_CUDA(cudaFuncSetAttribute(conv_slab_kernel<KCH, NCTA, LSTM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(SLAB_SMEM_BUDGET + 1024)));
    h->slab_attr_set[slot] = true;
  }
  const size_t smem = (size_t)p.xch_off + (size_t)2 * 4 * p.mt * p.xch_seg * 4 + (p.est_off ? SLAB_EST_BYTES : 0u) + 1024;
  const int units = h->sm_count / NCTA;  // CTAs, or CTA pairs
  const int grid = (p.num_tiles < units ? p.num_tiles : units) * NCTA;
  const SlabOperands o = {scale, shift, acc_in, (const __nv_bfloat16*)mask_y, mask_scale, out};
  IVF_CUDA(ivf_launch(conv_slab_kernel<KCH, NCTA, LSTM, false>, dim3(grid), dim3(SLAB_THREADS), smem, st, NCTA, ma, mb, p, o,
                      ma, mb, p, o, grid));
  IVF_LAUNCHED(h);
  return IVF_OK;
}

// two problems, one launch (GROUPED): single CTAs, 128-byte slab rows for both
int slab_launch_group(ivf_handle* h, const SlabParams (&p)[2], const CUtensorMap (&ma)[2], const CUtensorMap (&mb)[2],
                      const SlabOperands (&o)[2], int split, int grid, cudaStream_t st) {
  if (!h->slab_attr_set[15]) {
    IVF_CUDA(cudaFuncSetAttribute(conv_slab_kernel<64, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(SLAB_SMEM_BUDGET + 1024)));
    h->slab_attr_set[15] = true;
  }
  size_t smem = 0;
  for (int i = 0; i < 2; ++i)
    smem = std::max(smem, (size_t)p[i].xch_off + (size_t)2 * 4 * p[i].mt * p[i].xch_seg * 4 +
                              (p[i].est_off ? SLAB_EST_BYTES : 0u) + 1024);
  IVF_CUDA(ivf_launch(conv_slab_kernel<64, 1, false, true>, dim3(grid), dim3(SLAB_THREADS), smem, st, 1, ma[0], mb[0], p[0],
                      o[0], ma[1], mb[1], p[1], o[1], split));
  IVF_LAUNCHED(h);
  return IVF_OK;
}

}  // namespace

namespace {
// A 1x1x1 convolution has no halo: its (clip, depth, row) axes are one long row axis, so a tile is any run
// of th rows (full 128-row accumulators on the 14x14 and 7x7 maps too) and the kernel is a persistent GEMM
// with the epilogue overlapped.  Measured (round 1): per launch it ties with the im2col kernel's one-tile
// CTAs (1.09 vs 1.11 ms per step over the 64 1x1x1 launches) but a persistent 148-CTA kernel with ~200 KB of
// shared memory cannot share the SMs with the launches of the other Inception branches, and the step got
// slower (3.18 vs 3.08 ms) - so this route is opt-in (IVF_SLAB_1X1=1).
ivf_conv_desc slab_view(const ivf_conv_desc* d) {
  ivf_conv_desc v = *d;
  if (d->kd * d->kh * d->kw == 1) {
    v.ih = v.oh = d->n * d->id * d->ih;
    v.id = v.od = 1;
    v.n = 1;
  }
  return v;
}
}  // namespace

bool ivf_conv3d_slab_eligible(const ivf_handle* h, const ivf_conv_desc* d0) {
  const ivf_conv_desc dv = slab_view(d0);
  const ivf_conv_desc* d = &dv;
  if (env_int("IVF_SLAB", 1) == 0) return false;
  if (d->dtype != IVF_BF16 || d->transposed) return false;
  if (d->sd != 1 || d->sh != 1 || d->sw != 1) return false;
  if (d->od != d->id || d->oh != d->ih || d->ow != d->iw) return false;
  if (d->kd * d->kh * d->kw == 1 && env_int("IVF_SLAB_1X1", 0) == 0) return false;
  if ((long long)d0->n * d0->id * d0->ih >= (1ll << 30)) return false;
  if (d->cin % 8 || d->in_ld % 8 || d->in_coff % 8 || d->out_ld % 8 || d->out_coff % 8) return false;
  if ((d->flags & IVF_EP_MASK) && (d->mask_ld % 8 || d->mask_coff % 8)) return false;
  // 'same' geometry: front pad within the kernel, the rest is the back pad
  if (d->pd < 0 || d->pd >= d->kd || d->ph < 0 || d->ph >= d->kh || d->pw < 0 || d->pw >= d->kw) return false;
  if (d->iw < env_int("IVF_SLAB_MIN_W", 7)) return false;  // very narrow maps waste the padded-width tile
  if (d->iw + d->kw - 1 > 256) return false;
  if (d->kw > 5) return false;  // the issue loop is unrolled over at most five tap groups per filter row
  if (d->cout > SLAB_MAX_COUT) return false;
  if ((d->flags & IVF_EP_AFFINE) && (d->flags & IVF_EP_MASK)) return false;  // one shared scale vector
  SlabParams p;
  return slab_config(d, h->sm_count, &p);
}

namespace {
// the fields of SlabParams that come straight from the descriptor (after slab_config chose the plan)
void slab_fill(const ivf_conv_desc* d, SlabParams& p) {
  p.n = d->n; p.dd = d->id; p.hh = d->ih; p.ww = d->iw;
  p.kd = d->kd; p.kh = d->kh; p.kw = d->kw;
  p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
  p.cin = d->cin;
  p.cchunks = (d->cin + p.kch - 1) / p.kch;
  p.cin_pad = ivf_conv_bf16_cin_pad(d->cin);
  p.cout = d->cout;
  p.out_ld = d->out_ld; p.out_coff = d->out_coff;
  p.mask_ld = d->mask_ld; p.mask_coff = d->mask_coff;
  p.flags = d->flags;
  p.num_tiles = d->n * p.dgroups * ((p.htiles + p.ncta - 1) / p.ncta) * p.ntiles;
  p.diag = env_int("IVF_SLAB_DIAG", 0);
}
}  // namespace

int ivf_conv3d_slab_launch_pair(ivf_handle* h, const ivf_conv_desc* const d[2], const ivf_conv_operands op[2],
                                cudaStream_t st) {
  if (env_int("IVF_SLAB_GROUP", 1) == 0) return IVF_EUNSUPPORTED;
  int rc = ivf_load_driver_entry_points();
  if (rc) return rc;
  SlabParams p[2];
  CUtensorMap ma[2], mb[2];
  SlabOperands o[2];
  double cost[2];
  for (int i = 0; i < 2; ++i) {
    if (d[i]->kd * d[i]->kh * d[i]->kw == 1 || (d[i]->flags & IVF_EP_LSTM)) return IVF_EUNSUPPORTED;
    if (!ivf_conv3d_slab_eligible(h, d[i])) return IVF_EUNSUPPORTED;
    ivf_conv_desc dd = *d[i];
    dd.plan_ncta = 1;                   // single CTAs: the two problems share one grid
    dd.plan_kwm = dd.plan_kwm ? dd.plan_kwm : 0;
    g_force_kch = 64;                   // one kernel instantiation: 128-byte slab rows for both
    const bool ok = slab_config(&dd, h->sm_count, &p[i]);
    g_force_kch = 0;
    if (!ok || p[i].ncta != 1 || p[i].kch != 64) return IVF_EUNSUPPORTED;
    slab_fill(&dd, p[i]);
    p[i].lstm_c_prev = nullptr; p[i].lstm_c_next = nullptr; p[i].lstm_h_next = nullptr;
    if ((long long)dd.n * p[i].dgroups * p[i].htiles * p[i].ntiles >= (1ll << 31)) return IVF_EUNSUPPORTED;
    // relative cost of a problem: its MMAs (the cost model's per-instruction estimate) - decides the SM split
    int ksteps_total = 0;
    for (int cc = 0; cc < p[i].cchunks; ++cc) {
      const int crem = dd.cin - cc * 64;
      ksteps_total += crem >= 64 ? 4 : (crem + 15) / 16;
    }
    const int n_eff = p[i].kwm * p[i].bn * p[i].ds;
    cost[i] = (double)p[i].num_tiles * (dd.kd + p[i].ds - 1) * dd.kh * dd.kw / p[i].kwm * ksteps_total * p[i].mt *
                  std::max(64.0 + n_eff / 2.0, (double)n_eff / 2.0) +
              (double)p[i].num_tiles * 4000.0;  // per-tile fixed part (pipeline fill, epilogue)
  }
  // CTAs: as many as there are tiles, at most one per SM, split by cost with at least one CTA each
  const int total = std::min(h->sm_count, p[0].num_tiles + p[1].num_tiles);
  if (total < 2) return IVF_EUNSUPPORTED;
  int g0 = (int)std::lround(total * cost[0] / (cost[0] + cost[1]));
  g0 = std::max(1, std::min(g0, total - 1));
  g0 = std::min(g0, p[0].num_tiles);
  int g1 = std::min(total - g0, p[1].num_tiles);
  if (g0 + g1 < total) g0 = std::min(p[0].num_tiles, total - g1);
  for (int i = 0; i < 2; ++i) {
    rc = slab_map_a(h, d[i], op[i].in, p[i].wp, p[i].th + d[i]->kh - 1, 64, &ma[i]);
    if (rc) return rc;
    rc = slab_map_b(h, op[i].w, d[i]->kd * d[i]->kh * d[i]->kw * p[i].cin_pad, ivf_conv_bf16_cout_pad(d[i]->cout), p[i].bn,
                    64, &mb[i]);
    if (rc) return rc;
    o[i] = {op[i].scale, op[i].shift, op[i].acc_in, (const __nv_bfloat16*)op[i].mask_y, op[i].mask_scale, op[i].out};
  }
  if (env_int("IVF_SLAB_VERBOSE", 0))
    fprintf(stderr, "slab group: %d + %d CTAs | c%d->%d mt %d th %d items %d | c%d->%d mt %d th %d items %d\n", g0, g1,
            d[0]->cin, d[0]->cout, p[0].mt, p[0].th, p[0].num_tiles, d[1]->cin, d[1]->cout, p[1].mt, p[1].th, p[1].num_tiles);
  return slab_launch_group(h, p, ma, mb, o, g0, g0 + g1, st);
}

int ivf_conv3d_slab_launch(ivf_handle* h, const ivf_conv_desc* d0, const void* in, const void* w,
                           const float* scale, const float* shift, const float* acc_in,
                           const void* mask_y, const float* mask_scale, void* out, cudaStream_t st,
                           const float* lstm_c_prev, float* lstm_c_next, void* lstm_h_next) {
  const ivf_conv_desc dv = slab_view(d0);
  const ivf_conv_desc* d = &dv;
  int rc = ivf_load_driver_entry_points();
  if (rc) return rc;
  SlabParams p;
  if (!slab_config(d, h->sm_count, &p)) IVF_FAIL(IVF_EUNSUPPORTED, "conv(slab): no tile configuration fits");
  p.lstm_c_prev = lstm_c_prev;
  p.lstm_c_next = lstm_c_next;
  p.lstm_h_next = (__nv_bfloat16*)lstm_h_next;
  p.n = d->n; p.dd = d->id; p.hh = d->ih; p.ww = d->iw;
  p.kd = d->kd; p.kh = d->kh; p.kw = d->kw;
  p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
  p.cin = d->cin;
  p.cchunks = (d->cin + p.kch - 1) / p.kch;
  p.cin_pad = ivf_conv_bf16_cin_pad(d->cin);
  p.cout = d->cout;
  p.out_ld = d->out_ld; p.out_coff = d->out_coff;
  p.mask_ld = d->mask_ld; p.mask_coff = d->mask_coff;
  p.flags = d->flags;
  long long tiles = (long long)d->n * p.dgroups * ((p.htiles + p.ncta - 1) / p.ncta) * p.ntiles;  // work items
  IVF_REQUIRE(tiles < (1ll << 31), "conv(slab): too many tiles");
  p.num_tiles = (int)tiles;
  p.diag = env_int("IVF_SLAB_DIAG", 0);
  if (env_int("IVF_SLAB_VERBOSE", 0))
    fprintf(stderr, "slab: %dx%dx%d c%d->%d k%d%d%d | kch %d bn %d x%d kwm %d ds %d mt %d th %d acc %d a_st %d(%u) b_st %d(%u) items %d ncta %d\n",
            d->id, d->ih, d->iw, d->cin, d->cout, d->kd, d->kh, d->kw, p.kch, p.bn, p.ntiles, p.kwm, p.ds, p.mt, p.th,
            p.acc_stages, p.a_stages, p.a_stage_bytes, p.b_stages, p.b_stage_bytes, p.num_tiles, p.ncta);

  CUtensorMap ma, mb;
  rc = slab_map_a(h, d, in, p.wp, p.th + d->kh - 1, p.kch, &ma);
  if (rc) return rc;
  const int ntaps = d->kd * d->kh * d->kw;
  const int box_rows = (p.ncta == 2 && p.kwm == 1) ? p.bn / 2 : p.bn;
  rc = slab_map_b(h, w, ntaps * p.cin_pad, ivf_conv_bf16_cout_pad(d->cout), box_rows, p.kch, &mb);
  if (rc) return rc;
#define IVF_SLAB_GO(K, C, L) return slab_launch_t<K, C, L>(h, p, ma, mb, scale, shift, acc_in, mask_y, mask_scale, out, st)
  const bool lstm = (p.flags & IVF_EP_LSTM) != 0;
  if (p.ncta == 2) {
    if (p.kch == 64) { if (lstm) IVF_SLAB_GO(64, 2, true); IVF_SLAB_GO(64, 2, false); }
    if (lstm) IVF_SLAB_GO(32, 2, true);
    IVF_SLAB_GO(32, 2, false);
  }
  if (p.kch == 64) { if (lstm) IVF_SLAB_GO(64, 1, true); IVF_SLAB_GO(64, 1, false); }
  if (lstm) IVF_SLAB_GO(32, 1, true);
  IVF_SLAB_GO(32, 1, false);
#undef IVF_SLAB_GO
}

// diagnostic (no GPU needed): the tile plan the slab kernel would use for a layer, or 0 when the layer
// goes to the im2col kernel.  plan = {kch, bn, ntiles, mt, th, acc_stages, a_stages, b_stages, tiles, smem, kwm, ncta}
extern "C" int ivf_conv_slab_plan(const ivf_conv_desc* d, int sm_count, int* plan) {
  if (!d || !plan) return 0;
  ivf_handle fake;
  fake.sm_count = sm_count;
  if (!ivf_conv3d_slab_eligible(&fake, d)) return 0;
  const ivf_conv_desc dv = slab_view(d);
  d = &dv;
  SlabParams p;
  if (!slab_config(d, sm_count, &p)) return 0;
  plan[0] = p.kch; plan[1] = p.bn; plan[2] = p.ntiles; plan[3] = p.mt; plan[4] = p.th;
  plan[5] = p.acc_stages; plan[6] = p.a_stages; plan[7] = p.b_stages;
  plan[8] = d->n * p.dgroups * ((p.htiles + p.ncta - 1) / p.ncta) * p.ntiles;
  plan[9] = (int)(p.xch_off + 2u * 4u * (uint32_t)p.mt * (uint32_t)p.xch_seg * 4u + (p.est_off ? SLAB_EST_BYTES : 0u));
  plan[10] = p.kwm;
  plan[11] = p.ncta;
  return 1;
}

// depth stacking of the plan ivf_conv_slab_plan reports (1 or 2; 0 when the layer is not served by this kernel)
extern "C" int ivf_conv_slab_plan_ds(const ivf_conv_desc* d, int sm_count) {
  if (!d) return 0;
  ivf_handle fake;
  fake.sm_count = sm_count;
  if (!ivf_conv3d_slab_eligible(&fake, d)) return 0;
  const ivf_conv_desc dv = slab_view(d);
  SlabParams p;
  if (!slab_config(&dv, sm_count, &p)) return 0;
  return p.ds;
}
