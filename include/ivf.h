/*
 * ivf.h — C ABI of libivf.so: the B200 (sm_100a) kernels behind the
 * temporal-mask search and Grad-CAM hot path of interpreting-video-features.
 *
 * The reference is 100 % Python on PyTorch (no native code, no FFI), so there
 * is no existing binding to replace: each entry point below names the
 * reference call site (file:line under video_features_pytorch/, "pt/") whose
 * implicit ATen/cuDNN kernels it stands in for.  The reference-side binding a
 * maintainer would add is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless noted.
 *  - the caller owns all buffers; the library keeps only per-handle caches of
 *    TMA tensor maps (freed by ivf_destroy).
 *  - every launch goes to the caller's cudaStream_t (passed as void*), nothing
 *    synchronises, so a sequence of calls can be captured into a CUDA graph.
 *  - return 0 on success, an IVF_E* code otherwise; ivf_last_error() gives the
 *    message (thread local).  There is NO CPU fallback: an unsupported
 *    shape or a missing GPU is an error.
 *  - activations are channels-last (N,D,H,W,C) with an explicit per-pixel
 *    channel stride `ld` and channel offset `coff`, so a branch of an
 *    Inception block writes straight into its slice of the concat buffer
 *    (pt/models/I3D_doubled.py:146 torch.cat is never materialised).
 */
#ifndef IVF_H_
#define IVF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ivf_handle ivf_handle;

enum {
  IVF_OK = 0,
  IVF_EINVAL = 1,       /* bad argument / inconsistent descriptor            */
  IVF_EUNSUPPORTED = 2, /* valid request this build has no kernel for        */
  IVF_ECUDA = 3,        /* CUDA runtime / driver error                       */
  IVF_ENOGPU = 4,       /* no sm_100 device                                  */
  IVF_EWORKSPACE = 5    /* caller-owned workspace missing or too small       */
};

enum { IVF_F32 = 0, IVF_BF16 = 1, IVF_U8 = 2 /* uint8 frames: clip ingest / visualisation only */ };

/* epilogue flags shared by conv / pool-backward / head-backward */
enum {
  IVF_EP_AFFINE = 1,  /* v = v*scale[c] + shift[c]   (eval BatchNorm fold, pt/models/I3D_doubled.py:111) */
  IVF_EP_RELU = 2,    /* v = max(v,0)                (pt/models/I3D_doubled.py:113)                      */
  IVF_EP_ACCUM = 4,   /* v += acc_in[...]  (fp32)    (sum over the consumers of a tensor in backward)    */
  IVF_EP_MASK = 8,    /* v = mask_y[...]>0 ? v*mask_scale[c] : 0  (ReLU'+BN' of the producing Unit3D)     */
  IVF_EP_OUT_F32 = 16,/* store fp32 instead of the activation dtype                                       */
  IVF_EP_LSTM = 32    /* ConvLSTM gates in the epilogue (ivf_conv3d_lstm only)                             */
};

/* ---- lifetime ------------------------------------------------------------ */
int ivf_create(int device, ivf_handle** out);
int ivf_destroy(ivf_handle* h);
const char* ivf_last_error(void);
const char* ivf_version(void);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
int64_t ivf_launch_count(const ivf_handle* h);

/* ---- convolution (pt/models/I3D_doubled.py:83-118 Unit3D.forward and its autograd;
 *      pt/models/convolution_lstm.py:25-32 gate convolutions) ------------------------
 * One generalised gather-GEMM:  out[n,o,c'] = epilogue( sum_{tap,c} in[n, g(o,tap), c] * W[tap,c,c'] )
 *   transposed == 0:  g = o*stride - pad + tap          (forward; also the data-gradient of a
 *                                                        stride-1 conv with flipped weights)
 *   transposed == 1:  g = (o + pad - tap)/stride, taps with a remainder skipped
 *                                                       (data-gradient of a strided conv; fp32 kernel only)
 * Out-of-range gathers read 0 ('same' zero padding, pt/models/I3D_doubled.py:96-106 F.pad).
 *
 * dtype IVF_BF16: tcgen05/TMEM implicit GEMM, TMA-im2col fed, fp32 accumulate; requires
 *   transposed == 0 and stride 1 (strided layers are presented space-to-depth by the host side),
 *   cin/ld/coff multiples of 8.  Weights: bf16 [cout_pad][taps][cin_pad] (K-major) with
 *   cin_pad = ivf_conv_bf16_cin_pad(cin), cout_pad = ivf_conv_bf16_cout_pad(cout).
 * dtype IVF_F32: CUDA-core implicit GEMM (the 1e-4 "fp32 mode"); weights fp32 [taps][cin][cout].
 */
typedef struct ivf_conv_desc {
  int32_t n, id, ih, iw; /* gathered tensor: batch and spatial extent            */
  int32_t od, oh, ow;    /* produced tensor spatial extent                       */
  int32_t cin, cout;     /* channels reduced per tap, channels produced          */
  int32_t kd, kh, kw;
  int32_t sd, sh, sw;
  int32_t pd, ph, pw;    /* front padding                                        */
  int32_t transposed;
  int32_t in_ld, in_coff;
  int32_t out_ld, out_coff;
  int32_t mask_ld, mask_coff;
  int32_t flags;         /* IVF_EP_*                                             */
  int32_t dtype;         /* IVF_F32 | IVF_BF16                                   */
  /* Tile-plan request for the halo-slab kernel, 0 = let the cost model decide: kw taps merged into N,
   * accumulators per tile, TMEM stages (1|2), CTAs per work item (1|2), N tiles.  A request the layer cannot
   * satisfy falls back to the model.  The host side measures a few plans per layer shape once and passes
   * the fastest (interpreting_video_features_b200/tune.py). */
  int32_t plan_kwm, plan_mt, plan_acc, plan_ncta, plan_ntiles;
  int32_t plan_ds;       /* output depths stacked along the MMA's N by the halo-slab kernel (0 = its choice, 1, 2) */
} ivf_conv_desc;

int ivf_conv_bf16_kchunk(int cin);   /* channels per K stage: 16, 32 or 64 */
int ivf_conv_bf16_cin_pad(int cin);  /* cin rounded up to the K stage      */
int ivf_conv_bf16_ntile(int cout);   /* UMMA N of one CTA tile             */
int ivf_conv_bf16_cout_pad(int cout);

/* Diagnostic, needs no GPU: which bf16 kernel a layer gets.  Returns 1 and fills plan[12] =
 * {channels per slab row, N tile, N tiles, accumulators per tile, rows per tile, TMEM stages, slab stages,
 *  weight stages, tiles, dynamic smem bytes, kw taps merged into N, CTAs per work item (2 = cta_group::2 pairs)} when the halo-slab kernel serves it, 0 for the im2col kernel. */
int ivf_conv_slab_plan(const ivf_conv_desc* d, int sm_count, int* plan);
/* depth stacking (plan_ds) of that plan: 1 or 2, 0 when the layer does not go to the halo-slab kernel */
int ivf_conv_slab_plan_ds(const ivf_conv_desc* d, int sm_count);

int ivf_conv3d(ivf_handle* h, const ivf_conv_desc* d, const void* in, const void* w,
               const float* scale, const float* shift, const float* acc_in, const void* mask_y,
               const float* mask_scale, void* out, void* stream);

/* Two independent convolutions as one launch where the halo-slab kernel can group them (bf16, multi-tap, stride 1:
 * the two 3x3x3 branches of an InceptionModule, pt/models/I3D_doubled.py:136-146, forward or data gradient), otherwise
 * issued one after the other.  Same arguments as two ivf_conv3d calls. */
int ivf_conv3d_pair(ivf_handle* h, const ivf_conv_desc* d0, const void* in0, const void* w0, const float* scale0,
                    const float* shift0, const float* acc_in0, const void* mask_y0, const float* mask_scale0,
                    void* out0, const ivf_conv_desc* d1, const void* in1, const void* w1, const float* scale1,
                    const float* shift1, const float* acc_in1, const void* mask_y1, const float* mask_scale1,
                    void* out1, void* stream);

/* 1x1x1 stride-1 bf16 convolution with two destinations and/or two sources: the Inception bottleneck trio
 * b0 | b1a | b2a reads the same x (pt/models/I3D_doubled.py:136-146: three Unit3D calls on one input, b0's
 * result concatenated with the branch outputs), so ONE GEMM produces all three - channels [0, split_cout)
 * land in the concat buffer (out), the rest in the bottleneck buffer (out2) - and ONE data-gradient GEMM
 * reduces over [dz of b0 (in) | dz of the bottlenecks (in2)].  d->cin / d->cout are the totals; in_ld/in_coff
 * and out_ld/out_coff describe the first source / destination.  Weights: K-major [cout_pad][K] with
 * K = round_up(split_cin, 64) + (cin - split_cin), rounded up to 16; the K pad holds zeros.  split_cout must
 * be a multiple of 16.  Either split may be 0 (unused). */
typedef struct ivf_conv_split {
  int32_t split_cout, out2_ld, out2_coff;
  int32_t split_cin, in2_ld, in2_coff;
} ivf_conv_split;
int ivf_conv3d_split(ivf_handle* h, const ivf_conv_desc* d, const ivf_conv_split* sp, const void* in,
                     const void* in2, const void* w, const float* scale, const float* shift,
                     const float* acc_in, const void* mask_y, const float* mask_scale, void* out, void* out2,
                     void* stream);

/* ---- model load: weight packing and BatchNorm folding -------------------------------
 * ivf_pack_weights writes ONE source weight tensor (fp32 OIDHW on the device: the nn.Conv3d / nn.Conv2d
 * parameter of pt/models/I3D_doubled.py:66-72, pt/models/convolution_lstm.py:25-32, kd = 1 for 2-D) into the
 * operand matrix a convolution kernel reads, as the block starting at row n_off / reduction index k_off of a
 * packed matrix of n_pad rows and k_pad reduction entries per tap.  Several sources may fill one matrix (the
 * fused b0|b1a|b2a GEMM, the four ConvLSTM gates): the first call sets zero_first.
 *   dgrad == 0: rows = output channels, K = operand channels (the forward operand);
 *   dgrad == 1: rows = operand channels, K = output channels, taps flipped (data gradient as a stride-1
 *               convolution with flipped weights: the bf16 kernels);
 *   dgrad == 2: rows = operand channels, K = output channels, taps as stored (the fp32 kernel's transposed
 *               gather, ivf_conv_desc.transposed = 1).
 *   s2d_{d,h,w} == 2: the axis has stride 2 and is presented space-to-depth: operand channel =
 *     ((a*s2d_h + b)*s2d_w + c)*ci_stride + ch, operand tap = ceil(k/2) per axis, source tap = 2*tap + parity
 *     (taps beyond the kernel are zero).  ci_stride >= ci pads each parity block (0 = ci).
 *   layout IVF_PACK_KMAJOR:   dst[n_pad][taps][k_pad]  (bf16 tcgen05 kernels; n_pad = ivf_conv_bf16_cout_pad,
 *                             k_pad = ivf_conv_bf16_cin_pad of the totals)
 *          IVF_PACK_TAPMAJOR: dst[taps][k_pad][n_pad]  (fp32 kernel)
 *   dtype: element type of dst (IVF_BF16 | IVF_F32).                                                     */
enum { IVF_PACK_KMAJOR = 0, IVF_PACK_TAPMAJOR = 1 };
typedef struct ivf_pack_desc {
  int32_t co, ci, kd, kh, kw;
  int32_t ci_stride;
  int32_t s2d_d, s2d_h, s2d_w;
  int32_t dgrad;
  int32_t layout;
  int32_t dtype;
  int32_t n_pad, k_pad, n_off, k_off;
  int32_t zero_first;
  int32_t n_stride, k_stride; /* this source's rows / reduction entries land every n_stride / k_stride-th index
                               * (0 = 1): the unit-major gate interleave of the fused ConvLSTM epilogue */
} ivf_pack_desc;
int ivf_pack_weights(ivf_handle* h, const ivf_pack_desc* d, const float* src, void* dst, void* stream);
/* scale = gamma/sqrt(var+eps), shift = beta - mean*scale (+ scale*conv_bias): eval BatchNorm folded into the
 * conv epilogue (pt/models/I3D_doubled.py:75,111 eps 1e-3; pt/models/convolution_lstm.py:85 eps 1e-5).
 * gamma == NULL: scale 1, shift 0 (+ conv_bias) for units without BatchNorm (the logits layer).          */
int ivf_bn_fold(ivf_handle* h, const float* gamma, const float* beta, const float* mean, const float* var,
                float eps, const float* conv_bias, int c, float* scale, float* shift, void* stream);
/* dst[0 .. bytes) = the 32-bit pattern (buffer initialisation on the caller's stream; bytes % 4 == 0). */
int ivf_fill_u32(ivf_handle* h, void* dst, size_t bytes, uint32_t pattern, void* stream);
/* dst[i] = (float)src[i]: uint8 frames as the loaders decode them (pt/data_loader_jpg.py:27-37,
 * pt/data_loader_kth.py:20-43 produce 0..255 values) converted on the device, so a clip crosses PCIe as one
 * byte per value instead of four. */
int ivf_u8_to_f32(ivf_handle* h, const uint8_t* src, float* dst, size_t count, void* stream);
/* out[n][ncls] = one_hot(targets[n]) — the class selector of pt/FindMasksComparison_I3D_smth.py:205 and
 * pt/grad_cam_videos.py:73-79 as the upstream gradient of the head. */
int ivf_one_hot(ivf_handle* h, const int* targets, int n, int ncls, float* out, void* stream);
/* out[n] = index of the first maximum of row n (np.argmax of pt/grad_cam_videos.py:70-71), on the device. */
int ivf_argmax_rows(ivf_handle* h, const float* x, int n, int ncls, int* out, void* stream);

/* Diagnostic: copy the head of the handle's scratch buffer to the host after a device synchronise (kernel
 * phase traces written when IVF_TC_TRACE=1). */
int ivf_debug_read_scratch(ivf_handle* h, void* dst, size_t bytes);

/* ---- max-pool with TF-'same' ZERO padding (pt/models/I3D_doubled.py:8-40;
 *      nn.MaxPool2d of pt/models/convolution_lstm.py:79 with pad 0) ------------------
 * argmax: uint8 [n*od*oh*ow][c] window-scan index of the first maximum (ATen tie rule),
 * padded positions take part with value 0 exactly as F.pad + MaxPool3d does.           */
typedef struct ivf_pool_desc {
  int32_t n, id, ih, iw, c;
  int32_t od, oh, ow;
  int32_t kd, kh, kw;
  int32_t sd, sh, sw;
  int32_t pd, ph, pw;
  int32_t in_ld, in_coff;
  int32_t out_ld, out_coff;
  int32_t mask_ld, mask_coff; /* backward epilogue: mask tensor indexed like the pool INPUT */
  int32_t flags;              /* backward: IVF_EP_ACCUM | IVF_EP_MASK | IVF_EP_OUT_F32; forward: IVF_POOL_NONNEG */
  int32_t dtype;
} ivf_pool_desc;
/* forward flag: the caller vouches that no input element is negative (the input is a ReLU output, as at every
 * max-pool of pt/models/I3D_doubled.py:351-380): the packed bf16 kernels then compare bit patterns directly.
 * With the flag set and a negative input the result is unspecified. */
#define IVF_POOL_NONNEG 64

int ivf_maxpool3d_fwd(ivf_handle* h, const ivf_pool_desc* d, const void* in, void* out,
                      uint8_t* argmax, void* stream);
/* dx (at in_ld/in_coff) = epilogue( sum of dy over the windows whose argmax is this element ) */
int ivf_maxpool3d_bwd(ivf_handle* h, const ivf_pool_desc* d, const void* dy, const uint8_t* argmax,
                      const float* acc_in, const void* mask_y, const float* mask_scale, void* dx,
                      void* stream);

/* The same pair with a ReLU' bit mask: the forward also writes relu_bits[input pixel][c/8] (bit i of a byte =
 * element 8*(c/8)+i > 0; bf16, c % 8 == 0), and the backward's MASK epilogue reads that byte instead of the
 * 16 bytes of mask_y where its kernel supports it (the stride-2 pools; mask_y stays the fallback).  For the
 * stage pools of I3D, whose input is a Unit3D / Inception output (pt/models/I3D_doubled.py:244-290), this
 * removes ~230 MB of reads per 8-clip iteration.  relu_bits == NULL is the plain call. */
int ivf_maxpool3d_fwd_bits(ivf_handle* h, const ivf_pool_desc* d, const void* in, void* out, uint8_t* argmax,
                           uint8_t* relu_bits, void* stream);
int ivf_maxpool3d_bwd_bits(ivf_handle* h, const ivf_pool_desc* d, const void* dy, const uint8_t* argmax,
                           const float* acc_in, const void* mask_y, const uint8_t* relu_bits,
                           const float* mask_scale, void* dx, void* stream);

/* ---- I3D head (pt/models/I3D_doubled.py:360-371: avg_pool -> dropout(eval) -> 1x1x1 logits
 *      with bias -> squeeze -> softmax(dim=1)) ---------------------------------------
 * feat: [n][p][c] channels-last (p = all pooled positions; the pool must cover the whole map).
 * out: fp32 [n][ncls] probabilities (softmax != 0) or logits.  With p == 1 this is the
 * Linear(+Softmax) classifier of the ConvLSTM model (pt/models/CLSTM_4.py:78-83).      */
/* workspace: caller-owned fp32 buffer of ivf_i3d_head_workspace_bytes(n, c, ncls) bytes that holds the
 * per-chunk partial logits between the two launches of the forward.  It belongs to the CALLER (one per
 * engine), not to the handle: two engines driven from one thread on different streams (clip groups on
 * parallel CUDA-graph branches) must not share it. */
size_t ivf_i3d_head_workspace_bytes(int n, int c, int ncls);
int ivf_i3d_head_fwd(ivf_handle* h, int dtype, const void* feat, int n, int p, int c, int ld,
                     const float* w, const float* b, int ncls, int softmax, float* logits,
                     float* out, float* workspace, size_t workspace_bytes, void* stream);
/* dfeat[n][p][c] = epilogue( (1/p) * W^T * dlogits ), dlogits from dout through softmax'. */
int ivf_i3d_head_bwd(ivf_handle* h, int dtype, int n, int p, int c, int ld, const float* w,
                     int ncls, int softmax, const float* out, const float* dout, int flags,
                     const void* mask_y, int mask_ld, int mask_coff, const float* mask_scale,
                     void* dfeat, void* stream);

/* ---- temporal perturbation (pt/mask.py:4-56 perturb_sequence) ----------------------
 * x: fp32 [b][c][t][h][w] (the loader's layout, pt/data_loader_jpg.py:27-37).
 * mask: fp32, row `i*mask_bstride` for clip i (mask_bstride 0 = one mask for the batch,
 *       the reference's semantics).  mode 0 = 'freeze' (first-order recurrence over t),
 *       1 = 'reverse' (swap-blend inside each run of mask > 0.1, pt/mask.py:60-85).
 * out_fmt: IVF_PFMT_NCDHW_F32  fp32 [b][c][t][h][w]          (drop-in perturb_sequence result)
 *          IVF_PFMT_NDHWC_F32  fp32 [b][t][h][w][c]          (fp32 conv path)
 *          IVF_PFMT_S2D_BF16   bf16 [b][t/2][h/2][w/2][32]   (space-to-depth operand of the
 *              stride-2 7x7x7 stem, channel = ((dt*2+dh)*2+dw)*c + ch, 24..31 zero)
 *          IVF_PFMT_TBHWC_F32  fp32 [t][b][h][w][c]          (time-major frames: ConvLSTM, fp32 path)
 *          IVF_PFMT_S2D2_BF16  bf16 [t][b][h/2][w/2][16]     (time-major, 2-D space-to-depth operand of
 *              the stride-2 5x5 ConvLSTM x-convolution, channel = (dh*2+dw)*c + ch, 12..15 zero)
 */
enum {
  IVF_PFMT_NCDHW_F32 = 0,
  IVF_PFMT_NDHWC_F32 = 1,
  IVF_PFMT_S2D_BF16 = 2,
  IVF_PFMT_TBHWC_F32 = 3,
  IVF_PFMT_S2D2_BF16 = 4
};
int ivf_perturb_fwd(ivf_handle* h, int mode, const float* x, const float* mask, int mask_bstride,
                    int b, int c, int t, int hh, int ww, int out_fmt, void* out, void* stream);
/* dmask[b][t] (fp32, one row per clip) = d<gout, P>/dmask; gout is laid out like `out` above
 * (fp32 for the two F32 formats; for IVF_PFMT_S2D_BF16 gout is bf16 or fp32 [b][t/2][h/2][w/2][32]
 * selected by gout_dtype).                                                              */
int ivf_perturb_bwd(ivf_handle* h, int mode, const float* x, const float* mask, int mask_bstride,
                    int b, int c, int t, int hh, int ww, int out_fmt, int gout_dtype,
                    const void* gout, float* dmask, void* stream);

/* ---- mask objective + optimiser (pt/mask.py:88-100 calc_tv_norm;
 *      pt/FindMasksComparison_I3D_smth.py:191-214: sigmoid, L1, TV(p=q=3), Adam lr) ----
 * One launch per iteration for nclip independent masks:
 *   s = sigmoid(m); g = (lam1*sign(s) + lam2*dTV(s) + dclass) * s(1-s); Adam(m, g).
 * losses[nclip][3] = {lam1*L1, lam2*TV, sum}.  sig_out = sigmoid of the UPDATED m.
 * Adam's step number is `step` (1-based), or, when step_dev != NULL, the per-clip device
 * counter step_dev[clip]+1 which the kernel then stores back (CUDA-graph replay safe).
 * A constant mask makes dTV NaN exactly as the reference's pow chain does (pt/mask.py:163-165). */
int ivf_mask_loss_adam(ivf_handle* h, float* m, float* exp_avg, float* exp_avg_sq,
                       const float* dclass, int nclip, int t, int step, int* step_dev, float lam1,
                       float lam2,
                       float lr, float beta1, float beta2, float eps, float* losses,
                       float* sig_out, void* stream);
int ivf_sigmoid(ivf_handle* h, const float* m, float* out, int count, void* stream);
/* out[i] = probs[i][targets[i]] — the class score the drivers read after every forward
 * (pt/mask.py:128-129,140-143; pt/FindMasksComparison_I3D_smth.py:205), kept on the device. */
int ivf_select_scores(ivf_handle* h, const float* probs, const int* targets, int n, int ncls, float* out,
                      void* stream);
/* 'central' mask initialisation of pt/mask.py:121-154 for n clips from the class scores of every candidate:
 * scores is fp32 [1 + t/2][n] - row 0 the unperturbed clip, row 1 the fully frozen clip, row 1+i the centred
 * window with i frames switched off at both ends (i = 1 .. t/2-1).  raw[n][t] receives -5 / +5; chosen[n]
 * (optional) the selected i.  No host read-back: the search that follows is queued behind it. */
int ivf_init_mask_select(ivf_handle* h, const float* scores, int n, int t, float threshold, float* raw,
                         int* chosen, void* stream);
/* general-(p,q) TV norm used by the drop-in calc_tv_norm: val[0] and dval/dmask[t] */
int ivf_tv_norm(ivf_handle* h, const float* mask, int t, float p, float q, float* val,
                float* dmask, void* stream);

/* ---- Grad-CAM tail (pt/grad_cam_videos.py:85-140) -----------------------------------
 * act, grad: [n][tp][hp][wp][c] channels-last activations of the target layer and the
 * gradient of the class score w.r.t. them.  cam: fp32 [n][tp*step][hout][wout]:
 *   w_k = mean_{t,h,w} grad ; cam = relu(sum_k w_k act_k) ; bilinear (cv2 INTER_LINEAR,
 *   half-pixel) to hout x wout ; repeated `step` times along t ; min/max normalised per
 *   feature-time slice (per_frame != 0) or per clip.  One fused kernel.
 * cam_lowres (optional): fp32 [n][tp][hp][wp], the map before upsampling and normalisation (what a multi-GPU
 * job gathers); cam may be NULL when only that is wanted.                      */
int ivf_gradcam(ivf_handle* h, int act_dtype, int grad_dtype, const void* act, const void* grad,
                int n, int tp, int hp, int wp, int c, int ld, int step, int hout, int wout,
                int per_frame, float* cam, float* cam_lowres, void* stream);

/* ---- ConvLSTM (pt/models/convolution_lstm.py:38-60 cell, :96-132 stack) --------------
 * pre: fp32 [m][4*hid] gate pre-activations = x-conv(+bias) + h-conv (summed by the conv
 * epilogue, IVF_EP_ACCUM), gate order i,f,c,o along the channel axis; fused:
 *   i=s(.) f=s(.) c'=f*c+i*tanh(.) o=s(.) h'=o*tanh(c')      (zero peepholes, :52-54)
 * c_prev may be NULL (step 0, zero state).  gate_act (fp32 [m][4*hid]) keeps the activated
 * gates for the backward pass.                                                          */
/* unit_major != 0: pre / gate_act / dgates rows are [i0 f0 c0 o0 i1 f1 c1 o1 ...] instead of [i.. | f.. | c.. | o..]
 * (bf16 path, hid % 4 == 0): the channel order of the fused recurrent convolution below. */
int ivf_clstm_gates_fwd(ivf_handle* h, int dtype, const float* pre, const float* c_prev, int m,
                        int hid, float* c_next, void* h_next, float* gate_act, int unit_major, void* stream);
/* BPTT step: dh = dL/dh' (fp32 [m][hid]); dc_io holds dL/dc' carried from step t+1 on entry
 * and dL/dc for step t-1 on exit; dgates ([m][4*hid], activation dtype) = dL/dpre.      */
int ivf_clstm_gates_bwd(ivf_handle* h, int dtype, const float* gate_act, const float* c_prev,
                        const float* c_next, const float* dh, float* dc_io, int m, int hid,
                        void* dgates, int unit_major, void* stream);
/* The recurrent step as ONE kernel (pt/models/convolution_lstm.py:38-48): the h-convolution (bf16 tcgen05
 * implicit GEMM, stride 1, 'same') with the gates applied in its epilogue straight from the TMEM accumulator:
 *   pre = pre_x[pix] + conv(h_prev, Wh)      (pre_x: the x-convolution of this step incl. bias, fp32 [pix][4*hid])
 *   i,f,o = sigmoid, g = tanh;  c' = f*c_prev + i*g;  h' = o*tanh(c')
 * Output channels are UNIT-MAJOR (4*k + gate, weights packed with n_stride = 4): a 16-column accumulator chunk
 * holds four whole hidden units.  d describes the convolution (cout = 4*hid, out_ld = 4*hid, out_coff = 0);
 * gate_act (fp32 [pix][4*hid], unit-major) keeps the activated gates for the backward pass; c_prev / c_next fp32
 * [pix][hid]; h_next bf16 [pix][hid].  Replaces a convolution launch + a gate launch and the fp32 round trip of the
 * pre-activations between them. */
int ivf_conv3d_lstm(ivf_handle* h, const ivf_conv_desc* d, const void* h_prev, const void* w, const float* pre_x,
                    const float* c_prev, float* c_next, void* h_next, float* gate_act, void* stream);
/* eval BatchNorm2d affine + MaxPool2d(2) (pt/models/convolution_lstm.py:120-124);
 * x [n][hh][ww][c] -> y [n][hh/2][ww/2][c]; backward returns fp32 dx (+ acc_in if given).
 * s2d != 0: y (and dy) use the 2-D space-to-depth layout [n][hh/4][ww/4][4c] that the next layer's
 * stride-2 x-convolution reads as a stride-1 3x3 convolution (channel = (dy*2+dx)*c + k).  */
int ivf_bn_pool2d_fwd(ivf_handle* h, int dtype, const void* x, int n, int hh, int ww, int c,
                      const float* scale, const float* shift, void* y, uint8_t* argmax, int s2d,
                      void* stream);
int ivf_bn_pool2d_bwd(ivf_handle* h, int dtype, const void* dy, const uint8_t* argmax, int n, int hh,
                      int ww, int c, const float* scale, const float* acc_in, float* dx, int s2d,
                      void* stream);

/* ---- visualisation (pt/visualisation.py:96-130 create_image_arrays, :35-93 mask dots) ---------------------
 * One clip: clip [3][t][hh][ww] RGB 0..255 (fp32 or uint8), cam fp32 [t][hh][ww] in [0,1] (NaN allowed: an
 * all-zero Grad-CAM slice), pert fp32 [3][t][hh][ww] (the clip under the snapped mask).  out: uint8
 * [t][hh][3*ww][3] BGR - per frame [ frame | uint8(255*(JET(uint8(255*cam)) + frame)/max) | perturbed frame ],
 * bit-exact with the host pipeline of the reference.  draw_dots != 0 also draws the temporal-mask dots under
 * the third panel (mask fp32 [t], rounded at 0.5 as roundUpMask=True does).  Encoding stays on the host. */
int ivf_viz_triptych(ivf_handle* h, int clip_dtype, const void* clip, const float* cam, const float* pert,
                     const float* mask, int t, int hh, int ww, int draw_dots, uint8_t* out, void* stream);

/* ---- training step (pt/train_i3d_smth.py:192-250: model.train(), forward, CrossEntropyLoss, backward, step) -----
 * What the training path adds to the interpretation path: BatchNorm3d with batch statistics
 * (pt/models/I3D_doubled.py:75: eps 1e-3, momentum 0.01) forward and backward, the convolution weight gradient,
 * the classifier head with dropout + cross-entropy, and the optimizer update.  Tensors are channels-last
 * [rows][ld] with a channel offset (dtype IVF_F32 | IVF_BF16); statistics and parameter gradients are fp32.
 *
 * ivf_bn_train_fwd: y = relu?(gamma * (z - mean_batch) * rstd_batch + beta) over m rows of c channels; writes
 *   save_mean / save_rstd (fp32 [c]) for the backward pass and, when running_mean is given, updates the running
 *   statistics in place as nn.BatchNorm3d does (running_var with the unbiased batch variance).  ws: 2*c doubles.
 * ivf_bn_train_bwd: g = dy * [y > 0] (y NULL: no ReLU); dgamma = sum g*xhat, dbeta = sum g,
 *   dz = gamma * rstd * (g - mean(g) - xhat * mean(g*xhat)).  dz may alias dy or z.  ws: 2*c doubles.  dy_dtype: the
 *   element type of dy - IVF_F32 (gradients summed over the consumers of a tensor are kept in fp32 by the
 *   mixed-precision step) or the activation type.                                                            */
int ivf_bn_train_fwd(ivf_handle* h, int dtype, const void* z, int z_ld, int z_coff, long long m, int c,
                     const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                     float* running_var, float* save_mean, float* save_rstd, double* ws, void* y, int y_ld,
                     int y_coff, int relu, void* stream);
int ivf_bn_train_bwd(ivf_handle* h, int dtype, int dy_dtype, const void* dy, int dy_ld, int dy_coff, const void* y,
                     int y_ld, int y_coff, const void* z, int z_ld, int z_coff, long long m, int c, const float* gamma,
                     const float* save_mean, const float* save_rstd, double* ws, void* dz, int dz_ld, int dz_coff,
                     float* dgamma, float* dbeta, void* stream);
/* Weight gradient of the convolution d describes (autograd's convolution_backward, weight part, of
 * pt/models/I3D_doubled.py:109-113): x = the convolution's input (d->in_ld / in_coff), dz = the gradient w.r.t. its
 * output (d->out_ld / out_coff), dw = fp32 [cout][cin][kd][kh][kw] (the nn.Conv3d parameter's layout), overwritten.
 * d->pd/ph/pw are the front pads of the 'same' padding, as for ivf_conv3d; d->transposed must be 0.  d->dtype is
 * the element type of dz, x_dtype that of x (equal, or fp32 x with bf16 dz: the stem, whose tensor-core forward
 * reads a space-to-depth copy while the weight gradient reads the clip itself).                              */
int ivf_conv3d_wgrad(ivf_handle* h, const ivf_conv_desc* d, int x_dtype, const void* x, const void* dz, float* dw,
                     void* stream);
/* The same for a stride-2 layer whose tensor-core form is a stride-1 convolution over the 2x2x2 space-to-depth
 * record of its input (the stem, pt/models/I3D_doubled.py:233-235; the operand ivf_perturb_fwd writes as
 * IVF_PFMT_S2D_BF16 and ivf_pack_weights packs for with s2d = 2): d describes THAT convolution (bf16, cin = 8 * ci,
 * kernel = ceil(k / 2), stride 1), dw is written in the original [cout][ci][kd][kh][kw] layout.               */
int ivf_conv3d_wgrad_s2d(ivf_handle* h, const ivf_conv_desc* d, const void* x, const void* dz, float* dw, int ci,
                         int kd, int kh, int kw, void* stream);
/* Classifier head in training mode (pt/models/I3D_doubled.py:360-371 + nn.CrossEntropyLoss,
 * pt/train_i3d_smth.py:124-127): pooled = mean over the pix positions of a clip's feature map (the average
 * pool's window must cover the map) * drop (fp32 [batch][c] dropout mask already scaled by 1/keep, NULL: none);
 * logits = pooled . w^T + bias (w fp32 [classes][c]); loss = mean over the batch of -log softmax(logits)[target];
 * dlogits = dloss/dlogits.  The backward call returns dw, db and writes d feat (the same value at every position). */
int ivf_head_train_fwd(ivf_handle* h, int dtype, const void* feat, int ld, int coff, int batch, int pix, int c,
                       const float* drop, const float* w, const float* bias, const int* target, int classes,
                       float* pooled, float* logits, float* dlogits, float* loss, void* stream);
int ivf_head_train_bwd(ivf_handle* h, int dtype, const float* dlogits, const float* pooled, const float* drop,
                       const float* w, int batch, int pix, int c, int classes, float* dw, float* db, void* dfeat,
                       int ld, int coff, void* stream);
/* out[n] = 0 with probability p, else 1/(1-p): the scaled mask of nn.Dropout(p) (pt/models/I3D_doubled.py:319) from a
 * counter-based generator (seed + element index); the `drop` operand of the head calls above.               */
int ivf_dropout_mask(ivf_handle* h, float* out, long long n, float p, unsigned long long seed, void* stream);
/* In-place parameter update, kind 0 = torch.optim.SGD(momentum = beta1, weight_decay; s1 = momentum buffer, may be
 * NULL when beta1 == 0), kind 1 = torch.optim.Adam(betas, eps, weight_decay as L2; s1 / s2 = first / second moment);
 * step counts from 1 (pt/train_i3d_smth.py:131-138).                                                          */
int ivf_optim_step(ivf_handle* h, int kind, float* p, const float* g, float* s1, float* s2, long long n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, void* stream);

/* The same update for many tensors in ONE launch.  table (device): rows of five 64-bit words {p, g, s1, s2, n} -
 * device pointers of a parameter chunk, its gradient and its two state buffers (unused ones may be 0) and the
 * chunk's element count; one thread block per row, so long tensors are cut into chunks of a few thousand elements
 * by the caller.  grad_scale multiplies the gradient before weight decay (1 / world size after an all-reduce). */
int ivf_optim_step_multi(ivf_handle* h, int kind, const void* table, int rows, float lr, float beta1, float beta2,
                         float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ---- bring-up probes (tests only) ---------------------------------------------------
 * Loads one 128-pixel x kchunk im2col TMA tile exactly as the conv kernel does and
 * copies the shared-memory image (de-swizzled, [128][kchunk] bf16) to `tile_out`.      */
int ivf_probe_im2col(ivf_handle* h, const ivf_conv_desc* d, const void* in, int m0, int tap,
                     int c0, void* tile_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IVF_H_ */
