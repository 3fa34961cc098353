// libivf.so — handle lifetime, error reporting and the dtype dispatch of ivf_conv3d.
#include "common.cuh"

static thread_local char g_err[1024] = "";

void ivf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* ivf_last_error(void) { return g_err; }

bool ivf_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("IVF_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
extern "C" const char* ivf_version(void) { return "ivf-b200 0.1 (sm_100a)"; }

extern "C" int ivf_create(int device, ivf_handle** out) {
  if (!out) IVF_FAIL(IVF_EINVAL, "ivf_create: out is null");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    IVF_FAIL(IVF_ENOGPU, "ivf_create: no CUDA device (%s); libivf has no CPU fallback",
             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count)
    IVF_FAIL(IVF_EINVAL, "ivf_create: device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  IVF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    IVF_FAIL(IVF_ENOGPU, "ivf_create: device %d is sm_%d%d; libivf is built for sm_100a only",
             device, prop.major, prop.minor);
  ivf_handle* h = new ivf_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->cc_major = prop.major;
  h->cc_minor = prop.minor;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  h->scratch_bytes = 4u << 20;
  cudaError_t me = cudaMalloc(&h->scratch, h->scratch_bytes);
  cudaSetDevice(prev);
  if (me != cudaSuccess) {
    delete h;
    IVF_FAIL(IVF_ECUDA, "ivf_create: scratch allocation failed: %s", cudaGetErrorString(me));
  }
  *out = h;
  return IVF_OK;
}

extern "C" int ivf_destroy(ivf_handle* h) {
  if (h && h->scratch) cudaFree(h->scratch);
  delete h;
  return IVF_OK;
}

extern "C" int64_t ivf_launch_count(const ivf_handle* h) { return h ? h->launches : 0; }

extern "C" int ivf_conv3d(ivf_handle* h, const ivf_conv_desc* d, const void* in, const void* w,
                          const float* scale, const float* shift, const float* acc_in,
                          const void* mask_y, const float* mask_scale, void* out, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && d && in && w && out, "ivf_conv3d: null argument");
  IVF_REQUIRE(d->n > 0 && d->id > 0 && d->ih > 0 && d->iw > 0 && d->od > 0 && d->oh > 0 &&
                  d->ow > 0 && d->cin > 0 && d->cout > 0,
              "ivf_conv3d: non-positive extent");
  IVF_REQUIRE(d->kd > 0 && d->kh > 0 && d->kw > 0 && d->sd > 0 && d->sh > 0 && d->sw > 0,
              "ivf_conv3d: bad kernel/stride");
  IVF_REQUIRE(d->in_ld >= d->in_coff + d->cin && d->out_ld >= d->out_coff + d->cout,
              "ivf_conv3d: channel slice exceeds ld");
  if ((d->flags & IVF_EP_AFFINE)) IVF_REQUIRE(scale && shift, "ivf_conv3d: AFFINE needs scale/shift");
  if ((d->flags & IVF_EP_ACCUM)) IVF_REQUIRE(acc_in, "ivf_conv3d: ACCUM needs acc_in");
  if ((d->flags & IVF_EP_MASK)) IVF_REQUIRE(mask_y && mask_scale, "ivf_conv3d: MASK needs mask_y/mask_scale");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == IVF_F32)
    return ivf_conv3d_f32_launch(h, d, (const float*)in, (const float*)w, scale, shift, acc_in,
                                 (const float*)mask_y, mask_scale, (float*)out, st);
  if (d->dtype == IVF_BF16) {
    if (ivf_conv3d_slab_eligible(h, d))
      return ivf_conv3d_slab_launch(h, d, in, w, scale, shift, acc_in, mask_y, mask_scale, out, st);
    return ivf_conv3d_tc_launch(h, d, in, w, scale, shift, acc_in, mask_y, mask_scale, out, st);
  }
  IVF_FAIL(IVF_EINVAL, "ivf_conv3d: unknown dtype %d", d->dtype);
}

// Two independent convolutions - the two 3x3x3 branches of an Inception module (pt/models/I3D_doubled.py:136-146:
// b1b(b1a(x)) and b2b(b2a(x)) and their data gradients) - issued as ONE launch where the halo-slab kernel can group
// them (bf16, multi-tap, stride 1), otherwise one after the other on the same stream.  Results are those of two
// ivf_conv3d calls (up to the summation order of a different tile plan).
extern "C" int ivf_conv3d_pair(ivf_handle* h, const ivf_conv_desc* d0, const void* in0, const void* w0,
                               const float* scale0, const float* shift0, const float* acc_in0, const void* mask_y0,
                               const float* mask_scale0, void* out0, const ivf_conv_desc* d1, const void* in1,
                               const void* w1, const float* scale1, const float* shift1, const float* acc_in1,
                               const void* mask_y1, const float* mask_scale1, void* out1, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && d0 && d1 && in0 && in1 && w0 && w1 && out0 && out1, "ivf_conv3d_pair: null argument");
  if (d0->dtype == IVF_BF16 && d1->dtype == IVF_BF16 && !d0->transposed && !d1->transposed) {
    const ivf_conv_desc* ds[2] = {d0, d1};
    const ivf_conv_operands op[2] = {{in0, w0, scale0, shift0, acc_in0, mask_y0, mask_scale0, out0},
                                     {in1, w1, scale1, shift1, acc_in1, mask_y1, mask_scale1, out1}};
    bool ok = true;
    for (int i = 0; i < 2 && ok; ++i) {
      const ivf_conv_desc* d = ds[i];
      ok = d->n > 0 && d->id > 0 && d->ih > 0 && d->iw > 0 && d->cin > 0 && d->cout > 0 && d->kd > 0 && d->kh > 0 &&
           d->kw > 0 && d->in_ld >= d->in_coff + d->cin && d->out_ld >= d->out_coff + d->cout &&
           (!(d->flags & IVF_EP_AFFINE) || (op[i].scale && op[i].shift)) && (!(d->flags & IVF_EP_ACCUM) || op[i].acc_in) &&
           (!(d->flags & IVF_EP_MASK) || (op[i].mask_y && op[i].mask_scale));
    }
    if (ok) {
      const int rc = ivf_conv3d_slab_launch_pair(h, ds, op, (cudaStream_t)stream);
      if (rc != IVF_EUNSUPPORTED) return rc;
    }
  }
  int rc = ivf_conv3d(h, d0, in0, w0, scale0, shift0, acc_in0, mask_y0, mask_scale0, out0, stream);
  if (rc) return rc;
  return ivf_conv3d(h, d1, in1, w1, scale1, shift1, acc_in1, mask_y1, mask_scale1, out1, stream);
}

extern "C" int ivf_conv3d_split(ivf_handle* h, const ivf_conv_desc* d, const ivf_conv_split* sp, const void* in,
                                const void* in2, const void* w, const float* scale, const float* shift,
                                const float* acc_in, const void* mask_y, const float* mask_scale, void* out,
                                void* out2, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && d && sp && in && w && out, "ivf_conv3d_split: null argument");
  IVF_REQUIRE(d->dtype == IVF_BF16, "ivf_conv3d_split: bf16 only");
  IVF_REQUIRE(d->kd == 1 && d->kh == 1 && d->kw == 1 && d->sd == 1 && d->sh == 1 && d->sw == 1 && !d->transposed &&
                  d->pd == 0 && d->ph == 0 && d->pw == 0 && d->od == d->id && d->oh == d->ih && d->ow == d->iw,
              "ivf_conv3d_split: 1x1x1 stride-1 convolutions only");
  IVF_REQUIRE(d->n > 0 && d->id > 0 && d->ih > 0 && d->iw > 0 && d->cin > 0 && d->cout > 0,
              "ivf_conv3d_split: non-positive extent");
  IVF_REQUIRE(sp->split_cout >= 0 && sp->split_cout < d->cout && sp->split_cout % 16 == 0,
              "ivf_conv3d_split: split_cout %d must be a multiple of 16 below cout %d", sp->split_cout, d->cout);
  IVF_REQUIRE(sp->split_cin >= 0 && sp->split_cin < d->cin && sp->split_cin % 8 == 0 &&
                  (d->cin - sp->split_cin) % 8 == 0,
              "ivf_conv3d_split: split_cin %d must be a multiple of 8 below cin %d", sp->split_cin, d->cin);
  const int cout1 = sp->split_cout > 0 ? sp->split_cout : d->cout;
  const int cin1 = sp->split_cin > 0 ? sp->split_cin : d->cin;
  IVF_REQUIRE(d->in_ld >= d->in_coff + cin1 && d->out_ld >= d->out_coff + cout1,
              "ivf_conv3d_split: channel slice exceeds ld");
  if (sp->split_cout > 0)
    IVF_REQUIRE(out2 && sp->out2_ld % 8 == 0 && sp->out2_coff % 8 == 0 &&
                    sp->out2_ld >= sp->out2_coff + d->cout - sp->split_cout,
                "ivf_conv3d_split: bad second destination");
  if (sp->split_cin > 0)
    IVF_REQUIRE(in2 && sp->in2_ld % 8 == 0 && sp->in2_coff % 8 == 0 &&
                    sp->in2_ld >= sp->in2_coff + d->cin - sp->split_cin,
                "ivf_conv3d_split: bad second source");
  if ((d->flags & IVF_EP_AFFINE)) IVF_REQUIRE(scale && shift, "ivf_conv3d_split: AFFINE needs scale/shift");
  if ((d->flags & IVF_EP_ACCUM)) IVF_REQUIRE(acc_in, "ivf_conv3d_split: ACCUM needs acc_in");
  if ((d->flags & IVF_EP_MASK)) IVF_REQUIRE(mask_y && mask_scale, "ivf_conv3d_split: MASK needs mask_y/mask_scale");
  if (sp->split_cout > 0)
    IVF_REQUIRE(!(d->flags & (IVF_EP_ACCUM | IVF_EP_MASK)),
                "ivf_conv3d_split: two destinations support the forward epilogue (affine, ReLU) only");
  return ivf_conv3d_tc_launch(h, d, in, w, scale, shift, acc_in, mask_y, mask_scale, out, (cudaStream_t)stream, sp,
                              in2, out2);
}

extern "C" int ivf_conv3d_lstm(ivf_handle* h, const ivf_conv_desc* d0, const void* h_prev, const void* w,
                               const float* pre_x, const float* c_prev, float* c_next, void* h_next, float* gate_act,
                               void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && d0 && h_prev && w && pre_x && c_next && h_next && gate_act, "ivf_conv3d_lstm: null argument");
  ivf_conv_desc d = *d0;
  IVF_REQUIRE(d.dtype == IVF_BF16, "ivf_conv3d_lstm: bf16 only");
  IVF_REQUIRE(d.cout % 16 == 0 && d.out_ld == d.cout && d.out_coff == 0,
              "ivf_conv3d_lstm: cout = 4*hid must be a multiple of 16 and fill the gate rows (out_ld = cout, out_coff = 0)");
  IVF_REQUIRE(((uintptr_t)pre_x | (uintptr_t)c_prev | (uintptr_t)c_next | (uintptr_t)gate_act) % 16 == 0 &&
                  (uintptr_t)h_next % 8 == 0,
              "ivf_conv3d_lstm: state buffers must be 16-byte (h_next: 8-byte) aligned");
  d.flags = IVF_EP_ACCUM | IVF_EP_OUT_F32 | IVF_EP_LSTM;
  IVF_REQUIRE(ivf_conv3d_slab_eligible(h, &d),
              "ivf_conv3d_lstm: the recurrent convolution must be a stride-1 'same' convolution on a map >= 7 wide "
              "(the halo-slab kernel); use ivf_conv3d + ivf_clstm_gates_fwd otherwise");
  return ivf_conv3d_slab_launch(h, &d, h_prev, w, nullptr, nullptr, pre_x, nullptr, nullptr, gate_act,
                                (cudaStream_t)stream, c_prev, c_next, h_next);
}

// Diagnostic: copy the first `bytes` of the handle's scratch buffer to the host (kernel traces written under
// IVF_TC_TRACE=1).  Synchronises the device.
extern "C" int ivf_debug_read_scratch(ivf_handle* h, void* dst, size_t bytes) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && dst && bytes <= h->scratch_bytes, "ivf_debug_read_scratch: bad argument");
  IVF_CUDA(cudaDeviceSynchronize());
  IVF_CUDA(cudaMemcpy(dst, h->scratch, bytes, cudaMemcpyDeviceToHost));
  return IVF_OK;
}
