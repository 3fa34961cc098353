// GPU-side visualisation: the image triptych the drivers write per frame -
// [ input frame | Grad-CAM heat map blended over the frame | frame of the perturbed clip under the snapped mask ] -
// and the temporal-mask dots drawn under the third panel, as ONE kernel per clip instead of the reference's
// per-frame host loop (pt/visualisation.py:96-130 create_image_arrays: cv2.applyColorMap(uint8(255*cam), JET),
// heat + frame, divide by the frame's maximum, np.concatenate; :35-64 vizualize_results_on_gradcam and :67-93
// find_temp_mask_red_dots for the dots).  Encoding to JPEG/PNG/GIF stays on the host: it is file output.
// Output: uint8 [T][H][3W][3] in BGR order (what cv2.imwrite expects), bit-exact with the host pipeline.
#include "common.cuh"

namespace {

// cv2.COLORMAP_JET as BGR triples (generated once with cv2.applyColorMap(arange(256)), OpenCV 4.x)
__constant__ unsigned char kJet[256][3] = {
    {128,0,0}, {132,0,0}, {136,0,0}, {140,0,0}, {144,0,0}, {148,0,0}, {152,0,0}, {156,0,0},
    {160,0,0}, {164,0,0}, {168,0,0}, {172,0,0}, {176,0,0}, {180,0,0}, {184,0,0}, {188,0,0},
    {192,0,0}, {196,0,0}, {200,0,0}, {204,0,0}, {208,0,0}, {212,0,0}, {216,0,0}, {220,0,0},
    {224,0,0}, {228,0,0}, {232,0,0}, {236,0,0}, {240,0,0}, {244,0,0}, {248,0,0}, {252,0,0},
    {255,0,0}, {255,4,0}, {255,8,0}, {255,12,0}, {255,16,0}, {255,20,0}, {255,24,0}, {255,28,0},
    {255,32,0}, {255,36,0}, {255,40,0}, {255,44,0}, {255,48,0}, {255,52,0}, {255,56,0}, {255,60,0},
    {255,64,0}, {255,68,0}, {255,72,0}, {255,76,0}, {255,80,0}, {255,84,0}, {255,88,0}, {255,92,0},
    {255,96,0}, {255,100,0}, {255,104,0}, {255,108,0}, {255,112,0}, {255,116,0}, {255,120,0}, {255,124,0},
    {255,128,0}, {255,132,0}, {255,136,0}, {255,140,0}, {255,144,0}, {255,148,0}, {255,152,0}, {255,156,0},
    {255,160,0}, {255,164,0}, {255,168,0}, {255,172,0}, {255,176,0}, {255,180,0}, {255,184,0}, {255,188,0},
    {255,192,0}, {255,196,0}, {255,200,0}, {255,204,0}, {255,208,0}, {255,212,0}, {255,216,0}, {255,220,0},
    {255,224,0}, {255,228,0}, {255,232,0}, {255,236,0}, {255,240,0}, {255,244,0}, {255,248,0}, {255,252,0},
    {254,255,2}, {250,255,6}, {246,255,10}, {242,255,14}, {238,255,18}, {234,255,22}, {230,255,26}, {226,255,30},
    {222,255,34}, {218,255,38}, {214,255,42}, {210,255,46}, {206,255,50}, {202,255,54}, {198,255,58}, {194,255,62},
    {190,255,66}, {186,255,70}, {182,255,74}, {178,255,78}, {174,255,82}, {170,255,86}, {166,255,90}, {162,255,94},
    {158,255,98}, {154,255,102}, {150,255,106}, {146,255,110}, {142,255,114}, {138,255,118}, {134,255,122}, {130,255,126},
    {126,255,130}, {122,255,134}, {118,255,138}, {114,255,142}, {110,255,146}, {106,255,150}, {102,255,154}, {98,255,158},
    {94,255,162}, {90,255,166}, {86,255,170}, {82,255,174}, {78,255,178}, {74,255,182}, {70,255,186}, {66,255,190},
    {62,255,194}, {58,255,198}, {54,255,202}, {50,255,206}, {46,255,210}, {42,255,214}, {38,255,218}, {34,255,222},
    {30,255,226}, {26,255,230}, {22,255,234}, {18,255,238}, {14,255,242}, {10,255,246}, {6,255,250}, {1,255,254},
    {0,252,255}, {0,248,255}, {0,244,255}, {0,240,255}, {0,236,255}, {0,232,255}, {0,228,255}, {0,224,255},
    {0,220,255}, {0,216,255}, {0,212,255}, {0,208,255}, {0,204,255}, {0,200,255}, {0,196,255}, {0,192,255},
    {0,188,255}, {0,184,255}, {0,180,255}, {0,176,255}, {0,172,255}, {0,168,255}, {0,164,255}, {0,160,255},
    {0,156,255}, {0,152,255}, {0,148,255}, {0,144,255}, {0,140,255}, {0,136,255}, {0,132,255}, {0,128,255},
    {0,124,255}, {0,120,255}, {0,116,255}, {0,112,255}, {0,108,255}, {0,104,255}, {0,100,255}, {0,96,255},
    {0,92,255}, {0,88,255}, {0,84,255}, {0,80,255}, {0,76,255}, {0,72,255}, {0,68,255}, {0,64,255},
    {0,60,255}, {0,56,255}, {0,52,255}, {0,48,255}, {0,44,255}, {0,40,255}, {0,36,255}, {0,32,255},
    {0,28,255}, {0,24,255}, {0,20,255}, {0,16,255}, {0,12,255}, {0,8,255}, {0,4,255}, {0,0,255},
    {0,0,252}, {0,0,248}, {0,0,244}, {0,0,240}, {0,0,236}, {0,0,232}, {0,0,228}, {0,0,224},
    {0,0,220}, {0,0,216}, {0,0,212}, {0,0,208}, {0,0,204}, {0,0,200}, {0,0,196}, {0,0,192},
    {0,0,188}, {0,0,184}, {0,0,180}, {0,0,176}, {0,0,172}, {0,0,168}, {0,0,164}, {0,0,160},
    {0,0,156}, {0,0,152}, {0,0,148}, {0,0,144}, {0,0,140}, {0,0,136}, {0,0,132}, {0,0,128},
};

__device__ __forceinline__ unsigned char to_u8(float v) {
  // np.uint8(x) on a float array: C conversion, truncation toward zero (values here are within 0..255;
  // NaN converts to 0 on x86, which is what the reference's numpy does for an all-NaN Grad-CAM slice)
  if (!(v == v)) return 0;
  int i = (int)v;
  return (unsigned char)(i & 0xff);
}

template <typename TX>
__global__ void __launch_bounds__(256)
viz_triptych_kernel(const TX* __restrict__ clip, const float* __restrict__ cam, const float* __restrict__ pert,
                    int t, int hh, int ww, unsigned char* __restrict__ out) {
  __shared__ float red[8];
  __shared__ float fmax_s;
  const int f = blockIdx.x;
  const size_t plane = (size_t)hh * ww, cstride = (size_t)t * plane;
  const TX* x = clip + (size_t)f * plane;       // channel c at x[c*cstride + p] (RGB)
  const float* pz = pert + (size_t)f * plane;
  const float* cm = cam + (size_t)f * plane;
  // pass 1: the frame's maximum of heat + frame over all pixels and channels (cam = cam / np.max(cam))
  float mx = -INFINITY;
  for (int p = threadIdx.x; p < (int)plane; p += blockDim.x) {
    const unsigned char idx = to_u8(255.f * cm[p]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {  // BGR channel c is RGB channel 2-c
      const float v = (float)kJet[idx][c] + (float)x[(size_t)(2 - c) * cstride + p];
      mx = fmaxf(mx, v);  // np.max propagates NaN; the inputs here are finite (uint8-valued frames)
    }
  }
  mx = ivf_warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
    fmax_s = m;
  }
  __syncthreads();
  const float fm = fmax_s;
  // pass 2: the three panels
  unsigned char* o = out + (size_t)f * hh * (3 * ww) * 3;
  for (int p = threadIdx.x; p < (int)plane; p += blockDim.x) {
    const int y = p / ww, xx = p - y * ww;
    const unsigned char idx = to_u8(255.f * cm[p]);
    unsigned char* row = o + ((size_t)y * (3 * ww)) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float img = (float)x[(size_t)(2 - c) * cstride + p];
      const float blend = ((float)kJet[idx][c] + img) / fm;
      row[(size_t)xx * 3 + c] = to_u8(img);
      row[(size_t)(ww + xx) * 3 + c] = to_u8(255.f * blend);
      row[(size_t)(2 * ww + xx) * 3 + c] = to_u8(pz[(size_t)(2 - c) * cstride + p]);
    }
  }
}

// the temporal-mask dots of pt/visualisation.py:67-93 drawn into frame f: dot j spans the last dot_h rows and
// columns [2W + j*(dot_w+pad), +dot_w); channel 1 (green, BGR index 1) when mask[j] == 0, channel 2 (red) when 1;
// intensity 255 for the frame's own dot, 150 otherwise.
__global__ void viz_dots_kernel(unsigned char* __restrict__ img, const float* __restrict__ mask, int t, int hh, int ww,
                                int round_up) {
  const int f = blockIdx.x;
  const int dot_w = ww / (t + 4);
  const int pad = (ww - dot_w * t) / t;
  const int dot_h = hh / 20;
  unsigned char* o = img + (size_t)f * hh * (3 * ww) * 3;
  for (int i = threadIdx.x; i < t * dot_h * dot_w; i += blockDim.x) {
    const int j = i / (dot_h * dot_w), r = i - j * dot_h * dot_w;
    const int y = hh - dot_h + r / dot_w, xx = 2 * ww + j * (dot_w + pad) + r % dot_w;
    if (xx >= 3 * ww) continue;
    float m = mask[j];
    if (round_up) m = m > 0.5f ? 1.f : 0.f;
    const int ch = (m == 0.f) ? 1 : 2;
    unsigned char* px = o + ((size_t)y * (3 * ww) + xx) * 3;
    px[0] = px[1] = px[2] = 0;
    px[ch] = (j == f) ? 255 : 150;
  }
}

}  // namespace

extern "C" int ivf_viz_triptych(ivf_handle* h, int clip_dtype, const void* clip, const float* cam, const float* pert,
                                const float* mask, int t, int hh, int ww, int draw_dots, uint8_t* out, void* stream) {
  IVF_ON_DEVICE(h);
  IVF_REQUIRE(h && clip && cam && pert && out, "ivf_viz_triptych: null argument");
  IVF_REQUIRE(t > 0 && hh > 0 && ww > 0, "ivf_viz_triptych: bad extent");
  IVF_REQUIRE(clip_dtype == IVF_F32 || clip_dtype == IVF_U8, "ivf_viz_triptych: clip must be fp32 or uint8");
  if (draw_dots) IVF_REQUIRE(mask != nullptr, "ivf_viz_triptych: dots need the mask");
  cudaStream_t st = (cudaStream_t)stream;
  if (clip_dtype == IVF_F32)
    viz_triptych_kernel<float><<<t, 256, 0, st>>>((const float*)clip, cam, pert, t, hh, ww, out);
  else
    viz_triptych_kernel<unsigned char><<<t, 256, 0, st>>>((const unsigned char*)clip, cam, pert, t, hh, ww, out);
  IVF_LAUNCHED(h);
  if (draw_dots) {
    viz_dots_kernel<<<t, 256, 0, st>>>(out, mask, t, hh, ww, 1);
    IVF_LAUNCHED(h);
  }
  return IVF_OK;
}
