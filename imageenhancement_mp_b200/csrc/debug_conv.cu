// TESTS ONLY: a slow CUDA-core convolution with exactly the contract of ie_conv2d_nhwc_bf16
// (same rasters, packed weights, epilogues).  It exists so the tcgen05 kernel can be cross-checked
// on the GPU at sizes the CPU oracle cannot reach and so a descriptor bug in the tensor-core path
// can be told apart from a bug anywhere else.  The product path (model_library / eval) never
// calls it.
#include <cuda_bf16.h>

#include "ie_common.cuh"
#include "ie_ptx.cuh"

namespace ie {

int validate_conv_desc(const ie_conv_desc* d, const void* x, const void* w, void* y_bf16, float* y_f32);
int choose_n_tile(int cout, int epilogue);

struct NaiveParams {
  long long R;
  int plane, wp, hv, wv, ntaps, cin, x_pitch, x_coff, cout, y_pitch, y_coff, relu, epilogue;
  int tap_shift[9];
};

__device__ float naive_dot(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                           const NaiveParams& p, long long r, int co) {
  float acc = 0.f;
  const int ktot = p.ntaps * p.cin;
  for (int tap = 0; tap < p.ntaps; ++tap) {
    const long long rr = r + p.tap_shift[tap];
    if (rr < 0 || rr >= p.R) continue;
    const __nv_bfloat16* xp = x + rr * p.x_pitch + p.x_coff;
    const __nv_bfloat16* wp = w + (long long)co * ktot + tap * p.cin;
    for (int c = 0; c < p.cin; ++c) acc = fmaf(__bfloat162float(xp[c]), __bfloat162float(wp[c]), acc);
  }
  return acc;
}

__device__ bool naive_valid(const NaiveParams& p, long long r, int& img, int& y, int& x) {
  img = (int)(r / p.plane);
  const int pr = (int)(r - (long long)img * p.plane);
  y = pr / p.wp;
  x = pr - y * p.wp;
  return y >= 1 && y <= p.hv && x < p.wv;
}

__global__ void naive_conv_bf16_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                       const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, NaiveParams p) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= p.R * p.cout) return;
  const long long r = t / p.cout;
  const int co = (int)(t - r * p.cout);
  int img, yy, xx;
  float v = 0.f;
  if (naive_valid(p, r, img, yy, xx)) {
    v = naive_dot(x, w, p, r, co) + (bias ? bias[co] : 0.f);
    if (p.relu) v = fmaxf(v, 0.f);
  }
  y[r * p.y_pitch + p.y_coff + co] = __float2bfloat16_rn(v);
}

__global__ void naive_conv_f32_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                      const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ aux,
                                      NaiveParams p) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= p.R) return;
  int img, yy, xx;
  if (!naive_valid(p, r, img, yy, xx)) return;
  const long long pix = ((long long)img * p.hv + (yy - 1)) * p.wv + xx;
  float v[256];
  float mx = -INFINITY;
  for (int co = 0; co < p.cout; ++co) {
    float a = naive_dot(x, w, p, r, co) + (bias ? bias[co] : 0.f);
    if (p.relu) a = fmaxf(a, 0.f);
    v[co] = a;
    mx = fmaxf(mx, a);
  }
  if (p.epilogue == IE_EPI_F32_SOFTMAX) {
    float s = 0.f;
    for (int co = 0; co < p.cout; ++co) {
      if (aux) aux[pix * p.cout + co] = v[co];
      v[co] = expf(v[co] - mx);
      s += v[co];
    }
    for (int co = 0; co < p.cout; ++co) y[pix * p.cout + co] = v[co] / s;
  } else {
    const int pitch = p.y_pitch > 0 ? p.y_pitch : p.cout;          // fp32 channel slice (chunked wide layers)
    const int coff = p.y_pitch > 0 ? p.y_coff : 0;
    for (int co = 0; co < p.cout; ++co) y[pix * pitch + coff + co] = v[co];
  }
}

}  // namespace ie

extern "C" int ie_debug_conv2d_naive(const ie_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                                     void* y_bf16, float* y_f32, float* y_aux, void* stream) {
  using namespace ie;
  if (int rc = validate_conv_desc(d, x, w_packed, y_bf16, y_f32)) return rc;
  NaiveParams p{};
  const int wp = d->w + 1;
  if (d->kh == 3 && d->kw == 3) {
    p.ntaps = 9;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) p.tap_shift[i * 3 + j] = (i - 1) * wp + (j - 1);
  } else if (d->kh == 2 && d->kw == 2) {
    p.ntaps = 4;
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) p.tap_shift[i * 2 + j] = i * wp + j;
  } else if (d->kh == 1 && d->kw == 1) {
    p.ntaps = 1;
  } else {
    IE_REQUIRE(false, "debug conv: unsupported kernel size");
  }
  p.R = (long long)d->n_img * (d->h + 1) * wp;
  p.plane = (d->h + 1) * wp;
  p.wp = wp;
  p.hv = d->hv; p.wv = d->wv;
  p.cin = d->cin; p.x_pitch = d->x_pitch; p.x_coff = d->x_coff;
  p.cout = d->cout; p.y_pitch = d->y_pitch; p.y_coff = d->y_coff;
  p.relu = d->relu; p.epilogue = d->epilogue;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->epilogue == IE_EPI_BF16_RASTER) {
    const long long total = p.R * p.cout;
    naive_conv_bf16_kernel<<<ie_ceil_div(total, 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w_packed), bias,
        static_cast<__nv_bfloat16*>(y_bf16), p);
  } else {
    naive_conv_f32_kernel<<<ie_ceil_div(p.R, 128), 128, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                              static_cast<const __nv_bfloat16*>(w_packed), bias, y_f32,
                                                              y_aux, p);
  }
  IE_LAUNCH_CHECK();
  return IE_OK;
}
