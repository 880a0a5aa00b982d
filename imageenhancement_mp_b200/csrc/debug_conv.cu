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
  int h, w, b;          // image size and border (1 = shared-border raster, 0 = dense NHWC)
  int hv, wv, kh, kw, pad, cin, x_pitch, x_coff, cout, y_pitch, y_coff, relu, epilogue;
};

// row of pixel (img, y, x) in either layout
__device__ __forceinline__ long long naive_row(const NaiveParams& p, int img, int y, int x) {
  return ((long long)img * (p.h + p.b) + y + p.b) * (p.w + p.b) + x;
}

// plain definition of the convolution: taps outside the image contribute nothing (zero padding)
__device__ float naive_dot(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                           const NaiveParams& p, int img, int y, int xx, int co) {
  float acc = 0.f;
  const int ktot = p.kh * p.kw * p.cin;
  for (int i = 0; i < p.kh; ++i) {
    for (int j = 0; j < p.kw; ++j) {
      const int sy = y + i - p.pad, sx = xx + j - p.pad;
      if (sy < 0 || sy >= p.h || sx < 0 || sx >= p.w) continue;
      const __nv_bfloat16* xp = x + naive_row(p, img, sy, sx) * p.x_pitch + p.x_coff;
      const __nv_bfloat16* wp = w + (long long)co * ktot + (i * p.kw + j) * p.cin;
      for (int c = 0; c < p.cin; ++c) acc = fmaf(__bfloat162float(xp[c]), __bfloat162float(wp[c]), acc);
    }
  }
  return acc;
}

// row r -> pixel; false for border rows (raster) and for pixels outside the valid output extent
__device__ bool naive_valid(const NaiveParams& p, long long r, int& img, int& y, int& x) {
  const int wp = p.w + p.b, plane = (p.h + p.b) * wp;
  img = (int)(r / plane);
  const int pr = (int)(r - (long long)img * plane);
  y = pr / wp - p.b;
  x = pr - (y + p.b) * wp;
  return y >= 0 && y < p.hv && x < p.wv;
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
    v = naive_dot(x, w, p, img, yy, xx, co) + (bias ? bias[co] : 0.f);
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
  const long long pix = ((long long)img * p.hv + yy) * p.wv + xx;
  float v[256];
  float mx = -INFINITY;
  for (int co = 0; co < p.cout; ++co) {
    float a = naive_dot(x, w, p, img, yy, xx, co) + (bias ? bias[co] : 0.f);
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
                                     void* y_bf16, float* y_f32, float* y_aux, void* /*workspace*/,
                                     long long /*workspace_bytes*/, void* stream) {
  using namespace ie;
  if (int rc = validate_conv_desc(d, x, w_packed, y_bf16, y_f32)) return rc;
  NaiveParams p{};
  IE_REQUIRE((d->kh == 3 && d->kw == 3) || (d->kh == 2 && d->kw == 2) || (d->kh == 1 && d->kw == 1),
             "debug conv: unsupported kernel size");
  p.kh = d->kh; p.kw = d->kw;
  p.pad = d->kh == 3 ? 1 : 0;                     // 3x3 'same' centred; 2x2 'valid' taps at +0/+1; 1x1
  p.h = d->h; p.w = d->w; p.b = d->dense ? 0 : 1;
  p.R = (long long)d->n_img * (d->h + p.b) * (d->w + p.b);
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
