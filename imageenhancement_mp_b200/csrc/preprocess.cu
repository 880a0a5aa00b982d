// Synthetic-burst preprocessing arithmetic of /root/reference/data_utils.py:198-265 (with helpers
// :432-466) as one bandwidth kernel: uint8 -> (v/255)^degamma -> per-frame crop -> up x up AREA
// box mean -> channel mean -> white level -> read/shot noise -> noise-level channel -> NHWC.
// The reference draws crops / levels / normals from TF's RNG; here every draw is an input.
#include "ie_common.cuh"

namespace ie {

__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ src, int hs, int ws, int c, const int32_t* __restrict__ org, int up,
                     float degamma, const float* __restrict__ wl, const float* __restrict__ sig_read,
                     const float* __restrict__ sig_shot, const float* __restrict__ n_read,
                     const float* __restrict__ n_shot, int layer_type, int h, int w, int T, float* __restrict__ x,
                     float* __restrict__ truth, long long total) {
  __shared__ float lut[256];   // (v/255)^degamma has 256 possible values (data_utils.py:213)
  lut[threadIdx.x] = powf((float)threadIdx.x / 255.f, degamma);
  __syncthreads();
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int px = (int)(t % w);
  const int py = (int)((t / w) % h);
  const int n = (int)(t / ((long long)w * h));
  const int add = layer_type == 1 ? 1 : (layer_type == 2 ? 2 : 0);
  const float wln = wl[n], sr = sig_read[n], ss = sig_shot[n];
  const uint8_t* img = src + (long long)n * hs * ws * c;
  const float inv_area = 1.f / (float)(up * up);
  float* xo = x + t * (T + add);
  float noisy0 = 0.f;
  for (int f = 0; f < T; ++f) {
    const int oy = org[((long long)n * T + f) * 2] + py * up;
    const int ox = org[((long long)n * T + f) * 2 + 1] + px * up;
    float csum = 0.f;
    for (int ch = 0; ch < c; ++ch) {          // AREA mean per channel, then channel mean (:459, :220)
      float s = 0.f;
      for (int dy = 0; dy < up; ++dy) {
        const int sy = oy + dy;
        if (sy < 0 || sy >= hs) continue;      // zero padding of make_first_truth (:436-438)
        for (int dx = 0; dx < up; ++dx) {
          const int sx = ox + dx;
          if (sx < 0 || sx >= ws) continue;
          s += lut[img[((long long)sy * ws + sx) * c + ch]];
        }
      }
      csum += s * inv_area;
    }
    const float tr = wln * (csum / (float)c);                       // :230
    float v = tr;
    if (n_read != nullptr && n_shot != nullptr) {
      const long long ni = t * T + f;
      v = tr + sqrtf(tr) * ss * n_shot[ni] + sr * n_read[ni];       // :463-465
    }
    xo[f] = v;
    if (f == 0) {
      noisy0 = v;
      truth[t * 2] = tr;                                            // :248-250
      truth[t * 2 + 1] = wln;                                       // :252
    }
  }
  if (layer_type == 1) {
    xo[T] = sqrtf(sr * sr + fmaxf(0.f, noisy0) * ss * ss);          // :256
  } else if (layer_type == 2) {
    xo[T] = sr;                                                     // :257
    xo[T + 1] = ss;
  }
}

}  // namespace ie

extern "C" int ie_preprocess_u8(const uint8_t* src, int n, int hs, int ws, int c, const int32_t* org, int up,
                                float degamma, const float* wl, const float* sig_read, const float* sig_shot,
                                const float* n_read, const float* n_shot, int layer_type, int h, int w, int T,
                                float* x, float* truth, void* stream) {
  using namespace ie;
  IE_REQUIRE(src && org && wl && sig_read && sig_shot && x && truth, "preprocess_u8: null pointer");
  IE_REQUIRE(n > 0 && hs > 0 && ws > 0 && c > 0 && up >= 1 && h > 0 && w > 0 && T >= 1, "preprocess_u8: bad sizes");
  IE_REQUIRE(layer_type >= 0 && layer_type <= 2, "preprocess_u8: layer_type must be 0 (empty), 1 (singlestd), 2 (dualparams)");
  IE_REQUIRE((n_read == nullptr) == (n_shot == nullptr), "preprocess_u8: give both noise tensors or neither");
  const long long total = (long long)n * h * w;
  preprocess_u8_kernel<<<ie_ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, hs, ws, c, org, up, degamma, wl, sig_read, sig_shot, n_read, n_shot, layer_type, h, w, T, x, truth, total);
  IE_LAUNCH_CHECK();
  return IE_OK;
}
