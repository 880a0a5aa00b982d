// Synthetic-burst preprocessing arithmetic of /root/reference/data_utils.py:198-265 (with helpers
// :432-466) as one bandwidth kernel: uint8 -> (v/255)^degamma -> per-frame crop -> up x up AREA
// box mean -> channel mean -> white level -> read/shot noise -> noise-level channel -> NHWC.
// The reference draws crops / levels / normals from TF's RNG; here every draw is an input.
#include "ie_common.cuh"

namespace ie {

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"): counter-based, so the normals of
// element i depend only on (seed, i) - reproducible across grids, ranks and the numpy oracle (oracle/preprocess.py).
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f + 2.98023223876953125e-08f; }
// two standard normals for element `idx`: Box-Muller on (u0,u1) and (u2,u3)
__device__ __forceinline__ void normal_pair(unsigned long long seed, unsigned long long idx, float& z0, float& z1) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float two_pi = 6.283185307179586f;
  z0 = sqrtf(-2.f * logf(u01(c[0]))) * cosf(two_pi * u01(c[1]));
  z1 = sqrtf(-2.f * logf(u01(c[2]))) * cosf(two_pi * u01(c[3]));
}

__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ src, int hs, int ws, int c, const int32_t* __restrict__ org, int up,
                     float degamma, const float* __restrict__ wl, const float* __restrict__ sig_read,
                     const float* __restrict__ sig_shot, const float* __restrict__ n_read,
                     const float* __restrict__ n_shot, unsigned long long seed, int use_rng, int layer_type, int h,
                     int w, int T, float* __restrict__ x, float* __restrict__ truth, long long total) {
  __shared__ float lut[256];   // (v/255)^degamma has 256 possible values (data_utils.py:213)
  lut[threadIdx.x] = powf((float)threadIdx.x / 255.f, degamma);
  __syncthreads();
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y, n = blockIdx.z;
  if (px >= w) return;
  const long long t = ((long long)n * h + py) * w + px;
  (void)total;
  const int add = layer_type == 1 ? 1 : (layer_type == 2 ? 2 : 0);
  const float wln = wl[n], sr = sig_read[n], ss = sig_shot[n];
  const uint8_t* img = src + (long long)n * hs * ws * c;
  const bool img_aligned = (reinterpret_cast<uintptr_t>(img) & 3) == 0;
  const float inv_area = 1.f / (float)(up * up);
  float* xo = x + t * (T + add);
  float noisy0 = 0.f;
  for (int f = 0; f < T; ++f) {
    const int oy = org[((long long)n * T + f) * 2] + py * up;
    const int ox = org[((long long)n * T + f) * 2 + 1] + px * up;
    float csum = 0.f;
    // fast path (the reference defaults: upscale 4, grey source): the 4x4 window is inside the image -> per row one
    // unaligned 4-byte fetch assembled from two aligned words (funnel shift), no per-sample bounds checks
    const long long first = (long long)oy * ws + ox;
    if (up == 4 && c == 1 && oy >= 0 && oy + 4 <= hs && ox >= 0 && ox + 4 <= ws &&
        (((long long)(oy + 3) * ws + ox) & ~3ll) + 8 <= (long long)hs * ws && img_aligned) {
      float s = 0.f;
#pragma unroll
      for (int dy = 0; dy < 4; ++dy) {
        const long long off = first + (long long)dy * ws;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (off & ~3ll));
        const uint32_t lo = __ldg(wp), hi = __ldg(wp + 1);
        const uint32_t v = __funnelshift_r(lo, hi, 8 * (int)(off & 3));
        s += (lut[v & 255u] + lut[(v >> 8) & 255u]) + (lut[(v >> 16) & 255u] + lut[v >> 24]);
      }
      csum = s * inv_area;
    } else
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
    } else if (use_rng) {
      float zs, zr;                                                 // tf.random.normal x2 of add_read_shot_tf, on device
      normal_pair(seed, (unsigned long long)(t * T + f), zs, zr);
      v = tr + sqrtf(tr) * ss * zs + sr * zr;
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
  IE_REQUIRE(n <= 65535 && h <= 65535, "preprocess_u8: grid too large");
  preprocess_u8_kernel<<<dim3(ie_ceil_div(w, 256), h, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, hs, ws, c, org, up, degamma, wl, sig_read, sig_shot, n_read, n_shot, 0ull, 0, layer_type, h, w, T, x, truth, total);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

extern "C" int ie_preprocess_u8_rng(const uint8_t* src, int n, int hs, int ws, int c, const int32_t* org, int up,
                                    float degamma, const float* wl, const float* sig_read, const float* sig_shot,
                                    unsigned long long seed, int layer_type, int h, int w, int T, float* x, float* truth,
                                    void* stream) {
  using namespace ie;
  IE_REQUIRE(src && org && wl && sig_read && sig_shot && x && truth, "preprocess_u8_rng: null pointer");
  IE_REQUIRE(n > 0 && hs > 0 && ws > 0 && c > 0 && up >= 1 && h > 0 && w > 0 && T >= 1, "preprocess_u8_rng: bad sizes");
  IE_REQUIRE(layer_type >= 0 && layer_type <= 2, "preprocess_u8_rng: layer_type must be 0, 1 or 2");
  const long long total = (long long)n * h * w;
  IE_REQUIRE(n <= 65535 && h <= 65535, "preprocess_u8_rng: grid too large");
  preprocess_u8_kernel<<<dim3(ie_ceil_div(w, 256), h, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, hs, ws, c, org, up, degamma, wl, sig_read, sig_shot, nullptr, nullptr, seed, 1, layer_type, h, w, T, x, truth,
      total);
  IE_LAUNCH_CHECK();
  return IE_OK;
}
