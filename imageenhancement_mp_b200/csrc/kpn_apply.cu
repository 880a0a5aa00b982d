// Fused per-pixel kernel-prediction filter.  Replaces the filter synthesis + Convolve +
// Convolve_perlayer of /root/reference/model_library.py:439-451 (and :114-168).
//
// The reference materialises filts[n,y,x,i,j,t] = sum_b Bas[n,i,j,t,b] * Coef[n,y,x,b] (3.6 KB per
// pixel, after two 36 KB-per-pixel tiles) and then multiplies it with 225 shifted copies of the
// burst.  Re-associated, basis first:
//     G[b]        = sum_{i,j} Bas[n,i,j,t,b] * pad0(burst)[n, y+i-K/2, x+j-K/2, t]   (a per-image correlation)
//     out[..,1+t] = T * sum_b Coef[n,y,x,b] * G[b]                                   (model_library.py:164, 444)
//     out[..,0]   = sum_t out[..,1+t] / T                                            (model_library.py:134, 447)
// so nothing per-pixel is ever stored.  fp32 on the CUDA cores: 2*K*K*T*B FLOP per pixel.
//
// Block = 16 x 16 threads, each thread owns 4 consecutive x -> 64 x 16 output pixels of one image.
// Per frame: the (16+K-1) x (64+K-1) burst tile (zero padded) and the K*K*B basis slice sit in
// shared memory; per filter row a thread pulls its 4+K-1 burst values with 128-bit loads and
// every basis value is a warp-wide broadcast, giving 4*BC FFMA per 16-byte basis load.
#include "ie_common.cuh"

namespace ie {

constexpr int kTileW = 64, kTileH = 16, kPxPerThread = 4;
constexpr int kMaxK = 15;
constexpr int kRowRegs = ((kPxPerThread + kMaxK - 1 + 3) / 4) * 4;   // 20

template <int BC>
__global__ void __launch_bounds__(256, 2)
kpn_apply_kernel(const float* __restrict__ burst, int burst_pitch, const float* __restrict__ coef, int Hc, int Wc,
                 const float* __restrict__ bas, float* __restrict__ out, int H, int W, int T, int K, int B,
                 int tiles_x, int tiles_y) {
  extern __shared__ float smem[];
  const int halo = K - 1;
  const int sw = ((kTileW + halo + 3) / 4) * 4 + 4;   // smem row pitch (floats), 16-byte multiple
  const int sh = kTileH + halo;
  const int nchunk = (B + BC - 1) / BC;
  const int bpad = nchunk * BC;        // BC is a multiple of 4 or B%BC==0 with BC even -> see host
  float* s_burst = smem;               // [sh][sw]
  float* s_bas = smem + sh * sw;       // [K*K][bpad]

  int bid = blockIdx.x;
  const int tx_tile = bid % tiles_x; bid /= tiles_x;
  const int ty_tile = bid % tiles_y;
  const int img = bid / tiles_y;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int x0 = tx_tile * kTileW, y0 = ty_tile * kTileH;
  const int px = x0 + tx * kPxPerThread, py = y0 + ty;
  const int kpad = K / 2;

  float dsum[kPxPerThread];
#pragma unroll
  for (int p = 0; p < kPxPerThread; ++p) dsum[p] = 0.f;

  const float* burst_img = burst + (long long)img * H * W * burst_pitch;
  const float* bas_img = bas + (long long)img * K * K * T * B;
  const long long pix_base = ((long long)img * H + py) * W + px;
  const long long coef_base = ((long long)img * Hc + py) * Wc + px;

  for (int t = 0; t < T; ++t) {
    __syncthreads();
    // burst tile of frame t (zero outside the image = tf.pad at model_library.py:126)
    for (int i = threadIdx.x; i < sh * sw; i += 256) {
      const int ly = i / sw, lx = i - ly * sw;
      const int gy = y0 + ly - kpad, gx = x0 + lx - kpad;
      float v = 0.f;
      if (lx < kTileW + halo && gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = burst_img[((long long)gy * W + gx) * burst_pitch + t];
      s_burst[i] = v;
    }
    // basis slice of frame t: [tap][b] zero padded to bpad
    for (int i = threadIdx.x; i < K * K * bpad; i += 256) {
      const int tap = i / bpad, b = i - tap * bpad;
      s_bas[i] = (b < B) ? bas_img[((long long)tap * T + t) * B + b] : 0.f;
    }
    __syncthreads();

    float res[kPxPerThread];
#pragma unroll
    for (int p = 0; p < kPxPerThread; ++p) res[p] = 0.f;

    for (int ch = 0; ch < nchunk; ++ch) {
      float g[kPxPerThread][BC];
#pragma unroll
      for (int p = 0; p < kPxPerThread; ++p)
#pragma unroll
        for (int b = 0; b < BC; ++b) g[p][b] = 0.f;

      for (int i = 0; i < K; ++i) {
        float row[kRowRegs];
        const float4* rp = reinterpret_cast<const float4*>(s_burst + (ty + i) * sw + tx * kPxPerThread);
#pragma unroll
        for (int v = 0; v < kRowRegs / 4; ++v) {
          const float4 q = rp[v];
          row[4 * v] = q.x; row[4 * v + 1] = q.y; row[4 * v + 2] = q.z; row[4 * v + 3] = q.w;
        }
        const float* bp = s_bas + (i * K) * bpad + ch * BC;
#pragma unroll
        for (int j = 0; j < kMaxK; ++j) {
          if (j < K) {
            float bv[BC];
            if constexpr (BC % 4 == 0) {
#pragma unroll
              for (int v = 0; v < BC / 4; ++v) {
                const float4 q = *reinterpret_cast<const float4*>(bp + j * bpad + 4 * v);
                bv[4 * v] = q.x; bv[4 * v + 1] = q.y; bv[4 * v + 2] = q.z; bv[4 * v + 3] = q.w;
              }
            } else {
#pragma unroll
              for (int v = 0; v < BC / 2; ++v) {
                const float2 q = *reinterpret_cast<const float2*>(bp + j * bpad + 2 * v);
                bv[2 * v] = q.x; bv[2 * v + 1] = q.y;
              }
            }
#pragma unroll
            for (int p = 0; p < kPxPerThread; ++p)
#pragma unroll
              for (int b = 0; b < BC; ++b) g[p][b] = fmaf(row[j + p], bv[b], g[p][b]);
          }
        }
      }
      // mix with the per-pixel coefficients
      if (py < H) {
#pragma unroll
        for (int p = 0; p < kPxPerThread; ++p) {
          if (px + p < W) {
            const float* cp = coef + (coef_base + p) * B + ch * BC;
#pragma unroll
            for (int b = 0; b < BC; ++b)
              if (ch * BC + b < B) res[p] = fmaf(__ldg(cp + b), g[p][b], res[p]);
          }
        }
      }
    }
    if (py < H) {
#pragma unroll
      for (int p = 0; p < kPxPerThread; ++p) {
        if (px + p < W) {
          out[(pix_base + p) * (T + 1) + 1 + t] = res[p] * (float)T;
          dsum[p] += res[p];
        }
      }
    }
  }
  if (py < H) {
#pragma unroll
    for (int p = 0; p < kPxPerThread; ++p)
      if (px + p < W) out[(pix_base + p) * (T + 1)] = dsum[p];
  }
}

}  // namespace ie

extern "C" int ie_kpn_apply_f32(const float* burst, int burst_pitch, const float* coef, int hc, int wc, const float* bas,
                                float* out, int n, int h, int w, int T, int K, int B, void* stream) {
  using namespace ie;
  IE_REQUIRE(burst && coef && bas && out, "kpn_apply: null pointer");
  IE_REQUIRE(n > 0 && h > 0 && w > 0 && T > 0 && B > 0, "kpn_apply: bad sizes");
  IE_REQUIRE(K >= 1 && K <= kMaxK && (K & 1), "kpn_apply: K must be odd and <= %d (got %d)", kMaxK, K);
  IE_REQUIRE(burst_pitch >= T, "kpn_apply: burst_pitch < T");
  IE_REQUIRE(hc >= h && wc >= w, "kpn_apply: coef extent %dx%d smaller than the image %dx%d", hc, wc, h, w);
  const int tiles_x = (w + kTileW - 1) / kTileW, tiles_y = (h + kTileH - 1) / kTileH;
  const long long blocks = (long long)n * tiles_x * tiles_y;
  IE_REQUIRE(blocks < (1ll << 31), "kpn_apply: too many tiles");
  const int halo = K - 1;
  const int sw = ((kTileW + halo + 3) / 4) * 4 + 4, sh = kTileH + halo;
  const bool ten = (B % 10 == 0);
  const int BC = ten ? 10 : 8;
  const int bpad = ((B + BC - 1) / BC) * BC;
  const size_t smem = sizeof(float) * ((size_t)sh * sw + (size_t)K * K * bpad);
  IE_REQUIRE(smem <= 200 * 1024, "kpn_apply: B=%d needs %zu bytes of shared memory", B, smem);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ten) {
    IE_CUDA(cudaFuncSetAttribute(kpn_apply_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kpn_apply_kernel<10><<<(unsigned)blocks, 256, smem, st>>>(burst, burst_pitch, coef, hc, wc, bas, out, h, w, T, K, B,
                                                             tiles_x, tiles_y);
  } else {
    IE_CUDA(cudaFuncSetAttribute(kpn_apply_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kpn_apply_kernel<8><<<(unsigned)blocks, 256, smem, st>>>(burst, burst_pitch, coef, hc, wc, bas, out, h, w, T, K, B,
                                                            tiles_x, tiles_y);
  }
  IE_LAUNCH_CHECK();
  return IE_OK;
}
