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
// Block = TX x TY threads (TX*TY <= 256, chosen on the host so that tiles cover the image with little waste:
// 104-wide patches get TX = 26, one tile per row of tiles), each thread owns 4 consecutive x.  The burst tile
// of ALL T frames (zero padded halo of K-1) and the whole K*K*T*B basis of the image are staged in shared
// memory once (one barrier per block); per filter row a thread pulls its 4+K-1 burst values with 128-bit
// loads and every basis value is a warp-wide broadcast, giving 4*BC FFMA per 16-byte basis load.
#include "ie_common.cuh"

namespace ie {

constexpr int kMaxK = 15;

struct KpnTiling {
  int px;                // consecutive x per thread (4 or 8)
  int tx, ty;            // threads per tile row / rows per tile
  int tiles_x, tiles_y;
  int sw, sh;            // smem tile pitch (floats, multiple of 4) and rows
  int bpad;
  int tp;                // frames staged per pass (all T when they fit)
  size_t smem_bytes;
};

// KK / BPAD: compile-time kernel size and basis pitch of the hot configuration (15, 10) - every basis load then
// has an immediate offset and the tap loops carry no predicates; 0 = run-time values (any odd K <= 15, any B).
template <int BC, int KK, int BPAD, int kPxPerThread>
__global__ void __launch_bounds__(256, kPxPerThread == 4 ? 2 : 1)
kpn_apply_kernel(const float* __restrict__ burst, int burst_pitch, const float* __restrict__ coef, int Hc, int Wc,
                 const float* __restrict__ bas, float* __restrict__ out, int H, int W, int T, int K_rt, int B,
                 const KpnTiling tl) {
  extern __shared__ float smem[];
  constexpr int kRowRegs = ((kPxPerThread + kMaxK - 1 + 3) / 4) * 4;   // 20 / 24
  const int K = KK ? KK : K_rt;
  const int halo = K - 1;
  const int sw = tl.sw, sh = tl.sh, bpad = BPAD ? BPAD : tl.bpad;
  const int nchunk = bpad / BC;
  const int tile_w = tl.tx * kPxPerThread;
  const int TP = tl.tp;
  float* s_burst = smem;                    // [TP][sh][sw]
  float* s_bas = smem + TP * sh * sw;       // [TP][K*K][bpad]

  int bid = blockIdx.x;
  const int tx_tile = bid % tl.tiles_x; bid /= tl.tiles_x;
  const int ty_tile = bid % tl.tiles_y;
  const int img = bid / tl.tiles_y;
  const int ty = threadIdx.x / tl.tx, tx = threadIdx.x - ty * tl.tx;
  const int x0 = tx_tile * tile_w, y0 = ty_tile * tl.ty;
  const int px = x0 + tx * kPxPerThread, py = y0 + ty;
  const int kpad = K / 2;
  const bool active = ty < tl.ty;

  const float* burst_img = burst + (long long)img * H * W * burst_pitch;
  const float* bas_img = bas + (long long)img * K * K * T * B;

  const long long pix_base = ((long long)img * H + py) * W + px;
  const long long coef_base = ((long long)img * Hc + py) * Wc + px;
  float dsum[kPxPerThread];
#pragma unroll
  for (int p = 0; p < kPxPerThread; ++p) dsum[p] = 0.f;

  for (int t0 = 0; t0 < T; t0 += TP) {
  const int nt = min(TP, T - t0);
  if (t0) __syncthreads();
  // ---- stage the burst tile of frames [t0, t0+nt): a tile row is one contiguous run of (tile_w+halo)*pitch floats
  {
    // index splits use a float reciprocal (exact for these small non-negative ints) instead of integer division
    const int run = (tile_w + halo) * burst_pitch;
    const int gx0 = x0 - kpad;
    const float inv_run = 1.f / (float)run, inv_pitch = 1.f / (float)burst_pitch;
    for (int idx = threadIdx.x; idx < sh * run; idx += blockDim.x) {
      const int ly = (int)(((float)idx + 0.5f) * inv_run), e = idx - ly * run;
      const int lx = (int)(((float)e + 0.5f) * inv_pitch), ch = e - lx * burst_pitch - t0;
      if (ch >= 0 && ch < nt) {
        const int gy = y0 + ly - kpad, gx = gx0 + lx;
        float v = 0.f;                      // zero outside the image = tf.pad at model_library.py:126
        if (gy >= 0 && gy < H && gx >= 0 && gx < W)
          v = __ldg(burst_img + ((long long)gy * W + gx) * burst_pitch + t0 + ch);
        s_burst[(ch * sh + ly) * sw + lx] = v;
      }
    }
    // columns [tile_w+halo, sw) are read into registers (never used) by the 128-bit row loads: keep them finite
    const int extra = sw - (tile_w + halo);
    const float inv_extra = 1.f / (float)extra;
    for (int idx = threadIdx.x; idx < nt * sh * extra; idx += blockDim.x) {
      const int r = (int)(((float)idx + 0.5f) * inv_extra), c = idx - r * extra;
      s_burst[r * sw + tile_w + halo + c] = 0.f;
    }
    // basis of the image: global [tap][T][B] (read in memory order) -> smem [t][tap][bpad]
    const int tb = T * B;
    const float inv_tb = 1.f / (float)tb, inv_b = 1.f / (float)B;
    for (int idx = threadIdx.x; idx < K * K * tb; idx += blockDim.x) {
      const int tap = (int)(((float)idx + 0.5f) * inv_tb), r = idx - tap * tb;
      const int t = (int)(((float)r + 0.5f) * inv_b), b = r - t * B;
      const int tl_ = t - t0;
      if (tl_ >= 0 && tl_ < nt) s_bas[(tl_ * K * K + tap) * bpad + b] = __ldg(bas_img + idx);
    }
    if (bpad > B) {                          // zero the padding of the basis pitch
      const int padw = bpad - B;
      for (int idx = threadIdx.x; idx < nt * K * K * padw; idx += blockDim.x) {
        const int r = idx / padw, c = idx - r * padw;
        s_bas[r * bpad + B + c] = 0.f;
      }
    }
  }
  __syncthreads();
  if (!active) continue;

  for (int tl_ = 0; tl_ < nt; ++tl_) {
    const int t = t0 + tl_;
    float res[kPxPerThread];
#pragma unroll
    for (int p = 0; p < kPxPerThread; ++p) res[p] = 0.f;
    const float* sb_t = s_burst + (tl_ * sh + ty) * sw + tx * kPxPerThread;
    const float* bas_t = s_bas + tl_ * K * K * bpad;

    for (int ch = 0; ch < nchunk; ++ch) {
      // accumulators as fp32 PAIRS over the basis index: Blackwell's packed FFMA2 (fma.rn.f32x2) does two
      // IEEE fp32 FMAs per issue slot - a scalar FFMA stream tops out at half the FP32 lanes on sm_100.
      float2 g[kPxPerThread][BC / 2];
#pragma unroll
      for (int p = 0; p < kPxPerThread; ++p)
#pragma unroll
        for (int b = 0; b < BC / 2; ++b) g[p][b] = make_float2(0.f, 0.f);

      for (int i = 0; i < K; ++i) {
        float2 row[kRowRegs];               // every burst value duplicated into both halves of a pair
        const float4* rp = reinterpret_cast<const float4*>(sb_t + i * sw);
#pragma unroll
        for (int v = 0; v < kRowRegs / 4; ++v) {
          const float4 q = rp[v];
          row[4 * v] = make_float2(q.x, q.x); row[4 * v + 1] = make_float2(q.y, q.y);
          row[4 * v + 2] = make_float2(q.z, q.z); row[4 * v + 3] = make_float2(q.w, q.w);
        }
        const float* bp = bas_t + (i * K) * bpad + ch * BC;
#pragma unroll
        for (int j = 0; j < (KK ? KK : kMaxK); ++j) {
          if (KK || j < K) {
            float2 bv[BC / 2];
#pragma unroll
            for (int v = 0; v < BC / 2; ++v) bv[v] = *reinterpret_cast<const float2*>(bp + j * bpad + 2 * v);
#pragma unroll
            for (int p = 0; p < kPxPerThread; ++p)
#pragma unroll
              for (int b = 0; b < BC / 2; ++b) g[p][b] = __ffma2_rn(row[j + p], bv[b], g[p][b]);
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
              if (ch * BC + b < B) res[p] = fmaf(__ldg(cp + b), (b & 1) ? g[p][b >> 1].y : g[p][b >> 1].x, res[p]);
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
  }  // frame groups
  if (active && py < H) {
#pragma unroll
    for (int p = 0; p < kPxPerThread; ++p)
      if (px + p < W) out[(pix_base + p) * (T + 1)] = dsum[p];
  }
}

// Tiles: as few 128-px-wide columns of tiles as cover W, each an equal multiple of 4 px; rows balanced over H.
static KpnTiling kpn_tiling(int h, int w, int T, int K, int B, int BC, int px) {
  KpnTiling t{};
  t.px = px;
  const int max_tile_w = 32 * px;
  t.tiles_x = (w + max_tile_w - 1) / max_tile_w;
  const int tile_w = (((w + t.tiles_x - 1) / t.tiles_x) + px - 1) / px * px;
  t.tx = tile_w / px;
  int ty = 256 / t.tx;
  if (ty > 32) ty = 32;
  t.tiles_y = (h + ty - 1) / ty;
  t.ty = (h + t.tiles_y - 1) / t.tiles_y;
  const int halo = K - 1;
  t.sw = ((tile_w + halo + 3) / 4) * 4 + 4;
  t.sh = t.ty + halo;
  t.bpad = ((B + BC - 1) / BC) * BC;
  // frames per pass: all of them if the block then still fits twice per SM, else as many as ~100 KB holds
  const size_t per_frame = sizeof(float) * ((size_t)t.sh * t.sw + (size_t)K * K * t.bpad);
  int tp = (int)((100 * 1024) / per_frame);
  t.tp = tp < 1 ? 1 : (tp > T ? T : tp);
  t.smem_bytes = per_frame * t.tp;
  return t;
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
  const bool ten = (B % 10 == 0);
  const int BC = ten ? 10 : 8;
  const bool hot = ten && K == 15 && B == 10;                 // eval.py defaults (K=15, B=10)
  const int px = 4;      // 8 px per thread (one block of 8 warps per SM) measured slower on B200: 1.94 vs 1.38 ms at cfg2
  const KpnTiling tl = kpn_tiling(h, w, T, K, B, BC, px);
  const long long blocks = (long long)n * tl.tiles_x * tl.tiles_y;
  IE_REQUIRE(blocks < (1ll << 31), "kpn_apply: too many tiles");
  IE_REQUIRE(tl.smem_bytes <= 200 * 1024, "kpn_apply: T=%d B=%d needs %zu bytes of shared memory", T, B, tl.smem_bytes);
  const int threads = ((tl.tx * tl.ty + 31) / 32) * 32;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define IE_LAUNCH_KPN(BC_, KK_, BPAD_, PX_)                                                                         \
  do {                                                                                                                \
    IE_CUDA(cudaFuncSetAttribute(kpn_apply_kernel<BC_, KK_, BPAD_, PX_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)tl.smem_bytes));                                                                \
    kpn_apply_kernel<BC_, KK_, BPAD_, PX_><<<(unsigned)blocks, threads, tl.smem_bytes, st>>>(                         \
        burst, burst_pitch, coef, hc, wc, bas, out, h, w, T, K, B, tl);                                               \
  } while (0)
  if (hot && px == 8) IE_LAUNCH_KPN(10, 15, 10, 8);
  else if (hot) IE_LAUNCH_KPN(10, 15, 10, 4);
  else if (ten) IE_LAUNCH_KPN(10, 0, 0, 4);
  else IE_LAUNCH_KPN(8, 0, 0, 4);
#undef IE_LAUNCH_KPN
  IE_LAUNCH_CHECK();
  return IE_OK;
}

// -------------------------------------------------------------------------------------------------
// Convolve / cus_convolve / Convolve_perlayer with MATERIALISED per-pixel filters
// (model_library.py:114-168), for callers that build `filts` themselves.  One warp per pixel: the lanes
// stride over the pixel's K*K*T contiguous filter taps (coalesced 128-byte reads - the kernel is bound by the
// 4*K*K*T bytes of filter per pixel), the burst neighbourhood comes from L1/L2.
//   out[n,y,x,0]   = sum_{i,j,t} pad0(burst)[n,y+i-K/2,x+j-K/2,t] * filts[n,y,x,i,j,t]          (Convolve)
//   out[n,y,x,1+t] = T * sum_{i,j} pad0(burst)[n,y+i-K/2,x+j-K/2,t] * filts[n,y,x,i,j,t]        (Convolve_perlayer)
// -------------------------------------------------------------------------------------------------
namespace ie {
constexpr int kConvMaxT = 8;
__global__ void __launch_bounds__(256)
convolve_filts_kernel(const float* __restrict__ burst, int burst_pitch, const float* __restrict__ filts,
                      float* __restrict__ out, long long npix, int H, int W, int T, int K) {
  const long long pix = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= npix) return;
  const int lane = threadIdx.x & 31;
  const int x = (int)(pix % W);
  const int y = (int)((pix / W) % H);
  const long long img = pix / ((long long)W * H);
  const int taps = K * K * T, kpad = K / 2;
  const float* f = filts + pix * taps;
  const float* b = burst + img * H * W * burst_pitch;
  float acc[kConvMaxT];
#pragma unroll
  for (int t = 0; t < kConvMaxT; ++t) acc[t] = 0.f;
  for (int e = lane; e < taps; e += 32) {
    const int ij = e / T, t = e - ij * T;
    const int i = ij / K, j = ij - i * K;
    const int gy = y + i - kpad, gx = x + j - kpad;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const float v = __ldcs(f + e) * __ldg(b + ((long long)gy * W + gx) * burst_pitch + t);
#pragma unroll
      for (int k = 0; k < kConvMaxT; ++k)
        if (k == t) acc[k] += v;
    }
  }
  float total = 0.f;
#pragma unroll
  for (int t = 0; t < kConvMaxT; ++t) {
    if (t < T) {
      float s = acc[t];
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      total += s;
      if (lane == 0) out[pix * (T + 1) + 1 + t] = s * (float)T;
    }
  }
  if (lane == 0) out[pix * (T + 1)] = total;
}
}  // namespace ie

extern "C" int ie_convolve_filts_f32(const float* burst, int burst_pitch, const float* filts, float* out, int n, int h,
                                     int w, int T, int K, void* stream) {
  using namespace ie;
  IE_REQUIRE(burst && filts && out, "convolve_filts: null pointer");
  IE_REQUIRE(n > 0 && h > 0 && w > 0 && T >= 1 && T <= kConvMaxT, "convolve_filts: bad sizes (T=%d, max %d)", T, kConvMaxT);
  IE_REQUIRE(K >= 1 && (K & 1) && K <= 31 && burst_pitch >= T, "convolve_filts: bad K=%d / burst_pitch=%d", K, burst_pitch);
  const long long npix = (long long)n * h * w;
  const long long blocks = (npix + 7) / 8;
  IE_REQUIRE(blocks < (1ll << 31), "convolve_filts: too many pixels");
  convolve_filts_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(burst, burst_pitch, filts, out,
                                                                                         npix, h, w, T, K);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

// =================================================================================================
// Tensor-core variant of the fused filter (K = 15; one launch = 16 bases x 4 frames, more of either as further launches): G = Bas (*) burst as warp-level
// mma.sync m16n8k8 with TF32 operands and fp32 accumulation.
//
// Why mma.sync and not tcgen05: the A operand of this GEMM is the 225-tap im2col of a ONE-channel image, i.e. a
// Toeplitz matrix A[px][j] = row[px + j] whose rows are 4 bytes apart.  A UMMA shared-memory descriptor cannot
// express that pitch (core-matrix rows are 16 bytes apart), and materialising the im2col costs 900 B of shared
// memory per pixel and frame - more time than the fp32 kernel above needs in total.  With register operands the
// Toeplitz structure is free: thread (g, c) of a warp loads row[x + g + c + 4q] straight from the burst tile, and
// the same 18 loads serve all four 16-pixel M tiles and both K steps of a filter row.  Measured on B200: the
// legacy tensor path issues one m16n8k8 TF32 MMA per 8 cycles per scheduler = 512 MAC/clk/SM, 4x the FP32 pipe.
//
// Precision: burst and basis are rounded to TF32 (10-bit mantissa, cvt.rna) when they are staged; products are
// exact, accumulation and the coefficient mix are fp32.  The output is a convex combination of burst pixels, so
// the error is bounded by 2^-10 of the pixel range (measured < 2e-4 on [0,1] inputs) - inside the path's
// tolerance (max-abs 1e-2, PSNR 0.05 dB), but not the 1e-5 of ie_kpn_apply_f32, which stays available.
//
// Work split: a block owns a 128-px-wide column of one image over a range of rows and walks it in groups of 4
// rows (8 warps = 4 rows x 2 halves of 64 px = 4 M tiles each).  The basis lives in shared memory as ready-made
// B fragments for the whole block lifetime; the burst rows live in an 18-row ring (4 rows + K-1 halo) of which
// only the 4 new rows are loaded per group.
// =================================================================================================
namespace ie {

// ring: the 4 rows of the group being filtered + halo; pitch 144 floats
constexpr int kTfK = 15, kTfRing = 4 + kTfK - 1, kTfTileW = 128, kTfSw = kTfTileW + kTfK - 1 + 2;
constexpr int kTfMaxT = 4, kTfMaxB = 128;

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 4-byte asynchronous global->shared copy; src_bytes = 0 zero-fills (tf.pad, model_library.py:126)
__device__ __forceinline__ void cp_async4_tf(float* smem_dst, const float* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
               "r"(src_bytes)
               : "memory");
}

struct KpnTfParams {
  int H, W, Hc, Wc, T, B, burst_pitch;   // T = frames, B = bases (<= 16) handled by this launch
  int t0, Ttot, accumulate;              // first frame, frames of the burst, add the frame sum to out[...,0]
  int b0, Btot, accumulate_frames;       // first basis, bases of the model, add to out[...,1+t] (basis chunks > 0)
  int col_tiles, row_splits, rows_per_block;
};

// One output row segment of MT 16-px M tiles, all frames: the body of a warp's work in a row group.
// NT = N tiles of 8 bases: 2 for up to 16 bases, 1 when the launch's bases fit one tile (B <= 8: half the MMAs).
template <int MT, int NT>
__device__ __forceinline__ void kpn_tf32_row(const float2* __restrict__ s_bfrag, const float* __restrict__ s_ring,
                                             const float* __restrict__ coef, float* __restrict__ out,
                                             const KpnTfParams& P, int img, int y, int xpx, int s0, int xw, int lane) {
  const int T = P.T, B = P.B, W = P.W;
  const int g = lane >> 2, c = lane & 3;
  const float fT = (float)P.Ttot;
  float dsum[MT][2];
  float cf[MT][2][4];                       // coef[px][2c, 2c+1, 8+2c, 9+2c] of this thread's 2*MT pixels
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    dsum[m][0] = dsum[m][1] = 0.f;
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int px = xpx + 16 * m + g + 8 * hrow;
      const float* cp = coef + (((long long)img * P.Hc + y) * P.Wc + (px < W ? px : 0)) * P.Btot + P.b0;
#pragma unroll
      for (int q = 0; q < 2 * NT; ++q) {
        const int b = (q >> 1) * 8 + 2 * c + (q & 1);
        cf[m][hrow][q] = (px < W && b < B) ? __ldg(cp + b) : 0.f;
      }
    }
  }
  for (int t = 0; t < T; ++t) {
    float acc[MT][NT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[m][nt][e] = 0.f;
    const float2* bf_t = s_bfrag + (t * kTfK) * 128 + lane;
    const float* ring_t = s_ring + t * kTfRing * kTfSw + xw + g + c;
    for (int i = 0; i < kTfK; ++i) {
      int sl = s0 + i;
      if (sl >= kTfRing) sl -= kTfRing;
      const float* rp = ring_t + sl * kTfSw;
      uint32_t w[4 * MT + 2];
#pragma unroll
      for (int q = 0; q < 4 * MT + 2; ++q) w[q] = __float_as_uint(rp[4 * q]);            // already rounded to TF32
      const float2 b00 = bf_t[(i * 4 + 0) * 32], b10 = bf_t[(i * 4 + 2) * 32];   // nt 0: ks 0, 1
      float2 b01 = b00, b11 = b10;                                                // nt 1: ks 0, 1
      if (NT == 2) {
        b01 = bf_t[(i * 4 + 1) * 32];
        b11 = bf_t[(i * 4 + 3) * 32];
      }
      // A[r][k] = row[16 m + r + 8 ks + k]:  a0 (g, c), a1 (g+8, c), a2 (g, c+4), a3 (g+8, c+4).
      // All 2*MT accumulators of K step 0 first, then K step 1: dependent MMAs are 2*MT issues apart.
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        mma_tf32(acc[m][0], w[4 * m], w[4 * m + 2], w[4 * m + 1], w[4 * m + 3], __float_as_uint(b00.x), __float_as_uint(b00.y));
        if (NT == 2)
          mma_tf32(acc[m][NT - 1], w[4 * m], w[4 * m + 2], w[4 * m + 1], w[4 * m + 3], __float_as_uint(b01.x), __float_as_uint(b01.y));
      }
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        mma_tf32(acc[m][0], w[4 * m + 2], w[4 * m + 4], w[4 * m + 3], w[4 * m + 5], __float_as_uint(b10.x), __float_as_uint(b10.y));
        if (NT == 2)
          mma_tf32(acc[m][NT - 1], w[4 * m + 2], w[4 * m + 4], w[4 * m + 3], w[4 * m + 5], __float_as_uint(b11.x), __float_as_uint(b11.y));
      }
    }
    // mix with the per-pixel coefficients: thread holds G[px][2c, 2c+1] (nt 0) and G[px][8+2c, 9+2c] (nt 1)
    // for px = 16 m + g (e = 0, 1) and 16 m + g + 8 (e = 2, 3); the 4 lanes of a row add up through shuffles
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int px = xpx + 16 * m + g + 8 * hrow;
        float s = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) s = fmaf(cf[m][hrow][nt * 2 + e], acc[m][nt][2 * hrow + e], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (c == 0 && px < W) {
          float* of = out + (((long long)img * P.H + y) * W + px) * (P.Ttot + 1) + 1 + P.t0 + t;
          *of = P.accumulate_frames ? fmaf(s, fT, *of) : s * fT;
        }
        dsum[m][hrow] += s;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int px = xpx + 16 * m + g + 8 * hrow;
      if (c == 0 && px < W) {
        float* o0 = out + (((long long)img * P.H + y) * W + px) * (P.Ttot + 1);
        *o0 = P.accumulate ? *o0 + dsum[m][hrow] : dsum[m][hrow];
      }
    }
}

__global__ void __launch_bounds__(256, 2)
kpn_apply_tf32_kernel(const float* __restrict__ burst, const float* __restrict__ coef, const float* __restrict__ bas,
                      float* __restrict__ out, const KpnTfParams P) {
  extern __shared__ float smem[];
  const int T = P.T, B = P.B, H = P.H, W = P.W;
  float2* s_bfrag = reinterpret_cast<float2*>(smem);                       // [T][15][2 ks][2 nt][32 lanes]
  float* s_ring = smem + T * kTfK * 2 * 2 * 32 * 2;                        // [T][18][kTfSw]

  int bid = blockIdx.x;
  const int ct = bid % P.col_tiles; bid /= P.col_tiles;
  const int rs = bid % P.row_splits;
  const int img = bid / P.row_splits;
  const int x0 = ct * kTfTileW;
  const int yb0 = rs * P.rows_per_block, yb1 = min(yb0 + P.rows_per_block, H);
  if (yb0 >= H) return;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* burst_img = burst + (long long)img * H * W * P.burst_pitch;
  const float* bas_img = bas + (long long)img * kTfK * kTfK * P.Ttot * P.Btot;

  // ---- basis -> B fragments (once per block).  Fragment (t, i, ks, nt), lane (g, c):
  //      b0 = Bas[i][j = 8 ks + c][t][n = 8 nt + g],  b1 = Bas[i][j = 8 ks + c + 4][t][n]   (zero for j >= 15, n >= B)
  // The image's basis [tap][t][b] is read in memory order (coalesced) and scattered to its fragment slot.
  {
    float* bf = reinterpret_cast<float*>(s_bfrag);
    for (int idx = threadIdx.x; idx < T * kTfK * 2 * 2 * 32 * 2; idx += blockDim.x) bf[idx] = 0.f;
    __syncthreads();
    const int tb = T * B;                      // this launch's frames [t0, t0+T) x bases [b0, b0+B) of every tap
    const float inv_tb = 1.f / (float)tb, inv_b = 1.f / (float)B;
    for (int idx = threadIdx.x; idx < kTfK * kTfK * tb; idx += blockDim.x) {
      const int tap = (int)(((float)idx + 0.5f) * inv_tb), r = idx - tap * tb;
      const int t = (int)(((float)r + 0.5f) * inv_b), n = r - t * B;
      const int i = tap / kTfK, j = tap - i * kTfK;
      const int ks = j >> 3, cpos = j & 7, cc = cpos & 3, half = cpos >> 2;
      const int nt = n >> 3, gg = n & 7;
      bf[(((((t * kTfK + i) * 2 + ks) * 2 + nt) * 32) + gg * 4 + cc) * 2 + half] =
          to_tf32(__ldg(bas_img + ((long long)tap * P.Ttot + P.t0 + t) * P.Btot + P.b0 + n));
    }
  }

  // stages burst rows [gy0, gy1) of every frame into the ring (slot = (gy + 7) mod 18), rounded to TF32, zero outside
  // the image.  (cp.async into a deeper ring was tried: the copy cannot round, and cvt.rna.tf32 is a 4-instruction
  // sequence on sm_100 - converting the A fragments at load time cost more than the latency it hid.)
  auto stage_rows = [&](int gy0, int gy1) {
    const int run = (kTfTileW + kTfK - 1) * P.burst_pitch;
    const float inv_run = 1.f / (float)run, inv_pitch = 1.f / (float)P.burst_pitch;
    const int nrows = gy1 - gy0;
    // four independent loads in flight per thread before the first conversion / store
    for (int idx0 = threadIdx.x; idx0 < nrows * run; idx0 += 4 * blockDim.x) {
      float v[4];
      int dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = idx0 + u * blockDim.x;
        v[u] = 0.f;
        dst[u] = -1;
        if (idx < nrows * run) {
          const int r = (int)(((float)idx + 0.5f) * inv_run), e = idx - r * run;
          const int lx = (int)(((float)e + 0.5f) * inv_pitch), ch = e - lx * P.burst_pitch;
          if (ch < T) {                       // (channels t0 + ch of the pixel are read below)
            const int gy = gy0 + r, gx = x0 - kTfK / 2 + lx;
            dst[u] = (ch * kTfRing + (gy + kTfK / 2) % kTfRing) * kTfSw + lx;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W)
              v[u] = __ldg(burst_img + ((long long)gy * W + gx) * P.burst_pitch + P.t0 + ch);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (dst[u] >= 0) s_ring[dst[u]] = to_tf32(v[u]);
    }
  };
  // the 2 pad columns of every ring row are read by the A loads of the last M tile (never used for a valid pixel)
  for (int idx = threadIdx.x; idx < T * kTfRing * 2; idx += blockDim.x)
    s_ring[(idx / 2) * kTfSw + kTfTileW + kTfK - 1 + idx % 2] = 0.f;

  const int wr = warp >> 1;                  // row of the group this warp filters
  const int xw = (warp & 1) * 64;            // its 64-px half of the tile
  const int mt = min(4, (W - x0 - xw + 15) >> 4);                         // M tiles of this warp that hold real pixels

  for (int yg = yb0; yg < yb1; yg += 4) {
    __syncthreads();                                                      // previous group done with the ring
    if (yg == yb0) stage_rows(yg - kTfK / 2, yg + 4 + kTfK / 2);
    else stage_rows(yg + kTfK / 2, yg + 4 + kTfK / 2);                    // 4 new rows
    __syncthreads();
    const int y = yg + wr;
    if (y >= yb1) continue;
    const int s0 = y % kTfRing;                                           // slot of burst row y - 7 (filter row 0)
    // the warp's 64 px as two halves of 2 M tiles: the per-pixel coefficients of a half stay in registers across
    // the frames (they were re-loaded per frame at first: a third of the stall samples were their latency)
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int m_left = mt - 2 * half;                                   // warp-uniform
      const int xs = xw + 32 * half;
      if (B > 8) {                                                        // block-uniform
        if (m_left >= 2) kpn_tf32_row<2, 2>(s_bfrag, s_ring, coef, out, P, img, y, x0 + xs, s0, xs, lane);
        else if (m_left == 1) kpn_tf32_row<1, 2>(s_bfrag, s_ring, coef, out, P, img, y, x0 + xs, s0, xs, lane);
      } else {
        if (m_left >= 2) kpn_tf32_row<2, 1>(s_bfrag, s_ring, coef, out, P, img, y, x0 + xs, s0, xs, lane);
        else if (m_left == 1) kpn_tf32_row<1, 1>(s_bfrag, s_ring, coef, out, P, img, y, x0 + xs, s0, xs, lane);
      }
    }
  }
}

}  // namespace ie

extern "C" int ie_kpn_apply_tf32(const float* burst, int burst_pitch, const float* coef, int hc, int wc, const float* bas,
                                 float* out, int n, int h, int w, int T, int K, int B, void* stream) {
  using namespace ie;
  IE_REQUIRE(burst && coef && bas && out, "kpn_apply_tf32: null pointer");
  IE_REQUIRE(n > 0 && h > 0 && w > 0, "kpn_apply_tf32: bad sizes");
  IE_REQUIRE(K == kTfK && B >= 1 && B <= kTfMaxB && T >= 1 && T <= 2 * kTfMaxT,
             "kpn_apply_tf32: built for K = 15, B <= %d, T <= %d (got K=%d B=%d T=%d); use ie_kpn_apply_f32", kTfMaxB,
             2 * kTfMaxT, K, B, T);
  IE_REQUIRE(burst_pitch >= T && hc >= h && wc >= w, "kpn_apply_tf32: bad pitch / coef extent");
  KpnTfParams P{};
  P.H = h; P.W = w; P.Hc = hc; P.Wc = wc; P.Btot = B; P.burst_pitch = burst_pitch; P.Ttot = T;
  P.col_tiles = (w + kTfTileW - 1) / kTfTileW;
  // blocks: ~4 per SM so that the last wave is short, but at least 4 row groups each (the basis fragments and the
  // 14 halo rows are staged once per block)
  const long long base = (long long)n * P.col_tiles;
  const int groups = (h + 3) / 4;
  long long splits = (4ll * sm_count() + base - 1) / base;
  if (splits > groups / 4) splits = groups / 4;
  if (splits < 1) splits = 1;
  P.rows_per_block = (int)(((groups + splits - 1) / splits) * 4);
  P.row_splits = (h + P.rows_per_block - 1) / P.rows_per_block;
  const long long blocks = base * P.row_splits;
  IE_REQUIRE(blocks < (1ll << 31), "kpn_apply_tf32: too many blocks");
  // the basis fragments of 4 frames x 16 bases fill the shared memory: longer bursts run as two passes over the
  // frames, the second adding its frame sum to out[...,0]; more bases (Basis_kpn of the remote/ configs: up to 90)
  // run as chunks of 16 - the output is a sum over the bases - each later chunk adding to every output channel
  for (int b0 = 0; b0 < B; b0 += 16) {
    P.b0 = b0;
    P.B = (B - b0 < 16) ? B - b0 : 16;
    P.accumulate_frames = b0 > 0;
    for (int t0 = 0; t0 < T; t0 += kTfMaxT) {
      P.t0 = t0;
      P.T = (T - t0 < kTfMaxT) ? T - t0 : kTfMaxT;
      P.accumulate = t0 > 0 || b0 > 0;
      const size_t smem = sizeof(float) * ((size_t)P.T * kTfK * 2 * 2 * 32 * 2 + (size_t)P.T * kTfRing * kTfSw);
      IE_CUDA(cudaFuncSetAttribute(kpn_apply_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kpn_apply_tf32_kernel<<<(unsigned)blocks, 256, smem, static_cast<cudaStream_t>(stream)>>>(burst, coef, bas, out, P);
      IE_LAUNCH_CHECK();
    }
  }
  IE_LAUNCH_CHECK();
  return IE_OK;
}
