// Synthetic-burst preprocessing arithmetic of /root/reference/data_utils.py:198-265 (with helpers
// :432-466) as one bandwidth kernel: uint8 -> (v/255)^degamma -> per-frame crop -> up x up AREA
// box mean -> channel mean -> white level -> read/shot noise -> noise-level channel -> NHWC.
// The reference draws crops / levels / normals from TF's RNG; here every draw is an input.
#include "ie_common.cuh"
#include "ie_ptx.cuh"

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

// ------------------------------------------------------------------------------- 4 px per thread (reference defaults)
// upscale 4, grey source, w % 4 == 0, T in {4, 8}: a thread owns FOUR consecutive output pixels, i.e. per frame and
// source row one run of 16 source bytes.  What bounded the per-pixel kernel above was the de-gamma LUT: 64 lookups per
// output pixel at random addresses (3.5-way bank conflicts on average) and 32 scalar 4-byte fetches.  Here
//   * the LUT is replicated 32x with a 256-byte entry stride (64 KB, half used): lane l only ever reads bank l ->
//     conflict-free, and ONE byte-permute builds the address  sample * 256 + lane * 4  from the packed word;
//   * the 16 bytes come from two aligned 128-bit loads; the run's byte offset inside them is the same for all threads
//     of a row (thread stride = 16 bytes), so the word selection runs on warp-uniform predicates (11 selects + 4 funnel shifts);
//   * noise inputs and both outputs move as 128-bit vectors (4 px x T, 4 px x (T+add), 4 px x 2 floats are contiguous).
// Summation order is that of the kernel above (bit-identical truth); the noise terms use MUFU.SQRT (~1 ulp).
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// lut_b = table base (bytes); lane4 = lane * 4.  One PRMT builds the byte offset  b * 256 + lane * 4  of sample b.
__device__ __forceinline__ float lut1(const char* lut_b, uint32_t off) { return *reinterpret_cast<const float*>(lut_b + off); }
__device__ __forceinline__ float lut4(const char* lut_b, uint32_t lane4, uint32_t v) {
  const float a = lut1(lut_b, __byte_perm(v, lane4, 0x5504));
  const float b = lut1(lut_b, __byte_perm(v, lane4, 0x5514));
  const float c = lut1(lut_b, __byte_perm(v, lane4, 0x5524));
  const float d = lut1(lut_b, __byte_perm(v, lane4, 0x5534));
  return (a + b) + (c + d);
}
// The 16 bytes at byte offset `mis` of the 32-byte window (lo, hi): two select stages pick the 5 words they span
// (11 selects on 2 warp-uniform predicates; a 4-way switch was if-converted into 4 x 4 predicated funnel shifts).
__device__ __forceinline__ void row16(const uint4& lo, const uint4& hi, int mis, const char* lut_b, uint32_t lane4,
                                      float (&s)[4]) {
  const bool s2 = (mis & 8) != 0, s1 = (mis & 4) != 0;
  const int sh = (mis & 3) * 8;
  const uint32_t t0 = s2 ? lo.z : lo.x, t1 = s2 ? lo.w : lo.y, t2 = s2 ? hi.x : lo.z, t3 = s2 ? hi.y : lo.w,
                 t4 = s2 ? hi.z : hi.x, t5 = s2 ? hi.w : hi.y;
  const uint32_t u0 = s1 ? t1 : t0, u1 = s1 ? t2 : t1, u2 = s1 ? t3 : t2, u3 = s1 ? t4 : t3, u4 = s1 ? t5 : t4;
  s[0] += lut4(lut_b, lane4, __funnelshift_r(u0, u1, sh));
  s[1] += lut4(lut_b, lane4, __funnelshift_r(u1, u2, sh));
  s[2] += lut4(lut_b, lane4, __funnelshift_r(u2, u3, sh));
  s[3] += lut4(lut_b, lane4, __funnelshift_r(u3, u4, sh));
}

// Border quads (a window leaves the source): per-sample bounds checks, zero padding of make_first_truth (:436-438).
__device__ __noinline__ float4 box16_clipped(const uint8_t* __restrict__ img, int hs, int ws, int oy, int ox,
                                             const char* lut_b, uint32_t lane4) {
  float s[4];
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    float a = 0.f;
#pragma unroll 1
    for (int dy = 0; dy < 4; ++dy) {
      const int sy = oy + dy;
      if (sy < 0 || sy >= hs) continue;
#pragma unroll 1
      for (int dx = 0; dx < 4; ++dx) {
        const int sx = ox + p * 4 + dx;
        if (sx < 0 || sx >= ws) continue;
        a += lut1(lut_b, ((uint32_t)img[(long long)sy * ws + sx] << 8) + lane4);
      }
    }
    s[p] = a;
  }
  return make_float4(s[0], s[1], s[2], s[3]);
}

// 1-D TMA bulk copies (global <-> shared, no LSU wavefronts, no registers)
__device__ __forceinline__ void bulk_load(float* smem_dst, const float* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(float* gdst, const float* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

// Every WARP walks its own tiles of 32 consecutive quads (128 pixels, contiguous in every tensor).  What limited the
// first 4-px-per-thread version was the L1 data pipe (77 % busy): besides the 64 table lookups per pixel it carried the
// noise loads and the output stores as 128-bit accesses with a 64 / 80 / 32-byte lane stride, i.e. 16-20 wavefronts
// per instruction.  Here the tile's read / shot noise arrives by TMA bulk copy into the warp's slice of shared memory
// (the next tile's copy is in flight during the lookups) and x / truth leave by TMA bulk store from a staging slice;
// the threads touch them only with conflict-free 128-bit shared accesses (noise: lane-rotated order, un-rotated with
// selects).  Warps never meet at a block barrier after the table is built (a block-wide version of this pipeline was
// barrier- and latency-bound: one 512-thread block per SM, 0.9 stalled warps per issue at the barriers).
template <int T, int ADD, bool RNG, int NT>
__global__ void __launch_bounds__(NT)
preprocess_u8_tile_kernel(const uint8_t* __restrict__ src, int hs, int ws, const int32_t* __restrict__ org,
                          float degamma, const float* __restrict__ wl, const float* __restrict__ sig_read,
                          const float* __restrict__ sig_shot, const float* __restrict__ n_read,
                          const float* __restrict__ n_shot, unsigned long long seed, int n_img, int h, int w,
                          float* __restrict__ x, float* __restrict__ truth) {
  constexpr int CH = T + ADD;
  constexpr int ROTBITS = (T == 4) ? 2 : 3;
  static_assert(T == 4 || T == 8, "tile kernel: T must be 4 or 8");
  extern __shared__ float4 pre_dyn4[];
  float* lutr = reinterpret_cast<float*>(pre_dyn4);              // [256][64]: entry b, copy l at float b * 64 + l (l < 32)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = NT / 32;
  float* s_ns = lutr + 256 * 64 + warp * 32 * 4 * T;             // [NW][32][4 T] shot normals of the warp's tile
  float* s_nr = lutr + 256 * 64 + NT * 4 * T + warp * 32 * 4 * T;   // read normals
  float* s_x = lutr + 256 * 64 + 2 * NT * 4 * T + warp * 32 * 4 * CH;          // [NW][32][4 CH] staged x
  float* s_t = lutr + 256 * 64 + 2 * NT * 4 * T + NT * 4 * CH + warp * 32 * 8;  // [NW][32][8]    staged truth
  __shared__ uint64_t bars[NW];
  if (threadIdx.x < 256) {
    const float v = powf((float)threadIdx.x / 255.f, degamma);
    const int l = threadIdx.x & 31;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) lutr[threadIdx.x * 64 + ((l + j) & 31)] = v;
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NW; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  uint64_t* bar = &bars[warp];
  const char* lut_b = reinterpret_cast<const char*>(lutr);
  const uint32_t lane4 = lane * 4;
  const int n = blockIdx.z;
  const float wln = wl[n], sr = sig_read[n], ss = sig_shot[n];
  const long long img_off = (long long)n * hs * ws;
  const long long src_total = (long long)n_img * hs * ws;
  const int qw = w >> 2;
  const int nquads = h * qw;
  const long long img_px = (long long)n * h * w;
  const float inv_area = 1.f / 16.f;
  const bool have_noise = !RNG && (n_read != nullptr) && (n_shot != nullptr);
  int oys[T], oxs[T];
#pragma unroll
  for (int f = 0; f < T; ++f) {
    oys[f] = org[((long long)n * T + f) * 2];
    oxs[f] = org[((long long)n * T + f) * 2 + 1];
  }
  auto load_noise = [&](int q0) {                                // lane 0 only
    const uint32_t bytes = (uint32_t)min(32, nquads - q0) * 4u * T * 4u;
    mbar_arrive_expect_tx(bar, 2 * bytes);
    bulk_load(s_ns, n_shot + (img_px + 4ll * q0) * T, bytes, bar);
    bulk_load(s_nr, n_read + (img_px + 4ll * q0) * T, bytes, bar);
  };
  const int q_first = (blockIdx.x * NW + warp) * 32, q_step = gridDim.x * NW * 32;
  if (have_noise && lane == 0 && q_first < nquads) load_noise(q_first);
  // reads the thread's 4T noise floats (element p * T + f) of one tensor: unit u = float4 number u; slot j reads unit
  // (j + r) & (T-1) so that the 8 lanes of a shared-memory wavefront cover all 32 banks, then the rotation is undone
  // with log2(T) select stages
  auto read_noise = [&](const float* sbase, float (&out)[4 * T]) {
    const int r = (T == 4) ? ((lane >> 1) & 3) : (lane & 7);
    float4 a[T], a2[T];
#pragma unroll
    for (int j = 0; j < T; ++j) a[j] = reinterpret_cast<const float4*>(sbase + lane * 4 * T)[(j + r) & (T - 1)];
#pragma unroll
    for (int bit = 0; bit < ROTBITS; ++bit) {
      const bool on = (r >> bit) & 1;
#pragma unroll
      for (int u = 0; u < T; ++u) {
        const float4 o = a[(u - (1 << bit)) & (T - 1)], k = a[u];
        a2[u] = make_float4(on ? o.x : k.x, on ? o.y : k.y, on ? o.z : k.z, on ? o.w : k.w);
      }
#pragma unroll
      for (int u = 0; u < T; ++u) a[u] = a2[u];
    }
#pragma unroll
    for (int u = 0; u < T; ++u) { out[4 * u] = a[u].x; out[4 * u + 1] = a[u].y; out[4 * u + 2] = a[u].z; out[4 * u + 3] = a[u].w; }
  };
  int it = 0;
  for (int q0 = q_first; q0 < nquads; q0 += q_step, ++it) {
    const int q = q0 + lane;
    const bool active = q < nquads;
    const int py = q / qw, px = (q - py * qw) * 4;
    const long long t0 = img_px + 4ll * q;
    float xv[4 * CH], tv[8];

    // ---- phase 1: clean frames.  The source rows of frame f + 1 are requested right after frame f's 64 lookups.
    if (active) {
      auto in_range = [&](int f) {
        const int oy = oys[f] + py * 4, ox = oxs[f] + px * 4;
        const long long off_last = img_off + (long long)(oy + 3) * ws + ox;
        return oy >= 0 && oy + 4 <= hs && ox >= 0 && ox + 16 <= ws && (off_last & ~15ll) + 32 <= src_total;
      };
      auto fetch = [&](int f, uint4 (&buf)[8]) {
        const int oy = oys[f] + py * 4, ox = oxs[f] + px * 4;
#pragma unroll
        for (int dy = 0; dy < 4; ++dy) {
          const long long off = img_off + (long long)(oy + dy) * ws + ox;
          const uint4* p16 = reinterpret_cast<const uint4*>(src + (off & ~15ll));
          buf[2 * dy] = __ldg(p16);
          buf[2 * dy + 1] = __ldg(p16 + 1);
        }
      };
      uint4 cur[8];
      bool fast_cur = in_range(0);
      if (fast_cur) fetch(0, cur);
#pragma unroll
      for (int f = 0; f < T; ++f) {
        const int oy = oys[f] + py * 4, ox = oxs[f] + px * 4;
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        if (fast_cur) {
#pragma unroll
          for (int dy = 0; dy < 4; ++dy) {
            const long long off = img_off + (long long)(oy + dy) * ws + ox;
            row16(cur[2 * dy], cur[2 * dy + 1], (int)(off & 15), lut_b, lane4, s);
          }
        } else {
          const float4 b = box16_clipped(src + img_off, hs, ws, oy, ox, lut_b, lane4);
          s[0] = b.x; s[1] = b.y; s[2] = b.z; s[3] = b.w;
        }
        if (f + 1 < T) {                                         // (requesting them before the lookups needs 32 more
          fast_cur = in_range(f + 1);                            //  registers: spills at 512 threads, slower at 384)
          if (fast_cur) fetch(f + 1, cur);
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float csum = s[p] * inv_area;
          xv[p * CH + f] = wln * (csum / 1.f);                   // :230 (grey source: channel mean of one channel)
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) { tv[2 * p] = xv[p * CH]; tv[2 * p + 1] = wln; }   // :248-252
    }

    // ---- phase 2: noise (data_utils.py:462-466).  The tile's normals were requested a whole tile ago.
    if (have_noise) {
      float nsf[4 * T], nrf[4 * T];
      mbar_wait(bar, it & 1);
      read_noise(s_ns, nsf);
      read_noise(s_nr, nrf);
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int f = 0; f < T; ++f) {
          const float tr = xv[p * CH + f];
          xv[p * CH + f] = tr + sqrt_approx(tr) * ss * nsf[p * T + f] + sr * nrf[p * T + f];
        }
    } else if (RNG) {
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int f = 0; f < T; ++f) {
          const float tr = xv[p * CH + f];
          float zs, zr;
          normal_pair(seed, (unsigned long long)((t0 + p) * T + f), zs, zr);
          xv[p * CH + f] = tr + sqrt_approx(tr) * ss * zs + sr * zr;
        }
    }
    if (ADD == 1) {
#pragma unroll
      for (int p = 0; p < 4; ++p) xv[p * CH + T] = sqrt_approx(sr * sr + fmaxf(0.f, xv[p * CH]) * ss * ss);   // :256
    } else if (ADD == 2) {
#pragma unroll
      for (int p = 0; p < 4; ++p) { xv[p * CH + T] = sr; xv[p * CH + T + 1] = ss; }                     // :257
    }
    // the previous tile's bulk stores must have read the staging slice before it is overwritten
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();                                                // noise slice consumed, staging slice free
    if (have_noise && lane == 0 && q0 + q_step < nquads) load_noise(q0 + q_step);
    if (active) {
      float4* xo = reinterpret_cast<float4*>(s_x + lane * 4 * CH);
#pragma unroll
      for (int j = 0; j < CH; ++j) xo[j] = make_float4(xv[4 * j], xv[4 * j + 1], xv[4 * j + 2], xv[4 * j + 3]);
      float4* to = reinterpret_cast<float4*>(s_t + lane * 8);
      to[0] = make_float4(tv[0], tv[1], tv[2], tv[3]);
      to[1] = make_float4(tv[4], tv[5], tv[6], tv[7]);
    }
    fence_proxy_async_smem();                                    // generic-proxy writes -> visible to the bulk store
    __syncwarp();
    if (lane == 0) {
      const uint32_t nq = (uint32_t)min(32, nquads - q0);
      bulk_store(x + (img_px + 4ll * q0) * CH, s_x, nq * 4u * CH * 4u);
      bulk_store(truth + (img_px + 4ll * q0) * 2, s_t, nq * 8u * 4u);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static int g_pre_legacy = 0;
template <int T, int ADD>
static int launch_quad(const uint8_t* src, int n, int hs, int ws, const int32_t* org, float degamma, const float* wl,
                       const float* sig_read, const float* sig_shot, const float* n_read, const float* n_shot,
                       unsigned long long seed, int use_rng, int h, int w, float* x, float* truth, cudaStream_t st) {
  constexpr int NT = (T == 4) ? 512 : 256;                     // one block per SM (the table alone is 64 KB)
  constexpr int CH = T + ADD;
  const size_t smem = sizeof(float) * (256 * 64 + 2 * NT * 4 * T + NT * 4 * CH + NT * 8);
  auto kern = use_rng ? preprocess_u8_tile_kernel<T, ADD, true, NT> : preprocess_u8_tile_kernel<T, ADD, false, NT>;
  const int nt = NT;                                          // (576 / 640 threads measured the same)
  IE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nquads = (long long)h * (w / 4);
  long long gx = (nquads + NT - 1) / NT;
  long long cap = sm_count() / n;                              // persistent: the grid fills the SMs once, warps stride
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  kern<<<dim3((unsigned)gx, 1, n), nt, smem, st>>>(src, hs, ws, org, degamma, wl, sig_read, sig_shot, n_read, n_shot, seed,
                                                   n, h, w, x, truth);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

// Returns 1 if the quad kernel took the call, 0 if the caller must use the general kernel, <0 on error.
static int try_quad(const uint8_t* src, int n, int hs, int ws, int c, const int32_t* org, int up, float degamma,
                    const float* wl, const float* sig_read, const float* sig_shot, const float* n_read,
                    const float* n_shot, unsigned long long seed, int use_rng, int layer_type, int h, int w, int T,
                    float* x, float* truth, cudaStream_t st) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(x) |
                       reinterpret_cast<uintptr_t>(truth) | reinterpret_cast<uintptr_t>(n_read) |
                       reinterpret_cast<uintptr_t>(n_shot);
  if (g_pre_legacy || up != 4 || c != 1 || (w & 3) || (al & 15) || (long long)h * (w / 4) > 0x7fffffffll) return 0;
  int rc = 0;
#define IE_QUAD(TT, AA)                                                                                              \
  if (T == TT && layer_type == AA) {                                                                                 \
    rc = launch_quad<TT, AA>(src, n, hs, ws, org, degamma, wl, sig_read, sig_shot, n_read, n_shot, seed, use_rng, h, \
                             w, x, truth, st);                                                                       \
    return rc == IE_OK ? 1 : rc;                                                                                     \
  }
  IE_QUAD(4, 0) IE_QUAD(4, 1) IE_QUAD(4, 2) IE_QUAD(8, 0) IE_QUAD(8, 1) IE_QUAD(8, 2)
#undef IE_QUAD
  return 0;
}

}  // namespace ie

extern "C" int ie_preprocess_tune(int legacy) {
  ie::g_pre_legacy = legacy;
  return IE_OK;
}

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
  {
    const int q = try_quad(src, n, hs, ws, c, org, up, degamma, wl, sig_read, sig_shot, n_read, n_shot, 0ull, 0,
                           layer_type, h, w, T, x, truth, static_cast<cudaStream_t>(stream));
    if (q != 0) return q < 0 ? q : IE_OK;
  }
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
  {
    const int q = try_quad(src, n, hs, ws, c, org, up, degamma, wl, sig_read, sig_shot, nullptr, nullptr, seed, 1,
                           layer_type, h, w, T, x, truth, static_cast<cudaStream_t>(stream));
    if (q != 0) return q < 0 ? q : IE_OK;
  }
  preprocess_u8_kernel<<<dim3(ie_ceil_div(w, 256), h, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, hs, ws, c, org, up, degamma, wl, sig_read, sig_shot, nullptr, nullptr, seed, 1, layer_type, h, w, T, x, truth,
      total);
  IE_LAUNCH_CHECK();
  return IE_OK;
}
