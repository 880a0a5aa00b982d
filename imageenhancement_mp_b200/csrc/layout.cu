// Bandwidth kernels around the convolutions: weight / input packing, max-pool, bilinear up-sampling
// into concat slices, per-image channel means, broadcast, raster -> NHWC, basis softmax.
// All bf16 raster traffic moves as 16-byte vectors (8 channels); grids cover the tensor exactly.
#include <cuda_bf16.h>

#include "ie_common.cuh"
#include "ie_ptx.cuh"

namespace ie {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------- weights
__global__ void pack_weights_kernel(const float* __restrict__ hwio, int taps, int cin, int cout, int ktot_pad,
                                    __nv_bfloat16* __restrict__ out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)cout * ktot_pad;
  if (idx >= total) return;
  const int o = (int)(idx / ktot_pad);
  const int k = (int)(idx - (long long)o * ktot_pad);
  float v = 0.f;
  if (k < taps * cin) v = hwio[(long long)k * cout + o];   // HWIO flat index = ((i*kw+j)*cin + c)*cout + o
  out[idx] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------- input im2col
// grid = (ceil((w+1)*kvec / threads), h+1, n): blockIdx.y/z give the raster row and image, so no thread
// divides a 64-bit index.  One thread per (padded x, 16-byte chunk) = 8 consecutive k of the im2col row.
// For filter row dy the 3*c values k = dy*3c + i are CONTIGUOUS in the source: row y+dy-1, floats
// (x-1)*c + i, so no per-element tap arithmetic is needed.
// The source may be smaller than the raster (hs <= h, ws <= w): it is implicitly zero-padded at the
// bottom/right, which is how the boundary pads 100x100 patches to the network stride without a padded
// copy of the input (pixels of the padding still see their real neighbours, exactly like a padded input).
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const float* __restrict__ x, int hs, int ws, int h, int w, int c, uint4* __restrict__ out, int kvec) {
  const int wp = w + 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= wp * kvec) return;
  const int xp = idx / kvec, chunk = idx - xp * kvec;
  const int yp = blockIdx.y, img = blockIdx.z;
  const int y = yp - 1, xx = xp;                          // raster row 0 / column w are the shared zero border
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = 0.f;
  if (y >= 0 && y < h && xx >= 0 && xx < w) {
    const int c3 = 3 * c;
    const int lo = (xx > 0) ? 0 : c;                         // i < lo: left neighbour outside the image
    const int hi = (xx + 1 < ws) ? c3 : ((xx < ws) ? 2 * c : ((xx == ws) ? c : 0));   // i >= hi: right side outside
    // this chunk's 8 values span at most two filter rows: dy0 = first, switch to dy0+1 at k == (dy0+1)*c3
    const int k0 = chunk * 8;
    const int dy0 = (k0 >= c3) + (k0 >= 2 * c3);
    const float* centre = x + (((long long)img * hs + y) * ws + (xx - 1)) * c;      // (y, x-1, ch 0)
    const float* r0 = centre + (long long)(dy0 - 1) * ws * c - dy0 * c3;              // + k gives the element
    const float* r1 = r0 + (long long)ws * c - c3;
    const bool ok0 = (y + dy0 - 1 >= 0) && (y + dy0 - 1 < hs);
    const bool ok1 = (y + dy0 >= 0) && (y + dy0 < hs) && (dy0 < 2);
    const int sw_k = (dy0 + 1) * c3;                          // first k of the next filter row
    if (c3 >= 8) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = k0 + e;
        const bool second = k >= sw_k;
        const int i = k - (second ? sw_k : dy0 * c3);
        const bool ok = (second ? ok1 : ok0) && i >= lo && i < hi && k < 3 * c3;
        if (ok) f[e] = __ldg((second ? r1 : r0) + k);
      }
    } else {                                                  // c <= 2: a chunk may span all three filter rows
      for (int e = 0; e < 8; ++e) {
        const int k = k0 + e;
        const int dy = (k >= c3) + (k >= 2 * c3);
        const int i = k - dy * c3, sy = y + dy - 1;
        if (k < 3 * c3 && sy >= 0 && sy < hs && i >= lo && i < hi) f[e] = __ldg(centre + (long long)(dy - 1) * ws * c + i);
      }
    }
  }
  out[(((long long)img * (h + 1) + yp) * wp + xp) * kvec + chunk] = pack8(f);
}

// Ordered cross-block reduction: after a block has written its partial result it takes a ticket; the block that draws
// the last one sees every partial (fence + L2 loads), reduces them in a fixed order and re-arms the counter.
__device__ __forceinline__ bool stat_finish(int* counter, unsigned nblocks) {
  __shared__ int s_last;
  __syncthreads();                               // the block's partial sums are written (CTA scope)
  if (threadIdx.x == 0) {
    __threadfence();                             // cumulative: publishes them device-wide before the ticket is drawn
    const int ticket = atomicAdd(counter, 1);    // (one fence per block - a fence in every thread also waits for that
    s_last = (ticket == (int)nblocks - 1);       //  thread's output stores and cost 30 us per step in the max-pool)
    if (s_last) *counter = 0;                    // every block of this group has arrived: ready for the next launch
    __threadfence();
  }
  __syncthreads();
  return s_last != 0;
}

// ------------------------------------------------------------------------------- max-pool 2x2
// block = 256 threads = cvb channel chunks (16 B each) x 256/cvb pixel lanes; grid = (cvec/cvb, row groups, n).
// A block walks the padded output pixels of its rows; consecutive threads touch consecutive channel chunks
// (512-byte runs).  Optionally the per-image channel SUMS of the INPUT are accumulated on the way (each input
// pixel belongs to exactly one 2x2 window): this is the GlobalAveragePooling2D of Poolskip
// (model_library.py:110) on the very tensor the pool reads, so the basis branch needs no second pass over it.
constexpr int kPoolRows = 4;           // padded output rows per block (2: twice the tickets of the ordered statistics)
__global__ void __launch_bounds__(256)
maxpool2_kernel(const uint4* __restrict__ x, int h, int w, int cvec, int cvb, int x_pitch_v, int x_coff_v,
                uint4* __restrict__ y, int y_pitch_v, int y_coff_v, float* __restrict__ chan_sum, float scale,
                int stat_y0, int stat_y1, int bi, int bo, float* __restrict__ part, int* __restrict__ counter) {
  pdl_trigger();
  pdl_wait();
  // bi / bo: border of the input / output tensor (1 = shared-border raster, 0 = dense NHWC)
  const int ho = h >> 1, wo = w >> 1, wpo = wo + bo, wpi = w + bi;
  const int cl = threadIdx.x % cvb, pl = threadIdx.x / cvb, npl = blockDim.x / cvb;
  const int cv = blockIdx.x * cvb + cl;
  const int img = blockIdx.z;
  const int oy0 = blockIdx.y * kPoolRows;
  const int rows = min(kPoolRows, ho + bo - oy0);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (cv < cvec) {
    for (int p = pl; p < rows * wpo; p += npl) {
      const int ry = p / wpo, ox = p - ry * wpo;
      const int oyp = oy0 + ry, oy = oyp - bo;              // row of the (padded) output / of the output image
      uint4 res = make_uint4(0, 0, 0, 0);
      if (oy >= 0 && ox < wo) {
        const long long rin = ((long long)img * (h + bi) + (2 * oy + bi)) * wpi + 2 * ox;
        const uint4* q = x + rin * x_pitch_v + x_coff_v + cv;
        float a[8], b[8], c[8], d[8], m[8];
        unpack8(__ldg(q), a);
        unpack8(__ldg(q + x_pitch_v), b);
        unpack8(__ldg(q + (long long)wpi * x_pitch_v), c);
        unpack8(__ldg(q + (long long)(wpi + 1) * x_pitch_v), d);
        // statistics only over input rows [stat_y0, stat_y1) (both even): a spatial shard leaves its halo out
        const bool in_stats = (2 * oy >= stat_y0) && (2 * oy < stat_y1);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          m[e] = fmaxf(fmaxf(a[e], b[e]), fmaxf(c[e], d[e]));
          if (in_stats) acc[e] += (a[e] + b[e]) + (c[e] + d[e]);
        }
        res = pack8(m);
      }
      const long long ro = ((long long)img * (ho + bo) + oyp) * wpo + ox;
      y[ro * y_pitch_v + y_coff_v + cv] = res;
    }
  }
  if (chan_sum == nullptr) return;
  __shared__ float red[256][9];
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = acc[e];
  __syncthreads();
  // thread t < cvb*8 owns channel (t / 8 -> chunk, t % 8 -> element) and sums over the pixel lanes; the row groups
  // of an image meet in a FIXED order: every block stores its partial sums, the last one to arrive (ticket counter)
  // adds them up by row group - no floating-point atomics, the means are bit-reproducible
  const int t = threadIdx.x;
  const int C = cvec * 8;
  const bool owner = (t < cvb * 8) && (blockIdx.x * cvb + (t >> 3) < cvec);
  const int ch = (blockIdx.x * cvb + (t >> 3)) * 8 + (t & 7);
  if (owner) {
    const int chunk = t >> 3, e = t & 7;
    float s = 0.f;
    for (int l = 0; l < npl; ++l) s += red[l * cvb + chunk][e];
    part[((long long)img * gridDim.y + blockIdx.y) * C + ch] = s;
  }
  if (stat_finish(&counter[img * gridDim.x + blockIdx.x], gridDim.y)) {
    if (owner) {
      // loads eight at a time (independent), additions in block order
      float tot = 0.f;
      const int nb = (int)gridDim.y;
      const float* src = part + (long long)img * nb * C + ch;
      int by = 0;
      for (; by + 8 <= nb; by += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (long long)(by + u) * C);
#pragma unroll
        for (int u = 0; u < 8; ++u) tot += v[u];
      }
      for (; by < nb; ++by) tot += __ldcg(src + (long long)by * C);
      chan_sum[(long long)img * C + ch] = tot * scale;
    }
  }
}

// ------------------------------------------------------------------------------- bilinear upsample
// Half-pixel centres: src = (dst + 0.5)/s - 0.5; lower = max(floor(src),0), upper = min(ceil(src), size-1).
// grid = (ceil((wo+1)*cvec / 256), ho+1, n): the vertical taps / weights are block-uniform.
__global__ void __launch_bounds__(256)
upsample_kernel(const uint4* __restrict__ x, int h, int w, int cvec, int x_pitch_v, int x_coff_v, int s,
                uint4* __restrict__ y, int y_pitch_v, int y_coff_v, int bi, int bo) {
  pdl_trigger();
  pdl_wait();
  const int ho = h * s, wo = w * s, wpo = wo + bo;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= wpo * cvec) return;
  const int oxp = idx / cvec, cv = idx - oxp * cvec;
  const int oyp = blockIdx.y, img = blockIdx.z;
  const int oy = oyp - bo, ox = oxp;
  uint4 res = make_uint4(0, 0, 0, 0);
  if (oy >= 0 && oy < ho && ox >= 0 && ox < wo) {
    const float inv = 1.f / (float)s;
    const float sy = ((float)oy + 0.5f) * inv - 0.5f;
    const float sx = ((float)ox + 0.5f) * inv - 0.5f;
    const float fy = floorf(sy), fx = floorf(sx);
    const int y0 = max((int)fy, 0), y1 = min((int)ceilf(sy), h - 1);
    const int x0 = max((int)fx, 0), x1 = min((int)ceilf(sx), w - 1);
    const float ly = sy - fy, lx = sx - fx;
    const int wpi = w + bi;
    const long long base = (long long)img * (h + bi) * wpi;
    const uint4* p = x + x_coff_v + cv;
    float a[8], b[8], c[8], d[8], o[8];
    unpack8(__ldg(p + (base + (long long)(y0 + bi) * wpi + x0) * x_pitch_v), a);
    unpack8(__ldg(p + (base + (long long)(y0 + bi) * wpi + x1) * x_pitch_v), b);
    unpack8(__ldg(p + (base + (long long)(y1 + bi) * wpi + x0) * x_pitch_v), c);
    unpack8(__ldg(p + (base + (long long)(y1 + bi) * wpi + x1) * x_pitch_v), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float top = a[e] + (b[e] - a[e]) * lx;
      const float bot = c[e] + (d[e] - c[e]) * lx;
      o[e] = top + (bot - top) * ly;
    }
    res = pack8(o);
  }
  const long long ro = ((long long)img * (ho + bo) + oyp) * wpo + oxp;
  y[ro * y_pitch_v + y_coff_v + cv] = res;
}

// Scale 2: one thread per 2x2 OUTPUT block lying between four input pixels (i, j), (i, j+1), (i+1, j),
// (i+1, j+1), i in [-1, h-1], j in [-1, w-1] with clamped indices: 4 loads and 4 unpacks per 4 outputs instead
// of 16, and the blocks tile the padded output raster exactly (the out-of-image outputs of the edge blocks ARE
// the zero border).  Same lerp formula and weights (0.25 / 0.75) as the general kernel, on fp32 pairs (FFMA2).
__device__ __forceinline__ void unpack8_pairs(const uint4& v, float2 (&f)[4]) {
  f[0] = make_float2(bf16_lo(v.x), bf16_hi(v.x));
  f[1] = make_float2(bf16_lo(v.y), bf16_hi(v.y));
  f[2] = make_float2(bf16_lo(v.z), bf16_hi(v.z));
  f[3] = make_float2(bf16_lo(v.w), bf16_hi(v.w));
}
// a + (b - a) * l on fp32 pairs: two packed FFMA2 (sm_100 fma.rn.f32x2) per pair
__device__ __forceinline__ float2 lerp2(float2 a, float2 b, float2 l) {
  const float2 d = __ffma2_rn(a, make_float2(-1.f, -1.f), b);
  return __ffma2_rn(d, l, a);
}

// (the borders are template parameters: as run-time arguments they cost 8 registers and a quarter of the occupancy of
//  this write-bound kernel - 197 -> 232 us on its largest launch)
template <int bi, int bo>
__global__ void __launch_bounds__(256, 4)
upsample2_kernel(const uint4* __restrict__ x, int h, int w, int cvec, int x_pitch_v, int x_coff_v,
                 uint4* __restrict__ y, int y_pitch_v, int y_coff_v) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (w + 1) * cvec) return;
  const int jb = idx / cvec, cv = idx - jb * cvec;
  const int j = jb - 1, i = (int)blockIdx.y - 1, img = blockIdx.z;
  const int y0 = max(i, 0), y1 = min(i + 1, h - 1), x0 = max(j, 0), x1 = min(j + 1, w - 1);
  const int wpi = w + bi;
  const long long base = (long long)img * (h + bi) * wpi;
  const uint4* p = x + x_coff_v + cv;
  float2 a[4], b[4], c[4], d[4];
  unpack8_pairs(__ldg(p + (base + (long long)(y0 + bi) * wpi + x0) * x_pitch_v), a);
  unpack8_pairs(__ldg(p + (base + (long long)(y0 + bi) * wpi + x1) * x_pitch_v), b);
  unpack8_pairs(__ldg(p + (base + (long long)(y1 + bi) * wpi + x0) * x_pitch_v), c);
  unpack8_pairs(__ldg(p + (base + (long long)(y1 + bi) * wpi + x1) * x_pitch_v), d);
  const int wo = 2 * w, ho = 2 * h, wpo = wo + bo;
  // output pixel (oy, ox) = (2i+1+u, 2j+1+v), u, v in {0,1}.  Raster output (bo = 1): position (oy + 1, ox); oy = -1
  // is the shared zero row, ox = wo the shared zero column (both written as zeros); oy = ho and ox = -1 have no
  // slot.  Dense output (bo = 0): only the pixels inside the image exist.
  uint4* q = y + y_coff_v + cv;
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int ox = 2 * j + 1 + v;
    if (ox < 0) continue;                                   // left of the tensor
    const float lxs = v ? 0.75f : 0.25f;
    const float2 lx = make_float2(lxs, lxs);
    float2 top[4], bot[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      top[e] = lerp2(a[e], b[e], lx);
      bot[e] = lerp2(c[e], d[e], lx);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int oy = 2 * i + 1 + u;
      if (oy >= ho) continue;                               // below the tensor
      const bool inside = oy >= 0 && ox < wo;
      if (!inside && bo == 0) continue;                     // a dense tensor has no border entries
      uint4 res = make_uint4(0, 0, 0, 0);
      if (inside) {
        const float lys = u ? 0.75f : 0.25f;
        const float2 ly = make_float2(lys, lys);
        float2 o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = lerp2(top[e], bot[e], ly);
        res = make_uint4(pack_bf16x2(o[0].x, o[0].y), pack_bf16x2(o[1].x, o[1].y), pack_bf16x2(o[2].x, o[2].y),
                         pack_bf16x2(o[3].x, o[3].y));
      }
      q[(((long long)img * (ho + bo) + (oy + bo)) * wpo + ox) * y_pitch_v] = res;
    }
  }
}

// ------------------------------------------------------------------------------- channel means
// block = 256 threads = 32 pixel lanes x 8 vector lanes (64 channels); grid = (c/64, n, splits).
__global__ void channel_mean_kernel(const uint4* __restrict__ x, int h, int w, int x_pitch_v, int x_coff_v,
                                    float* __restrict__ mean, int c, float scale, int y0, int y1, int bi,
                                    float* __restrict__ part, int* __restrict__ counter) {
  pdl_trigger();
  pdl_wait();
  const int vl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int cb = blockIdx.x, img = blockIdx.y;
  const int wp = w + bi;
  const long long base = (long long)img * (h + bi) * wp;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const int npix = (y1 - y0) * w;                      // rows [y0, y1) only (a spatial shard leaves its halo out)
  for (int pix = blockIdx.z * 32 + pl; pix < npix; pix += gridDim.z * 32) {
    const int yr = pix / w, xx = pix - yr * w, yy = y0 + yr;
    float f[8];
    unpack8(x[(base + (long long)(yy + bi) * wp + xx) * x_pitch_v + x_coff_v + cb * 8 + vl], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
  }
  __shared__ float red[32][65];
#pragma unroll
  for (int e = 0; e < 8; ++e) red[pl][vl * 8 + e] = acc[e];
  __syncthreads();
  const int ch = cb * 64 + threadIdx.x;
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) s += red[i][threadIdx.x];
    part[((long long)img * gridDim.z + blockIdx.z) * c + ch] = s;
  }
  // the splits of an image meet in a fixed order in the last block to arrive (see stat_finish): reproducible means
  if (stat_finish(&counter[img * gridDim.x + cb], gridDim.z)) {
    if (threadIdx.x < 64) {
      float tot = 0.f;
      const int nb = (int)gridDim.z;
      const float* src = part + (long long)img * nb * c + ch;
      int z = 0;
      for (; z + 8 <= nb; z += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (long long)(z + u) * c);
#pragma unroll
        for (int u = 0; u < 8; ++u) tot += v[u];
      }
      for (; z < nb; ++z) tot += __ldcg(src + (long long)z * c);
      mean[(long long)img * c + ch] = tot * scale;
    }
  }
}

__global__ void broadcast_kernel(const float* __restrict__ vec, int n, int kh, int kw, int cvec, int c,
                                 uint4* __restrict__ y, int y_pitch_v, int y_coff_v, int bo) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * (kh + bo) * (kw + bo) * cvec;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int cv = (int)(t % cvec);
  const long long ro = t / cvec;
  const int wp = kw + bo, pl = (kh + bo) * wp;
  const int img = (int)(ro / pl);
  const int pr = (int)(ro - (long long)img * pl);
  const int oy = pr / wp, ox = pr % wp;
  uint4 res = make_uint4(0, 0, 0, 0);
  if (oy >= bo && oy < kh + bo && ox < kw) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = vec[(long long)img * c + cv * 8 + e];
    res = pack8(f);
  }
  y[ro * y_pitch_v + y_coff_v + cv] = res;
}

__global__ void raster_to_nhwc_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c, int x_pitch,
                                      int x_coff, float* __restrict__ y, int bi) {
  const long long total = (long long)n * h * w * c;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int ch = (int)(t % c);
  const long long pix = t / c;
  const int xx = (int)(pix % w);
  const int yy = (int)((pix / w) % h);
  const int img = (int)(pix / ((long long)w * h));
  const long long r = ((long long)img * (h + bi) + yy + bi) * (w + bi) + xx;
  y[t] = __bfloat162float(x[r * x_pitch + x_coff + ch]);
}

// ------------------------------------------------------------------------------- basis softmax
// One block per image: the threads tile the [taps][b] matrix with b fastest, so every pass reads and writes whole
// rows (coalesced); a thread owns column `col` of the rows g, g + groups, ..., and the per-column maxima / sums of
// the row groups meet in shared memory (fixed order).  The per-(image, basis) kernel below reads with a stride of
// b floats - one 32-byte sector per value - which was 0.6 ms at Basis_kpn's T = 8, B = 90 (166 MB of basis).
__global__ void softmax_taps_rows_kernel(const float* __restrict__ in, int taps, int b, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float red[];                         // [groups][b]
  const float* src = in + (long long)blockIdx.x * taps * b;
  float* dst = out + (long long)blockIdx.x * taps * b;
  const int groups = blockDim.x / b;
  const int g = threadIdx.x / b, col = threadIdx.x - g * b;
  const bool active = g < groups;
  float mx = -INFINITY;
  if (active) {
    for (int i = g; i < taps; i += groups) mx = fmaxf(mx, src[(long long)i * b + col]);
    red[g * b + col] = mx;
  }
  __syncthreads();
  if (active)
    for (int k = 0; k < groups; ++k) mx = fmaxf(mx, red[k * b + col]);
  __syncthreads();
  float sum = 0.f;
  if (active) {
    for (int i = g; i < taps; i += groups) sum += expf(src[(long long)i * b + col] - mx);
    red[g * b + col] = sum;
  }
  __syncthreads();
  if (active) {
    sum = 0.f;
    for (int k = 0; k < groups; ++k) sum += red[k * b + col];
    const float inv = 1.f / sum;
    for (int i = g; i < taps; i += groups) dst[(long long)i * b + col] = expf(src[(long long)i * b + col] - mx) * inv;
  }
}

// One block per (image, basis b): softmax over `taps` values strided by b (more than 256 bases).
__global__ void softmax_taps_kernel(const float* __restrict__ in, int taps, int b, float* __restrict__ out) {
  const int img = blockIdx.x / b, bi = blockIdx.x % b;
  const float* src = in + (long long)img * taps * b + bi;
  float* dst = out + (long long)img * taps * b + bi;
  __shared__ float red[32];
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < taps; i += blockDim.x) mx = fmaxf(mx, src[(long long)i * b]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < taps; i += blockDim.x) sum += expf(src[(long long)i * b] - mx);
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) sum += red[i];
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < taps; i += blockDim.x) dst[(long long)i * b] = expf(src[(long long)i * b] - mx) * inv;
}

// ------------------------------------------------------------------------------- cost_volume (data_utils.py:97-113)
// One block per image over Bas [taps][t][b]: the mean over the (tap, frame) rows of the variance across the b
// bases, and the mean over the (frame, basis) columns of (max(sum over taps, 0.75) - 0.75)^2; fp64 accumulation,
// fixed summation order.  per_image[img] = -variance + 0.1 * divergent.
__global__ void cost_volume_kernel(const float* __restrict__ bas, int taps, int tb, int b,
                                   double* __restrict__ per_image) {
  const float* src = bas + (long long)blockIdx.x * taps * tb;
  const int rows = taps * (tb / b);
  __shared__ double red[2][32];
  double var = 0.0, dvg = 0.0;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float* row = src + (long long)r * b;
    double s = 0.0, s2 = 0.0;
    for (int i = 0; i < b; ++i) {
      const double v = row[i];
      s += v;
      s2 += v * v;
    }
    const double m = s / b;
    var += s2 / b - m * m;
  }
  for (int c = threadIdx.x; c < tb; c += blockDim.x) {
    double s = 0.0;
    for (int i = 0; i < taps; ++i) s += src[(long long)i * tb + c];
    const double d = fmax(s, 0.75) - 0.75;
    dvg += d * d;
  }
  for (int o = 16; o; o >>= 1) {
    var += __shfl_xor_sync(0xffffffffu, var, o);
    dvg += __shfl_xor_sync(0xffffffffu, dvg, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = var;
    red[1][threadIdx.x >> 5] = dvg;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    var = dvg = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      var += red[0][i];
      dvg += red[1][i];
    }
    per_image[blockIdx.x] = -var / rows + 0.1 * dvg / tb;
  }
}

// mean of the per-image values (the batch-level reduce_mean of the reference), one warp, fixed order
__global__ void cost_volume_mean_kernel(const double* __restrict__ per_image, int n, double* __restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) s += per_image[i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) {
    out[0] = s / n;
    out[1] = s;
  }
}

}  // namespace ie

using namespace ie;

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

// Threads per block for a row of `total` work items: the fewest blocks of <= 256 threads, then the smallest
// warp multiple that still covers the row (848 items -> 4 blocks of 224, not 4 x 256).
static inline int row_block(long long total, int* blocks) {
  const int nb = ie_ceil_div(total, 256);
  const int per = ie_ceil_div(total, nb);
  *blocks = nb;
  return ((per + 31) / 32) * 32;
}

extern "C" int ie_pack_conv_weights(const float* hwio, int kh, int kw, int cin, int cout, int ktot_pad,
                                    void* packed_bf16, void* stream) {
  IE_REQUIRE(hwio && packed_bf16, "pack_conv_weights: null pointer");
  IE_REQUIRE(kh > 0 && kw > 0 && cin > 0 && cout > 0 && ktot_pad >= kh * kw * cin, "pack_conv_weights: bad sizes");
  const long long total = (long long)cout * ktot_pad;
  pack_weights_kernel<<<ie_ceil_div(total, 256), 256, 0, S(stream)>>>(hwio, kh * kw, cin, cout, ktot_pad,
                                                                     static_cast<__nv_bfloat16*>(packed_bf16));
  IE_LAUNCH_CHECK();
  return IE_OK;
}

extern "C" int ie_pack_input_im2col3x3(const float* x, int n, int hs, int ws, int c, int h, int w, void* raster_bf16,
                                       void* stream) {
  IE_REQUIRE(x && raster_bf16, "pack_input: null pointer");
  IE_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && 9 * c <= 1024, "pack_input: need 9*c <= 1024 (c=%d)", c);
  IE_REQUIRE(hs > 0 && ws > 0 && hs <= h && ws <= w, "pack_input: source %dx%d must fit the %dx%d raster", hs, ws, h, w);
  IE_REQUIRE(n <= 65535 && h + 1 <= 65535, "pack_input: grid too large");
  const int kvec = ((9 * c + 63) / 64) * 8;      // row width in 16-byte chunks: 9*c rounded up to 64 channels
  int nb;
  const int threads = row_block((long long)(w + 1) * kvec, &nb);
  dim3 grid(nb, h + 1, n);
  im2col3x3_kernel<<<grid, threads, 0, S(stream)>>>(x, hs, ws, h, w, c, static_cast<uint4*>(raster_bf16), kvec);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

static int check_slice(const char* who, int c, int pitch, int coff) {
  IE_REQUIRE(c > 0 && c % 8 == 0 && pitch % 8 == 0 && coff % 8 == 0 && coff + c <= pitch,
             "%s: bad channel slice (c %d, pitch %d, coff %d)", who, c, pitch, coff);
  return IE_OK;
}

// scratch of the ordered channel statistics: [n][groups][c] partial sums, then [n][ceil(c/64)... ] ticket counters
static long long stat_scratch_bytes(int n, int groups, int c, int counters_per_img) {
  return ((long long)n * groups * c + (long long)n * counters_per_img) * 4;
}
static int maxpool_row_groups(int h, int layout) {
  const int bo = (layout & IE_LAYOUT_Y_DENSE) ? 0 : 1;
  return (h / 2 + bo + kPoolRows - 1) / kPoolRows;
}
static int channel_mean_splits(int n, int h, int w, int c, int y0, int y1) {
  if (y1 <= 0) { y0 = 0; y1 = h; }
  const int npix = (y1 - y0) * w;
  int splits = (8 * sm_count() + (c / 64) * n - 1) / ((c / 64) * n);       // >= 8 blocks per SM in flight
  const int max_splits = (npix + 31) / 32;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  return splits;
}

extern "C" long long ie_maxpool2_stat_scratch_bytes(int n, int h, int w, int c, int layout) {
  (void)w;
  const int cvec = c / 8, cvb = cvec < 32 ? cvec : 32;
  return stat_scratch_bytes(n, maxpool_row_groups(h, layout), c, (cvec + cvb - 1) / (cvb > 0 ? cvb : 1));
}
extern "C" long long ie_channel_mean_scratch_bytes(int n, int h, int w, int c, int y0, int y1) {
  return stat_scratch_bytes(n, channel_mean_splits(n, h, w, c, y0, y1), c, c / 64);
}

extern "C" int ie_maxpool2_nhwc_bf16(const void* x, int n, int h, int w, int c, int x_pitch, int x_coff, void* y,
                                     int y_pitch, int y_coff, float* chan_mean, void* stat_scratch,
                                     long long stat_scratch_bytes_, int stat_y0, int stat_y1, long long stat_count,
                                     int layout, void* stream) {
  IE_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "maxpool2: bad arguments (h %d, w %d)", h, w);
  IE_REQUIRE(layout >= 0 && layout <= 3, "maxpool2: bad layout flags %d", layout);
  const int bi = (layout & IE_LAYOUT_X_DENSE) ? 0 : 1, bo = (layout & IE_LAYOUT_Y_DENSE) ? 0 : 1;
  if (int rc = check_slice("maxpool2(x)", c, x_pitch, x_coff)) return rc;
  if (int rc = check_slice("maxpool2(y)", c, y_pitch, y_coff)) return rc;
  const int cvec = c / 8;
  const int cvb = cvec < 32 ? cvec : 32;
  const int row_groups = (h / 2 + bo + kPoolRows - 1) / kPoolRows;
  IE_REQUIRE(n <= 65535 && row_groups <= 65535, "maxpool2: grid too large");
  dim3 grid(ie_ceil_div(cvec, cvb), row_groups, n);
  float* part = nullptr;
  int* counter = nullptr;
  if (chan_mean) {
    const long long need = stat_scratch_bytes(n, row_groups, c, (int)grid.x);
    IE_REQUIRE(stat_scratch && stat_scratch_bytes_ >= need, "maxpool2: the channel statistics need %lld bytes of scratch "
               "(ie_maxpool2_stat_scratch_bytes), got %lld", need, stat_scratch_bytes_);
    part = static_cast<float*>(stat_scratch);
    counter = reinterpret_cast<int*>(part + (size_t)n * row_groups * c);
    IE_CUDA(cudaMemsetAsync(counter, 0, sizeof(int) * (size_t)n * grid.x, S(stream)));
  }
  const int threads = (256 / cvb) * cvb;
  if (stat_y1 <= 0) { stat_y0 = 0; stat_y1 = h; }
  if (stat_count <= 0) stat_count = (long long)h * w;
  IE_REQUIRE(stat_y0 >= 0 && stat_y1 <= h && stat_y0 % 2 == 0 && stat_y1 % 2 == 0 && stat_y0 < stat_y1,
             "maxpool2: bad statistics row range [%d, %d)", stat_y0, stat_y1);
  IE_CUDA(launch_pdl(maxpool2_kernel, grid, dim3(threads), 0, S(stream), static_cast<const uint4*>(x), h, w, cvec, cvb,
                     x_pitch / 8, x_coff / 8, static_cast<uint4*>(y), y_pitch / 8, y_coff / 8, chan_mean,
                     1.f / (float)stat_count, stat_y0, stat_y1, bi, bo, part, counter));
  return IE_OK;
}

extern "C" int ie_upsample_bilinear_nhwc_bf16(const void* x, int n, int h, int w, int c, int x_pitch, int x_coff,
                                              int scale, void* y, int y_pitch, int y_coff, int layout, void* stream) {
  IE_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && scale >= 1, "upsample: bad arguments");
  IE_REQUIRE(layout >= 0 && layout <= 3, "upsample: bad layout flags %d", layout);
  const int bi = (layout & IE_LAYOUT_X_DENSE) ? 0 : 1, bo = (layout & IE_LAYOUT_Y_DENSE) ? 0 : 1;
  if (int rc = check_slice("upsample(x)", c, x_pitch, x_coff)) return rc;
  if (int rc = check_slice("upsample(y)", c, y_pitch, y_coff)) return rc;
  IE_REQUIRE(n <= 65535 && h * scale + 2 <= 65535, "upsample: grid too large");
  int nb;
  if (scale == 2) {
    const int threads = row_block((long long)(w + 1) * (c / 8), &nb);
    dim3 grid(nb, h + 1, n);
#define IE_UP2(BI_, BO_)                                                                                         \
  IE_CUDA(launch_pdl(upsample2_kernel<BI_, BO_>, grid, dim3(threads), 0, S(stream), static_cast<const uint4*>(x), h, w, \
                     c / 8, x_pitch / 8, x_coff / 8, static_cast<uint4*>(y), y_pitch / 8, y_coff / 8))
    if (bi && bo) IE_UP2(1, 1);
    else if (bi) IE_UP2(1, 0);
    else if (bo) IE_UP2(0, 1);
    else IE_UP2(0, 0);
#undef IE_UP2
    return IE_OK;
  }
  const int threads = row_block((long long)(w * scale + bo) * (c / 8), &nb);
  dim3 grid(nb, h * scale + bo, n);
  IE_CUDA(launch_pdl(upsample_kernel, grid, dim3(threads), 0, S(stream), static_cast<const uint4*>(x), h, w, c / 8,
                     x_pitch / 8, x_coff / 8, scale, static_cast<uint4*>(y), y_pitch / 8, y_coff / 8, bi, bo));
  return IE_OK;
}

extern "C" int ie_channel_mean_nhwc_bf16(const void* x, int n, int h, int w, int c, int x_pitch, int x_coff,
                                         float* mean, void* scratch, long long scratch_bytes, int y0, int y1,
                                         long long count, int layout, void* stream) {
  IE_REQUIRE(x && mean && n > 0 && h > 0 && w > 0, "channel_mean: bad arguments");
  IE_REQUIRE(layout == 0 || layout == IE_LAYOUT_X_DENSE, "channel_mean: bad layout flags %d", layout);
  const int bi = (layout & IE_LAYOUT_X_DENSE) ? 0 : 1;
  IE_REQUIRE(c % 64 == 0, "channel_mean: c must be a multiple of 64 (got %d)", c);
  if (int rc = check_slice("channel_mean(x)", c, x_pitch, x_coff)) return rc;
  const int splits = channel_mean_splits(n, h, w, c, y0, y1);
  if (y1 <= 0) { y0 = 0; y1 = h; }
  if (count <= 0) count = (long long)h * w;
  IE_REQUIRE(y0 >= 0 && y1 <= h && y0 < y1, "channel_mean: bad row range [%d, %d)", y0, y1);
  const long long need = stat_scratch_bytes(n, splits, c, c / 64);
  IE_REQUIRE(scratch && scratch_bytes >= need, "channel_mean: needs %lld bytes of scratch (ie_channel_mean_scratch_bytes), "
             "got %lld", need, scratch_bytes);
  float* part = static_cast<float*>(scratch);
  int* counter = reinterpret_cast<int*>(part + (size_t)n * splits * c);
  IE_CUDA(cudaMemsetAsync(counter, 0, sizeof(int) * (size_t)n * (c / 64), S(stream)));
  IE_REQUIRE(n <= 65535, "channel_mean: n too large");
  dim3 grid(c / 64, n, splits);
  IE_CUDA(launch_pdl(channel_mean_kernel, grid, dim3(256), 0, S(stream), static_cast<const uint4*>(x), h, w, x_pitch / 8,
                     x_coff / 8, mean, c, 1.f / (float)count, y0, y1, bi, part, counter));
  return IE_OK;
}

extern "C" int ie_broadcast_hw_bf16(const float* vec, int n, int kh, int kw, int c, void* y, int y_pitch, int y_coff,
                                    int layout, void* stream) {
  IE_REQUIRE(vec && y && n > 0 && kh > 0 && kw > 0, "broadcast: bad arguments");
  IE_REQUIRE(layout == 0 || layout == IE_LAYOUT_Y_DENSE, "broadcast: bad layout flags %d", layout);
  const int bo = (layout & IE_LAYOUT_Y_DENSE) ? 0 : 1;
  if (int rc = check_slice("broadcast(y)", c, y_pitch, y_coff)) return rc;
  const long long total = (long long)n * (kh + bo) * (kw + bo) * (c / 8);
  IE_CUDA(launch_pdl(broadcast_kernel, dim3(ie_ceil_div(total, 256)), dim3(256), 0, S(stream), vec, n, kh, kw, c / 8, c,
                     static_cast<uint4*>(y), y_pitch / 8, y_coff / 8, bo));
  return IE_OK;
}

extern "C" int ie_raster_to_nhwc_f32(const void* x, int n, int h, int w, int c, int x_pitch, int x_coff, float* y,
                                     int layout, void* stream) {
  IE_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && c > 0 && x_coff + c <= x_pitch, "raster_to_nhwc: bad arguments");
  IE_REQUIRE(layout == 0 || layout == IE_LAYOUT_X_DENSE, "raster_to_nhwc: bad layout flags %d", layout);
  const int bi = (layout & IE_LAYOUT_X_DENSE) ? 0 : 1;
  const long long total = (long long)n * h * w * c;
  raster_to_nhwc_kernel<<<ie_ceil_div(total, 256), 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(x), n, h, w,
                                                                       c, x_pitch, x_coff, y, bi);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

extern "C" int ie_softmax_taps_f32(const float* originbasis, int n, int taps, int b, float* bas, void* stream) {
  IE_REQUIRE(originbasis && bas && n > 0 && taps > 0 && b > 0, "softmax_taps: bad arguments");
  if (b <= 256) {
    const int threads = ((long long)taps * b > 16384) ? 1024 : 256;
    const size_t smem = sizeof(float) * (size_t)(threads / b) * b;
    IE_CUDA(launch_pdl(softmax_taps_rows_kernel, dim3(n), dim3(threads), smem, S(stream), originbasis, taps, b, bas));
  } else {
    softmax_taps_kernel<<<n * b, 256, 0, S(stream)>>>(originbasis, taps, b, bas);
  }
  IE_LAUNCH_CHECK();
  return IE_OK;
}

extern "C" int ie_cost_volume_f32(const float* bas, int n, int taps, int tb, int b, double* per_image, double* out,
                                  void* stream) {
  IE_REQUIRE(bas && per_image && out && n > 0 && taps > 0 && b > 0 && tb > 0 && tb % b == 0,
             "cost_volume: bad arguments");
  cost_volume_kernel<<<n, 256, 0, S(stream)>>>(bas, taps, tb, b, per_image);
  IE_LAUNCH_CHECK();
  cost_volume_mean_kernel<<<1, 32, 0, S(stream)>>>(per_image, n, out);
  IE_LAUNCH_CHECK();
  return IE_OK;
}
