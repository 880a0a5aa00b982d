// Per-pixel quality metrics of /root/reference/data_utils.py:24-164 as eval.py:139-182 uses them:
// white-level inversion, sRGB curve, 8-px crop, squared-error sums (PSNR) and the forward-difference
// gradient L1 term of basic_img_loss - in ONE pass over the tensors instead of the reference's
// 3T+4 separate invert_preproc calls.  Plus an SSIM extension (tf.image.ssim semantics).
// fp32 arithmetic per pixel, fp64 for everything that is accumulated across pixels.
#include "ie_common.cuh"
#include "ie_ptx.cuh"

namespace ie {

// sRGBforward, data_utils.py:24-36.
__device__ __forceinline__ float srgb_forward(float x) {
  const float b = .0031308f, a = .055f, k0 = 12.92f;
  const float gamma = 1.f / 2.4f;
  const float k1 = (1.f + a) * gamma;
  // x^gamma via exp2(gamma*log2(x)): x >= b > 0 so log2 is finite
  const float g = (1.f + a) * exp2f(gamma * log2f(fmaxf(x, b))) - a;
  float r = (x < b) ? k0 * x : g;
  if (x > 1.f) r = k1 * x - k1 + 1.f;
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum NQ per-thread floats over a 256-thread block and add them (fp64) to dst[0..NQ).
template <int NQ>
__device__ __forceinline__ void block_accumulate(const float (&v)[NQ], double* dst, int nq) {
  __shared__ float red[8][NQ];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float s = warp_sum(v[q]);
    if (lane == 0) red[warp][q] = s;
  }
  __syncthreads();
  if (threadIdx.x < nq) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += (double)red[w][threadIdx.x];
    atomicAdd(&dst[threadIdx.x], s);
  }
}

// ------------------------------------------------------------------------------- mean over H,W
__global__ void mean_hw_kernel(const float* __restrict__ x, long long npix, int pitch, int coff, float scale,
                               float* __restrict__ out) {
  const int img = blockIdx.y;
  const float* p = x + (long long)img * npix * pitch + coff;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x)
    acc += p[i * pitch];
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(&out[img], (float)(s * scale));
  }
}

// ------------------------------------------------------------------------------- invert_preproc
__device__ __forceinline__ float srgb_fast(float x);

// Contiguous single-channel images (pitch 1), w and the crop multiples of 4, 16-byte aligned: 4 px per thread with
// 128-bit loads / stores and the MUFU lg2/ex2 curve.  grid = (ceil(wc/4 / 128), hc, n).
__global__ void __launch_bounds__(128)
invert_preproc_vec_kernel(const float* __restrict__ img, const float* __restrict__ wl, int h, int w, int crop,
                          float* __restrict__ out) {
  const int wc = w - 2 * crop, hc = h - 2 * crop;
  const int xq = blockIdx.x * blockDim.x + threadIdx.x;
  if (xq * 4 >= wc) return;
  const int y = blockIdx.y, n = blockIdx.z;
  const float inv_wl = 1.f / wl[n];
  const float4 v = __ldcs(reinterpret_cast<const float4*>(img + ((long long)n * h + y + crop) * w + crop + xq * 4));
  float4 r;
  r.x = srgb_fast(v.x * inv_wl); r.y = srgb_fast(v.y * inv_wl);
  r.z = srgb_fast(v.z * inv_wl); r.w = srgb_fast(v.w * inv_wl);
  __stcs(reinterpret_cast<float4*>(out + ((long long)n * hc + y) * wc + xq * 4), r);
}

__global__ void invert_preproc_kernel(const float* __restrict__ img, int pitch, int coff, int nch,
                                      const float* __restrict__ wl, int h, int w, int crop, float* __restrict__ out,
                                      long long total) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int wc = w - 2 * crop, hc = h - 2 * crop;
  const int x = (int)(t % wc);
  const int y = (int)((t / wc) % hc);
  const int n = (int)(t / ((long long)wc * hc));
  const float* p = img + (((long long)n * h + y + crop) * w + x + crop) * pitch + coff;
  float v = p[0];
  if (nch > 1) {                      // tf.reduce_mean(burst, axis=-1) of psnr_average_f (data_utils.py:162)
    for (int c = 1; c < nch; ++c) v += p[c];
    v /= (float)nch;
  }
  out[t] = srgb_forward(v / wl[n]);
}

// ------------------------------------------------------------------------------- fused eval metrics
// A block owns a 32-px-wide column strip of one image's cropped area over a range of rows and walks it in
// sub-tiles of 32 rows (+1 halo row/column for the forward differences).  For every pixel of the haloed
// sub-tile the error images d_k = sRGB(e_k/wl) - sRGB(gt/wl) go to shared memory, then each interior thread
// accumulates d_k^2 and |d_k(y+1,x)-d_k(y,x)|/2 + |d_k(y,x+1)-d_k(y,x)|/2 in registers; ONE block reduction
// and 2T+4 fp64 atomics per block at the end.  The kernel is bound by instruction issue (T+4 sRGB curves per
// pixel), not by HBM: the curve uses the MUFU lg2/ex2 approximations (relative error < 1e-6).
constexpr int kMT_W = 32, kMT_H = 32, kMaxT = 8;     // sub-tile: 32 x 32 px, 4 rows per thread (+1 halo row / column)

__device__ __forceinline__ float srgb_fast(float x) {
  const float b = .0031308f, a = .055f, k0 = 12.92f;
  const float gamma = 1.f / 2.4f;
  const float k1 = (1.f + a) * gamma;
  float l, e;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(x, b)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(gamma * l));
  float r = (x < b) ? k0 * x : fmaf(1.f + a, e, -a);
  if (x > 1.f) r = fmaf(k1, x, 1.f - k1);
  return r;
}

__global__ void __launch_bounds__(256)
eval_metrics_kernel(const float* __restrict__ recon, const float* __restrict__ burst, int burst_pitch,
                    const float* __restrict__ truth, const float* __restrict__ wl, int h, int w, int T, int crop,
                    int rows_per_block, double* __restrict__ sums) {
  extern __shared__ float s_d[];                       // [T+1][kMT_H+1][kMT_W+1]
  constexpr int SW = kMT_W + 1, SH = kMT_H + 1;
  const int n = blockIdx.z;
  const int hc = h - 2 * crop, wc = w - 2 * crop;
  const int x0 = blockIdx.x * kMT_W;
  const int ys = blockIdx.y * rows_per_block, ye = min(ys + rows_per_block, hc);
  const float inv_wl = 1.f / wl[n];
  const float inv_T = 1.f / (float)T;
  const int nq = (T + 3) + (T + 1);

  // statically indexed accumulators (a run-time T in the index would push the array to local memory):
  // acc[k] squared error of e_k, acc[9] burst0, acc[10] burst mean, acc[11 + k] gradient L1 of e_k
  float acc[2 * kMaxT + 4];
#pragma unroll
  for (int q = 0; q < 2 * kMaxT + 4; ++q) acc[q] = 0.f;

  for (int y0 = ys; y0 < ye; y0 += kMT_H) {
    if (y0 != ys) __syncthreads();
    // 33 x 33 pixel evaluations by 256 threads: 4 full passes + a 65-pixel tail (6 % halo overhead)
    for (int i = threadIdx.x; i < SH * SW; i += 256) {
      const int ly = i / SW, lx = i - ly * SW;
      const int y = y0 + ly, x = x0 + lx;
      const bool inside = (y < hc) && (x < wc) && (y <= ye);         // the row at ye is only the halo of row ye-1
      // halo pixels are owned by the neighbouring tile / sub-tile
      const bool owner = inside && ly < kMT_H && lx < kMT_W && y < ye;
      if (inside) {
        const long long pix = ((long long)n * h + y + crop) * w + x + crop;
        const float g = srgb_fast(truth[pix * 2] * inv_wl);
        const float* rp = recon + pix * (T + 1);
        const float* bp = burst + pix * burst_pitch;
#pragma unroll
        for (int k = 0; k < kMaxT + 1; ++k) {
          if (k <= T) {
            const float d = srgb_fast(rp[k] * inv_wl) - g;
            s_d[(k * SH + ly) * SW + lx] = d;
            if (owner) acc[k] = fmaf(d, d, acc[k]);
          }
        }
        if (owner) {
          float b0 = 0.f, bsum = 0.f;
#pragma unroll
          for (int t = 0; t < kMaxT; ++t) {
            if (t < T) {
              const float v = bp[t];
              if (t == 0) b0 = v;
              bsum += v;
            }
          }
          const float d0 = srgb_fast(b0 * inv_wl) - g;                 // psnr_burst0, data_utils.py:152-154
          const float da = srgb_fast((bsum * inv_T) * inv_wl) - g;     // psnr_average_f, :162-164
          acc[kMaxT + 1] = fmaf(d0, d0, acc[kMaxT + 1]);
          acc[kMaxT + 2] = fmaf(da, da, acc[kMaxT + 2]);
        }
      }
    }
    __syncthreads();
    {
      const int lx = threadIdx.x & 31;
      const int x = x0 + lx;
#pragma unroll
      for (int r = 0; r < kMT_H / 8; ++r) {
        const int ly = (threadIdx.x >> 5) + 8 * r;
        const int y = y0 + ly;
        if (y < ye && y < hc - 1 && x < wc - 1) {
#pragma unroll
          for (int k = 0; k < kMaxT + 1; ++k) {
            if (k <= T) {
              const float c = s_d[(k * SH + ly) * SW + lx];
              acc[kMaxT + 3 + k] += .5f * fabsf(s_d[(k * SH + ly + 1) * SW + lx] - c) + .5f * fabsf(s_d[(k * SH + ly) * SW + lx + 1] - c);
            }
          }
        }
      }
    }
  }
  // block reduction; static slot q -> output index: k -> k, 9 -> T+1, 10 -> T+2, 11+k -> T+3+k
  __shared__ float red[8][2 * kMaxT + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 2 * kMaxT + 4; ++q) {
    const float v = warp_sum(acc[q]);
    if (lane == 0) red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2 * kMaxT + 4) {
    const int q = threadIdx.x;
    int dst = -1;
    if (q <= kMaxT) dst = (q <= T) ? q : -1;
    else if (q == kMaxT + 1) dst = T + 1;
    else if (q == kMaxT + 2) dst = T + 2;
    else dst = (q - kMaxT - 3 <= T) ? T + 3 + (q - kMaxT - 3) : -1;
    if (dst >= 0) {
      double v = 0.0;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) v += (double)red[wq][q];
      atomicAdd(&sums[(long long)n * nq + dst], v);
    }
  }
}

// ------------------------------------------------------------------------------- fused eval metrics, row-streaming
// Same contract as eval_metrics_kernel, organised for the HBM roofline.  The three tensors are interleaved NHWC
// (recon T+1 floats per pixel, burst `burst_pitch`, truth 2), so a per-pixel thread reads them with a 20-byte lane
// stride.  Here a block owns a strip of (warps x 31) columns and walks down its rows in batches of `rb` rows:
//   stage A  every row segment is ONE contiguous run of floats -> lane r of warp 0 moves row r of the next batch with
//            three 1-D TMA bulk copies (cp.async.bulk global -> shared, completion on an mbarrier): no per-thread
//            copy instructions (a 16-byte cp.async version spent more issue slots on index math than stage B on the
//            curves), fully coalesced, and batch i+1 is in flight while batch i is processed - DRAM latency is hidden
//            by the two-buffer pipeline, not by occupancy;
//   stage B  thread = pixel column: reads its pixel from shared memory (lane stride C floats: conflict-free for odd
//            C), evaluates the T+4 sRGB curves, squares the errors, takes the horizontal forward difference from the
//            right neighbour with a warp shuffle (warps overlap by one column, lane 31 is halo only) and the vertical
//            one from the previous row's errors kept in registers.
// Row segments start at arbitrary float offsets: the copy starts at the enclosing 16-byte boundary and the pixel data
// sit `mis` floats into the shared-memory row.  Requires 16-byte aligned base pointers (else the tile kernel above).
__device__ __forceinline__ void bulk_g2s(float* smem_dst, const float* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One row segment of `len` floats starting at element `s` of `base` (which holds `total` floats) is copied from the
// enclosing 16-byte boundary.  plan(): the aligned start and the float count of the bulk copy; a tail that would cross
// the end of the tensor (< 16 bytes, last row of the last image only) is moved with plain loads here, BEFORE the
// lane's mbarrier arrive releases them.
struct RowCopy {
  long long a0;
  int nfl;
  __device__ __forceinline__ void plan(float* sdst, const float* __restrict__ base, long long total, long long s, int len) {
    a0 = s & ~3ll;
    nfl = (((int)(s - a0) + len + 3) >> 2) << 2;
    if (a0 + nfl > total) {
      const int keep = (int)((total - a0) & ~3ll);
      for (int i = keep; a0 + i < total; ++i) sdst[i] = base[a0 + i];
      nfl = keep;
    }
    if (nfl < 0) nfl = 0;
  }
  __device__ __forceinline__ uint32_t bytes() const { return (uint32_t)nfl * 4u; }
  __device__ __forceinline__ void go(float* sdst, const float* __restrict__ base, uint64_t* bar) const {
    if (nfl > 0) bulk_g2s(sdst, base + a0, bytes(), bar);
  }
};

// sRGBforward(x * inv_wl) - data_utils.py:24-36 on the white-level-normalised value - with the scale folded in:
// power segment on max(x, b*wl) as ex2(gamma*lg2(x) + gamma*lg2(inv_wl)); the linear toe is min(k0*x', power) (the
// power curve is concave and meets the toe at b, so the toe is the smaller one exactly where x' < b); x' > 1 selects
// the tangent.  MUFU lg2/ex2: relative error < 1e-6.
struct SrgbScaled {
  float b_wl, gl, k0i, k1i, c1;
  __device__ __forceinline__ explicit SrgbScaled(float inv_wl) {
    const float gamma = 1.f / 2.4f, k1 = 1.055f * gamma;
    b_wl = .0031308f / inv_wl;
    gl = gamma * log2f(inv_wl);
    k0i = 12.92f * inv_wl;
    k1i = k1 * inv_wl;
    c1 = 1.f - k1;
  }
  __device__ __forceinline__ float operator()(float x, float wl_) const {
    float l, e;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(x, b_wl)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(1.f / 2.4f, l, gl)));
    float r = fminf(k0i * x, fmaf(1.055f, e, -.055f));
    if (x > wl_) r = fmaf(k1i, x, c1);
    return r;
  }
};

template <int TT>
__global__ void __launch_bounds__(128)
eval_metrics_rows_kernel(const float* __restrict__ recon, const float* __restrict__ burst, int burst_pitch,
                         const float* __restrict__ truth, const float* __restrict__ wl, int n_img, int h, int w,
                         int T_rt, int crop, int rows_per_block, int rb, double* __restrict__ sums,
                         float* __restrict__ crop_deblur, float* __restrict__ crop_gt) {
  extern __shared__ float4 s_dyn4[];
  float* sm = reinterpret_cast<float*>(s_dyn4);
  __shared__ uint64_t full[2];
  const int T = TT ? TT : T_rt;
  const int C = T + 1;
  const int nwarps = blockDim.x >> 5;
  const int two = nwarps * 31;                                  // output columns of the block
  const int n = blockIdx.z;
  const int hc = h - 2 * crop, wc = w - 2 * crop;
  const int x0 = blockIdx.x * two;
  const int ys = blockIdx.y * rows_per_block, ye = min(ys + rows_per_block, hc);
  const int ylast = min(ye, hc - 1);                            // row ye is only the vertical halo of row ye - 1
  const int nrows_total = ylast - ys + 1;
  const float wl_n = wl[n];
  const float inv_wl = 1.f / wl_n;
  const float inv_T = 1.f / (float)T;
  const SrgbScaled curve(inv_wl);

  // shared-memory rows: strides in floats (multiples of 4); +6 = up to 3 floats of misalignment at either end
  const int st_rc = (((two + 1) * C + 6) >> 2) << 2, st_tr = (((two + 1) * 2 + 6) >> 2) << 2,
            st_bu = ((two * burst_pitch + 6) >> 2) << 2;
  const int buf_floats = rb * (st_rc + st_tr + st_bu);
  const int px_h = min(two + 1, w - (x0 + crop));               // pixels of the row that exist (incl. right halo)
  const int px_b = min(two, w - (x0 + crop));
  const long long npix_all = (long long)n_img * h * w;

  if (threadIdx.x == 0) {
    mbar_init(&full[0], rb);
    mbar_init(&full[1], rb);
    fence_barrier_init();
  }
  __syncthreads();

  // warp 0, lane r: row r of the batch (every lane < rb arrives once per batch, rows that do not exist with 0 bytes)
  auto issue = [&](int batch) {
    const int r = threadIdx.x;
    if (r < rb) {
      const int yb = ys + batch * rb;
      float* b = sm + (batch & 1) * buf_floats;
      uint64_t* bar = &full[batch & 1];
      const int y = yb + r;
      if (y <= ylast) {
        const long long pix0 = ((long long)n * h + y + crop) * w + x0 + crop;
        const bool need_burst = y < ye;                          // the halo row needs no burst
        float* d_rc = b + r * st_rc;
        float* d_tr = b + rb * st_rc + r * st_tr;
        float* d_bu = b + rb * (st_rc + st_tr) + r * st_bu;
        RowCopy c_rc, c_tr, c_bu;
        c_rc.plan(d_rc, recon, npix_all * C, pix0 * C, px_h * C);
        c_tr.plan(d_tr, truth, npix_all * 2, pix0 * 2, px_h * 2);
        c_bu.nfl = 0;
        if (need_burst) c_bu.plan(d_bu, burst, npix_all * burst_pitch, pix0 * burst_pitch, px_b * burst_pitch);
        mbar_arrive_expect_tx(bar, c_rc.bytes() + c_tr.bytes() + c_bu.bytes());
        c_rc.go(d_rc, recon, bar);
        c_tr.go(d_tr, truth, bar);
        c_bu.go(d_bu, burst, bar);
      } else {
        mbar_arrive(bar);
      }
    }
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = warp * 31 + lane;
  const int x = x0 + lx;
  const bool col_own = (lane < 31) && (x < wc);
  const bool col_grad = (lane < 31) && (x < wc - 1);

  float sq[kMaxT + 1], gr[kMaxT + 1], prev[kMaxT + 1];
  float sq_b0 = 0.f, sq_avg = 0.f;
#pragma unroll
  for (int k = 0; k <= kMaxT; ++k) { sq[k] = 0.f; gr[k] = 0.f; prev[k] = 0.f; }

  const int nbatch = (nrows_total + rb - 1) / rb;
  issue(0);
  for (int bi = 0; bi < nbatch; ++bi) {
    if (bi + 1 < nbatch) issue(bi + 1);
    mbar_wait(&full[bi & 1], (bi >> 1) & 1);
    const float* b = sm + (bi & 1) * buf_floats;
    const int yb = ys + bi * rb;
    const int nr = min(rb, ylast - yb + 1);
    for (int r = 0; r < nr; ++r) {
      const int y = yb + r;
      const long long pix0 = ((long long)n * h + y + crop) * w + x0 + crop;
      const int mis_rc = (int)((pix0 * C) & 3), mis_tr = (int)((pix0 * 2) & 3), mis_bu = (int)((pix0 * burst_pitch) & 3);
      const float* rc = b + r * st_rc + mis_rc + lx * C;
      const float* tr = b + rb * st_rc + r * st_tr + mis_tr + lx * 2;
      const float* bu = b + rb * (st_rc + st_tr) + r * st_bu + mis_bu + lx * burst_pitch;
      const bool row_own = y < ye;                              // uniform
      const bool own = col_own && row_own;
      const bool hval = col_grad && row_own && (y < hc - 1);
      const bool vval = col_grad && (y > ys);
      const float g = curve(tr[0], wl_n);
      if (crop_gt != nullptr && own) {
        // by-product for the SSIM extension: invert_preproc(truth) and invert_preproc(deblurred) of this pixel
        // (eval.py:146-149) - the values this kernel forms anyway - as dense [n][hc][wc] crops
        const long long o = ((long long)n * hc + y) * wc + x;
        crop_gt[o] = g;
        crop_deblur[o] = curve(rc[0], wl_n);
      }
#pragma unroll
      for (int k = 0; k <= kMaxT; ++k) {
        if (k <= T) {
          const float c = curve(rc[k], wl_n) - g;
          const float cr = __shfl_down_sync(0xffffffffu, c, 1);
          if (own) sq[k] = fmaf(c, c, sq[k]);
          if (hval) gr[k] += fabsf(cr - c);
          if (vval) gr[k] += fabsf(c - prev[k]);
          prev[k] = c;
        }
      }
      if (own) {
        float b0 = 0.f, bsum = 0.f;
#pragma unroll
        for (int t = 0; t < kMaxT; ++t) {
          if (t < T) {
            const float v = bu[t];
            if (t == 0) b0 = v;
            bsum += v;
          }
        }
        const float d0 = curve(b0, wl_n) - g;                        // psnr_burst0, data_utils.py:152-154
        const float da = curve(bsum * inv_T, wl_n) - g;              // psnr_average_f, :162-164
        sq_b0 = fmaf(d0, d0, sq_b0);
        sq_avg = fmaf(da, da, sq_avg);
      }
    }
    __syncthreads();                                            // the buffer is refilled by the next iteration's issue
  }

  // block reduction: slot q -> output index: k -> k, 9 -> T+1, 10 -> T+2, 11+k -> T+3+k (gradient sums carry 1/2)
  __shared__ float red[4][2 * kMaxT + 4];
#pragma unroll
  for (int q = 0; q < 2 * kMaxT + 4; ++q) {
    float v;
    if (q <= kMaxT) v = sq[q];
    else if (q == kMaxT + 1) v = sq_b0;
    else if (q == kMaxT + 2) v = sq_avg;
    else v = gr[q - kMaxT - 3];
    v = warp_sum(v);
    if (lane == 0) red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2 * kMaxT + 4) {
    const int q = threadIdx.x;
    const int nq = 2 * T + 4;
    int dst = -1;
    double scale = 1.0;
    if (q <= kMaxT) dst = (q <= T) ? q : -1;
    else if (q == kMaxT + 1) dst = T + 1;
    else if (q == kMaxT + 2) dst = T + 2;
    else { dst = (q - kMaxT - 3 <= T) ? T + 3 + (q - kMaxT - 3) : -1; scale = 0.5; }
    if (dst >= 0) {
      double v = 0.0;
      for (int wq = 0; wq < nwarps; ++wq) v += (double)red[wq][q];
      atomicAdd(&sums[(long long)n * nq + dst], v * scale);
    }
  }
}

// ------------------------------------------------------------------------------- per-image sums -> totals
// One block: psnr_k[n] = -10 log10(sums[n][k] / npx) (data_utils.py:118-119), loss_k[n] = mse + grad-L1
// (:46-51); totals = [sum_n psnr_0..T+2, sum_n loss_0, sum_n sum_{k>=1} loss_k, n]  (fp64).
__global__ void metric_totals_kernel(const double* __restrict__ sums, int n, int T, double npx, double ngr,
                                     const double* __restrict__ ssim_sums, double ssim_px, double* __restrict__ totals) {
  const int nq = 2 * T + 4;
  const int nval = T + 5 + (ssim_sums != nullptr ? 1 : 0);      // values before the trailing image count
  __shared__ double red[32][2 * kMaxT + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = 0; k < nval; ++k) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double* s = sums + (long long)i * nq;
      if (k < T + 3) {
        acc += -10.0 * log10(s[k] / npx);
      } else if (k == T + 3) {
        acc += s[0] / npx + s[T + 3] / ngr;
      } else if (k == T + 4) {
        for (int f = 1; f <= T; ++f) acc += s[f] / npx + s[T + 3 + f] / ngr;
      } else {
        acc += ssim_sums[i] / ssim_px;                           // mean SSIM of image i (extension)
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp][k] = acc;
  }
  __syncthreads();
  if (threadIdx.x < nval) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
    totals[threadIdx.x] = s;
  }
  if (threadIdx.x == 0) totals[nval] = (double)n;
}

// ------------------------------------------------------------------------------- pair reductions
__global__ void sqdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, long long count,
                                  double* __restrict__ sums) {
  const int n = blockIdx.y;
  const float* pa = a + (long long)n * count;
  const float* pb = b + (long long)n * count;
  float acc[1] = {0.f};
  const bool vec = ((reinterpret_cast<uintptr_t>(pa) | reinterpret_cast<uintptr_t>(pb)) & 15) == 0;
  const long long nvec = vec ? count / 4 : 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float4* va = reinterpret_cast<const float4*>(pa);
  const float4* vb = reinterpret_cast<const float4*>(pb);
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // 4 independent 16-byte loads per operand in flight
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    float4 x[4], y[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { x[u] = __ldcs(va + i + u * stride); y[u] = __ldcs(vb + i + u * stride); }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float d0 = x[u].x - y[u].x, d1 = x[u].y - y[u].y, d2 = x[u].z - y[u].z, d3 = x[u].w - y[u].w;
      acc[0] += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
  }
  for (; i < nvec; i += stride) {
    const float4 x = __ldcs(va + i), y = __ldcs(vb + i);
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    acc[0] += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (long long j = nvec * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; j < count; j += stride) {
    const float d = pa[j] - pb[j];
    acc[0] += d * d;
  }
  block_accumulate<1>(acc, sums + n, 1);
}

__global__ void img_loss_sums_kernel(const float* __restrict__ a, const float* __restrict__ b, int h, int w,
                                     long long total, double* __restrict__ sums) {
  float acc[2] = {0.f, 0.f};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % w);
    const int y = (int)((t / w) % h);
    const float d = a[t] - b[t];
    acc[0] += d * d;
    if (y < h - 1 && x < w - 1) {
      const float dy = a[t + w] - b[t + w], dx = a[t + 1] - b[t + 1];
      acc[1] += .5f * fabsf(dy - d) + .5f * fabsf(dx - d);
    }
  }
  block_accumulate<2>(acc, sums, 2);
}

// Vectorised variant (w % 4 == 0, 16-byte aligned rows): a thread owns 4 consecutive x and walks kLossRows
// rows downwards, carrying the difference row it just loaded as the "current" row of the next step, so every
// element is loaded once (+1 halo row per strip, +1 scalar per row for the x+1 neighbour of the last lane).
constexpr int kLossRows = 32;
__global__ void __launch_bounds__(128)
img_loss_sums_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, int h, int w,
                         double* __restrict__ sums) {
  const int xq = blockIdx.x * blockDim.x + threadIdx.x;      // group of 4 pixels
  const int x = xq * 4;
  const int y0 = blockIdx.y * kLossRows;
  const long long img = (long long)blockIdx.z * h * w;
  float acc[2] = {0.f, 0.f};
  if (x < w) {
    const float* pa = a + img + x;
    const float* pb = b + img + x;
    const bool has_right = x + 4 < w;
    auto load = [&](int y, float (&d)[5]) {
      const float4 va = __ldcs(reinterpret_cast<const float4*>(pa + (long long)y * w));
      const float4 vb = __ldcs(reinterpret_cast<const float4*>(pb + (long long)y * w));
      d[0] = va.x - vb.x; d[1] = va.y - vb.y; d[2] = va.z - vb.z; d[3] = va.w - vb.w;
      d[4] = has_right ? __ldg(pa + (long long)y * w + 4) - __ldg(pb + (long long)y * w + 4) : 0.f;
    };
    // rows y+1 and y+2 are in flight while row y is consumed (two loads of latency per thread hidden)
    float cur[5], nxt[5], nn[5];
    const int yend = min(y0 + kLossRows, h);
    load(y0, cur);
    if (y0 + 1 < h) load(y0 + 1, nxt);
    for (int y = y0; y < yend; ++y) {
      const bool has_below = y + 1 < h;
      if (y + 2 < h && y + 2 <= yend) load(y + 2, nn);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[0] = fmaf(cur[e], cur[e], acc[0]);
        if (has_below && x + e < w - 1) acc[1] += .5f * fabsf(nxt[e] - cur[e]) + .5f * fabsf(cur[e + 1] - cur[e]);
      }
#pragma unroll
      for (int e = 0; e < 5; ++e) { cur[e] = nxt[e]; nxt[e] = nn[e]; }
    }
  }
  // 128-thread block: reduce with the 256-thread helper's layout (upper warps contribute zeros)
  __shared__ float red[4][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float s = warp_sum(acc[q]);
    if (lane == 0) red[warp][q] = s;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const double s = (double)red[0][threadIdx.x] + (double)red[1][threadIdx.x] + (double)red[2][threadIdx.x] +
                     (double)red[3][threadIdx.x];
    atomicAdd(&sums[threadIdx.x], s);
  }
}

// ------------------------------------------------------------------------------- SSIM (extension)
// tf.image.ssim semantics: 11-tap Gaussian (sigma 1.5) window, VALID, K1 = .01, K2 = .03, max_val 1.
constexpr int kSR = 5, kSTaps = 11;

// Streaming SSIM: a thread owns one output COLUMN of a 128-column strip and walks down the rows.  Per input row it
// filters its 11 horizontal neighbours (coalesced loads through L1: lane i reads x+i .. x+i+10) into the five moments
// (a, b, a^2, b^2, ab) as fp32 pairs (FFMA2), pushes them into an 11-row window held in REGISTERS, and once the window
// is full produces one output row with the vertical pass.  No shared memory, no barriers, no horizontal halo work;
// vertical halo = 10 rows per kSsimRows outputs.
constexpr int kSsimRows = 128;
__global__ void __launch_bounds__(128)
ssim_stream_kernel(const float* __restrict__ a, const float* __restrict__ b, int h, int w, double* __restrict__ sums) {
  const int n = blockIdx.z;
  const int ho = h - 2 * kSR, wo = w - 2 * kSR;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;            // output column
  const int y0 = blockIdx.y * kSsimRows;                          // first output row of this block
  const int y1 = min(y0 + kSsimRows, ho);
  float g[kSTaps];
  {
    float gs = 0.f;
#pragma unroll
    for (int k = 0; k < kSTaps; ++k) {
      const float c = (float)(k - kSR);
      g[k] = expf(-0.5f * c * c / (1.5f * 1.5f));
      gs += g[k];
    }
#pragma unroll
    for (int k = 0; k < kSTaps; ++k) g[k] /= gs;
  }
  float acc = 0.f;
  if (x < wo) {
    const float* pa = a + (long long)n * h * w + x;
    const float* pb = b + (long long)n * h * w + x;
    // vertical window: hm[r] = horizontally filtered moments of input row (current - r)
    float2 w_ab[kSTaps], w_sq[kSTaps];
    float w_x[kSTaps];
    const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
    // input rows y0 .. y1 + 9; output row (yi - 10) is complete after input row yi
    for (int yi0 = y0; yi0 < y1 + 2 * kSR; yi0 += kSTaps) {
#pragma unroll
      for (int slot = 0; slot < kSTaps; ++slot) {                  // the window rotates by one slot per input row
        const int yi = yi0 + slot;
        if (yi < y1 + 2 * kSR) {
          const float* ra = pa + (long long)yi * w;
          const float* rb = pb + (long long)yi * w;
          float2 m_ab = make_float2(0.f, 0.f), m_sq = make_float2(0.f, 0.f);
          float m_x = 0.f;
#pragma unroll
          for (int k = 0; k < kSTaps; ++k) {
            const float va = __ldg(ra + k), vb = __ldg(rb + k);
            const float2 gg = make_float2(g[k], g[k]);
            m_ab = __ffma2_rn(gg, make_float2(va, vb), m_ab);
            m_sq = __ffma2_rn(gg, make_float2(va * va, vb * vb), m_sq);
            m_x = fmaf(g[k], va * vb, m_x);
          }
          w_ab[slot] = m_ab; w_sq[slot] = m_sq; w_x[slot] = m_x;
          const int yo = yi - 2 * kSR;
          if (yo >= y0) {
            // vertical pass: input row yi - r sits in slot (slot - r) mod 11 and takes tap 10 - r
            float2 v_ab = make_float2(0.f, 0.f), v_sq = make_float2(0.f, 0.f);
            float v_x = 0.f;
#pragma unroll
            for (int r = 0; r < kSTaps; ++r) {
              const int sl = (slot - r + kSTaps) % kSTaps;           // compile-time
              const float2 gg = make_float2(g[kSTaps - 1 - r], g[kSTaps - 1 - r]);
              v_ab = __ffma2_rn(gg, w_ab[sl], v_ab);
              v_sq = __ffma2_rn(gg, w_sq[sl], v_sq);
              v_x = fmaf(g[kSTaps - 1 - r], w_x[sl], v_x);
            }
            const float mu_a = v_ab.x, mu_b = v_ab.y;
            const float num0 = 2.f * mu_a * mu_b, den0 = mu_a * mu_a + mu_b * mu_b;
            const float lum = (num0 + c1) / (den0 + c1);
            const float cs = (2.f * v_x - num0 + c2) / (v_sq.x + v_sq.y - den0 + c2);
            acc += lum * cs;
          }
        }
      }
    }
  }
  // 128-thread block reduction
  __shared__ float red[4];
  const float s = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(&sums[n], (double)red[0] + (double)red[1] + (double)red[2] + (double)red[3]);
}

// Two output columns per thread, four moments.  SSIM only needs  mu_a, mu_b, E[a^2 + b^2]  and  E[ab]
// (sigma_a^2 + sigma_b^2 = E[a^2+b^2] - mu_a^2 - mu_b^2), so the five filtered images of the textbook form become
// four, kept as two fp32 pairs (a, b) and (a^2+b^2, ab) that move through FFMA2.  A thread owns output columns
// (2c, 2c+1): the 12 input samples of a row and their products are formed once and feed both columns (11 of the 12
// each).  Per pixel: 12 loads, 18 product ops, 22 + 22 packed FMAs, ~14 for the SSIM quotient (one MUFU.RCP) - about
// 95 issue slots against ~170 of the one-column kernel above.  The kernel is bound by the FP32 pipe: 88 FMA-lane
// operations per pixel for the two separable passes alone put the ceiling at ~37 % of the HBM roofline.
// Rows staged by TMA.  With plain global loads (first version of this kernel, 17 % of the HBM roofline) every input
// row is fresh data: 24 scalar loads per thread and row, each stalling the warp on the long scoreboard (4.3 stalled
// warps per issued instruction at 16 warps per SM).  Here the block's row segments (256 + 10 columns of both images) arrive in shared memory by 1-D TMA bulk copies, 11
// rows (one turn of the register window) per batch, the next batch in flight while this one is filtered - the
// pipeline of eval_metrics_rows_kernel.  The filter reads shared memory only.
__global__ void __launch_bounds__(128)
ssim_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, int n_img, int h, int w, int rows_per_block,
                 double* __restrict__ sums) {
  extern __shared__ float4 ssim_dyn4[];
  float* sm = reinterpret_cast<float*>(ssim_dyn4);
  __shared__ uint64_t full[2];
  constexpr int kCols = 2 * 128 + 2 * kSR;                        // input columns of a block
  constexpr int kStride = ((kCols + 6) >> 2) << 2;                // floats per staged row (+ misalignment slack)
  constexpr int kBuf = 2 * kSTaps * kStride;                      // one batch: 11 rows of a, 11 rows of b
  const int n = blockIdx.z;
  const int ho = h - 2 * kSR, wo = w - 2 * kSR;
  const int x0 = 2 * blockIdx.x * 128;
  const int x = x0 + 2 * threadIdx.x;                             // first output column of the pair
  const int y0 = blockIdx.y * rows_per_block;
  const int y1 = min(y0 + rows_per_block, ho);
  const int yin_end = y1 + 2 * kSR;                               // input rows y0 .. yin_end - 1
  const int ncols = min(kCols, w - x0);
  const long long total = (long long)n_img * h * w;
  float g[kSTaps];
  {
    float gs = 0.f;
#pragma unroll
    for (int k = 0; k < kSTaps; ++k) {
      const float c = (float)(k - kSR);
      g[k] = expf(-0.5f * c * c / (1.5f * 1.5f));
      gs += g[k];
    }
#pragma unroll
    for (int k = 0; k < kSTaps; ++k) g[k] /= gs;
  }
  if (threadIdx.x == 0) {
    mbar_init(&full[0], kSTaps);
    mbar_init(&full[1], kSTaps);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int batch) {                                   // lane r of warp 0: row r of the batch
    const int r = threadIdx.x;
    if (r < kSTaps) {
      uint64_t* bar = &full[batch & 1];
      const int yi = y0 + batch * kSTaps + r;
      if (yi < yin_end) {
        float* da = sm + (batch & 1) * kBuf + r * kStride;
        float* db = da + kSTaps * kStride;
        const long long e = ((long long)n * h + yi) * w + x0;
        RowCopy ca, cb;
        ca.plan(da, a, total, e, ncols);
        cb.plan(db, b, total, e, ncols);
        mbar_arrive_expect_tx(bar, ca.bytes() + cb.bytes());
        ca.go(da, a, bar);
        cb.go(db, b, bar);
      } else {
        mbar_arrive(bar);
      }
    }
  };
  float acc = 0.f;
  const bool col_ok = x < wo;
  float2 w_ab[2][kSTaps], w_sp[2][kSTaps];                         // [column][window slot]
  const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
  const int nbatch = (yin_end - y0 + kSTaps - 1) / kSTaps;
  issue(0);
  for (int bi = 0; bi < nbatch; ++bi) {
    if (bi + 1 < nbatch) issue(bi + 1);
    mbar_wait(&full[bi & 1], (bi >> 1) & 1);
    const float* buf = sm + (bi & 1) * kBuf;
    const int yi0 = y0 + bi * kSTaps;
    if (col_ok) {
#pragma unroll
      for (int slot = 0; slot < kSTaps; ++slot) {
        const int yi = yi0 + slot;
        if (yi < yin_end) {
          const int mis = (int)((((long long)n * h + yi) * w + x0) & 3);
          const float* ra = buf + slot * kStride + mis + 2 * threadIdx.x;
          const float* rb = ra + kSTaps * kStride;
          float2 h_ab[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
          float2 h_sp[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
          for (int i = 0; i < kSTaps + 1; ++i) {
            const float va = ra[i], vb = rb[i];
            const float2 ab = make_float2(va, vb);
            const float2 sp = make_float2(fmaf(va, va, vb * vb), va * vb);
            if (i < kSTaps) {
              const float2 gg = make_float2(g[i], g[i]);
              h_ab[0] = __ffma2_rn(gg, ab, h_ab[0]);
              h_sp[0] = __ffma2_rn(gg, sp, h_sp[0]);
            }
            if (i >= 1) {
              const float2 gg = make_float2(g[i - 1], g[i - 1]);
              h_ab[1] = __ffma2_rn(gg, ab, h_ab[1]);
              h_sp[1] = __ffma2_rn(gg, sp, h_sp[1]);
            }
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) { w_ab[j][slot] = h_ab[j]; w_sp[j][slot] = h_sp[j]; }
          if (yi - 2 * kSR >= y0) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              float2 v_ab = make_float2(0.f, 0.f), v_sp = make_float2(0.f, 0.f);
#pragma unroll
              for (int r = 0; r < kSTaps; ++r) {
                const int sl = (slot - r + kSTaps) % kSTaps;         // compile-time
                const float2 gg = make_float2(g[kSTaps - 1 - r], g[kSTaps - 1 - r]);
                v_ab = __ffma2_rn(gg, w_ab[j][sl], v_ab);
                v_sp = __ffma2_rn(gg, w_sp[j][sl], v_sp);
              }
              const float num0 = 2.f * v_ab.x * v_ab.y, den0 = fmaf(v_ab.x, v_ab.x, v_ab.y * v_ab.y);
              const float num = (num0 + c1) * (2.f * v_sp.y - num0 + c2);
              const float den = (den0 + c1) * (v_sp.x - den0 + c2);
              acc += __fdividef(num, den);
            }
          }
        }
      }
    }
    __syncthreads();                                              // the buffer is refilled by the next iteration's issue
  }
  __shared__ float red[4];
  const float s = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(&sums[n], (double)red[0] + (double)red[1] + (double)red[2] + (double)red[3]);
}

}  // namespace ie

using namespace ie;
static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" int ie_mean_hw_f32(const float* x, int n, int h, int w, int pitch, int coff, float* out, void* stream) {
  IE_REQUIRE(x && out && n > 0 && n <= 65535 && h > 0 && w > 0 && coff >= 0 && coff < pitch, "mean_hw: bad arguments");
  IE_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * n, S(stream)));
  const long long npix = (long long)h * w;
  int gx = (int)((npix + 256 * 8 - 1) / (256 * 8));
  const int cap = (4 * sm_count() + n - 1) / n;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  mean_hw_kernel<<<dim3(gx, n), 256, 0, S(stream)>>>(x, npix, pitch, coff, 1.f / (float)npix, out);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

extern "C" int ie_invert_preproc_f32(const float* img, int pitch, int coff, int nch, const float* wl, int n, int h,
                                     int w, int crop, float* out, void* stream) {
  IE_REQUIRE(img && wl && out && n > 0 && crop >= 0 && h > 2 * crop && w > 2 * crop && coff >= 0 && nch >= 1 &&
                 coff + nch <= pitch,
             "invert_preproc: bad arguments");
  const long long total = (long long)n * (h - 2 * crop) * (w - 2 * crop);
  const bool vec = pitch == 1 && nch == 1 && w % 4 == 0 && crop % 4 == 0 && n <= 65535 && h - 2 * crop <= 65535 &&
                   ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) {
    dim3 grid(ie_ceil_div((w - 2 * crop) / 4, 128), h - 2 * crop, n);
    invert_preproc_vec_kernel<<<grid, 128, 0, S(stream)>>>(img, wl, h, w, crop, out);
    IE_LAUNCH_CHECK();
    return IE_OK;
  }
  invert_preproc_kernel<<<ie_ceil_div(total, 256), 256, 0, S(stream)>>>(img, pitch, coff, nch, wl, h, w, crop, out, total);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

// Tuning knobs of the row-streaming kernel (tools/metric_sweep.py sweeps them; 0 = default).
static int g_em_rb = 0, g_em_warps = 0, g_em_legacy = 0;
extern "C" int ie_eval_metrics_tune(int rows_per_batch, int warps, int legacy) {
  g_em_rb = rows_per_batch; g_em_warps = warps; g_em_legacy = legacy;
  return IE_OK;
}

template <int TT>
static int launch_eval_metrics_rows(const float* recon, const float* burst, int burst_pitch, const float* truth,
                                    const float* wl, int n, int h, int w, int T, int crop, double* sums,
                                    float* crop_deblur, float* crop_gt, void* stream) {
  const int hc = h - 2 * crop, wc = w - 2 * crop;
  int nwarps = g_em_warps;
  if (nwarps < 1 || nwarps > 4) {                              // narrowest block that wastes the fewest thread columns
    long long best = -1;
    for (int cand = 4; cand >= 2; --cand) {
      const long long cols = (long long)((wc + cand * 31 - 1) / (cand * 31)) * cand * 32;
      if (best < 0 || cols < best) { best = cols; nwarps = cand; }
    }
  }
  const int two = nwarps * 31;
  const int tiles_x = (wc + two - 1) / two;
  const int rb = g_em_rb > 0 ? g_em_rb : 2;
  long long ysplit = 32ll * sm_count() / ((long long)tiles_x * n);
  if (ysplit < 1) ysplit = 1;
  int rows_per_block = (int)((hc + ysplit - 1) / ysplit);
  if (rows_per_block < 16) rows_per_block = 16;
  if (rows_per_block > hc) rows_per_block = hc;
  const int gy = (hc + rows_per_block - 1) / rows_per_block;
  IE_REQUIRE(n <= 65535 && gy <= 65535, "eval_metrics: grid too large");
  const int C = T + 1;
  const int st = 4 * ((((two + 1) * C + 6) >> 2) + (((two + 1) * 2 + 6) >> 2) + ((two * burst_pitch + 6) >> 2));
  const size_t smem = sizeof(float) * 2 * (size_t)rb * st;
  IE_REQUIRE(smem <= 200 * 1024, "eval_metrics: burst_pitch %d needs %zu bytes of shared memory", burst_pitch, smem);
  IE_CUDA(cudaFuncSetAttribute(eval_metrics_rows_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  eval_metrics_rows_kernel<TT><<<dim3(tiles_x, gy, n), nwarps * 32, smem, S(stream)>>>(
      recon, burst, burst_pitch, truth, wl, n, h, w, T, crop, rows_per_block, rb, sums, crop_deblur, crop_gt);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

static int eval_metrics_impl(const float* recon, const float* burst, int burst_pitch, const float* truth,
                             const float* wl, int n, int h, int w, int T, int crop, double* sums, float* crop_deblur,
                             float* crop_gt, void* stream);

extern "C" int ie_eval_metrics_f32(const float* recon, const float* burst, int burst_pitch, const float* truth,
                                   const float* wl, int n, int h, int w, int T, int crop, double* sums, void* stream) {
  return eval_metrics_impl(recon, burst, burst_pitch, truth, wl, n, h, w, T, crop, sums, nullptr, nullptr, stream);
}

extern "C" int ie_eval_metrics_crops_f32(const float* recon, const float* burst, int burst_pitch, const float* truth,
                                         const float* wl, int n, int h, int w, int T, int crop, double* sums,
                                         float* crop_deblur, float* crop_gt, void* stream) {
  IE_REQUIRE(crop_deblur && crop_gt, "eval_metrics_crops: null crop pointer");
  return eval_metrics_impl(recon, burst, burst_pitch, truth, wl, n, h, w, T, crop, sums, crop_deblur, crop_gt, stream);
}

static int eval_metrics_impl(const float* recon, const float* burst, int burst_pitch, const float* truth,
                             const float* wl, int n, int h, int w, int T, int crop, double* sums, float* crop_deblur,
                             float* crop_gt, void* stream) {
  IE_REQUIRE(recon && burst && truth && wl && sums, "eval_metrics: null pointer");
  IE_REQUIRE(n > 0 && T >= 1 && T <= kMaxT && burst_pitch >= T, "eval_metrics: bad T=%d (max %d)", T, kMaxT);
  IE_REQUIRE(crop >= 0 && h > 2 * crop + 1 && w > 2 * crop + 1, "eval_metrics: image %dx%d too small for crop %d", h, w, crop);
  const int hc = h - 2 * crop, wc = w - 2 * crop;
  const bool aligned = ((reinterpret_cast<uintptr_t>(recon) | reinterpret_cast<uintptr_t>(burst) |
                         reinterpret_cast<uintptr_t>(truth)) & 15) == 0;
  if (aligned && !g_em_legacy && burst_pitch <= 2 * kMaxT + 2) {
    if (T == 4) return launch_eval_metrics_rows<4>(recon, burst, burst_pitch, truth, wl, n, h, w, T, crop, sums, crop_deblur, crop_gt, stream);
    if (T == 8) return launch_eval_metrics_rows<8>(recon, burst, burst_pitch, truth, wl, n, h, w, T, crop, sums, crop_deblur, crop_gt, stream);
    return launch_eval_metrics_rows<0>(recon, burst, burst_pitch, truth, wl, n, h, w, T, crop, sums, crop_deblur, crop_gt, stream);
  }
  // the tile kernel (unaligned views) has no crop by-product: the caller falls back to ie_invert_preproc_f32
  IE_REQUIRE(crop_deblur == nullptr, "eval_metrics_crops: needs 16-byte aligned recon / burst / truth (the row-streaming kernel)");
  const int tiles_x = (wc + kMT_W - 1) / kMT_W;
  // rows per block: a multiple of the 8-row sub-tile, as tall as still leaves ~4 blocks per SM in flight
  const int sub = (hc + kMT_H - 1) / kMT_H;
  long long want = 4ll * sm_count() / ((long long)tiles_x * n);
  if (want < 1) want = 1;
  if (want > sub) want = sub;
  const int subs_per_block = (int)((sub + want - 1) / want);
  const int rows_per_block = subs_per_block * kMT_H;
  const int ysplit = (hc + rows_per_block - 1) / rows_per_block;
  IE_REQUIRE(n <= 65535 && ysplit <= 65535, "eval_metrics: grid too large");
  const size_t smem = sizeof(float) * (size_t)(T + 1) * (kMT_H + 1) * (kMT_W + 1);
  IE_CUDA(cudaFuncSetAttribute(eval_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  eval_metrics_kernel<<<dim3(tiles_x, ysplit, n), 256, smem, S(stream)>>>(recon, burst, burst_pitch, truth, wl, h, w, T,
                                                                          crop, rows_per_block, sums);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

static int metric_totals(const double* sums, const double* ssim_sums, int n, int h, int w, int T, int crop,
                         double* totals, void* stream) {
  IE_REQUIRE(sums && totals && n > 0 && T >= 1 && T <= kMaxT, "metric_totals: bad arguments");
  IE_REQUIRE(crop >= 0 && h > 2 * crop + 1 && w > 2 * crop + 1, "metric_totals: image %dx%d too small for crop %d", h, w, crop);
  const double hc = h - 2 * crop, wc = w - 2 * crop;
  IE_REQUIRE(ssim_sums == nullptr || (hc >= kSTaps && wc >= kSTaps), "metric_totals: cropped image below 11x11 has no SSIM");
  metric_totals_kernel<<<1, 256, 0, S(stream)>>>(sums, n, T, hc * wc, (hc - 1) * (wc - 1) * 2, ssim_sums,
                                                 (hc - 2 * kSR) * (wc - 2 * kSR), totals);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

extern "C" int ie_metric_totals_f64(const double* sums, int n, int h, int w, int T, int crop, double* totals,
                                    void* stream) {
  return metric_totals(sums, nullptr, n, h, w, T, crop, totals, stream);
}

extern "C" int ie_metric_totals_ssim_f64(const double* sums, const double* ssim_sums, int n, int h, int w, int T,
                                         int crop, double* totals, void* stream) {
  IE_REQUIRE(ssim_sums, "metric_totals_ssim: null pointer");
  return metric_totals(sums, ssim_sums, n, h, w, T, crop, totals, stream);
}

extern "C" int ie_sqdiff_sum_f32(const float* a, const float* b, int n, long long count, double* sums, void* stream) {
  IE_REQUIRE(a && b && sums && n > 0 && n <= 65535 && count > 0, "sqdiff_sum: bad arguments");
  long long gx = (count / 4 + 256 * 4 - 1) / (256 * 4);
  const long long cap = (8ll * sm_count() + n - 1) / n;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  sqdiff_sum_kernel<<<dim3((unsigned)gx, n), 256, 0, S(stream)>>>(a, b, count, sums);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

extern "C" int ie_img_loss_sums_f32(const float* a, const float* b, int n, int h, int w, double* sums, void* stream) {
  IE_REQUIRE(a && b && sums && n > 0 && h > 1 && w > 1, "img_loss_sums: bad arguments");
  const long long total = (long long)n * h * w;
  const bool vec = (w % 4 == 0) && (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) &&
                   n <= 65535 && (h + kLossRows - 1) / kLossRows <= 65535;
  if (vec) {
    dim3 grid(ie_ceil_div(w / 4, 128), (h + kLossRows - 1) / kLossRows, n);
    img_loss_sums_vec_kernel<<<grid, 128, 0, S(stream)>>>(a, b, h, w, sums);
    IE_LAUNCH_CHECK();
    return IE_OK;
  }
  long long gx = (total + 255) / 256;
  if (gx > 8ll * sm_count()) gx = 8ll * sm_count();
  img_loss_sums_kernel<<<(unsigned)gx, 256, 0, S(stream)>>>(a, b, h, w, total, sums);
  IE_LAUNCH_CHECK();
  return IE_OK;
}

static int g_ssim_legacy = 0;
extern "C" int ie_ssim_tune(int legacy) {
  g_ssim_legacy = legacy;
  return IE_OK;
}

extern "C" int ie_ssim_f32(const float* a, const float* b, int n, int h, int w, double* sums, void* stream) {
  IE_REQUIRE(a && b && sums && n > 0 && h >= kSTaps && w >= kSTaps, "ssim: images must be at least 11x11");
  const int ho = h - 2 * kSR, wo = w - 2 * kSR;
  const int gy = (ho + kSsimRows - 1) / kSsimRows;
  IE_REQUIRE(n <= 65535 && gy <= 65535, "ssim: grid too large");
  // even width + 16-byte aligned images: two columns per thread, rows staged by TMA; anything else: one column per thread
  const bool rows_ok = g_ssim_legacy == 0 && (w % 2 == 0) &&
                       ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  if (rows_ok) {
    // rows per block: 128 (8 % vertical halo) for large images, down to 16 when small images would leave SMs idle
    const int gx = ie_ceil_div(wo / 2, 128);
    long long want = 4ll * sm_count() / ((long long)gx * n);
    if (want < 1) want = 1;
    int rows = (int)((ho + want - 1) / want);
    if (rows < 16) rows = 16;
    if (rows > kSsimRows) rows = kSsimRows;
    const size_t smem = sizeof(float) * 2 * 2 * kSTaps * ((((2 * 128 + 2 * kSR) + 6) >> 2) << 2);
    IE_CUDA(cudaFuncSetAttribute(ssim_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ssim_rows_kernel<<<dim3(gx, (ho + rows - 1) / rows, n), 128, smem, S(stream)>>>(a, b, n, h, w, rows, sums);
    IE_LAUNCH_CHECK();
    return IE_OK;
  }
  ssim_stream_kernel<<<dim3(ie_ceil_div(wo, 128), gy, n), 128, 0, S(stream)>>>(a, b, h, w, sums);
  IE_LAUNCH_CHECK();
  return IE_OK;
}
