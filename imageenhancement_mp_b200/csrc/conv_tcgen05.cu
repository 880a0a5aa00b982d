// Implicit-GEMM convolution for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma -> TMEM ->
// fused epilogue.  Replaces layers.Conv2D(c,3,'same',relu) / Conv2D(128,2,'valid',relu) of
// /root/reference/model_library.py:72-73,89-91,323-368.
//
// GEMM view.  Activations are bf16 "rasters" [R][pitch] (see include/imgenh_b200.h): one row per
// padded pixel, zero border.  For output row r and tap (i,j) the input row is r + shift(i,j), so
//     D[r][co] = sum_tap sum_c X[r + shift(tap)][c] * Wt[co][tap*cin + c]
// A tile  = 128 consecutive raster rows x 64 channels  (one 2-D TMA box, OOB rows zero-filled)
// B tile  = n_tile output channels x 64 K-elements of the packed weights [cout][ntaps*cin]
// Both K-major, SWIZZLE_128B, so one smem descriptor + 32-byte advance per UMMA_K=16 step.
// D lives in TMEM (128 lanes x n_tile fp32 columns), double-buffered (2 x 256 columns) so the
// epilogue of tile i overlaps the main loop of tile i+1.
//
// Warp roles (192 threads, persistent CTA, one per SM):
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..5  epilogue: tcgen05.ld -> bias/ReLU/border mask -> bf16 -> swizzled smem -> TMA store
//               (or fp32 global stores / per-pixel softmax for the two small heads)
#include "ie_common.cuh"
#include "ie_ptx.cuh"

namespace ie {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // bf16 elements = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kStgBytesPerWarp = 32 * 128;       // 32 rows x 64 bf16
constexpr int kMaxCout = 1024;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;      // TMEM columns per accumulator buffer

struct ConvKernelParams {
  int R;             // raster rows
  int plane;         // (h+2)*(w+2)
  int wp;            // w+2
  int hv, wv;        // valid output extent
  int ntaps;
  int tap_shift[9];
  int kblocks_per_tap;   // cin / 64
  int x_coff;
  int cin;
  int cout;
  int n_tile;
  int n_tiles;
  int m_tiles;
  int y_coff;
  int relu;
  int epilogue;
  int stages;
  int b_stage_bytes;     // n_tile*128 rounded up to 1024
  const float* bias;
  float* y_f32;
  float* y_aux;
};

struct SmemLayout {
  // dynamic smem, 1024-aligned base:
  //   [stages x (A 16 KiB | B b_stage_bytes)] [4 x 4 KiB staging] [bias 4 KiB] [barriers]
  uint8_t* base;
  int stages, b_bytes;
  __device__ uint8_t* a(int s) const { return base + s * (kABytes + b_bytes); }
  __device__ uint8_t* b(int s) const { return a(s) + kABytes; }
  __device__ uint8_t* stg(int warp) const { return base + stages * (kABytes + b_bytes) + warp * kStgBytesPerWarp; }
  __device__ float* bias() const { return reinterpret_cast<float*>(stg(4)); }
  __device__ uint64_t* bars() const { return reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias()) + kMaxCout * 4); }
};

static size_t conv_smem_bytes(int stages, int b_bytes) {
  return 1024 /*align slack*/ + static_cast<size_t>(stages) * (kABytes + b_bytes) + 4 * kStgBytesPerWarp +
         kMaxCout * 4 + (2 * kMaxStages + 4) * 8 + 16;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                  const __grid_constant__ CUtensorMap tm_y, const ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  SmemLayout sm;
  sm.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  sm.stages = p.stages;
  sm.b_bytes = p.b_stage_bytes;
  uint64_t* full_bar = sm.bars();
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int kblocks = p.ntaps * p.kblocks_per_tap;

  // ---- one-time setup
  for (int i = threadIdx.x; i < kMaxCout; i += kThreads) sm.bias()[i] = (i < p.cout && p.bias) ? p.bias[i] : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_y);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = kABytes + static_cast<uint32_t>(p.n_tile) * 128u;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles;
        const int n_idx = tile - m_tile * p.n_tiles;
        const int r0 = m_tile * kBlockM;
        const int n0 = n_idx * p.n_tile;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int row = r0 + p.tap_shift[tap];
          for (int kb = 0; kb < p.kblocks_per_tap; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_2d(sm.a(stage), &tm_a, &full_bar[stage], p.x_coff + kb * kBlockK, row);
            tma_load_2d(sm.b(stage), &tm_b, &full_bar[stage], tap * p.cin + kb * kBlockK, n0);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kBlockM, p.n_tile);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t use = static_cast<uint32_t>(it >> 1);
        mbar_wait(&tempty_bar[buf], (use & 1u) ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kAccStride);
        for (int kbi = 0; kbi < kblocks; ++kbi) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(sm.a(stage)));
          const uint64_t db = umma_desc_sw128(smem_u32(sm.b(stage)));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // +32 bytes along K inside the 128-byte swizzle row = +2 in the (addr >> 4) field
            umma_bf16_ss(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                         (kbi | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[buf]);       // accumulator complete
      }
    }
  } else {
    // ================================ epilogue ====================================
    const int q = warp & 3;                      // TMEM lane quadrant this warp may read
    const int row_in_tile = q * 32 + lane;
    uint8_t* stg = sm.stg(warp - 2);
    const float* sbias = sm.bias();
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_tile = tile / p.n_tiles;
      const int n_idx = tile - m_tile * p.n_tiles;
      const int r0 = m_tile * kBlockM;
      const int n0 = n_idx * p.n_tile;
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      const int r = r0 + row_in_tile;
      // position inside the image raster -> is this an interior (kept) output?
      const int img = r / p.plane;
      const int pr = r - img * p.plane;
      const int y = pr / p.wp;
      const int x = pr - y * p.wp;
      const bool valid = (r < p.R) && (y >= 1) && (y <= p.hv) && (x >= 1) && (x <= p.wv);

      mbar_wait(&tfull_bar[buf], use & 1u);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * kAccStride);

      if (p.epilogue == IE_EPI_BF16_RASTER) {
        const int chunks = p.n_tile >> 6;
        for (int c = 0; c < chunks; ++c) {
          uint32_t v0[32], v1[32];
          tmem_ld_x32(t_base + c * 64, v0);
          tmem_ld_x32(t_base + c * 64 + 32, v1);
          tmem_ld_wait();
          if (c == chunks - 1) {
            // all TMEM reads of this accumulator are done: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
          }
          uint32_t pk[32];
          const float* bs = sbias + n0 + c * 64;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(v0[2 * j]) + bs[2 * j];
            float b = __uint_as_float(v0[2 * j + 1]) + bs[2 * j + 1];
            if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            pk[j] = valid ? pack_bf16x2(a, b) : 0u;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(v1[2 * j]) + bs[32 + 2 * j];
            float b = __uint_as_float(v1[2 * j + 1]) + bs[32 + 2 * j + 1];
            if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            pk[16 + j] = valid ? pack_bf16x2(a, b) : 0u;
          }
          // staging buffer must have been read by the previous TMA store
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          // row `lane` of a 32x128B tile, 16-byte chunk j stored at j ^ (lane & 7)  (SWIZZLE_128B)
          uint8_t* rowp = stg + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 val = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) = val;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tm_y, stg, p.y_coff + n0 + c * 64, r0 + q * 32);
            tma_store_commit();
          }
        }
      } else {
        // small fp32 heads: n_tile <= 64, single N tile; each thread owns one pixel's channels
        float acc[64];
        {
          uint32_t v[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c * 16 < p.n_tile) {
              tmem_ld_x16(t_base + c * 16, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[c * 16 + j] = __uint_as_float(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[c * 16 + j] = 0.f;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        if (valid) {
          const long long pix = (static_cast<long long>(img) * p.hv + (y - 1)) * p.wv + (x - 1);
          float* dst = p.y_f32 + pix * p.cout;
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            if (j < p.cout) {
              float a = acc[j] + sbias[j];
              if (p.relu) a = fmaxf(a, 0.f);
              acc[j] = a;
              mx = fmaxf(mx, a);
            }
          }
          if (p.epilogue == IE_EPI_F32_SOFTMAX) {
            if (p.y_aux) {
              float* aux = p.y_aux + pix * p.cout;
#pragma unroll
              for (int j = 0; j < 64; ++j)
                if (j < p.cout) aux[j] = acc[j];
            }
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              if (j < p.cout) {
                acc[j] = __expf(acc[j] - mx);
                sum += acc[j];
              }
            }
            const float inv = 1.f / sum;
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (j < p.cout) dst[j] = acc[j] * inv;
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (j < p.cout) dst[j] = acc[j];
          }
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();   // all bulk stores complete before smem goes away
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// -------------------------------------------------------------------------------------------------
int choose_n_tile(int cout, int epilogue) {
  if (epilogue != IE_EPI_BF16_RASTER) return ((cout + 15) / 16) * 16;
  if (cout >= 256) return 256;
  if (cout > 64) return 128;
  return 64;
}

static int fill_taps(const ie_conv_desc* d, ConvKernelParams& p) {
  const int wp = d->w + 2;
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
    p.tap_shift[0] = 0;
  } else {
    return -1;
  }
  return 0;
}

int validate_conv_desc(const ie_conv_desc* d, const void* x, const void* w, void* y_bf16, float* y_f32) {
  IE_REQUIRE(d && x && w, "conv: null descriptor / input / weights");
  IE_REQUIRE(d->n_img > 0 && d->h > 0 && d->w > 0, "conv: bad raster size %d x %d x %d", d->n_img, d->h, d->w);
  IE_REQUIRE(d->cin > 0 && d->cin % 64 == 0, "conv: cin=%d must be a positive multiple of 64", d->cin);
  IE_REQUIRE(d->x_coff % 64 == 0 && d->x_coff + d->cin <= d->x_pitch, "conv: bad input slice (coff %d, cin %d, pitch %d)",
             d->x_coff, d->cin, d->x_pitch);
  IE_REQUIRE(d->x_pitch % 8 == 0, "conv: x_pitch must be a multiple of 8");
  IE_REQUIRE(d->cout > 0 && d->cout <= kMaxCout, "conv: cout=%d out of range", d->cout);
  IE_REQUIRE(d->hv >= 1 && d->hv <= d->h && d->wv >= 1 && d->wv <= d->w, "conv: bad valid extent %d x %d", d->hv, d->wv);
  IE_REQUIRE((long long)d->n_img * (d->h + 2) * (d->w + 2) < (1ll << 31) - 4096, "conv: raster too large for 32-bit rows");
  if (d->epilogue == IE_EPI_BF16_RASTER) {
    IE_REQUIRE(y_bf16, "conv: y_bf16 is null");
    IE_REQUIRE(d->cout % 64 == 0, "conv: bf16 raster epilogue needs cout %% 64 == 0 (got %d)", d->cout);
    IE_REQUIRE(d->y_coff % 64 == 0 && d->y_coff + d->cout <= d->y_pitch && d->y_pitch % 8 == 0,
               "conv: bad output slice (coff %d, cout %d, pitch %d)", d->y_coff, d->cout, d->y_pitch);
  } else if (d->epilogue == IE_EPI_F32_NHWC || d->epilogue == IE_EPI_F32_SOFTMAX) {
    IE_REQUIRE(y_f32, "conv: y_f32 is null");
    IE_REQUIRE(d->cout <= 64, "conv: fp32 epilogues need cout <= 64 (got %d)", d->cout);
  } else {
    IE_REQUIRE(false, "conv: unknown epilogue %d", d->epilogue);
  }
  return IE_OK;
}

}  // namespace ie

extern "C" int ie_conv2d_nhwc_bf16(const ie_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                                   void* y_bf16, float* y_f32, float* y_aux, void* stream) {
  using namespace ie;
  int rc = validate_conv_desc(d, x, w_packed, y_bf16, y_f32);
  if (rc) return rc;
  ConvKernelParams p{};
  IE_REQUIRE(fill_taps(d, p) == 0, "conv: unsupported kernel size %dx%d", d->kh, d->kw);
  const long long R = (long long)d->n_img * (d->h + 2) * (d->w + 2);
  p.R = (int)R;
  p.plane = (d->h + 2) * (d->w + 2);
  p.wp = d->w + 2;
  p.hv = d->hv;
  p.wv = d->wv;
  p.kblocks_per_tap = d->cin / 64;
  p.x_coff = d->x_coff;
  p.cin = d->cin;
  p.cout = d->cout;
  p.n_tile = choose_n_tile(d->cout, d->epilogue);
  p.n_tiles = (d->cout + p.n_tile - 1) / p.n_tile;
  p.m_tiles = (int)((R + kBlockM - 1) / kBlockM);
  p.y_coff = d->y_coff;
  p.relu = d->relu;
  p.epilogue = d->epilogue;
  p.bias = bias;
  p.y_f32 = y_f32;
  p.y_aux = y_aux;
  p.b_stage_bytes = ((p.n_tile * 128 + 1023) / 1024) * 1024;
  const size_t max_smem = 227 * 1024;
  int stages = kMaxStages;
  while (stages > 2 && conv_smem_bytes(stages, p.b_stage_bytes) > max_smem) --stages;
  p.stages = stages;
  const size_t smem = conv_smem_bytes(stages, p.b_stage_bytes);

  CUtensorMap tm_a, tm_b, tm_y;
  const int ktot = p.ntaps * d->cin;
  rc = make_tmap_2d_bf16(&tm_a, x, (uint64_t)d->x_pitch, (uint64_t)R, (uint64_t)d->x_pitch, 64, kBlockM);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tm_b, w_packed, (uint64_t)ktot, (uint64_t)(p.n_tiles * p.n_tile), (uint64_t)ktot, 64,
                         (uint32_t)p.n_tile);
  if (rc) return rc;
  if (d->epilogue == IE_EPI_BF16_RASTER) {
    rc = make_tmap_2d_bf16(&tm_y, y_bf16, (uint64_t)d->y_pitch, (uint64_t)R, (uint64_t)d->y_pitch, 64, 32);
    if (rc) return rc;
  } else {
    tm_y = tm_a;   // unused by the fp32 epilogues
  }

  IE_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  conv_igemm_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(tm_a, tm_b, tm_y, p);
  IE_LAUNCH_CHECK();
  return IE_OK;
}
