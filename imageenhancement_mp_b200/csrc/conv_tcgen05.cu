// Implicit-GEMM convolution for sm_100a: TMA -> 128B-swizzled smem -> tcgen05.mma -> TMEM -> fused
// epilogue.  Replaces layers.Conv2D(c,3,'same',relu) / Conv2D(128,2,'valid',relu) of
// /root/reference/model_library.py:72-73,89-91,323-368.
//
// GEMM view.  Activations are bf16 "rasters" [R][pitch] (see include/imgenh_b200.h): one row per
// padded pixel, zero border.  For output row r and tap (i,j) the input row is r + shift(i,j), so
//     D[r][co] = sum_tap sum_c X[r + shift(tap)][c] * Wt[co][tap*cin + c]
// Both operands are K-major with SWIZZLE_128B (rows of 64 bf16 = 128 B), so one shared-memory
// descriptor + a 32-byte advance per UMMA_K=16 step addresses every MMA.  D lives in TMEM
// (128 lanes x n_tile fp32 columns), double-buffered (2 x 256 columns) so the epilogue of tile i
// overlaps the main loop of tile i+1.  Persistent CTAs, one per SM, 192 threads (320 with a second epilogue set):
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + tcgen05.mma issuer (warp-uniform loop, one elected lane issues)
//   warps 2..5  epilogue: tcgen05.ld -> bias/ReLU/border mask -> bf16 -> swizzled smem -> TMA store
//               (or fp32 global stores / per-pixel softmax for the two small heads)
//
// Main-loop flavours (the first three share epilogue_loop):
//   conv_stream_kernel<IM2COL, MSUB>  wide layers (n_tile 128/256): every (tap, 64-channel block) is one pipeline stage =
//                         A box [128 rows x 64 ch] + B box [n_tile x 64]; MMA-bound (88-94 % tensor-pipe active in
//                         ncu).  IM2COL: dense NHWC activations, the A box is a TMA im2col-mode load (no border rows).
//                         MSUB = 2: two M tiles per work item share every B box (n_tile 128, the cout = 128 layers).
//   conv_streamT_kernel   cout = 128, cin >= 128: D^T = W X^T, 256 pixels as the N dimension, transposing epilogue.
//                         Split-K over CTAs for layers with very few tiles (+ splitk_finish_kernel).
//   conv_resident_kernel  1x1 / 2x2 / narrow 3x3 layers whose whole weight matrix fits in smem: loaded once per CTA;
//                         a stage is one A box of 128+2 rows per (filter row, channel block), and the horizontal
//                         taps read it at +0/+1/+2 rows through the descriptor start address.
//   conv_first_staged_kernel  first layer on the raw fp32 input (C = 3, 5, 10): stager warps bulk-copy the source rows a
//                         tile touches into a shared-memory ring, builder warps assemble the swizzled A tile from it.
//                         conv_first_kernel: the same with per-thread global gathers (unaligned sources).
//   conv_wide_kernel      3x3 layers with cout <= 64 and the fp32 coef head: the three horizontal taps are three column
//                         groups of ONE N = 192 accumulator, combined in the epilogue (own epilogues; see below).
// Every kernel is launched with programmatic stream serialization: cta_setup() ends with griddepcontrol.wait.
#include <climits>
#include <utility>
#include <cstdlib>

#include "ie_common.cuh"
#include "ie_ptx.cuh"

namespace ie {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // bf16 elements = 128 bytes = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kThreads = 192;        // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kThreads2 = 320;       // + warps 6-9: a second epilogue set (narrow layers: the epilogue is the critical path)
constexpr int kMaxStages = 8;
constexpr int kStgBytesPerWarp = 32 * 128;       // 32 rows x 64 bf16
constexpr int kMaxCout = 1024;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;      // TMEM columns per accumulator buffer
constexpr int kXchBytes = 2 * 4 * 2 * 32 * 4;     // wide-N exchange epilogue: [half][warp][up/down][32] fp32 edge rows
// tail of the dynamic smem for `nsets` epilogue sets (4 warps each): staging + bias + barriers + exchange rows
constexpr int tail_bytes(int nsets) {
  return nsets * 4 * kStgBytesPerWarp + kMaxCout * 4 + (2 * kMaxStages + 6) * 8 + 16 + nsets * kXchBytes;
}
constexpr int kTailBytes = tail_bytes(1);
constexpr int kTailBytes2 = tail_bytes(2);
constexpr size_t kMaxSmem = 227 * 1024;

// Output-side description shared by both kernels.
struct EpiParams {
  int R;             // raster rows
  int plane;         // (h+1)*(w+1)
  int wp;            // w+1
  int hv, wv;        // valid output extent
  int cout;
  int n_tile;
  int n_tiles;
  int m_tiles;
  int y_coff;
  int relu;
  int epilogue;
  const float* bias;
  float* y_f32;       // already advanced by the output slice's first channel
  float* y_aux;
  int f32_pitch;      // floats per pixel of y_f32 / y_aux (= cout unless the caller writes a channel slice)
  int dense;          // x / y are dense NHWC (R = n*h*w rows, no border rows to mask)
  // split-K: tiles [split_first, num_tiles) have their K loop cut into `ksplit` work items computed by different CTAs
  // (split_first = 0: every tile - tiny M, more SMs stream the weights; split_first > 0: only the tiles of the last,
  // partial wave of a persistent grid - the other SMs would idle while a few finish whole tiles).  A split item writes
  // its raw fp32 accumulator to its own slab part[item - split_first][128][n_tile]; splitk_finish_kernel sums the
  // slabs of a tile in a fixed order and applies bias / ReLU / mask (deterministic: no atomics).  ksplit = 1: off.
  int ksplit;
  int split_first;
  float* part;
  // M sub-tiles per work item (streaming kernel with N tile 128: two 128-row A tiles share every weight tile, their
  // accumulators sit side by side in one TMEM buffer; everything else: 1)
  int msub;
  // r / plane and pr / wp by multiply-high (FastDiv below): every epilogue warp decodes its raster row once per tile
  uint32_t plane_m, plane_s, wp_m, wp_s;
};

// floor(n / d) for 0 <= n < 2^31, d >= 2: with l = ceil(log2 d), m = floor(2^(31+l) / d) + 1, q = umulhi(n, m) >> (l - 1)
// (a runtime integer division is ~20 dependent instructions; the epilogue / builder / stager warps each run one long
//  dependent instruction stream per tile, which is what bounds the short-K kernels)
struct FastDiv {
  uint32_t m, s;
};
static FastDiv make_fastdiv(uint32_t d) {
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  FastDiv f;
  f.m = static_cast<uint32_t>((1ull << (31 + l)) / d + 1);
  f.s = l - 1;
  return f;
}
static void set_raster_dims(EpiParams& e, int plane, int wp) {
  e.plane = plane;
  e.wp = wp;
  const FastDiv fp = make_fastdiv((uint32_t)plane), fw = make_fastdiv((uint32_t)wp);
  e.plane_m = fp.m; e.plane_s = fp.s; e.wp_m = fw.m; e.wp_s = fw.s;
}
__device__ __forceinline__ int fastdiv(int n, uint32_t m, uint32_t s) {
  return static_cast<int>(__umulhi(static_cast<uint32_t>(n), m) >> s);
}

// Tail of the dynamic smem (after the operand buffers): staging, bias, barriers.
struct SmemTail {
  uint8_t* p;
  int nsets = 1;         // epilogue sets (4 warps each)
  __device__ uint8_t* stg(int warp) const { return p + warp * kStgBytesPerWarp; }
  __device__ float* bias() const { return reinterpret_cast<float*>(p + nsets * 4 * kStgBytesPerWarp); }
  __device__ uint64_t* full() const {
    return reinterpret_cast<uint64_t*>(p + nsets * 4 * kStgBytesPerWarp + kMaxCout * 4);
  }
  __device__ uint64_t* empty() const { return full() + kMaxStages; }
  __device__ uint64_t* tfull() const { return empty() + kMaxStages; }     // [2]
  __device__ uint64_t* tempty() const { return tfull() + 2; }              // [2]
  __device__ uint64_t* bres() const { return tempty() + 2; }               // [1] resident weights landed
  __device__ uint32_t* tmem_slot() const { return reinterpret_cast<uint32_t*>(bres() + 1); }
  __device__ float* xch() const { return reinterpret_cast<float*>(full() + 2 * kMaxStages + 8); }   // 16-byte aligned
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// Work item -> (tile, K-range index): items [0, split_first) are whole tiles, the rest are (tile, ks) pairs.
struct WorkItem {
  int tile, ks;
  bool split;
};
__device__ __forceinline__ int num_work_items(const EpiParams& e) {
  const int tiles = e.m_tiles * e.n_tiles;
  return e.split_first + (tiles - e.split_first) * e.ksplit;
}
__device__ __forceinline__ WorkItem decode_item(const EpiParams& e, int item) {
  WorkItem w;
  if (item < e.split_first) {
    w.tile = item; w.ks = 0; w.split = false;
  } else {
    const int j = item - e.split_first;
    w.tile = e.split_first + j / e.ksplit;
    w.ks = j - (w.tile - e.split_first) * e.ksplit;
    w.split = e.ksplit > 1;
  }
  return w;
}

// One-time CTA setup common to both kernels; returns the TMEM base address.
__device__ __forceinline__ uint32_t cta_setup(const SmemTail& t, const EpiParams& e, int stages, int warp, int lane,
                                              uint32_t full_count = 1) {
  pdl_trigger();                      // the next kernel's CTAs may be scheduled (they wait in their own pdl_wait)
  for (int i = threadIdx.x; i < kMaxCout; i += blockDim.x) t.bias()[i] = (i < e.cout && e.bias) ? e.bias[i] : 0.f;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&t.full()[s], full_count);
      mbar_init(&t.empty()[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&t.tfull()[b], 1);
      mbar_init(&t.tempty()[b], 4);   // one arrive per epilogue warp
    }
    mbar_init(t.bres(), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(t.tmem_slot(), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                         // everything above overlapped the previous kernel's tail; its outputs are visible now
  return *t.tmem_slot();
}

// 64 accumulator columns of one row + bias -> 32 packed bf16 pairs
template <bool RELU>
__device__ __forceinline__ void bias_pack_64(uint32_t (&pk)[32], const uint32_t (&v0)[32], const uint32_t (&v1)[32], uint32_t bs) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float4 bb = lds128(bs + j * 8);
    pk[j] = bias_act_pack<RELU>(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]), bb.x, bb.y);
    pk[j + 1] = bias_act_pack<RELU>(__uint_as_float(v0[2 * j + 2]), __uint_as_float(v0[2 * j + 3]), bb.z, bb.w);
  }
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float4 bb = lds128(bs + 128 + j * 8);
    pk[16 + j] = bias_act_pack<RELU>(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]), bb.x, bb.y);
    pk[16 + j + 1] = bias_act_pack<RELU>(__uint_as_float(v1[2 * j + 2]), __uint_as_float(v1[2 * j + 3]), bb.z, bb.w);
  }
}

// Epilogue warps (4 warps, one TMEM lane quadrant each): walk the CTA's tiles and drain the accumulator
// buffers as the MMA warp completes them.
// With two sets (warps 2-5 and 6-9) set s takes the CTA's tiles it = s, s+2, ... and therefore always drains TMEM
// buffer s: the sets never touch the same tile, staging buffer or barrier phase.
__device__ __forceinline__ void epilogue_loop(const EpiParams& p, const CUtensorMap* tm_y, const SmemTail& t,
                                              uint32_t tmem_base, int warp, int lane) {
  const int set = (warp - 2) >> 2;
  const int q = warp & 3;                      // TMEM lane quadrant this warp may read
  const int row_in_tile = q * 32 + lane;
  uint8_t* stg = t.stg(warp - 2);
  const float* sbias = t.bias();
  uint64_t* tfull_bar = t.tfull();
  uint64_t* tempty_bar = t.tempty();
  const int num_items = num_work_items(p);
  int it = set;
  for (int item = blockIdx.x + set * gridDim.x; item < num_items; item += t.nsets * gridDim.x, it += t.nsets) {
    const WorkItem wi = decode_item(p, item);
    const int tile = wi.tile;
    const int m_tile = tile / p.n_tiles;
    const int n_idx = tile - m_tile * p.n_tiles;
    const int r0 = m_tile * kBlockM;
    const int n0 = n_idx * p.n_tile;
    const int buf = it & 1;
    const uint32_t use = static_cast<uint32_t>(it >> 1);
    const int r = r0 + row_in_tile;
    // position inside the image raster -> is this an interior (kept) output?
    const int img = fastdiv(r, p.plane_m, p.plane_s);
    const int pr = r - img * p.plane;
    const int y = fastdiv(pr, p.wp_m, p.wp_s);
    const int x = pr - y * p.wp;
    const bool valid = p.dense ? (r < p.R) : ((r < p.R) && (y >= 1) && (y <= p.hv) && (x < p.wv));

    mbar_wait(&tfull_bar[buf], use & 1u);
    tc_fence_after();
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * kAccStride);

    if (wi.split) {
      // raw fp32 partial sums of this K range: row row_in_tile of the item's own slab [128][n_tile]
      float* dst = p.part + (static_cast<long long>(item - p.split_first) * kBlockM + row_in_tile) * p.n_tile;
      const int chunks = p.n_tile >> 5;
      for (int c = 0; c < chunks; ++c) {
        uint32_t v[32];
        tmem_ld_x32(t_base + c * 32, v);
        tmem_ld_wait();
        if (c == chunks - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(dst + c * 32 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    } else if (p.epilogue == IE_EPI_BF16_RASTER) {
      const int chunks = p.n_tile >> 6;
      for (int sub = 0; sub < p.msub; ++sub) {
        // sub-tile `sub` of the item: rows (m_tile * msub + sub) * 128 ..., accumulator columns sub * n_tile ...
        int rs0 = r0;
        bool valid_s = valid;
        if (p.msub > 1) {
          rs0 = (m_tile * p.msub + sub) * kBlockM;
          const int rs = rs0 + row_in_tile;
          const int img_s = fastdiv(rs, p.plane_m, p.plane_s);
          const int pr_s = rs - img_s * p.plane;
          const int y_s = fastdiv(pr_s, p.wp_m, p.wp_s);
          const int x_s = pr_s - y_s * p.wp;
          valid_s = p.dense ? (rs < p.R) : ((rs < p.R) && (y_s >= 1) && (y_s <= p.hv) && (x_s < p.wv));
        }
        const uint32_t ts_base = t_base + static_cast<uint32_t>(sub * p.n_tile);
        for (int c = 0; c < chunks; ++c) {
          uint32_t v0[32], v1[32];
          tmem_ld_x32(ts_base + c * 64, v0);
          tmem_ld_x32(ts_base + c * 64 + 32, v1);
          tmem_ld_wait();
          if (c == chunks - 1 && sub == p.msub - 1) {
            // all TMEM reads of this accumulator are done: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
          }
          uint32_t pk[32];
          const uint32_t bs = smem_u32(sbias + n0 + c * 64);
          if (p.relu) bias_pack_64<true>(pk, v0, v1, bs);
          else bias_pack_64<false>(pk, v0, v1, bs);
          // staging buffer must have been read by the previous TMA store
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          // row `lane` of a 32x128B tile, 16-byte chunk j stored at j ^ (lane & 7)  (SWIZZLE_128B); border / masked
          // rows are written as zeros (a branch per row instead of a select per word)
          const uint32_t rowa = smem_u32(stg) + lane * 128;
          if (valid_s) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128u(rowa + ((j ^ (lane & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) sts128u(rowa + (j << 4), 0u, 0u, 0u, 0u);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(tm_y, stg, p.y_coff + n0 + c * 64, rs0 + q * 32);
            tma_store_commit();
          }
        }
      }
    } else {
      // small fp32 heads (single N tile, cout <= 256): each thread owns one pixel's channels and walks
      // them 16 TMEM columns at a time; the softmax re-reads TMEM instead of holding them in registers.
      const int nch = (p.cout + 15) >> 4;
      const bool sm_mode = (p.epilogue == IE_EPI_F32_SOFTMAX);
      const long long pix = (static_cast<long long>(img) * p.hv + (y - 1)) * p.wv + x;
      if (nch == 1) {
        // <= 16 channels (the `coef` head): one TMEM read, everything in registers.  Valid pixels of a warp's
        // 32 raster rows are consecutive in the NHWC output (the skipped border pixels have no output slot), so
        // the rows are compacted through the warp's staging buffer and leave as fully coalesced 4-byte stores.
        uint32_t v[16];
        tmem_ld_x16(t_base, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        float a[16];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          a[j] = __uint_as_float(v[j]) + sbias[j];
          if (p.relu) a[j] = fmaxf(a[j], 0.f);
          if (j < p.cout) mx = fmaxf(mx, a[j]);
        }
        float e[16];
        if (sm_mode) {
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            e[j] = (j < p.cout) ? __expf(a[j] - mx) : 0.f;
            sum += e[j];
          }
          const float inv = 1.f / sum;
#pragma unroll
          for (int j = 0; j < 16; ++j) e[j] *= inv;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) e[j] = a[j];
        }
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        if (bal) {
          const int rank = __popc(bal & ((1u << lane) - 1u));
          const int nvalid = __popc(bal);
          const long long pix_first = __shfl_sync(0xffffffffu, pix, __ffs(bal) - 1);
          float* sf = reinterpret_cast<float*>(stg);
          const int total = nvalid * p.cout;
          for (int pass = 0; pass < ((sm_mode && p.y_aux) ? 2 : 1); ++pass) {
            __syncwarp();
            if (valid) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < p.cout) sf[rank * p.cout + j] = pass ? a[j] : e[j];
            }
            __syncwarp();
            float* dst = (pass ? p.y_aux : p.y_f32) + pix_first * p.cout;
            for (int i = lane; i < total; i += 32) dst[i] = sf[i];
          }
        }
        continue;
      }
      float* dst = p.y_f32 + pix * p.f32_pitch;
      float* aux = p.y_aux ? p.y_aux + pix * p.f32_pitch : nullptr;
      float mx = -INFINITY, inv = 1.f;
      uint32_t v[16];
      if (sm_mode) {
        for (int c = 0; c < nch; ++c) {
          tmem_ld_x16(t_base + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (c * 16 + j < p.cout) {
              float a = __uint_as_float(v[j]) + sbias[c * 16 + j];
              if (p.relu) a = fmaxf(a, 0.f);
              mx = fmaxf(mx, a);
            }
          }
        }
        float sum = 0.f;
        for (int c = 0; c < nch; ++c) {
          tmem_ld_x16(t_base + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (c * 16 + j < p.cout) {
              float a = __uint_as_float(v[j]) + sbias[c * 16 + j];
              if (p.relu) a = fmaxf(a, 0.f);
              sum += __expf(a - mx);
            }
          }
        }
        inv = 1.f / sum;
      }
      for (int c = 0; c < nch; ++c) {
        tmem_ld_x16(t_base + c * 16, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int ch = c * 16 + j;
            if (ch < p.cout) {
              float a = __uint_as_float(v[j]) + sbias[ch];
              if (p.relu) a = fmaxf(a, 0.f);
              if (sm_mode) {
                if (aux) aux[ch] = a;
                dst[ch] = __expf(a - mx) * inv;
              } else {
                dst[ch] = a;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  }
  if (lane == 0) tma_store_wait<0>();   // all bulk stores complete before smem goes away
}

// =================================================================================================
// Streaming kernel: stage = (tap, 64-channel block) -> A box [128 x 64] + B box [n_tile x 64]
// =================================================================================================
struct StreamParams {
  EpiParams e;
  int ntaps;
  int tap_shift[9];
  int kblocks_per_tap;   // cin / 64
  int x_coff;
  int cin;
  int stages;
  int b_stage_bytes;     // n_tile*128 rounded up to 1024
  // dense NHWC input (IM2COL): image size, filter width and zero padding - tap t is offset (t % kw, t / kw) of the
  // im2col tensor map, a tile is 128 consecutive pixels of the [n][h][w] index space
  int img_h, img_w, kw, pad;
};

// MSUB = 2 (N tile 128 only): a work item is TWO consecutive 128-row M tiles; every pipeline stage holds both A tiles and
// ONE weight tile, the two accumulators are columns [0, 128) and [128, 256) of the TMEM buffer.  The N = 128 layers
// were bound by what a stage moves (32 KB per 256 MMA cycles, 481 cycles per K block measured): per MMA cycle the pair
// moves 25 % less and every barrier round trip covers twice the math.
template <bool IM2COL, int MSUB>
__global__ void __launch_bounds__(kThreads, 1)
conv_stream_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const __grid_constant__ CUtensorMap tm_y, const StreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024(smem_raw);
  const int stage_bytes = MSUB * kABytes + p.b_stage_bytes;
  SmemTail t{base + p.stages * stage_bytes};
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.e.m_tiles * p.e.n_tiles;
  const int kblocks = p.ntaps * p.kblocks_per_tap;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_y);
  }
  const uint32_t tmem_base = cta_setup(t, p.e, p.stages, warp, lane);
  uint64_t* full_bar = t.full();
  uint64_t* empty_bar = t.empty();

  // work item = whole tile, or (tile, ks) = K blocks [ks * kb_per, ...) of a split tile
  const int num_items = num_work_items(p.e);
  const int kb_per = (kblocks + p.e.ksplit - 1) / p.e.ksplit;
  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = MSUB * kABytes + static_cast<uint32_t>(p.e.n_tile) * 128u;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const WorkItem wi = decode_item(p.e, item);
        const int tile = wi.tile;
        const int m_tile = tile / p.e.n_tiles;
        const int n_idx = tile - m_tile * p.e.n_tiles;
        const int r0 = m_tile * (kBlockM * MSUB);
        const int n0 = n_idx * p.e.n_tile;
        int px0[MSUB], py0[MSUB], pn0[MSUB];              // IM2COL: first pixel of each sub-tile in the pixel-box frame
#pragma unroll
        for (int sub = 0; sub < MSUB; ++sub) {
          px0[sub] = py0[sub] = pn0[sub] = 0;
          if constexpr (IM2COL) {
            const int rs = r0 + sub * kBlockM;
            const int plane = p.img_h * p.img_w;
            pn0[sub] = rs / plane;
            const int rem = rs - pn0[sub] * plane;
            py0[sub] = rem / p.img_w;
            px0[sub] = rem - py0[sub] * p.img_w - p.pad;
            py0[sub] -= p.pad;
          }
        }
        const int kb_begin = wi.split ? wi.ks * kb_per : 0;
        const int kb_end = (wi.split && kb_begin + kb_per < kblocks) ? kb_begin + kb_per : kblocks;
        int tap = kb_begin / p.kblocks_per_tap;
        int kb = kb_begin - tap * p.kblocks_per_tap;
        for (int kbi = kb_begin; kbi < kb_end; ++kbi) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          uint8_t* a_dst = base + stage * stage_bytes;
#pragma unroll
          for (int sub = 0; sub < MSUB; ++sub) {
            if constexpr (IM2COL) {
              const int ti = tap / p.kw, tj = tap - ti * p.kw;
              tma_load_im2col_4d(a_dst + sub * kABytes, &tm_a, &full_bar[stage], p.x_coff + kb * kBlockK, px0[sub], py0[sub],
                                 pn0[sub], tj, ti);
            } else {
              tma_load_2d(a_dst + sub * kABytes, &tm_a, &full_bar[stage], p.x_coff + kb * kBlockK,
                          r0 + sub * kBlockM + p.tap_shift[tap]);
            }
          }
          tma_load_2d(a_dst + MSUB * kABytes, &tm_b, &full_bar[stage], tap * p.cin + kb * kBlockK, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          if (++kb == p.kblocks_per_tap) { kb = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    // Warp-uniform loop (descriptors stay in uniform registers); one elected lane issues.
    const uint32_t idesc = umma_idesc_bf16(kBlockM, p.e.n_tile);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(base));
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(base + MSUB * kABytes));
    const uint32_t stage_stride = static_cast<uint32_t>(stage_bytes) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const WorkItem wi = decode_item(p.e, item);
      const int kb_begin = wi.split ? wi.ks * kb_per : 0;
      const int kb_end = (wi.split && kb_begin + kb_per < kblocks) ? kb_begin + kb_per : kblocks;
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      mbar_wait(&t.tempty()[buf], (use & 1u) ^ 1u);   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kAccStride);
      for (int kbi = kb_begin; kbi < kb_end; ++kbi) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo0 + stage * stage_stride;
          const uint32_t b_lo = b_lo0 + stage * stage_stride;
          // +32 bytes along K inside the 128-byte swizzle row = +2 in the (addr >> 4) field
#pragma unroll
          for (int sub = 0; sub < MSUB; ++sub) {
            const uint32_t a_s = a_lo + sub * (kABytes >> 4);
            const uint32_t d_s = d_tmem + static_cast<uint32_t>(sub * p.e.n_tile);
            umma_bf16_ss_lo(d_s, a_s, b_lo, idesc, kbi != kb_begin ? 1u : 0u);
            umma_bf16_ss_lo(d_s, a_s + 2, b_lo + 2, idesc, 1u);
            umma_bf16_ss_lo(d_s, a_s + 4, b_lo + 4, idesc, 1u);
            umma_bf16_ss_lo(d_s, a_s + 6, b_lo + 6, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);                            // frees the smem slot when these retire
          if (kbi == kb_end - 1) umma_commit(&t.tfull()[buf]);       // accumulator complete
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    epilogue_loop(p.e, &tm_y, t, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// Transposed flavour for the cout = 128 layers: D^T = W * X^T.  The A operand is the weight tile [128 cout][64 k], the
// B operand a box of 256 PIXELS [256][64 k] (same K-major 128B-swizzled tiles as everywhere), the accumulator is
// [128 lanes = output channels][256 columns = pixels].  Why: an M = 128, N = 128 MMA reads 8 KB of operands from shared
// memory in its 64 cycles - the whole 128 B/clk port, the N = 128 kernels stop at 60 - 67 % tensor-pipe - while the
// M = 128, N = 256 MMA used here reads 12 KB in 128 cycles like the big layers (96 B/clk, 95 %).  The price is a
// transposing epilogue: the thread that owns TMEM lane c holds ONE channel of 32 consecutive pixels per tcgen05.ld, so
// the bf16 values go to the [pixel][channel] staging tile with 2-byte stores (a warp's 32 channels of one pixel are 64
// contiguous bytes: one wavefront), the four epilogue warps meet at a named barrier and one thread issues the two
// 64-channel TMA stores of the 32 pixels.  Same accumulation order as the other kernels: bit-identical output.
// =================================================================================================
constexpr int kPixT = 256;                 // pixels (GEMM N) per work item
__device__ __forceinline__ void epi_bar_sync(int set) { asm volatile("bar.sync %0, 128;" ::"r"(set + 1) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<unsigned short>(v)) : "memory");
}

template <bool IM2COL>
__global__ void __launch_bounds__(kThreads, 1)
conv_streamT_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                    const __grid_constant__ CUtensorMap tm_y, const StreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024(smem_raw);
  constexpr int kWBytes = 128 * 128;                       // weight tile: 128 cout x 64 k
  constexpr int kXBytes = kPixT * 128;                     // pixel box: 256 x 64 k
  constexpr int stage_bytes = kWBytes + kXBytes;
  SmemTail t{base + p.stages * stage_bytes};
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = p.e.m_tiles;                       // items of 256 pixels
  const int kblocks = p.ntaps * p.kblocks_per_tap;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_y);
  }
  const uint32_t tmem_base = cta_setup(t, p.e, p.stages, warp, lane);
  uint64_t* full_bar = t.full();
  uint64_t* empty_bar = t.empty();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int r0 = item * kPixT;
        int px0 = 0, py0 = 0, pn0 = 0;
        if constexpr (IM2COL) {
          const int plane = p.img_h * p.img_w;
          pn0 = r0 / plane;
          const int rem = r0 - pn0 * plane;
          py0 = rem / p.img_w;
          px0 = rem - py0 * p.img_w - p.pad;
          py0 -= p.pad;
        }
        int tap = 0, kb = 0;
        for (int kbi = 0; kbi < kblocks; ++kbi) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          uint8_t* dst = base + stage * stage_bytes;
          tma_load_2d(dst, &tm_w, &full_bar[stage], tap * p.cin + kb * kBlockK, 0);
          if constexpr (IM2COL) {
            const int ti = tap / p.kw, tj = tap - ti * p.kw;
            tma_load_im2col_4d(dst + kWBytes, &tm_x, &full_bar[stage], p.x_coff + kb * kBlockK, px0, py0, pn0, tj, ti);
          } else {
            tma_load_2d(dst + kWBytes, &tm_x, &full_bar[stage], p.x_coff + kb * kBlockK, r0 + p.tap_shift[tap]);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          if (++kb == p.kblocks_per_tap) { kb = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    const uint32_t idesc = umma_idesc_bf16(128, kPixT);
    const uint32_t w_lo0 = umma_desc_lo(smem_u32(base));
    const uint32_t x_lo0 = umma_desc_lo(smem_u32(base + kWBytes));
    constexpr uint32_t stage_stride = static_cast<uint32_t>(stage_bytes) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      mbar_wait(&t.tempty()[buf], (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kAccStride);
      for (int kbi = 0; kbi < kblocks; ++kbi) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w_lo = w_lo0 + stage * stage_stride;
          const uint32_t x_lo = x_lo0 + stage * stage_stride;
          umma_bf16_ss_lo(d_tmem, w_lo, x_lo, idesc, kbi != 0 ? 1u : 0u);
          umma_bf16_ss_lo(d_tmem, w_lo + 2, x_lo + 2, idesc, 1u);
          umma_bf16_ss_lo(d_tmem, w_lo + 4, x_lo + 4, idesc, 1u);
          umma_bf16_ss_lo(d_tmem, w_lo + 6, x_lo + 6, idesc, 1u);
          umma_commit(&empty_bar[stage]);
          if (kbi == kblocks - 1) umma_commit(&t.tfull()[buf]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ================================ transposing epilogue ========================
    const EpiParams& e = p.e;
    const int q = warp & 3;                          // TMEM lane quadrant = output channels 32q .. 32q + 31
    const int co = q * 32 + lane;
    const float bias_c = t.bias()[co];
    uint8_t* stg0 = t.stg(0);                        // 16 KB: two [32 px][128 ch] bf16 buffers, each two 64-channel halves
    const bool issuer = (warp == 2 && lane == 0);
    // byte offset of channel `co` in pixel row j of a buffer: half (co / 64) * 4 KB + row * 128 + swizzled 16-byte chunk
    const uint32_t half_off = static_cast<uint32_t>(co >> 6) * 4096u;
    const uint32_t chunk16 = static_cast<uint32_t>((co & 63) >> 3);
    const uint32_t in_chunk = static_cast<uint32_t>(co & 7) * 2u;
    int it = 0;
    int nchunk = 0;                                  // staging buffers alternate over the whole kernel
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      const int r0 = item * kPixT;
      mbar_wait(&t.tfull()[buf], use & 1u);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * kAccStride);
      for (int c = 0; c < kPixT / 32; ++c, ++nchunk) {
        uint32_t v[32];
        tmem_ld_x32(t_base + c * 32, v);
        tmem_ld_wait();
        if (c == kPixT / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t.tempty()[buf]);
        }
        // which of the chunk's 32 pixels are kept outputs: lane j decodes pixel j, the ballot is every lane's mask
        const int r = r0 + c * 32 + lane;
        bool valid;
        if (e.dense) {
          valid = r < e.R;
        } else {
          const int img = fastdiv(r, e.plane_m, e.plane_s);
          const int pr = r - img * e.plane;
          const int y = fastdiv(pr, e.wp_m, e.wp_s);
          const int x = pr - y * e.wp;
          valid = (r < e.R) && (y >= 1) && (y <= e.hv) && (x < e.wv);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, valid);
        uint8_t* sb = stg0 + (nchunk & 1) * 8192;
        // the buffer was the source of the TMA stores two chunks ago
        if (issuer) tma_store_wait_read<1>();
        epi_bar_sync(0);
        const uint32_t colbase = smem_u32(sb) + half_off + in_chunk;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s = __uint_as_float(v[j]) + bias_c;
          uint32_t w = e.relu ? pack_relu_bf16x2(s, 0.f) : pack_bf16x2(s, 0.f);
          w = ((mask >> j) & 1u) ? w : 0u;
          sts16(colbase + j * 128 + ((chunk16 ^ (j & 7)) << 4), w);
        }
        fence_proxy_async_smem();
        epi_bar_sync(0);
        if (issuer) {
          tma_store_2d(&tm_y, sb, e.y_coff, r0 + c * 32);
          tma_store_2d(&tm_y, sb + 4096, e.y_coff + 64, r0 + c * 32);
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// Resident-weights kernel: all taps' weights stay in smem; stage = one A box of 128+ndx-1 rows per
// (filter row, channel block); the ndx horizontal taps address it at +dx rows.
// =================================================================================================
struct ResidentParams {
  EpiParams e;
  int ndy, ndx;
  int dy_shift[3];       // raster-row shift of each filter row's box start (includes the leftmost dx)
  int kb;                // cin / 64
  int x_coff;
  int cin;
  int stages;
  int box_rows;          // 128 + ndx - 1
  int a_box_bytes;       // box_rows*128 rounded up to 1024
  int a_slot_bytes;      // G boxes
  int b_tile_bytes;      // n_tile*128
};

// NDX = horizontal taps; G = filter rows fused into one pipeline stage (G = 3 needs kb == 1: the whole
// 3x3x64 tile is then ONE stage of 36 MMAs per barrier round trip).
template <int NDX, int G>
__global__ void __launch_bounds__(kThreads2, 1)
conv_resident_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                     const __grid_constant__ CUtensorMap tm_y, const ResidentParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024(smem_raw);
  const int ntaps = p.ndy * p.ndx;
  const int b_bytes = ntaps * p.kb * p.b_tile_bytes;      // resident weights, [tap][kb][n_tile x 64]
  uint8_t* a_base_ptr = base + b_bytes;
  SmemTail t{a_base_ptr + p.stages * p.a_slot_bytes, 2};
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.e.m_tiles;                      // single N tile
  const int stages_per_tile = p.ndy * p.kb / G;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_y);
  }
  const uint32_t tmem_base = cta_setup(t, p.e, p.stages, warp, lane);
  uint64_t* full_bar = t.full();
  uint64_t* empty_bar = t.empty();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      // weights: once per CTA
      mbar_arrive_expect_tx(t.bres(), static_cast<uint32_t>(b_bytes));
      for (int tap = 0; tap < ntaps; ++tap)
        for (int kb = 0; kb < p.kb; ++kb)
          tma_load_2d(base + (tap * p.kb + kb) * p.b_tile_bytes, &tm_b, t.bres(), tap * p.cin + kb * kBlockK, 0);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = static_cast<uint32_t>(p.box_rows) * 128u * G;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int r0 = tile * kBlockM;
        if constexpr (G == 1) {
          for (int dy = 0; dy < p.ndy; ++dy) {
            const int row = r0 + p.dy_shift[dy];
            for (int kb = 0; kb < p.kb; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1u);
              mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
              tma_load_2d(a_base_ptr + stage * p.a_slot_bytes, &tm_a, &full_bar[stage], p.x_coff + kb * kBlockK, row);
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
          }
        } else {
          // G filter rows (kb == 1) share one slot and ONE barrier: every mbarrier wait costs the MMA warp
          // ~100+ cycles, and at n_tile <= 64 those round trips, not the tensor pipe, bound the tile time
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
#pragma unroll
          for (int g = 0; g < G; ++g)
            tma_load_2d(a_base_ptr + stage * p.a_slot_bytes + g * p.a_box_bytes, &tm_a, &full_bar[stage], p.x_coff,
                        r0 + p.dy_shift[g]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    const uint32_t idesc = umma_idesc_bf16(kBlockM, p.e.n_tile);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(a_base_ptr));
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(base));
    const uint32_t a_stride = static_cast<uint32_t>(p.a_slot_bytes) >> 4;
    const uint32_t a_box = static_cast<uint32_t>(p.a_box_bytes) >> 4;
    const uint32_t b_tile = static_cast<uint32_t>(p.b_tile_bytes) >> 4;
    const uint32_t b_tap = static_cast<uint32_t>(p.kb) * b_tile;      // distance between horizontal taps
    mbar_wait(t.bres(), 0);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      mbar_wait(&t.tempty()[buf], (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kAccStride);
      uint32_t b_row = b_lo0;                                         // weights of filter row dy, block kb
      for (int si = 0; si < stages_per_tile; ++si) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_slot_lo = a_lo0 + stage * a_stride;
#pragma unroll
          for (int g = 0; g < G; ++g) {
#pragma unroll
            for (int dx = 0; dx < NDX; ++dx) {
              // tap (dy,dx) reads the box dx rows further down: +128 B (= +8) in the start address.  The
              // hardware swizzles on absolute smem address bits, so a start that is not 1024-byte aligned
              // needs no base-offset field (verified on B200: tools/try_resident.py)
              const uint32_t a = a_slot_lo + g * a_box + dx * 8;
              const uint32_t b = b_row + (g * NDX + dx) * b_tap;
              umma_bf16_ss_lo(d_tmem, a, b, idesc, (g == 0 && dx == 0) ? (si != 0 ? 1u : 0u) : 1u);
              umma_bf16_ss_lo(d_tmem, a + 2, b + 2, idesc, 1u);
              umma_bf16_ss_lo(d_tmem, a + 4, b + 4, idesc, 1u);
              umma_bf16_ss_lo(d_tmem, a + 6, b + 6, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (si == stages_per_tile - 1) umma_commit(&t.tfull()[buf]);
        }
        __syncwarp();
        // G == 1: the weights of (dy, kb+1) follow (dy, kb) by one tile and those of (dy+1, 0) follow
        // (dy, kb-1) by the remaining NDX-1 taps.  G == 3 (kb == 1): one batch per tile.
        b_row += ((si + 1) % p.kb == 0) ? b_tile + (NDX - 1) * b_tap : b_tile;
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    epilogue_loop(p.e, &tm_y, t, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}


// =================================================================================================
// Wide-N kernel: 3x3 convolutions with cout <= 64.
//
// At n_tile = 64 a 128x64x16 MMA needs 32 tensor cycles but reads 6 KB of operands from shared memory
// (48 cycles at 128 B/clk), and every horizontal tap re-reads the same A rows: the narrow layers are bound by
// shared-memory bandwidth, not by the tensor pipe.  Here the three horizontal taps become three 64-column
// groups of ONE N = 192 accumulator,
//     Z[r][dx*64 + co] = sum_dy sum_c X[r + (dy-1)*wp][c] * W[dy][dx][co][c],
// so each A row block is read once per filter row (12 MMAs of 10 KB per 128 pixels instead of 36 of 6 KB,
// tensor-bound at 96 cycles each), and the epilogue applies the horizontal shift on the OUTPUT side:
//     Y[r][co] = Z[r-1][co] + Z[r][64 + co] + Z[r+1][128 + co]
// Row r lives in TMEM lane r = thread r of the epilogue, so the shift is one __shfl_up and one __shfl_down per
// channel.  The 128 accumulator rows of a tile are FOUR SLABS of 32 raster rows, one per epilogue warp, that start 30
// rows apart: every warp owns the neighbours of its 30 inner rows itself (lanes 0 and 31 only feed the shuffles), so
// the warps never exchange rows - no shared-memory hand-off, no barrier between them, no per-lane edge selects (the
// round-1 kernel tiled 126 rows per 128 with a 2 KB exchange + bar.sync per 32 channels: 1 395 LSU wavefronts and
// ~1 000 dependent instructions per tile, which is what bounded the 64 -> 64 layers; 120 rows per tile cost 5 % more
// MMA work and four 32-row TMA boxes per filter row instead of one 128-row box).
// Weights: resident in smem when they fit (cin <= 128), else streamed with the A tiles.
// =================================================================================================
constexpr int kSlabRows = 32;          // accumulator rows (TMEM lanes) per epilogue warp
constexpr int kSlabOut = 30;           // of which are outputs
constexpr int kWideTileRows = 4 * kSlabOut;
constexpr int kXchTileRows = kBlockM - 2;          // exchange variant: one 128-row box, 126 outputs


struct WideParams {
  EpiParams e;
  int dy_shift[3];       // raster-row shift of filter row dy: (dy-1)*wp
  int kb;                // cin / 64
  int x_coff;
  int cin;
  int stages;
  int gw;                // accumulator columns per horizontal tap (64; 16..32 for the fp32 heads)
  int b_tile_bytes;      // 3*gw*128: weights of one (filter row, channel block): [3*gw rows][64 k]
  int a_slot_bytes;      // bytes per pipeline stage
  int nsets;             // epilogue sets in use (1: warps 6-9 idle, 16 KB more smem for the pipeline)
  int prefetch;          // L2-prefetch the next tile's A boxes (G == 1 variants)
};

// Horizontal tap combine of 32 accumulator columns (slab tiling: the neighbours are always in the warp): (up + centre)
// + down on fp32 pairs (FADD2), then bias + ReLU + bf16 rounding in two more instructions per pair.
template <bool RELU>
__device__ __forceinline__ void wide_combine_32(uint32_t (&pk)[32], int c, const uint32_t (&z0)[32], const uint32_t (&z1)[32],
                                                const uint32_t (&z2)[32], uint32_t bs) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 bb = lds128(bs + j * 4);
    const float bbv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int e = 0; e < 4; e += 2) {
      const float2 up = make_float2(__shfl_up_sync(0xffffffffu, __uint_as_float(z0[j + e]), 1),           // Z[r-1][co]
                                    __shfl_up_sync(0xffffffffu, __uint_as_float(z0[j + e + 1]), 1));
      const float2 dn = make_float2(__shfl_down_sync(0xffffffffu, __uint_as_float(z2[j + e]), 1),         // Z[r+1][128 + co]
                                    __shfl_down_sync(0xffffffffu, __uint_as_float(z2[j + e + 1]), 1));
      float2 a = __fadd2_rn(up, make_float2(__uint_as_float(z1[j + e]), __uint_as_float(z1[j + e + 1])));
      a = __fadd2_rn(a, dn);
      pk[c * 16 + ((j + e) >> 1)] = bias_act_pack<RELU>(a.x, a.y, bbv[e], bbv[e + 1]);
    }
  }
}

// Epilogue of the wide-N kernel, bf16 raster output, gw = 64.
__device__ __forceinline__ void epilogue_wide_bf16(const WideParams& wp_, const CUtensorMap* tm_y30, const SmemTail& t,
                                                   uint32_t tmem_base, int warp, int lane) {
  const EpiParams& p = wp_.e;
  const int q = warp & 3;
  uint8_t* stg = t.stg(warp - 2);
  const float* sbias = t.bias();
  const int set = (warp - 2) >> 2;
  const bool inner = (lane >= 1) && (lane <= kSlabOut);
  int it = set;
  for (int tile = blockIdx.x + set * gridDim.x; tile < p.m_tiles; tile += t.nsets * gridDim.x, it += t.nsets) {
    const int buf = it & 1;
    const uint32_t use = static_cast<uint32_t>(it >> 1);
    const int srow0 = tile * kWideTileRows - 1 + q * kSlabOut;   // raster row of this warp's lane 0
    const int r = srow0 + lane;
    const int rr = r < 0 ? 0 : r;
    const int img = fastdiv(rr, p.plane_m, p.plane_s);
    const int pr = rr - img * p.plane;
    const int y = fastdiv(pr, p.wp_m, p.wp_s);
    const int x = pr - y * p.wp;
    const bool valid = inner && (r < p.R) && (y >= 1) && (y <= p.hv) && (x < p.wv);
    mbar_wait(&t.tfull()[buf], use & 1u);
    tc_fence_after();
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * kAccStride);
    uint32_t pk[32];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t z0[32], z1[32], z2[32];
      tmem_ld_x32(t_base + c * 32, z0);
      tmem_ld_x32(t_base + 64 + c * 32, z1);
      tmem_ld_x32(t_base + 128 + c * 32, z2);
      tmem_ld_wait();
      if (c == 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t.tempty()[buf]);
      }
      const uint32_t bs = smem_u32(sbias + c * 32);
      if (p.relu) wide_combine_32<true>(pk, c, z0, z1, z2, bs);
      else wide_combine_32<false>(pk, c, z0, z1, z2, bs);
    }
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
    // lanes 1..30 hold the warp's 30 output rows: staged one row up so that the TMA source starts on a 1024-byte
    // boundary; 16-byte chunk j of a row goes to j ^ (row & 7)  (SWIZZLE_128B)
    if (inner) {
      const int srow = lane - 1;
      const uint32_t rowa = smem_u32(stg) + srow * 128;
      if (valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128u(rowa + ((j ^ (srow & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      } else {                                        // border / masked rows: zeros (a branch per row, not a select per word)
#pragma unroll
        for (int j = 0; j < 8; ++j) sts128u(rowa + (j << 4), 0u, 0u, 0u, 0u);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tm_y30, stg, p.y_coff, srow0 + 1);       // rows beyond the raster are clipped by TMA
      tma_store_commit();
    }
  }
  if (lane == 0) tma_store_wait<0>();
}


// Epilogue of the wide-N kernel for the small fp32 heads (cout <= 16, gw = 16: the `coef` conv + softmax,
// model_library.py:405-406).  Same horizontal combine as above on 16-column groups, then the single-chunk
// fp32 path of epilogue_loop: softmax in registers, rows compacted through the staging buffer, coalesced stores.
__device__ __forceinline__ void epilogue_wide_f32(const WideParams& wp_, const SmemTail& t, uint32_t tmem_base,
                                                  int warp, int lane) {
  const EpiParams& p = wp_.e;
  const int q = warp & 3;
  float* sf = reinterpret_cast<float*>(t.stg(warp - 2));
  const float* sbias = t.bias();
  const int set = (warp - 2) >> 2;
  const bool sm_mode = (p.epilogue == IE_EPI_F32_SOFTMAX);
  const bool inner = (lane >= 1) && (lane <= kSlabOut);
  int it = set;
  for (int tile = blockIdx.x + set * gridDim.x; tile < p.m_tiles; tile += t.nsets * gridDim.x, it += t.nsets) {
    const int buf = it & 1;
    const uint32_t use = static_cast<uint32_t>(it >> 1);
    const int srow0 = tile * kWideTileRows - 1 + q * kSlabOut;
    const int r = srow0 + lane;
    const int rr = r < 0 ? 0 : r;
    const int img = fastdiv(rr, p.plane_m, p.plane_s);
    const int pr = rr - img * p.plane;
    const int y = fastdiv(pr, p.wp_m, p.wp_s);
    const int x = pr - y * p.wp;
    const bool valid = inner && (r < p.R) && (y >= 1) && (y <= p.hv) && (x < p.wv);
    const long long pix = (static_cast<long long>(img) * p.hv + (y - 1)) * p.wv + x;
    mbar_wait(&t.tfull()[buf], use & 1u);
    tc_fence_after();
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * kAccStride);
    uint32_t z0[16], z1[16], z2[16];
    tmem_ld_x16(t_base, z0);
    tmem_ld_x16(t_base + 16, z1);
    tmem_ld_x16(t_base + 32, z2);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&t.tempty()[buf]);
    const uint32_t bs = smem_u32(sbias);
    float a[16];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 bb = lds128(bs + j * 4);
      const float bbv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(z0[j + e]), 1);
        const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(z2[j + e]), 1);
        float v = up + __uint_as_float(z1[j + e]) + dn + bbv[e];
        if (p.relu) v = fmaxf(v, 0.f);
        a[j + e] = v;
        if (j + e < p.cout) mx = fmaxf(mx, v);
      }
    }
    float e_[16];
    if (sm_mode) {
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        e_[j] = (j < p.cout) ? __expf(a[j] - mx) : 0.f;
        sum += e_[j];
      }
      const float inv = 1.f / sum;
#pragma unroll
      for (int j = 0; j < 16; ++j) e_[j] *= inv;
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) e_[j] = a[j];
    }
    // valid pixels of the warp's rows are consecutive in the NHWC output (border pixels have no slot): compact them
    // through the staging buffer and store fully coalesced
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (bal) {
      const int rank = __popc(bal & ((1u << lane) - 1u));
      const int nvalid = __popc(bal);
      const long long pix_first = __shfl_sync(0xffffffffu, pix, __ffs(bal) - 1);
      const int total = nvalid * p.cout;
      for (int pass = 0; pass < ((sm_mode && p.y_aux) ? 2 : 1); ++pass) {
        __syncwarp();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < p.cout) sf[rank * p.cout + j] = pass ? a[j] : e_[j];
        }
        __syncwarp();
        float* dst = (pass ? p.y_aux : p.y_f32) + pix_first * p.cout;
        for (int i = lane; i < total; i += 32) dst[i] = sf[i];
      }
    }
  }
}

// Epilogue of the wide-N kernel, bf16 raster output, gw = 64.
// EXCHANGE variant (layers with more than one channel block, whose long main loop hides it): ONE 128-row box per filter
// row, 126 outputs per tile; the rows on warp boundaries travel through a 2 KB shared-memory exchange.
__device__ __forceinline__ void epilogue_wide_bf16_xch(const WideParams& wp_, const CUtensorMap* tm_y32,
                                                   const CUtensorMap* tm_y31, const SmemTail& t, uint32_t tmem_base,
                                                   int warp, int lane) {
  const EpiParams& p = wp_.e;
  const int q = warp & 3;
  const int m = q * 32 + lane;                 // row inside the 128-row tile = TMEM lane
  uint8_t* stg = t.stg(warp - 2);
  const float* sbias = t.bias();
  const int set = (warp - 2) >> 2;
  float* xch = t.xch() + set * (kXchBytes / 4);
  int it = set;
  for (int tile = blockIdx.x + set * gridDim.x; tile < p.m_tiles; tile += t.nsets * gridDim.x, it += t.nsets) {
    const int buf = it & 1;
    const uint32_t use = static_cast<uint32_t>(it >> 1);
    const int row0 = tile * kXchTileRows - 1;  // raster row of tile-local row 0
    const int r = row0 + m;
    const int rr = r < 0 ? 0 : r;
    const int img = fastdiv(rr, p.plane_m, p.plane_s);
    const int pr = rr - img * p.plane;
    const int y = fastdiv(pr, p.wp_m, p.wp_s);
    const int x = pr - y * p.wp;
    const bool valid = (m >= 1) && (m <= kXchTileRows) && (r >= 0) && (r < p.R) && (y >= 1) && (y <= p.hv) &&
                       (x < p.wv);
    mbar_wait(&t.tfull()[buf], use & 1u);
    tc_fence_after();
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * kAccStride);
    uint32_t pk[32];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t z0[32], z1[32], z2[32];
      tmem_ld_x32(t_base + c * 32, z0);
      tmem_ld_x32(t_base + 64 + c * 32, z1);
      tmem_ld_x32(t_base + 128 + c * 32, z2);
      tmem_ld_wait();
      if (c == 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&t.tempty()[buf]);
        }
      }
      // rows on warp boundaries travel through shared memory
      const uint32_t mine = smem_u32(xch + ((c * 4 + q) * 2) * 32);
      if (lane == 31) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) sts128u(mine + j * 4, z0[j], z0[j + 1], z0[j + 2], z0[j + 3]);
      }
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) sts128u(mine + 128 + j * 4, z2[j], z2[j + 1], z2[j + 2], z2[j + 3]);
      }
      epi_bar_sync(set);
      // every lane loads the two edge rows (broadcast, no divergence) and lanes 0 / 31 select them
      const uint32_t prev = smem_u32(xch + ((c * 4 + ((q + 3) & 3)) * 2) * 32);       // lane 31 of the warp below (row m-1)
      const uint32_t next = smem_u32(xch + ((c * 4 + ((q + 1) & 3)) * 2 + 1) * 32);   // lane 0 of the warp above (row m+1)
      const uint32_t bs = smem_u32(sbias + c * 32);
      const bool first = (lane == 0), last = (lane == 31);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 pv = lds128(prev + j * 4), nx = lds128(next + j * 4), bb = lds128(bs + j * 4);
        const float pvv[4] = {pv.x, pv.y, pv.z, pv.w}, nxv[4] = {nx.x, nx.y, nx.z, nx.w}, bbv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          float2 up = make_float2(__shfl_up_sync(0xffffffffu, __uint_as_float(z0[j + e]), 1),
                                  __shfl_up_sync(0xffffffffu, __uint_as_float(z0[j + e + 1]), 1));
          float2 dn = make_float2(__shfl_down_sync(0xffffffffu, __uint_as_float(z2[j + e]), 1),
                                  __shfl_down_sync(0xffffffffu, __uint_as_float(z2[j + e + 1]), 1));
          up = first ? make_float2(pvv[e], pvv[e + 1]) : up;               // q == 0: row m = 0 is never stored
          dn = last ? make_float2(nxv[e], nxv[e + 1]) : dn;                // q == 3: row m = 127 is never stored
          float2 a = __fadd2_rn(up, make_float2(__uint_as_float(z1[j + e]), __uint_as_float(z1[j + e + 1])));
          a = __fadd2_rn(a, dn);
          // (one packed add + the plain conversion; ReLU as a packed max on the two bf16 halves would need HMNMX2 -
          //  the multi-block layers hide their epilogue anyway)
          const float2 sb = __fadd2_rn(a, make_float2(bbv[e], bbv[e + 1]));
          const float r0 = p.relu ? fmaxf(sb.x, 0.f) : sb.x, r1 = p.relu ? fmaxf(sb.y, 0.f) : sb.y;
          pk[c * 16 + ((j + e) >> 1)] = valid ? pack_bf16x2(r0, r1) : 0u;
        }
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
    // rows m = 0 and m = 127 belong to the neighbouring tiles: warps 0 and 3 store 31 rows; warp 0 shifts its
    // rows up by one so that every TMA source starts on a 1024-byte boundary
    const int srow = (q == 0) ? lane - 1 : lane;
    if (srow >= 0) {
      const uint32_t rowa = smem_u32(stg) + srow * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128u(rowa + ((j ^ (srow & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (q == 0) tma_store_2d(tm_y31, stg, p.y_coff, row0 + 1);
      else if (q == 3) tma_store_2d(tm_y31, stg, p.y_coff, row0 + 96);
      else tma_store_2d(tm_y32, stg, p.y_coff, row0 + q * 32);
      tma_store_commit();
    }
  }
  if (lane == 0) tma_store_wait<0>();
}


// EXCHANGE variant of the fp32-head epilogue (cout <= 16, more than one channel block).  Same horizontal combine as above on 16-column groups, then the single-chunk
// fp32 path of epilogue_loop: softmax in registers, rows compacted through the staging buffer, coalesced stores.
__device__ __forceinline__ void epilogue_wide_f32_xch(const WideParams& wp_, const SmemTail& t, uint32_t tmem_base,
                                                  int warp, int lane) {
  const EpiParams& p = wp_.e;
  const int q = warp & 3;
  const int m = q * 32 + lane;
  float* sf = reinterpret_cast<float*>(t.stg(warp - 2));
  const float* sbias = t.bias();
  const int set = (warp - 2) >> 2;
  float* xch = t.xch() + set * (kXchBytes / 4);
  const bool sm_mode = (p.epilogue == IE_EPI_F32_SOFTMAX);
  int it = set, local = 0;
  for (int tile = blockIdx.x + set * gridDim.x; tile < p.m_tiles; tile += t.nsets * gridDim.x, it += t.nsets, ++local) {
    const int buf = it & 1;
    const uint32_t use = static_cast<uint32_t>(it >> 1);
    const int row0 = tile * kXchTileRows - 1;
    const int r = row0 + m;
    const int rr = r < 0 ? 0 : r;
    const int img = fastdiv(rr, p.plane_m, p.plane_s);
    const int pr = rr - img * p.plane;
    const int y = fastdiv(pr, p.wp_m, p.wp_s);
    const int x = pr - y * p.wp;
    const bool valid = (m >= 1) && (m <= kXchTileRows) && (r >= 0) && (r < p.R) && (y >= 1) && (y <= p.hv) && (x < p.wv);
    const long long pix = (static_cast<long long>(img) * p.hv + (y - 1)) * p.wv + x;
    mbar_wait(&t.tfull()[buf], use & 1u);
    tc_fence_after();
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * kAccStride);
    uint32_t z0[16], z1[16], z2[16];
    tmem_ld_x16(t_base, z0);
    tmem_ld_x16(t_base + 16, z1);
    tmem_ld_x16(t_base + 32, z2);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&t.tempty()[buf]);
    const int par = local & 1;                                // exchange buffers alternate between this set's tiles
    const uint32_t mine = smem_u32(xch + ((par * 4 + q) * 2) * 32);
    if (lane == 31) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) sts128u(mine + j * 4, z0[j], z0[j + 1], z0[j + 2], z0[j + 3]);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) sts128u(mine + 128 + j * 4, z2[j], z2[j + 1], z2[j + 2], z2[j + 3]);
    }
    epi_bar_sync(set);
    const uint32_t prev = smem_u32(xch + ((par * 4 + ((q + 3) & 3)) * 2) * 32);
    const uint32_t next = smem_u32(xch + ((par * 4 + ((q + 1) & 3)) * 2 + 1) * 32);
    const uint32_t bs = smem_u32(sbias);
    const bool first = (lane == 0), last = (lane == 31);
    float a[16];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 pv = lds128(prev + j * 4), nx = lds128(next + j * 4), bb = lds128(bs + j * 4);
      const float pvv[4] = {pv.x, pv.y, pv.z, pv.w}, nxv[4] = {nx.x, nx.y, nx.z, nx.w}, bbv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float up = __shfl_up_sync(0xffffffffu, __uint_as_float(z0[j + e]), 1);
        float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(z2[j + e]), 1);
        up = first ? pvv[e] : up;
        dn = last ? nxv[e] : dn;
        float v = up + __uint_as_float(z1[j + e]) + dn + bbv[e];
        if (p.relu) v = fmaxf(v, 0.f);
        a[j + e] = v;
        if (j + e < p.cout) mx = fmaxf(mx, v);
      }
    }
    float e_[16];
    if (sm_mode) {
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        e_[j] = (j < p.cout) ? __expf(a[j] - mx) : 0.f;
        sum += e_[j];
      }
      const float inv = 1.f / sum;
#pragma unroll
      for (int j = 0; j < 16; ++j) e_[j] *= inv;
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) e_[j] = a[j];
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (bal) {
      const int rank = __popc(bal & ((1u << lane) - 1u));
      const int nvalid = __popc(bal);
      const long long pix_first = __shfl_sync(0xffffffffu, pix, __ffs(bal) - 1);
      const int total = nvalid * p.cout;
      for (int pass = 0; pass < ((sm_mode && p.y_aux) ? 2 : 1); ++pass) {
        __syncwarp();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < p.cout) sf[rank * p.cout + j] = pass ? a[j] : e_[j];
        }
        __syncwarp();
        float* dst = (pass ? p.y_aux : p.y_f32) + pix_first * p.cout;
        for (int i = lane; i < total; i += 32) dst[i] = sf[i];
      }
    }
  }
}

// RES: weights resident in smem.  G: filter rows per pipeline stage (3 needs RES and kb == 1).
// SLAB: tm_a is a box of kSlabRows rows (one slab of one filter row), 120 outputs per tile, tm_y1 = 30-row store box
// (tm_y2 unused); else tm_a is ONE 128-row box, 126 outputs per tile, tm_y1 / tm_y2 = 32- / 31-row store boxes.
template <bool RES, int G, bool SLAB>
__global__ void __launch_bounds__(kThreads2, 1)
conv_wide_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_y1, const __grid_constant__ CUtensorMap tm_y2, const WideParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024(smem_raw);
  const int b_res_bytes = RES ? 3 * p.kb * p.b_tile_bytes : 0;
  uint8_t* a_base_ptr = base + b_res_bytes;
  SmemTail t{a_base_ptr + p.stages * p.a_slot_bytes, p.nsets};
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.e.m_tiles;
  const int stages_per_tile = 3 * p.kb / G;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_y1);
    tma_prefetch_desc(&tm_y2);
  }
  const uint32_t tmem_base = cta_setup(t, p.e, p.stages, warp, lane);
  uint64_t* full_bar = t.full();
  uint64_t* empty_bar = t.empty();
  const int piece = p.gw * 128;                    // one horizontal tap's weights: [gw rows][64 k]
  constexpr int kSlabBytes = kSlabRows * 128;
  constexpr int kBoxes = SLAB ? 4 : 1;             // TMA boxes per filter row and channel block
  constexpr int kBoxStep = SLAB ? kSlabOut : 0;
  constexpr int kTileRows = SLAB ? kWideTileRows : kXchTileRows;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      if constexpr (RES) {
        mbar_arrive_expect_tx(t.bres(), static_cast<uint32_t>(b_res_bytes));
        for (int dy = 0; dy < 3; ++dy)
          for (int kb = 0; kb < p.kb; ++kb)
            for (int dx = 0; dx < 3; ++dx)
              tma_load_2d(base + (dy * p.kb + kb) * p.b_tile_bytes + dx * piece, &tm_b, t.bres(),
                          (dy * 3 + dx) * p.cin + kb * kBlockK, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * kTileRows - 1;                // raster row of accumulator row 0
        if constexpr (G == 3) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], 3u * kABytes);
#pragma unroll
          for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int sl = 0; sl < kBoxes; ++sl)
              tma_load_2d(a_base_ptr + stage * p.a_slot_bytes + g * kABytes + sl * kSlabBytes, &tm_a, &full_bar[stage],
                          p.x_coff, row0 + sl * kBoxStep + p.dy_shift[g]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        } else {
          // layers with several channel blocks keep only 3 - 5 A stages next to their weights and are bound by the
          // latency of the A boxes (HBM under load: ~2 us for 48 - 80 KB in flight per SM): pull the NEXT tile's
          // boxes into L2 while this tile's are consumed, so that the loads that fill the stages hit L2
          const int next_tile = tile + gridDim.x;
          const bool pf = p.prefetch && next_tile < num_tiles;
          const int next_row0 = next_tile * kTileRows - 1;
          for (int dy = 0; dy < 3; ++dy) {
            for (int kb = 0; kb < p.kb; ++kb) {
              if (pf) {
#pragma unroll
                for (int sl = 0; sl < kBoxes; ++sl)
                  tma_prefetch_2d(&tm_a, p.x_coff + kb * kBlockK, next_row0 + sl * kBoxStep + p.dy_shift[dy]);
              }
              mbar_wait(&empty_bar[stage], phase ^ 1u);
              mbar_arrive_expect_tx(&full_bar[stage], kABytes + (RES ? 0u : static_cast<uint32_t>(p.b_tile_bytes)));
              uint8_t* a_dst = a_base_ptr + stage * p.a_slot_bytes;
#pragma unroll
              for (int sl = 0; sl < kBoxes; ++sl)
                tma_load_2d(a_dst + sl * kSlabBytes, &tm_a, &full_bar[stage], p.x_coff + kb * kBlockK,
                            row0 + sl * kBoxStep + p.dy_shift[dy]);
              if constexpr (!RES) {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
                  tma_load_2d(a_dst + kABytes + dx * piece, &tm_b, &full_bar[stage],
                              (dy * 3 + dx) * p.cin + kb * kBlockK, 0);
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    const uint32_t idesc = umma_idesc_bf16(kBlockM, 3 * p.gw);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(a_base_ptr));
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(base));
    const uint32_t a_stride = static_cast<uint32_t>(p.a_slot_bytes) >> 4;
    const uint32_t b_tile = static_cast<uint32_t>(p.b_tile_bytes) >> 4;
    if constexpr (RES) {
      mbar_wait(t.bres(), 0);
      tc_fence_after();
    }
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      mbar_wait(&t.tempty()[buf], (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kAccStride);
      for (int si = 0; si < stages_per_tile; ++si) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_slot_lo = a_lo0 + stage * a_stride;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const uint32_t a = a_slot_lo + g * (kABytes >> 4);
            const uint32_t b = RES ? b_lo0 + (si * G + g) * b_tile : a_slot_lo + (kABytes >> 4);
            umma_bf16_ss_lo(d_tmem, a, b, idesc, (g == 0) ? (si != 0 ? 1u : 0u) : 1u);
            umma_bf16_ss_lo(d_tmem, a + 2, b + 2, idesc, 1u);
            umma_bf16_ss_lo(d_tmem, a + 4, b + 4, idesc, 1u);
            umma_bf16_ss_lo(d_tmem, a + 6, b + 6, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
          if (si == stages_per_tile - 1) umma_commit(&t.tfull()[buf]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 2 + 4 * p.nsets) {
    // second epilogue set not in use
  } else if (p.e.epilogue == IE_EPI_BF16_RASTER) {
    if constexpr (SLAB) epilogue_wide_bf16(p, &tm_y1, t, tmem_base, warp, lane);
    else epilogue_wide_bf16_xch(p, &tm_y1, &tm_y2, t, tmem_base, warp, lane);
  } else {
    if constexpr (SLAB) epilogue_wide_f32(p, t, tmem_base, warp, lane);
    else epilogue_wide_f32_xch(p, t, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}


// =================================================================================================
// First layer: Conv2D(64, 3, 'same', relu) on the raw fp32 NHWC input (model_library.py:323/376, 196/235)
// with the im2col fused into the kernel.  K = 9*C <= 64 is one 64-wide K block, so a tile is ONE stage of four
// MMAs; there is no bf16 im2col raster in HBM (368 MB written + read back at 256 x 104^2).  Four extra
// "builder" warps (one thread per tile row) gather each pixel's 3x3xC neighbourhood - three runs of 3*C
// contiguous floats - straight from the input, convert to bf16 and write the 128-byte swizzled A row that TMA
// would have written; generic-proxy writes are made visible to the tensor core with fence.proxy.async.
// The source may be smaller than the raster (implicit zero padding to the network stride, see im2col).
// =================================================================================================
struct FirstParams {
  EpiParams e;
  const float* x;
  int hs, ws;            // source size
  int stages;
  long long total_floats;      // n * hs * ws * C (staged flavour: copy ranges are clamped to the tensor)
};

constexpr int kFirstThreads = 448;   // gather flavour: warp 0 weights, 1 MMA, 2-5 epilogue, 6-13 two builder sets

template <int C>
__global__ void __launch_bounds__(kFirstThreads, 1)
conv_first_kernel(const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_y, const FirstParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024(smem_raw);
  constexpr int KB = (9 * C + 63) / 64;                   // 64-wide K blocks: 1 for C <= 7, 2 for C = 10 (Basis_kpn, T = 8)
  constexpr int kWTile = 64 * 128;                        // [64 cout][64 k] bf16 per K block
  constexpr int kWBytes = KB * kWTile;
  constexpr int kStageBytes = KB * kABytes;
  uint8_t* a_base_ptr = base + kWBytes;
  SmemTail t{a_base_ptr + p.stages * kStageBytes};
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.e.m_tiles;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_y);
  }
  const uint32_t tmem_base = cta_setup(t, p.e, p.stages, warp, lane, 4);
  uint64_t* full_bar = t.full();
  uint64_t* empty_bar = t.empty();

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(t.bres(), kWBytes);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(base + kb * kWTile, &tm_b, t.bres(), kb * kBlockK, 0);
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    const uint32_t idesc = umma_idesc_bf16(kBlockM, 64);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(a_base_ptr));
    const uint32_t b_lo = umma_desc_lo(smem_u32(base));
    mbar_wait(t.bres(), 0);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      mbar_wait(&t.tempty()[buf], (use & 1u) ^ 1u);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kAccStride);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t a = a_lo0 + stage * (kStageBytes >> 4) + kb * (kABytes >> 4);
          const uint32_t b = b_lo + kb * (kWTile >> 4);
          umma_bf16_ss_lo(d_tmem, a, b, idesc, kb == 0 ? 0u : 1u);
          umma_bf16_ss_lo(d_tmem, a + 2, b + 2, idesc, 1u);
          umma_bf16_ss_lo(d_tmem, a + 4, b + 4, idesc, 1u);
          umma_bf16_ss_lo(d_tmem, a + 6, b + 6, idesc, 1u);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&t.tfull()[buf]);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp < 6) {
    epilogue_loop(p.e, &tm_y, t, tmem_base, warp, lane);
  } else {
    // ================================ A-tile builders ==============================
    // two sets of four warps; set s builds the CTA's tiles it = s, s+2, ... into stage it % stages (a single warp
    // per scheduler cannot hide the latency of its gathers, exactly like the epilogue)
    const int bset = (warp - 6) >> 2;
    const int row_local = ((warp - 6) & 3) * 32 + lane;
    constexpr int C3 = 3 * C;
    int it = bset;
    for (int tile = blockIdx.x + bset * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, it += 2) {
      const int stage = it % p.stages;
      const uint32_t phase = static_cast<uint32_t>(it / p.stages) & 1u;
      const int r = tile * kBlockM + row_local;
      const int img = fastdiv(r, p.e.plane_m, p.e.plane_s);
      const int pr = r - img * p.e.plane;
      const int yp = fastdiv(pr, p.e.wp_m, p.e.wp_s);
      const int y = yp - 1, x = pr - yp * p.e.wp;
      const bool interior = (r < p.e.R) && (y >= 0) && (y < p.e.hv) && (x >= 0) && (x < p.e.wv);
      const float* centre = p.x + ((static_cast<long long>(img) * p.hs + y) * p.ws + (x - 1)) * C;   // (y, x-1, 0)
      bool rowok[3], colok[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        rowok[d] = interior && (y + d - 1 >= 0) && (y + d - 1 < p.hs);
        colok[d] = (x + d - 1 >= 0) && (x + d - 1 < p.ws);
      }
      uint32_t pk[KB][32];
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        float v[64];
#pragma unroll
        for (int kk = 0; kk < 64; ++kk) {
          const int k = kb * 64 + kk;
          v[kk] = 0.f;
          if (k < 9 * C) {
            const int dy = k / C3, i = k % C3, g = i / C;
            if (rowok[dy] && colok[g]) v[kk] = __ldg(centre + (dy - 1) * p.ws * C + i);
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) pk[kb][j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      }
      mbar_wait(&empty_bar[stage], phase ^ 1u);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint32_t rowa = smem_u32(a_base_ptr + stage * kStageBytes + kb * kABytes) + row_local * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128u(rowa + ((j ^ (row_local & 7)) << 4), pk[kb][4 * j], pk[kb][4 * j + 1], pk[kb][4 * j + 2], pk[kb][4 * j + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// Staged flavour of the first layer (the default): the gathers of the kernel above are 45 scalar global loads per
// pixel whose latency eight builder warps cannot hide (ncu: 183 us at 256 x 104^2, DRAM 24 %, no pipe above 33 %).
// Here warp 0 copies, per tile and filter row, the contiguous run of source pixels the tile's 128 raster positions
// touch (<= 130 pixels: in raster order the clamped source index (img, clamp(y + dy - 1), min(x, ws - 1)) moves by at
// most one per position) with ONE cp.async.bulk each, several tiles ahead, into a small ring; the builders read
// their 3 x 3C floats from shared memory (lane stride C floats: conflict-free for odd C) and write the swizzled A row
// as before.  The copies start on the enclosing 16-byte boundary; the float index of each slot's first element is
// passed through the ring's metadata words.  Requires a 16-byte aligned source whose float count is a multiple of 4.
template <int C>
struct FirstStaged {
  // ring depth = tiles whose source rows are in flight: the copies come from DRAM (1 - 2 us under the kernel's own
  // write traffic) and a tile takes ~0.5 us, so four stages left the builders waiting (measured: 148 us at 256 x 104^2)
  static constexpr int kInStages = C <= 5 ? 12 : 6;
  static constexpr int kAStages = C <= 7 ? 3 : 2;                            // built A tiles waiting for the MMA warp
  static constexpr int kSlot = ((130 * C * 4 + 32 + 127) / 128) * 128;     // one filter row of one tile
  static constexpr int kStage = 3 * kSlot + 128;                             // + metadata (3 ints), 128-byte aligned
  // + 2 * kInStages barriers + a 128-byte zero page (what masked taps read); a multiple of 1024 because the
  // epilogue's staging buffers behind it are 128B-swizzled
  static constexpr int kBarBytes = ((2 * kInStages * 8 + 127) / 128) * 128;
  static constexpr int kRing = ((kInStages * kStage + kBarBytes + 128 + 1023) / 1024) * 1024;
};

// im2col element k = tap * C + ch of a pixel: one ld.shared with the channel offset as an immediate
template <int C, int K>
__device__ __forceinline__ float load_tap(const uint32_t (&tapaddr)[9]) {
  if constexpr (K < 9 * C) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1 + %2];" : "=f"(v) : "r"(tapaddr[K / C]), "n"((K % C) * 4));
    return v;
  } else {
    return 0.f;
  }
}
template <int C, int KBASE, int... KK>
__device__ __forceinline__ void load_taps(float (&v)[64], const uint32_t (&tapaddr)[9], std::integer_sequence<int, KK...>) {
  ((v[KK] = load_tap<C, KBASE + KK>(tapaddr)), ...);
}

constexpr int kFirstStagedWarps = 19;      // warp 0 + warp 18: input stagers, 1: MMA, 2-9: two epilogue sets, 10-17: two builder sets
template <int C>
__global__ void __launch_bounds__(kFirstStagedWarps * 32, 1)
conv_first_staged_kernel(const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_y,
                         const FirstParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024(smem_raw);
  using L = FirstStaged<C>;
  constexpr int kInStages = L::kInStages;
  constexpr int KB = (9 * C + 63) / 64;
  constexpr int kWTile = 64 * 128;
  constexpr int kWBytes = KB * kWTile;
  constexpr int kStageBytes = KB * kABytes;
  uint8_t* a_base_ptr = base + kWBytes;
  uint8_t* ring = a_base_ptr + p.stages * kStageBytes;
  uint64_t* in_full = reinterpret_cast<uint64_t*>(ring + kInStages * L::kStage);
  uint64_t* in_empty = in_full + kInStages;
  uint8_t* zero_page = ring + kInStages * L::kStage + L::kBarBytes;
  SmemTail t{ring + L::kRing, 2};
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.e.m_tiles;
  if (warp == 2) reinterpret_cast<uint32_t*>(zero_page)[lane] = 0u;        // visible after cta_setup's __syncthreads
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_y);
    for (int s = 0; s < kInStages; ++s) {
      mbar_init(&in_full[s], 3);        // one arrive.expect_tx per filter row (lanes 0-2 of the stager)
      mbar_init(&in_empty[s], 4);       // one arrive per builder warp of the set that consumed the stage
    }
  }
  const uint32_t tmem_base = cta_setup(t, p.e, p.stages, warp, lane, 4);
  uint64_t* full_bar = t.full();
  uint64_t* empty_bar = t.empty();

  if (warp == 0 || warp == kFirstStagedWarps - 1) {
    if (warp == 0 && lane == 0) {
      mbar_arrive_expect_tx(t.bres(), kWBytes);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(base + kb * kWTile, &tm_b, t.bres(), kb * kBlockK, 0);
    }
    // ================================ input stagers ===============================
    // two warps, alternate tiles (one warp's per-tile latency - decode four positions per lane, six warp
    // reductions, plan, issue - was the kernel's critical path)
    const int pset = warp == 0 ? 0 : 1;
    int it = pset;
    for (int tile = blockIdx.x + pset * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, it += 2) {
      const int s = it % kInStages;
      const uint32_t phase = static_cast<uint32_t>(it / kInStages) & 1u;
      int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = tile * kBlockM + lane + 32 * j;
        if (r < p.e.R) {
          const int img = fastdiv(r, p.e.plane_m, p.e.plane_s);
          const int pr = r - img * p.e.plane;
          const int yp = fastdiv(pr, p.e.wp_m, p.e.wp_s);
          const int y = yp - 1, x = pr - yp * p.e.wp;
          const int cx = x < p.ws ? x : p.ws - 1;
          const bool interior = (y >= 0) && (y < p.e.hv) && (x < p.e.wv);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            // only positions that read filter row dy count: along those the source index moves by at most one per
            // raster position, so a tile's run is <= 128 pixels (+ one on either side for the horizontal taps)
            const int sy = y + dy - 1;
            if (interior && sy >= 0 && sy < p.hs) {
              const int v = (img * p.hs + sy) * p.ws + cx;
              lo[dy] = v < lo[dy] ? v : lo[dy];
              hi[dy] = v > hi[dy] ? v : hi[dy];
            }
          }
        }
      }
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        lo[dy] = __reduce_min_sync(0xffffffffu, lo[dy]);
        hi[dy] = __reduce_max_sync(0xffffffffu, hi[dy]);
      }
      mbar_wait(&in_empty[s], phase ^ 1u);
      if (lane < 3) {
        // lane dy plans and issues the copy of filter row dy (three short dependent chains side by side instead of
        // one long one: this warp's per-tile latency bounds the kernel)
        uint8_t* stage = ring + s * L::kStage;
        int* meta = reinterpret_cast<int*>(stage + 3 * L::kSlot);
        const int mylo = lane == 0 ? lo[0] : lane == 1 ? lo[1] : lo[2];
        const int myhi = lane == 0 ? hi[0] : lane == 1 ? hi[1] : hi[2];
        long long f0 = (static_cast<long long>(mylo) - 1) * C;
        long long f1 = (static_cast<long long>(myhi) + 2) * C;            // exclusive
        f0 = f0 < 0 ? 0 : f0;
        f1 = f1 > p.total_floats ? p.total_floats : f1;
        const long long b0 = (f0 * 4) & ~15ll;
        const uint32_t bytes = myhi < mylo ? 0u : static_cast<uint32_t>(((f1 * 4 + 15) & ~15ll) - b0);   // no reader: no copy
        meta[lane] = static_cast<int>(b0 >> 2);
        mbar_arrive_expect_tx(&in_full[s], bytes);
        if (bytes)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(stage + lane * L::kSlot)),
                       "l"(reinterpret_cast<const uint8_t*>(p.x) + b0), "r"(bytes), "r"(smem_u32(&in_full[s]))
                       : "memory");
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    const uint32_t idesc = umma_idesc_bf16(kBlockM, 64);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(a_base_ptr));
    const uint32_t b_lo = umma_desc_lo(smem_u32(base));
    mbar_wait(t.bres(), 0);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      mbar_wait(&t.tempty()[buf], (use & 1u) ^ 1u);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kAccStride);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t a = a_lo0 + stage * (kStageBytes >> 4) + kb * (kABytes >> 4);
          const uint32_t b = b_lo + kb * (kWTile >> 4);
          umma_bf16_ss_lo(d_tmem, a, b, idesc, kb == 0 ? 0u : 1u);
          umma_bf16_ss_lo(d_tmem, a + 2, b + 2, idesc, 1u);
          umma_bf16_ss_lo(d_tmem, a + 4, b + 4, idesc, 1u);
          umma_bf16_ss_lo(d_tmem, a + 6, b + 6, idesc, 1u);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&t.tfull()[buf]);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp < 10) {
    // two epilogue sets: with the gathers out of the way the tile rate is set by the epilogue's dependent
    // TMEM load -> pack -> stage -> TMA store stream (~2 300 cycles per tile and set)
    epilogue_loop(p.e, &tm_y, t, tmem_base, warp, lane);
  } else {
    // ================================ A-tile builders ==============================
    // two sets of four warps, alternate tiles
    const int bset = (warp - 10) >> 2;
    const int row_local = ((warp - 10) & 3) * 32 + lane;
    int it = bset;
    for (int tile = blockIdx.x + bset * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, it += 2) {
      const int stage = it % p.stages;
      const uint32_t phase = static_cast<uint32_t>(it / p.stages) & 1u;
      const int sin = it % kInStages;
      const uint32_t phase_in = static_cast<uint32_t>(it / kInStages) & 1u;
      const int r = tile * kBlockM + row_local;
      const int img = fastdiv(r, p.e.plane_m, p.e.plane_s);
      const int pr = r - img * p.e.plane;
      const int yp = fastdiv(pr, p.e.wp_m, p.e.wp_s);
      const int y = yp - 1, x = pr - yp * p.e.wp;
      const bool interior = (r < p.e.R) && (y >= 0) && (y < p.e.hv) && (x >= 0) && (x < p.e.wv);
      bool rowok[3], colok[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        rowok[d] = interior && (y + d - 1 >= 0) && (y + d - 1 < p.hs);
        colok[d] = (x + d - 1 >= 0) && (x + d - 1 < p.ws);
      }
      const uint8_t* stage_in = ring + sin * L::kStage;
      const int* meta = reinterpret_cast<const int*>(stage_in + 3 * L::kSlot);
      mbar_wait(&in_full[sin], phase_in);
      // one base address per (filter row, horizontal tap): the pixel's C floats in the staged row, or the zero page
      // when the tap is outside the source - 9 selects per pixel instead of a compare + select per element, and the
      // channel offset is an immediate of the load
      uint32_t tapaddr[9];
      const uint32_t zaddr = smem_u32(zero_page);
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        // float index of (img, y + dy - 1, x - 1, 0) relative to the slot's first float
        const int off = ((img * p.hs + (y + dy - 1)) * p.ws + (x - 1)) * C - meta[dy];
        const uint32_t rowaddr = smem_u32(stage_in + dy * L::kSlot) + static_cast<uint32_t>(off) * 4u;
#pragma unroll
        for (int g = 0; g < 3; ++g) tapaddr[dy * 3 + g] = (rowok[dy] && colok[g]) ? rowaddr + g * C * 4 : zaddr;
      }
      uint32_t pk[KB][32];
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        float v[64];
#pragma unroll
        for (int kk = 0; kk < 64; ++kk) v[kk] = 0.f;
        if (kb == 0) load_taps<C, 0>(v, tapaddr, std::make_integer_sequence<int, 64>{});
        else load_taps<C, 64>(v, tapaddr, std::make_integer_sequence<int, 64>{});
#pragma unroll
        for (int j = 0; j < 32; ++j) pk[kb][j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&in_empty[sin]);          // the values are in registers: the slot may be refilled
      mbar_wait(&empty_bar[stage], phase ^ 1u);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint32_t rowa = smem_u32(a_base_ptr + stage * kStageBytes + kb * kABytes) + row_local * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128u(rowa + ((j ^ (row_local & 7)) << 4), pk[kb][4 * j], pk[kb][4 * j + 1], pk[kb][4 * j + 2], pk[kb][4 * j + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// Second half of a split-K convolution: for every split tile, sums the `ksplit` fp32 slabs in a fixed order, adds the
// bias, applies ReLU and the border / valid-extent mask, and writes the bf16 rows (8 channels per thread).
__global__ void __launch_bounds__(256)
splitk_finish_kernel(EpiParams p, uint4* __restrict__ y, int y_pitch_v) {
  pdl_trigger();
  pdl_wait();
  const int vec_per_row = p.n_tile >> 3;
  const int per_tile = kBlockM * vec_per_row;
  const int tiles = p.m_tiles * p.n_tiles - p.split_first;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(tiles) * per_tile) return;
  const int lt = static_cast<int>(idx / per_tile);                 // split tile number
  const int rem = static_cast<int>(idx - static_cast<long long>(lt) * per_tile);
  const int row = rem / vec_per_row, cv = rem - row * vec_per_row;
  const int tile = p.split_first + lt;
  const int m_tile = tile / p.n_tiles, n_idx = tile - m_tile * p.n_tiles;
  const int r = m_tile * kBlockM + row;
  const int col = n_idx * p.n_tile + cv * 8;
  if (r >= p.R || col >= p.cout) return;
  bool valid = true;
  if (!p.dense) {
    const int img = fastdiv(r, p.plane_m, p.plane_s);
    const int pr = r - img * p.plane;
    const int yy = fastdiv(pr, p.wp_m, p.wp_s);
    const int xx = pr - yy * p.wp;
    valid = (yy >= 1) && (yy <= p.hv) && (xx < p.wv);
  }
  uint4 res = make_uint4(0, 0, 0, 0);
  if (valid) {
    const long long slab = static_cast<long long>(kBlockM) * p.n_tile;
    const float* src = p.part + static_cast<long long>(lt) * p.ksplit * slab + static_cast<long long>(row) * p.n_tile + cv * 8;
    float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
    for (int k = 1; k < p.ksplit; ++k) {
      const float4 c = *reinterpret_cast<const float4*>(src + k * slab);
      const float4 d = *reinterpret_cast<const float4*>(src + k * slab + 4);
      a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
      b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
    }
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] += p.bias ? p.bias[col + e] : 0.f;
      if (p.relu) v[e] = fmaxf(v[e], 0.f);
    }
    res = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
  y[static_cast<long long>(r) * y_pitch_v + ((p.y_coff + col) >> 3)] = res;
}

// -------------------------------------------------------------------------------------------------
int choose_n_tile(int cout, int epilogue) {
  if (epilogue != IE_EPI_BF16_RASTER) return ((cout + 15) / 16) * 16;
  if (cout >= 256) return 256;
  if (cout > 64) return 128;
  return 64;
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
  IE_REQUIRE((long long)d->n_img * (d->h + 1) * (d->w + 1) < (1ll << 31) - 4096, "conv: raster too large for 32-bit rows");
  IE_REQUIRE((d->kh == 3 && d->kw == 3) || (d->kh == 2 && d->kw == 2) || (d->kh == 1 && d->kw == 1),
             "conv: unsupported kernel size %dx%d", d->kh, d->kw);
  IE_REQUIRE(d->dense == 0 || d->dense == 1, "conv: dense must be 0 or 1 (got %d)", d->dense);
  if (d->dense) {
    IE_REQUIRE(d->kh != 2 && d->epilogue == IE_EPI_BF16_RASTER && d->hv == d->h && d->wv == d->w,
               "conv: dense NHWC tensors support 3x3 'same' and 1x1 convolutions with the bf16 epilogue only");
    IE_REQUIRE(d->h <= 32768 && d->w <= 32768, "conv: dense image too large for im2col coordinates");
  }
  if (d->epilogue == IE_EPI_BF16_RASTER) {
    IE_REQUIRE(y_bf16, "conv: y_bf16 is null");
    IE_REQUIRE(d->cout % 64 == 0, "conv: bf16 raster epilogue needs cout %% 64 == 0 (got %d)", d->cout);
    IE_REQUIRE(d->y_coff % 64 == 0 && d->y_coff + d->cout <= d->y_pitch && d->y_pitch % 8 == 0,
               "conv: bad output slice (coff %d, cout %d, pitch %d)", d->y_coff, d->cout, d->y_pitch);
  } else if (d->epilogue == IE_EPI_F32_NHWC || d->epilogue == IE_EPI_F32_SOFTMAX) {
    IE_REQUIRE(y_f32, "conv: y_f32 is null");
    IE_REQUIRE(d->cout <= 256, "conv: fp32 epilogues need cout <= 256 per launch (got %d)", d->cout);
    if (d->y_pitch > 0) {      // channel slice of a wider fp32 NHWC tensor (layers with more than 256 outputs run in chunks)
      IE_REQUIRE(d->epilogue == IE_EPI_F32_NHWC && d->cout > 16 && d->y_coff >= 0 && d->y_coff + d->cout <= d->y_pitch,
                 "conv: bad fp32 output slice (coff %d, cout %d, pitch %d; plain NHWC epilogue, cout > 16 only)",
                 d->y_coff, d->cout, d->y_pitch);
    }
  } else {
    IE_REQUIRE(false, "conv: unknown epilogue %d", d->epilogue);
  }
  return IE_OK;
}

// Tuning / test hooks (not part of the documented ABI surface): force a main-loop flavour.
static int g_force_mode = -1;        // -1 auto, 0 stream, 1 resident, 2 wide-N
static int g_fuse_rows = 1;
static int g_splitk = 1;             // 0: never split the K loop of the streaming kernel (tests / A-B timing)
static int g_splitk_tail = 1;        // 0: split-K only for tiny M, not for the last partial wave of larger grids
static int g_stream_transposed = 1;  // 0: cout = 128 layers through M-tile pairs instead of the transposed kernel (A-B timing)
static int g_stream_pairs = 1;       // 0: N = 128 streaming layers one M tile per work item (A-B timing)
static int g_first_gather = 0;       // 1: first layer with per-thread global gathers instead of staged source rows
static int g_splitk_wide = 1;        // 0: small grids always trade N-tile width for CTAs (the round-1 rule)
static bool splitk_env() {           // IE_SPLITK=0: A-B timing without a rebuild
  static const bool on = !(getenv("IE_SPLITK") && getenv("IE_SPLITK")[0] == '0');
  return on;
}
static bool splitk_wide_env() {      // IE_SPLITK_WIDE=0: the round-1 rule
  static const bool on = !(getenv("IE_SPLITK_WIDE") && getenv("IE_SPLITK_WIDE")[0] == '0');
  return on;
}
static int splitk_max() {            // most K slices per tile (the finisher reads one fp32 slab per slice)
  static const int v = getenv("IE_SPLITK_MAX") ? atoi(getenv("IE_SPLITK_MAX")) : 8;    // measured 8 / 16 / 36: 1x32x32 0.297 / 0.300 / 0.325 ms
  return v < 1 ? 1 : v;
}
static int g_wide_prefetch = 1;      // wide-N layers with several channel blocks: L2 prefetch of the next tile's A boxes
static int g_wide_flags = 0;         // tuning: bit 0 stream the weights even if they fit, bit 1 flip the number of
                                     // epilogue sets, bit 2 one filter row per stage even when cin = 64
// (measured on B200: the 128B swizzle is a function of the absolute smem address, so row-shifted descriptor starts
//  need NO base-offset field - setting it corrupts; the experiment lived behind flag bit 0, now unused)

}  // namespace ie

extern "C" int ie_conv_set_mode(int mode, int flags) {
  ie::g_force_mode = mode;
  ie::g_fuse_rows = (flags & 2) ? 0 : 1;
  ie::g_wide_flags = ((flags >> 2) & 7) | (((flags >> 11) & 1) << 3);      // bit 11: exchange epilogue for one-block layers
  ie::g_splitk = ((flags >> 9) & 1) ? 0 : 1;
  ie::pdl_set(((flags >> 10) & 1) == 0);
  ie::g_wide_prefetch = ((flags >> 12) & 1) ? 0 : 1;
  ie::g_splitk_tail = ((flags >> 13) & 1) ? 0 : 1;
  ie::g_splitk_wide = ((flags >> 14) & 1) ? 0 : 1;
  ie::g_first_gather = (flags >> 15) & 1;
  ie::g_stream_pairs = ((flags >> 16) & 1) ? 0 : 1;
  ie::g_stream_transposed = ((flags >> 17) & 1) ? 0 : 1;
  return IE_OK;
}

extern "C" int ie_conv2d_nhwc_bf16(const ie_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                                   void* y_bf16, float* y_f32, float* y_aux, void* workspace, long long workspace_bytes,
                                   void* stream) {
  using namespace ie;
  int rc = validate_conv_desc(d, x, w_packed, y_bf16, y_f32);
  if (rc) return rc;
  const long long R = d->dense ? (long long)d->n_img * d->h * d->w : (long long)d->n_img * (d->h + 1) * (d->w + 1);
  const int wp = d->w + 1;
  EpiParams e{};
  e.dense = d->dense;
  e.ksplit = 1;
  e.msub = 1;
  e.split_first = 0;                   // irrelevant while ksplit == 1 (every item is a whole tile)
  e.R = (int)R;
  set_raster_dims(e, (d->h + 1) * wp, wp);
  e.hv = d->hv;
  e.wv = d->wv;
  e.cout = d->cout;
  e.n_tile = choose_n_tile(d->cout, d->epilogue);
  e.m_tiles = (int)((R + kBlockM - 1) / kBlockM);
  // small rasters (the per-image basis branch: 256 images x 2x2 px = 8 M tiles): the grid must cover the SMs - a
  // 2048->512 layer on 36 CTAs ran 102 us with 112 SMs idle. Two ways: narrower N tiles (every CTA re-reads the whole
  // 128-row A tile for fewer output columns) or split-K over wide tiles (each A and weight byte is read once per
  // N tile). With a workspace and a deep K loop split-K provides the CTAs and the tile stays wide.
  const int kblocks_all = d->kh * d->kw * (d->cin / 64);
  const bool can_split = g_splitk && splitk_env() && workspace && d->epilogue == IE_EPI_BF16_RASTER && kblocks_all >= 8;
  if (d->epilogue == IE_EPI_BF16_RASTER) {
    while (e.n_tile > 64 && d->cout % (e.n_tile / 2) == 0 &&
           (long long)e.m_tiles * ((d->cout + e.n_tile - 1) / e.n_tile) * 2 <= sm_count()) {
      if (can_split && g_splitk_wide && splitk_wide_env()) {
        const long long tiles_now = (long long)e.m_tiles * ((d->cout + e.n_tile - 1) / e.n_tile);
        int max_split = kblocks_all / 4;
        if (max_split > splitk_max()) max_split = splitk_max();
        if (tiles_now * max_split * 2 > sm_count()) break;
      }
      e.n_tile /= 2;
    }
  }
  e.n_tiles = (d->cout + e.n_tile - 1) / e.n_tile;
  e.y_coff = d->y_coff;
  e.relu = d->relu;
  e.epilogue = d->epilogue;
  e.bias = bias;
  e.f32_pitch = d->cout;
  e.y_f32 = y_f32;
  if (d->epilogue != IE_EPI_BF16_RASTER && d->y_pitch > 0) {
    e.f32_pitch = d->y_pitch;
    e.y_f32 = y_f32 + d->y_coff;
  }
  e.y_aux = y_aux;
  const int ntaps = d->kh * d->kw;
  const int ktot = ntaps * d->cin;
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  CUtensorMap tm_a, tm_b, tm_y;
  rc = make_tmap_2d_bf16(&tm_b, w_packed, (uint64_t)ktot, (uint64_t)(e.n_tiles * e.n_tile), (uint64_t)ktot, 64,
                         (uint32_t)e.n_tile);
  if (rc) return rc;
  if (d->epilogue == IE_EPI_BF16_RASTER) {
    rc = make_tmap_2d_bf16(&tm_y, y_bf16, (uint64_t)d->y_pitch, (uint64_t)R, (uint64_t)d->y_pitch, 64, 32);
    if (rc) return rc;
  }

  // ---- resident-weights flavour: narrow layers whose whole weight matrix fits beside >= 3 A stages
  const int b_res_bytes = ntaps * (d->cin / 64) * e.n_tile * 128;
  const int box_rows = kBlockM + d->kw - 1;
  int a_slot = ((box_rows * 128 + 1023) / 1024) * 1024;
  const int res_room = (int)kMaxSmem - 1024 - kTailBytes2 - (b_res_bytes <= 200 * 1024 ? b_res_bytes : 200 * 1024);
  const int fuse_rows = (d->kh == 3 && d->cin == 64 && g_fuse_rows) ? 3 : 1;      // whole 3x3x64 tile in one stage
  const int a_box = a_slot;
  a_slot *= fuse_rows;
  const int res_stages = res_room > 0 ? res_room / a_slot : 0;
  bool resident = e.n_tiles == 1 && e.n_tile <= 64 && b_res_bytes <= 160 * 1024 && res_stages >= (fuse_rows == 3 ? 2 : 3);
  if (g_force_mode == 0 || d->dense) resident = false;          // dense tensors: the streaming kernel with im2col loads
  if (g_force_mode == 1)
    IE_REQUIRE(e.n_tiles == 1 && b_res_bytes <= 200 * 1024 && res_stages >= 2 && !d->dense,
               "conv: resident mode forced but the weights do not fit (or the tensors are dense)");
  if (g_force_mode == 1) resident = true;

  const int grid_cap = sm_count();

  // ---- wide-N flavour: 3x3, cout <= 64 (one 64-column group per horizontal tap), bf16 raster output
  const bool wide_ok = d->kh == 3 && d->kw == 3 && d->epilogue == IE_EPI_BF16_RASTER && e.n_tile == 64 && e.n_tiles == 1;
  // measured on B200 (tools/conv_bench.py, 256 x 104^2, us; resident = row-shifted descriptors at N = 64):
  //   cin  64: resident 206, wide-N 200 (two epilogue sets; 251 with one: the epilogue reads 3x the TMEM columns)
  //   cin 128: resident 527, wide-N 370 (weights resident, one epilogue set -> 3 A stages)
  //   cin 640: streaming N = 64 870, wide-N 417 (weights streamed with the A tiles)
  // small fp32 heads (cout <= 16): three 16-column groups, N = 48 - a third of the A reads of the N = 16 resident path
  const bool wide_f32 = d->kh == 3 && d->kw == 3 && d->epilogue != IE_EPI_BF16_RASTER && d->cout <= 16 && d->cin <= 128;
  bool wide = (wide_ok || wide_f32) && !d->dense;
  if (g_force_mode == 0 || g_force_mode == 1) wide = false;
  if (g_force_mode == 2) wide = (wide_ok || wide_f32) && !d->dense;
  if (g_force_mode == 2) IE_REQUIRE(wide, "conv: wide-N mode forced on an unsupported layer");
  if (wide) {
    WideParams p{};
    p.e = e;
    for (int i = 0; i < 3; ++i) p.dy_shift[i] = (i - 1) * wp;
    p.kb = d->cin / 64;
    p.x_coff = d->x_coff;
    p.cin = d->cin;
    p.gw = wide_f32 ? 16 : 64;
    p.b_tile_bytes = 3 * p.gw * 128;
    const int w_bytes = 3 * p.kb * p.b_tile_bytes;
    const bool res = w_bytes <= 150 * 1024 && !(g_wide_flags & 1);
    const bool fuse = res && p.kb == 1 && !(g_wide_flags & 4);
    // one channel block (the 64 -> 64 layers, the coef head): the tile's main loop is 12 MMAs and the epilogue is the
    // critical path -> per-warp slabs (no cross-warp exchange, 120 outputs per tile, four 32-row boxes per filter row).
    // Deeper layers hide the epilogue behind their main loop and are bound by TMA latency / issue instead -> one
    // 128-row box per stage, 126 outputs per tile, rows on warp boundaries exchanged through shared memory
    // (measured on B200 with slabs everywhere: 640 -> 64 417 -> 601 us, 128 -> 64 370 -> 479 us).
    const bool slab = fuse && !(g_wide_flags & 8);
    const int tile_rows = slab ? kWideTileRows : kXchTileRows;
    p.e.m_tiles = (int)((R + tile_rows - 1) / tile_rows);
    p.a_slot_bytes = fuse ? 3 * kABytes : kABytes + (res ? 0 : p.b_tile_bytes);
    // two epilogue sets when the main loop of a tile is short (one channel block); deeper layers hide the epilogue
    // anyway and need the 16 KB for pipeline stages (measured: 128->64 resident 378 us with one set, 487 with two)
    p.nsets = (p.kb == 1) ? 2 : 1;
    if (g_wide_flags & 2) p.nsets = 3 - p.nsets;
    p.prefetch = g_wide_prefetch && res;      // measured (B200): 128 -> 64 resident 379.6 -> 370.5 us; 640 -> 64 streamed 451 -> 454
    const int tail = tail_bytes(p.nsets);
    int stages = ((int)kMaxSmem - 1024 - tail - (res ? w_bytes : 0)) / p.a_slot_bytes;
    p.stages = stages > kMaxStages ? kMaxStages : stages;
    IE_REQUIRE(p.stages >= 2, "conv: wide-N pipeline does not fit in shared memory");
    CUtensorMap tm_y1, tm_y2;
    rc = make_tmap_2d_bf16(&tm_a, x, (uint64_t)d->x_pitch, (uint64_t)R, (uint64_t)d->x_pitch, 64, slab ? kSlabRows : kBlockM);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tm_b, w_packed, (uint64_t)ktot, (uint64_t)p.gw, (uint64_t)ktot, 64, (uint32_t)p.gw);
    if (rc) return rc;
    if (wide_f32) {
      tm_y1 = tm_a;                     // unused by the fp32 epilogues
      tm_y2 = tm_a;
    } else {
      rc = make_tmap_2d_bf16(&tm_y1, y_bf16, (uint64_t)d->y_pitch, (uint64_t)R, (uint64_t)d->y_pitch, 64, slab ? kSlabOut : 32);
      if (rc) return rc;
      rc = make_tmap_2d_bf16(&tm_y2, y_bf16, (uint64_t)d->y_pitch, (uint64_t)R, (uint64_t)d->y_pitch, 64, 31);
      if (rc) return rc;
    }
    const size_t smem = 1024 + (size_t)(res ? w_bytes : 0) + (size_t)p.stages * p.a_slot_bytes + tail;
    const int grid = p.e.m_tiles < grid_cap ? p.e.m_tiles : grid_cap;
#define IE_LAUNCH_WIDE(RES_, G_, SLAB_)                                                                           \
  do {                                                                                                            \
    IE_CUDA(cudaFuncSetAttribute(conv_wide_kernel<RES_, G_, SLAB_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                 (int)kMaxSmem));                                                                 \
    IE_CUDA(launch_pdl(conv_wide_kernel<RES_, G_, SLAB_>, dim3(grid), dim3(kThreads2), smem, st, tm_a, tm_b, tm_y1, \
                       tm_y2, p));                                                                                \
  } while (0)
    if (fuse && slab) IE_LAUNCH_WIDE(true, 3, true);
    else if (fuse) IE_LAUNCH_WIDE(true, 3, false);
    else if (res) IE_LAUNCH_WIDE(true, 1, false);
    else IE_LAUNCH_WIDE(false, 1, false);
#undef IE_LAUNCH_WIDE
    return IE_OK;
  }

  if (resident) {
    ResidentParams p{};
    p.e = e;
    p.ndy = d->kh;
    p.ndx = d->kw;
    for (int i = 0; i < d->kh; ++i) p.dy_shift[i] = (d->kh == 3) ? (i - 1) * wp - 1 : i * wp;
    p.kb = d->cin / 64;
    p.x_coff = d->x_coff;
    p.cin = d->cin;
    p.stages = res_stages > kMaxStages ? kMaxStages : res_stages;
    p.box_rows = box_rows;
    p.a_box_bytes = a_box;
    p.a_slot_bytes = a_slot;
    p.b_tile_bytes = e.n_tile * 128;
    rc = make_tmap_2d_bf16(&tm_a, x, (uint64_t)d->x_pitch, (uint64_t)R, (uint64_t)d->x_pitch, 64, (uint32_t)box_rows);
    if (rc) return rc;
    if (d->epilogue != IE_EPI_BF16_RASTER) tm_y = tm_a;
    const size_t smem = 1024 + (size_t)b_res_bytes + (size_t)p.stages * a_slot + kTailBytes2;
    const int grid = e.m_tiles < grid_cap ? e.m_tiles : grid_cap;
#define IE_LAUNCH_RES(NDX_, G_)                                                                                   \
  do {                                                                                                            \
    IE_CUDA(cudaFuncSetAttribute(conv_resident_kernel<NDX_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                 (int)kMaxSmem));                                                                 \
    IE_CUDA(launch_pdl(conv_resident_kernel<NDX_, G_>, dim3(grid), dim3(kThreads2), smem, st, tm_a, tm_b, tm_y, p)); \
  } while (0)
    if (d->kw == 3 && fuse_rows == 3) IE_LAUNCH_RES(3, 3);
    else if (d->kw == 3) IE_LAUNCH_RES(3, 1);
    else if (d->kw == 2) IE_LAUNCH_RES(2, 1);
    else IE_LAUNCH_RES(1, 1);
#undef IE_LAUNCH_RES
    IE_LAUNCH_CHECK();
    return IE_OK;
  }

  StreamParams p{};
  p.e = e;
  p.ntaps = ntaps;
  for (int i = 0; i < d->kh; ++i)
    for (int j = 0; j < d->kw; ++j)
      p.tap_shift[i * d->kw + j] = (d->kh == 3) ? (i - 1) * wp + (j - 1) : i * wp + j;
  p.kblocks_per_tap = d->cin / 64;
  p.x_coff = d->x_coff;
  p.cin = d->cin;
  p.b_stage_bytes = ((e.n_tile * 128 + 1023) / 1024) * 1024;
  // N tile 128 with one N tile (the cout = 128 layers) and enough rows to fill the machine twice: pairs of M tiles per
  // work item sharing every weight tile (conv_stream_kernel<.., 2>)
  const int msub = (g_stream_pairs && e.n_tile == 128 && e.n_tiles == 1 && d->epilogue == IE_EPI_BF16_RASTER &&
                    e.m_tiles >= 2 * sm_count()) ? 2 : 1;
  // ... or, better, the transposed flavour (pixels as the N = 256 dimension) when cout is exactly one 128-row A tile
  // (from cin = 128 up: with 9 K blocks per item the transposing epilogue - 16 named barriers per 256 pixels - is not
  //  hidden behind 4 600 MMA cycles; measured 64->128 @52^2: 99 us with pairs, 115 transposed; 128->128: 172 vs 153)
  const bool transposed = msub == 2 && g_stream_transposed && d->cout == 128 && kblocks_all >= 16;
  if (msub == 2 && !transposed) {
    p.e.msub = 2;
    p.e.m_tiles = (e.m_tiles + 1) / 2;
    e.m_tiles = p.e.m_tiles;
  }
  const int msub_k = transposed ? 1 : msub;            // (the transposed launch sizes its own pipeline)
  int stages = (int)((kMaxSmem - 1024 - kTailBytes) / (msub_k * kABytes + p.b_stage_bytes));
  p.stages = stages > kMaxStages ? kMaxStages : stages;
  p.img_h = d->h;
  p.img_w = d->w;
  p.kw = d->kw;
  p.pad = d->kh / 2;
  const size_t smem = 1024 + (size_t)p.stages * (msub_k * kABytes + p.b_stage_bytes) + kTailBytes;
  const int tiles = e.m_tiles * e.n_tiles;
  // split-K (see EpiParams): (a) tiny M - eval.py's default call is ONE 32 x 32 patch: the 1024-channel layers are
  // then 1 M tile x 16 N tiles with K = 9216 - 18432, i.e. 16 CTAs streaming 19 - 38 MB of weights while 132 SMs idle -
  // every tile is split; (b) a persistent grid whose last wave is less than half full (2048 -> 512 at 256 x 26 x 26:
  // 2704 tiles = 18.27 waves on 148 SMs): only the tiles of that wave are split, so all SMs finish together
  const int kblocks = ntaps * p.kblocks_per_tap;
  int ksplit = 1, split_first = tiles;
  const size_t slab_bytes = (size_t)kBlockM * e.n_tile * sizeof(float);
  if (can_split && msub == 1) {
    int tail = tiles * 2 <= grid_cap ? tiles : tiles % grid_cap;            // (a) all tiles, (b) the last wave
    if (tiles * 2 > grid_cap && (g_splitk_tail == 0 || tiles < grid_cap)) tail = 0;
    if (tail > 0 && tail * 2 <= grid_cap) {
      ksplit = grid_cap / tail;
      if (ksplit > kblocks / 4) ksplit = kblocks / 4;
      if (ksplit > splitk_max()) ksplit = splitk_max();
      while (ksplit > 1 && slab_bytes * tail * ksplit > (size_t)workspace_bytes) --ksplit;
      const int kb_per = (kblocks + ksplit - 1) / ksplit;
      ksplit = (kblocks + kb_per - 1) / kb_per;            // no empty split
      if (ksplit > 1) split_first = tiles - tail;
    }
  }
  if (ksplit == 1) split_first = tiles;
  p.e.ksplit = ksplit;
  p.e.split_first = split_first;
  p.e.part = static_cast<float*>(workspace);
  const int items = split_first + (tiles - split_first) * ksplit;
  const int grid = items < grid_cap ? items : grid_cap;
  if (transposed) {
    // D^T = W * X^T (conv_streamT_kernel): items of 256 pixels, weights as the A operand
    StreamParams pt = p;
    pt.e.msub = 1;
    pt.e.m_tiles = (int)((R + kPixT - 1) / kPixT);
    pt.stages = (int)((kMaxSmem - 1024 - kTailBytes) / (128 * 128 + kPixT * 128));
    if (pt.stages > kMaxStages) pt.stages = kMaxStages;
    const size_t smem_t = 1024 + (size_t)pt.stages * (128 * 128 + kPixT * 128) + kTailBytes;
    const int grid_t = pt.e.m_tiles < grid_cap ? pt.e.m_tiles : grid_cap;
    if (d->dense) {
      rc = make_tmap_im2col_bf16(&tm_a, x, d->n_img, d->h, d->w, (uint64_t)d->x_pitch, pt.pad, kPixT);
      if (rc) return rc;
      IE_CUDA(cudaFuncSetAttribute(conv_streamT_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
      IE_CUDA(launch_pdl(conv_streamT_kernel<true>, dim3(grid_t), dim3(kThreads), smem_t, st, tm_a, tm_b, tm_y, pt));
    } else {
      rc = make_tmap_2d_bf16(&tm_a, x, (uint64_t)d->x_pitch, (uint64_t)R, (uint64_t)d->x_pitch, 64, kPixT);
      if (rc) return rc;
      IE_CUDA(cudaFuncSetAttribute(conv_streamT_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
      IE_CUDA(launch_pdl(conv_streamT_kernel<false>, dim3(grid_t), dim3(kThreads), smem_t, st, tm_a, tm_b, tm_y, pt));
    }
    IE_LAUNCH_CHECK();
    return IE_OK;
  }
  if (d->dense) {
    rc = make_tmap_im2col_bf16(&tm_a, x, d->n_img, d->h, d->w, (uint64_t)d->x_pitch, p.pad, kBlockM);
    if (rc) return rc;
    if (msub == 2) {
      IE_CUDA(cudaFuncSetAttribute(conv_stream_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
      IE_CUDA(launch_pdl(conv_stream_kernel<true, 2>, dim3(grid), dim3(kThreads), smem, st, tm_a, tm_b, tm_y, p));
    } else {
      IE_CUDA(cudaFuncSetAttribute(conv_stream_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
      IE_CUDA(launch_pdl(conv_stream_kernel<true, 1>, dim3(grid), dim3(kThreads), smem, st, tm_a, tm_b, tm_y, p));
    }
  } else {
    rc = make_tmap_2d_bf16(&tm_a, x, (uint64_t)d->x_pitch, (uint64_t)R, (uint64_t)d->x_pitch, 64, kBlockM);
    if (rc) return rc;
    if (d->epilogue != IE_EPI_BF16_RASTER) tm_y = tm_a;
    if (msub == 2) {
      IE_CUDA(cudaFuncSetAttribute(conv_stream_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
      IE_CUDA(launch_pdl(conv_stream_kernel<false, 2>, dim3(grid), dim3(kThreads), smem, st, tm_a, tm_b, tm_y, p));
    } else {
      IE_CUDA(cudaFuncSetAttribute(conv_stream_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
      IE_CUDA(launch_pdl(conv_stream_kernel<false, 1>, dim3(grid), dim3(kThreads), smem, st, tm_a, tm_b, tm_y, p));
    }
  }
  IE_LAUNCH_CHECK();
  if (ksplit > 1) {
    const long long total = (long long)(tiles - split_first) * kBlockM * (e.n_tile / 8);
    IE_CUDA(launch_pdl(splitk_finish_kernel, dim3(ie_ceil_div(total, 256)), dim3(256), 0, st, p.e, static_cast<uint4*>(y_bf16),
                       d->y_pitch / 8));
  }
  return IE_OK;
}

extern "C" int ie_conv_first_layer_f32(const float* x, int n, int hs, int ws, int c, int h, int w, const void* w_packed,
                                       const float* bias, int cout, int relu, void* y_bf16, int y_pitch, int y_coff,
                                       void* stream) {
  using namespace ie;
  IE_REQUIRE(x && w_packed && y_bf16, "conv_first: null pointer");
  IE_REQUIRE(n > 0 && hs > 0 && ws > 0 && hs <= h && ws <= w, "conv_first: source %dx%d must fit the %dx%d raster", hs, ws, h, w);
  IE_REQUIRE(c == 3 || c == 5 || c == 10, "conv_first: fused first layer is built for 3, 5 or 10 input channels (got %d); "
                                         "use ie_pack_input_im2col3x3 + ie_conv2d_nhwc_bf16", c);
  IE_REQUIRE(cout == 64, "conv_first: cout must be 64 (got %d)", cout);
  IE_REQUIRE(y_coff % 64 == 0 && y_coff + cout <= y_pitch && y_pitch % 8 == 0, "conv_first: bad output slice");
  const long long R = (long long)n * (h + 1) * (w + 1);
  IE_REQUIRE(R < (1ll << 31) - 4096, "conv_first: raster too large for 32-bit rows");
  FirstParams p{};
  p.e.R = (int)R;
  set_raster_dims(p.e, (h + 1) * (w + 1), w + 1);
  p.e.hv = h;
  p.e.wv = w;
  p.e.cout = cout;
  p.e.n_tile = 64;
  p.e.n_tiles = 1;
  p.e.m_tiles = (int)((R + kBlockM - 1) / kBlockM);
  p.e.y_coff = y_coff;
  p.e.relu = relu;
  p.e.epilogue = IE_EPI_BF16_RASTER;
  p.e.ksplit = 1;
  p.e.msub = 1;
  p.e.split_first = 0;
  p.e.bias = bias;
  p.x = x;
  p.hs = hs;
  p.ws = ws;
  const int kb = (9 * c + 63) / 64;                       // K blocks of 64: w_packed is [64][64 * kb]
  p.total_floats = (long long)n * hs * ws * c;
  CUtensorMap tm_b, tm_y;
  int rc = make_tmap_2d_bf16(&tm_b, w_packed, 64 * kb, 64, 64 * kb, 64, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tm_y, y_bf16, (uint64_t)y_pitch, (uint64_t)R, (uint64_t)y_pitch, 64, 32);
  if (rc) return rc;
  // staged flavour (bulk copies of the source rows into shared memory) when the copies can be 16-byte aligned
  const bool staged = !g_first_gather && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && p.total_floats % 4 == 0 &&
                      p.total_floats < (1ll << 31);
  p.stages = !staged ? (kb == 1 ? 6 : 4) : c == 5 ? FirstStaged<5>::kAStages : c == 3 ? FirstStaged<3>::kAStages : FirstStaged<10>::kAStages;
  const int ring = !staged ? 0 : c == 5 ? FirstStaged<5>::kRing : c == 3 ? FirstStaged<3>::kRing : FirstStaged<10>::kRing;
  const size_t smem = 1024 + (size_t)kb * 64 * 128 + (size_t)p.stages * kb * kABytes + ring + (staged ? kTailBytes2 : kTailBytes);
  const int grid = p.e.m_tiles < sm_count() ? p.e.m_tiles : sm_count();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define IE_LAUNCH_FIRST(C_)                                                                                        \
  do {                                                                                                             \
    if (staged) {                                                                                                  \
      IE_CUDA(cudaFuncSetAttribute(conv_first_staged_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                   (int)kMaxSmem));                                                                \
      IE_CUDA(launch_pdl(conv_first_staged_kernel<C_>, dim3(grid), dim3(kFirstStagedWarps * 32), smem, st, tm_b, tm_y, p)); \
    } else {                                                                                                       \
      IE_CUDA(cudaFuncSetAttribute(conv_first_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem)); \
      IE_CUDA(launch_pdl(conv_first_kernel<C_>, dim3(grid), dim3(kFirstThreads), smem, st, tm_b, tm_y, p));        \
    }                                                                                                              \
  } while (0)
  if (c == 5) IE_LAUNCH_FIRST(5);
  else if (c == 3) IE_LAUNCH_FIRST(3);
  else IE_LAUNCH_FIRST(10);
#undef IE_LAUNCH_FIRST
  return IE_OK;
}
