// =================================================================================================
// Per-pixel filter on tcgen05 in the reference's own order of operations (model_library.py:439-451):
//
//   1. filter synthesis as a GEMM:  F[128 px][(tap, t)] = coef[128 px][B] * Bas[B][(tap, t)]   (kind::f16, fp16 operands)
//      A = the tile's coefficients (hand-swizzled K-major SWIZZLE_128B rows, one per pixel), B operand = the image's
//      basis, resident in shared memory for the CTA's lifetime as rows n = tap * 4 + t (the basis' memory order),
//      accumulators in TMEM: 240 columns = 4 filter rows x 15 x 4 frames per buffer, two buffers;
//   2. Convolve / Convolve_perlayer (:114-168) in the epilogue: the thread that owns TMEM lane p = pixel p of the 8 x 16
//      tile reads its filter values with tcgen05.ld and multiplies them with the burst window from a shared-memory
//      tile (one LDS.128 = the four frames of a tap).  The per-pixel filters never reach memory.
//
// Same 2*K*K*T*B MMA FLOP per pixel as kpn_apply_tf32_kernel (which computes G = Bas (*) burst with warp-level mma.sync
// and mixes with coef afterwards), but here K = B is tiny and N = the taps, so the GEMM runs at tcgen05 rates and the
// CUDA cores are left with K*K*T = 900 FMAs per pixel.  Warp roles: 0-3 / 4-7 two epilogue sets (set s drains TMEM
// buffer s = chunks s, s+2; set 1 hands its partial sums to set 0 through shared memory), 8 MMA issuer, 9-12 producers
// staging the next tile (coefficients -> fp16 -> swizzled A slot; burst window by 4-byte cp.async with zero fill).
//
// Persistent CTAs: the launch's (image, tile) list is cut into gridDim.x contiguous ranges; a CTA re-stages the basis
// whenever its range crosses into the next image (a CTA-wide barrier; ~1.7 images per CTA at 256 images on 148 SMs -
// one CTA per (image, strip) left 15 % of the time to the second, partial wave).
// Operand precision: coefficients and basis are softmax outputs in [0, 1]; fp16 (10-bit mantissa, round to nearest)
// carries them exactly as well as TF32 does (values below 6e-5 fall into fp16's subnormals: absolute error <= 3e-8 per
// tap), accumulation is fp32, the burst is never rounded.  fp16 rows hold 64 bases per 128-byte swizzle row instead of
// 32: Basis_kpn's B = 50 is ONE block, B = 90 two.
// More than 64 bases (Basis_kpn of remote/: B = 90): the filter is linear in the bases, so blocks of 64 bases run as
// further launches that ADD into the output; T > 4 likewise in passes of four frames.
//
// Measured on B200 (round 2): parity test green through the C ABI (coef at the padded extent, pitch != T + 1); the
// epilogue is bound by the TMEM read path - every pixel reads its 900 synthesised filter values (3.6 KB) with tcgen05.ld.
// Scope: K = 15, T a multiple of 4, any B (blocks of 64).
// =================================================================================================
#include "ie_common.cuh"
#include "ie_ptx.cuh"

namespace ie {
namespace tcf {

constexpr int kK = 15, kTaps = kK * kK, kTP = 4;          // frames per pass
constexpr int kTileH = 8, kTileW = 16;                    // 128 pixels = 128 TMEM lanes
constexpr int kSH = kTileH + kK - 1, kSW = kTileW + kK - 1;   // burst tile with halo: 22 x 30 pixels x 4 frames
constexpr int kRowsPerChunk = 4;                          // filter rows per accumulator buffer
constexpr int kChunkTaps = kRowsPerChunk * kK;            // 60
constexpr int kChunkN = kChunkTaps * kTP;                 // 240 columns
constexpr int kLastTaps = kTaps - 3 * kChunkTaps;         // 45
constexpr int kLastN = 192;                               // 180 real columns, N must be a multiple of 16
constexpr int kNumChunks = 4;
constexpr int kBRows = kTaps * kTP + (kLastN - kLastTaps * kTP);   // 900 + 12 rows the last MMA also reads
constexpr int kBBytes = kBRows * 128;
constexpr int kABytes = 128 * 128;
constexpr int kBurstBytes = kSH * kSW * kTP * 4;
constexpr int kPartBytes = 128 * 16;                        // partial sums of epilogue set 1, per slot
constexpr int kSmemBytes = 1024 + kBBytes + 2 * kABytes + 2 * kBurstBytes + 2 * kPartBytes + 128;
constexpr int kThreads = 416;        // warps 0-3 / 4-7: epilogue sets 0 / 1, warp 8: MMA issuer, warps 9-12: producers
constexpr int kMmaWarp = 8, kFirstProducerWarp = 9;
constexpr int kTmemCols = 512;

constexpr int kBlockBases = 64;                           // bases per launch: one 128-byte row of fp16

__device__ __forceinline__ uint32_t idesc_f16(int M, int N) {
  // c_format f32 (1 << 4), a_format / b_format F16 (0), both K-major, N >> 3 at bit 17, M >> 4 at 24
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint16_t to_h(float x) {
  uint16_t r;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return r;
}
// byte offset of element k (fp16) of row r inside a K-major SWIZZLE_128B operand: 128-byte rows, the 16-byte chunk
// index XOR-ed with the row's position in its 8-row group (what TMA writes; csrc/conv_tcgen05.cu builds its
// first-layer A tile the same way)
__device__ __forceinline__ uint32_t sw128_off_h(int r, int k) {
  return static_cast<uint32_t>(r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1)));
}

struct Params {
  const float* burst;
  const float* coef;
  const float* bas;
  float* out;
  int n, h, w, hc, wc, pitch, Ttot, B, t0;            // coef is [n][hc][wc][B], hc >= h, wc >= w
  int b0, nb;                                          // this launch mixes bases [b0, b0 + nb), nb <= 64
  int acc0, accf;                                      // add into out[..., 0] / into the per-frame channels
  int tiles_x, tiles_y, ksteps;
};

// 4-byte asynchronous global->shared copy; src_bytes = 0 zero-fills (tf.pad, model_library.py:126)
__device__ __forceinline__ void cp_async4_zfill(uint32_t smem_dst, const float* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int NTAPS>
__device__ __forceinline__ void consume_group(const uint32_t (&v)[16], int g, const float4* __restrict__ win, float (&acc)[kTP]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int tap = g * 4 + q;
    if (tap < NTAPS) {
      const int il = tap / kK, j = tap - il * kK;
      const float4 b = win[il * kSW + j];
      acc[0] = fmaf(__uint_as_float(v[q * 4 + 0]), b.x, acc[0]);
      acc[1] = fmaf(__uint_as_float(v[q * 4 + 1]), b.y, acc[1]);
      acc[2] = fmaf(__uint_as_float(v[q * 4 + 2]), b.z, acc[2]);
      acc[3] = fmaf(__uint_as_float(v[q * 4 + 3]), b.w, acc[3]);
    }
  }
}

// One accumulator buffer = NTAPS taps x 4 frames of this thread's pixel: multiply with the burst window.  The load of
// group g+1 is issued (after the wait that completes group g) before group g is consumed.
template <int NTAPS>
__device__ __forceinline__ void apply_chunk(uint32_t taddr, const float4* __restrict__ win, float (&acc)[kTP]) {
  constexpr int kGroups = (NTAPS * kTP + 15) / 16;
  uint32_t va[16], vb[16];
  tmem_ld_x16(taddr, va);
#pragma unroll
  for (int g = 0; g < kGroups; g += 2) {
    tmem_ld_wait();
    if (g + 1 < kGroups) tmem_ld_x16(taddr + (g + 1) * 16, vb);
    consume_group<NTAPS>(va, g, win, acc);
    if (g + 1 < kGroups) {
      tmem_ld_wait();
      if (g + 2 < kGroups) tmem_ld_x16(taddr + (g + 2) * 16, va);
      consume_group<NTAPS>(vb, g + 1, win, acc);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) kpn_tcgen05_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_b = base;                                    // [kBRows][128 B]  basis, rows n = tap * 4 + t
  uint8_t* s_a = base + kBBytes;                          // [2 slots][128 px][128 B]  coefficients of a tile
  uint8_t* s_burst = s_a + 2 * kABytes;                   // [2 slots][22][30] pixels x 4 frames
  float4* s_part = reinterpret_cast<float4*>(s_burst + 2 * kBurstBytes);   // [2 slots][128 px] partial sums of set 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_part) + 2 * kPartBytes);
  uint64_t* acc_full = bars;                              // [2] accumulator buffer complete (MMA -> epilogue set)
  uint64_t* acc_empty = bars + 2;                         // [2] accumulator buffer drained (epilogue set -> MMA)
  uint64_t* in_full = bars + 4;                           // [2] A + window slot staged (producers -> MMA, epilogue)
  uint64_t* in_empty = bars + 6;                          // [2] slot free again (both epilogue sets -> producers)
  uint64_t* part_full = bars + 8;                         // [2] partial sums written (set 1 -> set 0)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kused = p.ksteps * 16;                       // fp16 elements of K in use (k-steps of 16)
  pdl_trigger();

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);                        // the four warps of the set that owns the buffer
      mbar_init(&in_full[b], 4);                          // one arrive per producer warp
      mbar_init(&in_empty[b], 8);                         // one arrive per epilogue warp (both sets)
      mbar_init(&part_full[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                             // the basis, the coefficients and (later passes) out are ready
  const uint32_t tmem_base = *tmem_slot;
  const int tiles = p.tiles_x * p.tiles_y;
  // this CTA's contiguous range of the (image, tile) list
  const long long total = static_cast<long long>(p.n) * tiles;
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long g_begin = blockIdx.x * per;
  const long long g_end = (g_begin + per < total) ? g_begin + per : total;
  int it = 0;                                             // tiles this CTA has processed: slot / barrier phases

  for (long long g = g_begin; g < g_end;) {
    const int img = static_cast<int>(g / tiles);
    const int tile_lo = static_cast<int>(g - static_cast<long long>(img) * tiles);
    const long long seg_end = (static_cast<long long>(img + 1) * tiles < g_end) ? static_cast<long long>(img + 1) * tiles : g_end;
    const int tile_hi = tile_lo + static_cast<int>(seg_end - g);
    // ---- the image's basis -> B operand: Bas[img][tap][t0 + t][b0 + b] -> row tap * 4 + t, slot b, fp16.  Every MMA
    //      that read the previous image's rows has completed: the epilogue consumed its results before the barrier below.
    {
      const float* bas_img = p.bas + static_cast<long long>(img) * kTaps * p.Ttot * p.B;
      // eight independent loads in flight per thread (the loop is latency-bound: up to 58 368 scattered 4-byte loads
      // per image at 64 bases)
      constexpr int kU = 8;
      for (int idx0 = threadIdx.x; idx0 < kBRows * kused; idx0 += kThreads * kU) {
        float v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int idx = idx0 + u * kThreads;
          const int n = idx / kused, b = idx - n * kused;
          v[u] = 0.f;
          if (idx < kBRows * kused && n < kTaps * kTP && b < p.nb) {
            const int tap = n >> 2, t = n & 3;
            v[u] = __ldg(bas_img + (static_cast<long long>(tap) * p.Ttot + p.t0 + t) * p.B + p.b0 + b);
          }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int idx = idx0 + u * kThreads;
          if (idx < kBRows * kused) {
            const int n = idx / kused, b = idx - n * kused;
            *reinterpret_cast<uint16_t*>(s_b + sw128_off_h(n, b)) = to_h(v[u]);
          }
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

    if (warp >= kFirstProducerWarp) {
      // ================================ producers: stage tile `it` into slot it & 1 ==================
      const int pid = threadIdx.x - kFirstProducerWarp * 32;                        // 0..127 = pixel of the tile
      const float* burst_img = p.burst + static_cast<long long>(img) * p.h * p.w * p.pitch;
      for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
        const int slot = it & 1;
        const uint32_t use = static_cast<uint32_t>(it >> 1);
        const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        const int y0 = ty * kTileH, x0 = tx * kTileW;
        // this thread's pixel: its coefficients first into registers (all loads in flight), rounded to fp16 below
        const int y = y0 + (pid >> 4), x = x0 + (pid & 15);
        const bool inside = y < p.h && x < p.w;
        const float* cp = p.coef + ((static_cast<long long>(img) * p.hc + (inside ? y : 0)) * p.wc + (inside ? x : 0)) * p.B + p.b0;
        uint32_t pk[kBlockBases / 2];                                              // fp16 pairs
#pragma unroll
        for (int hb = 0; hb < kBlockBases; hb += 32) {                             // 32 loads in flight at a time
          float v[32];
#pragma unroll
          for (int b = 0; b < 32; ++b) v[b] = (hb + b < kused && hb + b < p.nb && inside) ? __ldg(cp + hb + b) : 0.f;
#pragma unroll
          for (int b = 0; b < 32; b += 2) pk[(hb + b) >> 1] = pack_h2(v[b], v[b + 1]);
        }
        mbar_wait(&in_empty[slot], (use & 1u) ^ 1u);                                // both epilogue sets are done with the slot
        // the burst window: 22 x 30 pixels x 4 frames, 4-byte asynchronous copies, zero outside the image
        {
          const uint32_t sb = smem_u32(s_burst + slot * kBurstBytes);
          for (int idx = pid; idx < kSH * kSW * kTP; idx += 128) {
            const int t = idx & 3, pix = idx >> 2;
            const int r = pix / kSW, c = pix - r * kSW;
            const int gy = y0 - kK / 2 + r, gx = x0 - kK / 2 + c;
            const bool ok = gy >= 0 && gy < p.h && gx >= 0 && gx < p.w;
            const float* src = burst_img + (static_cast<long long>(ok ? gy : 0) * p.w + (ok ? gx : 0)) * p.pitch + p.t0 + t;
            cp_async4_zfill(sb + idx * 4, src, ok ? 4 : 0);
          }
        }
        const uint32_t rowa = smem_u32(s_a + slot * kABytes) + pid * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)                                              // 8 fp16 coefficients per 16-byte chunk
          if (ch * 8 < kused)
            sts128u(rowa + ((ch ^ (pid & 7)) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        cp_async_wait_all();
        fence_proxy_async_smem();                                                   // A is read by the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&in_full[slot]);
      }
    } else if (warp == kMmaWarp) {
      // ================================ MMA issuer ==================================
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(s_a));
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(s_b));
      tc_fence_after();
      for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
        const int slot = it & 1;
        mbar_wait(&in_full[slot], static_cast<uint32_t>(it >> 1) & 1u);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + static_cast<uint32_t>((slot * kABytes) >> 4);
        for (int c = 0; c < kNumChunks; ++c) {
          const int buf = c & 1;                                                    // = the epilogue set that drains it
          const uint32_t use = static_cast<uint32_t>(it * 2 + (c >> 1));            // uses of this buffer so far
          mbar_wait(&acc_empty[buf], (use & 1u) ^ 1u);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * 256);
            const uint32_t idesc = idesc_f16(128, c == kNumChunks - 1 ? kLastN : kChunkN);
            const uint32_t b_lo = b_lo0 + static_cast<uint32_t>((c * kChunkN * 128) >> 4);
            for (int ks = 0; ks < p.ksteps; ++ks)                                    // 16 fp16 = 32 bytes per k-step
              umma_bf16_ss_lo(d_tmem, a_lo + 2 * ks, b_lo + 2 * ks, idesc, ks > 0 ? 1u : 0u);   // kind::f16, formats in idesc
            umma_commit(&acc_full[buf]);
          }
          __syncwarp();
        }
      }
    } else {
      // ================================ epilogue sets: apply the filter ====================
      const int set = warp >> 2;                                                     // 0: chunks 0, 2   1: chunks 1, 3
      const int px = (warp & 3) * 32 + lane;                                         // TMEM lane = pixel of the tile
      const int ry = px >> 4, cx = px & 15;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(set * 256);
      for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
        const int slot = it & 1;
        const uint32_t slot_use = static_cast<uint32_t>(it >> 1);
        const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        const int y0 = ty * kTileH, x0 = tx * kTileW;
        mbar_wait(&in_full[slot], slot_use & 1u);                                    // the burst window of the tile
        const float4* burst_tile = reinterpret_cast<const float4*>(s_burst + slot * kBurstBytes);
        float acc[kTP] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = set + 2 * cc;
          const uint32_t use = static_cast<uint32_t>(it * 2 + cc);
          mbar_wait(&acc_full[set], use & 1u);
          tc_fence_after();
          const float4* win = burst_tile + (ry + c * kRowsPerChunk) * kSW + cx;
          if (c < kNumChunks - 1) apply_chunk<kChunkTaps>(taddr, win, acc);
          else apply_chunk<kLastTaps>(taddr, win, acc);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[set]);
        }
        if (set == 1) {
          s_part[slot * 128 + px] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          __syncwarp();
          if (lane == 0) mbar_arrive(&part_full[slot]);
        } else {
          mbar_wait(&part_full[slot], slot_use & 1u);
          const float4 o1 = s_part[slot * 128 + px];
          acc[0] += o1.x; acc[1] += o1.y; acc[2] += o1.z; acc[3] += o1.w;
          const int y = y0 + ry, x = x0 + cx;
          if (y < p.h && x < p.w) {
            float* o = p.out + ((static_cast<long long>(img) * p.h + y) * p.w + x) * (p.Ttot + 1);
            const float fT = static_cast<float>(p.Ttot);
            float sum = 0.f;
#pragma unroll
            for (int t = 0; t < kTP; ++t) {
              const float v = acc[t] * fT;                                           // Convolve_perlayer, :164
              o[1 + p.t0 + t] = p.accf ? o[1 + p.t0 + t] + v : v;
              sum += acc[t];
            }
            o[0] = p.acc0 ? o[0] + sum : sum;                                        // Convolve = mean of the frames
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&in_empty[slot]);                                 // window / A slot / partial slot free
      }
    }
    // every role has finished the segment: the epilogue's last wait proves that no MMA still reads the basis
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    g = seg_end;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tcf
}  // namespace ie

extern "C" int ie_kpn_apply_tc(const float* burst, int burst_pitch, const float* coef, int hc, int wc, const float* bas,
                               float* out, int n, int h, int w, int T, int K, int B, void* stream) {
  using namespace ie;
  using namespace ie::tcf;
  IE_REQUIRE(burst && coef && bas && out, "kpn_apply_tc: null pointer");
  IE_REQUIRE(n > 0 && h > 0 && w > 0, "kpn_apply_tc: bad sizes");
  IE_REQUIRE(K == kK && T >= kTP && T % kTP == 0 && B >= 1 && B <= 1024,
             "kpn_apply_tc: built for K = 15, T a multiple of 4 (got K=%d T=%d B=%d); use ie_kpn_apply_tf32", K, T, B);
  IE_REQUIRE(burst_pitch >= T && hc >= h && wc >= w, "kpn_apply_tc: bad pitch / coef extent");
  Params p{};
  p.burst = burst; p.coef = coef; p.bas = bas; p.out = out;
  p.n = n; p.h = h; p.w = w; p.hc = hc; p.wc = wc; p.pitch = burst_pitch; p.Ttot = T; p.B = B;
  p.tiles_x = (w + kTileW - 1) / kTileW;
  p.tiles_y = (h + kTileH - 1) / kTileH;
  const long long total = (long long)n * p.tiles_x * p.tiles_y;
  IE_REQUIRE(total < (1ll << 40), "kpn_apply_tc: too many tiles");
  const int grid = total < sm_count() ? (int)total : sm_count();      // persistent: one CTA per SM
  IE_CUDA(cudaFuncSetAttribute(kpn_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  for (int b0 = 0; b0 < B; b0 += kBlockBases) {           // bases in blocks of 64 (one 128-byte K row of fp16)
    p.b0 = b0;
    p.nb = B - b0 < kBlockBases ? B - b0 : kBlockBases;
    p.ksteps = (p.nb + 15) / 16;
    for (int t0 = 0; t0 < T; t0 += kTP) {                 // frames in passes of four
      p.t0 = t0;
      p.accf = b0 > 0;                                    // later basis blocks add to everything,
      p.acc0 = b0 > 0 || t0 > 0;                          // later frame passes to the frame mean out[..., 0]
      IE_CUDA(launch_pdl(kpn_tcgen05_kernel, dim3((unsigned)grid), dim3(kThreads), (size_t)kSmemBytes,
                         static_cast<cudaStream_t>(stream), p));
    }
  }
  return IE_OK;
}
