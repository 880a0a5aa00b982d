// PROTOTYPE v2 (not part of libimgenh_b200.so yet): tools/micro/kpn_tcgen05.cu with the pipeline its measurement asked
// for.  Ran on a B200 at the very end of round 1: all cases PASS, 0.656 ms at cfg2 (shipped mma.sync kernel: 0.91 ms).  v1 (correct, 2.15 ms at
// cfg2) spends ~20 000 cycles per 128-px tile staging the coefficients and the burst window with dependent scalar loads
// on 5 warps, then computes, then synchronises.  Here:
//   * warps 9-12 are PRODUCERS: they stage tile i+1 (coefficients -> TF32, hand-swizzled A operand; burst window by
//     4-byte cp.async with zero fill) into the second A / window slot while tile i is being filtered;
//   * warps 0-3 and 4-7 are two EPILOGUE sets: set s drains accumulator buffer s (chunks s, s+2 of the tile), so two
//     warps per scheduler issue the 900 FMAs per pixel; set 1 hands its partial sums to set 0 through shared memory;
//   * tcgen05.ld of the next 16 columns is issued before the current 16 are consumed;
//   * no __syncthreads in the tile loop: mbarriers in_full / in_empty (A + window slots), acc_full / acc_empty (TMEM
//     buffers), part_full (partial sums).
// Same scope and the same self-check as v1 (K = 15, frames in passes of 4, B <= 32).
// The maintained copy of this kernel is imageenhancement_mp_b200/csrc/kpn_tcgen05.cu (ie_kpn_apply_tc, opt-in); this file
// stays as the standalone artefact behind profiles/r01_kpn_tcgen05_v2_prototype.txt.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o kpn_tcgen05_v2 kpn_tcgen05_v2.cu && ./kpn_tcgen05_v2
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../imageenhancement_mp_b200/csrc/ie_ptx.cuh"

using namespace ie;

namespace {

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

__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
  // c_format f32 (1 << 4), a_format / b_format TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at 24
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ss_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHi)
      : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// byte offset of element k (fp32 slot) of row r inside a K-major SWIZZLE_128B operand: 128-byte rows, the 16-byte
// chunk index XOR-ed with the row's position in its 8-row group (what TMA writes; csrc/conv_tcgen05.cu builds its
// first-layer A tile the same way)
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {
  return static_cast<uint32_t>(r * 128 + ((((k >> 2) ^ (r & 7)) << 4) | ((k & 3) << 2)));
}

struct Params {
  const float* burst;
  const float* coef;
  const float* bas;
  float* out;
  int h, w, pitch, Ttot, B, t0, accumulate;
  int tiles_x, tiles_y, strips, ksteps;
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
  const int img = blockIdx.x / p.strips, strip = blockIdx.x - img * p.strips;
  const int kused = p.ksteps * 8;

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
  // ---- the image's basis -> B operand (once per CTA): Bas[img][tap][t0 + t][b] -> row tap * 4 + t, slot b, TF32
  {
    const float* bas_img = p.bas + static_cast<long long>(img) * kTaps * p.Ttot * p.B;
    for (int idx = threadIdx.x; idx < kBRows * kused; idx += kThreads) {
      const int n = idx / kused, b = idx - n * kused;
      float v = 0.f;
      if (n < kTaps * kTP && b < p.B) {
        const int tap = n >> 2, t = n & 3;
        v = to_tf32(__ldg(bas_img + (static_cast<long long>(tap) * p.Ttot + p.t0 + t) * p.B + b));
      }
      *reinterpret_cast<float*>(s_b + sw128_off(n, b)) = v;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles = p.tiles_x * p.tiles_y;

  if (warp >= kFirstProducerWarp) {
    // ================================ producers: stage tile `it` into slot it & 1 ==================
    const int pid = threadIdx.x - kFirstProducerWarp * 32;                        // 0..127 = pixel of the tile
    const float* burst_img = p.burst + static_cast<long long>(img) * p.h * p.w * p.pitch;
    int it = 0;
    for (int tile = strip; tile < tiles; tile += p.strips, ++it) {
      const int slot = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
      const int y0 = ty * kTileH, x0 = tx * kTileW;
      // this thread's pixel: its coefficients first into registers (all loads in flight), rounded to TF32
      const int y = y0 + (pid >> 4), x = x0 + (pid & 15);
      const bool inside = y < p.h && x < p.w;
      const float* cp = p.coef + ((static_cast<long long>(img) * p.h + (inside ? y : 0)) * p.w + (inside ? x : 0)) * p.B;
      float v[32];
#pragma unroll
      for (int b = 0; b < 32; ++b) v[b] = (b < kused && b < p.B && inside) ? __ldg(cp + b) : 0.f;
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
      for (int ch = 0; ch < 8; ++ch)
        if (ch * 4 < kused)
          sts128(rowa + ((ch ^ (pid & 7)) << 4), to_tf32(v[4 * ch]), to_tf32(v[4 * ch + 1]), to_tf32(v[4 * ch + 2]),
                 to_tf32(v[4 * ch + 3]));
      cp_async_wait_all();
      fence_proxy_async_smem();                                                   // A is read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&in_full[slot]);
    }
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer ==================================
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(s_a));
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(s_b));
    int it = 0;
    for (int tile = strip; tile < tiles; tile += p.strips, ++it) {
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
          const uint32_t idesc = idesc_tf32(128, c == kNumChunks - 1 ? kLastN : kChunkN);
          const uint32_t b_lo = b_lo0 + static_cast<uint32_t>((c * kChunkN * 128) >> 4);
          for (int ks = 0; ks < p.ksteps; ++ks)                                    // 8 TF32 = 32 bytes per k-step
            umma_tf32_ss_lo(d_tmem, a_lo + 2 * ks, b_lo + 2 * ks, idesc, ks > 0 ? 1u : 0u);
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
    int it = 0;
    for (int tile = strip; tile < tiles; tile += p.strips, ++it) {
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
            o[1 + p.t0 + t] = acc[t] * fT;                                         // Convolve_perlayer, :164
            sum += acc[t];
          }
          o[0] = p.accumulate ? o[0] + sum : sum;                                  // Convolve = mean of the frames
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&in_empty[slot]);                                 // window / A slot / partial slot free
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

// naive fp32 reference: one thread per pixel, the algebra of model_library.py:439-451 spelled out
__global__ void kpn_naive_kernel(const float* burst, const float* coef, const float* bas, float* out, int n, int h, int w,
                                 int pitch, int T, int B) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(n) * h * w) return;
  const int x = static_cast<int>(i % w), y = static_cast<int>((i / w) % h), img = static_cast<int>(i / (static_cast<long long>(w) * h));
  const float* cf = coef + i * B;
  float sum = 0.f;
  for (int t = 0; t < T; ++t) {
    float a = 0.f;
    for (int ti = 0; ti < kK; ++ti)
      for (int tj = 0; tj < kK; ++tj) {
        const int gy = y + ti - kK / 2, gx = x + tj - kK / 2;
        if (gy < 0 || gy >= h || gx < 0 || gx >= w) continue;
        const float bv = burst[((static_cast<long long>(img) * h + gy) * w + gx) * pitch + t];
        const float* bb = bas + ((static_cast<long long>(img) * kTaps + ti * kK + tj) * T + t) * B;
        float f = 0.f;
        for (int b = 0; b < B; ++b) f = fmaf(cf[b], bb[b], f);
        a = fmaf(f, bv, a);
      }
    out[i * (T + 1) + 1 + t] = a * T;
    sum += a;
  }
  out[i * (T + 1)] = sum;
}

int launch(const float* burst, int pitch, const float* coef, const float* bas, float* out, int n, int h, int w, int T, int B,
           cudaStream_t st) {
  if (T % kTP != 0 || B < 1 || B > 32) return -1;
  Params p{};
  p.burst = burst; p.coef = coef; p.bas = bas; p.out = out;
  p.h = h; p.w = w; p.pitch = pitch; p.Ttot = T; p.B = B;
  p.tiles_x = (w + kTileW - 1) / kTileW;
  p.tiles_y = (h + kTileH - 1) / kTileH;
  p.ksteps = (B + 7) / 8;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int tiles = p.tiles_x * p.tiles_y;
  int strips = (sms + n - 1) / n;                          // ~one CTA per SM; every CTA stages one image's basis once
  if (strips > tiles) strips = tiles;
  if (strips < 1) strips = 1;
  p.strips = strips;
  cudaFuncSetAttribute(kpn_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  for (int t0 = 0; t0 < T; t0 += kTP) {
    p.t0 = t0;
    p.accumulate = t0 > 0;
    kpn_tcgen05_kernel<<<n * strips, kThreads, kSmemBytes, st>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

float frand(uint64_t& s) {
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return static_cast<float>((s >> 40) & 0xFFFFFF) / 16777216.f;
}

int run_case(int n, int h, int w, int T, int B, bool timing) {
  const int pitch = T + 1;
  const size_t npx = static_cast<size_t>(n) * h * w;
  std::vector<float> hb(npx * pitch), hc(npx * B), hbas(static_cast<size_t>(n) * kTaps * T * B);
  uint64_t s = 12345 + n * 7 + h * 3 + T * 11 + B;
  for (auto& v : hb) v = frand(s);
  for (size_t i = 0; i < npx; ++i) {                       // coefficients: a softmax over the bases
    float tot = 0.f;
    for (int b = 0; b < B; ++b) tot += (hc[i * B + b] = expf(3.f * frand(s)));
    for (int b = 0; b < B; ++b) hc[i * B + b] /= tot;
  }
  for (int im = 0; im < n; ++im)                           // basis: a softmax over the taps (and frames) per basis
    for (int b = 0; b < B; ++b) {
      float tot = 0.f;
      for (int k = 0; k < kTaps * T; ++k) tot += (hbas[(static_cast<size_t>(im) * kTaps * T + k) * B + b] = expf(4.f * frand(s)));
      for (int k = 0; k < kTaps * T; ++k) hbas[(static_cast<size_t>(im) * kTaps * T + k) * B + b] /= tot;
    }
  float *db, *dc, *dbas, *dout, *dref;
  cudaMalloc(&db, hb.size() * 4); cudaMalloc(&dc, hc.size() * 4); cudaMalloc(&dbas, hbas.size() * 4);
  cudaMalloc(&dout, npx * (T + 1) * 4); cudaMalloc(&dref, npx * (T + 1) * 4);
  cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dc, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dbas, hbas.data(), hbas.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, npx * (T + 1) * 4);
  int rc = launch(db, pitch, dc, dbas, dout, n, h, w, T, B, 0);
  kpn_naive_kernel<<<static_cast<unsigned>((npx + 127) / 128), 128>>>(db, dc, dbas, dref, n, h, w, pitch, T, B);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc || e != cudaSuccess) {
    printf("case n=%d %dx%d T=%d B=%d: launch failed (rc %d, %s)\n", n, h, w, T, B, rc, cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> o(npx * (T + 1)), r(npx * (T + 1));
  cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(r.data(), dref, r.size() * 4, cudaMemcpyDeviceToHost);
  double e0 = 0, e1 = 0;
  for (size_t i = 0; i < npx; ++i)
    for (int c = 0; c <= T; ++c) {
      const double d = fabs(static_cast<double>(o[i * (T + 1) + c]) - r[i * (T + 1) + c]);
      if (!(d == d)) e0 = e1 = 1e30;
      if (c == 0) e0 = d > e0 ? d : e0; else e1 = d > e1 ? d : e1;
    }
  const double bound = ldexp(1.0, -10) * 1.05;
  const bool ok = e0 <= bound && e1 <= T * bound;
  printf("case n=%d %dx%d T=%d B=%d: max-abs diff channel 0 %.3g, frames %.3g (bound %.3g / %.3g) %s\n", n, h, w, T, B, e0, e1,
         bound, T * bound, ok ? "PASS" : "FAIL");
  if (timing && ok) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch(db, pitch, dc, dbas, dout, n, h, w, T, B, 0);
    cudaEventRecord(a);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) launch(db, pitch, dc, dbas, dout, n, h, w, T, B, 0);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    printf("  %.4f ms per call, %.1f MP/s (%.3f ms per megapixel)\n", ms / reps, npx / 1e6 / (ms / reps / 1e3),
           ms / reps / (npx / 1e6));
  }
  cudaFree(db); cudaFree(dc); cudaFree(dbas); cudaFree(dout); cudaFree(dref);
  return ok ? 0 : 1;
}

}  // namespace

int main() {
  int bad = 0;
  bad += run_case(1, 8, 16, 4, 10, false);                 // exactly one tile
  bad += run_case(2, 19, 37, 4, 10, false);                // ragged edges, several tiles per CTA
  bad += run_case(1, 40, 72, 8, 10, false);                // two frame passes
  bad += run_case(3, 33, 47, 4, 32, false);                // four k-steps
  bad += run_case(256, 104, 104, 4, 10, true);             // BASELINE configs[1]
  bad += run_case(16, 256, 256, 8, 10, true);              // 1 MP at T = 8 (tools/kpn_bases_bench.py's case)
  printf(bad ? "FAILED\n" : "ALL PASS\n");
  return bad;
}
