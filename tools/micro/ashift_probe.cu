// What do `tcgen05.mma ... .ashift` (A operand in tensor memory) and `tcgen05.cp.128x256b` do?  (No offline PTX docs
// in this image; ptxas accepts both for sm_100a: SASS UTCHMMA.ASHIFT / UTCCP.)  One CTA of 128 threads:
//   A[r][k] = r for even k, k for odd k (bf16, exact), B[n][k] = (n == k): D[r][n] = A[r][n].
//   1. A written to TMEM with tcgen05.st (row r in lane r, column c = {A[r][2c], A[r][2c+1]}), 4 MMAs (K = 64)
//      plain -> D0, .ashift -> D1, plain -> D2, .ashift -> D3: direction of the shift, whether it persists in TMEM;
//   2. the same A as a 128B-swizzled K-major smem tile copied with 4 x tcgen05.cp.128x256b -> raw dump next to the
//      tcgen05.st image.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o ashift_probe ashift_probe.cu && ./ashift_probe
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float aval(int r, int k) { return (k & 1) ? (float)k : (float)r; }
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t id, uint32_t acc, bool shift) {
  if (shift)
    asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 db, {%2, %5};\n"
                 "tcgen05.mma.cta_group::1.kind::f16.ashift [%0], [%1], db, %3, p;\n}" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(id),
                 "r"(acc), "r"(kDescHi) : "memory");
  else
    asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 db, {%2, %5};\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n}" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(id),
                 "r"(acc), "r"(kDescHi) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
    asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0, 1, 0, P;}"
                 : "=r"(ok) : "r"(s32(bar)), "r"(parity));
  asm volatile("tcgen05.fence::after_thread_sync;");
  return ok;
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
}

// out: D[4][128][64] floats, then Araw[128][32], then Craw[128][32] (u32 as float bits), then flags
__global__ void __launch_bounds__(128) probe(float* out, int variant) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;                  // 128 rows x 128 B, 128B swizzle
  uint8_t* sB = base + 16384;          // 64 rows x 128 B
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, t = threadIdx.x;
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // smem tiles (what TMA with SWIZZLE_128B would have written): 16-byte chunk j of row r at chunk j ^ (r & 7)
  for (int j = 0; j < 8; ++j) {
    uint4 v;
    v.x = pack(aval(t, 8 * j + 0), aval(t, 8 * j + 1));
    v.y = pack(aval(t, 8 * j + 2), aval(t, 8 * j + 3));
    v.z = pack(aval(t, 8 * j + 4), aval(t, 8 * j + 5));
    v.w = pack(aval(t, 8 * j + 6), aval(t, 8 * j + 7));
    *reinterpret_cast<uint4*>(sA + t * 128 + ((j ^ (t & 7)) << 4)) = v;
    if (t < 64) {
      uint4 w;
      w.x = pack(t == 8 * j + 0, t == 8 * j + 1);
      w.y = pack(t == 8 * j + 2, t == 8 * j + 3);
      w.z = pack(t == 8 * j + 4, t == 8 * j + 5);
      w.w = pack(t == 8 * j + 6, t == 8 * j + 7);
      *reinterpret_cast<uint4*>(sB + t * 128 + ((j ^ (t & 7)) << 4)) = w;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = slot;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  constexpr uint32_t colA = 256, colC = 320;
  {
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = pack(aval(t, 2 * c), aval(t, 2 * c + 1));
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(tb + lane_off + colA),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]));
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t a_lo = ((s32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t b_lo = ((s32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
  if (t == 0) {
    const uint32_t id = idesc(128, 64);
    // variant 0: plain, ashift, plain, ashift.  variant 1: ashift only on the FIRST k-step of D1 / D3 (is the shift
    // per instruction = per 8-column slice?)
    for (int m = 0; m < 4; ++m) {
      for (int k = 0; k < 4; ++k) {
        bool sh = (m & 1);
        if (variant == 1 && k != 0) sh = false;
        mma_ts(tb + 64 * m, tb + colA + 8 * k, b_lo + 2 * k, id, k ? 1u : 0u, sh);
      }
    }
    // smem -> TMEM copy of the swizzled tile, one 128 x 32-byte slice per K step
    for (int k = 0; k < 4; ++k) {
      asm volatile("{\n.reg .b64 da;\nmov.b64 da, {%1, %2};\ntcgen05.cp.cta_group::1.128x256b [%0], da;\n}" ::"r"(tb + colC + 8 * k),
                   "r"(a_lo + 2 * k), "r"(kDescHi) : "memory");
    }
    commit(&bar);
  }
  const uint32_t ok = wait(&bar, 0);
  uint32_t r[32];
  for (int m = 0; m < 4; ++m)
    for (int h = 0; h < 2; ++h) {
      ld32(tb + lane_off + 64 * m + 32 * h, r);
      for (int c = 0; c < 32; ++c) out[(m * 128 + t) * 64 + 32 * h + c] = __uint_as_float(r[c]);
    }
  ld32(tb + lane_off + colA, r);
  for (int c = 0; c < 32; ++c) out[4 * 128 * 64 + t * 32 + c] = __uint_as_float(r[c]);
  ld32(tb + lane_off + colC, r);
  for (int c = 0; c < 32; ++c) out[4 * 128 * 64 + 128 * 32 + t * 32 + c] = __uint_as_float(r[c]);
  if (t == 0) out[4 * 128 * 64 + 2 * 128 * 32] = (float)ok;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

// Issue rate: `reps` x 36 MMAs per commit, as a tile of the shift-convolution would (3 filter rows x 4 K steps x 3 taps,
// N = 64, A in TMEM, taps 0 and 1 with .ashift), optionally preceded by the 12 tcgen05.cp of the tile; against
// 12 MMAs of N = 192 with A in shared memory (the wide-N kernel's tile).
__global__ void __launch_bounds__(128) rate(long long* cycles, int mode, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;                  // 3 x 16 KB
  uint8_t* sB = base + 49152;          // 72 KB of weights
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, t = threadIdx.x;
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = t; i < (49152 + 73728) / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = slot;
  const uint32_t a_lo = ((s32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t b_lo = ((s32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
  long long t0 = 0;
  if (t == 0) {
    t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      const uint32_t d = tb + 64 * (rep & 1);
      const uint32_t areg = tb + 256 + 96 * (rep & 1);
      if (mode == 0) {                 // wide-N: 12 MMAs N = 192, A and B in smem
        const uint32_t id = idesc(128, 192);
        const uint32_t dd = tb + 192 * (rep & 1);
        for (int dy = 0; dy < 3; ++dy)
          for (int k = 0; k < 4; ++k)
            asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}" ::"r"(dd), "r"(a_lo + dy * 1024 + 2 * k),
                         "r"(b_lo + dy * 1536 + 2 * k), "r"(id), "r"((dy | k) ? 1u : 0u), "r"(kDescHi) : "memory");
      } else {
        const uint32_t id = idesc(128, 64);
        if (mode >= 2)
          for (int dy = 0; dy < 3; ++dy)
            for (int k = 0; k < 4; ++k)
              asm volatile("{\n.reg .b64 da;\nmov.b64 da, {%1, %2};\ntcgen05.cp.cta_group::1.128x256b [%0], da;\n}" ::"r"(areg + 32 * dy + 8 * k),
                           "r"(a_lo + dy * 1024 + 2 * k), "r"(kDescHi) : "memory");
        for (int dy = 0; dy < 3; ++dy)
          for (int k = 0; k < 4; ++k)
            for (int dx = 0; dx < 3; ++dx)
              mma_ts(d, areg + 32 * dy + 8 * k, b_lo + dy * 1536 + dx * 512 + 2 * k, id, (dy | k | dx) ? 1u : 0u,
                     mode != 3 && dx < 2);
      }
    }
    commit(&bar);
  }
  wait(&bar, 0);
  if (t == 0) cycles[0] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

int main() {
  const int n = 4 * 128 * 64 + 2 * 128 * 32 + 1;
  float* d;
  cudaMalloc(&d, n * 4);
  static float h[n];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  const int rows[] = {0, 1, 2, 3, 30, 31, 32, 33, 34, 62, 63, 64, 65, 95, 96, 97, 126, 127};
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(d, 0, n * 4);
    probe<<<1, 128, 16384 + 8192 + 1024>>>(d, variant);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, n * 4, cudaMemcpyDeviceToHost);
    printf("== variant %d: %s, barrier_ok=%d\n", variant, cudaGetErrorString(e), (int)h[n - 1]);
    const char* names[] = {"D0 plain", "D1 ashift", "D2 plain again", "D3 ashift again"};
    for (int m = 0; m < 4; ++m) {
      printf("%-16s row: D[row][0] (= source row), D[row][1] (expect 1), D[row][16] (k-step 1), D[row][62]\n", names[m]);
      for (int rr : rows) {
        const float* p = h + (m * 128 + rr) * 64;
        printf("   r%-3d: %6.1f %6.1f %6.1f %6.1f\n", rr, p[0], p[1], p[16], p[62]);
      }
    }
    const uint32_t* A = reinterpret_cast<const uint32_t*>(h + 4 * 128 * 64);
    const uint32_t* C = A + 128 * 32;
    int same = 0, diff = 0;
    for (int i = 0; i < 128 * 32; ++i) (A[i] == C[i]) ? ++same : ++diff;
    printf("A region after the MMAs vs tcgen05.cp image: %d equal, %d different words\n", same, diff);
    for (int rr : {0, 1, 2, 31, 32, 33, 127})
      printf("   lane %3d: st-image %08x %08x %08x .. %08x | cp-image %08x %08x %08x .. %08x\n", rr, A[rr * 32], A[rr * 32 + 1],
             A[rr * 32 + 8], A[rr * 32 + 31], C[rr * 32], C[rr * 32 + 1], C[rr * 32 + 8], C[rr * 32 + 31]);
  }
  long long* dc;
  cudaMalloc(&dc, 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 73728 + 1024);
  const char* mn[] = {"wide-N: 12 MMAs N=192, A in smem", "36 MMAs N=64, A in TMEM, .ashift on taps 0/1", "12 tcgen05.cp + the same 36 MMAs",
                      "12 tcgen05.cp + 36 MMAs without .ashift"};
  for (int mode = 0; mode < 4; ++mode)
    for (int reps : {64, 256}) {
      rate<<<1, 128, 49152 + 73728 + 1024>>>(dc, mode, reps);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0;
      cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
      printf("rate mode %d (%s), %d tiles: %s, %.1f cycles per tile\n", mode, mn[mode], reps, cudaGetErrorString(e), (double)c / reps);
    }
  return 0;
}
