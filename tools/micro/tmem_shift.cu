// What does tcgen05.shift.cta_group::1.down do?  (PTX ISA 8.6, sm_100a; no offline docs in this image.)
// 128 threads fill 128 TMEM lanes x 32 columns with lane*1000 + col, one thread issues ONE shift at
// (lane_base, col_base), commit -> mbarrier, everybody reads back.  Prints the lane each value came from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_shift tmem_shift.cu && ./tmem_shift
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(int lane_base, int col_base, int nshift, float* out) {
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot;
  const uint32_t mine = base + ((uint32_t)(warp * 32) << 16);
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = __float_as_uint((float)(threadIdx.x * 1000 + c));
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(mine),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]));
  asm volatile("tcgen05.wait::st.sync.aligned;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x == 0) {
    for (int i = 0; i < nshift; ++i)
      asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(base + ((uint32_t)lane_base << 16) + col_base) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
    asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.u32 %0, 1, 0, P;}" : "=r"(ok) : "r"(s32(&bar)));
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(mine));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int c = 0; c < 32; ++c) out[threadIdx.x * 32 + c] = __uint_as_float(r[c]);
  out[128 * 32] = (float)ok;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base));
}
int main() {
  float* d; cudaMalloc(&d, (128 * 32 + 1) * 4);
  static float h[128 * 32 + 1];
  const int cases[][3] = {{0, 0, 1}, {0, 8, 1}, {32, 0, 1}, {0, 0, 2}, {0, 4, 1}};
  for (auto& cs : cases) {
    k<<<1, 128>>>(cs[0], cs[1], cs[2], d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("== lane_base %d col_base %d shifts %d: %s barrier_ok=%d\n", cs[0], cs[1], cs[2], cudaGetErrorString(e), (int)h[128 * 32]);
    // which columns changed, and for a changed column: source lane per lane
    int changed[32] = {0};
    for (int l = 0; l < 128; ++l) for (int c = 0; c < 32; ++c) if (h[l * 32 + c] != (float)(l * 1000 + c)) changed[c]++;
    printf("   columns changed (count of lanes):"); for (int c = 0; c < 32; ++c) if (changed[c]) printf(" c%d:%d", c, changed[c]); printf("\n");
    for (int c = 0; c < 32; ++c) if (changed[c]) {
      printf("   col %d source lane of lanes 0..5, 30..35, 62..66, 94..98, 125..127:", c);
      const int ls[] = {0,1,2,3,4,5,30,31,32,33,34,35,62,63,64,65,66,94,95,96,97,98,125,126,127};
      for (int l : ls) { float v = h[l * 32 + c]; printf(" %d<-%d", l, (int)((v - c) / 1000 + 0.5f)); }
      printf("\n"); break;
    }
  }
  return 0;
}
