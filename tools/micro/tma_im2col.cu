// How does TMA im2col mode (cuTensorMapEncodeIm2col + cp.async.bulk.tensor.4d...im2col) address pixels?
// (no offline PTX docs in this image: measured.)  Tensor [N][H][W][C = 64] bf16, element = a code of (n, h, w) in
// channel 0..; 3x3 'same' convolution: lower corner (-1,-1), upper corner (-1,-1), channelsPerPixel 64,
// pixelsPerColumn = ROWS.  One thread issues ONE load with coordinates {c, w, h, n} and offsets {ow, oh}; the block
// prints which (n, h, w) landed in each of the ROWS smem rows (or "zero").
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_im2col tma_im2col.cu -lcuda && ./tma_im2col
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int ROWS = 32;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ CUtensorMap tm, int c, int w, int h, int n, int ow, int oh, float* out) {
  __shared__ __align__(1024) __nv_bfloat16 tile[ROWS * 64];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = threadIdx.x; i < ROWS * 64; i += blockDim.x) tile[i] = __float2bfloat16(-7.f);
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(ROWS * 128));
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2], {%7, %8};" ::"r"(s32(tile)),
        "l"(&tm), "r"(s32(&bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)ow), "h"((uint16_t)oh)
        : "memory");
  }
  // wait (bounded)
  uint32_t ok = 0;
  for (int it = 0; it < 2000000 && !ok; ++it)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(s32(&bar)) : "memory");
  __syncthreads();
  if (threadIdx.x < ROWS) {
    out[threadIdx.x * 2] = ok ? __bfloat162float(tile[threadIdx.x * 64]) : -9.f;       // channel 0: code
    out[threadIdx.x * 2 + 1] = __bfloat162float(tile[threadIdx.x * 64 + 63]);           // channel 63: same code + 0.5 marker? (just non-zero)
  }
}

int main() {
  const int N = 2, H = 5, W = 6, C = 64;
  std::vector<__nv_bfloat16> hx((size_t)N * H * W * C);
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w)
        for (int c = 0; c < C; ++c) hx[(((size_t)n * H + h) * W + w) * C + c] = __float2bfloat16((float)(n * 100 + h * 10 + w + 1));
  __nv_bfloat16* dx;
  cudaMalloc(&dx, hx.size() * 2);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  float* dout;
  cudaMalloc(&dout, ROWS * 2 * 4);
  CUtensorMap tm;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lo[2] = {-1, -1}, hi[2] = {-1, -1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  cuInit(0);
  CUresult r = cuTensorMapEncodeIm2col(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, gdim, gstr, lo, hi, 64, ROWS, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  struct Case { int c, w, h, n, ow, oh; const char* what; };
  Case cases[] = {
      {0, -1, -1, 0, 0, 0, "base (-1,-1), offsets (0,0): tap (dy=-1,dx=-1) of output pixels 0.."},
      {0, -1, -1, 0, 1, 1, "base (-1,-1), offsets (1,1): centre tap -> should be pixels (0,0),(0,1).."},
      {0, -1, -1, 0, 2, 2, "base (-1,-1), offsets (2,2)"},
      {0, 2, 1, 0, 1, 1, "base (w=2,h=1), offsets (1,1): start mid-row, wraps rows and images?"},
      {0, 0, 0, 0, 0, 0, "base (0,0) offsets (0,0)"},
      {0, 3, 3, 1, 2, 0, "base (3,3) in image 1, offsets (ow=2, oh=0): runs off the end of the tensor"},
  };
  for (auto& cs : cases) {
    k<<<1, 128>>>(tm, cs.c, cs.w, cs.h, cs.n, cs.ow, cs.oh, dout);
    cudaError_t e = cudaDeviceSynchronize();
    float ho[ROWS * 2];
    cudaMemcpy(ho, dout, sizeof(ho), cudaMemcpyDeviceToHost);
    printf("%s  [%s]\n  rows:", cs.what, cudaGetErrorString(e));
    for (int i = 0; i < ROWS; ++i) {
      int v = (int)ho[2 * i];
      if (v == 0) printf(" zero");
      else if (v < 0) printf(" (%d)", v);
      else printf(" n%dh%dw%d", (v - 1) / 100, ((v - 1) % 100) / 10, (v - 1) % 10);
    }
    printf("\n");
  }
  return 0;
}
