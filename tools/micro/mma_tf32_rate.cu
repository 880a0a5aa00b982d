// Microbenchmark: issue rate of the legacy warp-level mma.sync (m16n8k8, tf32 in, fp32 accumulate) on sm_100a,
// register operands only.  Prints MMAs per cycle per SM and TFLOP/s.  nvcc -arch=sm_100a -O3 -o mma_tf32_rate ...
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, long long* cycles) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 20000;
    k<<<148, warps * 32>>>(out, 100, cyc); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<148, warps * 32>>>(out, iters, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double mmas = (double)warps * iters * 8;                 // per SM
    printf("warps/SM %2d: %.3f MMA/clk/SM (%.1f cycles per MMA per SMSP), %.1f TFLOP/s tf32, %.3f ms\n", warps, mmas / c,
           4.0 * c / mmas, 148.0 * mmas * 2 * 16 * 8 * 8 / (ms * 1e-3) / 1e12, ms);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
