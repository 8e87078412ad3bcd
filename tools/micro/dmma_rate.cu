// Microbenchmark: FP64 tensor (mma.sync m8n8k4) vs vector DFMA issue rate per SM on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__global__ void k_dmma(double *out, int iters, double a0, double b0) {
  double c[8][2];
  for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = 1.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma(c[i], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}
__global__ void k_dfma(double *out, int iters, double a0, double b0) {
  double c[16];
  for (int i = 0; i < 16; i++) c[i] = threadIdx.x + i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = fma(a, c[i], b);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 16; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}
int main() {
  double *d; cudaMalloc(&d, (148 * 1024 + 8) * sizeof(double));
  for (int warps : {1, 4, 8, 16}) {
    int iters = 2000; double clk;
    k_dmma<<<148, warps * 32>>>(d, iters, 1.0000001, 0.5); cudaDeviceSynchronize();
    cudaMemcpy(&clk, d + 148 * warps * 32, 8, cudaMemcpyDeviceToHost);
    printf("DMMA warps/SM=%2d: %.2f clk per DMMA per SM  -> %.1f FMA/clk/SM\n", warps, clk / (iters * 8.0 * warps), 256.0 * iters * 8 * warps / clk);
    k_dfma<<<148, warps * 32>>>(d, iters, 1.0000001, 0.5); cudaDeviceSynchronize();
    cudaMemcpy(&clk, d + 148 * warps * 32, 8, cudaMemcpyDeviceToHost);
    printf("DFMA warps/SM=%2d: %.2f clk per warp-DFMA per SM -> %.1f FMA/clk/SM\n", warps, clk / (iters * 16.0 * warps), 32.0 * iters * 16 * warps / clk);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
