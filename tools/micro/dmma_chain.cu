// Microbenchmark: FP64 mma.sync m8n8k4 issue rate per SM sub-partition as a function of the number of warps per
// sub-partition and of independent accumulator chains per warp (sm_100a), with and without an LDS feeding each mma.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int CH, bool LDSFEED>
__global__ void k(double *out, int iters, double a0, double b0) {
  __shared__ double sm[8 * 32 * 4];
  for (int i = threadIdx.x; i < 8 * 32 * 4; i += blockDim.x) sm[i] = a0 + i * 1e-9;
  __syncthreads();
  double c[CH][2];
  for (int i = 0; i < CH; i++) { c[i][0] = threadIdx.x; c[i][1] = 1.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  const double *sp = sm + (threadIdx.x & 31);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8 / CH; r++)
#pragma unroll
      for (int i = 0; i < CH; i++) {
        if (LDSFEED) a = sp[((it + r) & 3) * 256 + i * 32];
        dmma(c[i], a, b);
      }
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}
template <int CH, bool L>
void run(double *d, int warps) {
  int iters = 2000; double clk;
  k<CH, L><<<148, warps * 32>>>(d, iters, 1.0000001, 0.5); cudaDeviceSynchronize();
  cudaMemcpy(&clk, d + 148 * warps * 32, 8, cudaMemcpyDeviceToHost);
  printf("chains=%d lds=%d warps/SM=%2d: %.2f clk per DMMA per warp, %.2f clk per DMMA per sub-partition\n", CH, (int)L, warps, clk / (iters * 8.0),
         clk / (iters * 8.0 * warps / 4.0));
}
int main() {
  double *d; cudaMalloc(&d, (148 * 1024 + 8) * sizeof(double));
  for (int warps : {4, 8, 12, 16}) {
    run<1, false>(d, warps); run<2, false>(d, warps); run<4, false>(d, warps); run<8, false>(d, warps);
    run<8, true>(d, warps); run<4, true>(d, warps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
