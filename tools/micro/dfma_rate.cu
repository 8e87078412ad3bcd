// FP64 vector-pipe rates on one SM (B200): cycles per warp-instruction of DFMA / DADD / DMNMX per SM sub-partition for
// 1, 2, 4 warps per sub-partition, with 8 independent chains per thread (throughput) and one chain (dependent latency).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/dfma_rate tools/micro/dfma_rate.cu && tools/micro/dfma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int CH>
__global__ void k(double *out, long long *clk, int n, double a, double b) {
  double v[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) v[c] = threadIdx.x * 1e-3 + c;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) {
      if (OP == 0) v[c] = fma(v[c], a, b);
      else if (OP == 1) v[c] = v[c] + a;
      else v[c] = fmin(fmax(v[c], a), b + v[c]);      // two DMNMX and a DADD
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP, int CH>
void run(const char *name, int threads) {
  double *out; long long *clk, h;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&clk, 8);
  const int n = 4096;
  k<OP, CH><<<1, threads>>>(out, clk, n, 0.999, 1e-3);
  k<OP, CH><<<1, threads>>>(out, clk, n, 0.999, 1e-3);
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double warps_per_smsp = threads / 32 / 4.0;
  const double instr = (double)n * CH * (OP == 2 ? 3 : 1) * (warps_per_smsp < 1 ? 1 : warps_per_smsp);
  printf("%-28s %4d threads, %d chains: %8lld clk, %.2f clk per warp-instruction per sub-partition\n", name, threads, CH, h, h / instr);
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int th : {32, 128, 256, 512, 1024}) run<0, 8>("DFMA throughput", th);
  run<0, 1>("DFMA dependent chain", 32);
  for (int th : {128, 256, 512}) run<1, 8>("DADD throughput", th);
  for (int th : {128, 256, 512}) run<2, 8>("DMNMX+DMNMX+DADD", th);
  run<2, 1>("DMNMX chain", 32);
  return 0;
}
