// dp_latency.cu -- FP64 pipe / shared memory / barrier latencies on one SM (clock64 deltas).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dp_latency dp_latency.cu && ./dp_latency
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k(double *out, long long *cyc, double a, double b, int mode) {
  __shared__ __align__(16) double sm[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = (double)(i & 1) * 8;
  __syncthreads();
  double x0 = a + threadIdx.x, x1 = a * 2, x2 = a * 3, x3 = a * 4, x4 = a * 5, x5 = a * 6, x6 = a * 7, x7 = a * 8;
  long long t0 = clock64();
  if (mode == 0) {
#pragma unroll 16
    for (int i = 0; i < N; i++) x0 = fma(x0, b, a);
  } else if (mode == 1) {
#pragma unroll 4
    for (int i = 0; i < N; i++) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a); }
  } else if (mode == 2) {
#pragma unroll 2
    for (int i = 0; i < N; i++) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
      x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a); }
  } else if (mode == 3) {
#pragma unroll 16
    for (int i = 0; i < N; i++) x0 = x0 + b;
  } else if (mode == 4) {
#pragma unroll 16
    for (int i = 0; i < N; i++) x0 = fmin(x0 * b, a);   // DMUL + min
  } else if (mode == 5) {
    int idx = threadIdx.x & 1;
#pragma unroll 16
    for (int i = 0; i < N; i++) { double2 v = reinterpret_cast<double2 *>(sm)[idx]; idx = (int)v.x; x0 += v.y; }
  } else if (mode == 6) {
#pragma unroll 16
    for (int i = 0; i < N; i++) __syncthreads();
  } else if (mode == 7) {
#pragma unroll 16
    for (int i = 0; i < N; i++) x0 = __shfl_xor_sync(0xffffffffu, x0, 1);
  } else if (mode == 8) {
#pragma unroll 16
    for (int i = 0; i < N; i++) { sm[threadIdx.x] = x0; __syncthreads(); x0 = sm[(threadIdx.x + 1) % blockDim.x] + 1.0; }
  } else if (mode == 9) {
#pragma unroll 16
    for (int i = 0; i < N; i++) x0 = x0 * b;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096 * 8);
  const char *names[] = {"DFMA dependent chain", "DFMA 4 chains (per 4)", "DFMA 8 chains (per 8)", "DADD dependent",
                         "DMUL+fmin dependent", "LDS.128 dependent", "__syncthreads", "shfl64 dependent",
                         "STS+BAR+LDS+DADD round trip", "DMUL dependent"};
  int tpbs[] = {32, 64, 128, 256};
  for (int mode = 0; mode < 10; mode++)
    for (int ti = 0; ti < 4; ti++) {
      int tpb = tpbs[ti];
      long long h = 0;
      k<<<1, tpb>>>(out, cyc, 1.0, 1.0000001, mode);
      k<<<1, tpb>>>(out, cyc, 1.0, 1.0000001, mode);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-32s threads %3d : %.2f cycles/iter\n", names[mode], tpb, (double)h / N);
    }
  return 0;
}
