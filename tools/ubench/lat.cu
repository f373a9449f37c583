// Latency microbenchmarks behind the diagonal-block (potf2) design: dependent-chain cycles per op on one warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double* out, long long* cyc, double x0, int n) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = x0;
  __syncthreads();
  double x = x0 + threadIdx.x * 1e-9, y = 0.999;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    if (OP == 0) x = fma(x, y, 1e-3);
    if (OP == 1) x = rsqrt(x) + 1.0;
    if (OP == 2) x = sqrt(x) + 1.0;
    if (OP == 3) x = 1.0 / x + 1.0;
    if (OP == 4) { __syncthreads(); }
    if (OP == 5) x = sm[((int)x) & 63] + 1.0;
    if (OP == 6) x = (double)rsqrtf((float)x) + 1.0;
    if (OP == 7) x += __shfl_xor_sync(0xffffffffu, x, 1);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[threadIdx.x] = x;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  const char* names[] = {"dfma", "rsqrt(double)", "sqrt(double)", "1/x double", "syncthreads", "lds dependent (+cvt)", "rsqrtf via float", "shfl double"};
  const int n = 2000;
  for (int threads : {32, 128, 1024}) {
    for (int op = 0; op < 8; op++) {
      for (int rep = 0; rep < 2; rep++) {
        switch (op) {
          case 0: chain<0><<<1, threads>>>(out, cyc, 1.5, n); break;
          case 1: chain<1><<<1, threads>>>(out, cyc, 1.5, n); break;
          case 2: chain<2><<<1, threads>>>(out, cyc, 1.5, n); break;
          case 3: chain<3><<<1, threads>>>(out, cyc, 1.5, n); break;
          case 4: chain<4><<<1, threads>>>(out, cyc, 1.5, n); break;
          case 5: chain<5><<<1, threads>>>(out, cyc, 1.5, n); break;
          case 6: chain<6><<<1, threads>>>(out, cyc, 1.5, n); break;
          case 7: chain<7><<<1, threads>>>(out, cyc, 1.5, n); break;
        }
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("threads=%4d %-22s %7.1f cycles/op\n", threads, names[op], (double)c / n);
    }
  }
  return 0;
}
