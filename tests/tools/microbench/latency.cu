// Dependent-issue latencies on one warp (B200): DFMA, DADD, DMUL, MUFU.RSQ64H-based rsqrt, double divide, SHFL (64-bit =
// two 32-bit shuffles), LDS.64 / LDS.128 pointer chase, __syncwarp.  Build: nvcc -arch=sm_100a -O3 latency.cu -o latency
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[1024];
  const int lane = threadIdx.x;
  for (int i = lane; i < 1024; i += 32) sm[i] = (double)((i * 7 + 2) % 1024);
  __syncwarp();
  double x = seed + lane * 1e-9, y = 1.0000001, z = 1e-9;
  long long t0, t1; int s = 0;
  auto rec = [&](long long a, long long b) { if (lane == 0) cyc[s] = b - a; ++s; };
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, z);
  t1 = clock64(); rec(t0, t1);                      // 0 DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + z;
  t1 = clock64(); rec(t0, t1);                      // 1 DADD
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64(); rec(t0, t1);                      // 2 DMUL
  x = fabs(x) + 1.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.5;
  t1 = clock64(); rec(t0, t1);                      // 3 rsqrt + DADD
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = 1.0 / x + 1.5;
  t1 = clock64(); rec(t0, t1);                      // 4 reciprocal + DADD
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
  t1 = clock64(); rec(t0, t1);                      // 5 SHFL 64-bit
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + z;
  t1 = clock64(); rec(t0, t1);                      // 6 SHFL + DADD (one scan step)
  int p = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) p = (int)sm[p & 1023];
  t1 = clock64(); rec(t0, t1);                      // 7 LDS.64 + F2I chase
  x += p;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31] + z; __syncwarp(); }
  t1 = clock64(); rec(t0, t1);                      // 8 STS -> syncwarp -> LDS -> DADD -> syncwarp round trip
  float f = (float)x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) f = fmaf(f, 1.0000001f, 1e-9f);
  t1 = clock64(); rec(t0, t1);                      // 9 FFMA
  t0 = clock64();
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
#pragma unroll 16
  for (int i = 0; i < N; ++i) { a0 = fma(a0, y, z); a1 = fma(a1, y, z); a2 = fma(a2, y, z); a3 = fma(a3, y, z); a4 = fma(a4, y, z); a5 = fma(a5, y, z); a6 = fma(a6, y, z); a7 = fma(a7, y, z); }
  t1 = clock64(); rec(t0, t1);                      // 10 8 independent DFMA chains (throughput per warp)
  out[lane] = x + f + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 256);
  for (int r = 0; r < 2; ++r) k<<<1, 32>>>(out, cyc, 1.0);
  long long h[16]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const char* nm[] = {"DFMA dep", "DADD dep", "DMUL dep", "rsqrt(double)+DADD", "1/x (double)+DADD", "SHFL 64-bit", "SHFL 64-bit + DADD", "LDS.64 + F2I chase", "STS/sync/LDS/DADD/sync", "FFMA dep", "8 indep DFMA (per group of 8)"};
  for (int i = 0; i < 11; ++i) printf("%-32s %7.1f cycles\n", nm[i], (double)h[i] / N);
  return 0;
}
