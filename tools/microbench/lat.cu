// Dependent-chain latencies on sm_100a (profiling aid, not part of the product): DFMA, DADD, DMUL, MUFU.RSQ64H, rsqrt(),
// 1/x, sqrt(), SHFL(64-bit), LDS(64-bit), bar.sync with 128 threads.  One warp per block unless noted.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__device__ __forceinline__ double rsqrt_approx(double x) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ __forceinline__ double rcp_approx(double x) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
template <int OP> __global__ void k(double* out, long long* cyc, double a, double b) {
  __shared__ double sm[256];
  sm[threadIdx.x] = a + threadIdx.x; __syncthreads();
  double x = a + threadIdx.x * 1e-9;
  int idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, a, b);
    if (OP == 1) x = x + b;
    if (OP == 2) x = x * a;
    if (OP == 3) x = rsqrt_approx(x) + 1.5;
    if (OP == 4) x = rsqrt(x) + 1.5;
    if (OP == 5) x = 1.0 / x + 0.5;
    if (OP == 6) x = sqrt(x) + 1.5;
    if (OP == 7) x = __shfl_xor_sync(0xffffffffu, x, 1, 32);
    if (OP == 8) { idx = (int)sm[idx & 255] & 255; }
    if (OP == 9) __syncthreads();
    if (OP == 10) x = rcp_approx(x) + 0.5;
    if (OP == 11) x = (double)__any_sync(0xffffffffu, x > 0.5) + x * 1e-30;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x + idx;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  const char* names[] = {"DFMA", "DADD", "DMUL", "MUFU.RSQ64H+DADD", "rsqrt()+DADD", "1/x+DADD", "sqrt()+DADD", "SHFL64", "LDS64 chase", "bar.sync(128thr)", "MUFU.RCP64H+DADD", "vote.any+"};
  for (int op = 0; op < 12; ++op) {
    int thr = (op == 9) ? 128 : 32;
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) {
        case 0: k<0><<<1, thr>>>(out, cyc, 0.999, 1e-3); break; case 1: k<1><<<1, thr>>>(out, cyc, 0.999, 1e-3); break;
        case 2: k<2><<<1, thr>>>(out, cyc, 0.999, 1e-3); break; case 3: k<3><<<1, thr>>>(out, cyc, 2.0, 1e-3); break;
        case 4: k<4><<<1, thr>>>(out, cyc, 2.0, 1e-3); break; case 5: k<5><<<1, thr>>>(out, cyc, 2.0, 1e-3); break;
        case 6: k<6><<<1, thr>>>(out, cyc, 2.0, 1e-3); break; case 7: k<7><<<1, thr>>>(out, cyc, 2.0, 1e-3); break;
        case 8: k<8><<<1, thr>>>(out, cyc, 2.0, 1e-3); break; case 9: k<9><<<1, thr>>>(out, cyc, 2.0, 1e-3); break;
        case 10: k<10><<<1, thr>>>(out, cyc, 2.0, 1e-3); break; case 11: k<11><<<1, thr>>>(out, cyc, 2.0, 1e-3); break;
      }
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-22s %.1f cycles/iter\n", names[op], (double)c / N);
  }
  return 0;
}
