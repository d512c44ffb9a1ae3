// FP64 issue rate vs warps per SM sub-partition and ILP (profiling aid)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k(double* out, long long* cyc, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) x[u] = a + threadIdx.x * 1e-9 + u;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int u = 0; u < ILP; ++u) x[u] = fma(x[u], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int u = 0; u < ILP; ++u) s += x[u];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 4096);
  int iters = 2000;
  for (int warps = 1; warps <= 8; warps *= 2) {
    for (int ilp : {1, 2, 4, 8}) {
      int thr = 128 * warps;  // warps per SMSP
      for (int rep = 0; rep < 2; ++rep) {
        if (ilp == 1) k<1><<<148, thr>>>(out, cyc, 0.999, 1e-3, iters);
        if (ilp == 2) k<2><<<148, thr>>>(out, cyc, 0.999, 1e-3, iters);
        if (ilp == 4) k<4><<<148, thr>>>(out, cyc, 0.999, 1e-3, iters);
        if (ilp == 8) k<8><<<148, thr>>>(out, cyc, 0.999, 1e-3, iters);
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      double per_smsp = (double)iters * 8 * ilp * warps;  // warp-instructions per SMSP
      printf("warps/SMSP %d ILP %d: %.2f cycles per warp-DFMA per SMSP (%.1f DFMA lanes/clk/SM)\n", warps, ilp, c / per_smsp, 4 * 32.0 * per_smsp / c);
    }
  }
  return 0;
}
