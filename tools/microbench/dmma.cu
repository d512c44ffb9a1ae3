// FP64 tensor-core instruction (mma.sync.aligned.m8n8k4.row.col.f64 = DMMA, 256 FMAs per warp instruction) on sm_100a:
// issue rate per SM sub-partition versus warps and independent accumulators, and dependent-chain latency.
// Profiling aid for the question "would the small dense products / rank-4 updates of the DORT kernels gain from DMMA?"
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP> __global__ void k(double* out, long long* cyc, double a, double b, int iters) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) { c0[u] = threadIdx.x * 1e-9 + u; c1[u] = u; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int u = 0; u < ILP; ++u) dmma(c0[u], c1[u], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int u = 0; u < ILP; ++u) s += c0[u] + c1[u];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 4096);
  int iters = 500;
  for (int warps = 1; warps <= 4; warps *= 2) {
    for (int ilp : {1, 2, 4, 8}) {
      int thr = 128 * warps;  // warps per SMSP
      for (int rep = 0; rep < 2; ++rep) {
        if (ilp == 1) k<1><<<148, thr>>>(out, cyc, 1e-3, 1e-3, iters);
        if (ilp == 2) k<2><<<148, thr>>>(out, cyc, 1e-3, 1e-3, iters);
        if (ilp == 4) k<4><<<148, thr>>>(out, cyc, 1e-3, 1e-3, iters);
        if (ilp == 8) k<8><<<148, thr>>>(out, cyc, 1e-3, 1e-3, iters);
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      double per_smsp = (double)iters * 8 * ilp * warps;  // warp-instructions per SMSP
      printf("warps/SMSP %d ILP %d: %.2f cycles per DMMA per SMSP (%.1f FP64 FMA lanes/clk/SM; DFMA pipe peak is 64)\n", warps, ilp,
             c / per_smsp, 4 * 256.0 * per_smsp / c);
    }
  }
  return 0;
}
