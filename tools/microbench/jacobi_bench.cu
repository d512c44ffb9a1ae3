// Profiling aid (not part of the product): the one-sided Jacobi of the eigen kernel alone, on real operands M = C^T L
// dumped by gen_cases.py, with the launch shape of the eigen kernel (128 threads, 75 KB of shared memory -> 3 CTAs/SM).
//   ./jacobi_bench [items] [variant]     prints ms, items/s, average sweeps and the max relative error of sigma
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../smrt_b200/csrc/dort_linalg.cuh"

struct Case { int h; std::vector<double> M, sig; };

template <int NTHR, int JGV>
__global__ void __launch_bounds__(NTHR, 3) jacobi_kernel(const double* Mall, const int* hs, const long long* offs, int nitems,
                                                       double* sig_out, int* sweeps_out, int hmax) {
  extern __shared__ __align__(16) unsigned char raw[];
  double* zcol = reinterpret_cast<double*>(raw);
  double* nrm = zcol + 64;
  double* W = nrm + ((hmax + 8) & ~1);
  const int tid = threadIdx.x, NT = blockDim.x;
  for (int j = tid; j < 64; j += NT) zcol[j] = 0.0;
  __syncthreads();
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int h = hs[item];
    const double* M = Mall + offs[item];
    const int ld = jacobi_ld(h);
    for (int e = tid; e < h * (ld - 2); e += NT) {
      int i = e % (ld - 2), j = e / (ld - 2);
      W[(size_t)j * ld + i] = (i < h) ? M[(size_t)j * h + i] : 0.0;
    }
    __syncthreads();
    int sw = block_jacobi_svd_fast(W, ld, h, nrm, zcol, JGV);
    __syncthreads();
    for (int j = tid; j < h; j += NT) {
      double s2 = 0.0;
      for (int i = 0; i < h; ++i) s2 = fma(W[(size_t)j * ld + i], W[(size_t)j * ld + i], s2);
      sig_out[(size_t)item * hmax + j] = sqrt(s2);
    }
    if (tid == 0) sweeps_out[item] = sw;
    __syncthreads();
  }
}

int main(int argc, char** argv) {
  int nitems = argc > 1 ? atoi(argv[1]) : 20000;
  FILE* f = fopen("tools/microbench/cases.bin", "rb");
  if (!f) f = fopen("cases.bin", "rb");
  if (!f) { printf("cases.bin missing: run gen_cases.py\n"); return 1; }
  int nc, full; fread(&nc, 4, 1, f); fread(&full, 4, 1, f);
  std::vector<Case> cases(nc);
  for (auto& c : cases) {
    fread(&c.h, 4, 1, f);
    c.M.resize((size_t)c.h * c.h); c.sig.resize(c.h);
    std::vector<double> skip((size_t)2 * c.h * c.h);
    fread(c.M.data(), 8, c.M.size(), f); if (full) fread(skip.data(), 8, skip.size(), f); fread(c.sig.data(), 8, c.h, f);
  }
  fclose(f);
  const int hmax = 64;
  std::vector<double> Mall; std::vector<int> hs(nitems); std::vector<long long> offs(nitems);
  std::vector<long long> coff(nc);
  for (int c = 0; c < nc; ++c) { coff[c] = Mall.size(); Mall.insert(Mall.end(), cases[c].M.begin(), cases[c].M.end()); }
  for (int i = 0; i < nitems; ++i) { hs[i] = cases[i % nc].h; offs[i] = coff[i % nc]; }
  double *dM, *dsig; int *dh, *dsw; long long* doff;
  cudaMalloc(&dM, Mall.size() * 8); cudaMalloc(&dsig, (size_t)nitems * hmax * 8); cudaMalloc(&dh, nitems * 4);
  cudaMalloc(&dsw, nitems * 4); cudaMalloc(&doff, nitems * 8);
  cudaMemcpy(dM, Mall.data(), Mall.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dh, hs.data(), nitems * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(doff, offs.data(), nitems * 8, cudaMemcpyHostToDevice);
  size_t smem = 75 * 1024;
  int variant = argc > 2 ? atoi(argv[2]) : 0;
  auto kfn = variant == 0 ? jacobi_kernel<128, 8> : jacobi_kernel<256, 16>;
  int nthr = variant == 0 ? 128 : 256;
  cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, nthr, smem);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int grid = prop.multiProcessorCount * occ;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    kfn<<<grid, nthr, smem>>>(dM, dh, doff, nitems, dsig, dsw, hmax);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = std::min(best, ms);
  }
  cudaError_t err = cudaGetLastError();
  std::vector<double> sig((size_t)nitems * hmax); std::vector<int> sw(nitems);
  cudaMemcpy(sig.data(), dsig, sig.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(sw.data(), dsw, nitems * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, sws = 0;
  for (int i = 0; i < nitems; ++i) {
    const Case& c = cases[i % nc];
    std::vector<double> s(sig.begin() + (size_t)i * hmax, sig.begin() + (size_t)i * hmax + c.h);
    std::sort(s.begin(), s.end(), [](double a, double b) { return a > b; });
    for (int j = 0; j < c.h; ++j) maxerr = std::max(maxerr, fabs(s[j] - c.sig[j]) / c.sig[j]);
    sws += sw[i];
  }
  printf("variant %d occ %d grid %d items %d: %.3f ms  %.1f items/ms  avg sweeps %.2f  max rel err sigma %.2e  (%s)\n", variant, occ, grid,
         nitems, best, nitems / best, sws / nitems, maxerr, cudaGetErrorString(err));
  return 0;
}
