// dort_host.h — host-side planning shared by the CUDA library (capi.cu) and the CPU emulation driver used by the test
// suite: Gauss-Legendre nodes, workspace layout, kernel-argument assembly.  Plain C++ (no CUDA runtime calls).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

#include "../../include/smrt_dort_b200.h"
#include "dort_kernels.cuh"

namespace smrt_host {

// Positive roots of the Legendre polynomial P_{2n}, descending — what the reference obtains from
// scipy.special.roots_legendre(2n)[-1:n-1:-1] (smrt/rtsolver/streams.py:300-313, smrt/core/lib.py:669-684).
// Newton iteration on the three-term recurrence from the Tricomi initial guess, to machine precision.
inline void gauss_legendre_positive_nodes(int n, double* mu) {
  const int N = 2 * n;
  const double pi = 3.141592653589793238462643383279502884;
  for (int i = 0; i < n; ++i) {
    // i-th largest root
    long double x = std::cos(pi * (i + 0.75L) / (N + 0.5L));
    for (int it = 0; it < 100; ++it) {
      long double p0 = 1.0L, p1 = x;
      for (int k = 2; k <= N; ++k) {
        long double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0 = p1;
        p1 = pk;
      }
      long double dp = N * (x * p1 - p0) / (x * x - 1.0L);
      long double dx = p1 / dp;
      x -= dx;
      if (std::fabs((double)dx) < 1e-19) break;
    }
    mu[i] = (double)x;
  }
}

struct Layout {
  int n, m_max, mode, nmodes, hmax, K, nrhs_max;
  long long eig_stride;            // doubles per (problem, layer)
  int eig_off[SMRT_MAX_MODES];
  size_t eigen_smem_bytes, boundary_smem_bytes;            // dynamic shared memory of the shared-memory path
  size_t eigen_mid_smem_bytes;                              // ... of the eigen instantiation for 64 < h <= 128 (0: n/a)
  size_t eigen_small_packed_smem_bytes;                     // ... of the 4-CTAs-per-SM eigen instantiation (h <= 64; 0: n/a)
  size_t boundary_mid_smem_bytes;                           // ... of the boundary instantiation for 64 < h <= 128 (0: n/a)
  long long boundary_mid_scratch_doubles;                   // its per-CTA global scratch
  long long boundary_mid_arena;                             // doubles of its shared-memory matrix arena
  size_t boundary_stream_smem_bytes;                        // ... of the boundary instantiation that stages F and G (0: n/a)
  size_t eigen_vec_bytes, boundary_vec_bytes;              // vector region only (global-scratch path)
  long long eigen_scratch_doubles, boundary_scratch_doubles;  // per-CTA matrix scratch
};

inline int azimuth_half_samples(int m_max) {
  // nsamples = 2^ceil(4 + log2(m_max + 1)) (smrt/emmodel/common.py:401-414); K = nsamples / 2
  int e = 4;
  int v = 1;
  int extra = 0;
  while (v < m_max + 1) {
    v <<= 1;
    ++extra;
  }
  return (1 << (e + extra)) / 2;
}

inline Layout make_layout(const smrtb200_options& o) {
  Layout L;
  L.n = o.n_max_stream;
  L.mode = o.mode;
  L.m_max = (o.mode == SMRTB200_MODE_ACTIVE) ? o.m_max : 0;
  L.nmodes = L.m_max + 1;
  L.hmax = smrt_npol(L.m_max) * L.n;
  L.K = azimuth_half_samples(L.m_max);
  L.nrhs_max = (o.mode == SMRTB200_MODE_PASSIVE) ? 1 : 3 * 2 * o.n_inc;
  long long off = 0;
  for (int m = 0; m < SMRT_MAX_MODES; ++m) L.eig_off[m] = 0;
  for (int m = 0; m < L.nmodes; ++m) {
    long long hm = (long long)smrt_npol(m) * L.n;
    L.eig_off[m] = (int)off;
    off += smrt_even(hm) + 2 * smrt_even(hm * hm);
  }
  L.eig_stride = off;
  L.eigen_vec_bytes = eigen_vec_doubles(L.n, L.hmax, L.K) * sizeof(double);
  L.eigen_scratch_doubles = (long long)eigen_mat_doubles(L.hmax);
  L.eigen_smem_bytes = L.eigen_vec_bytes + (size_t)L.eigen_scratch_doubles * sizeof(double);
  L.eigen_mid_smem_bytes =
      (L.hmax > 64 && L.hmax <= 128) ? L.eigen_vec_bytes + eigen_mat_doubles(L.hmax, true) * sizeof(double) : 0;
  L.boundary_vec_bytes = boundary_vec_doubles(L.n, L.hmax, L.mode) * sizeof(double);
  L.boundary_scratch_doubles = (long long)boundary_mat_doubles(L.hmax, L.nrhs_max);
  L.boundary_smem_bytes = L.boundary_vec_bytes + (size_t)L.boundary_scratch_doubles * sizeof(double);
  L.eigen_small_packed_smem_bytes =
      (L.hmax <= 64) ? (eigen_vec_doubles(L.n, L.hmax, L.K, SMRT_PANEL_SMALL) + eigen_mat_doubles(L.hmax, true)) * sizeof(double)
                     : 0;
  const bool mid = L.hmax > 64 && L.hmax <= 128;
  // the matrix arena takes whatever the 227 KB opt-in limit leaves (at least the one resident matrix)
  {
    const long long limit = (227 * 1024 - 256) / 8;
    const long long fixed = (long long)(L.boundary_vec_bytes / 8 + boundary_mid_fixed_doubles(L.hmax, L.nrhs_max));
    const long long m1 = (long long)L.hmax * boundary_mid_ld(L.hmax) + SMRT_MID_STAGE;
    L.boundary_mid_arena = mid ? std::max(m1, (limit - fixed) & ~1LL) : 0;
  }
  L.boundary_mid_smem_bytes =
      mid ? L.boundary_vec_bytes + boundary_mid_smem_doubles(L.hmax, L.nrhs_max, (size_t)L.boundary_mid_arena) * sizeof(double)
          : 0;
  L.boundary_mid_scratch_doubles = mid ? (long long)boundary_mid_scratch_doubles(L.hmax) : 0;
  // F / G staged into [T | R] (deferred-store products): blocks of <= 64 unknowns only
  L.boundary_stream_smem_bytes =
      (L.hmax <= 64) ? L.boundary_vec_bytes + boundary_mat_doubles(L.hmax, L.nrhs_max, true) * sizeof(double) : 0;
  return L;
}

// Kernel arguments for problems [b0, b0 + nb) of `batch` (pointers may be host or device, the caller knows).
inline KArgs make_kargs(const smrtb200_options& o, const Layout& L, const smrtb200_batch& bt, int b0, int nb) {
  KArgs A;
  std::memset(&A, 0, sizeof(A));
  const size_t Ls = (size_t)o.max_layers;
  A.B = nb;
  A.L = o.max_layers;
  A.mode = o.mode;
  A.n = L.n;
  A.m_max = L.m_max;
  A.n_theta = o.n_theta;
  A.n_inc = o.n_inc;
  A.normalization = o.normalization;
  A.rayleigh_jeans = o.rayleigh_jeans;
  A.K = L.K;
  A.prune_tau = o.prune_deep_snowpack;
  A.phi = bt.phi;
  A.frequency = bt.frequency + b0;
  A.nlayer = bt.nlayer + b0;
  A.thickness = bt.thickness + b0 * Ls;
  A.temperature = bt.temperature + b0 * Ls;
  A.frac_volume = bt.frac_volume + b0 * Ls;
  A.eps_bg = bt.eps_bg + 2 * b0 * Ls;
  A.eps_sc = bt.eps_sc + 2 * b0 * Ls;
  A.emmodel = bt.emmodel + b0 * Ls;
  A.ms_kind = bt.ms_kind + b0 * Ls;
  A.ms_p0 = bt.ms_p0 + b0 * Ls;
  A.ms_p1 = bt.ms_p1 + b0 * Ls;
  A.interface_kind = bt.interface_kind + b0 * Ls;
  A.dense_corr = bt.dense_snow_correction + b0 * Ls;
  A.substrate_kind = bt.substrate_kind + b0;
  A.substrate_eps = bt.substrate_eps + 2 * (size_t)b0;
  A.substrate_temperature = bt.substrate_temperature + b0;
  A.substrate_params = bt.substrate_params ? bt.substrate_params + 4 * (size_t)b0 : nullptr;
  A.atmosphere = bt.atmosphere ? bt.atmosphere + 3 * (size_t)b0 : nullptr;
  A.inclusion = bt.inclusion ? bt.inclusion + 5 * b0 * Ls : nullptr;
  A.interface_params = bt.interface_params ? bt.interface_params + 4 * b0 * Ls : nullptr;
  A.theta = bt.theta;
  A.theta_inc = bt.theta_inc;
  const size_t nout = (o.mode == SMRTB200_MODE_PASSIVE) ? 2 * (size_t)o.n_theta : 9 * (size_t)o.n_inc;
  A.values = bt.values + b0 * nout;
  A.ks = bt.ks + b0 * Ls;
  A.ka = bt.ka + b0 * Ls;
  A.eps_eff = bt.eps_eff + 2 * b0 * Ls;
  A.n_streams_out = bt.n_streams_out + b0;
  A.stream_angles = bt.stream_angles + (size_t)b0 * L.n;
  A.optical_depth = bt.optical_depth + b0;
  A.status = bt.status + b0;
  A.eig_stride = L.eig_stride;
  A.mid_arena = L.boundary_mid_arena;
  for (int m = 0; m < SMRT_MAX_MODES; ++m) A.eig_off[m] = L.eig_off[m];
  return A;
}

inline const char* validate_options(const smrtb200_options& o) {
  if (o.abi_version != SMRTB200_ABI_VERSION) return "abi_version mismatch";
  if (o.mode != SMRTB200_MODE_PASSIVE && o.mode != SMRTB200_MODE_ACTIVE) return "mode must be passive or active";
  if (o.n_max_stream < 2 || o.n_max_stream > 256) return "n_max_stream must be in [2, 256]";
  if (o.mode == SMRTB200_MODE_ACTIVE && (o.m_max < 0 || o.m_max >= SMRT_MAX_MODES)) return "m_max must be in [0, 16]";
  if (o.max_layers < 1) return "max_layers must be >= 1";
  if (o.max_batch < 1) return "max_batch must be >= 1";
  if (o.mode == SMRTB200_MODE_PASSIVE && o.n_theta < 1) return "n_theta must be >= 1";
  if (o.mode == SMRTB200_MODE_ACTIVE && (o.n_inc < 1 || o.n_inc > SMRT_MAX_INC / 2)) return "n_inc must be in [1, 8]";
  if (o.normalization < 0 || o.normalization > 2) return "normalization must be 0, 1 or 2";
  return nullptr;
}

}  // namespace smrt_host
