// TEST INFRASTRUCTURE ONLY — runs the device code of smrt_b200/csrc on the CPU through the SIMT emulator so that the
// kernels can be checked against the oracle in the authoring container (no GPU).  Built by tests/simt_emu/Makefile into
// tests/simt_emu/libsmrt_emu.so; loaded by tests/test_simt_emulation.py with ctypes.  Same batch struct as the C ABI,
// host pointers.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dort_host.h"

extern "C" {

int emu_gauss_legendre(int n, double* mu) {
  smrt_host::gauss_legendre_positive_nodes(n, mu);
  return 0;
}

// runs optics + eigen + boundary for the whole batch with `threads` emulated threads per block
int emu_solve_batch(const smrtb200_options* opt, const smrtb200_batch* batch, int threads, int* sweeps_out) {
  const char* err = smrt_host::validate_options(*opt);
  if (err) {
    std::fprintf(stderr, "emu: %s\n", err);
    return -1;
  }
  smrt_host::Layout L = smrt_host::make_layout(*opt);
  const int B = batch->B;
  const size_t BL = (size_t)B * opt->max_layers;
  std::vector<double> gl(L.n), aux(BL * SMRT_AUX_STRIDE), eig(BL * (size_t)L.eig_stride), kmin(BL * SMRT_MAX_MODES);
  std::vector<int> scat(BL, 0), counters(2, 0);
  smrt_host::gauss_legendre_positive_nodes(L.n, gl.data());
  std::vector<double> scratch((size_t)std::max(std::max(L.eigen_scratch_doubles, L.boundary_scratch_doubles),
                                               L.boundary_mid_scratch_doubles) + 64);

  KArgs A = smrt_host::make_kargs(*opt, L, *batch, 0, B);
  A.gl_mu = gl.data();
  A.aux = aux.data();
  A.eig = eig.data();
  A.kmin = kmin.data();
  A.scat_flag = scat.data();
  A.counters = counters.data();
  A.scratch = scratch.data();
  A.scratch_stride = (long long)scratch.size();
  A.use_global_scratch = 1;  // the emulator's "shared memory" is 1 MiB: keep matrices in the scratch
  A.gj_single = 1;           // like the plans of the library
  int diag[2] = {0, 0};
  A.diag = diag;
  for (int b = 0; b < B; ++b) batch->status[b] = 0;

  simt::launch((unsigned)((BL + 127) / 128), 128, [&]() { optics_kernel(A); });
  // SMRT_EMU_EIGEN_MID=1: the shared-memory instantiation for 64 < h <= 128 (packed C, Jacobi groups of 16 lanes)
  const char* em = std::getenv("SMRT_EMU_EIGEN_MID");
  if (em && em[0] == '1' && L.eigen_mid_smem_bytes > 0)
    simt::launch(1, (unsigned)threads, [&]() { eigen_kernel<2>(A); });
  else
    simt::launch(1, (unsigned)threads, [&]() { eigen_kernel<1>(A); });
  // SMRT_EMU_STREAM_FG=1: the boundary instantiation that stages F and G into [T | R] (h <= 64)
  const char* sf = std::getenv("SMRT_EMU_STREAM_FG");
  // SMRT_EMU_BOUNDARY_MID=1: the boundary instantiation for 64 < h <= 128 (its tile maps need 512 threads)
  const char* bm = std::getenv("SMRT_EMU_BOUNDARY_MID");
  // (batches with rough interfaces run the kRough instantiations, like the CUDA host code)
  const bool rough = A.interface_params != nullptr;
  if (bm && bm[0] == '1' && L.boundary_mid_smem_bytes > 0) {
    if (rough)
      simt::launch(1, 512u, [&]() { boundary_kernel<false, 512, false, true, true>(A); });
    else
      simt::launch(1, 512u, [&]() { boundary_kernel<false, 512, false, true>(A); });
  } else if (sf && sf[0] == '1' && L.hmax <= 64) {
    if (rough)
      simt::launch(1, (unsigned)threads, [&]() { boundary_kernel<true, 512, true, false, true>(A); });
    else
      simt::launch(1, (unsigned)threads, [&]() { boundary_kernel<true, 512, true>(A); });
  } else {
    if (rough)
      simt::launch(1, (unsigned)threads, [&]() { boundary_kernel<true, 512, false, false, true>(A); });
    else
      simt::launch(1, (unsigned)threads, [&]() { boundary_kernel<true, 512>(A); });
  }
  if (sweeps_out) {
    sweeps_out[0] = diag[0];
    sweeps_out[1] = diag[1];
  }
  return 0;
}

// unit hooks -----------------------------------------------------------------------------------------------------
int emu_layer_optics(double frequency, double f, double e0r, double e0i, double esr, double esi, int emmodel,
                     int ms_kind, double p0, double p1, int invert, double* out /* eps_re, eps_im, ks, ka, iba */) {
  MicroParams mp;
  LayerOptics o = layer_optics(frequency, f, c_make(e0r, e0i), c_make(esr, esi), emmodel, ms_kind, p0, p1, invert, &mp);
  out[0] = o.eps_eff.re;
  out[1] = o.eps_eff.im;
  out[2] = o.ks;
  out[3] = o.ka;
  out[4] = o.iba_coeff;
  return o.status;
}

int emu_fresnel(int kind, double e1r, double e1i, double e2r, double e2i, double mu, double* out /* R[3], T[3] */) {
  FresnelRT f = fresnel_power(kind, c_make(e1r, e1i), c_make(e2r, e2i), mu);
  for (int p = 0; p < 3; ++p) {
    out[p] = f.R[p];
    out[3 + p] = f.T[p];
  }
  return 0;
}

// Fourier mode m of the IBA phase matrix for one stream pair: out[npol*npol]
int emu_iba_phase_mode(int m, int m_max, double mu_s, double mu_i, double iba_coeff, double kk, int ms_kind, double f,
                       double p0, double p1, double* out) {
  int K = smrt_host::azimuth_half_samples(m_max);
  std::vector<double> ct(2 * K), st(2 * K);
  for (int j = 0; j < 2 * K; ++j) sincospi((double)j / K, &st[j], &ct[j]);
  MicroParams mp = micro_prepare(ms_kind, f, p0, p1);
  iba_phase_mode(m, K, ct.data(), st.data(), mu_s, mu_i, iba_coeff, kk, mp, out);
  return 0;
}

// one-sided Jacobi on an h x h column-major matrix (ld = h); returns sweeps
int emu_jacobi(double* W, int h, int threads) {
  int sweeps = 0;
  std::vector<int> ctrl(8, 0);
  simt::launch(1, (unsigned)threads, [&]() {
    int s = block_jacobi_svd(W, h, h, ctrl.data());
    if (threadIdx.x == 0) sweeps = s;
  });
  return sweeps;
}

// register-blocked variant (h <= 64): W column-major with leading dimension emu_jacobi_ld(h), rows >= h zero
int emu_jacobi_ld(int h) { return jacobi_ld(h); }
int emu_jacobi_fast(double* W, int h, int ld, int threads) {
  int sweeps = 0;
  std::vector<double> nrm(h + 8, 0.0), zcol(128, 0.0);
  simt::launch(1, (unsigned)threads, [&]() {
    int s = block_jacobi_svd_fast(W, ld, h, nrm.data(), zcol.data(), h > 64 ? 16 : 8);
    if (threadIdx.x == 0) sweeps = s;
  });
  return sweeps;
}

// blocked Gauss-Jordan: X = A^-1 R for column-major A (h x h, ld = h) and R (h x nR, ld = h); returns the failure flag
int emu_gj_blocked(double* A, int h, double* R, int nR, int threads, double* X) {
  std::vector<int> rowof(h, -1), flag(1, 0);
  std::vector<double> ipiv(h, 0.0), Vbuf((size_t)2 * h * SMRT_GJ_NB + 8, 0.0);
  int rc = 0;
  simt::launch(1, (unsigned)threads, [&]() {
    int r = block_gj_rows_blocked<false>(A, h, R, h, h, nR, rowof.data(), ipiv.data(), Vbuf.data(), flag.data());
    if (threadIdx.x == 0) rc = r;
  });
  if (rc == 0)
    for (int k = 0; k < h; ++k)
      for (int c = 0; c < nR; ++c) X[(size_t)c * h + k] = R[(size_t)c * h + rowof[k]] * ipiv[k];
  return rc;
}

// left-looking two-column Cholesky of the lower triangle of A (h x h, ld = h), two concurrent teams like the kernel:
// team 0 factorises A, team 1 factorises A2.  Returns the failure flags (bit 0 / bit 1).
int emu_cholesky_pair(double* A, double* A2, int h, int threads) {
  std::vector<double> d0(h + 2, 0.0), d1(h + 2, 0.0);
  int bad[2] = {0, 0};
  simt::launch(1, (unsigned)threads, [&]() {
    Team tm;
    const int half = (int)blockDim.x / 2;
    const int which = (int)threadIdx.x / half;
    tm.size = half;
    tm.rank = (int)threadIdx.x % half;
    tm.bar_id = 1 + which;
    LowerMat<false> X;
    X.p = which == 0 ? A : A2;
    X.ld = h;
    X.h = h;
    int b = team_cholesky_fast(tm, X, h, which == 0 ? d0.data() : d1.data());
    if (tm.rank == 0) bad[which] = b;
  });
  return bad[0] | (bad[1] << 1);
}

// W <- C^-T W (C lower triangular h x h, ld = h; W h x h, ld = ldw)
int emu_backsolve_lt(const double* Cm, double* W, int h, int ldw, int threads) {
  std::vector<double> rdiag(h);
  for (int j = 0; j < h; ++j) rdiag[j] = 1.0 / Cm[(size_t)j * h + j];
  LowerMat<false> Cl;
  Cl.p = const_cast<double*>(Cm);
  Cl.ld = h;
  Cl.h = h;
  simt::launch(1, (unsigned)threads, [&]() { block_backsolve_lt(Cl, W, ldw, h, rdiag.data()); });
  return 0;
}

// y1 = A1 x, y2 = A2 x (M x K column-major, lda = M); A2 may be NULL
int emu_matvec_dual(const double* A1, const double* A2, int M, int K, const double* x, double* y1, double* y2, int threads) {
  std::vector<double> part((size_t)16 * M + 8, 0.0);
  simt::launch(1, (unsigned)threads, [&]() {
    block_matvec_dual(M, K, A1, A2, M, x, part.data(), [&](int i, double a, double b) {
      y1[i] = a;
      y2[i] = b;
    });
  });
  return 0;
}
}
