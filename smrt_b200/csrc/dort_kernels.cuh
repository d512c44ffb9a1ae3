// dort_kernels.cuh — the three kernels of the B200 DORT path.
//
//   optics_kernel    one THREAD per (problem, layer): effective permittivity, ks (65-point Romberg), ka, IBA coefficient
//   eigen_kernel     one CTA per (problem, layer) work item (persistent grid, atomic work counter): stream angles,
//                    Fourier modes of the phase matrix on the layer's streams, symmetrised half-rank eigenproblem by
//                    two concurrent Cholesky factorisations + one-sided Jacobi, eigenvectors written to the layer
//                    workspace in HBM (k, F, G with Eu = [F | G], Ed = D [G | F], beta = [k, -k])
//   boundary_kernel  one CTA per problem (persistent grid): Fresnel coefficients, bottom-up elimination of the
//                    block-tridiagonal boundary system with h x h blocks, emerging intensity, mode summation,
//                    inverse Planck, interpolation to the sensor angles
//
// Algebra and reference citations: DESIGN.md §3-§5; NumPy model of exactly this algorithm: oracle/b200_algorithm.py.
#pragma once
#include "dort_device.cuh"
#include "dort_linalg.cuh"

#define SMRT_MAX_MODES 17     // m = 0 .. 16 (the reference's tests go up to m_max = 16, rtsolver/test_dort.py:13-42)
#define SMRT_MAX_INC 16       // incident streams (<= 2 per incidence angle, n_inc <= 8)
#define SMRT_AUX_STRIDE 4     // per (problem, layer): iba_coeff, kk, f_eff, spare
#define SMRT_NT 256           // threads per CTA of the eigen kernel, global-scratch instantiation (large stream counts)
#define SMRT_NT_SMEM 128      // ... shared-memory instantiation (h <= 64): 16 Jacobi groups of 8 lanes, 3 CTAs per SM
#define SMRT_NT_MID 512       // ... shared-memory instantiation for 64 < h <= 128: 32 Jacobi groups of 16 lanes, 1 CTA per SM
#define SMRT_NT_B 512         // max threads per CTA of the boundary kernel

struct KArgs {
  int B, L;  // problems in this launch (all pointers are already offset to its first problem), row stride
  int mode, n, m_max, n_theta, n_inc, normalization, rayleigh_jeans, K;
  double prune_tau, phi;
  // inputs
  const double* frequency;
  const int* nlayer;
  const double *thickness, *temperature, *frac_volume, *eps_bg, *eps_sc;
  const int *emmodel, *ms_kind;
  const double *ms_p0, *ms_p1;
  const int *interface_kind, *dense_corr, *substrate_kind;
  const double *substrate_eps, *substrate_temperature, *theta, *theta_inc;
  const double *substrate_params, *atmosphere;  // optional (NULL): [B, 4] substrate model parameters, [B, 3] atmosphere
  const double* inclusion;                      // optional (NULL): [B, L, 5] inclusion shape weights, depolarisation factors
  const double* interface_params;               // optional (NULL): [B, L, 4] parameters of the rough interface above layer l
  // outputs
  double *values, *ks, *ka, *eps_eff;
  int* n_streams_out;
  double *stream_angles, *optical_depth;
  int* status;
  // workspace
  const double* gl_mu;  // [n] positive Gauss-Legendre nodes of order 2n, descending
  double* aux;          // [B, L, SMRT_AUX_STRIDE]
  double* eig;          // [B, L, eig_stride]
  double* kmin;         // [B, L, SMRT_MAX_MODES]
  int* scat_flag;       // [B, L]
  int* counters;        // [2]: eigen / boundary work counters
  int* diag;            // optional [2] diagnostics: total Jacobi sweeps, number of eigenproblems (may be NULL)
  unsigned long long* prof;  // optional [16] cycle counters of the boundary kernel phases (SMRT_B200_PROFILE; may be NULL)
  double* scratch;      // [gridDim.x, scratch_stride] when use_global_scratch
  long long eig_stride, scratch_stride;
  long long mid_arena;  // boundary kernel for 64 < h <= 128: doubles of the shared-memory matrix arena
  int use_global_scratch;
  int gj_single;  // one blocked Gauss-Jordan instantiation for every block of a plan (0: pick by block size)
  int eig_off[SMRT_MAX_MODES];  // offset of mode m inside one (problem, layer) eigen record
};

// layout of one mode record: k[hmax_even] | F[h*h] | G[h*h]  (F, G compact column-major with ld = h)
SMRT_HD int smrt_npol(int m) { return m == 0 ? 2 : 3; }
SMRT_HD long long smrt_even(long long x) { return (x + 1) & ~1LL; }

// 2-D element loop without integer division: warps walk the columns, lanes the rows (coalesced / conflict-free)
#define SMRT_FOR_2D(i, j, nrows, ncols)                                   \
  for (int j = (int)(threadIdx.x >> 5); j < (ncols); j += (int)(blockDim.x >> 5)) \
    for (int i = (int)(threadIdx.x & 31); i < (nrows); i += 32)

#ifdef __CUDACC__
#define SMRT_DYN_SMEM(ptr)                                        \
  extern __shared__ __align__(16) unsigned char smrt_smem_raw[]; \
  double* ptr = reinterpret_cast<double*>(smrt_smem_raw)
#else
#define SMRT_DYN_SMEM(ptr) double* ptr = reinterpret_cast<double*>(simt::dynamic_smem(0))
#endif

// Re sqrt(eps_star / eps_medium): the refraction index ratio that maps the most refringent layer's nodes to a medium
SMRT_DEV double real_index_of(cplx eps_star, cplx eps_medium) { return c_sqrt(c_div(eps_star, eps_medium)).re; }

SMRT_DEV void set_error(int* status, int b, int code) {
  // the error codes are enumerated values, not bits: the FIRST error wins (compare-and-swap on the low 4 bits; CTAs of
  // different layers of one problem run concurrently); warnings are OR-ed separately into the high bits
  int old = status[b];
  while ((old & 15) == 0) {
    const int seen = atomicCAS(&status[b], old, old | code);
    if (seen == old) break;
    old = seen;
  }
}

// --------------------------------------------------------------------------------------------------------------------
// kernel 1: layer optics
// --------------------------------------------------------------------------------------------------------------------
SMRT_GLOBAL void __launch_bounds__(128) optics_kernel(KArgs A) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= A.B * A.L) return;
  int b = idx / A.L, l = idx % A.L;
  if (l >= A.nlayer[b]) return;  // status[] is zeroed by the host before the launch
  cplx e0 = c_make(A.eps_bg[2 * idx], A.eps_bg[2 * idx + 1]);
  cplx es = c_make(A.eps_sc[2 * idx], A.eps_sc[2 * idx + 1]);
  MicroParams mp;
  LayerOptics o = layer_optics(A.frequency[b], A.frac_volume[idx], e0, es, A.emmodel[idx], A.ms_kind[idx],
                               A.ms_p0[idx], A.ms_p1[idx], A.dense_corr[idx], &mp,
                               A.inclusion ? A.inclusion + 5 * (size_t)idx : nullptr);
  A.eps_eff[2 * idx] = o.eps_eff.re;
  A.eps_eff[2 * idx + 1] = o.eps_eff.im;
  A.ks[idx] = o.ks;
  A.ka[idx] = o.ka;
  double* aux = A.aux + (size_t)idx * SMRT_AUX_STRIDE;
  aux[0] = o.iba_coeff;
  aux[1] = o.kk;
  aux[2] = o.f;
  aux[3] = 0.0;
  if (o.status != ST_OK) set_error(A.status, b, o.status);
}

// parameters of the (rough) interface above layer l_ of problem b (bL = b * L), or NULL
#define SMRT_IPAR(l_) (A.interface_params ? A.interface_params + 4 * (size_t)(bL + (l_)) : nullptr)

// Rough interfaces of a layer (boundary kernel): the entries the Fresnel pass wrote for a rough interface are replaced by
// the Kirchhoff coherent coefficients plus, in every pass but the coherent one, the diagonal diffuse (backscatter)
// reflection -- rtsolver_utils.py:480-522, 551-582, 690-709.  Same (entry -> thread) map as the Fresnel pass: the thread
// that wrote an entry rewrites it, no barrier in between.  NOT inlined: batches without rough interfaces (the pointer
// test is block-uniform) keep the register allocation of the kernel body.
// air - snow interface seen from the air on an air stream (rtsolver_utils.py:607-642); plain Fresnel unless it is rough
SMRT_DEV_NOINLINE FresnelRT air_interface_power(const KArgs& A, long long bL, double freq, cplx eps0, double mu, double w,
                                                int m_diff, int m_max) {
  const int ik = A.interface_kind[bL];
  if (ik >= IF_IEM_FUNG92 && A.interface_params)
    return rough_interface_power(ik, A.interface_params + 4 * (size_t)bL, freq, c_make(1.0, 0.0), eps0, mu, w, m_diff, m_max);
  return fresnel_power(ik, c_make(1.0, 0.0), eps0, mu);
}
struct RoughFix {
  int ik_top, ik_bot, l, nl, n_l, n, npol, m_diff, m_max;
  bool up_emits;
  const double *par_top, *par_bot, *eps_b, *mu, *gl_mu;
  double freq;
  cplx eps_l, eps_star;
  double *Rt, *Tt, *Rb, *Tb, *RbD, *Tup;
};
SMRT_DEV_NOINLINE void rough_interface_fixup(const RoughFix& f) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int span = 32 * ((f.n_l + 31) >> 5), npol = f.npol, l = f.l;
  for (int e = tid; e < 3 * span; e += NT) {
    const int which = e / span, j = e - which * span;
    if (j >= f.n_l) continue;
    const double wj = (f.n_l >= 2) ? stream_weight(f.mu, f.n_l, j) : 0.0;
    if (which == 0 && f.ik_top >= IF_IEM_FUNG92) {
      cplx eps_up = (l > 0) ? c_make(f.eps_b[2 * (l - 1)], f.eps_b[2 * (l - 1) + 1]) : c_make(1.0, 0.0);
      FresnelRT ft = rough_interface_power(f.ik_top, f.par_top, f.freq, f.eps_l, eps_up, f.mu[j], wj, f.m_diff, f.m_max);
      for (int p = 0; p < npol; ++p) {
        f.Rt[j * npol + p] = ft.R[p];
        f.Tt[j * npol + p] = ft.T[p];
      }
    } else if (which == 1 && f.ik_bot >= IF_IEM_FUNG92) {
      FresnelRT fb = rough_interface_power(f.ik_bot, f.par_bot, f.freq, f.eps_l,
                                           c_make(f.eps_b[2 * (l + 1)], f.eps_b[2 * (l + 1) + 1]), f.mu[j], wj, f.m_diff,
                                           f.m_max);
      for (int p = 0; p < npol; ++p) {
        f.Rb[j * npol + p] = fb.R[p];
        f.Tb[j * npol + p] = fb.T[p];
        f.RbD[j * npol + p] = (p == 2) ? -fb.R[p] : fb.R[p];
      }
    } else if (which == 2 && f.ik_top >= IF_IEM_FUNG92 && f.up_emits) {
      cplx eps_up = c_make(f.eps_b[2 * (l - 1)], f.eps_b[2 * (l - 1) + 1]);
      double ri_up = real_index_of(f.eps_star, eps_up);
      if (j < stream_count(ri_up, f.gl_mu, f.n)) {
        FresnelRT fu = rough_interface_power(f.ik_top, f.par_top, f.freq, eps_up, f.eps_l, stream_mu(ri_up, f.gl_mu, j),
                                             0.0, -1, 0);
        for (int p = 0; p < npol; ++p) f.Tup[j * npol + p] = fu.T[p];
      }
    }
  }
}

// --------------------------------------------------------------------------------------------------------------------
// kernel 2: per-layer eigenproblem
// --------------------------------------------------------------------------------------------------------------------
// shared-memory vector region (doubles): mu[n] w[n] norm0[2n] g[hmax] sdiag[hmax] dk[hmax] sigma[hmax] ctab[2K] stab[2K]
//                                        gq[hmax] nrm[hmax + 12] | g[hmax] dk[hmax] = panel[kPanel * hmax]; in front: zcol (zeros)
// matrix region: A1 (X- -> L -> M -> W -> E~+, hmax x (hmax + 3)), A2 (X+ -> C, hmax x (hmax + 1)): 68 KB at 32
// streams, so that three CTAs fit on an SM
#define SMRT_PANEL 8  // columns of L staged per step of the in-place product M = C^T L
// the column of zeros standing for missing Jacobi columns: 8 lanes x 8 rows, or 16 lanes x 8 rows beyond 64 unknowns
SMRT_HD int eigen_zcol_doubles(int hmax) { return hmax > 64 ? 128 : 64; }
// (the staging panel of M = C^T L lies over the scales g and ke / mu, which are dead by then; `panel` = its column count)
SMRT_HD size_t eigen_vec_doubles(int n, int hmax, int K, int panel = SMRT_PANEL) {
  return ((size_t)eigen_zcol_doubles(hmax) + 4 * n + 4 * hmax + 4 * K + (size_t)panel * hmax + 12 + 1) & ~(size_t)1;
}
SMRT_HD size_t eigen_mat1_doubles(int hmax) { return (size_t)hmax * jacobi_ld(hmax); }  // even: A2 16-byte aligned
// packed: the second matrix (X+ -> C) holds its lower triangle only, packed by columns
SMRT_HD size_t eigen_mat_doubles(int hmax, bool packed = false) {
  return eigen_mat1_doubles(hmax) + (packed ? ((size_t)hmax * (hmax + 1) / 2 + 1) & ~(size_t)1 : (size_t)hmax * (hmax + 1));
}

// kVariant 0: matrices in shared memory, h <= 64, 128 threads, three CTAs per SM
//          1: matrices in a per-CTA global (L2-resident) scratch, any h, 256 threads
//          2: matrices in shared memory, 64 < h <= 128, 512 threads, one CTA per SM: X+ / C packed (lower triangle),
//             Jacobi groups of 16 lanes with 6 or 8 rows per lane
//          3: as 0 with X+ / C packed and a 4-column staging panel: 56 KB and 128 registers, FOUR CTAs per SM
#define SMRT_PANEL_SMALL 4
template <int kVariant>
SMRT_GLOBAL void __launch_bounds__(kVariant == 1 ? SMRT_NT : (kVariant == 2 ? SMRT_NT_MID : SMRT_NT_SMEM),
                                   kVariant == 0 ? 3 : (kVariant == 3 ? 4 : 1)) eigen_kernel(KArgs A) {
  constexpr bool kGlobalScratch = kVariant == 1;
  constexpr bool kPacked = kVariant == 2 || kVariant == 3;
  constexpr int kPanel = kVariant == 3 ? SMRT_PANEL_SMALL : SMRT_PANEL;
  SMRT_DYN_SMEM(smem);
  SMRT_SHARED int s_item;
  SMRT_SHARED int s_ctrl[8];
  const int tid = threadIdx.x;
  const int NT = blockDim.x;
  const int n = A.n;
  const int nmodes = A.m_max + 1;
  const int hmax = smrt_npol(A.m_max) * n;
  const int K = A.K;

  double* zcol = smem;  // a column of zeros: stands for the missing columns of the register-blocked Jacobi
  const int nz = eigen_zcol_doubles(hmax);
  double* mu = zcol + nz;
  double* w = mu + n;
  double* norm0 = w + n;
  double* sdiag = norm0 + 2 * n;
  double* sigma = sdiag + hmax;
  double* ctab = sigma + hmax;
  double* stab = ctab + 2 * K;
  double* gq = stab + 2 * K;
  double* nrm = gq + hmax;  // tracked squared column norms of the Jacobi sweeps
  double* gvec = nrm + hmax + 12;
  double* dk = gvec + hmax;
  double* panel = gvec;  // kPanel * hmax doubles over [gvec | dk | ...]: the scales are dead when M = C^T L is formed
  // compile-time choice so that the shared-memory instantiation addresses its matrices with LDS/STS, not generic LD/ST
  double* mats = kGlobalScratch ? (A.scratch + (size_t)blockIdx.x * A.scratch_stride)
                                : (smem + eigen_vec_doubles(n, hmax, K, kPanel));
  double* A1 = mats;
  double* A2 = mats + eigen_mat1_doubles(hmax);

  for (int j = tid; j < nz; j += NT) zcol[j] = 0.0;
  for (int j = tid; j < 2 * K; j += NT) {
    double s, c;
    sincospi((double)j / (double)K, &s, &c);
    ctab[j] = c;
    stab[j] = s;
  }
  __syncthreads();

  const int nitems = A.B * A.L;
  for (;;) {
    if (tid == 0) s_item = atomicAdd(&A.counters[0], 1);
    __syncthreads();
    const int item = s_item;
    __syncthreads();
    if (item >= nitems) break;
    const int b = item / A.L, l = item % A.L;
    const int nl = A.nlayer[b];
    if (l >= nl) continue;
    const size_t bl = (size_t)b * A.L + l;
    const double* eps_b = A.eps_eff + (size_t)b * A.L * 2;
    const double ks = A.ks[bl], ke = A.ks[bl] + A.ka[bl];
    const int emmodel = A.emmodel[bl];

    // streams of this layer -------------------------------------------------------------------- streams.py:136-223
    const int kstar = most_refringent_layer(eps_b, nl);
    const cplx eps_star = c_make(eps_b[2 * kstar], eps_b[2 * kstar + 1]);
    const double rindex = real_index_of(eps_star, c_make(eps_b[2 * l], eps_b[2 * l + 1]));
    const int n_l = stream_count(rindex, A.gl_mu, n);
    double* kmin = A.kmin + bl * SMRT_MAX_MODES;
    if (n_l < 2) {
      if (tid == 0) {
        set_error(A.status, b, ST_INPUT);
        A.scat_flag[bl] = 0;
        for (int m = 0; m < nmodes; ++m) kmin[m] = 0.0;
      }
      continue;
    }
    for (int j = tid; j < n_l; j += NT) mu[j] = stream_mu(rindex, A.gl_mu, j);
    __syncthreads();
    for (int j = tid; j < n_l; j += NT) w[j] = stream_weight(mu, n_l, j);
    __syncthreads();

    const bool scattering = (ks != 0.0) && (emmodel != EM_NONSCATTERING);
    if (!scattering) {  // dort.py:716-717, 765-780: trivial solution, nothing to store
      if (tid == 0) {
        A.scat_flag[bl] = 0;
        for (int m = 0; m < nmodes; ++m) kmin[m] = ke / mu[0];
      }
      __syncthreads();
      continue;
    }

    const double* aux = A.aux + bl * SMRT_AUX_STRIDE;
    const double iba_coeff = aux[0], kk = aux[1];
    MicroParams mp = micro_prepare(A.ms_kind[bl], aux[2], A.ms_p0[bl], A.ms_p1[bl]);
    double* rec = A.eig + bl * A.eig_stride;
    bool failed = false;

    for (int m = 0; m < nmodes && !failed; ++m) {
      const int npol = smrt_npol(m);
      const int h = npol * n_l;
      // A2 (X+ -> C) is read row-wise by lanes: odd leading dimension (conflict-free); A1 (X- -> L -> M -> W) is the
      // Jacobi operand: even leading dimension (16-byte aligned columns), the pad row of an odd h is kept at zero
      const int ld = (h & 1) ? h : h + 1;
      const int ld1 = jacobi_ld(h);
      const double coef = (m == 0) ? 0.5 : 0.25;
      LowerMat<kPacked> C2;  // X+ -> C: full (ld) or packed lower triangle
      C2.p = A2;
      C2.ld = ld;
      C2.h = h;

      // phase matrix Fourier mode m on (mu_s > 0) x (mu_i > 0 | mu_i < 0): A1 <- P++, A2 <- (P+-) D.
      // Only the stream pairs js <= ji are evaluated: the phase matrix is reciprocal, P_ab(i, s) = P_ba(s, i) q_a / q_b
      // with q = (1, 1, 2) for (V, H, U), and for the backward half the same relation carries the signs D = (1, 1, -1)
      // of the third Stokes component on both sides (both follow from the sums of emmodel/common.py:87-129 term by
      // term), so the block (ji, js) is the mirrored, re-weighted block (js, ji).
      // (one work item = one stream pair and one half, forward or backward: twice as many, half as long items fill
      // the last pass over the threads better)
      for (int item2 = tid; item2 < n_l * (n_l + 1); item2 += NT) {
        const int pidx = item2 >> 1;
        const bool backward = (item2 & 1) != 0;
        int ji = (int)((sqrtf(8.0f * (float)pidx + 1.0f) - 1.0f) * 0.5f);
        while ((ji + 1) * (ji + 2) / 2 <= pidx) ++ji;
        while (ji * (ji + 1) / 2 > pidx) --ji;
        const int js = pidx - ji * (ji + 1) / 2;
        double pv[9];
        const double mui = backward ? -mu[ji] : mu[ji];
        if (em_is_iba(emmodel)) {
          iba_phase_mode(m, K, ctab, stab, mu[js], mui, iba_coeff, kk, mp, pv);
        } else {
          rayleigh_phase_mode(m, mu[js], mui, ks, pv);
        }
        for (int ps = 0; ps < npol; ++ps)
          for (int pi = 0; pi < npol; ++pi) {
            const int a = js * npol + ps, c = ji * npol + pi;
            double v = pv[ps * npol + pi];
            if (backward && pi == 2) v = -v;
            // mirrored block: row (ji, pi), column (js, ps)
            const double qr = ((pi == 2) ? 2.0 : 1.0) / ((ps == 2) ? 2.0 : 1.0);
            if (!backward) {
              SMRT_AT(A1, ld1, a, c) = v;
              if (js != ji) SMRT_AT(A1, ld1, c, a) = v * qr;
            } else if (!kPacked) {
              SMRT_AT(A2, ld, a, c) = v;
              if (js != ji) SMRT_AT(A2, ld, c, a) = v * qr;
            } else {  // lower triangle only
              if (a >= c) C2.at(a, c) = v;
              if (js != ji) C2.at(c, a) = v * qr;
            }
          }
      }
      if (tid == 0) s_ctrl[2] = 0;
      __syncthreads();

      // row normalisation (dort.py:782-819) and the symmetrising scales
      for (int a = tid; a < h; a += NT) {
        int j = a / npol, p = a % npol;
        double norm = 1.0;
        if (A.normalization != 0) {
          if (m == 0) {
            double rs = 0.0;
            if (!kPacked) {
              for (int c = 0; c < h; ++c) rs += (SMRT_AT(A1, ld1, a, c) + SMRT_AT(A2, ld, a, c)) * w[c / npol];
            } else {
              // the part of the backward half above the diagonal follows from reciprocity:
              // P(a, c) = P(c, a) q_a / q_c with q = (1, 1, 2) for (V, H, U)
              const double qa = (p == 2) ? 2.0 : 1.0;
              for (int c = 0; c <= a; ++c) rs += (SMRT_AT(A1, ld1, a, c) + C2.at(a, c)) * w[c / npol];
              for (int c = a + 1; c < h; ++c) {
                const double qc = (c % npol == 2) ? 2.0 : 1.0;
                rs += (SMRT_AT(A1, ld1, a, c) + C2.at(c, a) * (qa / qc)) * w[c / npol];
              }
            }
            // A row sum = -coef * rs ; norm_0 = -ks / rowsum
            norm = ks / (coef * rs);
            norm0[a] = norm;
            if (A.normalization == 1 && !(fabs(norm - 1.0) <= 0.3)) s_ctrl[2] = 1;
          } else {
            norm = (p < 2) ? norm0[2 * j + p] : sqrt(norm0[2 * j] * norm0[2 * j + 1]);
          }
        } else if (m == 0) {
          norm0[a] = 1.0;
        }
        double q = (p == 2) ? 2.0 : 1.0;
        double cw = coef * w[j];
        gvec[a] = sqrt(norm * q * cw / mu[j]);
        gq[a] = gvec[a] / q;
        sdiag[a] = sqrt(norm * q / (mu[j] * cw));
        dk[a] = ke / mu[j];
      }
      __syncthreads();
      if (s_ctrl[2]) {
        if (tid == 0) set_error(A.status, b, ST_NORMALIZATION);
        failed = true;
        break;
      }
      // X- = diag(ke/mu) - g Ps++ g + g Ps+-' g   (A1),   X+ = diag(ke/mu) - g Ps++ g - g Ps+-' g   (A2)
      // (gq = g / q: the reciprocity weight of the row is folded into the row scale)
      if (!kPacked) {
        SMRT_FOR_2D(a, c, h, h) {
          double sc = gq[a] * gvec[c];
          double x1 = sc * SMRT_AT(A1, ld1, a, c), x2 = sc * SMRT_AT(A2, ld, a, c);
          double d = (a == c) ? dk[a] : 0.0;
          SMRT_AT(A1, ld1, a, c) = d - x1 + x2;
          SMRT_AT(A2, ld, a, c) = d - x1 - x2;
        }
      } else {  // both matrices are symmetric and only their lower triangles are read from here on
        SMRT_FOR_2D(a, c, h, h) {
          if (a >= c) {
            double sc = gq[a] * gvec[c];
            double x1 = sc * SMRT_AT(A1, ld1, a, c), x2 = sc * C2.at(a, c);
            double d = (a == c) ? dk[a] : 0.0;
            SMRT_AT(A1, ld1, a, c) = d - x1 + x2;
            C2.at(a, c) = d - x1 - x2;
          }
        }
      }
      __syncthreads();

      // two concurrent Cholesky factorisations: X- = L L^T (A1), X+ = C C^T (A2)
      {
        Team tm;
        int half = NT / 2;
        tm.size = half;
        int which = tid / half;
        tm.rank = tid % half;
        tm.bar_id = 1 + which;
        int bad;
        if (which == 0) {
          LowerMat<false> X1;
          X1.p = A1;
          X1.ld = ld1;
          X1.h = h;
          bad = team_cholesky_fast(tm, X1, h, sigma);
        } else {
          bad = team_cholesky_fast(tm, C2, h, nrm);
        }
        if (tm.rank == 0) s_ctrl[4 + which] = bad;
      }
      __syncthreads();
      if (s_ctrl[4] | s_ctrl[5]) {
        if (tid == 0) set_error(A.status, b, ST_EIGEN);
        failed = true;
        break;
      }

      // M = C^T L in place over L (A1): M(i, j) = sum_{k >= max(i, j)} C(k, i) L(k, j) needs only column j of L,
      // staged kPanel columns at a time
      for (int jp = 0; jp < h; jp += kPanel) {
        const int pw = (h - jp < kPanel) ? (h - jp) : kPanel;
        for (int e = tid; e < h * pw; e += NT) {
          int k = e % h, jj = e / h;
          panel[jj * h + k] = (k >= jp + jj) ? SMRT_AT(A1, ld1, k, jp + jj) : 0.0;
        }
        __syncthreads();
        // one row of C^T against FOUR staged columns per thread: 5 shared-memory loads per 4 FMAs (the staged
        // columns are zero above their diagonal, so a common lower summation bound is exact)
        for (int e = tid; e < h * ((pw + 3) >> 2); e += NT) {
          const int i = e % h, jq = (e / h) << 2;
          const double* cc = C2.col(i);
          const double* p0 = panel + (size_t)jq * h;
          const int nj = (pw - jq < 4) ? (pw - jq) : 4;
          const double* p1 = p0 + ((nj > 1) ? h : 0);
          const double* p2 = p0 + ((nj > 2) ? 2 * h : 0);
          const double* p3 = p0 + ((nj > 3) ? 3 * h : 0);
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          const int k0 = (i > jp + jq) ? i : jp + jq;
          for (int k = k0; k < h; ++k) {
            const double c = cc[k];
            a0 = fma(c, p0[k], a0);
            a1 = fma(c, p1[k], a1);
            a2 = fma(c, p2[k], a2);
            a3 = fma(c, p3[k], a3);
          }
          SMRT_AT(A1, ld1, i, jp + jq) = a0;
          if (nj > 1) SMRT_AT(A1, ld1, i, jp + jq + 1) = a1;
          if (nj > 2) SMRT_AT(A1, ld1, i, jp + jq + 2) = a2;
          if (nj > 3) SMRT_AT(A1, ld1, i, jp + jq + 3) = a3;
        }
        __syncthreads();
      }

      // singular values / right rotations by one-sided Jacobi: A1 <- W = U Sigma
      {
        const bool fastj = (kVariant == 2) ? (h <= 128) : (h <= 64);
        if (fastj) {  // zero pad rows of the register-blocked Jacobi operand: rows [h, ld1 - 2)
          const int npad = ld1 - 2 - h;
          for (int e = tid; e < npad * h; e += NT) SMRT_AT(A1, ld1, h + e % npad, e / npad) = 0.0;
          __syncthreads();
        }
        int sw = fastj ? block_jacobi_svd_fast(A1, ld1, h, nrm, zcol, (kVariant == 2) ? 16 : 8)
                       : block_jacobi_svd(A1, ld1, h, s_ctrl);
        if (tid == 0 && A.diag) {
          atomicAdd(&A.diag[0], sw);
          atomicAdd(&A.diag[1], 1);
        }
        if (sw >= SMRT_JACOBI_MAX_SWEEPS) {  // not converged (never seen): report it instead of using the vectors
          if (tid == 0) set_error(A.status, b, ST_EIGEN);
          failed = true;
          break;
        }
      }
      __syncthreads();
      for (int j = tid; j < h; j += NT) {
        double s2 = 0.0;
        for (int i = 0; i < h; ++i) s2 = fma(SMRT_AT(A1, ld1, i, j), SMRT_AT(A1, ld1, i, j), s2);
        sigma[j] = sqrt(s2);
      }
      __syncthreads();

      double* rk = rec + A.eig_off[m];
      double* rF = rk + smrt_even(smrt_npol(m) * n);
      double* rG = rF + smrt_even((long long)h * h);
      // E~- = -C U = -C W Sigma^-1, staged in the G slot of the layer record (global memory; read back below)
      {
        Team tm = block_team();
        team_gemm(
            tm, h, h, h, [&](int i, int k) { return (k <= i) ? C2.at(i, k) : 0.0; },
            [&](int k, int j) { return SMRT_AT(A1, ld1, k, j); },
            [&](int i, int j, double acc) { rG[(size_t)j * h + i] = -acc / sigma[j]; });
      }
      __syncthreads();

      // E~+ = C^-T W: back substitution with the upper-triangular C^T, in place on the columns of A1 (two lanes per
      // column, blocks of 8 unknowns in registers); gq is dead since the formation of X+-: reciprocal diagonal of C
      for (int j = tid; j < h; j += NT) gq[j] = 1.0 / C2.at(j, j);
      __syncthreads();
      block_backsolve_lt(C2, A1, ld1, h, gq);
      __syncthreads();

      // store k, F = s (E~+ - E~-) / 2, G = s (E~+ + E~-) / 2
      {
        for (int j = tid; j < h; j += NT) rk[j] = sigma[j];
        SMRT_FOR_2D(a, j, h, h) {
          const size_t e = (size_t)j * h + a;
          double ep = SMRT_AT(A1, ld1, a, j), em = rG[e];
          double s = 0.5 * sdiag[a];
          rF[e] = s * (ep - em);
          rG[e] = s * (ep + em);
        }
        if (tid < 32) {
          double mn = 1e300;
          for (int j = tid; j < h; j += 32) mn = fmin(mn, sigma[j]);
          for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off, 32));
          if (tid == 0) kmin[m] = mn;
        }
      }
      __syncthreads();
    }
    if (tid == 0) A.scat_flag[bl] = failed ? 0 : 1;
    __syncthreads();
  }
}

// --------------------------------------------------------------------------------------------------------------------
// kernel 3: boundary system, mode summation, output stage
// --------------------------------------------------------------------------------------------------------------------
// vector region (doubles): mu[n] outmu[n] outw[n] kvec tvec Rt Tt Rb Tb Ttprev ipiv RbD Dsg Tup [hmax each]
//                          acc[9 * SMRT_MAX_INC] coh[4 * SMRT_MAX_INC] ; ints: rowstep[hmax] rowof[hmax] inc[SMRT_MAX_INC]
// (the accumulators of the active mode exist only in active plans)
SMRT_HD size_t boundary_vec_doubles(int n, int hmax, int mode = 1) {
  const size_t nacc = (mode == 1) ? 13 * SMRT_MAX_INC : 0;
  return ((size_t)3 * n + 11 * hmax + nacc + hmax /* two int arrays */ + SMRT_MAX_INC + 16 + 1) & ~(size_t)1;
}
// matrix region: BF, BG (compact), scratch of the blocked Gauss-Jordan (V, reciprocal pivots), BR (ld odd),
// T = [left | right | rhs] (ld odd), btop, svec, ytr, vvec
SMRT_HD size_t boundary_gj_doubles(int hmax, int nrhs_max, bool compact = false) {
  (void)nrhs_max;
  // V (double buffered; the same buffer is the scratch of the block matvecs: 2 * 8 * h doubles), reciprocal pivots;
  // compact: 2 * SMRT_GJ_NB * h = 8 h doubles, enough for the matvec scratch of blocks of <= 256 threads as well
  return (size_t)(compact ? 8 : 16) * hmax + ((hmax + 1) & ~1);
}
// stream = the instantiation that keeps only [T | R] resident and stages F and G into them (two CTAs per SM)
SMRT_HD size_t boundary_mat_doubles(int hmax, int nrhs_max, bool stream = false) {
  return (stream ? 0 : (size_t)2 * (((size_t)hmax * hmax + 1) & ~(size_t)1)) + boundary_gj_doubles(hmax, nrhs_max, stream) +
         (size_t)hmax * (hmax + 1) + (size_t)(hmax + 1) * (2 * hmax + nrhs_max) + 4 * (size_t)hmax * nrhs_max + 16;
}

// kMid instantiation (64 < h <= 128, 512 threads, one CTA per SM): shared memory holds ONE h x h matrix (the block being
// factorised in product form, then the B operand of the products), the staging panels of the GEMMs, the pivot-row
// exchange buffer and the right-hand sides; the other h x h blocks live in an L2-resident per-CTA global scratch.
SMRT_HD int boundary_mid_ld(int h) { return (h + 7) & ~7; }
#define SMRT_MID_STAGE 6144  // doubles of the GEMM staging ring (3 panels; inside the arena, placed per layer)
// `arena`: doubles of the matrix arena (>= hmax * ld + SMRT_MID_STAGE: the resident matrix M1 and the staging panels of
// the GEMMs; whatever shared memory is left goes to it,
// so that layers whose TWO blocks [A21 | A22 | rhs] fit — most layers keep fewer streams than n_max_stream — are
// eliminated in shared memory like the h <= 64 instantiations, with the right block riding along)
SMRT_HD size_t boundary_mid_fixed_doubles(int hmax, int nrhs_max) {
  const size_t ldm = boundary_mid_ld(hmax);
  return (size_t)1024 + ldm /* matvec scratch, reciprocal pivots */ +
         2 * 8 * 32 * 4 /* pivot-row exchange / V buffer */ + ldm * nrhs_max /* rhs block */ +
         4 * (size_t)hmax * nrhs_max + 16;
}
SMRT_HD size_t boundary_mid_smem_doubles(int hmax, int nrhs_max, size_t arena) {
  return boundary_mid_fixed_doubles(hmax, nrhs_max) + arena;
}
// global scratch: gA (A22 -> S), gB (Y~ -> R of the stack), gC (K), gF / gG (generated operands of non-scattering layers)
SMRT_HD size_t boundary_mid_scratch_doubles(int hmax) {
  return 3 * (size_t)hmax * boundary_mid_ld(hmax) + 2 * (size_t)hmax * hmax + 16;
}

struct BoundaryCtx {
  // per problem
  int b, nl, n_air, n_incs, npol_out;
  double freq;
  cplx eps_star;
};

// kStreamFG: the operands F and G of a layer are not kept in their own buffers but staged (TMA bulk copies from the
// layer record, L2-resident after the first touch) into the buffers the results of the step will occupy, and the
// products keep all their tiles in registers until the block has finished reading (block_gemm_*_deferred): 114 KB
// per problem at 32 streams instead of 186 KB, TWO problems per SM: the serial panel chain of one elimination is
// covered by the other problem's work.  Requires h <= 64 and 128 <= blockDim.x <= 256.
#ifdef SMRT_SIMT_EMULATION
#define SMRT_PHASE(id)
#else
// thread 0 charges the cycles since the previous mark to phase `id` (only when the plan was created with
// SMRT_B200_PROFILE set: A.prof != NULL)
#define SMRT_PHASE(id)                                             \
  if (A.prof && threadIdx.x == 0) {                                \
    const long long now_ = clock64();                              \
    atomicAdd(&A.prof[id], (unsigned long long)(now_ - prof_t0));  \
    prof_t0 = now_;                                                \
  }
#endif
// kRough: instantiation for batches that carry interface_params (rough interfaces).  The code of the rough surfaces
// is compiled into these instantiations only: in the kernel body it perturbed the register allocation of every batch
// by 1.2 - 1.7 % (profiles/r04_notes.txt), so batches of flat / transparent interfaces run kernels without it.
template <bool kGlobalScratch, int kMaxThreads, bool kStreamFG = false, bool kMid = false, bool kRough = false>
SMRT_GLOBAL void __launch_bounds__(kMaxThreads, kStreamFG ? 2 : 1) boundary_kernel(KArgs A) {
  SMRT_DYN_SMEM(smem);
  SMRT_SHARED int s_item;
  SMRT_SHARED int s_ctrl[8];
  SMRT_SHARED double s_tau;
  SMRT_SHARED int s_lend;
  SMRT_SHARED smrt_mbar_t s_mbar;  // completion barrier of the TMA bulk prefetch of the next layer's (F, G) record
  SMRT_SHARED smrt_mbar_t s_mbarG;  // kStreamFG: s_mbar completes the copies of F, s_mbarG those of G
  const int tid = threadIdx.x;
  const int NT = blockDim.x;
  const int n = A.n;
  const int hmax = smrt_npol(A.m_max) * n;
  const int nrhs_max = (A.mode == 0) ? 1 : 3 * 2 * A.n_inc;
  if (tid == 0) {
    smrt_mbar_init(&s_mbar, 1);
    smrt_mbar_init(&s_mbarG, 1);
  }
  unsigned pf_parity = 0;  // parity of the next phase to wait for
  unsigned pg_parity = 0;  // ... of s_mbarG
  __syncthreads();

  double* mu = smem;
  double* outmu = mu + n;
  double* outw = outmu + n;
  double* kvec = outw + n;
  double* tvec = kvec + hmax;
  double* Rt = tvec + hmax;
  double* Tt = Rt + hmax;
  double* Rb = Tt + hmax;
  double* Tb = Rb + hmax;
  double* Ttprev = Tb + hmax;
  double* ipiv = Ttprev + hmax;
  double* RbD = ipiv + hmax;
  double* Dsg = RbD + hmax;
  double* Tup = Dsg + hmax;                       // transmission of the layer above's emission into this layer
  const int nacc = (A.mode == 1) ? SMRT_MAX_INC : 0;  // the active-mode accumulators exist only in active plans
  double* acc_act = Tup + hmax;                   // [3][3][SMRT_MAX_INC]
  double* coh_act = acc_act + 9 * nacc;           // [2][2][SMRT_MAX_INC]
  int* rowstep = reinterpret_cast<int*>(coh_act + 4 * nacc);
  int* rowof = rowstep + hmax;
  int* inc = rowof + hmax;
  double* mats = kGlobalScratch ? (A.scratch + (size_t)blockIdx.x * A.scratch_stride)
                                : (smem + boundary_vec_doubles(n, hmax, A.mode));
  const size_t szc = ((size_t)hmax * hmax + 1) & ~(size_t)1, szp = (size_t)hmax * (hmax + 1);
  const size_t szr = (size_t)hmax * nrhs_max;
  // BF / BG: own buffers, or (kStreamFG) aliases of the staging areas inside T and R, set per layer and per phase
  double* BF = mats;  // 16-byte aligned (vector loads of the layer records)
  double* BG = BF + szc;
  double* GJV = kStreamFG ? mats : BG + szc;               // blocked Gauss-Jordan: V (2 x h x NB)
  double* pivinv = GJV + (size_t)(kStreamFG ? 8 : 16) * hmax;  // ... reciprocal pivots
  double* BR = GJV + boundary_gj_doubles(hmax, nrhs_max, kStreamFG);
  double* TT = BR + szp;  // h x (2h + nrhs), ld = ldp
  double* btop = TT + (size_t)(hmax + 1) * (2 * hmax + nrhs_max);
  double* svec = btop + szr;
  double* ytr = svec + szr;
  double* vvec = ytr + szr;
  // kMid: one resident matrix, the rest in the global scratch
  const int ldm_max = boundary_mid_ld(hmax);
  double* M1 = nullptr;
  double* stage = nullptr;
  double* xch = nullptr;
  double* rhsM = nullptr;
  double *gA = nullptr, *gB = nullptr, *gC = nullptr, *gF = nullptr, *gG = nullptr;
  if (kMid) {
    GJV = mats;  // scratch of the block matrix-vector products (1024 doubles)
    pivinv = GJV + 1024;
    rhsM = pivinv + ldm_max;
    btop = rhsM + (size_t)ldm_max * nrhs_max;
    svec = btop + szr;
    ytr = svec + szr;
    vvec = ytr + szr;
    M1 = vvec + szr;  // the matrix arena: [M1 | staging] or [M1 | right block | rhs] (A.mid_arena doubles)
    xch = M1 + A.mid_arena;
    gA = A.scratch + (size_t)blockIdx.x * A.scratch_stride;
    gB = gA + (size_t)hmax * ldm_max;
    gC = gB + (size_t)hmax * ldm_max;
    gF = gC + (size_t)hmax * ldm_max;
    gG = gF + (size_t)hmax * hmax;
    BR = gB;
  }

  long long prof_t0 = 0;
#ifndef SMRT_SIMT_EMULATION
  prof_t0 = clock64();
#endif
  (void)prof_t0;
  for (;;) {
    SMRT_PHASE(0)  // between problems: output stage of the previous one
    if (tid == 0) s_item = atomicAdd(&A.counters[1], 1);
    __syncthreads();
    const int b = s_item;
    __syncthreads();
    if (b >= A.B) break;
    const int nl = A.nlayer[b];
    const size_t bL = (size_t)b * A.L;
    const double freq = A.frequency[b];
    const int n_out = (A.mode == 0) ? 2 * A.n_theta : 9 * A.n_inc;
    double* out = A.values + (size_t)b * n_out;
    if (nl <= 0 || (A.status[b] & 15) != 0) {
      for (int i = tid; i < n_out; i += NT) out[i] = (nl <= 0) ? 0.0 : nan("");
      if (tid == 0) {
        A.n_streams_out[b] = 0;
        A.optical_depth[b] = 0.0;
      }
      continue;
    }
    const double* eps_b = A.eps_eff + bL * 2;
    const int kstar = most_refringent_layer(eps_b, nl);
    const cplx eps_star = c_make(eps_b[2 * kstar], eps_b[2 * kstar + 1]);
    const cplx eps0 = c_make(eps_b[0], eps_b[1]);
    // air streams ----------------------------------------------------------------------------- streams.py:156, 212-218
    const double rindex_air = c_sqrt(eps_star).re;
    const int n_air = stream_count(rindex_air, A.gl_mu, n);
    for (int j = tid; j < n_air; j += NT) outmu[j] = stream_mu(rindex_air, A.gl_mu, j);
    __syncthreads();
    if (n_air < 2) {
      for (int i = tid; i < n_out; i += NT) out[i] = nan("");
      if (tid == 0) {
        set_error(A.status, b, ST_INPUT);
        A.n_streams_out[b] = 0;
        A.optical_depth[b] = 0.0;
      }
      __syncthreads();
      continue;
    }
    for (int j = tid; j < n_air; j += NT) outw[j] = stream_weight(outmu, n_air, j);
    if (A.substrate_kind[b] == SUB_ROUGH_CHOUDHURY &&
        substrate_ksigma(freq, c_make(eps_b[2 * (nl - 1)], eps_b[2 * (nl - 1) + 1]),
                         A.substrate_params ? A.substrate_params[4 * (size_t)b] : 0.0) > 0.1) {
      // outside the validity range of the model: the reference raises (rough_choudhury79.py:29-31)
      for (int i = tid; i < n_out; i += NT) out[i] = nan("");
      if (tid == 0) {
        set_error(A.status, b, ST_SUBSTRATE);
        A.n_streams_out[b] = 0;
        A.optical_depth[b] = 0.0;
      }
      __syncthreads();
      continue;
    }
    // isotropic atmosphere (passive mode): downwelling / upwelling intensity and transmittance, identical for every
    // air stream and polarisation (atmosphere/simple_isotropic_atmosphere.py:55-77, core/atmosphere.py:134-162)
    double atm_down = 0.0, atm_up = 0.0, atm_trans = 1.0;
    if (A.mode == 0 && A.atmosphere) {
      atm_down = planck_function(freq, A.atmosphere[3 * (size_t)b], A.rayleigh_jeans);
      atm_up = planck_function(freq, A.atmosphere[3 * (size_t)b + 1], A.rayleigh_jeans);
      atm_trans = A.atmosphere[3 * (size_t)b + 2];
    }
    // incident streams (active) --------------------------------------------------------- rtsolver_utils.py:91-107
    if (tid == 0) {
      int cnt = 0;
      if (A.mode == 1) {
        for (int t = 0; t < A.n_inc; ++t) {
          double mu_inc = cos(A.theta_inc[t]);
          int i0 = 0;
          while (i0 < n_air && outmu[i0] > mu_inc) ++i0;  // searchsorted(-outmu, -mu_inc), left
          int cand[2], nc = 0;
          if (i0 == 0)
            cand[nc++] = 0;
          else if (i0 == n_air)
            cand[nc++] = n_air - 1;
          else {
            cand[nc++] = i0 - 1;
            cand[nc++] = i0;
          }
          for (int c = 0; c < nc; ++c) {
            bool found = false;
            for (int e = 0; e < cnt; ++e) found = found || (inc[e] == cand[c]);
            if (!found && cnt < SMRT_MAX_INC) inc[cnt++] = cand[c];
          }
        }
        // sort ascending (insertion sort, cnt <= 16)
        for (int i = 1; i < cnt; ++i) {
          int v = inc[i], j = i - 1;
          while (j >= 0 && inc[j] > v) {
            inc[j + 1] = inc[j];
            --j;
          }
          inc[j + 1] = v;
        }
      }
      s_ctrl[3] = cnt;
    }
    __syncthreads();
    const int n_incs = s_ctrl[3];
    for (int i = tid; i < 9 * nacc; i += NT) acc_act[i] = 0.0;
    for (int i = tid; i < 4 * nacc; i += NT) coh_act[i] = 0.0;
    __syncthreads();

    const int nruns = (A.mode == 0) ? 1 : (A.m_max + 2);  // active: coherent pass + modes 0..m_max
    bool failed = false;
    double tau_report = 0.0;
    bool shallow = false;

    for (int run = 0; run < nruns && !failed; ++run) {
      const bool coherent = (A.mode == 1 && run == 0);
      const int m = (A.mode == 0) ? 0 : (run == 0 ? 0 : run - 1);
      const int npol = smrt_npol(m);
      const int mdiff_max = (A.mode == 0) ? 0 : A.m_max;  // modes the diffuse backscatter of rough surfaces is spread over
      const int nrhs = (A.mode == 0) ? 1 : npol * n_incs;

      // optical depth, top-down, and the last layer kept ------------------------------------------ dort.py:444-452
      if (tid == 0) {
        double tau = 0.0;
        int lend = nl - 1;
        for (int l = 0; l < nl; ++l) {
          double kmn;
          if (!coherent && A.scat_flag[bL + l]) {
            kmn = A.kmin[(bL + l) * SMRT_MAX_MODES + m];
          } else {
            double ri = real_index_of(eps_star, c_make(eps_b[2 * l], eps_b[2 * l + 1]));
            kmn = (A.ks[bL + l] + A.ka[bL + l]) / stream_mu(ri, A.gl_mu, 0);
          }
          tau += kmn * A.thickness[bL + l];
          if (A.prune_tau > 0.0 && tau > A.prune_tau) {
            lend = l;
            break;
          }
        }
        s_tau = tau;
        s_lend = lend;
      }
      __syncthreads();
      const int l_end = s_lend;
      if (run == ((A.mode == 0) ? 0 : 1)) tau_report = s_tau;  // the mode-0 (scattering) run
      if (A.substrate_kind[b] == SUB_NONE && s_tau < 5.0) shallow = true;

      int h_prev = 0, ldr_prev = 1;
      bool have_prev = false;
      bool src_prev = false;  // the stack below carries a source vector (false for the source-free active layers)
      int pf_layer = -1;      // layer whose (F, G) record is in flight into BF / BG (TMA bulk copy), -1 = none

      SMRT_PHASE(1)  // problem setup
      for (int l = l_end; l >= 0 && !failed; --l) {
        const cplx eps_l = c_make(eps_b[2 * l], eps_b[2 * l + 1]);
        const double rindex = real_index_of(eps_star, eps_l);
        const int n_l = stream_count(rindex, A.gl_mu, n);
        const int h = npol * n_l;
        const int ldp = (h & 1) ? h : h + 1;
        const double thick = A.thickness[bL + l];
        const double ke = A.ks[bL + l] + A.ka[bL + l];
        const bool scat = (!coherent) && (A.scat_flag[bL + l] != 0);
        const int nr = (A.mode == 1 && l > 0) ? 0 : nrhs;  // active mode: sources only at the air-snow interface
        // threads used by the dual h x h product: a 4 x 4 tile of each product per thread
        const int gemm_thr = (16 * ((h + 3) / 4) < NT) ? 16 * ((h + 3) / 4) : NT;
        for (int j = tid; j < n_l; j += NT) mu[j] = stream_mu(rindex, A.gl_mu, j);
        __syncthreads();

        // eigen data: (k, F, G) from the workspace, or the trivial solution F = I, G = 0, k = ke / mu
        // (the eigenvalues are not part of the prefetch: their load is issued here and consumed after the Fresnel
        // evaluations, so that its DRAM latency overlaps them)
        const double* rk_l = A.eig + (bL + l) * A.eig_stride + A.eig_off[m];
        const double kpre = (scat && tid < h) ? rk_l[tid] : 0.0;
        // kStreamFG: where the operands of the layer come from, and the staging helper.  An operand is either copied by
        // the TMA engine (even h: 16-byte aligned destinations; completion on `bar`, returns true = wait needed),
        // copied by the threads (odd h), or generated (non-scattering layer: F = I, G = 0).  The destination must be
        // dead for every thread of the block (a block barrier precedes every call) and the caller places a block
        // barrier before the first use of a thread-written copy.
        const double* rF_l = rk_l + smrt_even(smrt_npol(m) * n);
        const double* rG_l = rF_l + smrt_even((long long)h * h);
        auto stage_operand = [&](double* dst, const double* src, bool identity, smrt_mbar_t* bar) -> bool {
#ifndef SMRT_STREAM_NO_TMA
          if (scat && (h & 1) == 0) {
            if (tid == 0) smrt_bulk_load1(bar, dst, src, (unsigned)((size_t)h * h * sizeof(double)));
            return true;
          }
#endif
          if (scat) {
            for (int e = tid; e < h * h; e += NT) dst[e] = src[e];
          } else {
            SMRT_FOR_2D(i, j, h, h) { dst[(size_t)j * h + i] = (identity && i == j) ? 1.0 : 0.0; }
          }
          return false;
        };
        bool waitF = false, waitG = false;
        if (kMid) {
          // the operands stay where they are (layer record in global memory, L2 after the first touch); a non-scattering
          // layer gets F = I, G = 0 generated in the global scratch
          if (scat) {
            BF = const_cast<double*>(rF_l);
            BG = const_cast<double*>(rG_l);
          } else {
            BF = gF;
            BG = gG;
            SMRT_FOR_2D(i, j, h, h) {
              gF[(size_t)j * h + i] = (i == j) ? 1.0 : 0.0;
              gG[(size_t)j * h + i] = 0.0;
            }
            for (int a = tid; a < h; a += NT) kvec[a] = ke / mu[a / npol];
          }
        } else if (kStreamFG) {
          // formation phase: F in the left block of T, G in the right block (compact, ld = h)
          BF = TT;
          BG = TT + (size_t)h * ldp;
          if (pf_layer == l) {  // F was prefetched while the layer below was being eliminated
            waitF = true;
            pf_layer = -1;
          } else {
            waitF = stage_operand(BF, rF_l, true, &s_mbar);
          }
          waitG = stage_operand(BG, rG_l, false, &s_mbarG);
          if (!scat)
            for (int a = tid; a < h; a += NT) kvec[a] = ke / mu[a / npol];
        } else if (scat) {
          const double* rk = rk_l;
          const double* rF = rk + smrt_even(smrt_npol(m) * n);
          const double* rG = rF + smrt_even((long long)h * h);
          if (pf_layer == l) {  // prefetched by the TMA engine while the layer below was being eliminated
            smrt_mbar_wait(&s_mbar, pf_parity);
            pf_parity ^= 1u;
            pf_layer = -1;
          } else {  // 4 independent 16-byte loads in flight per thread (records 16-byte aligned, h*h padded even)
            const int n2 = (h * h + 1) >> 1;
            const double2* sF = reinterpret_cast<const double2*>(rF);
            const double2* sG = reinterpret_cast<const double2*>(rG);
            double2* dF = reinterpret_cast<double2*>(BF);
            double2* dG = reinterpret_cast<double2*>(BG);
            int e = tid;
            for (; e + NT < n2; e += 2 * NT) {
              double2 f0 = sF[e], f1 = sF[e + NT], g0 = sG[e], g1 = sG[e + NT];
              dF[e] = f0;
              dF[e + NT] = f1;
              dG[e] = g0;
              dG[e + NT] = g1;
            }
            if (e < n2) {
              dF[e] = sF[e];
              dG[e] = sG[e];
            }
          }
        } else {
          SMRT_FOR_2D(i, j, h, h) {
            BF[(size_t)j * h + i] = (i == j) ? 1.0 : 0.0;
            BG[(size_t)j * h + i] = 0.0;
          }
          for (int a = tid; a < h; a += NT) kvec[a] = ke / mu[a / npol];
        }
        // interface coefficients on this layer's streams ------------------------ rtsolver_utils.py:473-644 (flat only)
        // one Fresnel evaluation per thread: (stream j) x (top of the layer | bottom of the layer | transmission of
        // the emission of the layer above into this layer, on the upper layer's streams: dort.py:383-395)
        const bool thermal = (A.mode == 0 && m == 0);
        // (`which` is uniform within a warp: the three branches never serialise)
        for (int e = tid; e < 3 * 32 * ((n_l + 31) >> 5); e += NT) {
          const int span = 32 * ((n_l + 31) >> 5);
          const int which = e / span, j = e - which * span;
          if (j >= n_l) continue;
          if (which == 0) {
            cplx eps_up = (l > 0) ? c_make(eps_b[2 * (l - 1)], eps_b[2 * (l - 1) + 1]) : c_make(1.0, 0.0);
            FresnelRT ft = fresnel_power(A.interface_kind[bL + l], eps_l, eps_up, mu[j]);
            for (int p = 0; p < npol; ++p) {
              Rt[j * npol + p] = ft.R[p];
              Tt[j * npol + p] = ft.T[p];
              Dsg[j * npol + p] = (p == 2) ? -1.0 : 1.0;
            }
          } else if (which == 1) {
            FresnelRT fb;
            if (l < nl - 1) {
              fb = fresnel_power(A.interface_kind[bL + l + 1], eps_l, c_make(eps_b[2 * (l + 1)], eps_b[2 * (l + 1) + 1]),
                                 mu[j]);
            } else if (A.substrate_kind[b] != SUB_NONE) {
              // the diagonal diffuse (backscatter) reflection of the rough substrates joins the specular one in every
              // pass but the coherent one (rtsolver_utils.py:690-709); the emissivity keeps the coherent value
              fb = substrate_power(A.substrate_kind[b], A.substrate_params ? A.substrate_params + 4 * (size_t)b : nullptr,
                                   freq, eps_l, c_make(A.substrate_eps[2 * b], A.substrate_eps[2 * b + 1]), mu[j],
                                   (n_l >= 2) ? stream_weight(mu, n_l, j) : 0.0, coherent ? -1 : m, mdiff_max);
            } else {
              fb.R[0] = fb.R[1] = fb.R[2] = 0.0;
              fb.T[0] = fb.T[1] = fb.T[2] = 0.0;
            }
            for (int p = 0; p < npol; ++p) {
              Rb[j * npol + p] = fb.R[p];
              Tb[j * npol + p] = fb.T[p];
              RbD[j * npol + p] = (p == 2) ? -fb.R[p] : fb.R[p];
            }
          } else {
            double tu[3] = {0.0, 0.0, 0.0};
            if (thermal && l > 0 && A.temperature[bL + l - 1] > 0.0) {
              cplx eps_up = c_make(eps_b[2 * (l - 1)], eps_b[2 * (l - 1) + 1]);
              double ri_up = real_index_of(eps_star, eps_up);
              if (j < stream_count(ri_up, A.gl_mu, n)) {
                FresnelRT fu = fresnel_power(A.interface_kind[bL + l], eps_up, eps_l, stream_mu(ri_up, A.gl_mu, j));
                tu[0] = fu.T[0];
                tu[1] = fu.T[1];
                tu[2] = fu.T[2];
              }
            }
            for (int p = 0; p < npol; ++p) Tup[j * npol + p] = tu[p];
          }
        }
        if constexpr (kRough) if (A.interface_params) {
          // rough interfaces (block-uniform: the batch carries parameters only when a snowpack has one)
          const int ik_top = A.interface_kind[bL + l], ik_bot = (l < nl - 1) ? A.interface_kind[bL + l + 1] : IF_FLAT;
          if (ik_top >= IF_IEM_FUNG92 || ik_bot >= IF_IEM_FUNG92) {
            RoughFix rf;
            rf.ik_top = ik_top, rf.ik_bot = ik_bot, rf.par_top = SMRT_IPAR(l), rf.par_bot = SMRT_IPAR(l + 1);
            rf.freq = freq, rf.eps_l = eps_l, rf.eps_star = eps_star, rf.eps_b = eps_b, rf.l = l, rf.nl = nl, rf.n_l = n_l;
            rf.n = n, rf.npol = npol, rf.m_diff = coherent ? -1 : m, rf.m_max = mdiff_max, rf.mu = mu, rf.gl_mu = A.gl_mu;
            rf.up_emits = thermal && l > 0 && A.temperature[bL + (l > 0 ? l - 1 : 0)] > 0.0;
            rf.Rt = Rt, rf.Tt = Tt, rf.Rb = Rb, rf.Tb = Tb, rf.RbD = RbD, rf.Tup = Tup;
            rough_interface_fixup(rf);
          }
        }
        if (scat) {
          if (tid < h) kvec[tid] = kpre;
          for (int a = tid + NT; a < h; a += NT) kvec[a] = rk_l[a];
        }
        if (kStreamFG) {
          if (waitF) {
            smrt_mbar_wait(&s_mbar, pf_parity);
            pf_parity ^= 1u;
          }
          if (waitG) {
            smrt_mbar_wait(&s_mbarG, pg_parity);
            pg_parity ^= 1u;
          }
        }
        __syncthreads();
        SMRT_PHASE(2)  // layer head: streams, record, Fresnel
        for (int a = tid; a < h; a += NT) tvec[a] = exp(-kvec[a] * thick);

        // right-hand sides ---------------------------------------------------------------------- dort.py:375-441
        const int r = have_prev ? (h < h_prev ? h : h_prev) : 0;
        const int ldm = boundary_mid_ld(h);            // kMid: leading dimension of the resident matrix and of the rhs block
        const int ldt = kMid ? ldm : ldp;
        // kMid: both blocks and the right-hand sides fit in the arena -> the right block rides along (resident path)
        // (the staging panels of the GEMMs then lie over the right block, which is dead or not yet formed whenever a GEMM
        // streams its operands; blocks so small that the panels would reach the right-hand sides get them behind)
        const long long need2 = (long long)(2 * h + nr) * ldm;
        const bool stage_behind = (long long)h * ldm < SMRT_MID_STAGE;
        const bool resident = kMid && need2 + (stage_behind ? SMRT_MID_STAGE : 0) <= A.mid_arena;
        double* R2 = kMid ? M1 + (size_t)h * ldm : nullptr;
        if (kMid) stage = (resident && stage_behind) ? M1 + ((need2 + 1) & ~1LL) : R2;
        double* Trhs = kMid ? (resident ? R2 + (size_t)h * ldm : rhsM)
                            : TT + (size_t)(2 * h) * ldp;  // b_bot lives in the augmented columns of T
        if (nr > 0) {
          const double Tl = A.temperature[bL + l];
          const double Bl = (thermal && Tl > 0.0) ? planck_function(freq, Tl, A.rayleigh_jeans) : 0.0;
          for (int e = tid; e < h * nr; e += NT) {
            int a = e % h, c = e / h;
            int j = a / npol, p = a % npol;
            double vt = 0.0, vb = 0.0;
            if (thermal) {
              if (Tl > 0.0) {
                vt -= (1.0 - Rt[a]) * Bl;
                vb -= (1.0 - Rb[a]) * Bl;
              }
              if (l > 0) {  // emission of the layer above transmitted into this layer (its own streams, truncated)
                const double Tabove = A.temperature[bL + l - 1];
                if (Tabove > 0.0) vt += Tup[a] * planck_function(freq, Tabove, A.rayleigh_jeans);
              }
              if (l < l_end && a < r) {
                double Tdn = A.temperature[bL + l + 1];
                if (Tdn > 0.0) vb += Ttprev[a] * planck_function(freq, Tdn, A.rayleigh_jeans);
              }
              if (l == nl - 1 && A.substrate_kind[b] != SUB_NONE) {
                vb += Tb[a] * planck_function(freq, A.substrate_temperature[b], A.rayleigh_jeans);
              }
            }
            if (l < l_end && a < r && src_prev) vb += Ttprev[a] * SMRT_AT(svec, h_prev, a, c);
            SMRT_AT(btop, h, a, c) = vt;
            SMRT_AT(Trhs, ldt, a, c) = vb;
          }
        }
        __syncthreads();
        if (l == 0 && A.mode == 0 && atm_down != 0.0) {
          // downwelling atmospheric radiation through the air-snow interface (dort.py:383-395): b_top += T_air I_down
          // on the air streams; rows beyond the layer's streams are truncated
          for (int a = tid; a < 2 * n_air && a < h; a += NT) {
            FresnelRT fa = (kRough ? air_interface_power(A, bL, freq, eps0, outmu[a >> 1], 0.0, -1, 0)
                                         : fresnel_power(A.interface_kind[bL], c_make(1.0, 0.0), eps0, outmu[a >> 1]));
            btop[a] += fa.T[a & 1] * atm_down;
          }
        }
        if (l == 0 && A.mode == 1) {
          // incident beams (rtsolver_utils.py:109-135) through the air-snow interface: b_top += T_air I_down
          for (int c = tid; c < nr; c += NT) {
            int jinc = c / npol, ipol = c % npol;
            int i = inc[jinc];
            if (i < n_l) {  // rows beyond the layer's streams are truncated (dort.py:391-395)
              double power = 1.0 / (2.0 * SMRT_PI * outw[i]);
              if (m > 0) power *= 2.0;
              FresnelRT fa = (kRough ? air_interface_power(A, bL, freq, eps0, outmu[i], 0.0, -1, 0)
                                         : fresnel_power(A.interface_kind[bL], c_make(1.0, 0.0), eps0, outmu[i]));
              SMRT_AT(btop, h, i * npol + ipol, c) += fa.T[ipol] * power;
            }
          }
        }
        if constexpr (kMid) {
          // ------------------------------------------------------------------------------------------------------------
          // 64 < h <= 128: one resident h x h matrix (M1), product-form eliminations, every other block in the L2-resident
          // global scratch; same algebra as the resident path below
          // ------------------------------------------------------------------------------------------------------------
          const double* Fo = BF;
          const double* Go = BG;
          int* kof = rowstep;  // inverse of the pivot order: kof[rowof[k]] = k
          // coupling operator of the stack below: R' = T_top(l+1) R(l+1) T_bottom(l) D on the common streams
          for (int k = tid; k < r; k += NT) ipiv[k] = Tb[k] * Dsg[k];  // (ipiv is free until the elimination)
          __syncthreads();
          scale_block_global(gB, ldr_prev, r, Ttprev, ipiv);
          __syncthreads();
          // A21 = F - Rb D G - R' G -> M1 ;  A22 = (G - Rb D F - R' F) t -> gA
          // (tile maps by block size: 32-row slabs x 16-column groups without padding)
          auto fg = [&](int i, int j, double& f, double& g) {  // the layer's F and G at (i, j), from the record
            const size_t e = (size_t)j * h + i;
            f = Fo[e];
            g = Go[e];
          };
          auto form21 = [&](int i, int j, double acc, double, double f, double g) {
            M1[(size_t)j * ldm + i] = f - RbD[i] * g - acc;
          };
          auto form22 = [&](int i, int j, double acc, double, double f, double g) {
            (resident ? R2 : gA)[(size_t)j * ldm + i] = (g - RbD[i] * f - acc) * tvec[j];
          };
          if (h <= 64) {
            mid_gemm<false, false, 4, 2>(h, r, h, 0, r, gB, nullptr, ldr_prev, Go, h, stage, fg, form21);
            mid_gemm<false, false, 4, 2>(h, r, h, 0, r, gB, nullptr, ldr_prev, Fo, h, stage, fg, form22);
          } else if (h <= 96) {
            mid_gemm<false, false, 6, 3>(h, r, h, 0, r, gB, nullptr, ldr_prev, Go, h, stage, fg, form21);
            mid_gemm<false, false, 6, 3>(h, r, h, 0, r, gB, nullptr, ldr_prev, Fo, h, stage, fg, form22);
          } else {
            mid_gemm<false, false, 8, 4>(h, r, h, 0, r, gB, nullptr, ldr_prev, Go, h, stage, fg, form21);
            mid_gemm<false, false, 8, 4>(h, r, h, 0, r, gB, nullptr, ldr_prev, Fo, h, stage, fg, form22);
          }
          __syncthreads();
          SMRT_PHASE(3)  // right-hand sides, formation of [A21 | A22]
          if (resident) {
            // [A21 | A22 | b_bot] in shared memory: the blocked elimination of the h <= 64 instantiations, whose look-ahead
            // hides the serial panel chain behind the updates of the right block; Y~ goes straight into M1
            if (block_gj_rows_blocked_mid(M1, ldm, R2, ldm, h, h + nr, rowof, pivinv, xch, &s_ctrl[6])) {
              failed = true;
              break;
            }
            for (int k = tid; k < h; k += NT) ipiv[k] = tvec[k] * pivinv[k];
            __syncthreads();
            SMRT_FOR_2D(k, c, h, nr) { SMRT_AT(ytr, h, k, c) = SMRT_AT(Trhs, ldm, rowof[k], c) * ipiv[k]; }
            SMRT_FOR_2D(k, c, h, h) { M1[(size_t)c * ldm + k] = R2[(size_t)c * ldm + rowof[k]] * ipiv[k]; }
          } else {
            // A21 in product form (the right-hand sides ride along), then A22 in one register-resident pass
            if (block_gj_factor(M1, ldm, Trhs, ldm, h, nr, rowof, pivinv, &s_ctrl[6])) {
              failed = true;
              break;
            }
            for (int k = tid; k < h; k += NT) {
              ipiv[k] = tvec[k] * pivinv[k];
              kof[rowof[k]] = k;
            }
            __syncthreads();
            SMRT_FOR_2D(k, c, h, nr) { SMRT_AT(ytr, h, k, c) = SMRT_AT(Trhs, ldm, rowof[k], c) * ipiv[k]; }
            // Y~ = diag(t) A21^-1 A22 -> gB (the operator of the stack below is dead)
            gj_apply_all(
                M1, ldm, h, rowof, xch, [&](int i, int c) { return gA[(size_t)c * ldm + i]; },
                [&](int i, int c, double v) {
                  const int k = kof[i];
                  gB[(size_t)c * ldm + k] = v * ipiv[k];
                });
            // Y~ -> M1 by ONE bulk copy of the TMA engine (the threads' own loop exposed the L2 latency)
            smrt_fence_async_global();
            __syncthreads();
            if (tid == 0) smrt_bulk_load1(&s_mbar, M1, gB, (unsigned)((size_t)h * ldm * sizeof(double)));
            smrt_mbar_wait(&s_mbar, pf_parity);
            pf_parity ^= 1u;
          }
          __syncthreads();
          SMRT_PHASE(4)  // first elimination
          // Y~ becomes the resident B operand of  P = F - G Y~,  K = G - F Y~ ;  S = D P - Rt K -> gA,  K -> gC
          // (l > 0: both TRANSPOSED, so that R_new = K S^-1 = (S^-T K^T)^T comes out of the same row elimination)
          const bool transposed = l > 0;
          auto prod = [&](int i, int j, double c1, double c2, double f, double g) {
            const double pv = f - c1;
            const double kv = g - c2;
            const double sv = Dsg[i] * pv - Rt[i] * kv;
            const size_t o = transposed ? (size_t)i * ldm + j : (size_t)j * ldm + i;
            gA[o] = sv;
            gC[o] = kv;
          };
          if (h <= 64) {
            for (int n0 = 0; n0 < h; n0 += 32) mid_gemm<true, true, 2, 2>(h, h, h, n0, h, Go, Fo, h, M1, ldm, stage, fg, prod);
          } else if (h <= 96) {
            for (int n0 = 0; n0 < h; n0 += 48) mid_gemm<true, true, 3, 3>(h, h, h, n0, h, Go, Fo, h, M1, ldm, stage, fg, prod);
          } else {
            for (int n0 = 0; n0 < h; n0 += 64) mid_gemm<true, true, 4, 4>(h, h, h, n0, h, Go, Fo, h, M1, ldm, stage, fg, prod);
          }
          // S (or S^T) -> M1 and, on the resident path, K^T -> the right block: bulk copies by the TMA engine (the products
          // were stored by other threads: proxy fence + barrier first; M1 was their B operand), in flight under b'
          smrt_fence_async_global();
          __syncthreads();
          if (tid == 0) {
            const unsigned bytes = (unsigned)((size_t)h * ldm * sizeof(double));
            if (resident && l > 0)
              smrt_bulk_load2(&s_mbar, M1, gA, R2, gC, bytes);
            else
              smrt_bulk_load1(&s_mbar, M1, gA, bytes);
          }
          // v = F y~r ;  b' = b_top - D (G y~r) + Rt v  (in the right-hand-side block)
          if (nr == 1) {
            block_matvec_dual(h, h, Go, Fo, h, ytr, GJV, [&](int i, double c1, double c2) {
              vvec[i] = c2;
              Trhs[i] = btop[i] - Dsg[i] * c1 + Rt[i] * c2;
            });
          } else {
            if (nr > 0) {
              block_gemm_dual(NT, h, nr, h, Go, Fo, h, ytr, h, [&](int i, int c, double c1, double c2) {
                SMRT_AT(vvec, h, i, c) = c2;
                SMRT_AT(Trhs, ldm, i, c) = SMRT_AT(btop, h, i, c) - Dsg[i] * c1 + Rt[i] * c2;
              });
            }
          }
          smrt_mbar_wait(&s_mbar, pf_parity);
          pf_parity ^= 1u;
          __syncthreads();
          SMRT_PHASE(5)  // extraction, products P / K / S, b'
          if (l > 0) {
            // [S^T | K^T]: rows of S^-T K^T = columns of R_new; b' is not touched
            if (resident) {
              if (block_gj_rows_blocked_mid(M1, ldm, R2, ldm, h, h, rowof, pivinv, xch, &s_ctrl[6])) {
                failed = true;
                break;
              }
              SMRT_FOR_2D(i, k, h, h) { gB[(size_t)k * ldm + i] = R2[(size_t)i * ldm + rowof[k]] * pivinv[k]; }
            } else {
              if (block_gj_factor(M1, ldm, Trhs, ldm, h, 0, rowof, pivinv, &s_ctrl[6])) {
                failed = true;
                break;
              }
              for (int k = tid; k < h; k += NT) kof[rowof[k]] = k;
              __syncthreads();
              gj_apply_all(
                  M1, ldm, h, rowof, xch, [&](int i, int c) { return gC[(size_t)c * ldm + i]; },
                  [&](int i, int c, double v) {
                    const int k = kof[i];
                    gB[(size_t)k * ldm + c] = v * pivinv[k];  // R_new(c, k)
                  });
            }
            __syncthreads();
            SMRT_PHASE(6)  // second elimination
            if (nr == 1) {  // s = v + R_new b'
              block_matvec_dual(h, h, gB, (const double*)nullptr, ldm, Trhs, GJV,
                                [&](int i, double acc, double) { svec[i] = vvec[i] + acc; });
            } else if (nr > 0) {
              block_gemm_ptr(NT, h, nr, h, gB, ldm, [&](int c) { return Trhs + (size_t)c * ldm; },
                             [&](int i, int c, double acc) { SMRT_AT(svec, h, i, c) = SMRT_AT(vvec, h, i, c) + acc; });
            }
            for (int a = tid; a < h; a += NT) Ttprev[a] = Tt[a];
            h_prev = h;
            ldr_prev = ldm;
            have_prev = true;
            src_prev = (nr > 0);
            __syncthreads();
            SMRT_PHASE(7)  // R of the stack, source vector
          } else {
            // top layer: z = S^-1 b' by row elimination of [S | b'], then s = v + K z
            if (block_gj_rows_blocked_mid(M1, ldm, Trhs, ldm, h, nr, rowof, pivinv, xch, &s_ctrl[6])) {
              failed = true;
              break;
            }
            SMRT_FOR_2D(k, c, h, nr) { SMRT_AT(ytr, h, k, c) = SMRT_AT(Trhs, ldm, rowof[k], c) * pivinv[k]; }
            __syncthreads();
            block_gemm_ptr(NT, h, nr, h, gC, ldm, [&](int c) { return ytr + (size_t)c * h; },
                           [&](int i, int c, double acc) { SMRT_AT(svec, h, i, c) = SMRT_AT(vvec, h, i, c) + acc; });
            __syncthreads();
          }
        } else {
        // T = [A21 | A22] without the coupling term:  A21 = F - Rb D G,  A22 = (G - Rb D F) t
        // (Dsg holds the sign D of the third Stokes component, RbD = Rb D)
        if (!kStreamFG) {
          SMRT_FOR_2D(i, j, h, h) {
            const double f = BF[(size_t)j * h + i], g = BG[(size_t)j * h + i];
            SMRT_AT(TT, ldp, i, j) = f - RbD[i] * g;
            SMRT_AT(TT, ldp, i, h + j) = (g - RbD[i] * f) * tvec[j];
          }
        }
        // coupling operator of the stack below: R' = T_top(l+1) R(l+1) T_bottom(l) D on the common streams
        SMRT_FOR_2D(i, k, r, r) { SMRT_AT(BR, ldr_prev, i, k) *= Ttprev[i] * (Tb[k] * Dsg[k]); }
        __syncthreads();
        if (kStreamFG) {
          // F and G sit (compact) in the two blocks of T: every thread evaluates its tiles of
          // [F - Rb D G - R' G | (G - Rb D F - R' F) t] in registers, the block synchronises, the tiles overwrite them
          const double* Fs = BF;
          const double* Gs = BG;
          block_gemm_split_deferred(
              h, r, h, r, BR, ldr_prev, Gs, Fs,
              [&](int i, int j, double acc) {
                if (j < h) return Fs[(size_t)j * h + i] - RbD[i] * Gs[(size_t)j * h + i] - acc;
                const int jj = j - h;
                return (Gs[(size_t)jj * h + i] - RbD[i] * Fs[(size_t)jj * h + i] - acc) * tvec[jj];
              },
              [&](int i, int j, double v) { SMRT_AT(TT, ldp, i, j) = v; });
          __syncthreads();
          // R of the stack below is dead: G for the products after the elimination goes there while it runs
          BG = BR;
          waitG = stage_operand(BG, rG_l, false, &s_mbarG);
        } else if (r > 0) {
          // one product over the 2h columns [G | F]: all threads busy with 4 x 4 tiles
          block_gemm_ptr(
              NT, r, 2 * h, r, BR, ldr_prev,
              [&](int j) { return (j < h) ? BG + (size_t)j * h : BF + (size_t)(j - h) * h; },
              [&](int i, int j, double acc) {
                SMRT_AT(TT, ldp, i, j) -= (j < h) ? acc : acc * tvec[j - h];
              });
          __syncthreads();
        }
        SMRT_PHASE(3)  // right-hand sides, formation of [A21 | A22]
        // [A21 | A22 | b_bot] -> [I | Y22 | Yr] (implicit row permutation, unscaled rows)
        const bool blocked = h <= 64;  // panel-blocked elimination (register tiles); larger blocks: one step at a time
        if (blocked ? block_gj_rows_blocked<!kGlobalScratch>(TT, ldp, TT + (size_t)h * ldp, ldp, h, h + nr, rowof, pivinv, GJV,
                                            &s_ctrl[6], A.gj_single ? hmax : 0)
                    : block_gj_rows(TT, ldp, h, 2 * h + nr, rowstep, rowof)) {
          if (kStreamFG && waitG) {  // drain the copy of G in flight before leaving
            smrt_mbar_wait(&s_mbarG, pg_parity);
            pg_parity ^= 1u;
          }
          failed = true;
          break;
        }
        SMRT_PHASE(4)  // first elimination
        for (int k = tid; k < h; k += NT)
          ipiv[k] = blocked ? tvec[k] * pivinv[k] : tvec[k] / SMRT_AT(TT, ldp, rowof[k], k);
        __syncthreads();
        // Y~ = diag(t) Y22 -> left block of T ;  y~r = diag(t) Yr -> ytr
        SMRT_FOR_2D(k, c, h, h) { SMRT_AT(TT, ldp, k, c) = SMRT_AT(TT, ldp, rowof[k], h + c) * ipiv[k]; }
        SMRT_FOR_2D(k, c, h, nr) { SMRT_AT(ytr, h, k, c) = SMRT_AT(Trhs, ldp, rowof[k], c) * ipiv[k]; }
        __syncthreads();
        // P = F - G Y~, K = G - F Y~ ;  Schur S = D P - Rt K -> right block of T ;  K -> BR
        double* TS = TT + (size_t)h * ldp;
        // (blocked path, l > 0: S and K are stored TRANSPOSED, so that R_new = K S^-1 = (S^-T K^T)^T comes out of the
        // same row elimination as above)
        const bool transposed = blocked && l > 0;
        if (kStreamFG) {
          // the right block of T is dead (Y~ was extracted): F is staged there, G already sits in R
          BF = TS;
          waitF = stage_operand(BF, rF_l, true, &s_mbar);
          if (waitF) {
            smrt_mbar_wait(&s_mbar, pf_parity);
            pf_parity ^= 1u;
          }
          if (waitG) {
            smrt_mbar_wait(&s_mbarG, pg_parity);
            pg_parity ^= 1u;
          }
          __syncthreads();
        } else {
          block_gemm_dual(gemm_thr, h, h, h, BG, BF, h, TT, ldp, [&](int i, int j, double c1, double c2) {
            double pv = SMRT_AT(BF, h, i, j) - c1;
            double kv = SMRT_AT(BG, h, i, j) - c2;
            double sv = Dsg[i] * pv - Rt[i] * kv;
            if (transposed) {
              SMRT_AT(TS, ldp, j, i) = sv;
              SMRT_AT(BR, ldp, j, i) = kv;
            } else {
              SMRT_AT(TS, ldp, i, j) = sv;
              SMRT_AT(BR, ldp, i, j) = kv;
            }
          });
        }
        // v = F y~r ;  b' = b_top - D (G y~r) + Rt v  (b' goes next to S, in the augmented columns)
        if (nr == 1 && NT >= h) {  // one right-hand side: whole-block matrix-vector products (GJV is free: scratch)
          block_matvec_dual(h, h, BG, BF, h, ytr, GJV, [&](int i, double c1, double c2) {
            vvec[i] = c2;
            Trhs[i] = btop[i] - Dsg[i] * c1 + Rt[i] * c2;
          });
        } else {
          if (nr > 0) {
            block_gemm_dual(NT, h, nr, h, BG, BF, h, ytr, h, [&](int i, int c, double c1, double c2) {
              SMRT_AT(vvec, h, i, c) = c2;
              SMRT_AT(Trhs, ldp, i, c) = SMRT_AT(btop, h, i, c) - Dsg[i] * c1 + Rt[i] * c2;
            });
          }
          __syncthreads();
        }
        if (kStreamFG) {
          // (the products come after the matrix-vector part here: their results overwrite the staged operands)
          const double* Fs = BF;
          const double* Gs = BG;
          block_gemm_dual_deferred(
              h, h, h, Gs, Fs, h, TT, ldp,
              [&](int i, int j, double& c1, double& c2) {  // (G Y~, F Y~) -> (S, K)
                const double pv = SMRT_AT(Fs, h, i, j) - c1;
                const double kv = SMRT_AT(Gs, h, i, j) - c2;
                c1 = Dsg[i] * pv - Rt[i] * kv;
                c2 = kv;
              },
              [&](int i, int j, double sv, double kv) {
                if (transposed) {
                  SMRT_AT(TS, ldp, j, i) = sv;
                  SMRT_AT(BR, ldp, j, i) = kv;
                } else {
                  SMRT_AT(TS, ldp, i, j) = sv;
                  SMRT_AT(BR, ldp, i, j) = kv;
                }
              });
          // the left block of T (Y~) is dead: F of the layer above is fetched into it during the second elimination
          // when it fits there (the layers have their own stream counts); G only as an L2 prefetch hint
          if (l > 0 && !coherent && A.scat_flag[bL + l - 1] != 0) {
            const cplx eps_u = c_make(eps_b[2 * (l - 1)], eps_b[2 * (l - 1) + 1]);
            const int h_u = npol * stream_count(real_index_of(eps_star, eps_u), A.gl_mu, n);
            const double* uk = A.eig + (bL + l - 1) * A.eig_stride + A.eig_off[m];
            const double* uF = uk + smrt_even(smrt_npol(m) * n);
            const double* uG = uF + smrt_even((long long)h_u * h_u);
            const unsigned bytes = (unsigned)((size_t)h_u * h_u * sizeof(double));
#ifdef SMRT_STREAM_NO_TMA
            if (false) {
#else
            if ((h_u & 1) == 0) {
#endif
              const bool fits = (size_t)h_u * h_u <= (size_t)h * ldp;
              if (tid == 0) {
                smrt_prefetch_l2(uG, bytes);
                if (fits)
                  smrt_bulk_load1(&s_mbar, TT, uF, bytes);
                else
                  smrt_prefetch_l2(uF, bytes);
              }
              if (fits) pf_layer = l - 1;
            }
          }
        }
        if (!kStreamFG && !kGlobalScratch && l > 0 && !coherent && A.scat_flag[bL + l - 1] != 0) {
          // BF / BG are dead from here on: let the TMA engine fetch the record of the layer above (cp.async.bulk,
          // completion on s_mbar) while this CTA runs the second elimination
          if (tid == 0) {
            const cplx eps_u = c_make(eps_b[2 * (l - 1)], eps_b[2 * (l - 1) + 1]);
            const int h_u = npol * stream_count(real_index_of(eps_star, eps_u), A.gl_mu, n);
            const double* uk = A.eig + (bL + l - 1) * A.eig_stride + A.eig_off[m];
            const double* uF = uk + smrt_even(smrt_npol(m) * n);
            const double* uG = uF + smrt_even((long long)h_u * h_u);
            smrt_bulk_load2(&s_mbar, BF, uF, BG, uG, (unsigned)(smrt_even((long long)h_u * h_u) * sizeof(double)));
          }
          pf_layer = l - 1;
        }
        SMRT_PHASE(5)  // extraction, products P / K / S, b'
        double* bounce = kStreamFG ? TS : TT;
        if (l > 0) {
          // keep b' (the column elimination below does not touch the augmented columns)
          // R_new = K S^-1 by column elimination of [S; K]
          if (transposed) {
            // [S^T | K^T] -> rows of S^-T K^T = columns of R_new
            if (block_gj_rows_blocked<!kGlobalScratch>(TS, ldp, BR, ldp, h, h, rowof, pivinv, GJV, &s_ctrl[6], A.gj_single ? hmax : 0)) {
              failed = true;
              break;
            }
            // (un-permuted through a dead block of T: the left one, or the right one when the left one receives the
            // prefetched F of the layer above)
            SMRT_PHASE(6)  // second elimination
            SMRT_FOR_2D(i, k, h, h) { SMRT_AT(bounce, ldp, i, k) = SMRT_AT(BR, ldp, rowof[k], i) * pivinv[k]; }
          } else {
            if (block_gj_cols(TS, ldp, BR, ldp, h, h, rowstep, rowof)) {
              failed = true;
              break;
            }
            for (int k = tid; k < h; k += NT) ipiv[k] = 1.0 / SMRT_AT(TS, ldp, k, rowof[k]);
            __syncthreads();
            // un-permute / scale through the (dead) left block of T, then back into BR
            SMRT_FOR_2D(i, k, h, h) { SMRT_AT(bounce, ldp, i, k) = SMRT_AT(BR, ldp, i, rowof[k]) * ipiv[k]; }
          }
          __syncthreads();
          SMRT_FOR_2D(i, k, h, h) { SMRT_AT(BR, ldp, i, k) = SMRT_AT(bounce, ldp, i, k); }
          __syncthreads();
          if (nr == 1 && NT >= h) {  // s = v + R_new b'
            block_matvec_dual(h, h, BR, (const double*)nullptr, ldp, Trhs, GJV,
                              [&](int i, double acc, double) { svec[i] = vvec[i] + acc; });
          } else if (nr > 0) {
            block_gemm_ptr(NT, h, nr, h, BR, ldp, [&](int c) { return Trhs + (size_t)c * ldp; },
                           [&](int i, int c, double acc) { SMRT_AT(svec, h, i, c) = SMRT_AT(vvec, h, i, c) + acc; });
          }
          for (int a = tid; a < h; a += NT) Ttprev[a] = Tt[a];
          h_prev = h;
          ldr_prev = ldp;
          have_prev = true;
          src_prev = (nr > 0);
          __syncthreads();
          SMRT_PHASE(7)  // R of the stack, source vector
        } else {
          // top layer: z = S^-1 b' by row elimination of [S | b'], then s = v + K z
          if (blocked ? block_gj_rows_blocked<!kGlobalScratch>(TS, ldp, Trhs, ldp, h, nr, rowof, pivinv, GJV, &s_ctrl[6], A.gj_single ? hmax : 0)
                      : block_gj_rows(TS, ldp, h, h + nr, rowstep, rowof)) {
            failed = true;
            break;
          }
          for (int k = tid; k < h; k += NT) ipiv[k] = blocked ? pivinv[k] : 1.0 / SMRT_AT(TS, ldp, rowof[k], k);
          __syncthreads();
          SMRT_FOR_2D(k, c, h, nr) { SMRT_AT(ytr, h, k, c) = SMRT_AT(Trhs, ldp, rowof[k], c) * ipiv[k]; }
          __syncthreads();
          block_gemm_ptr(NT, h, nr, h, BR, ldp, [&](int c) { return ytr + (size_t)c * h; },
                         [&](int i, int c, double acc) { SMRT_AT(svec, h, i, c) = SMRT_AT(vvec, h, i, c) + acc; });
          __syncthreads();
        }
        }  // resident / staged-operand layouts
      }  // layers
      if (pf_layer >= 0) {  // left the loop early with a bulk copy in flight: drain it before BF / BG are reused
        smrt_mbar_wait(&s_mbar, pf_parity);
        pf_parity ^= 1u;
        pf_layer = -1;
      }
      if (failed) break;

      // emerging intensity ---------------------------------------------------------------------------- dort.py:472-488
      // layer 0 quantities are still live: mu (layer 0), Tt (snow -> air), svec (h0 x nrhs)
      {
        const cplx eps_l0 = eps0;
        const double rindex0 = real_index_of(eps_star, eps_l0);
        const int n_0 = stream_count(rindex0, A.gl_mu, n);
        const int h0 = npol * n_0;
        const double T0 = A.temperature[bL];
        const double B0 = (A.mode == 0 && m == 0 && T0 > 0.0) ? planck_function(freq, T0, A.rayleigh_jeans) : 0.0;
        if (A.mode == 0) {
          // passive: I0up = T_top0 (I1up + B0) on the first npol*n_air rows; Tb(theta) by inverse Planck + interpolation
          // with an atmosphere: + R_air I_down (dort.py:481-484), then I_up(atm) + transmittance * I (rtsolver_utils.py:302-304)
          for (int a = tid; a < 2 * n_air; a += NT) {
            double v = (a < h0) ? Tt[a] * (svec[a] + B0) : 0.0;
            if (A.atmosphere) {
              if (atm_down != 0.0) {
                FresnelRT fa = (kRough ? air_interface_power(A, bL, freq, eps0, outmu[a >> 1], outw[a >> 1], coherent ? -1 : m,
                                                   mdiff_max)
                                         : fresnel_power(A.interface_kind[bL], c_make(1.0, 0.0), eps0, outmu[a >> 1]));
                v = fa.R[a & 1] * atm_down + v;
              }
              v = atm_up + atm_trans * v;
            }
            btop[a] = inverse_planck_function(freq, v, A.rayleigh_jeans);  // reuse btop: Tb per (stream, pol)
          }
          __syncthreads();
          for (int e = tid; e < 2 * A.n_theta; e += NT) {
            int p = e / A.n_theta, t = e % A.n_theta;
            double umu = cos(A.theta[t]);
            // rtsolver_utils.py:179-239: linear in mu with extrapolation; node mu = 1 inserted when needed
            bool ins = false;
            for (int tt = 0; tt < A.n_theta; ++tt) ins = ins || (cos(A.theta[tt]) > outmu[0]);
            int nn = n_air + (ins ? 1 : 0);
            auto xn = [&](int i) { return ins ? (i == 0 ? 1.0 : outmu[i - 1]) : outmu[i]; };
            auto yn = [&](int i) {
              if (ins) {
                if (i == 0) return 0.5 * (btop[0] + btop[1]);
                return btop[(i - 1) * 2 + p];
              }
              return btop[i * 2 + p];
            };
            // nodes are descending; find the segment [lo, lo+1] in ASCENDING order semantics of scipy interp1d
            // ascending index k = nn-1-i ; idx = searchsorted(x_asc, umu) clipped to [1, nn-1]
            int cnt = 0;  // number of nodes with x < umu
            for (int i = 0; i < nn; ++i) cnt += (xn(i) < umu) ? 1 : 0;
            int idx = cnt < 1 ? 1 : (cnt > nn - 1 ? nn - 1 : cnt);
            int ilo = nn - 1 - (idx - 1), ihi = nn - 1 - idx;  // descending-array indices of x_lo < x_hi
            double xlo = xn(ilo), xhi = xn(ihi), ylo = yn(ilo), yhi = yn(ihi);
            double slope = (yhi - ylo) / (xhi - xlo);
            out[p * A.n_theta + t] = slope * (umu - xlo) + ylo;
          }
          if (tid == 0) {
            A.n_streams_out[b] = n_air;
          }
          for (int j = tid; j < n; j += NT)
            A.stream_angles[(size_t)b * n + j] = (j < n_air) ? acos(outmu[j]) * (180.0 / SMRT_PI) : nan("");
        } else {
          // active: keep only the backscatter element (stream inc[j], beam j) of every (pol_out, pol_in) pair
          for (int e = tid; e < npol * npol * n_incs; e += NT) {
            int jinc = e / (npol * npol), rem = e % (npol * npol);
            int ps = rem / npol, pi = rem % npol;
            int i = inc[jinc];
            int row = i * npol + ps, col = jinc * npol + pi;
            double power = 1.0 / (2.0 * SMRT_PI * outw[i]);
            if (m > 0) power *= 2.0;
            FresnelRT fa = (kRough ? air_interface_power(A, bL, freq, eps0, outmu[i], outw[i], coherent ? -1 : m, mdiff_max)
                                         : fresnel_power(A.interface_kind[bL], c_make(1.0, 0.0), eps0, outmu[i]));
            double idn = (ps == pi) ? power : 0.0;
            double i1 = (row < h0) ? SMRT_AT(svec, h0, row, col) : 0.0;
            double v = fa.R[ps] * idn + ((row < h0) ? Tt[row] * i1 : 0.0);
            if (coherent) {
              coh_act[(ps * 2 + pi) * SMRT_MAX_INC + jinc] = v;
            } else {
              if (ps < 2 && pi < 2) v -= coh_act[(ps * 2 + pi) * SMRT_MAX_INC + jinc] * (m > 0 ? 2.0 : 1.0);
              double f;
              if (m == 0)
                f = 1.0;
              else
                f = (ps < 2) ? cos(m * A.phi) : sin(m * A.phi);
              acc_act[(ps * 3 + pi) * SMRT_MAX_INC + jinc] += f * v;
            }
          }
        }
        __syncthreads();
      }
    }  // runs

    if (failed) {
      for (int i = tid; i < n_out; i += NT) out[i] = nan("");
      if (tid == 0) {
        set_error(A.status, b, ST_SINGULAR);
        A.n_streams_out[b] = 0;
        A.optical_depth[b] = tau_report;
      }
      __syncthreads();
      continue;
    }
    if (A.mode == 1) {
      // interpolation over the incident streams ------------------------------------------- rtsolver_utils.py:179-239
      for (int e = tid; e < 9 * A.n_inc; e += NT) {
        int ps = e / (3 * A.n_inc), pi = (e / A.n_inc) % 3, t = e % A.n_inc;
        double umu = cos(A.theta_inc[t]);
        bool ins = false;
        for (int tt = 0; tt < A.n_inc; ++tt) ins = ins || (cos(A.theta_inc[tt]) > outmu[inc[0]]);
        int nn = n_incs + (ins ? 1 : 0);
        auto val = [&](int a, int c, int j) { return acc_act[(a * 3 + c) * SMRT_MAX_INC + j]; };
        auto xn = [&](int i) { return ins ? (i == 0 ? 1.0 : outmu[inc[i - 1]]) : outmu[inc[i]]; };
        auto yn = [&](int i) {
          if (ins) {
            if (i == 0) {  // inserted nadir node: co/cross-pol means of the steepest stream
              double copol = 0.5 * (val(0, 0, 0) + val(1, 1, 0));
              double cross = 0.5 * (val(1, 0, 0) + val(0, 1, 0));
              if (ps < 2 && pi < 2) return (ps == pi) ? copol : cross;
              return val(ps, pi, 0);
            }
            return val(ps, pi, i - 1);
          }
          return val(ps, pi, i);
        };
        double res;
        if (nn == 1) {
          res = yn(0);
        } else {
          int cnt = 0;
          for (int i = 0; i < nn; ++i) cnt += (xn(i) < umu) ? 1 : 0;
          int idx = cnt < 1 ? 1 : (cnt > nn - 1 ? nn - 1 : cnt);
          int ilo = nn - 1 - (idx - 1), ihi = nn - 1 - idx;
          double xlo = xn(ilo), xhi = xn(ihi), ylo = yn(ilo), yhi = yn(ihi);
          double slope = (yhi - ylo) / (xhi - xlo);
          res = slope * (umu - xlo) + ylo;
        }
        out[(ps * 3 + pi) * A.n_inc + t] = res;
      }
      if (tid == 0) A.n_streams_out[b] = n_incs;
      for (int j = tid; j < n; j += NT)
        A.stream_angles[(size_t)b * n + j] = (j < n_incs) ? acos(outmu[inc[j]]) * (180.0 / SMRT_PI) : nan("");
    }
    if (tid == 0) {
      A.optical_depth[b] = tau_report;
      if (shallow) atomicOr(&A.status[b], ST_WARN_SHALLOW);
    }
    __syncthreads();
  }
}
