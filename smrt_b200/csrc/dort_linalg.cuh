// dort_linalg.cuh — CTA-cooperative dense fp64 linear algebra on small column-major matrices held in shared memory
// (or, for stream counts whose blocks exceed 227 KB, in an L2-resident global scratch): register-tiled GEMMs, block
// matrix-vector products, left-looking Cholesky, one-sided Jacobi SVD (register-blocked for h <= 64, plain for larger
// blocks), Gauss-Jordan eliminations with partial pivoting (panel-blocked for h <= 64) and the triangular solve the
// DORT path needs.  All routines are called by every thread of the block (or of a half-block "team" for the two
// concurrent Cholesky factorisations) and synchronise internally.
#pragma once
#include "simt.h"
#include <math.h>

// A "team" is a contiguous range of threads of the block that synchronises on its own named barrier.
struct Team {
  int rank;      // thread index inside the team
  int size;      // number of threads (multiple of 32)
  int bar_id;    // named barrier id (0 = whole block -> __syncthreads)
  SMRT_DEV void sync() const {
    if (bar_id == 0)
      __syncthreads();
    else
      smrt_named_barrier(bar_id, size);
  }
};

SMRT_DEV Team block_team() {
  Team t;
  t.rank = threadIdx.x;
  t.size = blockDim.x;
  t.bar_id = 0;
  return t;
}

#define SMRT_AT(A, ld, i, j) (A)[(size_t)(j) * (ld) + (i)]

// ---------------------------------------------------------------------------------------------------------------- GEMM
// C(i, j) = epilogue(i, j, sum_k a(i, k) * b(k, j)),  i < M, j < N, k in [k0(i,j), K).
// Each thread owns a TM x TN register tile with rows strided by 16 and columns strided by 16 inside a 64 x 64 macro
// tile, so that consecutive lanes touch consecutive rows (conflict-free column-major shared-memory reads of `a`).
// a(i,k), b(k,j) are functors returning double (they may build operands on the fly); out-of-range indices are never
// requested.  `store(i, j, acc)` writes the result.  No synchronisation inside: the caller syncs before/after.
template <typename FA, typename FB, typename FS>
SMRT_DEV void team_gemm(const Team& tm, int M, int N, int K, FA a, FB b, FS store) {
  const int TX = 16;                 // threads along rows
  const int TY = tm.size / TX;       // threads along columns
  const int tx = tm.rank % TX, ty = tm.rank / TX;
  const int RM = 4, RN = 4;
  for (int j0 = 0; j0 < N; j0 += TY * RN) {
    for (int i0 = 0; i0 < M; i0 += TX * RM) {
      double acc[RM][RN];
#pragma unroll
      for (int ii = 0; ii < RM; ++ii)
#pragma unroll
        for (int jj = 0; jj < RN; ++jj) acc[ii][jj] = 0.0;
      int irow[RM], jcol[RN];
      bool iok[RM], jok[RN];
#pragma unroll
      for (int ii = 0; ii < RM; ++ii) {
        irow[ii] = i0 + tx + ii * TX;
        iok[ii] = irow[ii] < M;
      }
#pragma unroll
      for (int jj = 0; jj < RN; ++jj) {
        jcol[jj] = j0 + ty + jj * TY;
        jok[jj] = jcol[jj] < N;
      }
      for (int k = 0; k < K; ++k) {
        double av[RM], bv[RN];
#pragma unroll
        for (int ii = 0; ii < RM; ++ii) av[ii] = iok[ii] ? a(irow[ii], k) : 0.0;
#pragma unroll
        for (int jj = 0; jj < RN; ++jj) bv[jj] = jok[jj] ? b(k, jcol[jj]) : 0.0;
#pragma unroll
        for (int ii = 0; ii < RM; ++ii)
#pragma unroll
          for (int jj = 0; jj < RN; ++jj) acc[ii][jj] = fma(av[ii], bv[jj], acc[ii][jj]);
      }
#pragma unroll
      for (int ii = 0; ii < RM; ++ii)
#pragma unroll
        for (int jj = 0; jj < RN; ++jj)
          if (iok[ii] && jok[jj]) store(irow[ii], jcol[jj], acc[ii][jj]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------ Cholesky
// Lower-triangular operand (the factor C of X+ and, before it, the lower triangle of X+ itself), either inside a full
// column-major array (leading dimension ld) or PACKED by columns (column k holds rows k .. h-1: h (h + 1) / 2 doubles,
// half the shared memory; the 64 < h <= 128 instantiation of the eigen kernel could not hold two full matrices).
// col(k)[i] = element (i, k) for i >= k in both layouts.
template <bool kPacked>
struct LowerMat {
  double* p;
  int ld, h;
  SMRT_DEV double* col(int k) const {
    return kPacked ? p + ((k * (2 * h - k - 1)) >> 1) : p + (size_t)k * ld;
  }
  SMRT_DEV int col_step(int k) const { return kPacked ? h - k - 1 : ld; }  // col(k + 1) - col(k)
  SMRT_DEV double& at(int i, int k) const { return col(k)[i]; }
};

// In-place lower Cholesky A = L L^T of the h x h symmetric positive definite matrix whose LOWER triangle is stored in A;
// the strict upper triangle is neither read nor written.
// Left-looking with ONE team barrier per step (a right-looking version needs three per column and rewrites the
// trailing matrix at every step).  Thread t owns the rows t, t + size, ... (NR of them); at step j it forms
//   s_i = A(i, j) - sum_{k<j} L(i, k) L(j, k)      for its rows i > j
// and, redundantly (the L(j, k) operands are loaded anyway), the pivot s_jj = A(j, j) - sum_k L(j, k)^2, so that
// L(i, j) = s_i / sqrt(s_jj) can be written without a second synchronisation.  The diagonal of L is never read during
// the factorisation; it is collected in dvec (team-shared double[h]) and copied into A at the end, which removes the
// write-after-read race on A(j, j).  Only the lower triangle is read or written.  Returns 1 (to every thread of the
// team) when a pivot is not positive.
template <int NR, bool kPacked>
SMRT_DEV int team_cholesky_ll(const Team& tm, const LowerMat<kPacked>& A, int h, double* dvec) {
  int failed = 0;
  // two columns (j, j + 1) per step: the three operands L(i, k), L(j, k), L(j + 1, k) feed five FMAs, and the team
  // synchronises once per pair of columns
  int j = 0;
  for (; j + 1 < h; j += 2) {
    double p00[2] = {0.0, 0.0}, p10[2] = {0.0, 0.0}, p11[2] = {0.0, 0.0};
    double s0[NR][2], s1[NR][2];
    int row[NR];
    bool own[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int i = tm.rank + r * tm.size;
      own[r] = (i > j + 1) && (i < h);
      row[r] = own[r] ? i : j;
      s0[r][0] = s0[r][1] = s1[r][0] = s1[r][1] = 0.0;
    }
    const double* SMRT_RESTRICT ck = A.col(0);  // column k: ck[i] = L(i, k)
    int k = 0;
    for (; k + 1 < j; k += 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const double l0 = ck[j], l1 = ck[j + 1];
        p00[u] = fma(l0, l0, p00[u]);
        p10[u] = fma(l1, l0, p10[u]);
        p11[u] = fma(l1, l1, p11[u]);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const double li = ck[row[r]];
          s0[r][u] = fma(li, l0, s0[r][u]);
          s1[r][u] = fma(li, l1, s1[r][u]);
        }
        ck += A.col_step(k + u);
      }
    }
    if (k < j) {
      const double l0 = ck[j], l1 = ck[j + 1];
      p00[0] = fma(l0, l0, p00[0]);
      p10[0] = fma(l1, l0, p10[0]);
      p11[0] = fma(l1, l1, p11[0]);
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const double li = ck[row[r]];
        s0[r][0] = fma(li, l0, s0[r][0]);
        s1[r][0] = fma(li, l1, s1[r][0]);
      }
    }
    double* cj = A.col(j);
    double* cj1 = A.col(j + 1);
    // 2 x 2 diagonal block (identical values in every thread: uniform exits)
    const double d0 = cj[j] - (p00[0] + p00[1]);
    if (!(d0 > 0.0)) {
      failed = 1;
      break;
    }
    const double r0 = rsqrt(d0);
    const double l10 = (cj[j + 1] - (p10[0] + p10[1])) * r0;
    const double d1 = cj1[j + 1] - (p11[0] + p11[1]) - l10 * l10;
    if (!(d1 > 0.0)) {
      failed = 1;
      break;
    }
    const double r1 = rsqrt(d1);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int i = tm.rank + r * tm.size;
      if (own[r]) {
        const double x0 = (cj[i] - (s0[r][0] + s0[r][1])) * r0;
        const double x1 = (cj1[i] - (s1[r][0] + s1[r][1]) - x0 * l10) * r1;
        cj[i] = x0;
        cj1[i] = x1;
      }
      if (i == j) {
        dvec[j] = d0 * r0;
        dvec[j + 1] = d1 * r1;
      }
    }
    tm.sync();
    // L(j + 1, j) is read as an operand by the later steps but A(j + 1, j) was also an INPUT of this step for every
    // thread: written only after the barrier, by its owner, and published by the next step's barrier (the next step
    // reads column j only through rows >= j + 2)
    if (tm.rank == (j + 1) % tm.size) cj[j + 1] = l10;
  }
  if (!failed && j < h) {  // last column of an odd-sized matrix
    double p[2] = {0.0, 0.0};
    const double* SMRT_RESTRICT ck = A.col(0);
    for (int k = 0; k < j; ++k) {
      const double lj = ck[j];
      p[k & 1] = fma(lj, lj, p[k & 1]);
      ck += A.col_step(k);
    }
    const double d = A.at(j, j) - (p[0] + p[1]);
    if (!(d > 0.0))
      failed = 1;
    else if (tm.rank == 0)
      dvec[j] = sqrt(d);
  }
  tm.sync();
  if (failed) return 1;
  for (int jj = tm.rank; jj < h; jj += tm.size) A.at(jj, jj) = dvec[jj];
  tm.sync();
  return 0;
}
template <bool kPacked>
SMRT_DEV int team_cholesky_fast(const Team& tm, const LowerMat<kPacked>& A, int h, double* dvec) {
  if (h <= tm.size) return team_cholesky_ll<1, kPacked>(tm, A, h, dvec);
  if (h <= 2 * tm.size) return team_cholesky_ll<2, kPacked>(tm, A, h, dvec);
  if (h <= 4 * tm.size) return team_cholesky_ll<4, kPacked>(tm, A, h, dvec);
  return team_cholesky_ll<8, kPacked>(tm, A, h, dvec);
}

// ------------------------------------------------------------------------------------------------- one-sided Jacobi SVD
// Orthogonalises the columns of the h x h matrix W (column-major, ld) in place by plane rotations applied from the
// right (Hestenes): on exit W = M V with V orthogonal and mutually orthogonal columns, |w_j| = sigma_j.
// Round-robin ("circle") ordering: hp = h rounded up to even, hp - 1 rounds per sweep, hp / 2 disjoint column pairs per
// round, TPP threads per pair (power of two <= 32).  `ctrl` is a block-shared int[4] scratch.
// Returns the number of sweeps performed (every thread gets the same value).
#define SMRT_JACOBI_TOL2 1e-30       // rotate only when (w_p . w_q)^2 > tol^2 |w_p|^2 |w_q|^2   (tol = 1e-15)
#ifndef SMRT_JACOBI_QUAD2
#define SMRT_JACOBI_QUAD2 1e-18      // a sweep whose largest squared cosine (before its own rotations) stays below this
                                     // is the last one: the quadratic convergence of the cyclic method finishes the job
#endif
#define SMRT_JACOBI_MAX_SWEEPS 40

// one column pair: R rows per lane held in registers between the dot products and the rotation
template <int R>
SMRT_DEV int jacobi_pair(double* SMRT_RESTRICT wp, double* SMRT_RESTRICT wq, int h, int lane, int tpp) {
  // called by EVERY lane of the warp (groups without a pair pass h = 0): the shuffles use the full, compile-time mask
  double x[R], y[R];
  double a0 = 0.0, b0 = 0.0, g0 = 0.0, a1 = 0.0, b1 = 0.0, g1 = 0.0;
#pragma unroll
  for (int u = 0; u < R; ++u) {
    const int i = lane + u * tpp;
    const bool ok = i < h;
    x[u] = ok ? wp[i] : 0.0;
    y[u] = ok ? wq[i] : 0.0;
  }
#pragma unroll
  for (int u = 0; u < R; u += 2) {
    a0 = fma(x[u], x[u], a0);
    b0 = fma(y[u], y[u], b0);
    g0 = fma(x[u], y[u], g0);
    if (u + 1 < R) {
      a1 = fma(x[u + 1], x[u + 1], a1);
      b1 = fma(y[u + 1], y[u + 1], b1);
      g1 = fma(x[u + 1], y[u + 1], g1);
    }
  }
  double a = a0 + a1, b = b0 + b1, g = g0 + g1;
  for (int off = tpp >> 1; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off, 32);
    b += __shfl_xor_sync(0xffffffffu, b, off, 32);
    g += __shfl_xor_sync(0xffffffffu, g, off, 32);
  }
  const double g2 = g * g, ab = a * b;
  if (!(g2 > SMRT_JACOBI_TOL2 * ab)) return 0;
  // tan of the rotation angle: t = sgn(d) 2 g / (|d| + sqrt(d^2 + 4 g^2)), d = |w_q|^2 - |w_p|^2
  const double d = b - a;
  const double t = copysign(2.0 * g, (d >= 0.0) ? g : -g) / (fabs(d) + sqrt(fma(d, d, 4.0 * g2)));
  const double c = rsqrt(fma(t, t, 1.0));
  const double s = c * t;
#pragma unroll
  for (int u = 0; u < R; ++u) {
    const int i = lane + u * tpp;
    if (i < h) {
      wp[i] = fma(c, x[u], -s * y[u]);
      wq[i] = fma(s, x[u], c * y[u]);
    }
  }
  return (g2 > SMRT_JACOBI_QUAD2 * ab) ? 1 : 0;
}

SMRT_DEV_NOINLINE int block_jacobi_svd(double* W, int ld, int h, int* ctrl) {
  const int NT = blockDim.x;
  const int tid = threadIdx.x;
  const int hp = (h + 1) & ~1;
  const int hm1 = hp - 1;
  const int npairs = hp / 2;
  int tpp = 32;
  while (tpp > 1 && npairs * tpp > NT) tpp >>= 1;
  // if there are more pairs than groups (h > 2 NT) each group loops over several pairs
  const int ngroups = NT / tpp;
  const int grp = tid / tpp, lane = tid % tpp;
  const int rows = (h + tpp - 1) / tpp;  // rows per lane
  (void)ctrl;
  int sweeps = 0;
  for (;;) {
    int notconv = 0;
    for (int r = 0; r < hm1; ++r) {
      // uniform trip count: every thread of the block walks the same number of pair slots, threads without a pair
      // (padding column of an odd-sized problem, or more groups than pairs) run the pair code on an empty column
      for (int pi0 = 0; pi0 < npairs; pi0 += ngroups) {
        const int pi = pi0 + grp;
        int p, q;
        if (pi == 0) {
          p = r;
          q = hm1;
        } else {
          p = r + pi;
          if (p >= hm1) p -= hm1;
          q = r - pi;
          if (q < 0) q += hm1;
        }
        if (p > q) {
          int t = p;
          p = q;
          q = t;
        }
        const bool valid = (pi < npairs) && (q < h);
        const int hh = valid ? h : 0;
        double* wp = W + (size_t)(valid ? p : 0) * ld;
        double* wq = W + (size_t)(valid ? q : 0) * ld;
        if (rows <= 2)
          notconv |= jacobi_pair<2>(wp, wq, hh, lane, tpp);
        else if (rows <= 4)
          notconv |= jacobi_pair<4>(wp, wq, hh, lane, tpp);
        else if (rows <= 8)
          notconv |= jacobi_pair<8>(wp, wq, hh, lane, tpp);
        else {
          // large problems (global scratch path): plain loops
          double a = 0.0, b = 0.0, g = 0.0;
          for (int i = lane; i < hh; i += tpp) {
            double x = wp[i], y = wq[i];
            a = fma(x, x, a);
            b = fma(y, y, b);
            g = fma(x, y, g);
          }
          for (int off = tpp >> 1; off > 0; off >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, off, 32);
            b += __shfl_xor_sync(0xffffffffu, b, off, 32);
            g += __shfl_xor_sync(0xffffffffu, g, off, 32);
          }
          const double g2 = g * g, ab = a * b;
          if (g2 > SMRT_JACOBI_TOL2 * ab) {
            if (g2 > SMRT_JACOBI_QUAD2 * ab) notconv = 1;
            const double d = b - a;
            const double t = copysign(2.0 * g, (d >= 0.0) ? g : -g) / (fabs(d) + sqrt(fma(d, d, 4.0 * g2)));
            const double c = rsqrt(fma(t, t, 1.0));
            const double s = c * t;
            for (int i = lane; i < hh; i += tpp) {
              double x = wp[i], y = wq[i];
              wp[i] = fma(c, x, -s * y);
              wq[i] = fma(s, x, c * y);
            }
          }
        }
      }
      __syncthreads();
    }
    ++sweeps;
    if (!__syncthreads_or(notconv) || sweeps >= SMRT_JACOBI_MAX_SWEEPS) break;
  }
  return sweeps;
}

// ------------------------------------------------------------------------- register-blocked one-sided Jacobi (h <= 64)
// Same method as block_jacobi_svd, restructured around the two limits the profile of the first version showed
// (shared-memory wavefronts 65 % of peak with 1/3 of them bank conflicts, FP64 pipe 26 %):
//   * columns are handled in BLOCKS of two.  A group of 8 lanes loads two blocks (4 columns, R rows per lane, 16-byte
//     accesses: 8 lanes x 16 B = one conflict-free 128-byte wavefront) and performs all four cross rotations in
//     registers before storing them back: half the shared-memory traffic per rotation.  Blocks meet in round-robin
//     order (nb - 1 block rounds per sweep); the two columns of a block are rotated against each other at the start
//     of the sweep, together with the exact recomputation of every column norm.
//   * within a sweep the squared column norms are TRACKED (a' = a - t g, b' = b + t g) instead of recomputed for every
//     pair: one dot product per rotation instead of three.
//   * tan of the rotation angle from MUFU seeds (rsqrt / rcp, one Newton step each) instead of an IEEE sqrt and a
//     division; cos = rsqrt(1 + t^2) stays a full-precision rsqrt, so every rotation is orthogonal to rounding whatever
//     the accuracy of t.
//   * no predicates in the inner loops: the rows [h, 8 R) of every column are zero (and stay zero under rotations),
//     missing columns (odd h, padding blocks, idle groups) are read from a column of zeros and can never rotate, and
//     the rotation angles of the two pairs handled together are computed by the two halves of the group.
// W: column-major, leading dimension ld = jacobi_ld(h) (16-byte aligned columns, >= 8 R rows), rows [h, 8 R) zero on
// entry.  nrm: block-shared double[>= 2 * nb + 1]; zcol: >= 8 R zeros, 16-byte aligned.  Every thread of the block
// calls; the block size is a multiple of 32.  Returns the number of sweeps (same value in every thread).
// lanes per group: JG = 8 with 128-thread blocks, 16 with 256-thread blocks (more warps in flight for the same
// shared-memory footprint: the rotation chain is latency bound); R = rows per lane, JG * R = padded row count.
// leading dimension of the Jacobi operand: rows padded to a multiple of 16, = 2 (mod 4) so that column-strided accesses
// (two lanes per column in the back substitution) stay conflict-free
SMRT_HD int jacobi_ld(int h) {
  const int hr = (h + 1) & ~1;
  if (hr <= 16) return 18;
  if (hr <= 32) return 34;
  if (hr <= 48) return 50;
  if (hr <= 64) return 66;
  if (hr <= 96) return 98;    // 16 lanes x 6 rows
  if (hr <= 128) return 130;  // 16 lanes x 8 rows
  return ((hr & 3) == 2) ? hr : hr + 2;
}

// row owned by lane `lane` in register slot u: 16-byte pairs when R is even, single doubles otherwise
template <int JG, int R>
SMRT_DEV void jreg_load(const double* SMRT_RESTRICT col, int lane, double (&x)[R]) {
  if (R % 2 == 0) {
#pragma unroll
    for (int v = 0; v < R / 2; ++v) {
      const double2 t = *reinterpret_cast<const double2*>(col + 2 * lane + 2 * JG * v);
      x[2 * v] = t.x;
      x[2 * v + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int u = 0; u < R; ++u) x[u] = col[lane + JG * u];
  }
}
template <int JG, int R>
SMRT_DEV void jreg_store(double* SMRT_RESTRICT col, int lane, const double (&x)[R]) {
  if (R % 2 == 0) {
#pragma unroll
    for (int v = 0; v < R / 2; ++v) {
      double2 t;
      t.x = x[2 * v];
      t.y = x[2 * v + 1];
      *reinterpret_cast<double2*>(col + 2 * lane + 2 * JG * v) = t;
    }
  } else {
#pragma unroll
    for (int u = 0; u < R; ++u) col[lane + JG * u] = x[u];
  }
}
// sum over the JG lanes of a group (every lane of the warp takes part; identical result in all lanes of the group)
template <int JG>
SMRT_DEV double jreg_group_sum(double v) {
#pragma unroll
  for (int off = 1; off < JG; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off, 32);
  return v;
}
template <int JG, int R>
SMRT_DEV double jreg_dot(const double (&x)[R], const double (&y)[R]) {
  double g0 = 0.0, g1 = 0.0;
#pragma unroll
  for (int u = 0; u < R; u += 2) {
    g0 = fma(x[u], y[u], g0);
    if (u + 1 < R) g1 = fma(x[u + 1], y[u + 1], g1);
  }
  return jreg_group_sum<JG>(g0 + g1);
}
// tangent, cosine and sine of the rotation that orthogonalises a column pair with squared norms (a, b) and inner
// product g; rot = false gives the identity.  Branch-free.
SMRT_DEV void jreg_angle(double a, double b, double g, bool rot, double& t, double& c, double& s) {
  // t = sgn(d) 2 g / (|d| + sqrt(d^2 + 4 g^2)), d = |w_q|^2 - |w_p|^2
  const double d = b - a;
  const double q = rot ? fma(d, d, 4.0 * (g * g)) : 1.0;
  double r = smrt_rsqrt_approx(q);
  r = r * fma(-0.5 * q * r, r, 1.5);
  const double dd = fabs(d) + q * r;
  double rd = smrt_rcp_approx(dd);
  rd = fma(rd, fma(-dd, rd, 1.0), rd);
  t = rot ? copysign(2.0 * g, (d >= 0.0) ? g : -g) * rd : 0.0;
  c = rsqrt(fma(t, t, 1.0));
  s = c * t;
}
template <int R>
SMRT_DEV void jreg_apply(double (&x)[R], double (&y)[R], double c, double s) {
#pragma unroll
  for (int u = 0; u < R; ++u) {
    const double xu = x[u], yu = y[u];
    x[u] = fma(c, xu, -s * yu);
    y[u] = fma(s, xu, c * yu);
  }
}
// one rotation of the column pair (x, y) with squared norms (a, b), updated in place.
// returns bit 0: the pair was not yet orthogonal to quadratic-convergence level; bit 1: the columns were modified
template <int JG, int R>
SMRT_DEV int jreg_rotate(double (&x)[R], double (&y)[R], double& a, double& b) {
  const double g = jreg_dot<JG, R>(x, y);
  const double g2 = g * g, ab = a * b;
  if (!(g2 > SMRT_JACOBI_TOL2 * ab)) return 0;
  double t, c, s;
  jreg_angle(a, b, g, true, t, c, s);
  jreg_apply<R>(x, y, c, s);
  const double tg = t * g;
  a -= tg;
  b += tg;
  return (g2 > SMRT_JACOBI_QUAD2 * ab) ? 3 : 2;
}
// two INDEPENDENT rotations (x0, y0) and (x1, y1) handled together: the two dot products and their reductions
// interleave, the lower half of the group computes the angle of the first pair and the upper half that of the second
// one (the results are exchanged with three shuffles), then both rotations are applied.  Every lane of the WARP must
// call (the branch around the angle computation is warp-uniform).  Return codes as jreg_rotate, in rc0 / rc1.
template <int JG, int R>
SMRT_DEV void jreg_rotate2(int lane, double (&x0)[R], double (&y0)[R], double& a0, double& b0, double (&x1)[R],
                           double (&y1)[R], double& a1, double& b1, int& rc0, int& rc1) {
  double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;
#pragma unroll
  for (int u = 0; u < R; u += 2) {
    p0 = fma(x0[u], y0[u], p0);
    q0 = fma(x1[u], y1[u], q0);
    if (u + 1 < R) {
      p1 = fma(x0[u + 1], y0[u + 1], p1);
      q1 = fma(x1[u + 1], y1[u + 1], q1);
    }
  }
  // the lower half of the group collects the inner product of the first pair, the upper half that of the second one:
  // one exchange across the halves, then log2(JG / 2) butterfly stages on ONE value (half the shuffles of reducing
  // both sums in every lane; each half computes the angle of its own pair anyway)
  const bool hi = (lane & (JG / 2)) != 0;
  const double g0p = p0 + p1, g1p = q0 + q1;
  double gm = (hi ? g1p : g0p) + __shfl_xor_sync(0xffffffffu, hi ? g0p : g1p, JG / 2, 32);
#pragma unroll
  for (int off = 1; off < JG / 2; off <<= 1) gm += __shfl_xor_sync(0xffffffffu, gm, off, 32);
  const double am = hi ? a1 : a0, bm = hi ? b1 : b0;
  const double g2m = gm * gm, abm = am * bm;
  const bool rotm = g2m > SMRT_JACOBI_TOL2 * abm;
  const bool quadm = g2m > SMRT_JACOBI_QUAD2 * abm;
  // the flags of both pairs for every lane of the group: two warp votes, bit = first lane of the half
  const unsigned vrot = __ballot_sync(0xffffffffu, rotm), vquad = __ballot_sync(0xffffffffu, quadm);
  rc0 = rc1 = 0;
  if (vrot == 0u) return;  // warp-uniform
  const int g0lane = ((int)threadIdx.x & 31) & ~(JG - 1);
  const bool rot0 = (vrot >> g0lane) & 1u, rot1 = (vrot >> (g0lane + JG / 2)) & 1u;
  const bool quad0 = (vquad >> g0lane) & 1u, quad1 = (vquad >> (g0lane + JG / 2)) & 1u;
  double t, c, s;
  jreg_angle(am, bm, gm, rotm, t, c, s);
  const double tg = t * gm;
  const double co = __shfl_xor_sync(0xffffffffu, c, JG / 2, 32);
  const double so = __shfl_xor_sync(0xffffffffu, s, JG / 2, 32);
  const double tgo = __shfl_xor_sync(0xffffffffu, tg, JG / 2, 32);
  // both rotations are applied unconditionally: a pair that does not rotate got the exact identity (c = 1, s = 0,
  // t g = 0), and straight-line code keeps the 4 R operands in place (the conditional version paid one register move
  // per operand to merge the two paths: 13 % of the instructions of the sweep)
  {
    jreg_apply<R>(x0, y0, hi ? co : c, hi ? so : s);
    const double tg0 = hi ? tgo : tg;
    a0 -= tg0;
    b0 += tg0;
    rc0 = rot0 ? (quad0 ? 3 : 2) : 0;
  }
  {
    jreg_apply<R>(x1, y1, hi ? c : co, hi ? s : so);
    const double tg1 = hi ? tg : tgo;
    a1 -= tg1;
    b1 += tg1;
    rc1 = rot1 ? (quad1 ? 3 : 2) : 0;
  }
}

template <int JG, int R>
SMRT_DEV int block_jacobi_svd_reg(double* W, int ld, int h, double* nrm, const double* zcol) {
  const int NT = blockDim.x, tid = threadIdx.x;
  const int ngroups = NT / JG, grp = tid / JG, lane = tid % JG;
  const int ncb = (h + 1) >> 1;   // blocks of two columns (the last one holds a single column when h is odd)
  const int nb = (ncb + 1) & ~1;  // padded to an even number of blocks; blocks >= ncb are empty
  const int nb1 = nb - 1, npairs = nb >> 1;
  const int nodummy = 2 * nb;     // norm slot of the missing columns (always zero)
  // logical column c is stored column h - 1 - c: the operands M = C^T L of this solver come with column norms that
  // grow with the stream index, and the cyclic method converges faster when it meets the large columns first
  // (de Rijk's ordering; 5.4 -> 5.0 sweeps on the cfg-2 layers, profiles/r01_microbench.txt)
  double* const Wlast = W + (size_t)(h - 1) * ld;
  if (tid == 0) nrm[nodummy] = 0.0;
  int sweeps = 0;
  for (;;) {
    int notconv = 0;
    // the two columns of every block against each other, with exact norms (they are tracked from here on)
    for (int b0 = 0; b0 < nb; b0 += ngroups) {
      const int blk = b0 + grp;
      const int c0 = 2 * blk, c1 = c0 + 1;
      const bool v0 = (blk < nb) && (c0 < h), v1 = (blk < nb) && (c1 < h);
      double* w0 = v0 ? Wlast - (size_t)c0 * ld : const_cast<double*>(zcol);
      double* w1 = v1 ? Wlast - (size_t)c1 * ld : const_cast<double*>(zcol);
      double x[R], y[R];
      jreg_load<JG, R>(w0, lane, x);
      jreg_load<JG, R>(w1, lane, y);
      double a = jreg_dot<JG, R>(x, x), b = jreg_dot<JG, R>(y, y);
      const int rc = jreg_rotate<JG, R>(x, y, a, b);
      if (rc & 2) {  // only real columns can rotate
        jreg_store<JG, R>(w0, lane, x);
        jreg_store<JG, R>(w1, lane, y);
      }
      if (lane == 0) {
        nrm[v0 ? c0 : nodummy] = a;
        nrm[v1 ? c1 : nodummy] = b;
      }
      notconv |= rc & 1;
    }
    __syncthreads();
    // block rounds: round-robin tournament of the nb blocks, four cross rotations per meeting
    for (int r = 0; r < nb1; ++r) {
      for (int p0 = 0; p0 < npairs; p0 += ngroups) {
        const int pg = p0 + grp;
        int P, Q;
        if (pg == 0) {
          P = r;
          Q = nb1;
        } else {
          P = r + pg;
          if (P >= nb1) P -= nb1;
          Q = r - pg;
          if (Q < 0) Q += nb1;
        }
        if (P > Q) {
          const int t = P;
          P = Q;
          Q = t;
        }
        const bool act = pg < npairs;
        const int cp0 = 2 * P, cp1 = cp0 + 1, cq0 = 2 * Q, cq1 = cq0 + 1;
        const bool vp0 = act && cp0 < h, vp1 = act && cp1 < h, vq0 = act && cq0 < h, vq1 = act && cq1 < h;
        double* wp0 = vp0 ? Wlast - (size_t)cp0 * ld : const_cast<double*>(zcol);
        double* wp1 = vp1 ? Wlast - (size_t)cp1 * ld : const_cast<double*>(zcol);
        double* wq0 = vq0 ? Wlast - (size_t)cq0 * ld : const_cast<double*>(zcol);
        double* wq1 = vq1 ? Wlast - (size_t)cq1 * ld : const_cast<double*>(zcol);
        const int np0 = vp0 ? cp0 : nodummy, np1 = vp1 ? cp1 : nodummy;
        const int nq0 = vq0 ? cq0 : nodummy, nq1 = vq1 ? cq1 : nodummy;
        double x0[R], x1[R], y0[R], y1[R];
        jreg_load<JG, R>(wp0, lane, x0);
        jreg_load<JG, R>(wp1, lane, x1);
        jreg_load<JG, R>(wq0, lane, y0);
        jreg_load<JG, R>(wq1, lane, y1);
        double a0 = nrm[np0], a1 = nrm[np1], b0 = nrm[nq0], b1 = nrm[nq1];
        int r00, r11, r01, r10;
        jreg_rotate2<JG, R>(lane, x0, y0, a0, b0, x1, y1, a1, b1, r00, r11);
        jreg_rotate2<JG, R>(lane, x0, y1, a0, b1, x1, y0, a1, b0, r01, r10);
#ifndef SMRT_SIMT_EMULATION  // (the emulated shuffles above are warp barriers already; a pthread barrier per round is slow)
        __syncwarp();  // every lane of the group has read the tracked norms before lane 0 replaces them
#endif
        // a column that rotated is a real column (missing ones have zero norm): unpredicated stores
        if ((r00 | r01) & 2) {
          jreg_store<JG, R>(wp0, lane, x0);
          if (lane == 0) nrm[np0] = a0;
        }
        if ((r11 | r10) & 2) {
          jreg_store<JG, R>(wp1, lane, x1);
          if (lane == 0) nrm[np1] = a1;
        }
        if ((r00 | r10) & 2) {
          jreg_store<JG, R>(wq0, lane, y0);
          if (lane == 0) nrm[nq0] = b0;
        }
        if ((r11 | r01) & 2) {
          jreg_store<JG, R>(wq1, lane, y1);
          if (lane == 0) nrm[nq1] = b1;
        }
        notconv |= (r00 | r11 | r01 | r10) & 1;
      }
      __syncthreads();
    }
    ++sweeps;
    if (!__syncthreads_or(notconv) || sweeps >= SMRT_JACOBI_MAX_SWEEPS) break;
  }
  return sweeps;
}

template <int JG, int R>
SMRT_DEV int block_jacobi_svd_reg_sb(double* W, int ld, int h, double* nrm, const double* zcol) {
  // Ordering of a sweep: the G = 32 / JG lane groups of a warp own a SUPER-BLOCK pair (2 x G column blocks).  After the
  // exact-norm pass (the two columns of every block against each other) come
  //   * the meetings inside every super-block (round-robin of its G blocks, G - 1 rounds, two super-blocks per warp),
  //   * a round-robin tournament of the super-blocks: in a super-round a warp holds the super-blocks (SA, SB) and its
  //     group g meets block g of SA with the blocks g, g + 1, ... (mod G) of SB in G rounds.
  // Rounds of one warp are separated by __syncwarp only; the block-wide barrier comes once per super-round
  // (nb / G instead of nb - 1 per sweep).  Used with 16-lane groups (G = 2, 512 threads: 16 warps meet at every
  // barrier); with 8-lane groups the padding to whole super-block pairs costs more rounds than the barriers save, and
  // keeping the SA block in registers through a super-round spills (both measured: profiles/r04_notes.txt).
  constexpr int G = 32 / JG;
  const int NT = blockDim.x, tid = threadIdx.x;
  const int nwarps = NT >> 5, warp = tid >> 5;
  const int grp = (tid & 31) / JG, lane = tid % JG;
  const int ncb = (h + 1) >> 1;                      // blocks of two columns (the last one holds one column when h is odd)
  const int nsb = (((ncb + G - 1) / G) + 1) & ~1;    // super-blocks, padded to an even number (blocks >= ncb are empty)
  const int nsb1 = nsb - 1, nsp = nsb >> 1;
  const int nb = nsb * G;
  const int ngroups = NT / JG;
  const int nodummy = 2 * ncb;    // norm slot of the missing columns (always zero)
  // logical column c is stored column h - 1 - c: the operands M = C^T L of this solver come with column norms that
  // grow with the stream index, and the cyclic method converges faster when it meets the large columns first
  // (de Rijk's ordering; 5.4 -> 5.0 sweeps on the cfg-2 layers, profiles/r01_microbench.txt)
  double* const Wlast = W + (size_t)(h - 1) * ld;
  if (tid == 0) nrm[nodummy] = 0.0;
  int sweeps = 0;
  for (;;) {
    int notconv = 0;
    // the two columns of every block against each other, with exact norms (they are tracked from here on)
    for (int b0 = 0; b0 < nb; b0 += ngroups) {
      const int blk = b0 + tid / JG;
      const int c0 = 2 * blk, c1 = c0 + 1;
      const bool v0 = (blk < nb) && (c0 < h), v1 = (blk < nb) && (c1 < h);
      double* w0 = v0 ? Wlast - (size_t)c0 * ld : const_cast<double*>(zcol);
      double* w1 = v1 ? Wlast - (size_t)c1 * ld : const_cast<double*>(zcol);
      double x[R], y[R];
      jreg_load<JG, R>(w0, lane, x);
      jreg_load<JG, R>(w1, lane, y);
      double a = jreg_dot<JG, R>(x, x), b = jreg_dot<JG, R>(y, y);
      const int rc = jreg_rotate<JG, R>(x, y, a, b);
      if (rc & 2) {  // only real columns can rotate
        jreg_store<JG, R>(w0, lane, x);
        jreg_store<JG, R>(w1, lane, y);
      }
      if (lane == 0) {
        nrm[v0 ? c0 : nodummy] = a;
        nrm[v1 ? c1 : nodummy] = b;
      }
      notconv |= rc & 1;
    }
    __syncthreads();
    // sr = -1: meetings inside the super-blocks; sr >= 0: super-round sr of the tournament of the super-blocks
    for (int sr = -1; sr < nsb1; ++sr) {
      const bool intra = sr < 0;
      const int nrounds = intra ? G - 1 : G;
      for (int sp = warp; sp < nsp; sp += nwarps) {
        int SA, SB;
        if (intra) {
          SA = SB = 2 * sp + grp / (G / 2);
        } else if (sp == 0) {
          SA = sr;
          SB = nsb1;
        } else {
          SA = sr + sp;
          if (SA >= nsb1) SA -= nsb1;
          SB = sr - sp;
          if (SB < 0) SB += nsb1;
          if (SA > SB) {
            const int t = SA;
            SA = SB;
            SB = t;
          }
        }
        for (int k = 0; k < nrounds; ++k) {
          int P, Q;
          if (intra) {  // round-robin of the G blocks of the super-block: G / 2 groups, pair index gi
            constexpr int G1 = G - 1, GH = G / 2;
            const int gi = grp % GH;
            if (gi == 0) {
              P = k;
              Q = G1;
            } else {
              P = (k + gi) % G1;
              Q = (k - gi + G1) % G1;
            }
            if (P > Q) {
              const int t = P;
              P = Q;
              Q = t;
            }
          } else {
            P = grp;
            Q = (grp + k) % G;
          }
          P += SA * G;
          Q += SB * G;
          const int cp0 = 2 * P, cp1 = cp0 + 1, cq0 = 2 * Q, cq1 = cq0 + 1;
          const bool vp0 = cp0 < h, vp1 = cp1 < h, vq0 = cq0 < h, vq1 = cq1 < h;
          double* wp0 = vp0 ? Wlast - (size_t)cp0 * ld : const_cast<double*>(zcol);
          double* wp1 = vp1 ? Wlast - (size_t)cp1 * ld : const_cast<double*>(zcol);
          double* wq0 = vq0 ? Wlast - (size_t)cq0 * ld : const_cast<double*>(zcol);
          double* wq1 = vq1 ? Wlast - (size_t)cq1 * ld : const_cast<double*>(zcol);
          const int np0 = vp0 ? cp0 : nodummy, np1 = vp1 ? cp1 : nodummy;
          const int nq0 = vq0 ? cq0 : nodummy, nq1 = vq1 ? cq1 : nodummy;
          double x0[R], x1[R], y0[R], y1[R];
          jreg_load<JG, R>(wp0, lane, x0);
          jreg_load<JG, R>(wp1, lane, x1);
          jreg_load<JG, R>(wq0, lane, y0);
          jreg_load<JG, R>(wq1, lane, y1);
          double a0 = nrm[np0], a1 = nrm[np1], b0 = nrm[nq0], b1 = nrm[nq1];
          int r00, r11, r01, r10;
          jreg_rotate2<JG, R>(lane, x0, y0, a0, b0, x1, y1, a1, b1, r00, r11);
          jreg_rotate2<JG, R>(lane, x0, y1, a0, b1, x1, y0, a1, b0, r01, r10);
#ifndef SMRT_SIMT_EMULATION  // (the emulated shuffles above are warp barriers already)
          __syncwarp();  // every lane of the group has read the tracked norms before lane 0 replaces them
#endif
          // a column that rotated is a real column (missing ones have zero norm): unpredicated stores
          if ((r00 | r01) & 2) {
            jreg_store<JG, R>(wp0, lane, x0);
            if (lane == 0) nrm[np0] = a0;
          }
          if ((r11 | r10) & 2) {
            jreg_store<JG, R>(wp1, lane, x1);
            if (lane == 0) nrm[np1] = a1;
          }
          if ((r00 | r10) & 2) {
            jreg_store<JG, R>(wq0, lane, y0);
            if (lane == 0) nrm[nq0] = b0;
          }
          if ((r11 | r01) & 2) {
            jreg_store<JG, R>(wq1, lane, y1);
            if (lane == 0) nrm[nq1] = b1;
          }
          notconv |= (r00 | r11 | r01 | r10) & 1;
          __syncwarp();  // the blocks move to another group of this warp in the next round
        }
      }
      __syncthreads();
    }
    ++sweeps;
    if (!__syncthreads_or(notconv) || sweeps >= SMRT_JACOBI_MAX_SWEEPS) break;
  }
  return sweeps;
}

// dispatch on the lanes per group (8 or 16) and the rows per lane: JG x R covers the padded rows of the operand
SMRT_DEV int block_jacobi_svd_fast(double* W, int ld, int h, double* nrm, const double* zcol, int jg = 8) {
  if (jg == 16) {
    if (ld <= 18) return block_jacobi_svd_reg_sb<16, 1>(W, ld, h, nrm, zcol);
    if (ld <= 34) return block_jacobi_svd_reg_sb<16, 2>(W, ld, h, nrm, zcol);
    if (ld <= 50) return block_jacobi_svd_reg_sb<16, 3>(W, ld, h, nrm, zcol);
    if (ld <= 66) return block_jacobi_svd_reg_sb<16, 4>(W, ld, h, nrm, zcol);
    if (ld <= 98) return block_jacobi_svd_reg_sb<16, 6>(W, ld, h, nrm, zcol);
    return block_jacobi_svd_reg_sb<16, 8>(W, ld, h, nrm, zcol);  // ld = 130: up to 128 rows (zcol holds 128 zeros)
  }
  if (ld <= 18) return block_jacobi_svd_reg<8, 2>(W, ld, h, nrm, zcol);
  if (ld <= 34) return block_jacobi_svd_reg<8, 4>(W, ld, h, nrm, zcol);
  if (ld <= 50) return block_jacobi_svd_reg<8, 6>(W, ld, h, nrm, zcol);
  return block_jacobi_svd_reg<8, 8>(W, ld, h, nrm, zcol);
}

// ------------------------------------------------------------------------------- triangular solve C^T Z = W, in place
// C: h x h lower triangular (column-major, ldc); W: h x h (column-major, ldw), overwritten by Z = C^-T W; rdiag[j] =
// 1 / C(j, j) (block-shared).  Every column of W is an independent back substitution: TWO adjacent lanes own a column
// and walk it bottom-up in blocks of 8 unknowns held in registers (both lanes solve the 8 x 8 diagonal block
// redundantly, then split the update of the rows above between them), so the only synchronisation is one __syncwarp
// per block and all C operands are warp-wide broadcasts.  Called by every thread of the block.
template <bool kPacked>
SMRT_DEV void block_backsolve_lt(const LowerMat<kPacked>& C, double* W, int ldw, int h,
                                 const double* SMRT_RESTRICT rdiag) {
  const int NT = blockDim.x, tid = threadIdx.x;
  const int half = tid & 1;
  const int nblk = (h + 7) >> 3;
  for (int c0 = 0; c0 < h; c0 += NT / 2) {
    const int c = c0 + (tid >> 1);
    const bool act = c < h;
    double* x = W + (size_t)(act ? c : 0) * ldw;
    for (int jbk = nblk - 1; jbk >= 0; --jbk) {
      const int jb = jbk * 8;
      const int nv = (h - jb < 8) ? (h - jb) : 8;
      double z[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) z[jj] = (act && jj < nv) ? x[jb + jj] : 0.0;
      __syncwarp();  // the partner lane has read the block too before lane 0 overwrites it with the solution
#pragma unroll
      for (int jj = 7; jj >= 0; --jj) {
        if (jj < nv) {
          z[jj] *= rdiag[jb + jj];
#pragma unroll
          for (int ii = 0; ii < jj; ++ii) z[ii] = fma(-C.at(jb + jj, jb + ii), z[jj], z[ii]);
        }
      }
      if (act && half == 0) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          if (jj < nv) x[jb + jj] = z[jj];
      }
      // rows above the block, interleaved between the two lanes, two rows in flight per lane
      if (act) {
        int i = half;
        for (; i + 2 < jb; i += 4) {
          const double* ca = C.col(i) + jb;
          const double* cb = C.col(i + 2) + jb;
          double xa = x[i], xb = x[i + 2];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            if (jj < nv) {
              xa = fma(-ca[jj], z[jj], xa);
              xb = fma(-cb[jj], z[jj], xb);
            }
          }
          x[i] = xa;
          x[i + 2] = xb;
        }
        if (i < jb) {
          const double* ca = C.col(i) + jb;
          double xa = x[i];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            if (jj < nv) xa = fma(-ca[jj], z[jj], xa);
          x[i] = xa;
        }
      }
      __syncwarp();
    }
  }
}

// =====================================================================================================================
// Second-generation primitives used by the boundary kernel: pointer-operand GEMMs without guards in the inner loop and
// Gauss-Jordan eliminations whose every step is ONE block barrier (pivot search done redundantly by every warp,
// implicit permutation instead of physical swaps, no separate scaling pass).
// =====================================================================================================================

// C(i,j) = epi(i, j, sum_{k<K} A(i,k) B(k,j)) for i < M, j < N; A column-major (lda), column j of B at bcol(j).
// 4 x 4 register tile per thread (rows strided by 16 so that consecutive lanes read consecutive rows), threads arranged
// 16 x (nthr/16); only the first `nthr` threads of the block work.  FP64 FMA rate vs shared-memory bandwidth needs
// >= 16 FMA per 8 loaded doubles, hence no smaller tiles.  Loads of out-of-range rows / columns are clamped to the last
// valid one (harmless), stores are guarded.
template <typename FB, typename FE>
SMRT_DEV void block_gemm_ptr(int nthr, int M, int N, int K, const double* SMRT_RESTRICT Am, int lda, FB bcol, FE epi) {
  const int tid = threadIdx.x;
  if (tid >= nthr || M <= 0 || N <= 0) return;
  const int TX = 16, TY = nthr / TX;
  const int tx = tid % TX, ty = tid / TX;
  for (int j0 = 0; j0 < N; j0 += TY * 4) {
    for (int i0 = 0; i0 < M; i0 += TX * 4) {
      int ir[4], jc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ir[u] = i0 + tx + u * TX;
        jc[u] = j0 + ty + u * TY;
      }
      if (ir[0] >= M || jc[0] >= N) continue;
      const double* ap[4];
      const double* bp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ap[u] = Am + (ir[u] < M ? ir[u] : M - 1);
        bp[u] = bcol(jc[u] < N ? jc[u] : N - 1);
      }
      double acc[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        double av[4], bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) av[u] = ap[u][(size_t)k * lda];
#pragma unroll
        for (int v = 0; v < 4; ++v) bv[v] = bp[v][k];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (ir[u] < M && jc[v] < N) epi(ir[u], jc[v], acc[u][v]);
    }
  }
}

// Two products sharing the B operand: C1 = A1 B, C2 = A2 B (M x N, inner K); 4 x 4 tile of each product per thread.
template <typename FE>
SMRT_DEV void block_gemm_dual(int nthr, int M, int N, int K, const double* SMRT_RESTRICT A1,
                              const double* SMRT_RESTRICT A2, int lda, const double* SMRT_RESTRICT Bm, int ldb, FE epi) {
  const int tid = threadIdx.x;
  if (tid >= nthr || M <= 0 || N <= 0) return;
  const int TX = 16, TY = nthr / TX;
  const int tx = tid % TX, ty = tid / TX;
  for (int j0 = 0; j0 < N; j0 += TY * 4) {
    for (int i0 = 0; i0 < M; i0 += TX * 4) {
      int ir[4], jc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ir[u] = i0 + tx + u * TX;
        jc[u] = j0 + ty + u * TY;
      }
      if (ir[0] >= M || jc[0] >= N) continue;
      size_t ao[4];
      const double* bp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ao[u] = (size_t)(ir[u] < M ? ir[u] : M - 1);
        bp[u] = Bm + (size_t)(jc[u] < N ? jc[u] : N - 1) * ldb;
      }
      double c1[4][4], c2[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) c1[u][v] = c2[u][v] = 0.0;
#pragma unroll 2
      for (int k = 0; k < K; ++k) {
        double a1[4], a2[4], bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a1[u] = A1[ao[u] + (size_t)k * lda];
          a2[u] = A2[ao[u] + (size_t)k * lda];
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) bv[v] = bp[v][k];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            c1[u][v] = fma(a1[u], bv[v], c1[u][v]);
            c2[u][v] = fma(a2[u], bv[v], c2[u][v]);
          }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (ir[u] < M && jc[v] < N) epi(ir[u], jc[v], c1[u][v], c2[u][v]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// "Deferred store" variants for the boundary kernel that keeps only [T | R] resident (two CTAs per SM): the operands F
// and G of a layer are staged by the TMA engine INTO THE BUFFERS THE RESULTS WILL OCCUPY, every thread computes ALL its
// output tiles in registers, the block synchronises, and only then the results overwrite the operands.
// Tiles: 4 x 4 per thread, rows tx + 16 u (M <= 64: one row pass), columns j0 + ty + TY v; at most SMRT_DEF_TILES column
// passes per thread (N <= 4 TY SMRT_DEF_TILES).
#ifdef SMRT_SIMT_EMULATION
#define SMRT_DEF_TILES 8  // the CPU tests run 64-thread blocks (TY = 4): 128 columns
#else
#define SMRT_DEF_TILES 4  // >= 128 threads on the device (TY >= 8)
#endif

// out(i, j) = pre(i, j, sum_{k < K} A(i, k) B(k, j)) for i < M (<= 64), j < 2 h, evaluated in registers; then a block
// barrier; then post(i, j, value).  A column-major in shared memory (lda); B = [Bg | Bf], two compact h x h blocks
// (column-major, leading dimension h; K <= h rows used).  rows i >= Mk contribute no product (the sum is zero there).
// Every thread of the block must call.  The k loop is the outer one: the four A operands of a step serve all the
// column tiles of the thread.  The first half of a thread's tiles covers columns of Bg, the second half the SAME
// columns of Bf: one set of 2 SMRT_DEF_TILES element offsets and two running pointers address all the B operands
// (the version with one pointer per column re-derived its 16 addresses in every k step: the kernel sits at the
// register limit).  Requires h <= 2 TY SMRT_DEF_TILES.
template <typename FPRE, typename FPOST>
SMRT_DEV void block_gemm_split_deferred(int M, int Mk, int h, int K, const double* SMRT_RESTRICT Am, int lda,
                                        const double* SMRT_RESTRICT Bg, const double* SMRT_RESTRICT Bf, FPRE pre,
                                        FPOST post) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int TX = 16, TY = NT / TX;
  const int tx = tid % TX, ty = tid / TX;
  constexpr int HT = SMRT_DEF_TILES / 2;  // tiles per half
  double val[SMRT_DEF_TILES][4][4];
  int ao[4], bo[HT][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = tx + u * TX;
    ao[u] = i < Mk ? i : (Mk > 0 ? Mk - 1 : 0);
  }
#pragma unroll
  for (int t = 0; t < HT; ++t)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int jl = ty + TY * (4 * t + v);
      bo[t][v] = (jl < h ? jl : h - 1) * h;  // clamped: the reads stay inside the operand
      SMRT_KEEP_INT(bo[t][v]);
#pragma unroll
      for (int u = 0; u < 4; ++u) val[t][u][v] = val[HT + t][u][v] = 0.0;
    }
  if (Mk > 0 && tx < M) {
    const double* ap = Am;
    const double* gp = Bg;
    const double* fp = Bf;
#pragma unroll 1
    for (int k = 0; k < K; ++k, ap += lda, ++gp, ++fp) {
      double av[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) av[u] = ap[ao[u]];
#pragma unroll
      for (int t = 0; t < HT; ++t) {
        if (ty + TY * 4 * t < h) {  // uniform over the threads of a tile column
          double bv[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) bv[v] = gp[bo[t][v]];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) val[t][u][v] = fma(av[u], bv[v], val[t][u][v]);
#pragma unroll
          for (int v = 0; v < 4; ++v) bv[v] = fp[bo[t][v]];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) val[HT + t][u][v] = fma(av[u], bv[v], val[HT + t][u][v]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < SMRT_DEF_TILES; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int i = tx + u * TX, jl = ty + TY * (4 * (t % HT) + v), j = (t < HT) ? jl : h + jl;
        if (i < M && jl < h) val[t][u][v] = pre(i, j, (i < Mk) ? val[t][u][v] : 0.0);
      }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < SMRT_DEF_TILES; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int i = tx + u * TX, jl = ty + TY * (4 * (t % HT) + v), j = (t < HT) ? jl : h + jl;
        if (i < M && jl < h) post(i, j, val[t][u][v]);
      }
}

// Two products sharing the B operand, C1 = A1 B and C2 = A2 B (M x N, inner K, M <= 64, N <= 4 TY SMRT_DEF_TILES / 2):
// pre(i, j, c1, c2) turns the two sums IN PLACE into the two values to keep (it may read A1 / A2), then a block
// barrier, then post(i, j, v1, v2) stores them (it may overwrite A1 / A2).  Every thread of the block must call.
template <typename FPRE, typename FPOST>
SMRT_DEV void block_gemm_dual_deferred(int M, int N, int K, const double* SMRT_RESTRICT A1, const double* SMRT_RESTRICT A2,
                                       int lda, const double* SMRT_RESTRICT Bm, int ldb, FPRE pre, FPOST post) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int TX = 16, TY = NT / TX;
  const int tx = tid % TX, ty = tid / TX;
  constexpr int NTL = SMRT_DEF_TILES / 2;
  double v1[NTL][4][4], v2[NTL][4][4];
  int ao[4], bo[NTL][4];  // element offsets; the operands are addressed through three running pointers
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = tx + u * TX;
    ao[u] = i < M ? i : M - 1;
  }
#pragma unroll
  for (int t = 0; t < NTL; ++t)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = t * TY * 4 + ty + v * TY;
      bo[t][v] = (j < N ? j : N - 1) * ldb;
      SMRT_KEEP_INT(bo[t][v]);
#pragma unroll
      for (int u = 0; u < 4; ++u) v1[t][u][v] = v2[t][u][v] = 0.0;
    }
  if (tx < M) {
    const double* a1p = A1;
    const double* a2p = A2;
    const double* bq = Bm;
#pragma unroll 1
    for (int k = 0; k < K; ++k, a1p += lda, a2p += lda, ++bq) {
      double a1[4], a2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a1[u] = a1p[ao[u]];
        a2[u] = a2p[ao[u]];
      }
#pragma unroll
      for (int t = 0; t < NTL; ++t) {
        if (t * TY * 4 + ty < N) {
          double bv[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) bv[v] = bq[bo[t][v]];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              v1[t][u][v] = fma(a1[u], bv[v], v1[t][u][v]);
              v2[t][u][v] = fma(a2[u], bv[v], v2[t][u][v]);
            }
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < NTL; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int i = tx + u * TX, j = t * TY * 4 + ty + v * TY;
        if (i < M && j < N) pre(i, j, v1[t][u][v], v2[t][u][v]);
      }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < NTL; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int i = tx + u * TX, j = t * TY * 4 + ty + v * TY;
        if (i < M && j < N) post(i, j, v1[t][u][v], v2[t][u][v]);
      }
}

// Matrix-vector products by the whole block (the 4 x 4 tiled GEMMs above keep only 16 threads busy for one right-hand
// side): y1 = A1 x and y2 = A2 x for column-major M x K matrices (lda), x in shared memory.  Thread (i, p) sums the
// columns k = p, p + tpr, ... of row i (lanes along rows: conflict-free), the tpr partial sums per row meet in `part`
// (block-shared double[2 * tpr * M], tpr = min(blockDim.x / M, 8) >= 1).  epi(i, y1_i, y2_i) is called once per row.
// A2 may be NULL (single product, y2 = 0).  Synchronises internally (two barriers); requires blockDim.x >= M.
template <typename FE>
SMRT_DEV void block_matvec_dual(int M, int K, const double* SMRT_RESTRICT A1, const double* SMRT_RESTRICT A2, int lda,
                                const double* SMRT_RESTRICT x, double* part, FE epi) {
  const int NT = blockDim.x, tid = threadIdx.x;
  int tpr = NT / M;
  if (tpr > 8) tpr = 8;
  const int i = tid % M, p = tid / M;
  if (p < tpr) {
    double s1a = 0.0, s1b = 0.0, s2a = 0.0, s2b = 0.0;
    int k = p;
    for (; k + tpr < K; k += 2 * tpr) {
      const double x0 = x[k], x1 = x[k + tpr];
      s1a = fma(A1[(size_t)k * lda + i], x0, s1a);
      s1b = fma(A1[(size_t)(k + tpr) * lda + i], x1, s1b);
      if (A2) {
        s2a = fma(A2[(size_t)k * lda + i], x0, s2a);
        s2b = fma(A2[(size_t)(k + tpr) * lda + i], x1, s2b);
      }
    }
    if (k < K) {
      const double x0 = x[k];
      s1a = fma(A1[(size_t)k * lda + i], x0, s1a);
      if (A2) s2a = fma(A2[(size_t)k * lda + i], x0, s2a);
    }
    part[(size_t)(2 * p) * M + i] = s1a + s1b;
    part[(size_t)(2 * p + 1) * M + i] = s2a + s2b;
  }
  __syncthreads();
  if (tid < M) {
    double y1 = 0.0, y2 = 0.0;
    for (int q = 0; q < tpr; ++q) {
      y1 += part[(size_t)(2 * q) * M + tid];
      y2 += part[(size_t)(2 * q + 1) * M + tid];
    }
    epi(tid, y1, y2);
  }
  __syncthreads();
}

// warp-wide argmax of (value, index) with ties to the smaller index; every lane returns the winner
SMRT_DEV void warp_argmax(double& best, int& bi) {
  for (int off = 16; off > 0; off >>= 1) {
    double ob = __shfl_xor_sync(0xffffffffu, best, off, 32);
    int oi = __shfl_xor_sync(0xffffffffu, bi, off, 32);
    if (ob > best || (ob == best && oi < bi)) {
      best = ob;
      bi = oi;
    }
  }
}

// elimination pass of one Gauss-Jordan step: lanes along rows (RPL rows per lane), warps along columns, two columns per
// iteration for instruction-level parallelism
template <int RPL>
SMRT_DEV void gj_rows_update(double* T, int ld, int h, int W, int j, int p, double inv, int lane, int warp, int nwarp) {
  const double* colj = T + (size_t)j * ld;
  double mrow[RPL];
#pragma unroll
  for (int u = 0; u < RPL; ++u) {
    const int i = lane + 32 * u;
    mrow[u] = (i < h && i != p) ? -(colj[i] * inv) : 0.0;
  }
  int c = j + 1 + warp;
  for (; c + nwarp < W; c += 2 * nwarp) {
    double* col0 = T + (size_t)c * ld;
    double* col1 = col0 + (size_t)nwarp * ld;
    const double p0 = col0[p], p1 = col1[p];
#pragma unroll
    for (int u = 0; u < RPL; ++u) {
      const int i = lane + 32 * u;
      if (i < h) {
        const double v0 = col0[i], v1 = col1[i];
        col0[i] = fma(mrow[u], p0, v0);
        col1[i] = fma(mrow[u], p1, v1);
      }
    }
  }
  if (c < W) {
    double* col0 = T + (size_t)c * ld;
    const double p0 = col0[p];
#pragma unroll
    for (int u = 0; u < RPL; ++u) {
      const int i = lane + 32 * u;
      if (i < h) col0[i] = fma(mrow[u], p0, col0[i]);
    }
  }
}

// Gauss-Jordan elimination by ROWS with partial pivoting on T = [A | R] (h rows, W >= h columns, column-major, ld):
// afterwards, for every unknown k, row rowof[k] of the right block holds piv_k * (A^-1 R)(k, :), piv_k = T(rowof[k], k).
// No physical swaps, no scaling pass, one barrier per step.  rowstep/rowof: block-shared int[h].
// Returns 1 if a pivot vanishes (singular / non finite), else 0 — the same value in every thread.
SMRT_DEV_NOINLINE int block_gj_rows(double* T, int ld, int h, int W, int* rowstep, int* rowof) {
  const int NT = blockDim.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
  for (int i = tid; i < h; i += NT) rowstep[i] = -1;
  __syncthreads();
  for (int j = 0; j < h; ++j) {
    const double* colj = T + (size_t)j * ld;
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = lane; i < h; i += 32) {
      int st = rowstep[i];
      if (st < 0) {
        double v = fabs(colj[i]);
        if (v > best) {
          best = v;
          bi = i;
        }
      }
    }
    warp_argmax(best, bi);
    if (!(best > 0.0) || !(best < 1e300)) return 1;  // identical decision in every warp
    const int p = bi;
    const double inv = 1.0 / colj[p];
    // columns still to update: (j, h) of the left block and the whole right block
    if (h <= 64)
      gj_rows_update<2>(T, ld, h, W, j, p, inv, lane, warp, nwarp);
    else if (h <= 128)
      gj_rows_update<4>(T, ld, h, W, j, p, inv, lane, warp, nwarp);
    else
      gj_rows_update<8>(T, ld, h, W, j, p, inv, lane, warp, nwarp);
    __syncthreads();
    // the row is marked only now: slower warps were still scanning rowstep[] for this step's pivot (racecheck)
    if (tid == 0) {
      rowstep[p] = j;
      rowof[j] = p;
    }
    __syncthreads();
  }
  return 0;
}

// Gauss-Jordan elimination by COLUMNS with partial (column) pivoting on the stacked pair [S; K] (S: h x h, K: m x h):
// afterwards (K S^-1)(:, k) = K(:, colof[k]) / S(k, colof[k]).  One barrier per step.
SMRT_DEV_NOINLINE int block_gj_cols(double* S, int lds, double* Km, int ldk, int h, int m, int* colstep, int* colof) {
  const int NT = blockDim.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
  for (int i = tid; i < h; i += NT) colstep[i] = -1;
  __syncthreads();
  for (int j = 0; j < h; ++j) {
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int c = lane; c < h; c += 32) {
      int st = colstep[c];
      if (st < 0) {
        double v = fabs(SMRT_AT(S, lds, j, c));
        if (v > best) {
          best = v;
          bi = c;
        }
      }
    }
    warp_argmax(best, bi);
    if (!(best > 0.0) || !(best < 1e300)) return 1;
    const int p = bi;
    const double inv = 1.0 / SMRT_AT(S, lds, j, p);
    const double* sp = S + (size_t)p * lds;
    const double* kp = Km + (size_t)p * ldk;
    // warps along columns (two per iteration for ILP), lanes along rows
    for (int c0 = warp; c0 < h; c0 += 2 * nwarp) {
      const int c1 = c0 + nwarp;
      const bool ok0 = (c0 != p), ok1 = (c1 < h) && (c1 != p);
      double* s0 = S + (size_t)c0 * lds;
      double* s1 = S + (size_t)(ok1 ? c1 : c0) * lds;
      double* k0 = Km + (size_t)c0 * ldk;
      double* k1 = Km + (size_t)(ok1 ? c1 : c0) * ldk;
      const double m0 = ok0 ? -(s0[j] * inv) : 0.0;
      const double m1 = ok1 ? -(s1[j] * inv) : 0.0;
      if (ok0 && ok1) {
        for (int i = j + 1 + lane; i < h; i += 32) {
          const double pv = sp[i];
          s0[i] = fma(m0, pv, s0[i]);
          s1[i] = fma(m1, pv, s1[i]);
        }
        for (int i = lane; i < m; i += 32) {
          const double pv = kp[i];
          k0[i] = fma(m0, pv, k0[i]);
          k1[i] = fma(m1, pv, k1[i]);
        }
      } else if (ok0 || ok1) {
        double* sc = ok0 ? s0 : s1;
        double* kc = ok0 ? k0 : k1;
        const double mm = ok0 ? m0 : m1;
        for (int i = j + 1 + lane; i < h; i += 32) sc[i] = fma(mm, sp[i], sc[i]);
        for (int i = lane; i < m; i += 32) kc[i] = fma(mm, kp[i], kc[i]);
      }
      // S(j, c) is mathematically zero now; it is never read again, and must NOT be written here: slower warps may
      // still be scanning row j for the pivot of this step
    }
    __syncthreads();
    if (tid == 0) {  // marked only now (slower warps were still scanning colstep[] for this step's pivot)
      colstep[p] = j;
      colof[j] = p;
    }
    __syncthreads();
  }
  return 0;
}


// =====================================================================================================================
// BLOCKED Gauss-Jordan by rows with partial pivoting (h <= 64).  The unblocked eliminations above rewrite the whole
// trailing matrix at every step (one FMA per shared-memory load + store: the boundary kernel spent 60 % of its time
// there at 13 % FP64 utilisation).  Here the steps are grouped in panels of SMRT_GJ_NB columns:
//   1. warp 0 factorises the panel with its columns in REGISTERS (lanes along rows; the pivot is found by ONE REDUX on
//      a key packing the exponent / leading mantissa bits with the lane number; the pivot row is broadcast by shuffles;
//      the reciprocal comes from a MUFU seed + two Newton steps; straight-line code for the 8 steps, the singularity
//      test is deferred to the end of the panel) and builds V (h x nb) such that the nb elimination steps applied to
//      any other column x amount to  x <- x + V x[P]  (P = the panel's pivot rows, old values).  Derivation: every
//      step is G_k = I + m_k e_{p_k}^T, so G_nb ... G_1 differs from I only in the columns P, and
//      V(:, k) = G e_{p_k} - e_{p_k} obeys the same update rule as an ordinary column, starting from m_k.
//   2. the rank-nb update of a column is done by one half-warp (lanes along rows, 4 rows per lane, V held in
//      registers for the whole panel): it reads the old pivot-row entries of its column, then rewrites the column:
//      nb FMAs per shared-memory load + store of an element.
//   3. look-ahead: warp 0 updates the columns of the NEXT panel first and factorises it while the other warps update
//      the remaining columns: one block barrier per panel, and the (serial) panel factorisations are the only
//      critical path.
// Conventions: two column blocks (left: the h x h system, right: nR further columns), implicit
// row permutation, unscaled rows:  (A^-1 R)(k, :) = R(rowof[k], :) * ipiv[k].  The left block is destroyed.
// Scratch (block-shared): Vbuf double[2 * h * SMRT_GJ_NB], rowof int[h], ipiv double[h], flag int[1].
// Returns 1 in every thread if a pivot vanishes / is not finite.  blockDim.x >= 64.
// =====================================================================================================================
#ifndef SMRT_GJ_NB
#define SMRT_GJ_NB 4
#endif

// pivot of a panel column held in registers (col[u] = row lane + 32 u): largest |value| among the rows not used yet.
// Key = high word of |value| (exponent + 20 mantissa bits) with the 5 low bits replaced by 31 - lane: ONE REDUX gives
// the maximum and its (lowest) lane; any element within 2^-15 of the maximum is as good a pivot.  Every lane computes
// the reciprocal of its own candidate while the REDUX is in flight (MUFU seed + two Newton steps), so the winner's
// reciprocal arrives with the same shuffle round as the pivot itself.  Returns the lane and register slot of the
// pivot row, the (signed) pivot and its reciprocal; no candidate / zero column -> pv = 0.
// split in two halves so that the caller can place independent work between the REDUX and the shuffles that depend on it
template <int RPL>
SMRT_DEV unsigned gj_pivot_begin(const double (&col)[RPL], unsigned used, int lane, int h, int& bu, double& mine,
                                 double& inv) {
  double bv = -1.0;
  mine = 0.0;
  bu = 0;
#pragma unroll
  for (int u = 0; u < RPL; ++u) {
    const double a = fabs(col[u]);
    if ((lane + 32 * u) < h && !((used >> u) & 1u) && a > bv) {
      bv = a;
      bu = u;
      mine = col[u];
    }
  }
  const unsigned key = (bv >= 0.0) ? (((unsigned)__double2hiint(bv) & ~31u) | (unsigned)(31 - lane)) : 0u;
  const unsigned mx = __reduce_max_sync(0xffffffffu, key);
  inv = smrt_rcp_approx(mine);
  inv = fma(inv, fma(-mine, inv, 1.0), inv);
  inv = fma(inv, fma(-mine, inv, 1.0), inv);
  return mx;
}
SMRT_DEV void gj_pivot_finish(unsigned mx, int bu, double mine, double myinv, int& pl, int& pu, double& pv,
                              double& pinv) {
  pl = 31 - (int)(mx & 31u);
  pu = __shfl_sync(0xffffffffu, bu, pl, 32);
  pv = __shfl_sync(0xffffffffu, mine, pl, 32);
  pinv = __shfl_sync(0xffffffffu, myinv, pl, 32);
}

// Panel factorisation by ONE warp, columns in registers.  The step loop is rolled (small instruction footprint): the
// panel columns shift left by one position per step so that the current column is always slot 0, the V columns shift
// right (the column created at step k ends in slot npc - 1 - k), and the pivot search of the NEXT step is issued as
// soon as its column is up to date, ahead of the other updates of the current step (software pipelining of the only
// loop-carried dependency).
template <int RPL>
SMRT_DEV void gj_panel_warp(const double* Lb, int ldl, int h, int j0, int npc, unsigned& used, int lane,
                            int* rowof, double* ipiv, double* Vout, int ldv, int* flag) {
  double pc[RPL][SMRT_GJ_NB], v[RPL][SMRT_GJ_NB];
#pragma unroll
  for (int u = 0; u < RPL; ++u) {
    const int row = lane + 32 * u;
#pragma unroll
    for (int c = 0; c < SMRT_GJ_NB; ++c) {
      pc[u][c] = (row < h && c < npc) ? Lb[(size_t)(j0 + c) * ldl + row] : 0.0;
      v[u][c] = 0.0;
    }
  }
  int bad = 0;
  int pl, pu;
  double pv, inv;
  {
    double col0[RPL], mine, myinv;
    int bu;
#pragma unroll
    for (int u = 0; u < RPL; ++u) col0[u] = pc[u][0];
    const unsigned mx = gj_pivot_begin<RPL>(col0, used, lane, h, bu, mine, myinv);
    gj_pivot_finish(mx, bu, mine, myinv, pl, pu, pv, inv);
  }
#pragma unroll 1
  for (int k = 0; k < npc; ++k) {
    bad |= (!(fabs(pv) > 0.0)) | (!(fabs(pv) < 1e300));
    if (lane == 0) {
      rowof[j0 + k] = pl + 32 * pu;
      ipiv[j0 + k] = inv;
    }
    if (lane == pl) used |= 1u << pu;
    double m[RPL], nxt[RPL];
#pragma unroll
    for (int u = 0; u < RPL; ++u) m[u] = (lane == pl && u == pu) ? 0.0 : -(pc[u][0] * inv);
    {  // the next column first, and its pivot search right away
      double sel = pc[0][1];
#pragma unroll
      for (int u = 1; u < RPL; ++u) sel = (pu == u) ? pc[u][1] : sel;
      const double pr = __shfl_sync(0xffffffffu, sel, pl, 32);
#pragma unroll
      for (int u = 0; u < RPL; ++u) nxt[u] = fma(m[u], pr, pc[u][1]);
    }
    int bu2;
    double mine2, myinv2;
    const unsigned mx2 = gj_pivot_begin<RPL>(nxt, used, lane, h, bu2, mine2, myinv2);
#pragma unroll
    for (int c = 2; c < SMRT_GJ_NB; ++c) {
      double sel = pc[0][c];
#pragma unroll
      for (int u = 1; u < RPL; ++u) sel = (pu == u) ? pc[u][c] : sel;
      const double pr = __shfl_sync(0xffffffffu, sel, pl, 32);
#pragma unroll
      for (int u = 0; u < RPL; ++u) pc[u][c - 1] = fma(m[u], pr, pc[u][c]);
    }
#pragma unroll
    for (int u = 0; u < RPL; ++u) {
      pc[u][SMRT_GJ_NB - 1] = 0.0;
      pc[u][0] = nxt[u];
    }
#pragma unroll
    for (int c = SMRT_GJ_NB - 1; c >= 1; --c) {
      double sel = v[0][c - 1];
#pragma unroll
      for (int u = 1; u < RPL; ++u) sel = (pu == u) ? v[u][c - 1] : sel;
      const double vr = __shfl_sync(0xffffffffu, sel, pl, 32);
#pragma unroll
      for (int u = 0; u < RPL; ++u) v[u][c] = fma(m[u], vr, v[u][c - 1]);
    }
#pragma unroll
    for (int u = 0; u < RPL; ++u) v[u][0] = m[u];
    gj_pivot_finish(mx2, bu2, mine2, myinv2, pl, pu, pv, inv);
  }
  if (bad && lane == 0) *flag = 1;
#pragma unroll
  for (int u = 0; u < RPL; ++u) {
    const int row = lane + 32 * u;
    if (row < h) {
#pragma unroll
      for (int c = 0; c < SMRT_GJ_NB; ++c) {
        const int kcol = npc - 1 - c;  // slot c holds the column created at step npc - 1 - c
        if (kcol >= 0) Vout[(size_t)kcol * ldv + row] = v[u][c];
      }
    }
  }
}

// rank-npc update  x <- x + V x[P]  of the columns c = cbeg + hw, cbeg + hw + nhw, ... < cend by half-warp number hw
// (of nhw, even; the two half-warps of a warp have hw = 2 i and 2 i + 1): lane lx owns the rows lx + 16 u (u < RT).
// prow: the npc pivot rows of the panel.  The trip count is WARP-uniform (a half-warp without a column of its own
// repeats its sibling's: same reads before the warp barrier, same values stored after it), so that the barrier between
// the reads of the pivot-row entries and the stores is a plain full-mask one; kFull: npc == SMRT_GJ_NB, no predicates
// on the panel index.
template <int RT, bool kFull>
SMRT_DEV void gj_update_cols_t(double* Lb, int ldl, double* Rb, int ldr, int h, int cbeg, int cend, int hw, int nhw,
                               int lx, int npc, const double* Vin, int ldv, const int* SMRT_RESTRICT prow) {
  const int cw0 = cbeg + (hw & ~1);
  if (cw0 >= cend) return;  // warp uniform
  double Vr[RT][SMRT_GJ_NB];
  int pr[SMRT_GJ_NB];
  bool rowok[RT];
#pragma unroll
  for (int k = 0; k < SMRT_GJ_NB; ++k) pr[k] = (kFull || k < npc) ? prow[k] : 0;
#pragma unroll
  for (int u = 0; u < RT; ++u) {
    const int row = lx + 16 * u;
    rowok[u] = row < h;
#pragma unroll
    for (int k = 0; k < SMRT_GJ_NB; ++k) Vr[u][k] = (rowok[u] && (kFull || k < npc)) ? Vin[k * ldv + row] : 0.0;
  }
  for (int cw = cw0; cw < cend; cw += nhw) {
    const int cc = cw + (hw & 1);
    const int c = (cc < cend) ? cc : cw;
    double* col = (c < h) ? Lb + c * ldl : Rb + (c - h) * ldr;
    double tp[SMRT_GJ_NB], acc[RT];
#pragma unroll
    for (int k = 0; k < SMRT_GJ_NB; ++k) tp[k] = (kFull || k < npc) ? col[pr[k]] : 0.0;  // old pivot-row entries
#pragma unroll
    for (int u = 0; u < RT; ++u) acc[u] = rowok[u] ? col[lx + 16 * u] : 0.0;
    __syncwarp();  // every lane has read the column before it is rewritten
#pragma unroll
    for (int k = 0; k < SMRT_GJ_NB; ++k)
#pragma unroll
      for (int u = 0; u < RT; ++u) acc[u] = fma(Vr[u][k], tp[k], acc[u]);
#pragma unroll
    for (int u = 0; u < RT; ++u)
      if (rowok[u]) col[lx + 16 * u] = acc[u];
  }
}
template <int RT>
SMRT_DEV void gj_update_cols(double* Lb, int ldl, double* Rb, int ldr, int h, int cbeg, int cend, int hw, int nhw,
                             int lx, int npc, const double* Vin, int ldv, const int* SMRT_RESTRICT prow) {
  if (npc == SMRT_GJ_NB)
    gj_update_cols_t<RT, true>(Lb, ldl, Rb, ldr, h, cbeg, cend, hw, nhw, lx, npc, Vin, ldv, prow);
  else
    gj_update_cols_t<RT, false>(Lb, ldl, Rb, ldr, h, cbeg, cend, hw, nhw, lx, npc, Vin, ldv, prow);
}

// one instance per (rows per lane of the panel warp, rows per lane of the update tiles); NOT inlined: the boundary
// kernel calls it from three places and the straight-line panel code is large (instruction-cache footprint)
// kKeepV: PRODUCT FORM.  The V of a panel is stored into the panel's own (dead) columns of the left block instead of
// the double buffer, so that after the call the left block holds the whole transformation: applying the panels in
// order, x <- x + V_p x[P_p] (P_p = rowof[4p .. 4p+3], old values), maps ANY further column b to the unscaled solution
// of the system (gj_apply_block): right-hand blocks that do not fit next to the left block in shared memory are
// eliminated afterwards, chunk by chunk.
// kSpare (blocks of 16 warps): the warps that share the panel warp's scheduler (warp % 4 == 0) take no part in the
// updates of the remaining columns: the serial panel chain is the critical path and runs faster alone on its scheduler
template <int RPL, int RT, bool kShared, bool kKeepV = false, bool kSpare = false>
SMRT_DEV_NOINLINE int block_gj_rows_blocked_t(double* Lb, int ldl, double* Rb, int ldr, int h, int nR, int* rowof,
                                              double* ipiv, double* Vbuf, int* flag) {
  if (kShared) {  // every operand lives in the block's shared memory
    SMRT_ASSUME_SHARED(Lb);
    SMRT_ASSUME_SHARED(Rb);
    SMRT_ASSUME_SHARED(ipiv);
    SMRT_ASSUME_SHARED(Vbuf);
  }
  SMRT_ASSUME_SHARED(rowof);
  SMRT_ASSUME_SHARED(flag);
  const int NT = blockDim.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
  const int lx = tid & 15;
  const int W = h + nR;
  unsigned used = 0u;  // warp 0: bit u = row lane + 32 u already served as a pivot
  if (tid == 0) *flag = 0;
  __syncthreads();
  int buf = 1;
  // iteration j0 = -NB only factorises the first panel; iteration j0 >= 0 applies panel j0 (V in Vbuf[buf]) while warp 0
  // looks ahead to the next one
  for (int j0 = -SMRT_GJ_NB; j0 < h; j0 += SMRT_GJ_NB, buf ^= 1) {
    const bool cur = j0 >= 0;
    const int npc = cur ? ((h - j0 < SMRT_GJ_NB) ? (h - j0) : SMRT_GJ_NB) : 0;
    const int cstart = cur ? j0 + npc : 0;
    const double* Vin = kKeepV ? Lb + (size_t)(cur ? j0 : 0) * ldl : Vbuf + (size_t)buf * h * SMRT_GJ_NB;
    const int ldv = kKeepV ? ldl : h;
    const bool more = cstart < h;
    const int npn = more ? ((h - cstart < SMRT_GJ_NB) ? (h - cstart) : SMRT_GJ_NB) : 0;  // width of the next panel
    if (cur) {
      // the next panel is brought up to date first: by warps 0 and 1 together (one column per half-warp, then a
      // 64-thread named barrier) when the block has more than two warps, else by warp 0 alone; the warps >= 1 share
      // the remaining columns
      if (more) {
        if (nwarp > 2) {
          if (warp < 2) {
            gj_update_cols<RT>(Lb, ldl, Rb, ldr, h, cstart, cstart + npn, tid >> 4, 4, lx, npc, Vin, ldv, rowof + j0);
            smrt_named_barrier(1, 64);
          }
        } else if (warp == 0) {
          gj_update_cols<RT>(Lb, ldl, Rb, ldr, h, cstart, cstart + npn, (tid >> 4) & 1, 2, lx, npc, Vin, ldv, rowof + j0);
        }
      }
      if (kSpare) {
        if ((warp & 3) != 0)  // 3 update warps per scheduler group of 4: index warp - 1 - warp / 4 among 3 nwarp / 4
          gj_update_cols<RT>(Lb, ldl, Rb, ldr, h, cstart + npn, W, 2 * (warp - 1 - (warp >> 2)) + ((tid >> 4) & 1),
                             2 * (nwarp - (nwarp >> 2)), lx, npc, Vin, ldv, rowof + j0);
      } else if (warp > 0)
        gj_update_cols<RT>(Lb, ldl, Rb, ldr, h, cstart + npn, W, (tid >> 4) - 2, 2 * (nwarp - 1), lx, npc, Vin, ldv,
                           rowof + j0);
    }
    if (warp == 0 && more) {
      __syncwarp();
      gj_panel_warp<RPL>(Lb, ldl, h, cstart, npn, used, lane, rowof, ipiv,
                         kKeepV ? Lb + (size_t)cstart * ldl : Vbuf + (size_t)(buf ^ 1) * h * SMRT_GJ_NB, ldv, flag);
    }
    __syncthreads();
    if (*flag) return 1;
  }
  return 0;
}
// blockDim.x >= 64 (warp 0 factorises, the others update).  kShared: Lb, Rb, ipiv, Vbuf are in shared memory (rowof and
// flag always are)
// hcode: the largest block of the plan (h_max): with 32 < h_max <= 64 every block above 32 unknowns uses the ONE
// instantiation sized for 64 rows (layers of a snowpack keep different stream counts: alternating between the <2, 3> and
// <2, 4> copies of the hot loop thrashed the instruction cache; a few masked rows cost less) — 0: pick by h alone
template <bool kShared>
SMRT_DEV int block_gj_rows_blocked(double* Lb, int ldl, double* Rb, int ldr, int h, int nR, int* rowof, double* ipiv,
                                   double* Vbuf, int* flag, int hcode = 0) {
  if (hcode > 48 && h > 32) return block_gj_rows_blocked_t<2, 4, kShared>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Vbuf, flag);
  if (h <= 16) return block_gj_rows_blocked_t<1, 1, kShared>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Vbuf, flag);
  if (h <= 32) return block_gj_rows_blocked_t<1, 2, kShared>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Vbuf, flag);
  if (h <= 48) return block_gj_rows_blocked_t<2, 3, kShared>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Vbuf, flag);
  return block_gj_rows_blocked_t<2, 4, kShared>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Vbuf, flag);
}

// =====================================================================================================================
// Blocks of 64 < h <= 128 unknowns (boundary kernel, one CTA of 512 threads per SM).  [A | B] with two h x h blocks no
// longer fits in shared memory, so the elimination is split: (1) the LEFT block is factorised in shared memory in
// product form (block_gj_factor: the blocked Gauss-Jordan above, every panel's V kept in the panel's columns), with the
// few right-hand-side columns riding along as before; (2) the RIGHT block is streamed through afterwards in chunks of
// up to 96 columns (gj_apply_block), each thread holding 8 rows of up to 3 columns in registers for the whole pass.
// =====================================================================================================================
SMRT_DEV int block_gj_factor(double* Lb, int ldl, double* Rb, int ldr, int h, int nR, int* rowof, double* ipiv, int* flag) {
  // (the V buffer argument is unused in product form: any shared-memory pointer)
  if (h <= 32) return block_gj_rows_blocked_t<1, 2, true, true, true>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Lb, flag);
  if (h <= 64) return block_gj_rows_blocked_t<2, 4, true, true, true>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Lb, flag);
  if (h <= 96) return block_gj_rows_blocked_t<3, 6, true, true, true>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Lb, flag);
  return block_gj_rows_blocked_t<4, 8, true, true, true>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Lb, flag);
}

// the plain (two resident blocks) elimination for up to 128 unknowns: Vbuf double[>= 8 h]
SMRT_DEV int block_gj_rows_blocked_mid(double* Lb, int ldl, double* Rb, int ldr, int h, int nR, int* rowof, double* ipiv,
                                       double* Vbuf, int* flag) {
  if (h <= 32) return block_gj_rows_blocked_t<1, 2, true, false, true>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Vbuf, flag);
  if (h <= 64) return block_gj_rows_blocked_t<2, 4, true, false, true>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Vbuf, flag);
  if (h <= 96) return block_gj_rows_blocked_t<3, 6, true, false, true>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Vbuf, flag);
  return block_gj_rows_blocked_t<4, 8, true, false, true>(Lb, ldl, Rb, ldr, h, nR, rowof, ipiv, Vbuf, flag);
}

// Apply the product-form factorisation held in M (h x h, leading dimension ldm, a multiple of 2; panels of SMRT_GJ_NB = 4
// columns, pivot rows rowof[]) to ncols <= 32 NC columns at once.  Warp w owns the rows 8 w .. 8 w + 7, lane c the
// columns c, c + 32, ... (NC of them): a thread keeps its 8 x NC entries in registers through all the panels and every
// 16-byte V operand (a warp-wide broadcast) feeds 2 NC FMAs.  TWO panels per block barrier: the owners of the 8 pivot
// rows of a panel pair publish their entries as they are before the pair (xch: block-shared double[2 * 8 * 32 * NC],
// double buffered); every thread derives the entries the second panel meets from them,
//     z1 = x[P1] + V0[P1, :] z0,   z0 = x[P0]          (16 FMAs per column, V0[P1, :] are broadcast loads),
// and applies both panels,  x <- x + V0 z0 + V1 z1.
// load(i, c) gives the initial entry (i < h, c < ncols); store(i, c, v) receives the transformed entry of row i (the
// unscaled solution: row rowof[k] holds piv_k * x_k).  Every thread of the block must call; blockDim.x >= 32 ceil(h / 8).
template <int NC, typename FLoad, typename FStore>
SMRT_DEV_NOINLINE void gj_apply_block(const double* M, int ldm, int h, const int* rowof, int ncols, double* xch, FLoad load,
                             FStore store) {
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int r0 = 8 * w;
  const bool active = r0 < h;  // warp-uniform
  double x[NC][8];
#pragma unroll
  for (int q = 0; q < NC; ++q)
#pragma unroll
    for (int u = 0; u < 8; ++u)
      x[q][u] = (active && r0 + u < h && lane + 32 * q < ncols) ? load(r0 + u, lane + 32 * q) : 0.0;
  int buf = 0;
  for (int j0 = 0; j0 < h; j0 += 2 * SMRT_GJ_NB, buf ^= 1) {
    const int npc = (h - j0 < 2 * SMRT_GJ_NB) ? (h - j0) : 2 * SMRT_GJ_NB;  // pivot rows of this pair of panels
    double* xb = xch + buf * (8 * 32 * NC);
    int pr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pr[k] = (k < npc) ? rowof[j0 + k] : -8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if ((pr[k] >> 3) == w) {  // warp-uniform: this warp owns the pivot row
#pragma unroll
        for (int q = 0; q < NC; ++q) {
          // (an opaque select chain: written as a conditional expression the compiler turns it back into a dynamically
          // indexed array and moves x[][] to local memory)
          double val = x[q][0];
#pragma unroll
          for (int u = 1; u < 8; ++u) val = smrt_select_eq(pr[k] & 7, u, x[q][u], val);
          xb[(k * NC + q) * 32 + lane] = val;
        }
      }
    }
    __syncthreads();
    if (active) {
      // (register budget: the 4 x NC pivot values of the first panel stay live, those of the second panel are formed
      // and consumed one column at a time; the two updates commute)
      double z0[NC][4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int q = 0; q < NC; ++q) z0[q][k] = (k < npc) ? xb[(k * NC + q) * 32 + lane] : 0.0;
      const double* vcol = M + (size_t)j0 * ldm + r0;
#pragma unroll
      for (int k = 4; k < 8; ++k) {
        if (k < npc) {
          // the second panel meets its pivot row after the first one has been applied
          double zk[NC];
#pragma unroll
          for (int q = 0; q < NC; ++q) zk[q] = xb[(k * NC + q) * 32 + lane];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const double wv = M[(size_t)(j0 + kk) * ldm + pr[k]];
#pragma unroll
            for (int q = 0; q < NC; ++q) zk[q] = fma(wv, z0[q][kk], zk[q]);
          }
          const double2* vc = reinterpret_cast<const double2*>(vcol + (size_t)k * ldm);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const double2 v = vc[t];
#pragma unroll
            for (int q = 0; q < NC; ++q) {
              x[q][2 * t] = fma(v.x, zk[q], x[q][2 * t]);
              x[q][2 * t + 1] = fma(v.y, zk[q], x[q][2 * t + 1]);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < npc) {
          const double2* vc = reinterpret_cast<const double2*>(vcol + (size_t)k * ldm);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const double2 v = vc[t];
#pragma unroll
            for (int q = 0; q < NC; ++q) {
              x[q][2 * t] = fma(v.x, z0[q][k], x[q][2 * t]);
              x[q][2 * t + 1] = fma(v.y, z0[q][k], x[q][2 * t + 1]);
            }
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int q = 0; q < NC; ++q)
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (r0 + u < h && lane + 32 * q < ncols) store(r0 + u, lane + 32 * q, x[q][u]);
  }
  __syncthreads();  // the exchange buffers are free again
}
// all the h columns of an h x h right block (h <= 128): one pass of up to 96 columns (3 per thread: 8 x 3 entries and
// 8 x 3 pivot values per thread fit the 128-register budget of a 512-thread block), or two passes of 64
template <typename FLoad, typename FStore>
SMRT_DEV void gj_apply_all(const double* M, int ldm, int h, const int* rowof, double* xch, FLoad load, FStore store) {
  if (h <= 32) {
    gj_apply_block<1>(M, ldm, h, rowof, h, xch, load, store);
  } else if (h <= 64) {
    gj_apply_block<2>(M, ldm, h, rowof, h, xch, load, store);
  } else if (h <= 96) {
    gj_apply_block<3>(M, ldm, h, rowof, h, xch, load, store);
  } else {
    gj_apply_block<2>(M, ldm, h, rowof, 64, xch, load, store);
    gj_apply_block<2>(
        M, ldm, h, rowof, h - 64, xch, [&](int i, int c) { return load(i, c + 64); },
        [&](int i, int c, double v) { store(i, c + 64, v); });
  }
}

// C1 = A1 B (and C2 = A2 B when kDual) for M x N results with inner dimension K, M <= 128 rows, the columns
// [n0, n0 + 64) of the result per call... see the tile map below.  A1 / A2: column-major in GLOBAL memory (lda), streamed
// through shared-memory panels of SMRT_MG_KP columns (double buffered, the next panel is fetched into registers while the
// current one is multiplied: ONE block barrier per panel); B: column-major (ldb) in global memory (kBShared = false:
// its row panels are staged too) or in shared memory (kBShared = true: read in place).  512 threads: tx = tid % 32 owns
// the rows tx + 32 u (u < 4), ty = tid / 32 the columns n0 + ty + 16 v (v < NV).  epi(i, j, c1, c2) is called for i < M,
// j < n0 + 16 NV, j < N.  Rows of A beyond Ma (the product has only Ma <= M non-zero rows) read as zero.
// stage: block-shared double[SMRT_MG_STAGES * SMRT_MG_KP * (32 MU (kDual ? 2 : 1) + (kBShared ? 0 : 16 NV))] <= 6144.
#define SMRT_MG_KP 8
#define SMRT_MG_STAGES 3
// MU: row slabs of 32 (M <= 32 MU), NV: columns per thread (the call covers the columns [n0, n0 + 16 NV)): the caller picks
// both from the block size, so that the inner loop carries no guards and no padded tiles (guards inside the loop were
// measured: 2x slower)
// pre(i, j, f, g) loads the two values the epilogue needs besides the sums (from global memory): they are fetched for
// a whole row slab before the first epi(i, j, c1, c2, f, g) of the slab stores anything, so that their L2 latency is
// paid once per slab, not once per element (the stores of an element and the loads of the next cannot be reordered
// by the compiler).
template <bool kDual, bool kBShared, int NV, int MU, typename FP, typename FE>
SMRT_DEV void mid_gemm(int M, int Ma, int N, int n0, int K, const double* SMRT_RESTRICT A1, const double* SMRT_RESTRICT A2,
                       int lda, const double* Bm, int ldb, double* stage, FP pre, FE epi) {
  constexpr int KP = SMRT_MG_KP;
  constexpr int NA = kDual ? 2 : 1;
  constexpr int BW = 16 * NV;                           // columns of the result handled by this call
  constexpr int SA = 32 * MU;                           // row stride of a staged A panel
  constexpr int BUF = KP * (SA * NA + (kBShared ? 0 : BW));
  const int tid = threadIdx.x, NT = blockDim.x;
  const int tx = tid & 31, ty = tid >> 5;
  double c1[MU][NV], c2[MU][NV];  // (c2 is dead code unless kDual)
#pragma unroll
  for (int u = 0; u < MU; ++u)
#pragma unroll
    for (int v = 0; v < NV; ++v) c1[u][v] = c2[u][v] = 0.0;
  // staged element e of a panel: A part, e in [0, KP * SA * NA): (which = e / (KP * SA), kk = (e / SA) % KP, i = e % SA);
  // B part: (kk = e % KP, jj = e / KP) -> B(k0 + kk, n0 + jj).  The panels travel by asynchronous 8-byte copies (LDGSTS)
  // through a ring of SMRT_MG_STAGES buffers: the operands come from L2 / HBM (microseconds under load), one panel of
  // look-ahead in registers left the loop latency bound (profiles/r04_notes.txt)
  constexpr int NS = SMRT_MG_STAGES;
  constexpr int EA = (KP * SA * NA + 511) / 512, EB = kBShared ? 0 : (KP * BW + 511) / 512;
  auto issue = [&](int k0, double* buf) {
#pragma unroll
    for (int q = 0; q < EA; ++q) {
      const int e = tid + q * NT;
      if (e < KP * SA * NA) {
        const int i = e % SA, kk = (e / SA) % KP, which = e / (KP * SA);
        const double* Ap = (kDual && which == 1) ? A2 : A1;
        const bool ok = i < Ma && k0 + kk < K;
        smrt_cp_async8(buf + e, ok ? Ap + (size_t)(k0 + kk) * lda + i : Ap, ok);
      }
    }
#pragma unroll
    for (int q = 0; q < EB; ++q) {
      const int e = tid + q * NT;
      if (e < KP * BW) {
        const int kk = e % KP, jj = e / KP;  // consecutive threads walk down a column of B: contiguous in memory
        const bool ok = k0 + kk < K && n0 + jj < N;
        smrt_cp_async8(buf + KP * SA * NA + kk * BW + jj, ok ? Bm + (size_t)(n0 + jj) * ldb + k0 + kk : Bm, ok);
      }
    }
  };
#pragma unroll
  for (int st = 0; st < NS - 1; ++st) {
    if (st * KP < K) issue(st * KP, stage + st * BUF);
    smrt_cp_async_commit();
  }
  int cur = 0, nxt = NS - 1;
  for (int k0 = 0; k0 < K; k0 += KP) {
    smrt_cp_async_wait<NS - 2>();  // this thread's copies of the current panel have landed ...
    __syncthreads();               // ... and everybody's; the buffer computed last is free again
    if (k0 + (NS - 1) * KP < K) issue(k0 + (NS - 1) * KP, stage + nxt * BUF);
    smrt_cp_async_commit();
    const double* buf = stage + cur * BUF;
    const int kn = (K - k0 < KP) ? (K - k0) : KP;
#pragma unroll 2
    for (int kk = 0; kk < kn; ++kk) {
      double a1[MU], a2[MU], bv[NV];
#pragma unroll
      for (int u = 0; u < MU; ++u) {
        a1[u] = buf[kk * SA + tx + 32 * u];
        a2[u] = kDual ? buf[KP * SA + kk * SA + tx + 32 * u] : 0.0;
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (kBShared) {
          const int j = n0 + ty + 16 * v;
          bv[v] = Bm[(size_t)(j < N ? j : N - 1) * ldb + k0 + kk];
        } else {
          bv[v] = buf[KP * SA * NA + kk * BW + ty + 16 * v];
        }
      }
#pragma unroll
      for (int u = 0; u < MU; ++u)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          c1[u][v] = fma(a1[u], bv[v], c1[u][v]);
          if (kDual) c2[u][v] = fma(a2[u], bv[v], c2[u][v]);  // compile-time condition
        }
    }
    cur = (cur + 1 == NS) ? 0 : cur + 1;
    nxt = (nxt + 1 == NS) ? 0 : nxt + 1;
  }
  smrt_cp_async_wait<0>();
  __syncthreads();  // the staging ring is free (the caller may overwrite it)
#pragma unroll
  for (int u = 0; u < MU; ++u) {
    const int i = tx + 32 * u;
    double fv[NV], gv[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int j = n0 + ty + 16 * v;
      fv[v] = gv[v] = 0.0;
      if (i < M && j < N) pre(i, j, fv[v], gv[v]);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int j = n0 + ty + 16 * v;
      if (i < M && j < N) epi(i, j, c1[u][v], kDual ? c2[u][v] : 0.0, fv[v], gv[v]);
    }
  }
}

// elementwise scaling of an r x r block in GLOBAL memory, A(i, k) *= rs[i] cs[k]: four independent elements per thread
// and trip (a plain read-modify-write loop pays the L2 latency once per element)
SMRT_DEV void scale_block_global(double* A, int ld, int r, const double* rs, const double* cs) {
  const int NT = blockDim.x, tid = threadIdx.x;
  const int total = r * r;
  for (int e0 = tid; e0 < total; e0 += 4 * NT) {
    double v[4];
    int idx[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = e0 + q * NT;
      const int k = e / r, i = e - k * r;
      idx[q] = (e < total) ? k * ld + i : -1;
      v[q] = (e < total) ? A[idx[q]] * (rs[i] * cs[k]) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (idx[q] >= 0) A[idx[q]] = v[q];
  }
}
