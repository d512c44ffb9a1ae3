// dort_linalg.cuh — CTA-cooperative dense fp64 linear algebra on small column-major matrices held in shared memory
// (or, for stream counts whose blocks exceed 227 KB, in an L2-resident global scratch): register-tiled GEMM with
// functor operands, Cholesky, one-sided Jacobi SVD, LU with partial pivoting and the triangular solves the DORT path
// needs.  All routines are called by every thread of the block (or of a half-block "team" for the two concurrent
// Cholesky factorisations) and synchronise internally.
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
// In-place lower Cholesky A = L L^T of the h x h symmetric positive definite matrix whose LOWER triangle is stored in A
// (column-major, leading dimension ld).  The strict upper triangle is neither read nor written.
// Returns (to every thread of the team) 0 on success, 1 if a pivot is not positive (matrix not SPD).
// `flag` is a team-shared int scratch.
SMRT_DEV int team_cholesky(const Team& tm, double* A, int ld, int h, int* flag) {
  if (tm.rank == 0) *flag = 0;
  tm.sync();
  for (int j = 0; j < h; ++j) {
    // column j: every thread reads the pivot, scales its rows
    double d = SMRT_AT(A, ld, j, j);
    if (!(d > 0.0)) {
      if (tm.rank == 0) *flag = 1;
      tm.sync();
      return 1;
    }
    double piv = sqrt(d);
    double inv = 1.0 / piv;
    tm.sync();  // everybody has read A(j,j) before it is overwritten
    for (int i = j + tm.rank; i < h; i += tm.size) {
      double v = SMRT_AT(A, ld, i, j);
      SMRT_AT(A, ld, i, j) = (i == j) ? piv : v * inv;
    }
    tm.sync();
    // trailing update of the lower triangle: A(i, c) -= L(i, j) L(c, j) for j < c <= i < h
    int nt = h - j - 1;  // trailing size
    if (nt > 0) {
      // enumerate (c, i) with c in (j, h), i in [c, h): work split by columns-of-rows pairs
      for (int c = j + 1 + (tm.rank / 32); c < h; c += tm.size / 32) {
        double lc = SMRT_AT(A, ld, c, j);
        for (int i = c + (tm.rank & 31); i < h; i += 32) {
          SMRT_AT(A, ld, i, c) = fma(-SMRT_AT(A, ld, i, j), lc, SMRT_AT(A, ld, i, c));
        }
      }
    }
    tm.sync();
  }
  return *flag;
}

// ------------------------------------------------------------------------------------------------- one-sided Jacobi SVD
// Orthogonalises the columns of the h x h matrix W (column-major, ld) in place by plane rotations applied from the
// right (Hestenes): on exit W = M V with V orthogonal and mutually orthogonal columns, |w_j| = sigma_j.
// Round-robin ("circle") ordering: hp = h rounded up to even, hp - 1 rounds per sweep, hp / 2 disjoint column pairs per
// round, TPP threads per pair (power of two <= 32).  `ctrl` is a block-shared int[4] scratch.
// Returns the number of sweeps performed (every thread gets the same value).
#define SMRT_JACOBI_TOL 1e-15        // rotate only when |w_p . w_q| > tol |w_p| |w_q|
#define SMRT_JACOBI_DONE 1e-13       // a sweep whose largest cosine is below this ends the iteration
#define SMRT_JACOBI_QUAD 1e-9        // ... or below this before its own rotations (quadratic convergence finishes it)
#define SMRT_JACOBI_MAX_SWEEPS 40

SMRT_DEV int block_jacobi_svd(double* W, int ld, int h, int* ctrl) {
  const int NT = blockDim.x;
  const int tid = threadIdx.x;
  const int hp = (h + 1) & ~1;
  const int npairs = hp / 2;
  int tpp = 32;
  while (tpp > 1 && npairs * tpp > NT) tpp >>= 1;
  // if there are more pairs than threads (h > 2 NT) each group loops over several pairs
  const int ngroups = NT / tpp;
  const int grp = tid / tpp, lane = tid % tpp;
  // lanes of this thread's group inside its warp (shuffles are issued per group, groups may diverge)
  const unsigned gmask = (tpp == 32) ? 0xffffffffu : (((1u << tpp) - 1u) << ((tid & 31) & ~(tpp - 1)));
  int sweeps = 0;
  for (; sweeps < SMRT_JACOBI_MAX_SWEEPS; ++sweeps) {
    if (tid == 0) ctrl[0] = 0;  // max cosine of the sweep, as ordered int bits of a non-negative double's high word
    __syncthreads();
    double maxcos = 0.0;
    for (int r = 0; r < hp - 1; ++r) {
      for (int pi = grp; pi < npairs; pi += ngroups) {
        int p, q;
        if (pi == 0) {
          p = r;
          q = hp - 1;
        } else {
          p = (r + pi) % (hp - 1);
          q = (r - pi + (hp - 1)) % (hp - 1);
        }
        if (p > q) {
          int t = p;
          p = q;
          q = t;
        }
        if (q < h) {  // skip the padding column of an odd-sized problem
          double* wp = W + (size_t)p * ld;
          double* wq = W + (size_t)q * ld;
          double a = 0.0, b = 0.0, g = 0.0;
          for (int i = lane; i < h; i += tpp) {
            double x = wp[i], y = wq[i];
            a = fma(x, x, a);
            b = fma(y, y, b);
            g = fma(x, y, g);
          }
          for (int off = tpp >> 1; off > 0; off >>= 1) {
            a += __shfl_xor_sync(gmask, a, off, 32);
            b += __shfl_xor_sync(gmask, b, off, 32);
            g += __shfl_xor_sync(gmask, g, off, 32);
          }
          double denom = sqrt(a * b);
          double cosv = (denom > 0.0) ? fabs(g) / denom : 0.0;
          maxcos = fmax(maxcos, cosv);
          if (cosv > SMRT_JACOBI_TOL) {
            double zeta = (b - a) / (2.0 * g);
            double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            double c = 1.0 / sqrt(1.0 + t * t);
            double s = c * t;
            for (int i = lane; i < h; i += tpp) {
              double x = wp[i], y = wq[i];
              wp[i] = c * x - s * y;
              wq[i] = s * x + c * y;
            }
          }
        }
      }
      __syncthreads();
    }
    // block-wide max of maxcos (non-negative doubles order like their bit patterns: compare the high 32 bits + 1)
    {
      unsigned long long bits;
      memcpy(&bits, &maxcos, 8);
      int hi = (int)(bits >> 33);  // drop the sign bit position, keep 31 bits: monotone for non-negative values
      atomicMax(&ctrl[0], hi);
    }
    __syncthreads();
    int hi = ctrl[0];
    __syncthreads();
    unsigned long long bits = ((unsigned long long)(unsigned)hi) << 33;
    double mc;
    memcpy(&mc, &bits, 8);  // lower bound of the max cosine (truncated mantissa)
    if (mc < SMRT_JACOBI_QUAD) {
      ++sweeps;
      break;
    }
  }
  return sweeps;
}

// ------------------------------------------------------------------------------------------ LU with partial pivoting
// In-place LU of the h x h matrix A (column-major, ld) with row interchanges applied physically; perm[j] = pivot row
// chosen at step j (LAPACK ipiv convention, 0-based).  Returns 0, or 1 if a pivot is exactly zero / not finite.
// `ctrl` is block-shared int[4].
SMRT_DEV int block_lu(double* A, int ld, int h, int* perm, int* ctrl) {
  const int NT = blockDim.x;
  const int tid = threadIdx.x;
  if (tid == 0) ctrl[1] = 0;
  __syncthreads();
  for (int j = 0; j < h; ++j) {
    // pivot search by warp 0
    if (tid < 32) {
      double best = -1.0;
      int bi = j;
      for (int i = j + tid; i < h; i += 32) {
        double v = fabs(SMRT_AT(A, ld, i, j));
        if (v > best) {
          best = v;
          bi = i;
        }
      }
      for (int off = 16; off > 0; off >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, off, 32);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off, 32);
        if (ob > best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      if (tid == 0) {
        perm[j] = bi;
        if (!(best > 0.0) || !isfinite(best)) ctrl[1] = 1;
      }
    }
    __syncthreads();
    if (ctrl[1]) return 1;
    int pr = perm[j];
    // swap rows j and pr across all columns
    if (pr != j) {
      for (int c = tid; c < h; c += NT) {
        double t = SMRT_AT(A, ld, j, c);
        SMRT_AT(A, ld, j, c) = SMRT_AT(A, ld, pr, c);
        SMRT_AT(A, ld, pr, c) = t;
      }
      __syncthreads();
    }
    double inv = 1.0 / SMRT_AT(A, ld, j, j);
    __syncthreads();
    // scale the column and update the trailing matrix: A(i, c) -= l_i * A(j, c)
    // thread layout: 32 lanes along rows, warps along columns
    int nrows = h - j - 1;
    if (nrows > 0) {
      for (int c = j + 1 + (tid >> 5); c < h; c += (NT >> 5)) {
        double ujc = SMRT_AT(A, ld, j, c);
        for (int i = j + 1 + (tid & 31); i < h; i += 32) {
          double lij = SMRT_AT(A, ld, i, j) * inv;
          SMRT_AT(A, ld, i, c) = fma(-lij, ujc, SMRT_AT(A, ld, i, c));
        }
      }
      __syncthreads();
      for (int i = j + 1 + tid; i < h; i += NT) SMRT_AT(A, ld, i, j) *= inv;
      __syncthreads();
    }
  }
  return 0;
}

// Solve op(A) X = B in place for the h x nrhs block B (column-major, ldb), A = P^T L U from block_lu.
//   transposed == 0:  A X = B    (row interchanges, forward with unit-lower L, backward with U)
//   transposed == 1:  A^T X = B  (forward with U^T, backward with unit-upper L^T, interchanges undone in reverse)
// Columns are independent: `tpc` threads cooperate on one column through the axpy form of the substitutions.
SMRT_DEV void block_lu_solve(const double* LU, int ld, int h, const int* perm, double* Bm, int ldb, int nrhs,
                             int transposed) {
  const int NT = blockDim.x;
  const int tid = threadIdx.x;
  int tpc = 32;
  while (tpc > 1 && nrhs * tpc > NT) tpc >>= 1;
  const int ngroups = NT / tpc;
  const int grp = tid / tpc, lane = tid % tpc;
  const unsigned gmask = (tpc == 32) ? 0xffffffffu : (((1u << tpc) - 1u) << ((tid & 31) & ~(tpc - 1)));
  for (int c = grp; c < nrhs; c += ngroups) {
    double* x = Bm + (size_t)c * ldb;
    if (!transposed) {
      if (lane == 0) {
        for (int j = 0; j < h; ++j) {
          int pr = perm[j];
          if (pr != j) {
            double t = x[j];
            x[j] = x[pr];
            x[pr] = t;
          }
        }
      }
      __syncwarp(gmask);
      for (int j = 0; j < h; ++j) {  // forward: x_i -= L(i, j) x_j, i > j
        double xj = x[j];
        for (int i = j + 1 + lane; i < h; i += tpc) x[i] = fma(-SMRT_AT(LU, ld, i, j), xj, x[i]);
        __syncwarp(gmask);
      }
      for (int j = h - 1; j >= 0; --j) {  // backward: x_j /= U(j, j); x_i -= U(i, j) x_j, i < j
        if (lane == 0) x[j] = x[j] / SMRT_AT(LU, ld, j, j);
        __syncwarp(gmask);
        double xj = x[j];
        for (int i = lane; i < j; i += tpc) x[i] = fma(-SMRT_AT(LU, ld, i, j), xj, x[i]);
        __syncwarp(gmask);
      }
    } else {
      for (int j = 0; j < h; ++j) {  // U^T w = b: w_j = b_j / U(j, j); b_i -= U(j, i) w_j, i > j
        if (lane == 0) x[j] = x[j] / SMRT_AT(LU, ld, j, j);
        __syncwarp(gmask);
        double xj = x[j];
        for (int i = j + 1 + lane; i < h; i += tpc) x[i] = fma(-SMRT_AT(LU, ld, j, i), xj, x[i]);
        __syncwarp(gmask);
      }
      for (int j = h - 1; j >= 0; --j) {  // L^T z = w: z_j = w_j; w_i -= L(j, i) z_j, i < j
        double xj = x[j];
        for (int i = lane; i < j; i += tpc) x[i] = fma(-SMRT_AT(LU, ld, j, i), xj, x[i]);
        __syncwarp(gmask);
      }
      if (lane == 0) {
        for (int j = h - 1; j >= 0; --j) {
          int pr = perm[j];
          if (pr != j) {
            double t = x[j];
            x[j] = x[pr];
            x[pr] = t;
          }
        }
      }
      __syncwarp(gmask);
    }
  }
  __syncthreads();
}
