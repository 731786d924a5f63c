/*
 * oracle/hif_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product).
 *
 * Plain-C CPU restatement of the reference's preconditioner-application path,
 * working on the same plain-pointer level description the device backend is
 * attached with (LhfdGpuLevel, include/hifir_b200.h).  Each function cites the
 * reference lines it follows.  Loop orders are the reference's own (CCS column
 * sweeps), so that the only arithmetic difference to the reference build is the
 * compiler's -ffast-math freedom.
 *
 * PARITY PINNED: tests/test_oracle.py checks this file against
 *   (a) the UNMODIFIED reference compiled in oracle/_ref (HIF::solve, hifir,
 *       fgmres_hifir on the very same factorized object), and
 *   (b) golden vectors produced by that reference and committed in tests/golden/
 *       (script: tests/golden/make_golden.py), and
 *   (c) the reference's own known-answer test for the dense level
 *       (tests/test_sss_qrcp.cpp:16-197, 20x20 system, |x - x_ref| <= 1e-10).
 *
 * The dense last level lives in LAPACK in the reference (dormqr + dtrsv, any
 * LAPACK; OpenBLAS 0.3.15 in this image).  It is restated here with the
 * published unblocked algorithms (dorm2r: apply H_1 ... H_rk from the left,
 * transposed; dtrsv 'U','N','N': column-oriented back substitution).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hifir_b200.h"

/* CCS::solve_as_strict_lower -- CompressedStorage.hpp:2267-2279 */
static void ccs_solve_strict_lower(const LhfdGpuCcs *L, double *y) {
  const size_t n = L->ncols;
  if (!L->col_start) return;
  for (size_t j = 0; j < n; ++j) {
    const double yj = y[j];
    for (LhfIndPtr k = L->col_start[j]; k < L->col_start[j + 1]; ++k)
      y[L->row_ind[k]] -= L->vals[k] * yj;
  }
}

/* CCS::solve_as_strict_upper -- CompressedStorage.hpp:2356-2369
 * (columns n-1 ... 1, entries of a column in reverse order) */
static void ccs_solve_strict_upper(const LhfdGpuCcs *U, double *y) {
  const size_t n = U->ncols;
  if (!U->col_start || !n) return;
  for (size_t j = n - 1; j != 0; --j) {
    const double yj = y[j];
    for (LhfIndPtr k = U->col_start[j + 1]; k > U->col_start[j]; --k)
      y[U->row_ind[k - 1]] -= U->vals[k - 1] * yj;
  }
}

/* CCS::multiply_nt_low -- CompressedStorage.hpp:2078-2096 */
static void ccs_multiply(const LhfdGpuCcs *A, const double *x, double *y) {
  for (size_t i = 0; i < A->nrows; ++i) y[i] = 0.0;
  if (!A->col_start) return;
  for (size_t j = 0; j < A->ncols; ++j) {
    const double t = x[j];
    for (LhfIndPtr k = A->col_start[j]; k < A->col_start[j + 1]; ++k)
      y[A->row_ind[k]] += t * A->vals[k];
  }
}

/* internal::prec_solve_ldu -- prec_solve.hpp:204-223 */
static void solve_ldu(const LhfdGpuLevel *P, double *y) {
  if (!P->m) return;
  ccs_solve_strict_lower(&P->L_B, y);
  for (size_t i = 0; i < P->m; ++i) y[i] /= P->d_B[i];
  ccs_solve_strict_upper(&P->U_B, y);
}

/* QRCP::_solve_nt<1> -- small_scale/QRCP.hpp:370-411
 * mat: nm x nm column-major; tau; jpvt 1-based; rank: 0 -> numerical rank,
 * > nm -> nm (QRCP.hpp:376-377) */
void hif_oracle_qrcp_solve(size_t nm, size_t num_rank, const double *mat, const double *tau,
                           const LhfInt *jpvt, size_t rank, double *x) {
  const size_t rk = rank == 0 ? num_rank : (rank > nm ? nm : rank);
  /* x <- Q(:,1:rk)^T x : dormqr('L','T') == H_rk ... H_1 x applied H_1 first (dorm2r) */
  for (size_t k = 0; k < rk; ++k) {
    const double *v = mat + k * nm; /* v_k = [1; mat(k+1:nm, k)] */
    double        dot = x[k];
    for (size_t i = k + 1; i < nm; ++i) dot += v[i] * x[i];
    const double sigma = tau[k] * dot;
    x[k] -= sigma;
    for (size_t i = k + 1; i < nm; ++i) x[i] -= sigma * v[i];
  }
  /* dtrsv('U','N','N', rk): column-oriented back substitution (reference BLAS) */
  for (size_t jj = rk; jj > 0; --jj) {
    const size_t j = jj - 1;
    if (x[j] != 0.0) {
      x[j] /= mat[j + j * nm];
      const double t = x[j];
      for (size_t i = 0; i < j; ++i) x[i] -= t * mat[i + j * nm];
    }
  }
  /* work[jpvt[i]-1] = x[i] (i<rk), 0 for rk<=i<nm -- QRCP.hpp:401-409 */
  double *work = (double *)malloc(sizeof(double) * (nm ? nm : 1));
  for (size_t i = 0; i < rk; ++i) work[jpvt[i] - 1] = x[i];
  for (size_t i = rk; i < nm; ++i) work[jpvt[i] - 1] = 0.0;
  memcpy(x, work, sizeof(double) * nm);
  free(work);
}

/* prec_solve -- prec_solve.hpp:332-412 (recursive over levels) */
static void prec_solve(size_t nlevels, const LhfdGpuLevel *lv, size_t l, const double *b, size_t rank,
                       double *y, double *work) {
  const LhfdGpuLevel *P = lv + l;
  const size_t        m = P->m, n = P->n, nm = n - m;
  const int last = (P->dense_n != 0) || (m == n) || (l + 1 == nlevels); /* Prec.hpp:197-199 */
  for (size_t i = 0; i < m; ++i) work[i] = P->s[P->p[i]] * b[P->p[i]]; /* :359 */
  if (nm) {
    solve_ldu(P, work);                                                            /* :364 */
    ccs_multiply(&P->E, work, y + m);                                              /* :366 */
    for (size_t i = m; i < n; ++i) y[i] = P->s[P->p[i]] * b[P->p[i]] - y[i];       /* :368 */
  }
  if (last) {
    if (nm) /* :376-377 */
      hif_oracle_qrcp_solve(P->dense_n, P->dense_rank, P->qr_mat, P->qr_tau, P->qr_jpvt, rank, y + m);
  } else {
    memcpy(work + m, y + m, sizeof(double) * nm);                     /* :386 */
    prec_solve(nlevels, lv, l + 1, work + m, rank, y + m, work + n);  /* :388 */
  }
  memcpy(work + m, y + m, sizeof(double) * nm); /* :392 */
  if (P->F.ncols) {                             /* :395-399 */
    ccs_multiply(&P->F, y + m, work);
    for (size_t i = 0; i < m; ++i) work[i] = P->s[P->p[i]] * b[P->p[i]] - work[i];
  } else if (nm) {
    for (size_t i = 0; i < m; ++i) work[i] = P->s[P->p[i]] * b[P->p[i]];
  }
  solve_ldu(P, work);                                                   /* :406 */
  for (size_t i = 0; i < n; ++i) y[i] = P->t[i] * work[P->q_inv[i]];    /* :411 */
}

/* compute_prec_work_space -- prec_solve.hpp:620-626 */
size_t hif_oracle_work_size(size_t nlevels, const LhfdGpuLevel *lv) {
  size_t s = 0;
  for (size_t l = 0; l < nlevels; ++l) s += lv[l].n;
  return s;
}

/* NspFilter::_const_filter -- NspFilter.hpp:161-175 */
static void nsp_const_filter(double *x, size_t n, size_t start, size_t end) {
  if (end == (size_t)-1 || end < start) end = n;
  if (start == end) return;
  double sum = 0.0;
  for (size_t i = start; i < end; ++i) sum += x[i];
  const double shift = sum / (double)(end - start);
  for (size_t i = start; i < end; ++i) x[i] -= shift;
}

typedef struct HifOracleNsp {
  int    enabled;
  size_t start, end;
} HifOracleNsp;

/* HIF::solve -- builder.hpp:409-423 */
int hif_oracle_solve(size_t nlevels, const LhfdGpuLevel *lv, const double *b, size_t rank,
                     const HifOracleNsp *nsp, double *x) {
  if (!nlevels) return -1;
  double *work = (double *)malloc(sizeof(double) * (hif_oracle_work_size(nlevels, lv) + 1));
  if (!work) return -2;
  prec_solve(nlevels, lv, 0, b, rank, x, work);
  free(work);
  if (nsp && nsp->enabled) nsp_const_filter(x, lv[0].n, nsp->start, nsp->end);
  return 0;
}

/* ---- user matrix in CRS: y = A x  (CRS::multiply_nt_low, CompressedStorage.hpp:1108-1127) */
typedef struct HifOracleCrs {
  size_t           n;
  const LhfIndPtr *row_start;
  const LhfInt *   col_ind;
  const double *   vals;
} HifOracleCrs;

void hif_oracle_spmv(const HifOracleCrs *A, const double *x, double *y) {
  for (size_t i = 0; i < A->n; ++i) {
    double t = 0.0;
    for (LhfIndPtr k = A->row_start[i]; k < A->row_start[i + 1]; ++k) t += A->vals[k] * x[A->col_ind[k]];
    y[i] = t;
  }
}

/* norm2 -- utils/math.hpp:112-137 (max-scaled, two passes) */
double hif_oracle_norm2(const double *v, size_t n) {
  if (!n) return 0.0;
  double mx = 0.0;
  for (size_t i = 0; i < n; ++i)
    if (fabs(v[i]) > mx) mx = fabs(v[i]);
  double tmp = 0.0;
  if (mx == 0.0) {
    for (size_t i = 0; i < n; ++i) tmp += fabs(v[i]);
    return tmp;
  }
  const double alpha = 1.0 / mx;
  for (size_t i = 0; i < n; ++i) {
    const double a = v[i] * alpha;
    tmp += a * a;
  }
  return mx * sqrt(tmp);
}

/* IterRefine::iter_refine, fixed count -- IterRefine.hpp:77-105 */
int hif_oracle_hifir(size_t nlevels, const LhfdGpuLevel *lv, const HifOracleCrs *A, const double *b,
                     size_t N, size_t rank, const HifOracleNsp *nsp, double *x) {
  const size_t n = lv[0].n;
  if (N <= 1) return hif_oracle_solve(nlevels, lv, b, rank, nsp, x);
  double *xk = (double *)malloc(sizeof(double) * n), *r = (double *)malloc(sizeof(double) * n);
  for (size_t j = 0; j < n; ++j) x[j] = 0.0;
  for (size_t i = 0; i < N; ++i) {
    memcpy(xk, x, sizeof(double) * n);
    if (i) {
      hif_oracle_spmv(A, xk, x);
      for (size_t j = 0; j < n; ++j) x[j] = b[j] - x[j];
    } else
      memcpy(x, b, sizeof(double) * n);
    hif_oracle_solve(nlevels, lv, x, rank, nsp, r);
    for (size_t j = 0; j < n; ++j) x[j] = r[j] + xk[j];
  }
  free(xk);
  free(r);
  return 0;
}

/* IterRefine::iter_refine with residual bounds -- IterRefine.hpp:121-165
 * out[0] = iters, out[1] = flag */
int hif_oracle_hifir_betas(size_t nlevels, const LhfdGpuLevel *lv, const HifOracleCrs *A,
                           const double *b, size_t N, const double *betas, size_t rank,
                           const HifOracleNsp *nsp, double *x, long *out) {
  const size_t n = lv[0].n;
  if (N <= 1) {
    hif_oracle_solve(nlevels, lv, b, rank, nsp, x);
    out[0] = 1;
    out[1] = -1;
    return 0;
  }
  const double bnorm = hif_oracle_norm2(b, n);
  if (bnorm == 0.0) {
    for (size_t j = 0; j < n; ++j) x[j] = 0.0;
    out[0] = 0;
    out[1] = 0;
    return 0;
  }
  double *xk = (double *)malloc(sizeof(double) * n), *r = (double *)malloc(sizeof(double) * n);
  for (size_t j = 0; j < n; ++j) x[j] = 0.0;
  memcpy(r, b, sizeof(double) * n);
  size_t iters = 0;
  int    flag  = 0;
  for (;;) {
    hif_oracle_solve(nlevels, lv, r, rank, nsp, xk);
    for (size_t j = 0; j < n; ++j) x[j] += xk[j];
    if (++iters >= N) {
      flag = -1;
      break;
    }
    hif_oracle_spmv(A, x, xk);
    for (size_t j = 0; j < n; ++j) r[j] = b[j] - xk[j];
    const double res = hif_oracle_norm2(r, n) / bnorm;
    if (res <= betas[0]) break;
    if (res > betas[1]) {
      flag = 1;
      break;
    }
  }
  free(xk);
  free(r);
  out[0] = (long)iters;
  out[1] = flag;
  return 0;
}

/* fgmres_hifir / gmres_hif -- examples/advanced/gmres.hpp:126-230 / 18-122
 * flexible != 0: FGMRES with nirs = 2^outer refinements (full_rank -> rank -1 else 0);
 * flexible == 0: GMRES with the plain apply and x += M^{-1}(Q y).
 * out[0]=flag (0 ok, 1 stagnated, 2 diverged) out[1]=iters out[2]=num_mv */
int hif_oracle_krylov(size_t nlevels, const LhfdGpuLevel *lv, const HifOracleCrs *A, const double *b,
                      int flexible, int restart, double rtol, int maxit, int full_rank,
                      const HifOracleNsp *nsp, double *x, int *out) {
  const size_t n     = lv[0].n;
  const size_t rr    = full_rank ? (size_t)-1 : 0;
  int          iter = 0, flag = 0, num_mv = 0;
  const double beta0 = hif_oracle_norm2(b, n);
  for (size_t i = 0; i < n; ++i) x[i] = 0.0;
  out[0] = out[1] = out[2] = 0;
  if (beta0 == 0.0) return 0;
  double *v = (double *)malloc(sizeof(double) * n), *w = (double *)malloc(sizeof(double) * n);
  double *y  = (double *)calloc(restart + 1, sizeof(double));
  double *w2 = (double *)calloc(restart, sizeof(double));
  double *Q  = (double *)malloc(sizeof(double) * n * restart);
  double *Z  = flexible ? (double *)malloc(sizeof(double) * n * restart) : NULL;
  double *R  = (double *)calloc((size_t)restart * restart, sizeof(double));
  double *J  = (double *)calloc((size_t)restart * 2, sizeof(double));
  const int max_outer = (int)ceil((double)maxit / restart);
  double    resid     = 1.0;
  for (int it_outer = 0; it_outer < max_outer; ++it_outer) {
    if (iter) {
      hif_oracle_spmv(A, x, v);
      for (size_t i = 0; i < n; ++i) v[i] = b[i] - v[i];
    } else
      memcpy(v, b, sizeof(double) * n);
    const double beta = hif_oracle_norm2(v, n);
    y[0]              = beta;
    for (size_t i = 0; i < n; ++i) Q[i] = v[i] / beta;
    int          j    = 0;
    const size_t nirs = (size_t)1 << it_outer;
    for (;;) {
      memcpy(v, Q + (size_t)j * n, sizeof(double) * n);
      if (flexible) {
        hif_oracle_hifir(nlevels, lv, A, v, nirs, rr, nsp, w);
        num_mv += (int)nirs;
        memcpy(Z + (size_t)j * n, w, sizeof(double) * n);
      } else {
        hif_oracle_solve(nlevels, lv, v, rr, nsp, w);
        num_mv += 1;
      }
      hif_oracle_spmv(A, w, v);
      for (int k = 0; k <= j; ++k) {
        const double *qk = Q + (size_t)k * n;
        double        t  = 0.0;
        for (size_t i = 0; i < n; ++i) t += v[i] * qk[i];
        w2[k] = t;
        for (size_t i = 0; i < n; ++i) v[i] -= t * qk[i];
      }
      double v_norm2 = 0.0;
      for (size_t i = 0; i < n; ++i) v_norm2 += v[i] * v[i];
      const double v_norm = sqrt(v_norm2);
      if (j + 1 < restart)
        for (size_t i = 0; i < n; ++i) Q[(size_t)(j + 1) * n + i] = v[i] / v_norm;
      for (int c = 0; c + 1 <= j; ++c) {
        const double tmp = w2[c];
        w2[c]            = J[c] * tmp + J[restart + c] * w2[c + 1];
        w2[c + 1]        = -J[restart + c] * tmp + J[c] * w2[c + 1];
      }
      const double rho = sqrt(w2[j] * w2[j] + v_norm2);
      J[j]             = w2[j] / rho;
      J[restart + j]   = v_norm / rho;
      y[j + 1]         = -J[restart + j] * y[j];
      y[j]             = J[j] * y[j];
      w2[j]            = rho;
      for (int k = 0; k <= j; ++k) R[(size_t)j * restart + k] = w2[k];
      const double resid_prev = resid;
      resid                   = fabs(y[j + 1]) / beta0;
      if (resid >= resid_prev * (1.0 - 1e-8)) {
        flag = 1; /* STAGNATED */
        break;
      } else if (iter >= maxit) {
        flag = 2; /* DIVERGED */
        break;
      }
      ++iter;
      if (resid <= rtol || j + 1 >= restart) break;
      ++j;
    }
    for (int k = j; k > -1; --k) {
      y[k] /= R[(size_t)k * restart + k];
      const double tmp = y[k];
      for (int i = k - 1; i > -1; --i) y[i] -= tmp * R[(size_t)k * restart + i];
    }
    if (flexible) {
      for (int i = 0; i <= j; ++i) {
        const double tmp = y[i];
        for (size_t k = 0; k < n; ++k) x[k] += tmp * Z[(size_t)i * n + k];
      }
    } else { /* gmres.hpp:111-118: w = Q y ; x += M^{-1} w */
      for (size_t k = 0; k < n; ++k) w[k] = 0.0;
      for (int i = 0; i <= j; ++i) {
        const double tmp = y[i];
        for (size_t k = 0; k < n; ++k) w[k] += tmp * Q[(size_t)i * n + k];
      }
      hif_oracle_solve(nlevels, lv, w, rr, nsp, v);
      for (size_t k = 0; k < n; ++k) x[k] += v[k];
    }
    if (resid <= rtol || flag != 0) break;
  }
  free(v); free(w); free(y); free(w2); free(Q); free(Z); free(R); free(J);
  out[0] = flag;
  out[1] = iter;
  out[2] = num_mv;
  return 0;
}

/* CCS -> CRS conversion with ascending column order inside each row, the form
 * hif::CRS(const CCS&) produces (CompressedStorage.hpp:861-890); used by the tests to
 * check the device backend's index handling bit-exactly. */
void hif_oracle_ccs_to_crs(const LhfdGpuCcs *A, int64_t *row_start, LhfInt *col_ind, double *vals) {
  const size_t nr = A->nrows, nc = A->ncols;
  for (size_t i = 0; i <= nr; ++i) row_start[i] = 0;
  if (!A->col_start) return;
  const LhfIndPtr nnz = A->col_start[nc];
  for (LhfIndPtr k = 0; k < nnz; ++k) ++row_start[A->row_ind[k] + 1];
  for (size_t i = 0; i < nr; ++i) row_start[i + 1] += row_start[i];
  int64_t *next = (int64_t *)malloc(sizeof(int64_t) * (nr + 1));
  memcpy(next, row_start, sizeof(int64_t) * (nr + 1));
  for (size_t j = 0; j < nc; ++j)
    for (LhfIndPtr k = A->col_start[j]; k < A->col_start[j + 1]; ++k) {
      const int64_t pos = next[A->row_ind[k]]++;
      col_ind[pos]      = (LhfInt)j;
      vals[pos]         = A->vals[k];
    }
  free(next);
}

/* ======================================================================================
 * Transpose solve and multilevel products: LHF_SH, LHF_M, LHF_MH
 * ====================================================================================== */

/* CCS::multiply_t_low -- CompressedStorage.hpp:2161-2178 : y = A^T x, one dot product per column */
static void ccs_multiply_t(const LhfdGpuCcs *A, const double *x, double *y) {
  for (size_t j = 0; j < A->ncols; ++j) {
    double tmp = 0.0;
    if (A->col_start)
      for (LhfIndPtr k = A->col_start[j]; k < A->col_start[j + 1]; ++k) tmp += x[A->row_ind[k]] * A->vals[k];
    y[j] = tmp;
  }
}

/* CCS::solve_as_strict_upper_tran -- CompressedStorage.hpp:2399-2413 (forward, dot form) */
static void ccs_solve_strict_upper_tran(const LhfdGpuCcs *U, double *y) {
  if (!U->col_start) return;
  for (size_t j = 1; j < U->ncols; ++j) {
    double tmp = 0.0;
    for (LhfIndPtr k = U->col_start[j]; k < U->col_start[j + 1]; ++k) tmp += U->vals[k] * y[U->row_ind[k]];
    y[j] -= tmp;
  }
}

/* CCS::solve_as_strict_lower_tran -- CompressedStorage.hpp:2307-2322 (backward, dot form) */
static void ccs_solve_strict_lower_tran(const LhfdGpuCcs *L, double *y) {
  if (!L->col_start || !L->ncols) return;
  for (size_t j = L->ncols - 1; j != 0; --j) {
    const size_t j1 = j - 1;
    double       tmp = 0.0;
    for (LhfIndPtr k = L->col_start[j1]; k < L->col_start[j1 + 1]; ++k) tmp += L->vals[k] * y[L->row_ind[k]];
    y[j1] -= tmp;
  }
}

/* internal::prec_solve_utdlt -- prec_solve.hpp:284-303 */
static void solve_utdlt(const LhfdGpuLevel *P, double *y) {
  if (!P->m) return;
  ccs_solve_strict_upper_tran(&P->U_B, y);
  for (size_t i = 0; i < P->m; ++i) y[i] /= P->d_B[i];
  ccs_solve_strict_lower_tran(&P->L_B, y);
}

static size_t qr_rank(size_t nm, size_t num_rank, size_t rank) {
  return rank == 0 ? num_rank : (rank > nm ? nm : rank);
}

/* x <- Q(:,1:rk) x(1:rk): dormqr('L','N') with rk reflectors = H_1 ... H_rk x (apply H_rk first) */
static void qr_apply_q(size_t nm, size_t rk, const double *mat, const double *tau, double *x) {
  for (size_t kk = rk; kk > 0; --kk) {
    const size_t  k = kk - 1;
    const double *v = mat + k * nm;
    double        dot = x[k];
    for (size_t i = k + 1; i < nm; ++i) dot += v[i] * x[i];
    const double sigma = tau[k] * dot;
    x[k] -= sigma;
    for (size_t i = k + 1; i < nm; ++i) x[i] -= sigma * v[i];
  }
}
static void qr_apply_qt(size_t nm, size_t rk, const double *mat, const double *tau, double *x) {
  for (size_t k = 0; k < rk; ++k) {
    const double *v = mat + k * nm;
    double        dot = x[k];
    for (size_t i = k + 1; i < nm; ++i) dot += v[i] * x[i];
    const double sigma = tau[k] * dot;
    x[k] -= sigma;
    for (size_t i = k + 1; i < nm; ++i) x[i] -= sigma * v[i];
  }
}

/* QRCP::_solve_t<1> -- QRCP.hpp:417-453 */
void hif_oracle_qrcp_solve_t(size_t nm, size_t num_rank, const double *mat, const double *tau,
                             const LhfInt *jpvt, size_t rank, double *x) {
  const size_t rk = qr_rank(nm, num_rank, rank);
  double *     w  = (double *)calloc(nm ? nm : 1, sizeof(double));
  for (size_t i = 0; i < rk; ++i) w[i] = x[jpvt[i] - 1];
  /* dtrsv('U','C','N'): forward substitution with R^T, dot form (reference BLAS) */
  for (size_t j = 0; j < rk; ++j) {
    double temp = w[j];
    for (size_t i = 0; i < j; ++i) temp -= mat[i + j * nm] * w[i];
    w[j] = temp / mat[j + j * nm];
  }
  qr_apply_q(nm, rk, mat, tau, w);
  memcpy(x, w, sizeof(double) * nm);
  free(w);
}

/* QRCP::_multiply_nt<1> -- QRCP.hpp:459-492 : x <- Q R(1:rk,1:rk) (P^T x)(1:rk) */
void hif_oracle_qrcp_multiply(size_t nm, size_t num_rank, const double *mat, const double *tau,
                              const LhfInt *jpvt, size_t rank, double *x) {
  const size_t rk = qr_rank(nm, num_rank, rank);
  double *     w  = (double *)calloc(nm ? nm : 1, sizeof(double));
  for (size_t i = 0; i < rk; ++i) w[i] = x[jpvt[i] - 1];
  /* dtrmv('U','N','N'), reference BLAS column form */
  for (size_t j = 0; j < rk; ++j) {
    if (w[j] != 0.0) {
      const double temp = w[j];
      for (size_t i = 0; i < j; ++i) w[i] += temp * mat[i + j * nm];
      w[j] *= mat[j + j * nm];
    }
  }
  qr_apply_q(nm, rk, mat, tau, w);
  memcpy(x, w, sizeof(double) * nm);
  free(w);
}

/* QRCP::_multiply_t<1> -- QRCP.hpp:494-540 : x <- P [R^T (Q^T x)(1:rk) ; 0] */
void hif_oracle_qrcp_multiply_t(size_t nm, size_t num_rank, const double *mat, const double *tau,
                                const LhfInt *jpvt, size_t rank, double *x) {
  const size_t rk = qr_rank(nm, num_rank, rank);
  qr_apply_qt(nm, rk, mat, tau, x);
  /* dtrmv('U','C','N'), reference BLAS: j descending, dot with rows i < j (descending) */
  for (size_t jj = rk; jj > 0; --jj) {
    const size_t j = jj - 1;
    double       temp = x[j] * mat[j + j * nm];
    for (size_t ii = j; ii > 0; --ii) temp += mat[(ii - 1) + j * nm] * x[ii - 1];
    x[j] = temp;
  }
  double *w = (double *)calloc(nm ? nm : 1, sizeof(double));
  for (size_t i = 0; i < rk; ++i) w[jpvt[i] - 1] = x[i];
  memcpy(x, w, sizeof(double) * nm);
  free(w);
}

/* prec_solve_tran -- prec_solve.hpp:541-612 */
static void prec_solve_tran(size_t nlevels, const LhfdGpuLevel *lv, size_t l, const double *b, size_t rank,
                            double *y, double *work) {
  const LhfdGpuLevel *P = lv + l;
  const size_t        m = P->m, n = P->n, nm = n - m;
  const int last = (P->dense_n != 0) || (m == n) || (l + 1 == nlevels);
  for (size_t i = 0; i < m; ++i) work[i] = P->t[P->q[i]] * b[P->q[i]];
  if (P->F.ncols) {
    solve_utdlt(P, work);
    ccs_multiply_t(&P->F, work, y + m);
    for (size_t i = m; i < n; ++i) y[i] = P->t[P->q[i]] * b[P->q[i]] - y[i];
  } else if (nm)
    for (size_t i = m; i < n; ++i) y[i] = P->t[P->q[i]] * b[P->q[i]];
  if (last) {
    if (nm) hif_oracle_qrcp_solve_t(P->dense_n, P->dense_rank, P->qr_mat, P->qr_tau, P->qr_jpvt, rank, y + m);
  } else {
    memcpy(work + m, y + m, sizeof(double) * nm);
    prec_solve_tran(nlevels, lv, l + 1, work + m, rank, y + m, work + n);
  }
  memcpy(work + m, y + m, sizeof(double) * nm);
  if (nm) {
    ccs_multiply_t(&P->E, y + m, work);
    for (size_t i = 0; i < m; ++i) work[i] = P->t[P->q[i]] * b[P->q[i]] - work[i];
  } else
    for (size_t i = 0; i < m; ++i) work[i] = P->t[P->q[i]] * b[P->q[i]];
  solve_utdlt(P, work);
  for (size_t i = 0; i < n; ++i) y[i] = P->s[i] * work[P->p_inv[i]];
}

/* prec_prod -- prec_prod.hpp:54-134 */
static void prec_prod(size_t nlevels, const LhfdGpuLevel *lv, size_t l, const double *b, size_t rank, double *y,
                      double *work) {
  const LhfdGpuLevel *P = lv + l;
  const size_t        m = P->m, n = P->n, nm = n - m;
  const int last = (P->dense_n != 0) || (m == n) || (l + 1 == nlevels);
  for (size_t i = m; i < n; ++i) work[i] = b[P->q[i]] / P->t[P->q[i]];
  if (last) {
    if (nm) {
      memcpy(y + m, work + m, sizeof(double) * nm);
      hif_oracle_qrcp_multiply(P->dense_n, P->dense_rank, P->qr_mat, P->qr_tau, P->qr_jpvt, rank, y + m);
    }
  } else
    prec_prod(nlevels, lv, l + 1, work + m, rank, y + m, work + n);
  for (size_t i = 0; i < m; ++i) work[i] = b[P->q[i]] / P->t[P->q[i]];
  ccs_multiply(&P->U_B, work, y);
  for (size_t i = 0; i < m; ++i) y[i] = (y[i] + work[i]) * P->d_B[i];
  ccs_multiply(&P->L_B, y, work);
  for (size_t i = 0; i < m; ++i) work[i] += y[i];
  if (P->F.ncols) {
    ccs_multiply(&P->F, work + m, y);
    for (size_t i = 0; i < m; ++i) work[i] += y[i];
  }
  if (nm) {
    solve_ldu(P, y);
    for (size_t i = 0; i < m; ++i) y[i] += b[P->q[i]] / P->t[P->q[i]];
    ccs_multiply(&P->E, y, work + m);
    for (size_t i = m; i < n; ++i) work[i] += y[i];
  }
  for (size_t i = 0; i < n; ++i) y[i] = work[P->p_inv[i]] / P->s[i];
}

/* prec_prod_tran -- prec_prod.hpp:148-230 */
static void prec_prod_tran(size_t nlevels, const LhfdGpuLevel *lv, size_t l, const double *b, size_t rank,
                           double *y, double *work) {
  const LhfdGpuLevel *P = lv + l;
  const size_t        m = P->m, n = P->n, nm = n - m;
  const int last = (P->dense_n != 0) || (m == n) || (l + 1 == nlevels);
  for (size_t i = m; i < n; ++i) work[i] = b[P->p[i]] / P->s[P->p[i]];
  if (last) {
    if (nm) {
      memcpy(y + m, work + m, sizeof(double) * nm);
      hif_oracle_qrcp_multiply_t(P->dense_n, P->dense_rank, P->qr_mat, P->qr_tau, P->qr_jpvt, rank, y + m);
    }
  } else
    prec_prod_tran(nlevels, lv, l + 1, work + m, rank, y + m, work + n);
  for (size_t i = 0; i < m; ++i) work[i] = b[P->p[i]] / P->s[P->p[i]];
  ccs_multiply_t(&P->L_B, work, y);
  for (size_t i = 0; i < m; ++i) y[i] = (y[i] + work[i]) * P->d_B[i];
  ccs_multiply_t(&P->U_B, y, work);
  for (size_t i = 0; i < m; ++i) work[i] += y[i];
  if (nm) {
    ccs_multiply_t(&P->E, work + m, y);
    for (size_t i = 0; i < m; ++i) work[i] += y[i];
  }
  if (nm) {
    solve_utdlt(P, y);
    for (size_t i = 0; i < m; ++i) y[i] += b[P->p[i]] / P->s[P->p[i]];
    ccs_multiply_t(&P->F, y, work + m);
    for (size_t i = m; i < n; ++i) work[i] += y[i];
  }
  for (size_t i = 0; i < n; ++i) y[i] = work[P->q_inv[i]] / P->t[i];
}

/* HIF::solve(trans=true) / HIF::mmultiply -- builder.hpp:409-423, 502-512
 * op: 1 = S^H, 2 = M, 3 = M^H (LhfOperationType values) */
int hif_oracle_apply_op(size_t nlevels, const LhfdGpuLevel *lv, int op, const double *b, size_t rank,
                        double *x) {
  if (!nlevels) return -1;
  double *work = (double *)malloc(sizeof(double) * (hif_oracle_work_size(nlevels, lv) + 1));
  if (!work) return -2;
  if (op == 1)
    prec_solve_tran(nlevels, lv, 0, b, rank, x, work);
  else if (op == 2)
    prec_prod(nlevels, lv, 0, b, rank, x, work);
  else if (op == 3)
    prec_prod_tran(nlevels, lv, 0, b, rank, x, work);
  else {
    free(work);
    return -3;
  }
  free(work);
  return 0;
}
