// oracle/ref_bridge.cpp -- TEST INFRASTRUCTURE / FACTOR PRODUCER, not product code.
//
// Compiles the UNMODIFIED reference (hifirworks/hifir v0.2.0, header-only C++)
// from where it lies under /root/reference/src into oracle/_ref/libhifir_ref.so
// and exposes it through a flat C interface so that Python (ctypes) tests and
// bench.py can
//   (1) run the reference's host-side Crout factorization (north_star keeps the
//       factorization on the host: builder.hpp:263-366),
//   (2) read the per-level factors exactly as they sit in hif::Prec
//       (alg/Prec.hpp:309-323) -- the upload contract of the device backend,
//   (3) run the reference's own CPU M^-1 apply / hifir / FGMRES as the parity
//       oracle and CPU baseline (builder.hpp:409-489, IterRefine.hpp:77-165,
//       examples/advanced/gmres.hpp:18-230).
// No reference source text is copied; only its public API is called.
//
// Build: see oracle/Makefile (g++ -std=c++11 -O3 -ffast-math -DNDEBUG -fopenmp,
// the reference's own release flags, Makefile.in:26).

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#define HIF_THROW  // errors become std::runtime_error (log.hpp:110-131)
#include "hifir.hpp"

#include "advanced/gmres.hpp"  // examples/advanced/gmres.hpp (fgmres_hifir, gmres_hif)

// the C++ drop-in adapter of the product (header-only, plain pointers out)
#include "hifir_b200.hpp"

namespace {

using crs_t   = hif::CRS<double, int>;
using array_t = hif::Array<double>;

template <class V>
struct QrAccT : hif::QRCP<V> {
  using hif::QRCP<V>::_tau;
  using hif::QRCP<V>::_jpvt;
};

// V = value type of the factors: double (hifref_*) or float (hifrefs_*: hif::HIF<float,int>, the
// mixed-precision set-up of examples/intermediate/demo_mixedprecision.cpp -- the user matrix and
// all vectors stay double, exactly as lhfsdApply uses it, libhifir.cpp:1192-1218)
template <class V>
struct RefHandleT {
  typedef hif::HIF<V, int>             prec_t;
  typedef typename prec_t::prec_type   level_t;
  prec_t M;
  // the user matrix (CRS), wrapped around caller-kept arrays copied here
  std::vector<std::ptrdiff_t> indptr;
  std::vector<int>            indices;
  std::vector<double>         vals;
  crs_t                       A;
  std::size_t                 n;
};
using RefHandle = RefHandleT<double>;
using prec_t    = RefHandle::prec_t;
using level_t   = RefHandle::level_t;
using QrAcc     = QrAccT<double>;

thread_local std::string g_err;

template <class V>
const typename RefHandleT<V>::level_t &level_at(const RefHandleT<V> *h, int lvl) {
  auto it = h->M.precs().cbegin();
  std::advance(it, lvl);
  return *it;
}

}  // namespace

#define REF_TRY try {
#define REF_CATCH                  \
  }                                \
  catch (const std::exception &e) { \
    g_err = e.what();              \
    return -1;                     \
  }                                \
  return 0;

extern "C" {

const char *hifref_last_error() { return g_err.c_str(); }

const char *hifref_version() {
  static std::string v = hif::version();
  return v.c_str();
}

// params[0..7] = tau_L, tau_U, kappa_d, kappa, alpha_L, alpha_U, dense_thres(<=0: keep), threads
// any entry < 0 keeps the reference default (Options.h:135-164).
void *hifref_create(std::size_t n, const std::int64_t *indptr, const int *indices,
                    const double *vals, const double *params, int verbose) {
  try {
    std::unique_ptr<RefHandle> h(new RefHandle());
    h->n = n;
    h->indptr.assign(indptr, indptr + n + 1);
    const std::size_t nnz = (std::size_t)indptr[n];
    h->indices.assign(indices, indices + nnz);
    h->vals.assign(vals, vals + nnz);
    h->A = crs_t(n, n, h->indptr.data(), h->indices.data(), h->vals.data(), true);
    auto opts = hif::get_default_options();
    if (params) {
      if (params[0] >= 0) opts.tau_L = params[0];
      if (params[1] >= 0) opts.tau_U = params[1];
      if (params[2] >= 0) opts.kappa_d = params[2];
      if (params[3] >= 0) opts.kappa = params[3];
      if (params[4] >= 0) opts.alpha_L = params[4];
      if (params[5] >= 0) opts.alpha_U = params[5];
      if (params[6] > 0) opts.dense_thres = (int)params[6];
      if (params[7] > 0) opts.threads = (int)params[7];
    }
    opts.verbose = verbose ? hif::VERBOSE_INFO : hif::VERBOSE_NONE;
    h->M.factorize(h->A, opts);
    return h.release();
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

void hifref_destroy(void *hdl) { delete static_cast<RefHandle *>(hdl); }

// number of hif::Prec nodes (the dense block lives inside the last one)
int hifref_num_precs(const void *hdl) {
  return (int)static_cast<const RefHandle *>(hdl)->M.precs().size();
}
std::size_t hifref_levels(const void *hdl) {
  return static_cast<const RefHandle *>(hdl)->M.levels();
}
std::size_t hifref_nnz(const void *hdl) {
  return static_cast<const RefHandle *>(hdl)->M.nnz();
}

// sizes[0..9] = m, n, nnz(L_B), nnz(U_B), nnz(E), nnz(F), dense_nrows, dense_rank,
//               has_symm_dense, is_last_level
int hifref_level_sizes(const void *hdl, int lvl, std::size_t *sizes) {
  REF_TRY
  const auto &P = level_at(static_cast<const RefHandle *>(hdl), lvl);
  sizes[0]      = P.m;
  sizes[1]      = P.n;
  sizes[2]      = P.L_B.nnz();
  sizes[3]      = P.U_B.nnz();
  sizes[4]      = P.E.nnz();
  sizes[5]      = P.F.nnz();
  sizes[6]      = P.dense_solver.empty() ? 0 : P.dense_solver.mat().nrows();
  sizes[7]      = P.dense_solver.empty() ? 0 : P.dense_solver.rank();
  sizes[8]      = !P.symm_dense_solver.empty();
  sizes[9]      = P.is_last_level();
  REF_CATCH
}

// which: 0=L_B 1=U_B 2=E 3=F ; copies the native CCS arrays verbatim
int hifref_export_ccs(const void *hdl, int lvl, int which, std::int64_t *col_start,
                      int *row_ind, double *vals) {
  REF_TRY
  const auto &P = level_at(static_cast<const RefHandle *>(hdl), lvl);
  const level_t::mat_type *M =
      which == 0 ? &P.L_B : which == 1 ? &P.U_B : which == 2 ? &P.E : &P.F;
  const std::size_t nc = M->ncols();
  if (M->col_start().size() >= nc + 1)
    for (std::size_t i = 0; i <= nc; ++i) col_start[i] = M->col_start()[i];
  else
    for (std::size_t i = 0; i <= nc; ++i) col_start[i] = 0;
  const std::size_t nnz = M->nnz();
  std::copy_n(M->row_ind().cbegin(), nnz, row_ind);
  std::copy_n(M->vals().cbegin(), nnz, vals);
  REF_CATCH
}

// shape of one of the four blocks: dims[0]=nrows dims[1]=ncols
int hifref_block_shape(const void *hdl, int lvl, int which, std::size_t *dims) {
  REF_TRY
  const auto &P = level_at(static_cast<const RefHandle *>(hdl), lvl);
  const level_t::mat_type *M =
      which == 0 ? &P.L_B : which == 1 ? &P.U_B : which == 2 ? &P.E : &P.F;
  dims[0] = M->nrows();
  dims[1] = M->ncols();
  REF_CATCH
}

int hifref_export_vectors(const void *hdl, int lvl, double *d, double *s, double *t,
                          int *p, int *p_inv, int *q, int *q_inv) {
  REF_TRY
  const auto &P = level_at(static_cast<const RefHandle *>(hdl), lvl);
  std::copy_n(P.d_B.cbegin(), P.m, d);
  std::copy_n(P.s.cbegin(), P.n, s);
  std::copy_n(P.t.cbegin(), P.n, t);
  std::copy_n(P.p.cbegin(), P.n, p);
  std::copy_n(P.p_inv.cbegin(), P.n, p_inv);
  std::copy_n(P.q.cbegin(), P.n, q);
  std::copy_n(P.q_inv.cbegin(), P.n, q_inv);
  REF_CATCH
}

// QRCP state of the last level: mat (nm x nm col-major, R on/above diag, reflectors
// below), tau[nm], jpvt[nm] (1-based, verbatim) -- QRCP.hpp:544-555
int hifref_export_dense(const void *hdl, int lvl, double *mat, double *tau, int *jpvt) {
  REF_TRY
  const auto &P  = level_at(static_cast<const RefHandle *>(hdl), lvl);
  const auto &qr = static_cast<const QrAcc &>(P.dense_solver);
  const std::size_t nm = qr.mat().nrows();
  std::copy_n(qr.mat().data(), nm * qr.mat().ncols(), mat);
  std::copy_n(qr._tau.cbegin(), qr._tau.size(), tau);
  for (std::size_t i = 0; i < qr._jpvt.size(); ++i) jpvt[i] = (int)qr._jpvt[i];
  REF_CATCH
}

// NspFilter in constant mode (NspFilter.hpp:161-175); end = size_t(-1) -> n
int hifref_set_nsp_const(void *hdl, std::size_t start, std::size_t end) {
  REF_TRY
  auto *h  = static_cast<RefHandle *>(hdl);
  h->M.nsp = hif::create_nsp_filter(start, end);
  REF_CATCH
}
int hifref_clear_nsp(void *hdl) {
  static_cast<RefHandle *>(hdl)->M.nsp.reset();
  return 0;
}

// x = M^{-1} b   (HIF::solve, builder.hpp:409-423); rank: 0 = numerical rank
int hifref_solve(const void *hdl, const double *b, double *x, std::size_t rank) {
  REF_TRY
  const auto *h = static_cast<const RefHandle *>(hdl);
  const array_t bb(h->n, const_cast<double *>(b), true);
  array_t       xx(h->n, x, true);
  h->M.solve(bb, xx, false, rank);
  REF_CATCH
}

// y = M x (HIF::mmultiply, builder.hpp:502-512) -- used for the libhifir
// round-trip invariant (libhifir/tests/test_real.c:146)
int hifref_mmultiply(const void *hdl, const double *x, double *y, std::size_t rank) {
  REF_TRY
  const auto *h = static_cast<const RefHandle *>(hdl);
  const array_t xx(h->n, const_cast<double *>(x), true);
  array_t       yy(h->n, y, true);
  h->M.mmultiply(xx, yy, false, rank);
  REF_CATCH
}

// the remaining lhf?Apply operations (libhifir.cpp:447-472): op 1 = S^H (solve, trans),
// 2 = M (mmultiply), 3 = M^H (mmultiply, trans)   -- builder.hpp:409-423, 502-512
int hifref_apply_op(const void *hdl, int op, const double *b, double *x, std::size_t rank) {
  REF_TRY
  const auto *h = static_cast<const RefHandle *>(hdl);
  const array_t bb(h->n, const_cast<double *>(b), true);
  array_t       xx(h->n, x, true);
  if (op == 1)
    h->M.solve(bb, xx, true, rank);
  else if (op == 2)
    h->M.mmultiply(bb, xx, false, rank);
  else if (op == 3)
    h->M.mmultiply(bb, xx, true, rank);
  else
    throw std::invalid_argument("bad op");
  REF_CATCH
}

// HIF::hifir fixed-count variant (builder.hpp:458-465), A = the matrix given at create
int hifref_hifir(const void *hdl, const double *b, std::size_t N, double *x,
                 std::size_t rank) {
  REF_TRY
  const auto *h = static_cast<const RefHandle *>(hdl);
  const array_t bb(h->n, const_cast<double *>(b), true);
  array_t       xx(h->n, x, true);
  h->M.hifir(h->A, bb, N, xx, false, rank);
  REF_CATCH
}

// HIF::hifir residual-bounded variant (builder.hpp:481-489); out[0]=iters out[1]=flag
int hifref_hifir_betas(const void *hdl, const double *b, std::size_t N,
                       const double *betas, double *x, std::size_t rank, long *out) {
  REF_TRY
  const auto *h = static_cast<const RefHandle *>(hdl);
  const array_t bb(h->n, const_cast<double *>(b), true);
  array_t       xx(h->n, x, true);
  auto          r = h->M.hifir(h->A, bb, N, betas, xx, false, rank);
  out[0]          = (long)r.first;
  out[1]          = r.second;
  REF_CATCH
}

// y = A x through the reference's CRS multiply
int hifref_spmv(const void *hdl, const double *x, double *y) {
  REF_TRY
  const auto *h = static_cast<const RefHandle *>(hdl);
  const array_t xx(h->n, const_cast<double *>(x), true);
  array_t       yy(h->n, y, true);
  h->A.multiply(xx, yy);
  REF_CATCH
}

// the reference's example Krylov drivers, unmodified (gmres.hpp:18-230)
// which: 0 = gmres_hif, 1 = fgmres_hifir ; out[0]=flag out[1]=iters out[2]=num_mv
int hifref_krylov(const void *hdl, int which, const double *b, int restart, double rtol,
                  int maxit, double *x, int *out) {
  REF_TRY
  const auto *h = static_cast<const RefHandle *>(hdl);
  const array_t bb(h->n, const_cast<double *>(b), true);
  if (which == 0) {
    array_t xs;
    int     flag, iters;
    std::tie(xs, flag, iters) = gmres_hif(h->A, bb, h->M, restart, rtol, maxit, 0);
    std::copy_n(xs.cbegin(), h->n, x);
    out[0] = flag;
    out[1] = iters;
    out[2] = iters;
  } else {
    array_t xs;
    int     flag, iters, nmv;
    std::tie(xs, flag, iters, nmv) =
        fgmres_hifir(h->A, bb, h->M, restart, rtol, maxit, 0);
    std::copy_n(xs.cbegin(), h->n, x);
    out[0] = flag;
    out[1] = iters;
    out[2] = nmv;
  }
  REF_CATCH
}

// hif::QRCP<double> on a dense row-major n x n matrix: factorize (QRCP.hpp:107-179) and
// export its state, plus the reference's own solve of b (QRCP.hpp:211-226) -- used to
// pin the dense-level restatement on the reference's known-answer test
// (tests/test_sss_qrcp.cpp:16-197).
int hifref_qrcp_factor_solve(std::size_t n, const double *a_rowmajor, const double *b, double *mat,
                             double *tau, int *jpvt, std::size_t *rank, double *x) {
  REF_TRY
  hif::DenseMatrix<double> D(n, n);
  for (std::size_t i = 0; i < n; ++i)
    for (std::size_t j = 0; j < n; ++j) D(i, j) = a_rowmajor[i * n + j];
  hif::QRCP<double> qr;
  qr.set_matrix(D);
  qr.factorize(hif::get_default_options());
  const auto &acc = static_cast<const QrAcc &>(qr);
  std::copy_n(qr.mat().data(), n * n, mat);
  std::copy_n(acc._tau.cbegin(), n, tau);
  for (std::size_t i = 0; i < n; ++i) jpvt[i] = (int)acc._jpvt[i];
  *rank = qr.rank();
  array_t xx(n);
  std::copy_n(b, n, xx.begin());
  qr.solve(xx);
  std::copy_n(xx.cbegin(), n, x);
  REF_CATCH
}

// the reference's own CCS -> CRS conversion (hif::CRS(const CCS&)) of one factor block,
// to check the device backend's index handling bit-exactly. which: 0=L_B 1=U_B 2=E 3=F
int hifref_export_crs(const void *hdl, int lvl, int which, std::int64_t *row_start, int *col_ind,
                      double *vals) {
  REF_TRY
  const auto &P = level_at(static_cast<const RefHandle *>(hdl), lvl);
  const level_t::mat_type *M =
      which == 0 ? &P.L_B : which == 1 ? &P.U_B : which == 2 ? &P.E : &P.F;
  const level_t::crs_type R(*M);
  for (std::size_t i = 0; i <= R.nrows(); ++i) row_start[i] = R.row_start()[i];
  std::copy_n(R.col_ind().cbegin(), R.nnz(), col_ind);
  std::copy_n(R.vals().cbegin(), R.nnz(), vals);
  REF_CATCH
}

double hifref_norm2(const double *v, std::size_t n) {
  const array_t vv(n, const_cast<double *>(v), true);
  return hif::norm2(vv);
}

// Attach the device backend to the factorized hif::HIF object through the
// product's header-only adapter (include/hifir_b200.hpp). `api` is the table of
// C-ABI entry points resolved by the caller from libhifir_b200.so, so that this
// library has no link-time dependency on CUDA.
int hifref_gpu_attach(const void *hdl, const hifir_b200::AttachApi *api, int device,
                      void **out_gpu_handle) {
  REF_TRY
  const auto *h = static_cast<const RefHandle *>(hdl);
  const int   st = hifir_b200::attach(h->M, *api, device, out_gpu_handle);
  if (st != 0) {
    g_err = "lhfdGpu attach failed with status " + std::to_string(st);
    return st;
  }
  REF_CATCH
}

// ---- the same bridge for a single-precision preconditioner, hif::HIF<float,int> ----------------
// (factors float, user matrix and all vectors double: the reference's mixed-precision path)

using RefHandleS = RefHandleT<float>;

void *hifrefs_create(std::size_t n, const std::int64_t *indptr, const int *indices, const double *vals,
                     const double *params, int verbose) {
  try {
    std::unique_ptr<RefHandleS> h(new RefHandleS());
    h->n = n;
    h->indptr.assign(indptr, indptr + n + 1);
    const std::size_t nnz = (std::size_t)indptr[n];
    h->indices.assign(indices, indices + nnz);
    h->vals.assign(vals, vals + nnz);
    h->A      = crs_t(n, n, h->indptr.data(), h->indices.data(), h->vals.data(), true);
    auto opts = hif::get_default_options();
    if (params) {
      if (params[0] >= 0) opts.tau_L = params[0];
      if (params[1] >= 0) opts.tau_U = params[1];
      if (params[2] >= 0) opts.kappa_d = params[2];
      if (params[3] >= 0) opts.kappa = params[3];
      if (params[4] >= 0) opts.alpha_L = params[4];
      if (params[5] >= 0) opts.alpha_U = params[5];
      if (params[6] > 0) opts.dense_thres = (int)params[6];
      if (params[7] > 0) opts.threads = (int)params[7];
    }
    opts.verbose = verbose ? hif::VERBOSE_INFO : hif::VERBOSE_NONE;
    h->M.factorize(h->A, opts);  // double matrix -> float factors (demo_mixedprecision.cpp)
    return h.release();
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

void        hifrefs_destroy(void *hdl) { delete static_cast<RefHandleS *>(hdl); }
int         hifrefs_num_precs(const void *hdl) { return (int)static_cast<const RefHandleS *>(hdl)->M.precs().size(); }
std::size_t hifrefs_levels(const void *hdl) { return static_cast<const RefHandleS *>(hdl)->M.levels(); }
std::size_t hifrefs_nnz(const void *hdl) { return static_cast<const RefHandleS *>(hdl)->M.nnz(); }

int hifrefs_level_sizes(const void *hdl, int lvl, std::size_t *sizes) {
  REF_TRY
  const auto &P = level_at(static_cast<const RefHandleS *>(hdl), lvl);
  sizes[0]      = P.m;
  sizes[1]      = P.n;
  sizes[2]      = P.L_B.nnz();
  sizes[3]      = P.U_B.nnz();
  sizes[4]      = P.E.nnz();
  sizes[5]      = P.F.nnz();
  sizes[6]      = P.dense_solver.empty() ? 0 : P.dense_solver.mat().nrows();
  sizes[7]      = P.dense_solver.empty() ? 0 : P.dense_solver.rank();
  sizes[8]      = !P.symm_dense_solver.empty();
  sizes[9]      = P.is_last_level();
  REF_CATCH
}

int hifrefs_block_shape(const void *hdl, int lvl, int which, std::size_t *dims) {
  REF_TRY
  const auto &P = level_at(static_cast<const RefHandleS *>(hdl), lvl);
  const RefHandleS::level_t::mat_type *M = which == 0 ? &P.L_B : which == 1 ? &P.U_B : which == 2 ? &P.E : &P.F;
  dims[0] = M->nrows();
  dims[1] = M->ncols();
  REF_CATCH
}

int hifrefs_export_ccs(const void *hdl, int lvl, int which, std::int64_t *col_start, int *row_ind, float *vals) {
  REF_TRY
  const auto &P = level_at(static_cast<const RefHandleS *>(hdl), lvl);
  const RefHandleS::level_t::mat_type *M = which == 0 ? &P.L_B : which == 1 ? &P.U_B : which == 2 ? &P.E : &P.F;
  const std::size_t nc = M->ncols();
  const bool        has = M->col_start().size() >= nc + 1;
  for (std::size_t i = 0; i <= nc; ++i) col_start[i] = has ? M->col_start()[i] : 0;
  std::copy_n(M->row_ind().cbegin(), M->nnz(), row_ind);
  std::copy_n(M->vals().cbegin(), M->nnz(), vals);
  REF_CATCH
}

int hifrefs_export_vectors(const void *hdl, int lvl, float *d, float *s, float *t, int *p, int *p_inv, int *q,
                           int *q_inv) {
  REF_TRY
  const auto &P = level_at(static_cast<const RefHandleS *>(hdl), lvl);
  std::copy_n(P.d_B.cbegin(), P.m, d);
  std::copy_n(P.s.cbegin(), P.n, s);
  std::copy_n(P.t.cbegin(), P.n, t);
  std::copy_n(P.p.cbegin(), P.n, p);
  std::copy_n(P.p_inv.cbegin(), P.n, p_inv);
  std::copy_n(P.q.cbegin(), P.n, q);
  std::copy_n(P.q_inv.cbegin(), P.n, q_inv);
  REF_CATCH
}

int hifrefs_export_dense(const void *hdl, int lvl, float *mat, float *tau, int *jpvt) {
  REF_TRY
  const auto &P  = level_at(static_cast<const RefHandleS *>(hdl), lvl);
  const auto &qr = static_cast<const QrAccT<float> &>(P.dense_solver);
  const std::size_t nm = qr.mat().nrows();
  std::copy_n(qr.mat().data(), nm * qr.mat().ncols(), mat);
  std::copy_n(qr._tau.cbegin(), qr._tau.size(), tau);
  for (std::size_t i = 0; i < qr._jpvt.size(); ++i) jpvt[i] = (int)qr._jpvt[i];
  REF_CATCH
}

int hifrefs_set_nsp_const(void *hdl, std::size_t start, std::size_t end) {
  REF_TRY
  static_cast<RefHandleS *>(hdl)->M.nsp = hif::create_nsp_filter(start, end);
  REF_CATCH
}
int hifrefs_clear_nsp(void *hdl) {
  static_cast<RefHandleS *>(hdl)->M.nsp.reset();
  return 0;
}

// lhfsdApply (libhifir.cpp:1192-1218): op 0 = S, 1 = S^H, 2 = M, 3 = M^H on double vectors
int hifrefs_apply_op(const void *hdl, int op, const double *b, double *x, std::size_t rank) {
  REF_TRY
  const auto *  h = static_cast<const RefHandleS *>(hdl);
  const array_t bb(h->n, const_cast<double *>(b), true);
  array_t       xx(h->n, x, true);
  if (op == 0 || op == 1)
    h->M.solve(bb, xx, op == 1, rank);
  else if (op == 2 || op == 3)
    h->M.mmultiply(bb, xx, op == 3, rank);
  else
    throw std::invalid_argument("bad op");
  REF_CATCH
}

// lhfsApply (libhifir.h:841-854) with op = LHF_S: float vectors
int hifrefs_solve_f32(const void *hdl, const float *b, float *x, std::size_t rank) {
  REF_TRY
  const auto *            h = static_cast<const RefHandleS *>(hdl);
  const hif::Array<float> bb(h->n, const_cast<float *>(b), true);
  hif::Array<float>       xx(h->n, x, true);
  h->M.solve(bb, xx, false, rank);
  REF_CATCH
}

// mixed-precision refinement: float preconditioner, residuals with the double matrix (lhfsdUpdate + lhfsdApply)
int hifrefs_hifir(const void *hdl, const double *b, std::size_t N, double *x, std::size_t rank) {
  REF_TRY
  const auto *  h = static_cast<const RefHandleS *>(hdl);
  const array_t bb(h->n, const_cast<double *>(b), true);
  array_t       xx(h->n, x, true);
  h->M.hifir(h->A, bb, N, xx, false, rank);
  REF_CATCH
}

int hifrefs_gpu_attach(const void *hdl, const hifir_b200::AttachApi *api, int device, void **out_gpu_handle) {
  REF_TRY
  const auto *h  = static_cast<const RefHandleS *>(hdl);
  const int   st = hifir_b200::attach(h->M, *api, device, out_gpu_handle);
  if (st != 0) {
    g_err = "lhfsGpu attach failed with status " + std::to_string(st);
    return st;
  }
  REF_CATCH
}

}  // extern "C"
