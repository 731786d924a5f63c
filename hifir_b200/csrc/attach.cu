// attach.cu -- host side of the device backend: take the plain description of an
// already factorized hif::HIF (LhfdGpuLevel, one per hif::Prec, reference
// alg/Prec.hpp:309-323), convert each CCS block to the row-gather CSR form the kernels
// consume, analyse the triangular dependency structure, upload everything once.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <exception>
#include <memory>
#include <thread>

#include "hifgpu.h"

namespace hifgpu {

// CCS -> CSR, ascending column index inside each row.  Same result as the reference's
// own converting constructor hif::CRS(const CCS&) (ds/CompressedStorage.hpp:861-890);
// integers are moved verbatim (bit-exact index handling).
HostCsr ccs_to_csr(const LhfdGpuCcs &c, const char *name) {
  HostCsr r;
  r.nrows = c.nrows;
  r.ncols = c.ncols;
  r.ptr.assign(c.nrows + 1, 0u);
  if (!c.col_start || !c.ncols) return r;
  const LhfIndPtr nnz = c.col_start[c.ncols];
  if (c.col_start[0] != 0) throw std::invalid_argument(std::string(name) + ": col_start[0] is not 0");
  if (nnz < 0 || static_cast<unsigned long long>(nnz) > 0x7fffffffull)
    throw std::invalid_argument(std::string(name) + ": more than 2^31-1 nonzeros in one block");
  if (nnz && (!c.row_ind || !c.vals)) throw std::invalid_argument(std::string(name) + ": null index/value array");
  for (LhfIndPtr k = 0; k < nnz; ++k) {
    const LhfInt i = c.row_ind[k];
    if (i < 0 || static_cast<std::size_t>(i) >= c.nrows)
      throw std::invalid_argument(std::string(name) + ": row index out of range");
    ++r.ptr[i + 1];
  }
  for (std::size_t i = 0; i < c.nrows; ++i) r.ptr[i + 1] += r.ptr[i];
  r.col.resize(nnz);
  r.val.resize(nnz);
  std::vector<unsigned> next(r.ptr.begin(), r.ptr.end() - 1);
  for (std::size_t j = 0; j < c.ncols; ++j) {
    if (c.col_start[j + 1] < c.col_start[j]) throw std::invalid_argument(std::string(name) + ": col_start not monotone");
    for (LhfIndPtr k = c.col_start[j]; k < c.col_start[j + 1]; ++k) {
      const unsigned pos = next[c.row_ind[k]]++;
      r.col[pos]         = static_cast<int>(j);
      r.val[pos]         = c.vals[k];
    }
  }
  return r;
}

namespace {

void upload_csr(const HostCsr &h, DevCsr &d, std::size_t *tally) {
  d.nrows = h.nrows;
  d.ncols = h.ncols;
  d.nnz   = h.col.size();
  d.ptr.upload(h.ptr, tally);
  d.col.upload(h.col, tally);
  d.val.upload(h.val, tally);
}

// number of level sets of a strictly triangular CSR (row i depends on its columns)
std::size_t dag_depth(const HostCsr &T, bool upper) {
  const std::size_t m = T.nrows;
  if (!m) return 0;
  std::vector<int> lev(m, 0);
  int              mx = 0;
  auto             row = [&](std::size_t i) {
    int l = 0;
    for (unsigned k = T.ptr[i]; k < T.ptr[i + 1]; ++k) l = std::max(l, lev[T.col[k]] + 1);
    lev[i] = l;
    mx     = std::max(mx, l);
  };
  if (!upper)
    for (std::size_t i = 0; i < m; ++i) row(i);
  else
    for (std::size_t i = m; i-- > 0;) row(i);
  return static_cast<std::size_t>(mx) + 1;
}

void check_triangular(const HostCsr &T, bool upper, const char *name) {
  for (std::size_t i = 0; i < T.nrows; ++i)
    for (unsigned k = T.ptr[i]; k < T.ptr[i + 1]; ++k) {
      const std::size_t j = static_cast<std::size_t>(T.col[k]);
      if (upper ? j <= i : j >= i)
        throw std::invalid_argument(std::string(name) + ": block is not strictly triangular");
    }
}

// explicit Q = H_1 H_2 ... H_nm from the Householder vectors stored below the diagonal
// of the QRCP factor and tau (backward accumulation; LAPACK dorg2r's published scheme).
// Column k of Q is row k of Q^T: the device computes (Q^T b)_k as a dot product.
std::vector<double> form_q(std::size_t nm, const double *mat, const double *tau) {
  std::vector<double> Q(nm * nm, 0.0);
  for (std::size_t i = 0; i < nm; ++i) Q[i + i * nm] = 1.0;
  std::vector<double> w(nm);
  for (std::size_t kk = nm; kk-- > 0;) {
    const double  tk = tau[kk];
    if (tk == 0.0) continue;
    const double *v = mat + kk * nm;  // v = [1; mat(kk+1:nm, kk)] on rows kk..nm-1
    // Q(kk:, kk:) -= tk * v * (v^T Q(kk:, kk:))
    for (std::size_t c = kk; c < nm; ++c) {
      double *qc  = Q.data() + c * nm;
      double  dot = qc[kk];
      for (std::size_t i = kk + 1; i < nm; ++i) dot += v[i] * qc[i];
      w[c] = tk * dot;
    }
    for (std::size_t c = kk; c < nm; ++c) {
      double *     qc = Q.data() + c * nm;
      const double wc = w[c];
      qc[kk] -= wc;
      for (std::size_t i = kk + 1; i < nm; ++i) qc[i] -= wc * v[i];
    }
  }
  return Q;
}

}  // namespace

Handle *attach_levels(int device, std::size_t nlevels, const LhfdGpuLevel *lv, bool dense_transposed, bool f32) {
  if (!nlevels || !lv) throw std::invalid_argument("empty preconditioner (no levels)");
  int ndev = 0;
  HIF_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) throw std::invalid_argument("invalid CUDA device ordinal");
  HIF_CUDA(cudaSetDevice(device));

  std::unique_ptr<Handle> h(new Handle());
  h->device = device;
  HIF_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  h->f32    = f32;
  HIF_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  h->levels.resize(nlevels);
  std::size_t *tally = &h->device_bytes;
  std::vector<HostCsr> Lrs(nlevels), Urs(nlevels), Ers(nlevels), Frs(nlevels);  // row-gather forms of the blocks
  // sv = size of a factor value as the reference holds it (4 for hif::HIF<float>);
  // the vectors of an apply are double here whatever the factor precision
  const std::size_t     sv = f32 ? sizeof(float) : sizeof(double);
  constexpr std::size_t sx = sizeof(double);

  for (std::size_t l = 0; l < nlevels; ++l) {
    const LhfdGpuLevel &P = lv[l];
    DevLevel &          D = h->levels[l];
    const std::string   tag = "level " + std::to_string(l);
    if (P.has_symm_dense)
      throw std::invalid_argument(tag + ": symmetric dense factor (SYEIG) is not supported by the device backend");
    if (P.m > P.n) throw std::invalid_argument(tag + ": m > n");
    if (l + 1 < nlevels && P.dense_n) throw std::invalid_argument(tag + ": dense factor on a non-final level");
    if (l + 1 < nlevels && lv[l + 1].n != P.n - P.m)
      throw std::invalid_argument(tag + ": next level size does not match n-m");
    if (P.n > 0x7fffffffull) throw std::invalid_argument(tag + ": n exceeds int32");
    D.m  = P.m;
    D.n  = P.n;
    D.nm = P.n - P.m;
    if (l + 1 == nlevels && D.nm && P.dense_n != D.nm)
      throw std::invalid_argument(tag + ": last level has a Schur block but no dense QRCP factor of that order");
    if (!P.s || !P.t || !P.p || !P.q_inv || (P.m && !P.d_B))
      throw std::invalid_argument(tag + ": missing scaling/permutation/diagonal array");
    for (std::size_t i = 0; i < P.n; ++i)
      if (P.p[i] < 0 || static_cast<std::size_t>(P.p[i]) >= P.n || P.q_inv[i] < 0 ||
          static_cast<std::size_t>(P.q_inv[i]) >= P.n)
        throw std::invalid_argument(tag + ": permutation entry out of range");

    HostCsr &Lr = Lrs[l], &Ur = Urs[l], &Er = Ers[l], &Fr = Frs[l];
    Lr = ccs_to_csr(P.L_B, "L_B"), Ur = ccs_to_csr(P.U_B, "U_B");
    Er = ccs_to_csr(P.E, "E"), Fr = ccs_to_csr(P.F, "F");
    // hif leaves default-constructed (0 x 0) blocks when a part is empty
    Lr.nrows = Lr.ncols = Ur.nrows = Ur.ncols = P.m;
    Lr.ptr.resize(P.m + 1, Lr.ptr.empty() ? 0u : Lr.ptr.back());
    Ur.ptr.resize(P.m + 1, Ur.ptr.empty() ? 0u : Ur.ptr.back());
    Er.ptr.resize(D.nm + 1, Er.ptr.empty() ? 0u : Er.ptr.back());
    Er.nrows = D.nm, Er.ncols = P.m;
    Fr.ptr.resize(P.m + 1, Fr.ptr.empty() ? 0u : Fr.ptr.back());
    Fr.nrows = P.m, Fr.ncols = D.nm;
    check_triangular(Lr, false, "L_B");
    check_triangular(Ur, true, "U_B");
    for (int c : Er.col)
      if (static_cast<std::size_t>(c) >= P.m) throw std::invalid_argument(tag + ": E column out of range");
    for (int c : Fr.col)
      if (static_cast<std::size_t>(c) >= D.nm) throw std::invalid_argument(tag + ": F column out of range");
    D.depthL = dag_depth(Lr, false);
    D.depthU = dag_depth(Ur, true);
  }

  // ---- attach-time analysis of all triangular factors (sweep form + algebraic level merging,
  // merge.cu: 0.5-1 s per factor at 128^3), one host thread per factor.  The level loop below
  // consumes the results in order through the plan cache -- the mechanism an arena file with plans
  // uses (arena.cu); with such a file the analysis is not run at all.
  PlanCache analysed;
  struct CacheScope {
    bool on = false;
    ~CacheScope() {
      if (on) tls_plan_cache = nullptr;
    }
  } cache_scope;
  {
    const MergeParams mp = MergeParams::from_env();
    const char *      et = std::getenv("HIFIR_B200_ATTACH_THREADS");
    if (!tls_plan_cache && mp.enabled && nlevels && (!et || std::atoi(et) != 1)) {
      analysed.f.resize(2 * nlevels);
      std::vector<std::exception_ptr> errs(2 * nlevels);
      std::vector<std::thread>        pool;
      for (std::size_t k = 0; k < 2 * nlevels; ++k)
        pool.emplace_back([&, k] {
          try {
            const bool    upper = (k & 1u) != 0;
            MergedFactor &mf    = analysed.f[k];
            mf.upper            = upper;
            mf.S                = merged_sweep_form(upper ? Urs[k / 2] : Lrs[k / 2], upper, mp, &mf.st);
          } catch (...) {
            errs[k] = std::current_exception();
          }
        });
      for (std::thread &t : pool) t.join();
      for (const std::exception_ptr &e : errs)
        if (e) std::rethrow_exception(e);
      tls_plan_cache = &analysed;
      cache_scope.on = true;
    }
  }

  for (std::size_t l = 0; l < nlevels; ++l) {
    const LhfdGpuLevel &P = lv[l];
    DevLevel &          D = h->levels[l];
    HostCsr &Lr = Lrs[l], &Ur = Urs[l], &Er = Ers[l], &Fr = Frs[l];
    const std::string tag = "level " + std::to_string(l);

    D.L.f32 = D.U.f32 = f32;  // streamed sweep values in single precision
    const char *ef   = std::getenv("HIFIR_B200_WS_FUSE");
    const bool  fuse = sweep_kind() == 2 && P.m && (!ef || std::atoi(ef) != 0);
    if (fuse) {
      // one launch per U^{-1} D^{-1} L^{-1} solve: L's segments followed by U's in every warp's stream
      const MergeParams mp = MergeParams::from_env();
      D.LU.f32             = f32;
      HostCsr SL = merged_sweep_form(Lr, false, mp, &D.L.merge);
      HostCsr SU = merged_sweep_form(Ur, true, mp, &D.U.merge);
      build_ws_ldu_plan(SL, SU, D.LU, D.L, D.U, tally, static_cast<unsigned>(h->num_sms));
    } else {
      // L first: the U sweep reads its right-hand side (L's tagged result) at L's solution slots
      build_sweep_plan(Lr, false, D.L, tally, h->num_sms);
      build_sweep_plan(Ur, true, D.U, tally, h->num_sms, D.L.slot_of.empty() ? nullptr : D.L.slot_of.data());
    }
    if (std::getenv("HIFIR_B200_VERBOSE")) {
      for (const SweepPlan *pl : {&D.L, &D.U})
        std::fprintf(stderr,
                     "[hifir_b200] level %zu %s: rows %zu nnz %zu depth %zu -> merged rows %zu nnz %zu depth %zu "
                     "(steps %zu); packed: slices %u depth %zu padded %zu bytes %.1f MB\n",
                     l, pl->upper ? "U" : "L", pl->merge.rows, pl->merge.nnz, pl->merge.depth, pl->merge.ext_rows,
                     pl->merge.ext_nnz, pl->merge.ext_depth, pl->merge.super_levels, pl->nblocks, pl->st_depth,
                     pl->st_padded, pl->slab_bytes / 1e6);
    }
    {
      // consumers of the renumbered solution slots
      std::vector<unsigned> ls(P.m), us(P.m);
      for (std::size_t r = 0; r < P.m; ++r) {
        ls[r] = D.L.slot_of.empty() ? static_cast<unsigned>(r) : D.L.slot_of[r];
        us[r] = D.U.slot_of.empty() ? static_cast<unsigned>(r) : D.U.slot_of[r];
      }
      std::vector<double> dls(2 * P.m, 1.0);
      for (std::size_t r = 0; r < P.m; ++r) dls[ls[r]] = P.d_B[r];
      D.d_ls.upload(dls, tally);
      std::vector<int> xc(Er.col.size());
      for (std::size_t k = 0; k < xc.size(); ++k) xc[k] = static_cast<int>(us[Er.col[k]]);
      D.E_xcol.upload(xc, tally);
      std::vector<int> qs(P.n);
      for (std::size_t i = 0; i < P.n; ++i) {
        const std::size_t j = static_cast<std::size_t>(P.q_inv[i]);
        qs[i] = j < P.m ? static_cast<int>(us[j]) : -static_cast<int>(j - P.m) - 1;  // < 0: entry of ychild
      }
      D.q_slot.upload(qs, tally);
      D.L_slot.upload(ls, tally);
      D.U_slot.upload(us, tally);
    }
    D.L.nnz = Lr.col.size();
    D.U.nnz = Ur.col.size();
    D.hostL = std::move(Lr);
    D.hostU = std::move(Ur);
    D.h_d.assign(P.d_B, P.d_B + P.m);
    D.h_s.assign(P.s, P.s + P.n);
    D.h_t.assign(P.t, P.t + P.n);
    D.h_p.assign(P.p, P.p + P.n);
    D.h_qinv.assign(P.q_inv, P.q_inv + P.n);
    if (P.p_inv && P.q) {
      D.h_pinv.assign(P.p_inv, P.p_inv + P.n);
      D.h_q.assign(P.q, P.q + P.n);
      for (std::size_t i = 0; i < P.n; ++i)
        if (P.p_inv[i] < 0 || static_cast<std::size_t>(P.p_inv[i]) >= P.n || P.q[i] < 0 ||
            static_cast<std::size_t>(P.q[i]) >= P.n)
          throw std::invalid_argument(tag + ": permutation entry out of range");
    }
    upload_csr(Er, D.E, tally);
    upload_csr(Fr, D.F, tally);
    {
      std::vector<unsigned> rows, cptr;
      for (std::size_t i = 0; i < Fr.nrows; ++i)
        if (Fr.ptr[i + 1] > Fr.ptr[i]) rows.push_back(static_cast<unsigned>(i)), cptr.push_back(Fr.ptr[i]);
      cptr.push_back(Fr.nrows ? Fr.ptr[Fr.nrows] : 0u);
      D.F_rows.upload(rows, tally);
      D.F_cptr.upload(cptr, tally);
    }
    D.hostE = std::move(Er);
    D.hostF = std::move(Fr);
    D.d.upload(P.d_B, P.m, tally);
    D.s.upload(P.s, P.n, tally);
    {
      std::vector<double> sp(P.n);
      for (std::size_t i = 0; i < P.n; ++i) sp[i] = P.s[P.p[i]];
      D.sp.upload(sp, tally);
    }
    D.t.upload(P.t, P.n, tally);
    D.p.upload(reinterpret_cast<const int *>(P.p), P.n, tally);
    D.q_inv.upload(reinterpret_cast<const int *>(P.q_inv), P.n, tally);

    D.bhat.alloc(P.n, tally);
    // tagged sweep results: m solution slots + m slots for the auxiliary unknowns of merge.cu
    D.xL_dn.alloc(2 * P.m, tally);
    D.xU_dn.alloc(2 * P.m, tally);
    D.xL_up.alloc(2 * P.m, tally);
    D.xU_up.alloc(2 * P.m, tally);
    D.r.alloc(D.nm, tally);
    D.ychild.alloc(D.nm, tally);

    // algorithmic bytes per apply, SURVEY.md section 8(d)
    const std::size_t k   = D.nm ? 2 : 1;
    const std::size_t nLU = D.L.nnz + D.U.nnz, nEF = D.E.nnz + D.F.nnz;
    h->bytes_factors += k * (nLU * (sv + 4) + 2 * (P.m + 1) * 4 + P.m * sv) + nEF * (sv + 4) + (P.n + 2) * 4 +
                        P.n * (2 * sv + 2 * 4);
    h->bytes_vec += sx * (D.nm ? (7 * P.n + 8 * P.m) : 8 * P.n);
    h->nnz_total += nLU + nEF + P.m;
  }

  const LhfdGpuLevel &last = lv[nlevels - 1];
  if (last.dense_n) {
    const std::size_t nm = last.dense_n;
    if (!last.qr_mat || !last.qr_tau || !last.qr_jpvt) throw std::invalid_argument("dense level: null QRCP arrays");
    if (last.dense_rank > nm) throw std::invalid_argument("dense level: rank exceeds order");
    std::vector<char> seen(nm, 0);
    for (std::size_t i = 0; i < nm; ++i) {
      const LhfInt j = last.qr_jpvt[i];
      if (j < 1 || static_cast<std::size_t>(j) > nm || seen[j - 1])
        throw std::invalid_argument("dense level: jpvt is not a 1-based permutation");
      seen[j - 1] = 1;
    }
    h->dense.nm   = nm;
    h->dense.rank = last.dense_rank;
    h->dense.transposed = dense_transposed;
    h->dense.h_mat.assign(last.qr_mat, last.qr_mat + nm * nm);
    h->dense.h_tau.assign(last.qr_tau, last.qr_tau + nm);
    h->dense.h_jpvt.assign(last.qr_jpvt, last.qr_jpvt + nm);
    h->dense.Q.upload(form_q(nm, last.qr_mat, last.qr_tau), tally);
    h->dense.R.upload(last.qr_mat, nm * nm, tally);
    {
      // inverses of the 32x32 diagonal tiles of R (upper triangular): column j of the inverse by
      // back substitution on e_j, in double, the tile's own columns only
      const std::size_t   nt = (nm + 31) / 32;
      std::vector<double> tinv(nt * 1024, 0.0);
      for (std::size_t t = 0; t < nt; ++t) {
        const std::size_t j0 = 32 * t, w = std::min<std::size_t>(32, nm - j0);
        double *          T  = tinv.data() + t * 1024;
        for (std::size_t j = 0; j < w; ++j) {
          double col[32] = {0.0};
          col[j]         = 1.0;
          for (std::size_t i = j + 1; i-- > 0;) {
            double v = col[i];
            for (std::size_t k = i + 1; k <= j; ++k) v -= last.qr_mat[(j0 + i) + (j0 + k) * nm] * col[k];
            col[i] = v / last.qr_mat[(j0 + i) + (j0 + i) * nm];
          }
          // a zero or subnormal R(j,j) -- a column beyond the numerical rank of a singular Schur block --
          // gives inf / NaN here; the reference's dtrsv on R(0:rk,0:rk) never touches such a column and
          // the device kernel masks it, so it is stored as 0 (inf * 0 must not poison the tile)
          for (std::size_t i = 0; i <= j; ++i) T[i + 32 * j] = std::isfinite(col[i]) ? col[i] : 0.0;
        }
      }
      h->dense.tinv.upload(tinv, tally);
    }
    h->dense.jpvt.upload(reinterpret_cast<const int *>(last.qr_jpvt), nm, tally);
    h->dense.c.alloc(nm, tally);
    h->bytes_dense = nm * nm * sv + nm * (sv + 4);
    h->nnz_total += nm * nm;
  }

  for (const DevLevel &D : h->levels)
    h->tick_stride = std::max<std::size_t>(
        {h->tick_stride, kSyncStride * (2 + D.L.st_depth), kSyncStride * (2 + D.U.st_depth), kSyncStride * (2 + D.LU.st_depth)});
  h->tickets.alloc((8 * nlevels + 8) * h->tick_stride, tally);
  h->error_flag.alloc(8, tally);  // [0] flag, [1..7] coordinates of the first failing wait (wsweep.cu)
  HIF_CUDA(cudaMallocHost(&h->h_error, sizeof(int)));
  *h->h_error = 0;
  HIF_CUDA(cudaMallocHost(&h->h_scal, 256 * sizeof(double)));
  HIF_CUDA(cudaDeviceSynchronize());
  return h.release();
}

// The transposed preconditioner M^T described with the same structures: with
// P S A T Q = [B F; E C], B ~ L D U   =>   (.)^T has B^T ~ U^T D L^T, E' = F^T, F' = E^T and the
// roles of (s, p, p_inv) and (t, q, q_inv) exchanged.  prec_solve / prec_prod on this description
// are exactly prec_solve_tran / prec_prod_tran of the reference (prec_solve.hpp:541-612,
// prec_prod.hpp:148-230) on the original one, so LHF_SH and LHF_MH reuse every kernel of LHF_S and
// LHF_M.  The CCS arrays of a transposed block are the CSR arrays of the block itself.
Handle *ensure_twin(Handle *h) {
  if (h->twin) return h->twin;
  if (h->is_twin) throw std::logic_error("transpose of a transposed handle requested");
  const std::size_t nl = h->levels.size();
  std::vector<LhfdGpuLevel>            tl(nl);
  std::vector<std::vector<LhfIndPtr>> ptrs;
  ptrs.reserve(4 * nl);
  auto as_ccs = [&](const HostCsr &R, std::size_t nrows_t, std::size_t ncols_t) {
    // R = CSR of X (rows = ncols_t of X^T ... ) -> CCS of X^T
    ptrs.emplace_back(R.ptr.begin(), R.ptr.end());
    if (ptrs.back().size() < ncols_t + 1) ptrs.back().resize(ncols_t + 1, ptrs.back().empty() ? 0 : ptrs.back().back());
    LhfdGpuCcs c;
    c.nrows     = nrows_t;
    c.ncols     = ncols_t;
    c.col_start = ptrs.back().data();
    c.row_ind   = reinterpret_cast<const LhfInt *>(R.col.data());
    c.vals      = R.val.data();
    return c;
  };
  for (std::size_t l = 0; l < nl; ++l) {
    DevLevel &D = h->levels[l];
    if (D.h_pinv.empty() || D.h_q.empty())
      throw std::logic_error("transpose operations need p_inv and q of every level (not given at attach)");
    LhfdGpuLevel &T = tl[l];
    std::memset(&T, 0, sizeof(T));
    T.m     = D.m;
    T.n     = D.n;
    T.L_B   = as_ccs(D.hostU, D.m, D.m);   // (U_B)^T is strictly lower
    T.U_B   = as_ccs(D.hostL, D.m, D.m);   // (L_B)^T is strictly upper
    T.E     = as_ccs(D.hostF, D.nm, D.m);  // F^T : (n-m) x m
    T.F     = as_ccs(D.hostE, D.m, D.nm);  // E^T : m x (n-m)
    T.d_B   = D.h_d.data();
    T.s     = D.h_t.data();
    T.t     = D.h_s.data();
    T.p     = reinterpret_cast<const LhfInt *>(D.h_q.data());
    T.p_inv = reinterpret_cast<const LhfInt *>(D.h_qinv.data());
    T.q     = reinterpret_cast<const LhfInt *>(D.h_p.data());
    T.q_inv = reinterpret_cast<const LhfInt *>(D.h_pinv.data());
    if (l + 1 == nl && h->dense.nm) {
      T.dense_n    = h->dense.nm;
      T.dense_rank = h->dense.rank;
      T.qr_mat     = h->dense.h_mat.data();
      T.qr_tau     = h->dense.h_tau.data();
      T.qr_jpvt    = reinterpret_cast<const LhfInt *>(h->dense.h_jpvt.data());
    }
  }
  Handle *t  = attach_levels(h->device, nl, tl.data(), true, h->f32);
  t->is_twin = true;
  h->twin    = t;
  h->device_bytes += t->device_bytes;
  if (h->has_A)  // A^T: the CSR arrays of A are the CCS arrays of A^T and vice versa
    set_matrix(t, !h->hA_rowmajor, h->n0(), h->hA_ptr.data(), h->hA_idx.data(), h->hA_val.data());
  return t;
}

void destroy_handle(Handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  clear_apply_graphs(h);
  if (h->aio.h2d) {
    cudaStreamSynchronize(h->aio.h2d);
    cudaStreamSynchronize(h->aio.d2h);
    for (int s = 0; s < 2; ++s) {
      cudaEventDestroy(h->aio.copied_in[s]);
      cudaEventDestroy(h->aio.applied[s]);
      cudaEventDestroy(h->aio.copied_out[s]);
    }
    cudaStreamDestroy(h->aio.h2d);
    cudaStreamDestroy(h->aio.d2h);
  }
  if (h->twin) destroy_handle(h->twin);
  if (h->h_error) cudaFreeHost(h->h_error);
  if (h->h_scal) cudaFreeHost(h->h_scal);
  cudaStream_t s = h->own_stream;
  delete h;
  if (s) cudaStreamDestroy(s);
}

void set_matrix(Handle *h, bool rowmajor, std::size_t n, const LhfIndPtr *indptr, const LhfInt *indices,
                const double *vals) {
  if (n != h->n0()) throw std::length_error("matrix size does not match the preconditioner");
  if (!indptr) throw std::invalid_argument("null indptr");
  HIF_CUDA(cudaSetDevice(h->device));
  HostCsr A;
  if (rowmajor) {
    const LhfIndPtr nnz = indptr[n];
    if (nnz < 0 || static_cast<unsigned long long>(nnz) > 0x7fffffffull)
      throw std::invalid_argument("A: more than 2^31-1 nonzeros");
    A.nrows = A.ncols = n;
    A.ptr.resize(n + 1);
    for (std::size_t i = 0; i <= n; ++i) A.ptr[i] = static_cast<unsigned>(indptr[i]);
    A.col.assign(indices, indices + nnz);
    A.val.assign(vals, vals + nnz);
    for (int c : A.col)
      if (c < 0 || static_cast<std::size_t>(c) >= n) throw std::invalid_argument("A: column index out of range");
  } else {
    LhfdGpuCcs c{n, n, indptr, indices, vals};
    A = ccs_to_csr(c, "A");
  }
  h->A.n   = n;
  h->A.nnz = A.col.size();
  upload_csr(A, h->A.A, &h->device_bytes);
  h->has_A = true;
  if (!h->is_twin) {  // keep the caller's arrays for the twin's A^T
    const std::size_t nnz = static_cast<std::size_t>(indptr[n]);
    h->hA_ptr.assign(indptr, indptr + n + 1);
    h->hA_idx.assign(indices, indices + nnz);
    h->hA_val.assign(vals, vals + nnz);
    h->hA_rowmajor = rowmajor;
    if (h->twin) set_matrix(h->twin, !rowmajor, n, indptr, indices, vals);
  }
}

}  // namespace hifgpu
