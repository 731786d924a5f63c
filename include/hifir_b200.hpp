// hifir_b200.hpp -- header-only C++ adapter: attach the B200 device backend to an
// already factorized hif::HIF object.
//
// The reference's header-only C++ API stays the front door (src/hifir.hpp).  This
// header is what a HIFIR user adds next to `#include <hifir.hpp>`:
//
//     hif::HIF<double, int> M;  M.factorize(A, opts);          // host, unchanged
//     LhfdGpuHdl G;  hifir_b200::attach(M, /*device*/ 0, &G);   // upload factors once
//     lhfdGpuSolve(G, b, x);                                    // = M.solve(b, x)
//
// It walks M.precs() (builder.hpp:152-155), describes each hif::Prec
// (alg/Prec.hpp:309-323) with plain pointers in LhfdGpuLevel and calls the C-ABI
// lhfdGpuAttachLevels.  Nothing of the host object is modified or copied here.
//
// A single-precision preconditioner (hif::HIF<float, int>, the mixed-precision set-up of
// examples/intermediate/demo_mixedprecision.cpp) attaches the same way:
//
//     hif::HIF<float, int> Ms;  Ms.factorize(A);                // A in double
//     LhfsGpuHdl Gs;  hifir_b200::attach(Ms, 0, &Gs);
//     lhfsdGpuSolve(Gs, b, x);                                   // = Ms.solve(b, x), b/x double
//
// The QRCP members _tau/_jpvt are protected with no getter (QRCP.hpp:544-555);
// they are read through a derived accessor type, the same device the reference
// itself uses for its own internals (utils/common.hpp:268-290).
#ifndef HIFIR_B200_HPP_
#define HIFIR_B200_HPP_

#include <cstddef>
#include <type_traits>
#include <vector>

#include "hifir_b200.h"

namespace hifir_b200 {

/// table of C-ABI entry points, for callers that resolve libhifir_b200.so at run
/// time (dlopen / ctypes) instead of linking against it
struct AttachApi {
  LhfStatus (*attach_levels)(int, std::size_t, const LhfdGpuLevel *, LhfdGpuHdl *);
  LhfStatus (*attach_levels_s)(int, std::size_t, const LhfsGpuLevel *, LhfsGpuHdl *);  // may be null
};

/// the C-ABI types that describe factors of value type V
template <class V>
struct AbiOf;
template <>
struct AbiOf<double> {
  typedef LhfdGpuLevel level_type;
  typedef LhfdGpuCcs   ccs_type;
  typedef LhfdGpuHdl   handle_type;
};
template <>
struct AbiOf<float> {
  typedef LhfsGpuLevel level_type;
  typedef LhfsGpuCcs   ccs_type;
  typedef LhfsGpuHdl   handle_type;
};

namespace detail {

template <class Qr>
struct QrAccess : Qr {
  using Qr::_jpvt;
  using Qr::_tau;
};

template <class Ccs>
inline typename AbiOf<typename Ccs::value_type>::ccs_type describe_ccs(const Ccs &M) {
  static_assert(sizeof(typename Ccs::index_type) == sizeof(LhfInt), "int32 indices");
  static_assert(sizeof(typename Ccs::indptr_type) == sizeof(LhfIndPtr), "ptrdiff_t indptr");
  typename AbiOf<typename Ccs::value_type>::ccs_type c;
  c.nrows = M.nrows();
  c.ncols = M.ncols();
  // a default-constructed block has no ind_start array at all
  const bool has_ptr = M.col_start().size() >= M.ncols() + 1;
  c.col_start = has_ptr ? reinterpret_cast<const LhfIndPtr *>(M.col_start().data()) : nullptr;
  c.row_ind   = M.nnz() ? reinterpret_cast<const LhfInt *>(M.row_ind().data()) : nullptr;
  c.vals      = M.nnz() ? M.vals().data() : nullptr;
  return c;
}

}  // namespace detail

/// describe every level of a factorized preconditioner; \a scratch keeps
/// converted pivot arrays alive until the attach call returned
template <class Hif>
inline std::vector<typename AbiOf<typename Hif::value_type>::level_type> describe(
    const Hif &M, std::vector<std::vector<LhfInt>> &scratch) {
  using prec_type  = typename Hif::prec_type;
  using qr_type    = typename prec_type::sss_solver_type;
  using level_type = typename AbiOf<typename Hif::value_type>::level_type;
  std::vector<level_type> lv;
  lv.reserve(M.precs().size());
  scratch.clear();
  scratch.reserve(M.precs().size());
  for (const auto &P : M.precs()) {
    level_type L;
    L.m     = P.m;
    L.n     = P.n;
    L.L_B   = detail::describe_ccs(P.L_B);
    L.d_B   = P.d_B.data();
    L.U_B   = detail::describe_ccs(P.U_B);
    L.E     = detail::describe_ccs(P.E);
    L.F     = detail::describe_ccs(P.F);
    L.s     = P.s.data();
    L.t     = P.t.data();
    L.p     = reinterpret_cast<const LhfInt *>(P.p.data());
    L.p_inv = reinterpret_cast<const LhfInt *>(P.p_inv.data());
    L.q     = reinterpret_cast<const LhfInt *>(P.q.data());
    L.q_inv = reinterpret_cast<const LhfInt *>(P.q_inv.data());
    L.dense_n = L.dense_rank = 0;
    L.qr_mat = L.qr_tau = nullptr;
    L.qr_jpvt           = nullptr;
    L.has_symm_dense    = !P.symm_dense_solver.empty();
    if (!P.dense_solver.empty()) {
      const auto &qr = static_cast<const detail::QrAccess<qr_type> &>(P.dense_solver);
      L.dense_n      = qr.mat().nrows();
      L.dense_rank   = qr.rank();
      L.qr_mat       = qr.mat().data();
      L.qr_tau       = qr._tau.data();
      // hif_lapack_int may be wider than LhfInt: convert once, verbatim values
      scratch.emplace_back(qr._jpvt.size());
      for (std::size_t i = 0; i < qr._jpvt.size(); ++i)
        scratch.back()[i] = static_cast<LhfInt>(qr._jpvt[i]);
      L.qr_jpvt = scratch.back().data();
    }
    lv.push_back(L);
  }
  return lv;
}

namespace detail {
inline LhfStatus call_attach(const AttachApi &api, int device, std::size_t nl, const LhfdGpuLevel *lv, void **out) {
  return api.attach_levels(device, nl, lv, reinterpret_cast<LhfdGpuHdl *>(out));
}
inline LhfStatus call_attach(const AttachApi &api, int device, std::size_t nl, const LhfsGpuLevel *lv, void **out) {
  if (!api.attach_levels_s) return LHF_NULL_OBJ;
  return api.attach_levels_s(device, nl, lv, reinterpret_cast<LhfsGpuHdl *>(out));
}
}  // namespace detail

/// attach through a run-time resolved entry point
template <class Hif>
inline LhfStatus attach(const Hif &M, const AttachApi &api, int device, void **out) {
  std::vector<std::vector<LhfInt>> scratch;
  const auto lv = describe(M, scratch);
  return detail::call_attach(api, device, lv.size(), lv.data(), out);
}

/// attach when the program links against libhifir_b200.so (double factors)
template <class Hif>
inline LhfStatus attach(const Hif &M, int device, LhfdGpuHdl *out) {
  std::vector<std::vector<LhfInt>> scratch;
  const auto lv = describe(M, scratch);
  return lhfdGpuAttachLevels(device, lv.size(), lv.data(), out);
}

/// the same for a single-precision preconditioner, hif::HIF<float, int>
template <class Hif>
inline LhfStatus attach(const Hif &M, int device, LhfsGpuHdl *out) {
  std::vector<std::vector<LhfInt>> scratch;
  const auto lv = describe(M, scratch);
  return lhfsGpuAttachLevels(device, lv.size(), lv.data(), out);
}

}  // namespace hifir_b200

#endif  // HIFIR_B200_HPP_
