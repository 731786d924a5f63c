// arena.cu -- factor arena serialization.
//
// The reference can write matrices (utils/io.hpp:76-303) but has no on-disk form of a factorized
// preconditioner: every run pays the Crout factorization again (20+ s at 128^3), and every attach
// pays the attach-time analysis (algebraic level merging, merge.cu: ~1.5-2.5 s per factor).  This
// file gives the device backend its own arena format:
//
//   * the per-level description exactly as attach consumes it -- the members of hif::Prec
//     (alg/Prec.hpp:309-323) and the QRCP state (small_scale/QRCP.hpp:544-555): CCS arrays, integers
//     verbatim, values in the precision of the preconditioner (double, or float for hif::HIF<float>);
//   * optionally the PLANS: the merged sweep-form factors (to_sweep_form + merge_levels) of every
//     L_B / U_B, so that attaching from the file skips the analysis as well.
//
// Writing needs no GPU (plain host code on the plain description); attaching from a file is
// attach_levels() on the arrays read back, with the plans injected through tls_plan_cache.
//
// Layout (native little endian): header {magic "HIFB200A", u32 version, u32 flags, u64 nlevels},
// then per level {u64 m, n, dense_n, dense_rank, flags; blocks L_B U_B E F as {u64 nrows, ncols,
// nnz, has_ptr; i64 col_start[ncols+1]; i32 row_ind[nnz]; V vals[nnz]}; V d[m] s[n] t[n]; i32 p[n]
// q_inv[n] (p_inv[n] q[n]); V qr_mat[dn*dn] qr_tau[dn]; i32 qr_jpvt[dn]}, then, if flagged, per
// level and per factor (L, U) {u64 nrows, orig_rows, nnz; u32 ptr[nrows+1]; i32 col[nnz];
// f64 val[nnz]; u32 gid[nrows]; u64 stats[7]}; trailer u64 = FNV-1a of everything before it.
#include <cstdio>
#include <cstring>
#include <exception>
#include <memory>
#include <thread>

#include "hifgpu.h"

namespace hifgpu {

thread_local PlanCache *tls_plan_cache = nullptr;

// to_sweep_form + merge_levels of one triangular factor, or -- while attaching from an arena file
// that carries plans -- the stored result of exactly that computation
HostCsr merged_sweep_form(const HostCsr &Tnat, bool upper, const MergeParams &mp, MergeStats *ms) {
  if (PlanCache *pc = tls_plan_cache) {
    // plans are stored in attach order; a factor that is never analysed (the U_B of a level with
    // m = 0) leaves its entry unused, so look forward for the first entry that belongs to this factor
    for (std::size_t k = pc->next; mp.enabled && k < pc->f.size(); ++k) {
      MergedFactor &mf = pc->f[k];
      if (mf.S.orig_rows != Tnat.nrows || mf.upper != upper || mf.st.nnz != Tnat.col.size()) continue;
      pc->next = k + 1;
      if (ms) *ms = mf.st;
      return std::move(mf.S);
    }
  }
  HostCsr S = to_sweep_form(Tnat, upper);
  validate_sweep_form(S);  // strictly triangular input: every entry references an earlier sweep row
  if (mp.enabled) S = merge_levels(S, mp, ms, upper);
  return S;
}

namespace {

constexpr char     kMagic[8]  = {'H', 'I', 'F', 'B', '2', '0', '0', 'A'};
constexpr unsigned kVersion   = 1;
constexpr unsigned kFlagF32   = 1u, kFlagPlans = 2u;
constexpr std::uint64_t kLvHasPinvQ = 1u, kLvSymmDense = 2u;

struct Writer {
  std::FILE *   f    = nullptr;
  std::uint64_t hash = 1469598103934665603ull;
  explicit Writer(const char *path) : f(std::fopen(path, "wb")) {
    if (!f) throw std::runtime_error(std::string("cannot open for writing: ") + path);
  }
  ~Writer() {
    if (f) std::fclose(f);
  }
  void raw(const void *p, std::size_t bytes) {
    if (!bytes) return;
    if (std::fwrite(p, 1, bytes, f) != bytes) throw std::runtime_error("arena file: write failed");
    const unsigned char *b = static_cast<const unsigned char *>(p);
    // FNV-1a over 8-byte words (the tail byte-wise): cheap enough for GB-sized arenas
    std::size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
      std::uint64_t w;
      std::memcpy(&w, b + i, 8);
      hash = (hash ^ w) * 1099511628211ull;
    }
    for (; i < bytes; ++i) hash = (hash ^ b[i]) * 1099511628211ull;
  }
  void u64(std::uint64_t v) { raw(&v, 8); }
  void u32(std::uint32_t v) { raw(&v, 4); }
  template <class T>
  void arr(const T *p, std::size_t n) {
    raw(p, n * sizeof(T));
  }
  // values: double in memory, stored in the precision of the preconditioner
  void vals(const double *p, std::size_t n, bool f32) {
    if (!f32) return arr(p, n);
    std::vector<float> v(p, p + n);
    arr(v.data(), n);
  }
  void finish() {
    const std::uint64_t h = hash;
    if (std::fwrite(&h, 1, 8, f) != 8 || std::fflush(f) != 0) throw std::runtime_error("arena file: write failed");
  }
};

struct Reader {
  std::FILE *   f    = nullptr;
  std::uint64_t hash = 1469598103934665603ull;
  explicit Reader(const char *path) : f(std::fopen(path, "rb")) {
    if (!f) throw std::runtime_error(std::string("cannot open arena file: ") + path);
  }
  ~Reader() {
    if (f) std::fclose(f);
  }
  void raw(void *p, std::size_t bytes) {
    if (!bytes) return;
    if (std::fread(p, 1, bytes, f) != bytes) throw std::runtime_error("arena file: truncated");
    const unsigned char *b = static_cast<const unsigned char *>(p);
    std::size_t          i = 0;
    for (; i + 8 <= bytes; i += 8) {
      std::uint64_t w;
      std::memcpy(&w, b + i, 8);
      hash = (hash ^ w) * 1099511628211ull;
    }
    for (; i < bytes; ++i) hash = (hash ^ b[i]) * 1099511628211ull;
  }
  std::uint64_t u64() {
    std::uint64_t v;
    raw(&v, 8);
    return v;
  }
  std::uint32_t u32() {
    std::uint32_t v;
    raw(&v, 4);
    return v;
  }
  // a count read from the file, bounded so that a corrupt header can not drive an allocation
  std::size_t count(std::uint64_t limit = 1ull << 33) {
    const std::uint64_t v = u64();
    if (v > limit) throw std::runtime_error("arena file: implausible array length (corrupt file?)");
    return static_cast<std::size_t>(v);
  }
  template <class T>
  void arr(std::vector<T> &v, std::size_t n) {
    v.resize(n);
    raw(v.data(), n * sizeof(T));
  }
  void vals(std::vector<double> &v, std::size_t n, bool f32) {
    if (!f32) return arr(v, n);
    std::vector<float> t;
    arr(t, n);
    v.assign(t.begin(), t.end());
  }
  void finish() {
    const std::uint64_t h = hash;
    std::uint64_t       stored;
    if (std::fread(&stored, 1, 8, f) != 8) throw std::runtime_error("arena file: truncated (no checksum)");
    if (stored != h) throw std::runtime_error("arena file: checksum mismatch (corrupt file)");
  }
};

void write_ccs(Writer &w, const LhfdGpuCcs &c, bool f32) {
  const bool          has_ptr = c.col_start && c.ncols;
  const std::uint64_t nnz     = has_ptr ? static_cast<std::uint64_t>(c.col_start[c.ncols]) : 0;
  w.u64(c.nrows), w.u64(c.ncols), w.u64(nnz), w.u64(has_ptr ? 1 : 0);
  if (has_ptr) w.arr(c.col_start, c.ncols + 1);
  if (nnz && (!c.row_ind || !c.vals)) throw std::invalid_argument("null index/value array in a factor block");
  w.arr(c.row_ind, nnz);
  w.vals(c.vals, nnz, f32);
}

// the CSR the attach code derives from a triangular CCS block (attach_levels does the same)
HostCsr factor_csr(const LhfdGpuCcs &c, std::size_t m, const char *name) {
  HostCsr r = ccs_to_csr(c, name);
  r.nrows = r.ncols = m;
  r.ptr.resize(m + 1, r.ptr.empty() ? 0u : r.ptr.back());
  return r;
}

void write_plan(Writer &w, const HostCsr &S, const MergeStats &st) {
  w.u64(S.nrows), w.u64(S.orig_rows), w.u64(S.col.size());
  w.arr(S.ptr.data(), S.ptr.size());
  w.arr(S.col.data(), S.col.size());
  w.arr(S.val.data(), S.val.size());
  w.arr(S.gid.data(), S.gid.size());
  const std::uint64_t s[7] = {st.rows, st.ext_rows, st.nnz, st.ext_nnz, st.depth, st.ext_depth, st.super_levels};
  w.arr(s, 7);
}

}  // namespace

void save_levels_file(const char *path, std::size_t nlevels, const LhfdGpuLevel *lv, bool f32, bool with_plans) {
  if (!nlevels || !lv) throw std::invalid_argument("empty preconditioner (no levels)");
  const MergeParams mp = MergeParams::from_env();
  if (with_plans && !mp.enabled) with_plans = false;  // nothing to store: attach will not merge either
  Writer w(path);
  w.raw(kMagic, 8);
  w.u32(kVersion);
  w.u32((f32 ? kFlagF32 : 0u) | (with_plans ? kFlagPlans : 0u));
  w.u64(nlevels);
  for (std::size_t l = 0; l < nlevels; ++l) {
    const LhfdGpuLevel &P = lv[l];
    if (P.m > P.n) throw std::invalid_argument("level " + std::to_string(l) + ": m > n");
    if (!P.s || !P.t || !P.p || !P.q_inv || (P.m && !P.d_B))
      throw std::invalid_argument("level " + std::to_string(l) + ": missing scaling/permutation/diagonal array");
    const bool pq = P.p_inv && P.q;
    w.u64(P.m), w.u64(P.n), w.u64(P.dense_n), w.u64(P.dense_rank);
    w.u64((pq ? kLvHasPinvQ : 0) | (P.has_symm_dense ? kLvSymmDense : 0));
    write_ccs(w, P.L_B, f32), write_ccs(w, P.U_B, f32), write_ccs(w, P.E, f32), write_ccs(w, P.F, f32);
    w.vals(P.d_B, P.m, f32), w.vals(P.s, P.n, f32), w.vals(P.t, P.n, f32);
    w.arr(P.p, P.n), w.arr(P.q_inv, P.n);
    if (pq) w.arr(P.p_inv, P.n), w.arr(P.q, P.n);
    if (P.dense_n) {
      if (!P.qr_mat || !P.qr_tau || !P.qr_jpvt) throw std::invalid_argument("dense level: null QRCP arrays");
      w.vals(P.qr_mat, P.dense_n * P.dense_n, f32), w.vals(P.qr_tau, P.dense_n, f32);
      w.arr(P.qr_jpvt, P.dense_n);
    }
  }
  if (with_plans) {
    // the analysis of every factor on its own host thread (as attach_levels does), written in attach order
    std::vector<MergedFactor>       mf(2 * nlevels);
    std::vector<std::exception_ptr> errs(2 * nlevels);
    std::vector<std::thread>        pool;
    for (std::size_t k = 0; k < 2 * nlevels; ++k)
      pool.emplace_back([&, k] {
        try {
          const bool    upper = (k & 1u) != 0;
          const HostCsr T     = factor_csr(upper ? lv[k / 2].U_B : lv[k / 2].L_B, lv[k / 2].m, upper ? "U_B" : "L_B");
          mf[k].S             = merged_sweep_form(T, upper, mp, &mf[k].st);
        } catch (...) {
          errs[k] = std::current_exception();
        }
      });
    for (std::thread &t : pool) t.join();
    for (const std::exception_ptr &e : errs)
      if (e) std::rethrow_exception(e);
    for (const MergedFactor &f : mf) write_plan(w, f.S, f.st);
  }
  w.finish();
}

struct ArenaFile::Impl {
  std::vector<std::vector<LhfIndPtr>> ptrs;
  std::vector<std::vector<LhfInt>>    ints;
  std::vector<std::vector<double>>    dbls;
};

ArenaFile::ArenaFile() : impl(new Impl()) {}
ArenaFile::~ArenaFile() { delete impl; }

void ArenaFile::load(const char *path, bool want_plans) {
  Reader r(path);
  char   magic[8];
  r.raw(magic, 8);
  if (std::memcmp(magic, kMagic, 8) != 0) throw std::invalid_argument("not a hifir_b200 arena file");
  const unsigned ver = r.u32();
  if (ver != kVersion) throw std::invalid_argument("arena file: unsupported version " + std::to_string(ver));
  const unsigned flags = r.u32();
  f32                  = (flags & kFlagF32) != 0;
  has_plans            = (flags & kFlagPlans) != 0;
  const std::size_t nl = r.count(1u << 20);
  if (!nl) throw std::invalid_argument("arena file: no levels");
  lv.assign(nl, LhfdGpuLevel());
  impl->ptrs.reserve(4 * nl), impl->ints.reserve(10 * nl), impl->dbls.reserve(12 * nl);
  auto ints = [&](std::size_t n) -> const LhfInt * {
    impl->ints.emplace_back();
    r.arr(impl->ints.back(), n);
    return impl->ints.back().data();
  };
  auto dbls = [&](std::size_t n) -> const double * {
    impl->dbls.emplace_back();
    r.vals(impl->dbls.back(), n, f32);
    return impl->dbls.back().data();
  };
  auto ccs = [&]() {
    LhfdGpuCcs c;
    c.nrows               = r.count();
    c.ncols               = r.count();
    const std::size_t nnz = r.count();
    const bool        hp  = r.u64() != 0;
    c.col_start           = nullptr;
    if (hp) {
      impl->ptrs.emplace_back();
      r.arr(impl->ptrs.back(), c.ncols + 1);
      c.col_start = impl->ptrs.back().data();
      if (static_cast<std::size_t>(c.col_start[c.ncols]) != nnz) throw std::runtime_error("arena file: inconsistent block");
    } else if (nnz) {
      throw std::runtime_error("arena file: inconsistent block");
    }
    c.row_ind = ints(nnz);
    c.vals    = dbls(nnz);
    return c;
  };
  nnz_total = 0;
  for (std::size_t l = 0; l < nl; ++l) {
    LhfdGpuLevel &P = lv[l];
    std::memset(&P, 0, sizeof(P));
    P.m = r.count(), P.n = r.count(), P.dense_n = r.count(1u << 20), P.dense_rank = r.count(1u << 20);
    const std::uint64_t lf = r.u64();
    if (P.m > P.n) throw std::runtime_error("arena file: m > n");
    P.L_B = ccs(), P.U_B = ccs(), P.E = ccs(), P.F = ccs();
    P.d_B = dbls(P.m), P.s = dbls(P.n), P.t = dbls(P.n);
    P.p = ints(P.n), P.q_inv = ints(P.n);
    if (lf & kLvHasPinvQ) P.p_inv = ints(P.n), P.q = ints(P.n);
    if (P.dense_n) {
      P.qr_mat = dbls(P.dense_n * P.dense_n), P.qr_tau = dbls(P.dense_n);
      P.qr_jpvt = ints(P.dense_n);
    }
    P.has_symm_dense = (lf & kLvSymmDense) ? 1 : 0;
    for (const LhfdGpuCcs *c : {&P.L_B, &P.U_B, &P.E, &P.F})
      if (c->col_start) nnz_total += static_cast<std::size_t>(c->col_start[c->ncols]);
  }
  plans.f.clear();
  plans.next = 0;
  if (has_plans) {
    for (std::size_t l = 0; l < nl; ++l) {
      for (int upper = 0; upper < 2; ++upper) {
        MergedFactor mf;
        mf.upper          = upper != 0;
        const std::size_t nrows = r.count(), orig = r.count(), nnz = r.count();
        mf.S.nrows = mf.S.ncols = nrows;
        mf.S.orig_rows          = orig;
        r.arr(mf.S.ptr, nrows + 1);
        r.arr(mf.S.col, nnz);
        r.arr(mf.S.val, nnz);
        r.arr(mf.S.gid, nrows);
        std::uint64_t s[7];
        r.raw(s, sizeof(s));
        mf.st.rows = s[0], mf.st.ext_rows = s[1], mf.st.nnz = s[2], mf.st.ext_nnz = s[3];
        mf.st.depth = s[4], mf.st.ext_depth = s[5], mf.st.super_levels = s[6];
        if (mf.S.ptr.empty() || mf.S.ptr.back() != nnz || orig != lv[l].m)
          throw std::runtime_error("arena file: inconsistent plan");
        try {
          validate_sweep_form(mf.S);  // a stale or foreign file must not become an out-of-bounds device access
        } catch (const std::exception &e) {
          throw std::runtime_error(std::string("arena file: inconsistent plan (") + e.what() + ")");
        }
        if (want_plans) plans.f.push_back(std::move(mf));
      }
    }
  }
  r.finish();
}

Handle *attach_file(int device, const char *path, bool want_f32) {
  ArenaFile A;
  A.load(path, true);
  if (A.f32 != want_f32)
    throw std::invalid_argument(A.f32 ? "the arena file holds a single-precision preconditioner: use lhfsGpuAttachFile"
                                      : "the arena file holds a double-precision preconditioner: use lhfdGpuAttachFile");
  struct Scope {  // the plans are consumed by exactly this attach (not by a twin built later)
    explicit Scope(PlanCache *p) { tls_plan_cache = p; }
    ~Scope() { tls_plan_cache = nullptr; }
  } scope(A.has_plans ? &A.plans : nullptr);
  return attach_levels(device, A.lv.size(), A.lv.data(), false, A.f32);
}

}  // namespace hifgpu
