// mrhs.cu -- batched multi-rhs apply: X = M^{-1} B for row-interleaved blocks B[i*nrhs + k]
// (the Array<std::array<T,Nrhs>> layout of hif::HIF::solve_mrhs, reference builder.hpp:433-445).
//
// Semantics = nrhs independent hif::HIF::solve calls (the reference's own multilevel mrhs
// driver prec_solve_mrhs is defective, prec_solve.hpp:489-497 / SURVEY.md App. B-1; its leaf
// kernels -- CompressedStorage.hpp:2117-2137, 2286-2301, 2376-2393, QRCP.hpp:229-242 -- fix the
// per-column arithmetic that is reproduced here).
//
// Default path (warp-stream plans): the block is processed in passes of W = 16, 32 or 64 columns
// (HIFIR_B200_MRHS_WIDE) on the fused L-then-U plans of the single-rhs apply -- wsweep_cols_kernel
// (wsweep.cu) puts the lanes ACROSS the columns, every dependency is one coalesced row of W values
// and every factor is streamed once per W columns.  The glue kernels read the caller's block in
// place (row stride nrhs, column offset of the pass).  Streaming plans (HIFIR_B200_SWEEP=stream)
// keep round 1's passes of kMrhsWidth (= 8) columns.  The schedule is the single-rhs one (apply.cu).
#include <algorithm>
#include <cstdlib>

#include "hifgpu.h"

namespace hifgpu {

constexpr unsigned NR = kMrhsWidth;

namespace {

inline unsigned cdiv(std::size_t a, std::size_t b) { return static_cast<unsigned>((a + b - 1) / b); }

// chunk of user columns [k0, k0+NR) <-> internal [n][NR] block (zero padded)
__global__ void chunk_extract_kernel(const unsigned n, const unsigned nrhs, const unsigned k0,
                                     const double *__restrict__ B, double *__restrict__ out) {
  const std::size_t g = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= static_cast<std::size_t>(n) * NR) return;
  const std::size_t i = g / NR;
  const unsigned    c = static_cast<unsigned>(g % NR);
  out[g]              = k0 + c < nrhs ? B[i * nrhs + k0 + c] : 0.0;
}
__global__ void chunk_insert_kernel(const unsigned n, const unsigned nrhs, const unsigned k0,
                                    const double *__restrict__ in, double *__restrict__ X) {
  const std::size_t g = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= static_cast<std::size_t>(n) * NR) return;
  const std::size_t i = g / NR;
  const unsigned    c = static_cast<unsigned>(g % NR);
  if (k0 + c < nrhs) X[i * nrhs + k0 + c] = in[g];
}

// bhat[i][c] = s[p[i]] * b[p[i]][k0 + c]     (prec_solve.hpp:359, 368, 399); b has row stride ldb and
// nvalid columns from k0 on (the rest of the pass is zero padding)
__global__ void gather_scale_m_kernel(const unsigned n, const unsigned nc, const int *__restrict__ p,
                                      const double *__restrict__ s, const double *__restrict__ b, const unsigned ldb,
                                      const unsigned k0, const unsigned nvalid, double *__restrict__ bhat) {
  const std::size_t g = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= static_cast<std::size_t>(n) * nc) return;
  const std::size_t i = g / nc, c = g % nc;
  const int         pi = p[i];
  bhat[g]              = c < nvalid ? s[pi] * b[static_cast<std::size_t>(pi) * ldb + k0 + c] : 0.0;
}

// The glue of the column-parallel path: a thread owns FOUR columns of a row (32 bytes in flight per thread;
// one element per thread left these kernels latency bound at 3 TB/s: p -> s, b is a chain of two round trips).
// bhat[i][c] = s[p[i]] * b[p[i]][k0 + c], zero beyond nvalid columns; nc is a multiple of 4
__global__ void gather_scale_w_kernel(const unsigned n, const unsigned nc, const int *__restrict__ p,
                                      const double *__restrict__ s, const double *__restrict__ b, const unsigned ldb,
                                      const unsigned k0, const unsigned nvalid, double *__restrict__ bhat) {
  const std::size_t g = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const unsigned    q = nc >> 2;
  if (g >= static_cast<std::size_t>(n) * q) return;
  const std::size_t i = g / q;
  const unsigned    c = static_cast<unsigned>(g % q) * 4u;
  const int         pi = p[i];
  const double      si = s[i];  // s[p[i]], formed at attach
  const double *    src = b + static_cast<std::size_t>(pi) * ldb + k0 + c;
  double            v[4];
#pragma unroll
  for (unsigned j = 0; j < 4; ++j) v[j] = c + j < nvalid ? si * src[j] : 0.0;
  double2 *dst = reinterpret_cast<double2 *>(bhat + i * nc + c);
  dst[0]       = make_double2(v[0], v[1]);
  dst[1]       = make_double2(v[2], v[3]);
}

// y[i][k0 + c] = t[i] * [xU; ychild][q_slot[i]][c] for the first nvalid columns
__global__ void scatter_scale_w_kernel(const unsigned n, const unsigned nc, const int *__restrict__ q_slot,
                                       const double *__restrict__ t, const unsigned long long *__restrict__ xU,
                                       const double *__restrict__ ychild, double *__restrict__ y, const unsigned ldy,
                                       const unsigned k0, const unsigned nvalid) {
  const std::size_t g = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const unsigned    q = nc >> 2;
  if (g >= static_cast<std::size_t>(n) * q) return;
  const std::size_t i = g / q;
  const unsigned    c = static_cast<unsigned>(g % q) * 4u;
  if (c >= nvalid) return;
  const int    j  = q_slot[i];
  const double ti = t[i];
  double       v[4];
  if (j >= 0) {
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(xU + static_cast<std::size_t>(j) * nc + c);
    const ulonglong2  a = src[0], b2 = src[1];
    v[0] = tag_value(a.x), v[1] = tag_value(a.y), v[2] = tag_value(b2.x), v[3] = tag_value(b2.y);
  } else {
    const double2 *src = reinterpret_cast<const double2 *>(ychild + static_cast<std::size_t>(-j - 1) * nc + c);
    const double2  a = src[0], b2 = src[1];
    v[0] = a.x, v[1] = a.y, v[2] = b2.x, v[3] = b2.y;
  }
  double *dst = y + i * ldy + k0 + c;
#pragma unroll
  for (unsigned k = 0; k < 4; ++k)
    if (c + k < nvalid) dst[k] = ti * v[k];
}

// g = bhat - F y in place on the rows of F that have entries (apply.cu: spmv_sub_rows_kernel), two columns
// per thread
__global__ void spmv_sub_rows_m_kernel(const unsigned nrows, const unsigned nc, const unsigned *__restrict__ rows,
                                       const unsigned *__restrict__ cptr, const int *__restrict__ col,
                                       const double *__restrict__ val, const double *__restrict__ x,
                                       double *__restrict__ inout) {
  const std::size_t g = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const unsigned    q = nc >> 1;
  if (g >= static_cast<std::size_t>(nrows) * q) return;
  const std::size_t k = g / q;
  const unsigned    c = static_cast<unsigned>(g % q) * 2u;
  double            a0 = 0.0, a1 = 0.0;
  for (unsigned e = cptr[k], end = cptr[k + 1]; e < end; ++e) {
    const double2 xv = *reinterpret_cast<const double2 *>(x + static_cast<std::size_t>(col[e]) * nc + c);
    const double  v  = val[e];
    a0               = fma(v, xv.x, a0);
    a1               = fma(v, xv.y, a1);
  }
  double2 *io = reinterpret_cast<double2 *>(inout + static_cast<std::size_t>(rows[k]) * nc + c);
  double2  w  = *io;
  w.x -= a0, w.y -= a1;
  *io = w;
}

// out[i][c] = base[i][c] - sum_j A(i,j) x[j][c]; NR consecutive threads = the NR columns of one row
template <bool TAGGED>
__global__ void spmv_resid_m_kernel(const unsigned nrows, const unsigned nc, const unsigned *__restrict__ ptr,
                                    const int *__restrict__ col, const double *__restrict__ val,
                                    const void *__restrict__ xin, const double *__restrict__ base,
                                    double *__restrict__ out) {
  const std::size_t g = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= static_cast<std::size_t>(nrows) * nc) return;
  const std::size_t row = g / nc, c = g % nc;
  double            acc = 0.0;
  const unsigned    e   = ptr[row + 1];
  for (unsigned k = ptr[row]; k < e; ++k) {
    const std::size_t j = static_cast<std::size_t>(col[k]) * nc + c;
    const double      xj =
        TAGGED ? tag_value(static_cast<const unsigned long long *>(xin)[j]) : static_cast<const double *>(xin)[j];
    acc = fma(val[k], xj, acc);
  }
  out[g] = base[g] - acc;
}

// y[i][c] = t[i] * [xU; ychild][q_inv[i]][c]   (prec_solve.hpp:392, 411)
__global__ void scatter_scale_m_kernel(const unsigned n, const unsigned nc, const unsigned m,
                                       const int *__restrict__ q_inv, const double *__restrict__ t,
                                       const unsigned long long *__restrict__ xU, const double *__restrict__ ychild,
                                       double *__restrict__ y) {
  const std::size_t g = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= static_cast<std::size_t>(n) * nc) return;
  const std::size_t i = g / nc, c = g % nc;
  const unsigned    j = static_cast<unsigned>(q_inv[i]);
  const double      w = j < m ? tag_value(xU[static_cast<std::size_t>(j) * nc + c])
                              : ychild[static_cast<std::size_t>(j - m) * nc + c];
  y[g]                = t[i] * w;
}

// dense level: cq[k][c] = Q(:,k)^T x[:, c] : one warp per column k of Q, NR accumulators
// (row stride nc, blockIdx.y = group of NR columns)
__global__ void dense_qt_m_kernel(const unsigned nm, const unsigned rk, const unsigned nc, const double *__restrict__ Q,
                                  const double *__restrict__ x, double *__restrict__ cq) {
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (warp >= rk) return;
  const unsigned c0 = blockIdx.y * NR;
  const double * q  = Q + static_cast<std::size_t>(warp) * nm;
  double         acc[NR];
#pragma unroll
  for (unsigned c = 0; c < NR; ++c) acc[c] = 0.0;
  for (unsigned i = lane; i < nm; i += 32) {
    const double qi = q[i];
#pragma unroll
    for (unsigned c = 0; c < NR; ++c) acc[c] = fma(qi, x[static_cast<std::size_t>(i) * nc + c0 + c], acc[c]);
  }
#pragma unroll
  for (unsigned c = 0; c < NR; ++c) {
    const double v = warp_sum(acc[c]);
    if (lane == 0) cq[static_cast<std::size_t>(warp) * nc + c0 + c] = v;
  }
}

// work vectors of the multi-rhs paths with row stride nc.  The tagged ones are only valid for ONE stride:
// after a change a word of an older layout could carry the tag two epochs later -> zero them, epoch 0.
void ensure_cols(Handle *h, std::size_t nc) {
  if (h->wide_nc == nc) return;
  std::size_t *tally = &h->device_bytes;
  HIF_CUDA(cudaStreamSynchronize(h->stream));
  if (h->wide_cap < nc) {
    for (DevLevel &D : h->levels) {
      D.m_bhat.alloc(D.n * nc, tally);
      D.m_r.alloc(D.nm * nc, tally);
      D.m_ychild.alloc(D.nm * nc, tally);
      // m solution slots + m slots for the auxiliary unknowns of merge.cu, nc values each
      D.m_xL_dn.alloc(2 * D.m * nc, tally);
      D.m_xU_dn.alloc(2 * D.m * nc, tally);
      D.m_xL_up.alloc(2 * D.m * nc, tally);
      D.m_xU_up.alloc(2 * D.m * nc, tally);
    }
    const std::size_t n = h->n0();
    h->mr_b.alloc(n * nc, tally);
    h->mr_x.alloc(n * nc, tally);
    h->mr_c.alloc(h->dense.nm * nc + 1, tally);
    h->wide_cap = nc;
  } else {
    for (DevLevel &D : h->levels)
      for (DevBuf<unsigned long long> *b : {&D.m_xL_dn, &D.m_xU_dn, &D.m_xL_up, &D.m_xU_up})
        if (b->p) HIF_CUDA(cudaMemset(b->p, 0, b->n * sizeof(unsigned long long)));
    HIF_CUDA(cudaDeviceSynchronize());
  }
  h->wide_nc = nc;
  h->epoch_m = 0;  // fresh (zeroed) tagged buffers
}

void ensure_mrhs(Handle *h) {
  if (!h->mrhs_ready) {
    std::size_t *tally = &h->device_bytes;
    for (DevLevel &D : h->levels) {
      // the streaming plans (stream.cu) serve any width; warp-stream plans that are not fused serve one
      // column: the multi-rhs sweeps then run on streaming plans of their own (identity slots)
      if (!D.L.stream) {
        D.Lm.f32 = D.Um.f32 = h->f32;
        build_sweep_plan(D.hostL, false, D.Lm, tally, static_cast<unsigned>(h->num_sms), nullptr, 1);
        build_sweep_plan(D.hostU, true, D.Um, tally, static_cast<unsigned>(h->num_sms), nullptr, 1);
      }
    }
    h->mrhs_ready = true;
  }
  ensure_cols(h, NR);
  for (DevLevel &D : h->levels)  // only these passes keep g = bhat - F y in a buffer of its own
    if (D.m_g.n < D.m * NR) D.m_g.alloc(D.m * NR, &h->device_bytes);
}

}  // namespace

// dense_trsv_kernel of apply.cu, one CTA per column of the chunk
void launch_dense_trsv_cols(Handle *h, unsigned nm, unsigned rk, const double *c, double *out, unsigned ncols);

// one chunk of NR columns: in/out are [n][NR] device blocks
static void apply_chunk(Handle *h, const double *d_in, double *d_out, std::size_t rank) {
  const std::size_t nl = h->levels.size();
  // the multi-rhs work vectors have their own epoch: a tagged buffer must be rewritten on
  // EVERY increment of the epoch whose parity it is checked against
  ++h->epoch_m;
  const unsigned parity = h->epoch_m & 1u;
  HIF_CUDA(cudaMemsetAsync(h->tickets.p, 0, h->tickets.n * sizeof(int), h->stream));
  constexpr int T = 256;
  const double *b = d_in;
  for (std::size_t l = 0; l < nl; ++l) {
    DevLevel &D = h->levels[l];
    if (D.n) {
      gather_scale_m_kernel<<<cdiv(D.n * NR, T), T, 0, h->stream>>>(static_cast<unsigned>(D.n), NR, D.p.p, D.s.p, b, NR, 0u,
                                                                     NR, D.m_bhat.p);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
    }
    if (D.nm) {
      if (D.m) {
        launch_sweep(h, D.L.stream ? D.L : D.Lm, D.m_bhat.p, nullptr, nullptr, D.m_xL_dn.p, parity, h->tick(4 * l), NR);
        launch_sweep(h, D.U.stream ? D.U : D.Um, nullptr, D.m_xL_dn.p, D.d.p, D.m_xU_dn.p, parity,
                     h->tick(4 * l + 1), NR);
      }
      spmv_resid_m_kernel<true><<<cdiv(D.nm * NR, T), T, 0, h->stream>>>(
          static_cast<unsigned>(D.nm), NR, D.E.ptr.p, D.E.col.p, D.E.val.p, D.m_xU_dn.p, D.m_bhat.p + D.m * NR, D.m_r.p);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
      b = D.m_r.p;
    }
  }
  DevLevel &last = h->levels[nl - 1];
  if (last.nm) {
    DevDense &     Q  = h->dense;
    const unsigned nm = static_cast<unsigned>(Q.nm);
    const unsigned rk = static_cast<unsigned>(rank == 0 ? Q.rank : (rank > Q.nm ? Q.nm : rank));
    if (rk) {
      dense_qt_m_kernel<<<cdiv(static_cast<std::size_t>(rk) * 32, T), T, 0, h->stream>>>(nm, rk, NR, Q.Q.p, last.m_r.p,
                                                                                        h->mr_c.p);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
    }
    launch_dense_trsv_cols(h, nm, rk, h->mr_c.p, last.m_ychild.p, NR);
  }
  for (std::size_t l = nl; l-- > 0;) {
    DevLevel &    D   = h->levels[l];
    double *      y   = l == 0 ? d_out : h->levels[l - 1].m_ychild.p;
    const double *rhs = D.m_bhat.p;
    if (D.nm && D.F.nnz && D.m) {
      spmv_resid_m_kernel<false><<<cdiv(D.m * NR, T), T, 0, h->stream>>>(
          static_cast<unsigned>(D.m), NR, D.F.ptr.p, D.F.col.p, D.F.val.p, D.m_ychild.p, D.m_bhat.p, D.m_g.p);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
      rhs = D.m_g.p;
    }
    if (D.m) {
      launch_sweep(h, D.L.stream ? D.L : D.Lm, rhs, nullptr, nullptr, D.m_xL_up.p, parity, h->tick(4 * l + 2),
                   NR);
      launch_sweep(h, D.U.stream ? D.U : D.Um, nullptr, D.m_xL_up.p, D.d.p, D.m_xU_up.p, parity,
                   h->tick(4 * l + 3), NR);
    }
    if (D.n) {
      scatter_scale_m_kernel<<<cdiv(D.n * NR, T), T, 0, h->stream>>>(
          static_cast<unsigned>(D.n), NR, static_cast<unsigned>(D.m), D.q_inv.p, D.t.p, D.m_xU_up.p, D.m_ychild.p, y);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
    }
  }
}

void launch_dense_trsv_cols(Handle *h, unsigned nm, unsigned rk, const double *c, double *out, unsigned ncols);

// one pass of the column-parallel path: columns [k0, k0 + nvalid) of the caller's blocks (row stride nrhs),
// nc = 16 | 32 | 64 internal columns (zero padded), every factor streamed once
static void apply_cols(Handle *h, unsigned nc, const double *d_B, double *d_X, unsigned nrhs, unsigned k0, unsigned nvalid,
                       std::size_t rank) {
  const std::size_t nl = h->levels.size();
  ++h->epoch_m;
  const unsigned parity = h->epoch_m & 1u;
  HIF_CUDA(cudaMemsetAsync(h->tickets.p, 0, h->tickets.n * sizeof(int), h->stream));
  constexpr int T = 256;
  const double *b = d_B;
  unsigned      ldb = nrhs, kb = k0, nvb = nvalid;
  for (std::size_t l = 0; l < nl; ++l) {
    DevLevel &D = h->levels[l];
    if (D.n) {
      gather_scale_w_kernel<<<cdiv(D.n * (nc / 4), T), T, 0, h->stream>>>(static_cast<unsigned>(D.n), nc, D.p.p, D.sp.p, b, ldb,
                                                                           kb, nvb, D.m_bhat.p);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
    }
    if (D.nm) {
      if (D.m) launch_ws_sweep_cols(h, D.LU, D.m_bhat.p, D.d_ls.p, D.m_xL_dn.p, D.m_xU_dn.p, parity, h->tick(8 * l), nc);
      spmv_resid_m_kernel<true><<<cdiv(D.nm * nc, T), T, 0, h->stream>>>(
          static_cast<unsigned>(D.nm), nc, D.E.ptr.p, D.E_xcol.p, D.E.val.p, D.m_xU_dn.p,
          D.m_bhat.p + D.m * static_cast<std::size_t>(nc), D.m_r.p);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
      b = D.m_r.p, ldb = nc, kb = 0u, nvb = nc;
    }
  }
  DevLevel &last = h->levels[nl - 1];
  if (last.nm) {
    DevDense &     Q  = h->dense;
    const unsigned nm = static_cast<unsigned>(Q.nm);
    const unsigned rk = static_cast<unsigned>(rank == 0 ? Q.rank : (rank > Q.nm ? Q.nm : rank));
    if (rk) {
      dense_qt_m_kernel<<<dim3(cdiv(static_cast<std::size_t>(rk) * 32, T), nc / NR), T, 0, h->stream>>>(
          nm, rk, nc, Q.Q.p, last.m_r.p, h->mr_c.p);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
    }
    launch_dense_trsv_cols(h, nm, rk, h->mr_c.p, last.m_ychild.p, nc);
  }
  for (std::size_t l = nl; l-- > 0;) {
    DevLevel &    D   = h->levels[l];
    double *      y   = l == 0 ? d_X : h->levels[l - 1].m_ychild.p;
    if (D.nm && D.F_rows.n && D.m) {
      spmv_sub_rows_m_kernel<<<cdiv(D.F_rows.n * (nc / 2), T), T, 0, h->stream>>>(
          static_cast<unsigned>(D.F_rows.n), nc, D.F_rows.p, D.F_cptr.p, D.F.col.p, D.F.val.p, D.m_ychild.p, D.m_bhat.p);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
    }
    if (D.m) launch_ws_sweep_cols(h, D.LU, D.m_bhat.p, D.d_ls.p, D.m_xL_up.p, D.m_xU_up.p, parity, h->tick(8 * l + 2), nc);
    if (D.n) {
      scatter_scale_w_kernel<<<cdiv(D.n * (nc / 4), T), T, 0, h->stream>>>(
          static_cast<unsigned>(D.n), nc, D.q_slot.p, D.t.p, D.m_xU_up.p, D.m_ychild.p, y, l == 0 ? nrhs : nc,
          l == 0 ? k0 : 0u, l == 0 ? nvalid : nc);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
    }
  }
}

void apply_mrhs_dev(Handle *h, std::size_t nrhs, const double *d_B, double *d_X, std::size_t rank) {
  HIF_CUDA(cudaSetDevice(h->device));
  if (nrhs == 1) {
    apply_dev(h, d_B, d_X, rank);
    return;
  }
  static const int wide_max = [] {  // most columns per pass of the column-parallel path (0: passes of 8 on streaming plans)
    const char *e = std::getenv("HIFIR_B200_MRHS_WIDE");
    const int   w = e ? std::atoi(e) : 64;
    return w >= 64 ? 64 : w >= 32 ? 32 : w >= 16 ? 16 : 0;
  }();
  bool fused = wide_max > 0 && !h->levels.empty();
  for (const DevLevel &D : h->levels) fused = fused && (!D.m || (D.LU.nblocks && D.LU.ws_warps == 24u && D.LU.ws_stages == 2u));
  if (fused) {
    const std::size_t n = h->n0();
    // one internal width per call: the widest pass that the block fills (narrow blocks: 16 or 8 columns)
    const unsigned nc = nrhs >= 64u && wide_max >= 64 ? 64u : nrhs > 16u && wide_max >= 32 ? 32u : nrhs > 8u ? 16u : 8u;
    if (static_cast<unsigned long long>(n) * std::max<std::size_t>(nc, nrhs) > 0xffffffffull) throw std::length_error("multi-rhs block too large");
    ensure_cols(h, nc);
    try {
      for (std::size_t k0 = 0; k0 < nrhs; k0 += nc)
        apply_cols(h, nc, d_B, d_X, static_cast<unsigned>(nrhs), static_cast<unsigned>(k0),
                   static_cast<unsigned>(std::min<std::size_t>(nc, nrhs - k0)), rank);
    } catch (...) {
      reset_tagged_state(h);
      throw;
    }
    return;
  }
  ensure_mrhs(h);
  const std::size_t n = h->n0();
  constexpr int     T = 256;
  for (std::size_t k0 = 0; k0 < nrhs; k0 += NR) {
    chunk_extract_kernel<<<cdiv(n * NR, T), T, 0, h->stream>>>(static_cast<unsigned>(n), static_cast<unsigned>(nrhs),
                                                              static_cast<unsigned>(k0), d_B, h->mr_b.p);
    HIF_KERNEL_CHECK();
    apply_chunk(h, h->mr_b.p, h->mr_x.p, rank);
    chunk_insert_kernel<<<cdiv(n * NR, T), T, 0, h->stream>>>(static_cast<unsigned>(n), static_cast<unsigned>(nrhs),
                                                             static_cast<unsigned>(k0), h->mr_x.p, d_X);
    HIF_KERNEL_CHECK();
    h->launch_count += 2;
  }
}

}  // namespace hifgpu
