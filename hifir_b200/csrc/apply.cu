// apply.cu -- the multilevel M^{-1} apply (reference alg/prec_solve.hpp:332-412) as a
// static schedule of hand-written sm_100a kernels.
//
// The reference's recursion over levels unrolls into a down-sweep, the dense solve and
// an up-sweep.  Per level l (m, n, nm = n-m):
//   down:  bhat = s[p] .* b[p]                                  (gather_scale_kernel)
//          xL   = L^{-1} bhat[0:m]                              (sptrsv_slab_kernel<LOWER>, sptrsv.cu)
//          xU   = U^{-1} (xL ./ d)                              (sptrsv_slab_kernel<UPPER>)
//          r    = bhat[m:n] - E xU         -> b of level l+1    (spmv_resid_kernel)
//   last:  ychild = P R^{-1} Q^T r                              (dense_qt_kernel, dense_trsv_kernel)
//   up:    g    = bhat[0:m] - F ychild  (in place, rows of F with entries: spmv_sub_rows_kernel)
//          xL'  = L^{-1} g ; xU' = U^{-1} (xL' ./ d)
//          y[i] = t[i] * [xU'; ychild][q_inv[i]]                (scatter_scale_kernel)
// `bhat` is evaluated once per level (the reference evaluates s[p]*b[p] up to 3 times,
// prec_solve.hpp:359/368/399).
#include <cstdlib>

#include "hifgpu.h"

namespace hifgpu {

// ============================================================================
// glue kernels
// ============================================================================

// bhat[i] = s[p[i]] * b[p[i]]   (prec_solve.hpp:359, 368, 399)
// (sp[i] = s[p[i]] is formed at attach: one scattered load per row, not two -- the kernel is bound by
// scattered sector requests)
__global__ void gather_scale_kernel(const unsigned n, const int *__restrict__ p, const double *__restrict__ sp,
                                    const double *__restrict__ b, double *__restrict__ bhat) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bhat[i] = sp[i] * b[p[i]];
}

// out[i] = base[i] - sum_j A(i,j) x[j]  with x either tagged (result of a sweep) or plain
// E step: prec_solve.hpp:366-368 ; F step: :395-399 ; accumulation in ascending j like
// CCS::multiply_nt_low (CompressedStorage.hpp:2078-2096).  kLanes lanes cooperate on a row.
template <int kLanes, bool TAGGED>
__global__ void spmv_resid_kernel(const unsigned nrows, const unsigned *__restrict__ ptr,
                                  const int *__restrict__ col, const double *__restrict__ val,
                                  const void *__restrict__ xin, const double *__restrict__ base,
                                  double *__restrict__ out) {
  const unsigned gid  = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned row  = gid / kLanes;
  const unsigned lane = gid % kLanes;
  double         acc  = 0.0;
  if (row < nrows) {
    const unsigned e = ptr[row + 1];
    for (unsigned k = ptr[row] + lane; k < e; k += kLanes) {
      const int j = col[k];
      double    xj;
      if (TAGGED)
        xj = tag_value(static_cast<const unsigned long long *>(xin)[j]);
      else
        xj = static_cast<const double *>(xin)[j];
      acc = fma(val[k], xj, acc);
    }
  }
#pragma unroll
  for (int o = kLanes / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (row < nrows && lane == 0) out[row] = base[row] - acc;
}

// y[i] = t[i] * [xU; ychild][q_inv[i]]   (prec_solve.hpp:392, 411)
// q_slot = q_inv mapped at attach: >= 0 -> solution slot of the U sweep, < 0 -> entry -q-1 of ychild
__global__ void scatter_scale_kernel(const unsigned n, const int *__restrict__ q_slot,
                                     const double *__restrict__ t, const unsigned long long *__restrict__ xU,
                                     const double *__restrict__ ychild, double *__restrict__ y) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int    j = q_slot[i];
    const double w = j >= 0 ? tag_value(xU[j]) : ychild[-j - 1];
    y[i]           = t[i] * w;
  }
}

// ============================================================================
// dense last level: x <- P [R11^{-1} (Q^T x)(1:rk) ; 0]   (small_scale/QRCP.hpp:370-411)
// ============================================================================

// c[k] = Q(:,k)^T x, k < rk : one warp per column of the explicit Q (coalesced)
__global__ void dense_qt_kernel(const unsigned nm, const unsigned rk, const double *__restrict__ Q,
                                const double *__restrict__ x, double *__restrict__ c) {
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (warp >= rk) return;
  const double *q   = Q + static_cast<std::size_t>(warp) * nm;
  double        acc = 0.0;
  for (unsigned i = lane; i < nm; i += 32) acc = fma(q[i], x[i], acc);
  acc = warp_sum(acc);
  if (lane == 0) c[warp] = acc;
}

// Back substitution on R(0:rk,0:rk) (reference: dtrsv 'U','N','N', QRCP.hpp:393), blocked by 32 from
// row 0: the 32x32 diagonal tiles are inverted once at attach (attach.cu; the leading k x k part of
// an upper triangular inverse is the inverse of the leading part, so a truncated rank needs no other
// tiles), hence a tile is a 32x32 matrix-vector product without any dependent chain; then the whole
// CTA applies the tile's 32 columns to the rows above (coalesced column reads, all 32 loads of a
// thread in flight).  Finally out[jpvt[i]-1] = x_i, zero for i >= rk (QRCP.hpp:401-409).
constexpr int kTrsvThreads = 512;
__global__ void __launch_bounds__(kTrsvThreads, 1)
    dense_trsv_kernel(const unsigned nm, const unsigned rk, const double *__restrict__ R,
                      const double *__restrict__ tinv, const double *__restrict__ c,
                      const int *__restrict__ jpvt, double *__restrict__ out, const unsigned stride,
                      const double *__restrict__ Qfull) {
  // one CTA per right-hand side column (blockIdx.x) of a row-interleaved block of `stride` columns
  extern __shared__ double xs[];  // rk values (+ nm values of the input when the Q^T product is fused)
  const unsigned           tid = threadIdx.x, colr = blockIdx.x;
  if (Qfull) {
    // single-rhs path: c = Q(:,1:rk)^T x in the same launch -- one warp per column of the explicit Q,
    // the input vector staged in shared memory (a second launch cost more than this product)
    double *xin = xs + rk;
    for (unsigned i = tid; i < nm; i += kTrsvThreads) xin[i] = c[i];
    __syncthreads();
    const unsigned warp = tid >> 5, lane = tid & 31u;
    for (unsigned k = warp; k < rk; k += kTrsvThreads / 32) {
      const double *q   = Qfull + static_cast<std::size_t>(k) * nm;
      double        acc = 0.0;
      for (unsigned i = lane; i < nm; i += 32) acc = fma(q[i], xin[i], acc);
      acc = warp_sum(acc);
      if (lane == 0) xs[k] = acc;
    }
  } else {
    for (unsigned i = tid; i < rk; i += kTrsvThreads) xs[i] = c[static_cast<std::size_t>(i) * stride + colr];
  }
  __syncthreads();
  const unsigned ntile = (rk + 31u) / 32u;
  for (unsigned t = ntile; t-- > 0;) {
    const unsigned j0 = 32u * t, w = min(32u, rk - j0);  // tile [j0, j0 + w)
    if (tid < 32) {
      // x_i = sum_{j >= i} Tinv(i, j) c_j : 4 independent partial sums
      const double *Ti = tinv + static_cast<std::size_t>(t) * 1024u + tid;  // column-major 32x32, lane = row
      double        a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      if (tid < w) {
#pragma unroll 8
        for (unsigned j = 0; j < 32u; j += 4) {
          // masked columns (beyond the rank, below the diagonal) enter as 0 * finite: the tile inverses
          // hold no inf / NaN (attach.cu stores 0 where a zero pivot beyond the numerical rank gave one),
          // and the loads stay unconditional so that all of them are in flight together
          a0 = fma(Ti[(j + 0u) * 32u], (j + 0u < w && j + 0u >= tid) ? xs[j0 + j + 0u] : 0.0, a0);
          a1 = fma(Ti[(j + 1u) * 32u], (j + 1u < w && j + 1u >= tid) ? xs[j0 + j + 1u] : 0.0, a1);
          a2 = fma(Ti[(j + 2u) * 32u], (j + 2u < w && j + 2u >= tid) ? xs[j0 + j + 2u] : 0.0, a2);
          a3 = fma(Ti[(j + 3u) * 32u], (j + 3u < w && j + 3u >= tid) ? xs[j0 + j + 3u] : 0.0, a3);
        }
      }
      const double xv = (a0 + a1) + (a2 + a3);
      __syncwarp();
      if (tid < w) xs[j0 + tid] = xv;
    }
    __syncthreads();
    // rows above the tile: x[i] -= sum_{col in tile} R(i,col) * x[col]
    for (unsigned i = tid; i < j0; i += kTrsvThreads) {
      const double *Ri = R + i + static_cast<std::size_t>(j0) * nm;
      double        r[32];
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) r[jj] = static_cast<unsigned>(jj) < w ? Ri[static_cast<std::size_t>(jj) * nm] : 0.0;
      double b0 = 0.0, b1 = 0.0;
#pragma unroll
      for (int jj = 0; jj < 32; jj += 2) {  // r[jj] = 0 for a masked column (loaded that way above)
        b0 = fma(r[jj], static_cast<unsigned>(jj) < w ? xs[j0 + jj] : 0.0, b0);
        b1 = fma(r[jj + 1], static_cast<unsigned>(jj + 1) < w ? xs[j0 + jj + 1] : 0.0, b1);
      }
      xs[i] -= b0 + b1;
    }
    __syncthreads();
  }
  for (unsigned i = tid; i < nm; i += kTrsvThreads)
    out[static_cast<std::size_t>(jpvt[i] - 1) * stride + colr] = i < rk ? xs[i] : 0.0;
}

// ============================================================================
// null-space filter, constant mode: x[start:end] -= mean   (NspFilter.hpp:161-175)
// ============================================================================
constexpr int kRedBlocks = 2 * kNumSMs, kRedThreads = 512;

__global__ void __launch_bounds__(kRedThreads) sum_partial_kernel(const double *__restrict__ x, const unsigned n,
                                                                  double *__restrict__ part) {
  __shared__ double sm[kRedThreads / 32];
  double            acc = 0.0;
  for (unsigned i = blockIdx.x * kRedThreads + threadIdx.x; i < n; i += gridDim.x * kRedThreads) acc += x[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < kRedThreads / 32 ? sm[threadIdx.x] : 0.0;
    acc = warp_sum(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
  }
}

__global__ void __launch_bounds__(kRedThreads) shift_kernel(double *__restrict__ x, const unsigned n,
                                                            const double *__restrict__ part, const unsigned nparts) {
  // every block re-reduces the (few) partials in the same fixed order -> deterministic
  __shared__ double sm[kRedThreads / 32];
  __shared__ double shift;
  double            acc = 0.0;
  for (unsigned i = threadIdx.x; i < nparts; i += kRedThreads) acc += part[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < kRedThreads / 32 ? sm[threadIdx.x] : 0.0;
    acc = warp_sum(acc);
    if (threadIdx.x == 0) shift = acc / static_cast<double>(n);
  }
  __syncthreads();
  const double sh = shift;
  for (unsigned i = blockIdx.x * kRedThreads + threadIdx.x; i < n; i += gridDim.x * kRedThreads) x[i] -= sh;
}

// ============================================================================
// host-side schedule
// ============================================================================
namespace {

inline unsigned cdiv(std::size_t a, std::size_t b) { return static_cast<unsigned>((a + b - 1) / b); }

// profiling mode only: an event after each kernel of the schedule
void mark(Handle *h, const std::string &name) {
  if (!h->profiling) return;
  cudaEvent_t e;
  HIF_CUDA(cudaEventCreate(&e));
  HIF_CUDA(cudaEventRecord(e, h->stream));
  h->prof_marks.emplace_back(name, e);
}

void launch_sweeps(Handle *h, DevLevel &D, const double *rhs, unsigned long long *xL, unsigned long long *xU,
                   unsigned parity, int *tickets, const std::string &tag, int trace_base = -100) {
  if (!D.m) return;
  if (D.LU.nblocks) {  // L then U in one launch (wsweep.cu, MODE 2); d permuted to the L sweep's slots
    launch_ws_sweep(h, D.LU, rhs, nullptr, D.d_ls.p, xL, parity, tickets, nullptr, xU);
    mark(h, tag + "LU");
    return;
  }
  const bool tl = h->trace_level >= 0 && &h->levels[h->trace_level] == &D;
  launch_sweep(h, D.L, rhs, nullptr, nullptr, xL, parity, tickets, 0,
               tl && h->trace_which == trace_base ? h->trace_buf.p : nullptr);
  mark(h, tag + "L");
  // the U sweep reads x_L at the L sweep's slots and divides by d (permuted to those slots)
  launch_sweep(h, D.U, nullptr, xL, D.d_ls.p, xU, parity, tickets + h->tick_stride, 0,
               tl && h->trace_which == trace_base + 1 ? h->trace_buf.p : nullptr);
  mark(h, tag + "U");
}

// g = bhat - F y in place: only the rows of F that have entries change (bhat[0:m] is not needed again in
// this apply); one thread per such row (1.6 entries on average)
__global__ void spmv_sub_rows_kernel(const unsigned nrows, const unsigned *__restrict__ rows, const unsigned *__restrict__ cptr,
                                     const int *__restrict__ col, const double *__restrict__ val,
                                     const double *__restrict__ x, double *__restrict__ inout) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nrows) return;
  double acc = 0.0;
  for (unsigned e = cptr[k], end = cptr[k + 1]; e < end; ++e) acc = fma(val[e], x[col[e]], acc);
  const unsigned r = rows[k];
  inout[r]         = inout[r] - acc;
}

template <bool TAGGED>
void launch_spmv_resid(Handle *h, const DevCsr &A, const int *col, const void *x, const double *base, double *out,
                       const std::string &tag) {
  if (!A.nrows) return;
  const double avg = A.nrows ? static_cast<double>(A.nnz) / static_cast<double>(A.nrows) : 0.0;
  constexpr int T  = 256;
  if (avg > 12.0) {
    spmv_resid_kernel<8, TAGGED><<<cdiv(A.nrows * 8, T), T, 0, h->stream>>>(
        static_cast<unsigned>(A.nrows), A.ptr.p, col, A.val.p, x, base, out);
  } else if (avg > 3.0) {
    spmv_resid_kernel<4, TAGGED><<<cdiv(A.nrows * 4, T), T, 0, h->stream>>>(
        static_cast<unsigned>(A.nrows), A.ptr.p, col, A.val.p, x, base, out);
  } else {
    spmv_resid_kernel<1, TAGGED><<<cdiv(A.nrows, T), T, 0, h->stream>>>(static_cast<unsigned>(A.nrows), A.ptr.p,
                                                                        col, A.val.p, x, base, out);
  }
  HIF_KERNEL_CHECK();
  mark(h, tag);
  ++h->launch_count;
}

}  // namespace

namespace {
// parts of the schedule: the only kernel that reads the caller's b (level-0 gather), the only ones that write
// the caller's x (level-0 scatter, null-space filter), and everything in between (handle-internal buffers only)
constexpr unsigned kHead = 1u, kBody = 2u, kTail = 4u;
void apply_schedule(Handle *h, const double *d_b, double *d_x, std::size_t rank, unsigned parity, unsigned parts);
}
void apply_dev_impl(Handle *h, const double *d_b, double *d_x, std::size_t rank);
void reset_tagged_state(Handle *h);

// The schedule of one apply is static and only its first and last kernels touch the caller's vectors: the
// kernels in between are captured into a CUDA graph per (rank, tag parity) and replayed for every b and x
// (the kernels of an apply are short -- 8 to 270 us -- and there are 14 of them).
void apply_dev(Handle *h, const double *d_b, double *d_x, std::size_t rank) {
  try {
    apply_dev_impl(h, d_b, d_x, rank);
  } catch (...) {
    reset_tagged_state(h);
    throw;
  }
}

void apply_dev_impl(Handle *h, const double *d_b, double *d_x, std::size_t rank) {
  HIF_CUDA(cudaSetDevice(h->device));
  ++h->epoch;
  const unsigned parity = h->epoch & 1u;
  static const bool graphs_on = [] {
    const char *e = std::getenv("HIFIR_B200_GRAPH");
    return !e || std::atoi(e) != 0;
  }();
  const std::size_t launches0 = h->launch_count;
  if (!graphs_on || h->graphs_off || h->profiling || h->trace_level >= 0) {
    apply_schedule(h, d_b, d_x, rank, parity, kHead | kBody | kTail);
    h->kernels_per_apply = h->launch_count - launches0;
    return;
  }
  // The graph holds the BODY of the schedule -- every kernel between the level-0 gather (the only reader of b)
  // and the level-0 scatter / null-space filter (the only writers of x) -- so one graph per (rank, tag parity)
  // serves every pair of vectors: the basis vectors of FGMRES, the staging slots of the pipelined host
  // solves, a refinement loop.  Capturing and instantiating costs ~1.7 ms, a graph launch saves ~30 us: the
  // body is captured on the third use of its key.
  ApplyGraph *g = nullptr;
  for (ApplyGraph &c : h->graphs)
    if (c.rank == rank && c.parity == parity) g = &c;
  if (!g) {
    if (h->graphs.size() >= 256) clear_apply_graphs(h);
    h->graphs.push_back(ApplyGraph{nullptr, nullptr, rank, parity, false, 0, 0, 0, nullptr, 0, 0});
    g = &h->graphs.back();
  }
  if (!g->exec && ++g->uses < 3u) {
    apply_schedule(h, d_b, d_x, rank, parity, kHead | kBody | kTail);
    h->kernels_per_apply = h->launch_count - launches0;
    return;
  }
  apply_schedule(h, d_b, d_x, rank, parity, kHead);
  if (!g->exec) {
    const std::size_t body0 = h->launch_count;
    cudaGraph_t       graph = nullptr;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      // e.g. the legacy default stream can not be captured: this handle launches its kernels one by one
      cudaGetLastError();
      h->graphs_off = true;
      apply_schedule(h, d_b, d_x, rank, parity, kBody | kTail);
      h->kernels_per_apply = h->launch_count - launches0;
      return;
    }
    try {
      apply_schedule(h, d_b, d_x, rank, parity, kBody);
    } catch (...) {
      cudaStreamEndCapture(h->stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    HIF_CUDA(cudaStreamEndCapture(h->stream, &graph));
    const cudaError_t e = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    cuda_check(e, "cudaGraphInstantiate", __FILE__, __LINE__);
    g->launches     = h->launch_count - body0;
    h->launch_count = body0;
  }
  HIF_CUDA(cudaGraphLaunch(g->exec, h->stream));
  h->launch_count += g->launches;
  apply_schedule(h, d_b, d_x, rank, parity, kTail);
  h->kernels_per_apply = h->launch_count - launches0;
}

// An apply that failed half way (launch error, exception) has bumped the epoch but rewritten only some
// of the tagged buffers: two applies later their stale slots would read as ready.  Start over: zero
// every tagged buffer, epoch 0.
void reset_tagged_state(Handle *h) {
  cudaStreamSynchronize(h->stream);
  cudaGetLastError();
  auto zero = [&](DevBuf<unsigned long long> &b) {
    if (b.p) cudaMemset(b.p, 0, b.n * sizeof(unsigned long long));
  };
  for (DevLevel &D : h->levels)
    for (DevBuf<unsigned long long> *b : {&D.xL_dn, &D.xU_dn, &D.xL_up, &D.xU_up, &D.p_xL, &D.p_xU, &D.m_xL_dn, &D.m_xU_dn,
                                          &D.m_xL_up, &D.m_xU_up})
      zero(*b);
  h->epoch = h->epoch_m = h->epoch_p = 0;
  cudaDeviceSynchronize();
}

void clear_apply_graphs(Handle *h) {
  for (ApplyGraph &c : h->graphs)
    if (c.exec) cudaGraphExecDestroy(c.exec);
  h->graphs.clear();
}

namespace {
void apply_schedule(Handle *h, const double *d_b, double *d_x, std::size_t rank, unsigned parity, unsigned parts) {
  const std::size_t nl   = h->levels.size();
  const bool        body = (parts & kBody) != 0u;
  if (body) HIF_CUDA(cudaMemsetAsync(h->tickets.p, 0, h->tickets.n * sizeof(int), h->stream));
  constexpr int T = 256;
  if (parts & kHead) mark(h, "begin");

  // ---- down-sweep
  const double *b = d_b;
  for (std::size_t l = 0; l < nl; ++l) {
    DevLevel &D = h->levels[l];
    if (D.n && (parts & (l == 0 ? kHead : kBody))) {
      gather_scale_kernel<<<cdiv(D.n, T), T, 0, h->stream>>>(static_cast<unsigned>(D.n), D.p.p, D.sp.p, b, D.bhat.p);
      HIF_KERNEL_CHECK();
      mark(h, "lv" + std::to_string(l) + ".gather");
      ++h->launch_count;
    }
    if (D.nm) {
      if (body) {
        launch_sweeps(h, D, D.bhat.p, D.xL_dn.p, D.xU_dn.p, parity, h->tick(8 * l), "lv" + std::to_string(l) + ".down.", 0);
        launch_spmv_resid<true>(h, D.E, D.E_xcol.p, D.xU_dn.p, D.bhat.p + D.m, D.r.p, "lv" + std::to_string(l) + ".E");
      }
      b = D.r.p;
    }
  }
  // ---- dense last level (QRCP.hpp:370-411); rank: 0 -> numerical, > nm -> nm
  DevLevel &last = h->levels[nl - 1];
  if (last.nm && body) {
    dense_solve_dev(h, last.r.p, last.ychild.p, rank);
    mark(h, "dense");
  }
  // ---- up-sweep
  for (std::size_t l = nl; l-- > 0;) {
    DevLevel &    D      = h->levels[l];
    double *      y      = l == 0 ? d_x : h->levels[l - 1].ychild.p;
    const double *rhs    = D.bhat.p;
    if (D.nm && D.F_rows.n && body) {
      spmv_sub_rows_kernel<<<cdiv(D.F_rows.n, T), T, 0, h->stream>>>(static_cast<unsigned>(D.F_rows.n), D.F_rows.p, D.F_cptr.p,
                                                                     D.F.col.p, D.F.val.p, D.ychild.p, D.bhat.p);
      HIF_KERNEL_CHECK();
      mark(h, "lv" + std::to_string(l) + ".F");
      ++h->launch_count;
    }
    if (body) launch_sweeps(h, D, rhs, D.xL_up.p, D.xU_up.p, parity, h->tick(8 * l + 2), "lv" + std::to_string(l) + ".up.", 2);
    if (D.n && (parts & (l == 0 ? kTail : kBody))) {
      scatter_scale_kernel<<<cdiv(D.n, T), T, 0, h->stream>>>(static_cast<unsigned>(D.n), D.q_slot.p, D.t.p,
                                                             D.xU_up.p, D.ychild.p, y);
      HIF_KERNEL_CHECK();
      mark(h, "lv" + std::to_string(l) + ".scatter");
      ++h->launch_count;
    }
  }
  // ---- null-space filter (builder.hpp:419-420)
  if (h->nsp_on && (parts & kTail)) {
    const std::size_t n     = h->n0();
    std::size_t       start = h->nsp_start, end = h->nsp_end;
    if (end == static_cast<std::size_t>(-1) || end < start) end = n;
    if (end > n) throw std::out_of_range("null-space filter range exceeds the system size");
    if (end > start) {
      if (h->kr_part.n < static_cast<std::size_t>(kRedBlocks)) h->kr_part.alloc(kRedBlocks, &h->device_bytes);
      const unsigned len = static_cast<unsigned>(end - start);
      const unsigned nb  = std::min<unsigned>(kRedBlocks, cdiv(len, kRedThreads));
      sum_partial_kernel<<<nb, kRedThreads, 0, h->stream>>>(d_x + start, len, h->kr_part.p);
      HIF_KERNEL_CHECK();
      shift_kernel<<<nb, kRedThreads, 0, h->stream>>>(d_x + start, len, h->kr_part.p, nb);
      HIF_KERNEL_CHECK();
      mark(h, "nsp");
      h->launch_count += 2;
    }
  }
}
}  // namespace

// ---- the four dense operations of QRCP (small_scale/QRCP.hpp:370-540) ------------------
// single-CTA kernels for the transposed solve and the two products: off the BASELINE path,
// written for clarity; arithmetic order = the reference BLAS (dtrsv / dtrmv column forms)
constexpr int kDenseT = 512;

// out = Q(:,1:rk) * R(1:rk,1:rk)^{-T} * (P^T x)(1:rk)      QRCP::_solve_t (QRCP.hpp:417-453)
__global__ void __launch_bounds__(kDenseT, 1)
    dense_solve_t_kernel(const unsigned nm, const unsigned rk, const double *__restrict__ R,
                         const double *__restrict__ Q, const int *__restrict__ jpvt, const double *__restrict__ x,
                         double *__restrict__ out) {
  extern __shared__ double w[];
  const unsigned           tid = threadIdx.x;
  for (unsigned i = tid; i < rk; i += kDenseT) w[i] = x[jpvt[i] - 1];
  __syncthreads();
  for (unsigned j = 0; j < rk; ++j) {  // forward substitution with R^T, updates in ascending j
    if (tid == 0) w[j] /= R[j + static_cast<std::size_t>(j) * nm];
    __syncthreads();
    const double wj = w[j];
    for (unsigned i = j + 1 + tid; i < rk; i += kDenseT) w[i] -= R[j + static_cast<std::size_t>(i) * nm] * wj;
    __syncthreads();
  }
  for (unsigned i = tid; i < nm; i += kDenseT) {
    double acc = 0.0;
    for (unsigned k = 0; k < rk; ++k) acc = fma(Q[i + static_cast<std::size_t>(k) * nm], w[k], acc);
    out[i] = acc;
  }
}

// out = Q(:,1:rk) * R(1:rk,1:rk) * (P^T x)(1:rk)            QRCP::_multiply_nt (QRCP.hpp:459-492)
__global__ void __launch_bounds__(kDenseT, 1)
    dense_mult_nt_kernel(const unsigned nm, const unsigned rk, const double *__restrict__ R,
                         const double *__restrict__ Q, const int *__restrict__ jpvt, const double *__restrict__ x,
                         double *__restrict__ out) {
  extern __shared__ double w[];  // 2 * rk
  double *                 v   = w + rk;
  const unsigned           tid = threadIdx.x;
  for (unsigned i = tid; i < rk; i += kDenseT) w[i] = x[jpvt[i] - 1];
  __syncthreads();
  for (unsigned i = tid; i < rk; i += kDenseT) {  // dtrmv 'U','N','N': diagonal term first, then ascending j
    double acc = R[i + static_cast<std::size_t>(i) * nm] * w[i];
    for (unsigned j = i + 1; j < rk; ++j) acc = fma(R[i + static_cast<std::size_t>(j) * nm], w[j], acc);
    v[i] = acc;
  }
  __syncthreads();
  for (unsigned i = tid; i < nm; i += kDenseT) {
    double acc = 0.0;
    for (unsigned k = 0; k < rk; ++k) acc = fma(Q[i + static_cast<std::size_t>(k) * nm], v[k], acc);
    out[i] = acc;
  }
}

// out = P [R(1:rk,1:rk)^T c ; 0] with c = Q(:,1:rk)^T x given   QRCP::_multiply_t (QRCP.hpp:494-540)
__global__ void __launch_bounds__(kDenseT, 1)
    dense_mult_t_kernel(const unsigned nm, const unsigned rk, const double *__restrict__ R,
                        const double *__restrict__ c, const int *__restrict__ jpvt, double *__restrict__ out) {
  extern __shared__ double w[];
  const unsigned           tid = threadIdx.x;
  for (unsigned i = tid; i < rk; i += kDenseT) w[i] = c[i];
  __syncthreads();
  for (unsigned j = tid; j < nm; j += kDenseT) {
    double v = 0.0;
    if (j < rk) {  // dtrmv 'U','C','N': diagonal term, then rows i < j in descending order
      v = w[j] * R[j + static_cast<std::size_t>(j) * nm];
      for (unsigned i = j; i-- > 0;) v = fma(R[i + static_cast<std::size_t>(j) * nm], w[i], v);
    }
    out[jpvt[j] - 1] = v;
  }
}

static unsigned dense_rank_of(const DevDense &Q, std::size_t rank) {
  return static_cast<unsigned>(rank == 0 ? Q.rank : (rank > Q.nm ? Q.nm : rank));
}

// QRCP::solve(x, rank, tran) with tran = the handle's orientation
void dense_solve_dev(Handle *h, const double *d_in, double *d_out, std::size_t rank) {
  DevDense &     Q  = h->dense;
  const unsigned nm = static_cast<unsigned>(Q.nm), rk = dense_rank_of(Q, rank);
  if (!Q.transposed) {
    // Q^T product and back substitution in one launch only for tiny dense levels: one SM reading the
    // whole explicit Q is slow (measured at nm = 215: 54 us in one launch against 34 us with the
    // multi-CTA product in front)
    const bool one_launch = nm <= 64u;
    if (rk && !one_launch) {
      dense_qt_kernel<<<cdiv(static_cast<std::size_t>(rk) * 32, 256), 256, 0, h->stream>>>(nm, rk, Q.Q.p, d_in, Q.c.p);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
    }
    if (one_launch)
      dense_trsv_kernel<<<1, kTrsvThreads, (rk + nm + 1) * sizeof(double), h->stream>>>(nm, rk, Q.R.p, Q.tinv.p, d_in,
                                                                                       Q.jpvt.p, d_out, 1u, Q.Q.p);
    else
      dense_trsv_kernel<<<1, kTrsvThreads, (rk ? rk : 1) * sizeof(double), h->stream>>>(nm, rk, Q.R.p, Q.tinv.p, Q.c.p,
                                                                                       Q.jpvt.p, d_out, 1u, nullptr);
  } else {
    dense_solve_t_kernel<<<1, kDenseT, (rk ? rk : 1) * sizeof(double), h->stream>>>(nm, rk, Q.R.p, Q.Q.p, Q.jpvt.p, d_in,
                                                                                   d_out);
  }
  HIF_KERNEL_CHECK();
  ++h->launch_count;
}

// QRCP::multiply(x, rank, tran)
void dense_multiply_dev(Handle *h, const double *d_in, double *d_out, std::size_t rank) {
  DevDense &     Q  = h->dense;
  const unsigned nm = static_cast<unsigned>(Q.nm), rk = dense_rank_of(Q, rank);
  if (!Q.transposed) {
    dense_mult_nt_kernel<<<1, kDenseT, (rk ? 2 * rk : 2) * sizeof(double), h->stream>>>(nm, rk, Q.R.p, Q.Q.p, Q.jpvt.p,
                                                                                       d_in, d_out);
  } else {
    if (rk) {
      dense_qt_kernel<<<cdiv(static_cast<std::size_t>(rk) * 32, 256), 256, 0, h->stream>>>(nm, rk, Q.Q.p, d_in, Q.c.p);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
    }
    dense_mult_t_kernel<<<1, kDenseT, (rk ? rk : 1) * sizeof(double), h->stream>>>(nm, rk, Q.R.p, Q.c.p, Q.jpvt.p, d_out);
  }
  HIF_KERNEL_CHECK();
  ++h->launch_count;
}

void launch_ldu_solve(Handle *h, DevLevel &D, const double *rhs, unsigned long long *xL, unsigned long long *xU,
                      unsigned parity, int *tickets) {
  launch_sweeps(h, D, rhs, xL, xU, parity, tickets, "prod.");
}

void launch_dense_trsv_cols(Handle *h, unsigned nm, unsigned rk, const double *c, double *out, unsigned ncols) {
  DevDense &Q = h->dense;
  dense_trsv_kernel<<<ncols, kTrsvThreads, (rk ? rk : 1) * sizeof(double), h->stream>>>(nm, rk, Q.R.p, Q.tinv.p, c,
                                                                                       Q.jpvt.p, out, ncols, nullptr);
  HIF_KERNEL_CHECK();
  ++h->launch_count;
}

void check_sweep_error(Handle *h) {
  HIF_CUDA(cudaMemcpyAsync(h->h_error, h->error_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  HIF_CUDA(cudaStreamSynchronize(h->stream));
  if (*h->h_error) {
    const int kind = *h->h_error;
    int       info[8] = {0};
    cudaMemcpy(info, h->error_flag.p, sizeof(info), cudaMemcpyDeviceToHost);
    HIF_CUDA(cudaMemsetAsync(h->error_flag.p, 0, 8 * sizeof(int), h->stream));
    reset_tagged_state(h);  // the aborted sweep left slots unwritten
    std::string where;
    if (info[1])  // the first failing wait left its coordinates (wsweep.cu, ws_fail)
      where = " [warp " + std::to_string(info[2]) + " segment " + std::to_string(info[3]) + " level " +
              std::to_string(info[4]) + " kind " + std::to_string(info[5]) + " detail " + std::to_string(info[6]) +
              " parity " + std::to_string(info[7]) + "]";
    throw std::runtime_error("a triangular sweep exceeded its spin limit (dependency never became ready; kind " +
                             std::to_string(kind) + ")" + where);
  }
}

}  // namespace hifgpu
