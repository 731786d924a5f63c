// product.cu -- the multilevel matrix-vector product y = M x (LHF_M; with the transposed twin
// handle: LHF_MH), reference src/hif/alg/prec_prod.hpp:54-134 (and :148-230 through the twin).
//
// Per level (m, n, nm = n-m), with w = x[q] ./ t[q]:
//   y2       = (next level product, or dense Q R P^T) applied to w[m:n]
//   v1       = L (D (U w1))            with the implicit unit diagonals, w1 = w[0:m]
//   f        = F w[m:n] ;  v1 += f
//   z        = (L D U)^{-1} f + w1     (the level's one triangular solve: the sweep kernels)
//   v2       = E z + y2
//   y[i]     = [v1; v2][p_inv[i]] / s[i]
#include "hifgpu.h"

namespace hifgpu {

namespace {

inline unsigned cdiv(std::size_t a, std::size_t b) { return static_cast<unsigned>((a + b - 1) / b); }
constexpr int   T = 256;

// w[i] = x[q[i]] / t[q[i]]       (prec_prod.hpp:74, 96)
__global__ void gather_div_kernel(const unsigned n, const int *__restrict__ q, const double *__restrict__ t,
                                  const double *__restrict__ x, double *__restrict__ w) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int qi = q[i];
    w[i]         = x[qi] / t[qi];
  }
}

// out[i] = ((sum_j A(i,j) x[j]) + (UNIT ? x[i] : 0) + (add ? add[i] : 0)) * (scale ? scale[i] : 1)
// 4 lanes per row; covers U*w (+unit diagonal, *d), L*y (+unit diagonal), F*w2, E*z + y2
template <bool UNIT, bool TAGX>
__global__ void spmv_axpby_kernel(const unsigned nrows, const unsigned *__restrict__ ptr, const int *__restrict__ col,
                                  const double *__restrict__ val, const void *__restrict__ xin,
                                  const double *__restrict__ add, const double *__restrict__ scale,
                                  double *__restrict__ out) {
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x, row = gid >> 2, lane = gid & 3u;
  double         acc = 0.0;
  auto           X   = [&](unsigned j) {
    return TAGX ? tag_value(static_cast<const unsigned long long *>(xin)[j]) : static_cast<const double *>(xin)[j];
  };
  if (row < nrows) {
    const unsigned e = ptr[row + 1];
    for (unsigned k = ptr[row] + lane; k < e; k += 4) acc = fma(val[k], X(col[k]), acc);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (row < nrows && lane == 0) {
    if (UNIT) acc += X(row);
    if (add) acc += add[row];
    if (scale) acc *= scale[row];
    out[row] = acc;
  }
}

__global__ void vec_add_kernel(const unsigned n, const double *a, const double *b, double *out) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// z[i] = value(xU[i]) + w1[i]      (prec_prod.hpp:124)
// (row i of the U sweep's result lives at slot[i])
__global__ void add_tagged_kernel(const unsigned m, const unsigned long long *__restrict__ xU,
                                  const unsigned *__restrict__ slot, const double *__restrict__ w1,
                                  double *__restrict__ z) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) z[i] = tag_value(xU[slot[i]]) + w1[i];
}

// y[i] = [v1; v2][p_inv[i]] / s[i]   (prec_prod.hpp:133)
__global__ void scatter_div_kernel(const unsigned n, const unsigned m, const int *__restrict__ p_inv,
                                   const double *__restrict__ s, const double *__restrict__ v1,
                                   const double *__restrict__ v2, double *__restrict__ y) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const unsigned j = static_cast<unsigned>(p_inv[i]);
    y[i]             = (j < m ? v1[j] : v2[j - m]) / s[i];
  }
}

void upload_host_csr(const HostCsr &hc, DevCsr &d, std::size_t *tally) {
  d.nrows = hc.nrows;
  d.ncols = hc.ncols;
  d.nnz   = hc.col.size();
  d.ptr.upload(hc.ptr, tally);
  d.col.upload(hc.col, tally);
  d.val.upload(hc.val, tally);
}

void ensure_prod(Handle *h) {
  if (h->prod_ready) return;
  std::size_t *tally = &h->device_bytes;
  for (DevLevel &D : h->levels) {
    if (D.h_pinv.empty() || D.h_q.empty())
      throw std::logic_error("the multilevel product needs p_inv and q of every level (not given at attach)");
    upload_host_csr(D.hostL, D.Lcsr, tally);
    upload_host_csr(D.hostU, D.Ucsr, tally);
    D.q_dev.upload(D.h_q, tally);
    D.pinv_dev.upload(D.h_pinv, tally);
    D.pw.alloc(D.n, tally);
    D.pw1.alloc(D.m, tally);
    D.pw2.alloc(D.nm, tally);
    D.pf.alloc(D.m, tally);
    D.py.alloc(D.n, tally);
    D.p_xL.alloc(2 * D.m, tally);
    D.p_xU.alloc(2 * D.m, tally);
  }
  h->prod_ready = true;
}

template <bool UNIT, bool TAGX>
void spmv(Handle *h, const DevCsr &A, const void *x, const double *add, const double *scale, double *out) {
  if (!A.nrows) return;
  spmv_axpby_kernel<UNIT, TAGX><<<cdiv(A.nrows * 4, T), T, 0, h->stream>>>(static_cast<unsigned>(A.nrows), A.ptr.p,
                                                                          A.col.p, A.val.p, x, add, scale, out);
  HIF_KERNEL_CHECK();
  ++h->launch_count;
}

void prod_level(Handle *h, std::size_t l, const double *x, double *y, std::size_t rank, unsigned parity) {
  DevLevel &     D = h->levels[l];
  const unsigned n = static_cast<unsigned>(D.n), m = static_cast<unsigned>(D.m);
  if (!n) return;
  gather_div_kernel<<<cdiv(n, T), T, 0, h->stream>>>(n, D.q_dev.p, D.t.p, x, D.pw.p);
  HIF_KERNEL_CHECK();
  ++h->launch_count;
  const double *w1 = D.pw.p, *w2 = D.pw.p + m;
  double *      y2 = D.py.p + m;  // product of the trailing block
  const bool    last = l + 1 == h->levels.size();
  if (D.nm) {
    if (last)
      dense_multiply_dev(h, w2, y2, rank);
    else
      prod_level(h, l + 1, w2, y2, rank, parity);
  }
  double *v1 = D.pw1.p;
  if (m) {
    // t1 = D (U w1 + w1) ; v1 = L t1 + t1        (prec_prod.hpp:100-107); t1 lives in py[0:m]
    spmv<true, false>(h, D.Ucsr, w1, nullptr, D.d.p, D.py.p);
    spmv<true, false>(h, D.Lcsr, D.py.p, nullptr, nullptr, v1);
    if (D.nm) {
      // f = F w2 ; v1 += f                         (:111-114)
      spmv<false, false>(h, D.F, w2, nullptr, nullptr, D.pf.p);
      vec_add_kernel<<<cdiv(m, T), T, 0, h->stream>>>(m, v1, D.pf.p, v1);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
      // z = (LDU)^{-1} f + w1                       (:120-124); z lives in pf
      launch_ldu_solve(h, D, D.pf.p, D.p_xL.p, D.p_xU.p, parity, h->tick(8 * l));
      add_tagged_kernel<<<cdiv(m, T), T, 0, h->stream>>>(m, D.p_xU.p, D.U_slot.p, w1, D.pf.p);
      HIF_KERNEL_CHECK();
      ++h->launch_count;
    }
  }
  double *v2 = D.pw2.p;
  if (D.nm) {
    if (m)
      spmv<false, false>(h, D.E, D.pf.p, y2, nullptr, v2);  // v2 = E z + y2   (:126-128)
    else
      HIF_CUDA(cudaMemcpyAsync(v2, y2, D.nm * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  scatter_div_kernel<<<cdiv(n, T), T, 0, h->stream>>>(n, m, D.pinv_dev.p, D.s.p, v1, v2, y);
  HIF_KERNEL_CHECK();
  ++h->launch_count;
}

}  // namespace

void prod_dev(Handle *h, const double *d_x, double *d_y, std::size_t rank) {
  HIF_CUDA(cudaSetDevice(h->device));
  ensure_prod(h);
  ++h->epoch_p;
  HIF_CUDA(cudaMemsetAsync(h->tickets.p, 0, h->tickets.n * sizeof(int), h->stream));
  prod_level(h, 0, d_x, d_y, rank, h->epoch_p & 1u);
}

}  // namespace hifgpu
