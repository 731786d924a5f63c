"""ctypes front-end of oracle/_ref/libhifir_ref.so (the UNMODIFIED reference compiled
by oracle/Makefile).  TEST INFRASTRUCTURE: imported only by tests/, bench.py (factor
producer + cpu_baseline leg + reference arm) and __graft_entry__.smoke().  The product
package hifir_b200/ never imports this module.

What it gives:
  * RefHif(A, params).factorize  -> hif::HIF<double,int>::factorize (builder.hpp:263-366)
  * .levels()                    -> per-level factors verbatim from hif::Prec (Prec.hpp:309-323)
  * .solve/.hifir/.krylov        -> the reference's CPU path, the parity oracle proper
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libhifir_ref.so")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                f"{LIB_PATH} missing: run `make -C oracle` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        L.hifref_last_error.restype = C.c_char_p
        L.hifref_version.restype = C.c_char_p
        L.hifref_create.restype = C.c_void_p
        L.hifref_create.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hifref_destroy.argtypes = [C.c_void_p]
        L.hifref_num_precs.argtypes = [C.c_void_p]
        L.hifref_levels.restype = C.c_size_t
        L.hifref_levels.argtypes = [C.c_void_p]
        L.hifref_nnz.restype = C.c_size_t
        L.hifref_nnz.argtypes = [C.c_void_p]
        L.hifref_level_sizes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hifref_export_ccs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hifref_block_shape.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.hifref_export_vectors.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.hifref_export_dense.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hifref_set_nsp_const.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.hifref_clear_nsp.argtypes = [C.c_void_p]
        L.hifref_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.hifref_mmultiply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.hifref_hifir.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.hifref_hifir_betas.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                         C.c_size_t, C.c_void_p]
        L.hifref_spmv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hifref_krylov.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                    C.c_void_p, C.c_void_p]
        L.hifref_norm2.restype = C.c_double
        L.hifref_norm2.argtypes = [C.c_void_p, C.c_size_t]
        L.hifref_gpu_attach.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        # single-precision preconditioner (hif::HIF<float,int>), same bridge with float factor arrays
        L.hifrefs_create.restype = C.c_void_p
        L.hifrefs_create.argtypes = L.hifref_create.argtypes
        L.hifrefs_destroy.argtypes = [C.c_void_p]
        L.hifrefs_num_precs.argtypes = [C.c_void_p]
        L.hifrefs_levels.restype = C.c_size_t
        L.hifrefs_levels.argtypes = [C.c_void_p]
        L.hifrefs_nnz.restype = C.c_size_t
        L.hifrefs_nnz.argtypes = [C.c_void_p]
        L.hifrefs_level_sizes.argtypes = L.hifref_level_sizes.argtypes
        L.hifrefs_export_ccs.argtypes = L.hifref_export_ccs.argtypes
        L.hifrefs_block_shape.argtypes = L.hifref_block_shape.argtypes
        L.hifrefs_export_vectors.argtypes = L.hifref_export_vectors.argtypes
        L.hifrefs_export_dense.argtypes = L.hifref_export_dense.argtypes
        L.hifrefs_set_nsp_const.argtypes = L.hifref_set_nsp_const.argtypes
        L.hifrefs_clear_nsp.argtypes = [C.c_void_p]
        L.hifrefs_apply_op.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.hifrefs_solve_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.hifrefs_hifir.argtypes = L.hifref_hifir.argtypes
        L.hifrefs_gpu_attach.argtypes = L.hifref_gpu_attach.argtypes
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


FULL_RANK = (1 << 64) - 1  # size_t(-1)


class RefHif:
    """A factorized hif::HIF<double,int> -- or, with dtype=np.float32, hif::HIF<float,int> (float
    factors, double user matrix and vectors: the reference's mixed-precision path) -- living in the
    reference library."""

    def __init__(self, A, params=None, dense_thres=0, threads=0, verbose=0, dtype=np.float64):
        n, indptr, indices, vals = A
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype(np.float64), np.dtype(np.float32))
        self._pfx = "hifref_" if self.dtype == np.float64 else "hifrefs_"
        self.n = int(n)
        self._A = (np.ascontiguousarray(indptr, dtype=np.int64),
                   np.ascontiguousarray(indices, dtype=np.int32),
                   np.ascontiguousarray(vals, dtype=np.float64))
        pr = np.full(8, -1.0)
        if params:
            for k, name in enumerate(("tau_L", "tau_U", "kappa_d", "kappa", "alpha_L", "alpha_U")):
                if name in params:
                    pr[k] = params[name]
        pr[6] = dense_thres
        pr[7] = threads
        self._h = self._f("create")(self.n, _p(self._A[0]), _p(self._A[1]), _p(self._A[2]), _p(pr),
                                      int(verbose))
        if not self._h:
            raise RuntimeError("reference factorize failed: " + lib().hifref_last_error().decode())

    def _f(self, name):
        return getattr(lib(), self._pfx + name)

    def close(self):
        if getattr(self, "_h", None):
            self._f("destroy")(self._h)
            self._h = None

    __del__ = close

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("reference error: " + lib().hifref_last_error().decode())

    @property
    def num_precs(self):
        return self._f("num_precs")(self._h)

    @property
    def num_levels(self):
        return self._f("levels")(self._h)

    @property
    def nnz(self):
        return self._f("nnz")(self._h)

    def level_sizes(self, lvl):
        s = np.zeros(10, dtype=np.uint64)
        self._chk(self._f("level_sizes")(self._h, lvl, _p(s)))
        keys = ("m", "n", "nnz_L", "nnz_U", "nnz_E", "nnz_F", "dense_n", "dense_rank", "has_symm_dense",
                "is_last")
        return {k: int(v) for k, v in zip(keys, s)}

    def levels(self):
        """Per-level factors, copied verbatim: list of dicts with CCS blocks
        ('L','U','E','F' -> (nrows, ncols, col_start[int64], row_ind[int32], vals)), 'd','s','t',
        'p','p_inv','q','q_inv', and on the last level 'qr_mat' (column-major, flattened),
        'qr_tau','qr_jpvt' (1-based), 'dense_n','dense_rank'."""
        out = []
        for lvl in range(self.num_precs):
            sz = self.level_sizes(lvl)
            m, n = sz["m"], sz["n"]
            L = dict(m=m, n=n, has_symm_dense=sz["has_symm_dense"])
            for which, (name, nnzk) in enumerate((("L", "nnz_L"), ("U", "nnz_U"), ("E", "nnz_E"), ("F", "nnz_F"))):
                dims = np.zeros(2, dtype=np.uint64)
                self._chk(self._f("block_shape")(self._h, lvl, which, _p(dims)))
                nr, nc = int(dims[0]), int(dims[1])
                cs = np.zeros(nc + 1, dtype=np.int64)
                ri = np.zeros(sz[nnzk], dtype=np.int32)
                va = np.zeros(sz[nnzk], dtype=self.dtype)
                self._chk(self._f("export_ccs")(self._h, lvl, which, _p(cs), _p(ri), _p(va)))
                L[name] = (nr, nc, cs, ri, va)
            d = np.zeros(m, dtype=self.dtype); s = np.zeros(n, dtype=self.dtype); t = np.zeros(n, dtype=self.dtype)
            p, p_inv, q, q_inv = (np.zeros(n, dtype=np.int32) for _ in range(4))
            self._chk(self._f("export_vectors")(self._h, lvl, _p(d), _p(s), _p(t), _p(p), _p(p_inv), _p(q),
                                                  _p(q_inv)))
            L.update(d=d, s=s, t=t, p=p, p_inv=p_inv, q=q, q_inv=q_inv)
            nm = sz["dense_n"]
            L["dense_n"], L["dense_rank"] = nm, sz["dense_rank"]
            if nm:
                mat = np.zeros(nm * nm, dtype=self.dtype); tau = np.zeros(nm, dtype=self.dtype)
                jp = np.zeros(nm, dtype=np.int32)
                self._chk(self._f("export_dense")(self._h, lvl, _p(mat), _p(tau), _p(jp)))
                L.update(qr_mat=mat, qr_tau=tau, qr_jpvt=jp)
            out.append(L)
        return out

    def set_nsp_const(self, start=0, end=FULL_RANK):
        self._chk(self._f("set_nsp_const")(self._h, start, end))

    def clear_nsp(self):
        self._f("clear_nsp")(self._h)

    def solve(self, b, rank=0):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        if self.dtype == np.float32:
            self._chk(lib().hifrefs_apply_op(self._h, 0, _p(b), _p(x), rank))
        else:
            self._chk(lib().hifref_solve(self._h, _p(b), _p(x), rank))
        return x

    def solve_f32(self, b, rank=0):
        """lhfsSolve: float factors AND float vectors (single-precision handles only)"""
        assert self.dtype == np.float32
        b = np.ascontiguousarray(b, dtype=np.float32)
        x = np.empty_like(b)
        self._chk(lib().hifrefs_solve_f32(self._h, _p(b), _p(x), rank))
        return x

    def mmultiply(self, x, rank=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        if self.dtype == np.float32:
            self._chk(lib().hifrefs_apply_op(self._h, 2, _p(x), _p(y), rank))
        else:
            self._chk(lib().hifref_mmultiply(self._h, _p(x), _p(y), rank))
        return y

    def apply_op(self, op, b, rank=0):
        """op 1 = S^H (solve trans), 2 = M (mmultiply), 3 = M^H -- LhfOperationType values"""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        f = self._f("apply_op")
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        self._chk(f(self._h, op, _p(b), _p(x), rank))
        return x

    def hifir(self, b, nirs, rank=FULL_RANK):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        self._chk(self._f("hifir")(self._h, _p(b), nirs, _p(x), rank))
        return x

    def hifir_betas(self, b, nirs, betas, rank=FULL_RANK):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        bt = np.asarray(betas, dtype=np.float64)
        out = np.zeros(2, dtype=np.int64)
        self._chk(lib().hifref_hifir_betas(self._h, _p(b), nirs, _p(bt), _p(x), rank, _p(out)))
        return x, int(out[0]), int(out[1])

    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self._chk(lib().hifref_spmv(self._h, _p(x), _p(y)))
        return y

    def krylov(self, b, which="fgmres", restart=30, rtol=1e-6, maxit=500):
        """gmres_hif / fgmres_hifir of examples/advanced/gmres.hpp, unmodified.
        Returns x, flag, iters, num_mv."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        out = np.zeros(3, dtype=np.int32)
        self._chk(lib().hifref_krylov(self._h, 1 if which == "fgmres" else 0, _p(b), restart, rtol, maxit,
                                      _p(x), _p(out)))
        return x, int(out[0]), int(out[1]), int(out[2])

    def gpu_attach(self, attach_levels_fnptr, device=0, attach_levels_s_fnptr=None):
        """Attach the device backend through the C++ adapter include/hifir_b200.hpp,
        i.e. the way a hif::HIF user would.  Returns the raw LhfdGpuHdl / LhfsGpuHdl (int)."""
        api = (C.c_void_p * 2)(C.cast(attach_levels_fnptr, C.c_void_p),
                               C.cast(attach_levels_s_fnptr, C.c_void_p) if attach_levels_s_fnptr else None)
        out = C.c_void_p()
        rc = self._f("gpu_attach")(self._h, api, device, C.byref(out))
        if rc != 0:
            raise RuntimeError(f"gpu attach failed ({rc}): " + lib().hifref_last_error().decode())
        return out.value


def norm2(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().hifref_norm2(_p(v), v.size)


def qrcp_factor_solve(a_rowmajor, b):
    """hif::QRCP<double>: factorize a dense matrix, export (mat, tau, jpvt, rank) and the
    reference's own solution of b."""
    a = np.ascontiguousarray(a_rowmajor, dtype=np.float64)
    n = a.shape[0]
    b = np.ascontiguousarray(b, dtype=np.float64)
    mat = np.zeros(n * n); tau = np.zeros(n); jp = np.zeros(n, dtype=np.int32); x = np.zeros(n)
    rank = C.c_size_t()
    f = lib().hifref_qrcp_factor_solve
    f.argtypes = [C.c_size_t] + [C.c_void_p] * 7
    if f(n, _p(a), _p(b), _p(mat), _p(tau), _p(jp), C.byref(rank), _p(x)) != 0:
        raise RuntimeError(lib().hifref_last_error().decode())
    return mat, tau, jp, rank.value, x


def export_crs(M, lvl, which, nrows, nnz):
    """the reference's own CRS(CCS) conversion of block `which` (0=L,1=U,2=E,3=F) of level lvl"""
    rs = np.zeros(nrows + 1, dtype=np.int64); ci = np.zeros(nnz, dtype=np.int32); va = np.zeros(nnz)
    f = lib().hifref_export_crs
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    if f(M._h, lvl, which, _p(rs), _p(ci), _p(va)) != 0:
        raise RuntimeError(lib().hifref_last_error().decode())
    return rs, ci, va
