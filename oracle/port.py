"""ctypes front-end of oracle/_build/liboracle.so (oracle/hif_oracle.c, the plain-C
restatement of the reference apply path).  TEST INFRASTRUCTURE: imported only by tests/,
bench.py's cpu_baseline/reference legs and __graft_entry__.smoke()."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from hifir_b200 import LhfdGpuCcs, make_level_structs

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
FULL_RANK = (1 << 64) - 1

_lib = None


class _Nsp(C.Structure):
    _fields_ = [("enabled", C.c_int), ("start", C.c_size_t), ("end", C.c_size_t)]


class _Crs(C.Structure):
    _fields_ = [("n", C.c_size_t), ("row_start", C.c_void_p), ("col_ind", C.c_void_p), ("vals", C.c_void_p)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "_build/liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.hif_oracle_norm2.restype = C.c_double
        L.hif_oracle_work_size.restype = C.c_size_t
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleHif:
    """The C port applied to a list of per-level factor dicts (same layout the device
    backend is attached with)."""

    def __init__(self, levels, A=None):
        self._arr, self._keep = make_level_structs(levels)
        self.nl = len(levels)
        self.n = levels[0]["n"]
        self._nsp = _Nsp(0, 0, 0)
        self._A = None
        if A is not None:
            self.set_matrix(A)

    def set_matrix(self, A):
        n, ip, ix, va = A
        self._Aarr = (np.ascontiguousarray(ip, dtype=np.int64), np.ascontiguousarray(ix, dtype=np.int32),
                      np.ascontiguousarray(va, dtype=np.float64))
        self._A = _Crs(n, _p(self._Aarr[0]), _p(self._Aarr[1]), _p(self._Aarr[2]))

    def set_nsp_const(self, start=0, end=FULL_RANK):
        self._nsp = _Nsp(1, start, end)

    def clear_nsp(self):
        self._nsp = _Nsp(0, 0, 0)

    def solve(self, b, rank=0):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        rc = lib().hif_oracle_solve(C.c_size_t(self.nl), self._arr, _p(b), C.c_size_t(rank), C.byref(self._nsp),
                                    _p(x))
        assert rc == 0
        return x

    def apply_op(self, op, b, rank=0):
        """op 1 = S^H, 2 = M, 3 = M^H (prec_solve_tran / prec_prod / prec_prod_tran)"""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        rc = lib().hif_oracle_apply_op(C.c_size_t(self.nl), self._arr, C.c_int(op), _p(b), C.c_size_t(rank), _p(x))
        assert rc == 0
        return x

    def solve_mrhs(self, B, rank=0):
        """multi-RHS semantics = a loop of single solves"""
        return np.stack([self.solve(np.ascontiguousarray(B[:, k]), rank) for k in range(B.shape[1])], axis=1)

    def hifir(self, b, nirs, rank=FULL_RANK):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        rc = lib().hif_oracle_hifir(C.c_size_t(self.nl), self._arr, C.byref(self._A), _p(b), C.c_size_t(nirs),
                                    C.c_size_t(rank), C.byref(self._nsp), _p(x))
        assert rc == 0
        return x

    def hifir_betas(self, b, nirs, betas, rank=FULL_RANK):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        bt = np.asarray(betas, dtype=np.float64)
        out = np.zeros(2, dtype=np.int64)
        rc = lib().hif_oracle_hifir_betas(C.c_size_t(self.nl), self._arr, C.byref(self._A), _p(b),
                                          C.c_size_t(nirs), _p(bt), C.c_size_t(rank), C.byref(self._nsp), _p(x),
                                          _p(out))
        assert rc == 0
        return x, int(out[0]), int(out[1])

    def krylov(self, b, which="fgmres", restart=30, rtol=1e-6, maxit=500, full_rank=False):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        out = np.zeros(3, dtype=np.int32)
        rc = lib().hif_oracle_krylov(C.c_size_t(self.nl), self._arr, C.byref(self._A), _p(b),
                                     C.c_int(1 if which == "fgmres" else 0), C.c_int(restart), C.c_double(rtol),
                                     C.c_int(maxit), C.c_int(int(full_rank)), C.byref(self._nsp), _p(x), _p(out))
        assert rc == 0
        return x, int(out[0]), int(out[1]), int(out[2])

    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        lib().hif_oracle_spmv(C.byref(self._A), _p(x), _p(y))
        return y


def norm2(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().hif_oracle_norm2(_p(v), C.c_size_t(v.size))


def qrcp_solve(mat, tau, jpvt, num_rank, b, rank=0):
    nm = len(tau)
    x = np.array(b, dtype=np.float64)
    mat = np.ascontiguousarray(mat, dtype=np.float64)
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    jpvt = np.ascontiguousarray(jpvt, dtype=np.int32)
    lib().hif_oracle_qrcp_solve(C.c_size_t(nm), C.c_size_t(num_rank), _p(mat), _p(tau), _p(jpvt),
                                C.c_size_t(rank), _p(x))
    return x


def ccs_to_crs(block):
    nr, nc, cs, ri, va = block
    cs = np.ascontiguousarray(cs, dtype=np.int64)
    ri = np.ascontiguousarray(ri, dtype=np.int32)
    va = np.ascontiguousarray(va, dtype=np.float64)
    c = LhfdGpuCcs(nr, nc, _p(cs) if cs.size else None, _p(ri) if ri.size else None, _p(va) if va.size else None)
    rs = np.zeros(nr + 1, dtype=np.int64)
    ci = np.zeros(len(ri), dtype=np.int32)
    v = np.zeros(len(ri), dtype=np.float64)
    lib().hif_oracle_ccs_to_crs(C.byref(c), _p(rs), _p(ci), _p(v))
    return rs, ci, v
