"""Developer micro-benchmark (GPU box): dependent-step latency of the triangular sweep on a
synthetic pure chain (row i depends on row i-1 [+ `extra` far, long-finished rows])."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def chain_level(m, extra=0, far=5000):
    rows, cols = [], []
    i = np.arange(1, m)
    rows.append(i); cols.append(i - 1)
    for k in range(extra):
        ii = np.arange(far + k * 37, m)
        rows.append(ii); cols.append(ii - far - k * 37)
    rows = np.concatenate(rows); cols = np.concatenate(cols)
    order = np.lexsort((rows, cols))
    rows, cols = rows[order], cols[order]
    cs = np.zeros(m + 1, dtype=np.int64)
    np.add.at(cs, cols + 1, 1)
    cs = np.cumsum(cs)
    vals = np.full(len(rows), -0.3 / (1 + extra))
    empty = (m, m, np.zeros(m + 1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0))
    ident = np.arange(m, dtype=np.int32)
    return dict(m=m, n=m, L=(m, m, cs, rows.astype(np.int32), vals), U=empty,
                E=(0, m, np.zeros(m + 1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0)),
                F=(m, 0, np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0)),
                d=np.ones(m), s=np.ones(m), t=np.ones(m), p=ident, p_inv=ident, q=ident, q_inv=ident,
                dense_n=0, dense_rank=0)


def main():
    import torch

    import hifir_b200 as hb
    from hifir_b200 import build
    build.build()
    m = 60000
    for extra in (0, 6):
        lv = chain_level(m, extra)
        G = hb.GpuHif([lv])
        b = torch.ones(m, dtype=torch.float64, device="cuda")
        x = torch.empty_like(b)
        for _ in range(3):
            G.solve_dev(b.data_ptr(), x.data_ptr())
        G.synchronize()
        t = dict(G.profile_solve_dev(b.data_ptr(), x.data_ptr()))
        # reference: forward substitution on the CPU
        nr, nc, cs, ri, va = lv["L"]
        ref = np.ones(m)
        import scipy.sparse as sp
        Lc = sp.csc_matrix((va, ri, cs), shape=(m, m)).tocsr()
        for i in range(m):
            s, e = Lc.indptr[i], Lc.indptr[i + 1]
            ref[i] -= Lc.data[s:e] @ ref[Lc.indices[s:e]]
        err = np.abs(x.cpu().numpy() - ref).max()
        print(f"T={os.environ.get('HIFIR_B200_ROW_THREADS', 'dflt')} backoff={os.environ.get('HIFIR_B200_BACKOFF', '1')} "
              f"chain m={m} extra={extra}: L sweep {t['lv0.up.L'] * 1e3:.0f} us = {t['lv0.up.L'] * 1e6 / m:.1f} ns/row, "
              f"err {err:.1e}")
        G.close()


if __name__ == "__main__":
    main()
