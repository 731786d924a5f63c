"""Developer analysis: fill of the super-level merge W^{-1} (levels grouped B at a time)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
import bench

def levels_of(csr):
    m = csr.shape[0]; ip, ix = csr.indptr, csr.indices
    lev = np.zeros(m, dtype=np.int32)
    # process in index order (lower triangular: deps have smaller index)
    for i in range(m):
        a, b = ip[i], ip[i+1]
        if b > a: lev[i] = lev[ix[a:b]].max() + 1
    return lev

def analyse(M, name):
    m = M.shape[0]
    t0 = time.time(); lev = levels_of(M); depth = lev.max() + 1
    cnt = np.bincount(lev)
    print(f"{name}: m={m} nnz={M.nnz} depth={depth}  ({time.time()-t0:.1f}s)")
    coo = M.tocoo()
    for B in (4, 8, 16, 32, 64):
        sl = lev // B
        within = sl[coo.row] == sl[coo.col]
        N = sp.csr_matrix((coo.data[within], (coo.row[within], coo.col[within])), shape=(m, m))
        nX = int((~within).sum())
        Mn = -N
        S = sp.identity(m, format="csr") + Mn
        P = Mn
        k = 1
        while k < B:
            P = P @ P
            if P.nnz == 0: break
            S = S + S @ P
            k *= 2
        S.eliminate_zeros()
        fill = S.nnz - m
        split = int((np.diff(S.indptr) > 1).sum())
        print(f"   B={B:3d} superlevels={int(sl.max())+1:4d} nnz(X)={nX:9d} nnz(N)={N.nnz:9d} nnz(Winv-I)={fill:10d} "
              f"ratio (X+Winv-I)/nnz={(nX+fill)/M.nnz:.2f} split rows={split} ({split/m:.2%}) maxrow={np.diff(S.indptr).max()}")

if __name__ == "__main__":
    wl, size = sys.argv[1], int(sys.argv[2])
    A, lv = bench.cached_levels(wl, size)
    for li, L in enumerate(lv):
        for nm, upper in (("L", False), ("U", True)):
            nr, nc, cs, ri, va = L[nm]
            if nr == 0 or len(ri) == 0: continue
            M = sp.csc_matrix((va, ri, cs), shape=(nr, nc)).tocsr()
            if upper:  # mirror: dependencies point to smaller indices
                M = M[::-1, ::-1].tocsr()
            M.sort_indices()
            analyse(M, f"level {li} {nm}")
