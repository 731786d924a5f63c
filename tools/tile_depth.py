"""Developer analysis: dependency depth of a triangular factor at row level and at tile level
(tiles of T consecutive rows whose diagonal block is inverted at attach time)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
import bench

def tile_depth(csr, T, upper=False):
    m = csr.shape[0]
    indptr, idx = csr.indptr, csr.indices
    rows = np.repeat(np.arange(m), np.diff(indptr))
    if upper:  # mirror so that dependencies point to smaller indices
        rows, cols = m - 1 - rows, m - 1 - idx
    else:
        cols = idx
    tr, tc = rows // T, cols // T
    off = tr != tc
    # unique tile edges
    key = np.unique(tr[off].astype(np.int64) * ((m + T - 1) // T) + tc[off])
    nt = (m + T - 1) // T
    etr, etc = key // nt, key % nt
    order = np.argsort(etr, kind="stable")
    etr, etc = etr[order], etc[order]
    ptr = np.searchsorted(etr, np.arange(nt + 1))
    depth = np.zeros(nt, dtype=np.int32)
    for t in range(nt):
        a, b = ptr[t], ptr[t + 1]
        if b > a:
            depth[t] = depth[etc[a:b]].max() + 1
    return depth, len(key)

if __name__ == "__main__":
    wl, size = sys.argv[1], int(sys.argv[2])
    A, lv = bench.cached_levels(wl, size)
    for li, L in enumerate(lv):
        for name, upper in (("L", False), ("U", True)):
            nr, nc, cs, ri, va = L[name]
            if nr == 0 or len(ri) == 0: continue
            M = sp.csc_matrix((va, ri, cs), shape=(nr, nc)).tocsr()
            print(f"level {li} {name}: m={nr} nnz={M.nnz}")
            for T in (1, 8, 16, 32, 64, 128, 256, 512, 1024):
                d, ne = tile_depth(M, T, upper)
                print(f"   T={T:5d} tiles={len(d):8d} depth={d.max()+1:5d} tile-edges={ne:9d} ({ne/len(d):.1f}/tile)")
