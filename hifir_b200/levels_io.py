"""(De)serialise per-level factor descriptions (the dict layout taken by
``hifir_b200.make_level_structs``) to a flat ``.npz`` -- used for the committed golden
fixtures and as an on-disk cache of factors (factorization of the 128^3 configs takes
~20 s on the host)."""
from __future__ import annotations

import numpy as np

_BLOCKS = ("L", "U", "E", "F")
_VECS = ("d", "s", "t", "p", "p_inv", "q", "q_inv")
_DENSE = ("qr_mat", "qr_tau", "qr_jpvt")


def levels_to_arrays(levels, prefix=""):
    out = {prefix + "nlevels": np.array(len(levels))}
    for k, L in enumerate(levels):
        pre = f"{prefix}lv{k}_"
        out[pre + "dims"] = np.array([L["m"], L["n"], L.get("dense_n", 0), L.get("dense_rank", 0),
                                      L.get("has_symm_dense", 0)], dtype=np.int64)
        for b in _BLOCKS:
            nr, nc, cs, ri, va = L[b]
            out[pre + b + "_shape"] = np.array([nr, nc], dtype=np.int64)
            out[pre + b + "_cs"] = np.asarray(cs, dtype=np.int64)
            out[pre + b + "_ri"] = np.asarray(ri, dtype=np.int32)
            out[pre + b + "_va"] = np.asarray(va)  # float64, or float32 for hif::HIF<float> factors
        for v in _VECS:
            out[pre + v] = np.asarray(L[v])
        if L.get("dense_n", 0):
            for v in _DENSE:
                out[pre + v] = np.asarray(L[v])
    return out


def arrays_to_levels(z, prefix=""):
    levels = []
    for k in range(int(z[prefix + "nlevels"])):
        pre = f"{prefix}lv{k}_"
        m, n, dn, dr, hs = (int(v) for v in z[pre + "dims"])
        L = dict(m=m, n=n, dense_n=dn, dense_rank=dr, has_symm_dense=hs)
        for b in _BLOCKS:
            nr, nc = (int(v) for v in z[pre + b + "_shape"])
            L[b] = (nr, nc, z[pre + b + "_cs"], z[pre + b + "_ri"], z[pre + b + "_va"])
        for v in _VECS:
            L[v] = z[pre + v]
        if dn:
            for v in _DENSE:
                L[v] = z[pre + v]
        levels.append(L)
    return levels


def save_levels(path, levels, **extra):
    arrs = levels_to_arrays(levels)
    arrs.update({k: np.asarray(v) for k, v in extra.items()})
    np.savez_compressed(path, **arrs)


def load_levels(path):
    """Returns (levels, extras dict)."""
    with np.load(path) as z:
        data = {k: z[k] for k in z.files}
    levels = arrays_to_levels(data)
    extra = {k: v for k, v in data.items() if not (k == "nlevels" or k.startswith("lv"))}
    return levels, extra
