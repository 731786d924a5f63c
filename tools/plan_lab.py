"""Developer tool (host only): gather locality of candidate row orders / slot numberings.

    python tools/plan_lab.py poisson 128 [level ...]

Prints, per factor and configuration, L2 sector requests per stored entry (what bounds the
streaming sweep: one sector request per SM per clock), chunk-scope distinct sectors (what an L1
could capture) and the ELL padding."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
import hifir_b200 as hb


def lab(block, upper, key, opts):
    nr, nc, cs, ri, va = block
    cs = np.ascontiguousarray(cs, dtype=np.int64)
    ri = np.ascontiguousarray(ri, dtype=np.int32)
    va = np.ascontiguousarray(va, dtype=np.float64)
    c = hb.LhfdGpuCcs(nr, nc, hb._ptr(cs), hb._ptr(ri), hb._ptr(va))
    o = np.zeros(8)
    o[:len(opts)] = opts
    out = np.zeros(16)
    k = np.ascontiguousarray(key, dtype=np.int32) if key is not None else None
    f = hb.lib().lhfdGpuDebugPlanLab
    f.argtypes = [C.c_void_p] * 1 + [C.c_int] + [C.c_void_p] * 3
    st = f(C.byref(c), int(upper), hb._ptr(k) if k is not None else None, hb._ptr(o), hb._ptr(out))
    assert st == 0, hb.lib().lhfGpuGetErrorMsg()
    return out


CONFIGS_ALL = [
    ("r1: (lev,len) orig slots", (0, 0, 0, 0)),
    ("(lev,cls,key) sweep-index key, lvl-major slots", (1, 1, 0, 0)),
    ("(lev,cls,key) user key p, lvl-major slots", (1, 1, 1, 0)),
    ("(lev,cls,key) user key p, packed slots", (1, 2, 1, 0)),
    ("(lev,cls,bucket,key) user key p, lvl-major slots", (2, 1, 1, 0)),
    ("barycenter x1 seeded p, lvl-major slots", (1, 1, 3, 1)),
    ("barycenter x3 seeded p, lvl-major slots", (1, 1, 3, 3)),
    ("barycenter x3 seeded sweep idx, lvl-major slots", (1, 1, 2, 3)),
    ("barycenter x3 seeded p, packed slots", (1, 2, 3, 3)),
]


CONFIGS = CONFIGS_ALL if os.environ.get("PLAN_LAB_ALL") else CONFIGS_ALL[:2]


def main():
    wl, size = sys.argv[1], int(sys.argv[2])
    which = [int(a) for a in sys.argv[3:]] or None
    A, lv = bench.cached_levels(wl, size)
    for li, L in enumerate(lv):
        if which is not None and li not in which:
            continue
        m = L["m"]
        p = np.asarray(L["p"])[:m]
        q = np.asarray(L["q"])[:m] if L.get("q") is not None else np.argsort(np.asarray(L["q_inv"]))[:m]
        for nm, upper, key in (("L", False, p), ("U", True, q)):
            blk = L[nm]
            if blk[0] == 0 or len(blk[3]) == 0:
                continue
            for name, opts in CONFIGS:
                t0 = time.time()
                o = lab(blk, upper, key, opts)
                ent = o[1]
                print(f"lv{li} {nm} [{name}]: rows {int(o[0])} entries {int(ent)} pad {o[2] / ent:.2%} depth {int(o[7])} "
                      f"sectors/entry {o[4] / ent:.3f} lines/entry {o[5] / ent:.3f} chunk-sectors/entry {o[6] / ent:.3f} "
                      f"pub sectors/row {o[8] / o[0]:.3f} rhs sectors/row {o[9] / o[0]:.3f}  ({time.time() - t0:.1f}s)",
                      flush=True)
                if name.startswith("r1"):
                    print(f"      fan-out: most references to one row {int(o[10])}; entries on rows referenced >= 32x "
                          f"{o[11] / ent:.1%}, >= 256x {o[12] / ent:.1%} ({int(o[13])} such rows); entries at level "
                          f"distance <= 1 {o[14] / ent:.1%}, <= 4 {o[15] / ent:.1%}", flush=True)


if __name__ == "__main__":
    main()
