"""Developer tool (GPU box): per-chunk timeline of the streaming sweeps (stream.cu).

    python tools/trace_stream.py --size 128 --out gpurun_out/stream_trace.npz

record per chunk: [0] ticket taken, [1] factor entries in registers, [2] admitted,
[3] last warp published, [4] level set, [5] SM id (times in ns, globaltimer)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def summarize(name, t):
    t = t[t[:, 0] > 0].astype(np.int64)
    if not len(t):
        return
    t0 = t[:, 0].min()
    tk, pf, ad, dn, lv = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, (t[:, 3] - t0) / 1e3, t[:, 4]
    nl = int(lv.max()) + 1
    g1, g2 = (t[:, 6] - t0) / 1e3, (t[:, 7] - t0) / 1e3
    print(f"{name}: chunks {len(t)} levels {nl} sweep {dn.max():.1f} us | prefetch mean {np.mean(pf - tk):.2f} us, "
          f"admission wait mean {np.mean(ad - pf):.2f}, admitted->done mean {np.mean(dn - ad):.2f} max {np.max(dn - ad):.2f}"
          f" | warp0: admitted->gathered {np.mean(g1 - ad):.2f}, gathered->consumed {np.mean(g2 - g1):.2f}, "
          f"consumed->last publish {np.mean(dn - g2):.2f}")
    # per level: first admission, last done
    first_ad = np.full(nl, np.inf); last_dn = np.zeros(nl); cnt = np.zeros(nl, dtype=int)
    np.minimum.at(first_ad, lv, ad); np.maximum.at(last_dn, lv, dn); np.add.at(cnt, lv, 1)
    last_ad = np.zeros(nl); np.maximum.at(last_ad, lv, ad)
    step = np.diff(last_dn)
    print(f"   level completion step: mean {step.mean():.2f} us median {np.median(step):.2f} max {step.max():.2f}")
    rows = []
    for l in list(range(min(nl, 6))) + list(range(max(6, nl // 2 - 2), min(nl, nl // 2 + 2))) + list(range(max(6, nl - 5), nl)):
        rows.append(f"     L{l}: chunks {cnt[l]} first_adm {first_ad[l]:.1f} last_adm {last_ad[l]:.1f} done {last_dn[l]:.1f}")
    print("\n".join(rows))


def main():
    import torch

    import hifir_b200 as hb
    from bench import factorize, make_problem
    from hifir_b200 import build, problems as P

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="poisson")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--out", default="gpurun_out/stream_trace.npz")
    args = ap.parse_args()
    build.build()
    A = make_problem(args.workload, args.size)
    M = factorize(A, threads=os.cpu_count() or 1)
    levels = M.levels()
    G = hb.GpuHif(levels)
    b = torch.from_numpy(P.seeded_rhs(A[0], 0)).cuda()
    x = torch.empty_like(b)
    for _ in range(3):
        G.solve_dev(b.data_ptr(), x.data_ptr())
    G.synchronize()
    out = {}
    names = ("downL", "downU", "upL", "upU")
    for lvl in range(len(levels)):
        if not levels[lvl]["m"]:
            continue
        for which in range(2):
            t = G.trace_sweep(b.data_ptr(), x.data_ptr(), lvl, which, max_blocks=1 << 20)
            out[f"lv{lvl}_{names[which]}"] = t
            summarize(f"lv{lvl} {names[which]}", t)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    np.savez_compressed(args.out, **out)


if __name__ == "__main__":
    main()
