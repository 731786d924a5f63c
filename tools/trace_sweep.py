"""Developer tool (GPU box): per-block timeline of the triangular sweeps.

    python tools/trace_sweep.py --size 128 --out gpurun_out/trace128.npz

Runs one traced apply per (level, sweep) and stores [nblocks, 8] arrays
(start ns, slab-loaded ns, last-row-done ns, SM id, polls of last row, its thread, poller passes)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
    import torch

    import hifir_b200 as hb
    from bench import cached_levels
    from hifir_b200 import build, problems as P

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="poisson")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--out", default="gpurun_out/trace.npz")
    ap.add_argument("--no-trace", action="store_true")
    args = ap.parse_args()
    build.build()
    A, levels = cached_levels(args.workload, args.size)
    G = hb.GpuHif(levels)
    b = torch.from_numpy(P.seeded_rhs(A[0], 0)).cuda()
    x = torch.empty_like(b)
    for _ in range(3):
        G.solve_dev(b.data_ptr(), x.data_ptr())
    G.synchronize()
    G.set_stream(torch.cuda.current_stream().cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        G.solve_dev(b.data_ptr(), x.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    print(f"ROW_THREADS={os.environ.get('HIFIR_B200_ROW_THREADS', 'default')}: apply {e0.elapsed_time(e1) / 20:.3f} ms")
    prof = {}
    for _ in range(3):
        for name, ms in G.profile_solve_dev(b.data_ptr(), x.data_ptr()):
            prof[name] = ms
    print("  " + " ".join(f"{k}={v * 1e3:.0f}us" for k, v in prof.items()))
    if args.no_trace:
        return
    out = {}
    names = ("downL", "downU", "upL", "upU")
    for lvl in range(len(levels)):
        if not levels[lvl]["m"]:
            continue
        for which in range(4):
            t = G.trace_sweep(b.data_ptr(), x.data_ptr(), lvl, which)
            out[f"lv{lvl}_{names[which]}"] = t
            t0 = t[:, 0].min()
            dur = (t[:, 2].max() - t0) / 1e3
            life = (t[:, 2] - t[:, 0]) / 1e3
            load = (t[:, 1].astype(np.int64) - t[:, 0].astype(np.int64)) / 1e3
            halo_wait = (t[:, 7].astype(np.int64) - t[:, 1].astype(np.int64)).clip(min=0) / 1e3
            tail = (t[:, 2].astype(np.int64) - np.maximum(t[:, 7], t[:, 1]).astype(np.int64)) / 1e3
            print(f"   halo wait after load mean {halo_wait.mean():.1f} us; in-block tail after halo mean {tail.mean():.1f} "
                  f"max {tail.max():.1f} us")
            print(f"lv{lvl} {names[which]}: blocks {len(t)} sweep {dur:.1f} us; block life mean {life.mean():.1f} "
                  f"max {life.max():.1f} us; load mean {load.mean():.2f} us; poller passes mean {t[:, 6].mean():.0f} "
                  f"max {t[:, 6].max()}; last-row polls mean {t[:, 4].mean():.0f} max {t[:, 4].max()}; "
                  f"SMs used {len(np.unique(t[:, 3]))}")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    np.savez_compressed(args.out, **out)


if __name__ == "__main__":
    main()
