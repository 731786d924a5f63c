"""Developer tool (GPU box): multi-rhs apply throughput (config 4: conv-diff 96^3, nrhs 16/64)."""
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
    build.build()
    workload, size = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("convdiff", 96)
    A, levels = cached_levels(workload, size)
    n = A[0]
    G = hb.GpuHif(levels)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    G.set_stream(side.cuda_stream)
    st = G.stats()
    print("HIFIR_B200_MRHS_WIDE", os.environ.get("HIFIR_B200_MRHS_WIDE"), "HIFIR_B200_MRHS_BATCH", os.environ.get("HIFIR_B200_MRHS_BATCH"))
    for nrhs in (1, 8, 16, 32, 64):
        B = torch.from_numpy(P.seeded_rhs(n, 0, nrhs=nrhs)).cuda() if nrhs > 1 else torch.from_numpy(P.seeded_rhs(n, 0)).cuda()
        X = torch.empty_like(B)
        for _ in range(2):
            G.solve_mrhs_dev(nrhs, B.data_ptr(), X.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            G.solve_mrhs_dev(nrhs, B.data_ptr(), X.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        G.synchronize()
        ms = e0.elapsed_time(e1) / reps
        bytes_apply = st["bytes_factors"] + st["bytes_dense"] + nrhs * st["bytes_vec_per_rhs"]
        print(f"{workload} {size} nrhs={nrhs}: {ms:.3f} ms per block apply = {nrhs / ms * 1e3:.0f} column-applies/s; "
              f"algorithmic {bytes_apply / 1e9:.2f} GB -> {bytes_apply / ms / 1e6:.0f} GB/s")
    # parity of some columns against single-rhs applies
    for nrhs in (16, 64, 70):
        B = torch.from_numpy(P.seeded_rhs(n, 0, nrhs=nrhs)).cuda()
        X = torch.empty_like(B)
        G.solve_mrhs_dev(nrhs, B.data_ptr(), X.data_ptr())
        errs = []
        for c in (0, 5, nrhs // 2 + 1, nrhs - 1):
            bc = B[:, c].contiguous()
            xc = torch.empty_like(bc)
            G.solve_dev(bc.data_ptr(), xc.data_ptr())
            G.synchronize()
            errs.append(float((X[:, c] - xc).norm() / xc.norm()))
        print(f"nrhs={nrhs}: columns vs single-rhs applies:", errs)


if __name__ == "__main__":
    main()
