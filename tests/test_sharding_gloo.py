"""N>1 host logic on CPU: world_size-2 gloo run of the column sharding + result gather used by the
multi-GPU benchmark (each rank holds a factor replica and applies to its own columns; here the
apply is the C oracle, the GPU apply has its own parity tests)."""
import os
import socket

import numpy as np
import pytest

from conftest import load_golden, relerr
from hifir_b200.sharding import shard_range


def test_shard_range_is_a_partition():
    for total in (0, 1, 7, 16, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, nrhs, out_path):
    import torch
    import torch.distributed as dist

    from hifir_b200 import problems as P
    from hifir_b200.sharding import gather_columns, local_columns
    from oracle import port as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden("convdiff14_ml")
    B = P.seeded_rhs(g.n, 0, nrhs=nrhs)
    Xl = O.OracleHif(g.levels).solve_mrhs(local_columns(B, world, rank))
    X = gather_columns(Xl, nrhs, world, rank, dist=dist)
    if rank == 0:
        np.save(out_path, X.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nrhs", [4, 5])
def test_two_rank_column_sharding(tmp_path, nrhs):
    import torch.multiprocessing as mp

    from hifir_b200 import problems as P
    from oracle import port as O
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "X.npy")
    mp.spawn(_worker, args=(2, port, nrhs, out), nprocs=2, join=True)
    g = load_golden("convdiff14_ml")
    B = P.seeded_rhs(g.n, 0, nrhs=nrhs)
    ref = O.OracleHif(g.levels).solve_mrhs(B)
    X = np.load(out)
    assert X.shape == ref.shape and relerr(X, ref) <= 1e-15
