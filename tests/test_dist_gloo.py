"""N > 1 path on CPU: world_size-2 gloo process group, contiguous sharding + final all_gather (smrt_b200/dist.py).
The per-shard solve is the SIMT-emulated device code (test infrastructure); on GPUs it is the CUDA library."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from emu_util import load_golden


def _worker(rank, world, port, ret):
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emu_util import emu_solve
    from smrt_b200 import dist as sdist

    d, batch, opts = load_golden("iba_multiangle_passive")  # 6 problems, 16 streams
    batch = batch.subset(slice(0, 5))  # uneven split: 3 + 2
    values, status = sdist.solve_sharded(batch, opts, solve_fn=lambda shard, o: emu_solve(shard, o, threads=64))
    ret[rank] = (values, status)
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from smrt_b200.dist import shard_bounds

    for B in (0, 1, 5, 60000):
        for w in (1, 2, 3, 8):
            cuts = [shard_bounds(B, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1


def test_two_rank_gloo_gather_matches_reference_fixture():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29517 + os.getpid() % 1000, ret), nprocs=world, join=True)
    d, _, _ = load_golden("iba_multiangle_passive")
    for rank in range(world):
        values, status = ret[rank]
        assert values.shape == (5, 2, 7) and np.all(status == 0)
        np.testing.assert_allclose(values, d["ref_values"][:5], rtol=1e-10)
