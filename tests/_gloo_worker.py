"""world_size-2 gloo worker for tests/test_host_cpu.py::test_sharded_ensemble_gloo_world2."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowtrain_stochastic_interpolation_b200 import sharding  # noqa: E402


def fake_sample(i):  # CPU stand-in for solve+decode of sample i: a deterministic volume
    g = torch.Generator().manual_seed(42 + i)
    return torch.randint(0, 15, (4, 4, 4), generator=g)


def main():
    rank, _, world = sharding.init_distributed("gloo")
    n = 5  # not divisible by world: ragged shards
    mine = list(sharding.shard_indices(n, rank, world))
    local = torch.stack([fake_sample(i) for i in mine])
    full = sharding.gather_samples(local, n, rank, world)
    # stand-in for the per-rank histogram the decode -> vote kernel accumulates (EnsembleVotes.counts)
    hist = torch.zeros(15, 4, 4, 4, dtype=torch.int32)
    hist.scatter_add_(0, local, torch.ones_like(local, dtype=torch.int32))
    hist = sharding.reduce_votes(hist, world, dst=0)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    if rank == 0:
        want = torch.stack([fake_sample(i) for i in range(n)])
        assert torch.equal(full, want)
        want_hist = torch.zeros(15, 4, 4, 4, dtype=torch.int32)
        want_hist.scatter_add_(0, want, torch.ones_like(want, dtype=torch.int32))
        assert torch.equal(hist, want_hist)
        assert int(hist.sum()) == n * 64
        print("GLOO_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
