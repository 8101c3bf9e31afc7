"""world_size-2 gloo worker: the bucketed gradient all-reduce that FlowTrainer drives from inside the backward
(BucketAllReduce), on CPU tensors.  The 'engine' here is a stand-in that reports parameter ranges back to front
exactly like ftb_unet3d_backward's bucket callback."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowtrain_stochastic_interpolation_b200 import sharding  # noqa: E402
from flowtrain_stochastic_interpolation_b200.training import BucketAllReduce  # noqa: E402


def main():
    rank, _, world = sharding.init_distributed("gloo")
    n = 10_000
    g = torch.Generator().manual_seed(7 + rank)
    flat = torch.randn(n, generator=g)
    mine = flat.clone()
    # ranges as the engine reports them: last layers first, with one out-of-order neighbour (the mid block
    # sits between the ups and the final block in state_dict order)
    ranges = [(9000, 1000), (6000, 2000), (5000, 1000), (4000, 1000), (8000, 1000), (100, 3900), (0, 100)]
    red = BucketAllReduce(flat, min_elems=1500)
    for off, cnt in ranges:
        red(None, off, cnt)
    red.finish()
    covered = torch.zeros(n, dtype=torch.int32)
    for lo, hi in red.launched:
        covered[lo:hi] += 1
    assert torch.all(covered == 1), "every element must be reduced exactly once"
    assert len(red.launched) < len(ranges), "adjacent ranges must be coalesced"
    others = [torch.zeros(n) for _ in range(world)]
    dist.all_gather(others, mine)
    want = sum(others)
    assert torch.allclose(flat, want, atol=1e-6)
    if rank == 0:
        print("BUCKET_OK", red.launched)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
