"""torchrun worker (one rank per GPU, NCCL): sharded ensemble votes.  Every rank decodes its shard of the samples into
a vote histogram (ftb_decode_vote), one all-reduce(sum) combines them; the result must equal the single-process
histogram over all samples, and the statistics must be identical on every rank."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowtrain_stochastic_interpolation_b200 as ftb  # noqa: E402
from flowtrain_stochastic_interpolation_b200 import sharding  # noqa: E402
from oracle import synth  # noqa: E402  (synthetic inputs only)


def main():
    rank, local, world = sharding.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    W = ftb.simplex_embedding(15, 15)
    n_samples, shape = 11, (8, 16, 16)          # 11 samples over 2 ranks: uneven shards
    x = synth.synth_input((n_samples, 15) + shape, 77, "ens")
    mine = list(sharding.shard_indices(n_samples, rank, world))
    v = ftb.EnsembleVotes(W, shape, dev)
    v.add(x[mine].to(dev))
    v.all_reduce()
    ref = ftb.EnsembleVotes(W, shape, dev)
    ref.add(x.to(dev))
    assert v.samples == n_samples and torch.equal(v.counts, ref.counts), f"rank {rank}: histogram mismatch"
    a, b = v.finalize(), ref.finalize()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    dist.barrier()
    if rank == 0:
        print("ENSEMBLE_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
