"""Multi-GPU plumbing for the parts of the path that shard naturally.

Sampling and ensembles are independent per sample: sample ``i`` goes to rank ``i % world`` and no
collective touches the data path.  The only exchanges are optional, after the solve: gathering
the decoded int64 volumes and summing the per-voxel category vote histogram
(project/geodata-3d-conditional/inference_demo.ipynb cell 21) onto rank 0.  One process per GPU,
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_indices(n_samples: int, rank: int, world: int) -> range:
    """Sample indices owned by ``rank``: i with i % world == rank."""
    return range(rank, n_samples, world)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_distributed(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for world 1)."""
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def gather_samples(local: torch.Tensor, n_samples: int, rank: int, world: int, dst: int = 0):
    """Gather per-rank stacks of per-sample tensors (rank r holds samples r, r+world, ...) into the
    original sample order on ``dst``.  Ranks may hold different counts (n_samples % world != 0)."""
    if world == 1:
        return local
    per = (n_samples + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n_samples,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = list(shard_indices(n_samples, r, world))
        out[idx] = bufs[r][: len(idx)]
    return out


def reduce_votes(counts: torch.Tensor, world: int, dst: int = None, group=None):
    """The one collective of ensemble sampling: sum the per-rank int32 vote histograms [n_cat, X, Y, Z] that
    ``EnsembleVotes.add`` (decode -> histogram kernel) accumulated.  ``dst=None``: all-reduce (every rank gets the
    ensemble statistics); otherwise a reduce onto ``dst``.  In place; no-op for world 1."""
    if world > 1:
        if dst is None:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(counts, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return counts
