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


def vote_histogram(decoded_local: torch.Tensor, n_cat: int, world: int, dst: int = 0):
    """Per-voxel category counts over the whole ensemble: local one-hot sum, then one reduce."""
    hist = torch.zeros((n_cat,) + tuple(decoded_local.shape[1:]), dtype=torch.int32, device=decoded_local.device)
    hist.scatter_add_(0, decoded_local.long().clamp(0, n_cat - 1),
                      torch.ones_like(decoded_local, dtype=torch.int32))
    if world > 1:
        dist.reduce(hist, dst=dst, op=dist.ReduceOp.SUM)
    return hist
