"""Ensemble statistics of the conditional project on the GPU: per-voxel vote probabilities, entropy
and most probable category of an ensemble of generated volumes
(project/geodata-3d-conditional/model_inference_experiments.py:442-459; inference_demo.ipynb cell 21).

The reference decodes every sample to an int64 volume, one-hot encodes them ([S,15,64^3] floats) and averages.
Here ``EnsembleVotes.add`` decodes a batch of samples and adds them to an int32 vote histogram in ONE kernel
(``ftb_decode_vote``: same bit-exact decode order as ``decode``), ranks combine their histograms with one
``all_reduce(sum)`` (NCCL), and ``finalize`` turns counts into the statistics.  No CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib


class EnsembleVotes:
    def __init__(self, embedding_weight: torch.Tensor, spatial, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("EnsembleVotes runs on CUDA only (no CPU fallback)")
        self.en = F.normalize(embedding_weight.detach().float().cpu(), dim=1).to(self.device).contiguous()  # decode :384
        self.ncat, self.E = self.en.shape
        self.spatial = tuple(int(s) for s in spatial)
        self.n = 1
        for s in self.spatial:
            self.n *= s
        self.counts = torch.zeros((self.ncat,) + self.spatial, dtype=torch.int32, device=self.device)
        self.samples = 0

    def add(self, x: torch.Tensor, return_decoded: bool = False) -> Optional[torch.Tensor]:
        """x: [S, E, X, Y, Z] generated samples (embedding space).  Adds their votes; optionally returns the decoded
        int64 volumes [S, X, Y, Z] (written by the same kernel)."""
        if x.device != self.device or x.dim() != 5 or x.shape[1] != self.E or tuple(x.shape[2:]) != self.spatial:
            raise ValueError(f"expected [S, {self.E}, {self.spatial}] on {self.device}, got {tuple(x.shape)} on {x.device}")
        xin = x.detach().float().contiguous()
        S = xin.shape[0]
        dec = torch.empty((S,) + self.spatial, dtype=torch.int64, device=self.device) if return_decoded else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.ftb_decode_vote(_lib.ptr(xin), _lib.ptr(self.en), S, self.E, self.ncat, self.n,
                                                _lib.ptr(dec), _lib.ptr(self.counts), _lib.stream_ptr()))
        self.samples += S
        return dec

    def all_reduce(self, group=None):
        """Sum the histograms (and sample counts) of all ranks: the only collective of ensemble sampling."""
        import torch.distributed as dist
        from .sharding import reduce_votes
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            reduce_votes(self.counts, dist.get_world_size(group), group=group)
            tot = torch.tensor([self.samples], dtype=torch.int64, device=self.device)
            dist.all_reduce(tot, group=group)
            self.samples = int(tot.item())
        return self

    def finalize(self, probabilities: bool = True):
        """Returns a dict: ``probability_vector`` [ncat,X,Y,Z] (one-hot mean, :442-447), ``entropy`` [X,Y,Z]
        (-sum p log(p + 1e-8), :449-452), ``most_probable`` int64 [X,Y,Z] (first argmax - 1, :454-455) and
        ``entropy_masked`` (entropy with -1 where the most probable category is air, :458-459)."""
        if self.samples < 1:
            raise RuntimeError("no samples were added")
        dev = self.device
        probs = torch.empty((self.ncat,) + self.spatial, dtype=torch.float32, device=dev) if probabilities else None
        ent = torch.empty(self.spatial, dtype=torch.float32, device=dev)
        entm = torch.empty(self.spatial, dtype=torch.float32, device=dev)
        most = torch.empty(self.spatial, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.ftb_vote_finalize(_lib.ptr(self.counts), self.samples, self.ncat, self.n, -1,
                                                  _lib.ptr(probs), _lib.ptr(ent), _lib.ptr(most), _lib.ptr(entm),
                                                  _lib.stream_ptr()))
        return {"probability_vector": probs, "entropy": ent, "most_probable": most, "entropy_masked": entm}
