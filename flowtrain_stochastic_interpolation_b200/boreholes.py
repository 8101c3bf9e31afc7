"""Conditioning front-end of the conditional project on the GPU: drop-ins for
``project/geodata-3d-conditional/boreholes.py`` (``make_boreholes_mask`` :45-75, ``make_surface_mask`` :77-111,
``make_combined_mask`` :114-129) plus the fused ``embed -> mask -> ATb = X1 * mask`` step of the training / inference
scripts (model_train_sh_inference_cond.py:413-420, model_inference_experiments.py:228-232).

The reference fills the masks with Python loops and one ``.item()`` device sync per borehole; here the random borehole
columns are drawn on the host from a CPU generator (no device sync) and ONE kernel (``ftb_cond_frontend``) applies the
mask rule, the embedding lookup and the product.  The mask rule is bit-exact given the same borehole coordinates; the
random stream itself is this build's own (the reference draws from the device generator, one call per coordinate).
There is no CPU path: every function raises without the CUDA library / a GPU.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib

MAX_BORES = 64


def jittered_grid_points(X: int, Y: int, n_bores: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """``_jittered_grid_points`` (boreholes.py:9-42): one point per cell of an n_x x n_y grid, jittered uniformly
    inside the cell, clamped to the volume, truncated to integers; the first ``n_bores`` in (i, j) order.
    Returns a CPU int64 tensor [n, 2]."""
    n_x = int(math.floor(math.sqrt(n_bores)))
    n_y = int(math.ceil(n_bores / n_x))
    cw_x, cw_y = X / n_x, Y / n_y
    r = torch.rand(n_x * n_y, 2, generator=generator)        # (rand_x, rand_y) per cell, cells in (i, j) order
    i = torch.arange(n_x, dtype=torch.float32).repeat_interleave(n_y)
    j = torch.arange(n_y, dtype=torch.float32).repeat(n_x)
    px = ((i + 0.5) * cw_x + (r[:, 0] * cw_x - cw_x / 2)).clamp(min=0, max=X - 1)
    py = ((j + 0.5) * cw_y + (r[:, 1] * cw_y - cw_y / 2)).clamp(min=0, max=Y - 1)
    return torch.stack((px, py), dim=1)[:n_bores].to(torch.long)


def draw_boreholes(B: int, X: int, Y: int, generator: Optional[torch.Generator] = None,
                   lo: int = 8, hi: int = 32) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per sample: n_bores ~ randint(lo, hi) (boreholes.py:67) jittered-grid columns.  Returns pinned CPU tensors
    ``(bores int32 [B, MAX_BORES, 2], n_bores int32 [B])`` ready for an asynchronous upload."""
    bores = torch.zeros(B, MAX_BORES, 2, dtype=torch.int32)
    nb = torch.zeros(B, dtype=torch.int32)
    for b in range(B):
        n = int(torch.randint(lo, hi, (1,), generator=generator))
        pts = jittered_grid_points(X, Y, n, generator)
        bores[b, : pts.shape[0]] = pts.to(torch.int32)
        nb[b] = pts.shape[0]
    if torch.cuda.is_available():
        bores, nb = bores.pin_memory(), nb.pin_memory()
    return bores, nb


def _cats(X: torch.Tensor) -> torch.Tensor:
    if not X.is_cuda:
        raise RuntimeError("the conditioning front-end runs on CUDA only (no CPU fallback)")
    if X.dim() != 5:
        raise ValueError(f"expected [B, C, X, Y, Z], got {tuple(X.shape)}")
    return X[:, 0].long().contiguous()


def _frontend(X, bores, n_bores, surface, weight=None, want_x1=False, want_atb=False):
    cats = _cats(X)
    B, sx, sy, sz = cats.shape
    dev = X.device
    if bores is not None:
        bores = bores.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
        n_bores = n_bores.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
        if bores.dim() != 3 or bores.shape[0] != B or bores.shape[2] != 2 or n_bores.shape[0] != B:
            raise ValueError("bores must be [B, max_bores, 2] and n_bores [B]")
        max_b = bores.shape[1]
    else:
        max_b = 0
    mask = torch.empty((B, 1, sx, sy, sz), dtype=torch.uint8, device=dev)
    w = x1 = atb = None
    E = ncat = 1
    if weight is not None:
        w = weight.detach().to(device=dev, dtype=torch.float32).contiguous()
        ncat, E = w.shape
        if want_x1:
            x1 = torch.empty((B, E, sx, sy, sz), dtype=torch.float32, device=dev)
        if want_atb:
            atb = torch.empty((B, E, sx, sy, sz), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.ftb_cond_frontend(_lib.ptr(cats), _lib.ptr(bores), _lib.ptr(n_bores), max_b, _lib.ptr(w),
                                              B, E, ncat, 1, sx, sy, sz, 1 if surface else 0, _lib.ptr(mask),
                                              _lib.ptr(x1), _lib.ptr(atb), _lib.stream_ptr()))
    return mask.view(torch.bool), x1, atb


def make_boreholes_mask(X: torch.Tensor, bores=None, n_bores=None, generator=None) -> torch.Tensor:
    """bool [B,1,X,Y,Z]: full-depth columns at the borehole (x, y) positions (boreholes.py:45-75).  ``bores`` /
    ``n_bores`` as returned by ``draw_boreholes``; drawn here when omitted."""
    if bores is None:
        bores, n_bores = draw_boreholes(X.shape[0], X.shape[2], X.shape[3], generator)
    return _frontend(X, bores, n_bores, surface=False)[0]


def make_surface_mask(X: torch.Tensor) -> torch.Tensor:
    """bool [B,1,X,Y,Z]: top z-slice, every air voxel (category -1) and the voxel below it (boreholes.py:77-111)."""
    return _frontend(X, None, None, surface=True)[0]


def make_combined_mask(X: torch.Tensor, bores=None, n_bores=None, generator=None) -> torch.Tensor:
    """``make_boreholes_mask | make_surface_mask`` (boreholes.py:114-129) in one kernel."""
    if bores is None:
        bores, n_bores = draw_boreholes(X.shape[0], X.shape[2], X.shape[3], generator)
    return _frontend(X, bores, n_bores, surface=True)[0]


def conditioning_frontend(batch: torch.Tensor, weight: torch.Tensor, bores=None, n_bores=None, generator=None):
    """``X1 = embed(batch); mask = make_combined_mask(batch); ATb = X1 * mask``
    (model_train_sh_inference_cond.py:413-420) in ONE kernel.  Returns ``(X1, ATb, mask)`` with mask bool [B,1,X,Y,Z]."""
    if bores is None:
        bores, n_bores = draw_boreholes(batch.shape[0], batch.shape[2], batch.shape[3], generator)
    mask, x1, atb = _frontend(batch, bores, n_bores, surface=True, weight=weight, want_x1=True, want_atb=True)
    return x1, atb, mask
