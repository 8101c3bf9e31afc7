"""Task-level pieces of ``Geo3DStochInterp`` that sit on the hot path
(project/geodata-3d-unconditional/model_train_inference.py:264-484): the simplex category
embedding, ``embed`` / ``decode``, the flow-matching loss and the EMA shadow update — each as a
single CUDA kernel behind the C ABI.  Lightning plumbing, logging, checkpoints and data loading
stay with the reference (out of scope).
"""
from __future__ import annotations

from typing import Any, List, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .interpolation import LinearInterpolant, StochasticInterpolator
from .unet3d import Unet3D, Unet3DCond


def simplex_embedding(n_cats: int, n_dims: int) -> torch.Tensor:
    """Centred-simplex embedding with unit rows (_initialize_embedding, :330-356); 270 numbers,
    built once on the host."""
    eye = torch.eye(n_cats)
    m = torch.zeros(n_cats, n_dims)
    m[:, :n_cats] = eye - 1.0 / n_cats
    return m / m.norm(dim=1, keepdim=True)


def embed(weight: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """[B,1,X,Y,Z] integer categories (-1 .. n_cat-2) -> [B,E,X,Y,Z] fp32 (embed, :361-370)."""
    if not x.is_cuda:
        raise RuntimeError("embed runs on CUDA only (no CPU fallback)")
    cats = x.squeeze(1).long().contiguous()
    B = cats.shape[0]
    n = cats[0].numel()
    w = weight.detach().to(device=x.device, dtype=torch.float32).contiguous()
    ncat, E = w.shape
    out = torch.empty((B, E) + tuple(cats.shape[1:]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.ftb_embed(_lib.ptr(cats), _lib.ptr(w), _lib.ptr(out), B, E, ncat, n, 1,
                                      _lib.stream_ptr()))
    return out


def decode(weight: torch.Tensor, x: torch.Tensor, return_logits: bool = False) -> torch.Tensor:
    """[B,E,X,Y,Z] -> [B,X,Y,Z] int64 nearest category by cosine similarity (decode, :373-404).
    One kernel, no [B,ncat,E,X,Y,Z] temporary; integer output bit-exact w.r.t. the reference's
    CPU op order.  The 15x18 embedding matrix is row-normalised on the host (F.normalize, :384)."""
    if not x.is_cuda:
        raise RuntimeError("decode runs on CUDA only (no CPU fallback)")
    en = F.normalize(weight.detach().float().cpu(), dim=1).to(x.device).contiguous()
    ncat, E = en.shape
    if x.shape[1] != E:
        raise ValueError(f"expected {E} embedding channels, got {x.shape[1]}")
    xin = x.detach().float().contiguous()
    B = xin.shape[0]
    n = xin[0, 0].numel()
    if return_logits:   # :398-399: the [B, ncat, X, Y, Z] cosine logits instead of their argmax
        logits = torch.empty((B, ncat) + tuple(xin.shape[2:]), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib.ftb_decode_logits(_lib.ptr(xin), _lib.ptr(en), _lib.ptr(logits), B, E, ncat, n,
                                                  _lib.stream_ptr()))
        return logits
    out = torch.empty((B,) + tuple(xin.shape[2:]), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.ftb_decode(_lib.ptr(xin), _lib.ptr(en), _lib.ptr(out), B, E, ncat, n,
                                       _lib.stream_ptr()))
    return out


def flow_loss(VT: torch.Tensor, VT_hat: torch.Tensor) -> torch.Tensor:
    """mse(VT, VT_hat) / mse(VT, 0) (:443) as one reduction kernel; returns a 0-d fp32 tensor."""
    if not VT.is_cuda:
        raise RuntimeError("flow_loss runs on CUDA only (no CPU fallback)")
    v = VT.detach().float().contiguous()
    vh = VT_hat.detach().float().contiguous()
    acc = torch.zeros(2, dtype=torch.float64, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib.ftb_mse_ratio_accumulate(_lib.ptr(v), _lib.ptr(vh), v.numel(), _lib.ptr(acc),
                                                     _lib.stream_ptr()))
    return (acc[0] / acc[1]).float()


def ema_update_(shadow: torch.Tensor, param: torch.Tensor, decay: float) -> torch.Tensor:
    """shadow <- decay*shadow + (1-decay)*param in place (EMACallback.on_train_batch_end,
    project/geodata-3d-conditional/callbacks.py:263-266)."""
    if not (shadow.is_cuda and param.is_cuda):
        raise RuntimeError("ema_update_ runs on CUDA only (no CPU fallback)")
    assert shadow.is_contiguous() and shadow.dtype == torch.float32 and shadow.shape == param.shape
    p = param.detach().float().contiguous()
    with torch.cuda.device(shadow.device):
        _lib.check(_lib.lib.ftb_ema_update(_lib.ptr(shadow), _lib.ptr(p), shadow.numel(), float(decay),
                                           _lib.stream_ptr()))
    return shadow


class EMAShadow:
    """Per-batch EMA of the trainable parameters (callbacks.py:225-268 semantics: nothing before
    ``start_step``; the first eligible update clones the parameter, later ones blend)."""

    def __init__(self, decay=0.9999, start_step=15000, update_every=1):
        self.decay, self.start_step, self.update_every = decay, start_step, update_every
        self.shadow = {}
        self.step = 0

    @torch.no_grad()
    def update(self, module: nn.Module):
        self.step += 1
        if self.step < self.start_step or self.step % self.update_every != 0:
            return
        for name, p in module.named_parameters():
            if not p.requires_grad:
                continue
            if name not in self.shadow:
                self.shadow[name] = p.detach().clone().float().contiguous()
            else:
                ema_update_(self.shadow[name], p, self.decay)

    @torch.no_grad()
    def apply_to(self, module: nn.Module):
        for name, p in module.named_parameters():
            if name in self.shadow:
                p.copy_(self.shadow[name])


class Geo3DStochInterp(nn.Module):
    """The hot-path surface of the reference LightningModule (:264-484) as a plain nn.Module:
    ``net`` (B200 Unet3D), frozen simplex ``embedding``, ``interpolator``, ``embed``, ``decode``,
    ``forward``.  Attribute names match so ``net.*`` / ``embedding.weight`` checkpoint keys load."""

    def __init__(self, data_shape: Tuple[int, int, int] = (32, 32, 32),
                 time_range: List[float] = [0.0005, 0.9995], num_categories: int = 15,
                 embedding_dim: int = 20, lambda_angle: float = 0.1, learning_rate=None, lr_decay=None,
                 **model_params: Any):
        super().__init__()
        self.data_shape = data_shape
        self.time_range = time_range
        self.num_categories = num_categories
        self.embedding_dim = embedding_dim
        self.lambda_angle = lambda_angle
        self.embedding = nn.Embedding(num_categories, embedding_dim)
        with torch.no_grad():
            self.embedding.weight.copy_(simplex_embedding(num_categories, embedding_dim))
        self.embedding.weight.requires_grad = False
        model_params["data_channels"] = embedding_dim
        self.net = Unet3D(**model_params)
        self.ema_shadow = {}
        self.interpolant = LinearInterpolant(one_sided=True)
        self.interpolator = StochasticInterpolator(self.interpolant)

    def forward(self, x, t):
        return self.net(x, t)

    def embed(self, x):
        return embed(self.embedding.weight, x)

    def decode(self, x, return_logits=False):
        return decode(self.embedding.weight, x, return_logits)

    @torch.no_grad()
    def flow_matching_loss(self, batch, noise1=None, X0=None, T=None):
        """Forward half of training_step (:417-457): the loss value for given (or drawn) noise.
        The backward pass through the B200 network is not implemented yet."""
        X1 = self.embed(batch)
        X1 = X1 + 1e-3 * (torch.randn_like(X1) if noise1 is None else noise1)
        X0 = torch.randn_like(X1) if X0 is None else X0
        if T is None:
            T = torch.empty(X1.size(0), device=X1.device).uniform_(self.time_range[0], self.time_range[1])
        XT, VT = self.interpolator.flow_objective(T, X0, X1)
        return flow_loss(VT, self.net(XT, T))


class Geo3DStochInterpCond(nn.Module):
    """Hot-path surface of the conditional project's LightningModule
    (project/geodata-3d-conditional/model_train_sh_inference_cond.py:247-495): ``net`` is the B200 ``Unet3DCond``,
    the simplex ``embedding`` is frozen (:302), ``forward(x, ATb, t)``; ``conditioning(batch)`` is the fused
    embed + combined mask + ``ATb = X1 * mask`` front-end of training_step (:413-420)."""

    def __init__(self, data_shape: Tuple[int, int, int] = (32, 32, 32),
                 time_range: List[float] = [0.0001, 0.9999], num_categories: int = 15, embedding_dim: int = 20,
                 lambda_reconstruct: float = 1.0, learning_rate: float = 2e-3, lr_decay: float = 0.999,
                 **model_params: Any):
        super().__init__()
        self.data_shape = data_shape
        self.time_range = time_range
        self.num_categories = num_categories
        self.embedding_dim = embedding_dim
        self.lambda_reconstruct = lambda_reconstruct
        self.learning_rate, self.lr_decay = learning_rate, lr_decay
        self.embedding = nn.Embedding(num_categories, embedding_dim)
        with torch.no_grad():
            self.embedding.weight.copy_(simplex_embedding(num_categories, embedding_dim))
        self.embedding.weight.requires_grad = False
        model_params["data_channels"] = embedding_dim
        self.net = Unet3DCond(**model_params)
        self.ema_shadow = {}
        self.interpolant = LinearInterpolant(one_sided=True)
        self.interpolator = StochasticInterpolator(self.interpolant)

    def forward(self, x, ATb, t):
        return self.net(x, ATb, t)

    def embed(self, x):
        return embed(self.embedding.weight, x)

    def decode(self, x, return_logits=False):
        return decode(self.embedding.weight, x, return_logits)

    def conditioning(self, batch, bores=None, n_bores=None, generator=None):
        """(X1, ATb, mask) of a category batch [B,1,X,Y,Z] in one kernel (see boreholes.conditioning_frontend)."""
        from .boreholes import conditioning_frontend
        return conditioning_frontend(batch, self.embedding.weight, bores, n_bores, generator)


# --------------------------------------------------------------------------------------- checkpoint interop
def load_model_with_ema_option(module: nn.Module, ckpt, map_location="cpu", use_ema: bool = False, strict: bool = True):
    """Counterpart of ``load_model_with_ema_option`` (project/geodata-3d-conditional/model_inference_experiments.py:
    387-403) for the B200 modules: ``ckpt`` is a Lightning ``.ckpt`` path or an already loaded dict with
    ``"state_dict"`` (keys ``net.*``, ``embedding.weight``: the reference LightningModule's own names, which
    ``Geo3DStochInterp`` / ``Geo3DStochInterpCond`` share) and optionally ``"ema_shadow"`` (EMACallback.on_save_checkpoint,
    callbacks.py:295-303: name -> tensor over ``pl_module.named_parameters()``).  With ``use_ema`` the shadow replaces the
    parameters it covers, as the reference does.  Pure host-side plumbing (no math): returns ``module``."""
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        ckpt = torch.load(ckpt, map_location=map_location, weights_only=False)
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    module.load_state_dict(sd, strict=strict)
    if use_ema and "ema_shadow" in ckpt:
        shadow = ckpt["ema_shadow"]
        with torch.no_grad():
            for name, param in module.named_parameters():
                if name in shadow:
                    param.data.copy_(shadow[name].to(param.device))
    elif use_ema:
        print("WARNING: 'ema_shadow' not found in checkpoint. Using regular weights.")
    return module


def lightning_checkpoint(module: nn.Module, trainer=None, update_on_cpu: bool = True) -> dict:
    """The two entries of a Lightning ``.ckpt`` the reference's loaders read (``state_dict``, ``ema_shadow`` keyed like
    ``pl_module.named_parameters()``, ``ema_update_on_cpu``; callbacks.py:295-303), from a B200 module and, when given,
    the EMA buffer of a ``FlowTrainer`` / ``CondFlowTrainer`` — so weights trained here load into the reference."""
    ck = {"state_dict": {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}}
    if trainer is not None:
        ck["ema_shadow"] = {f"net.{k}": (v.cpu() if update_on_cpu else v) for k, v in trainer.ema_state_dict().items()}
        ck["ema_update_on_cpu"] = update_on_cpu
    return ck
