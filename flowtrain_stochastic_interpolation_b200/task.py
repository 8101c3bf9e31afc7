"""Task-level pieces of ``Geo3DStochInterp`` that sit on the hot path
(project/geodata-3d-unconditional/model_train_inference.py:264-484): the simplex category
embedding, ``embed`` / ``decode``, the flow-matching loss and the EMA shadow update — each as a
single CUDA kernel behind the C ABI — and the LightningModule-shaped surface of the two task
modules (``training_step``, ``configure_optimizers``, ``on_save_checkpoint`` / ``on_load_checkpoint``,
``on_train_epoch_end``; :417-484 and model_train_sh_inference_cond.py:401-495).  When ``lightning`` /
``pytorch_lightning`` is importable the modules ARE LightningModules (so ``Trainer.fit`` drives them
as it drives the reference); otherwise they are plain ``nn.Module``s with the same methods (``log`` /
``log_dict`` collect into ``self.logged``).  Loggers, data loading and plotting stay with the reference.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, List, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .interpolation import LinearInterpolant, StochasticInterpolator
from .unet3d import Unet3D, Unet3DCond

try:                                   # the reference's base class (pyproject.toml:16-17), when it is installed
    from lightning import LightningModule as _TaskBase
    HAVE_LIGHTNING = True
except ImportError:                    # pragma: no cover - depends on the environment
    try:
        from pytorch_lightning import LightningModule as _TaskBase
        HAVE_LIGHTNING = True
    except ImportError:
        _TaskBase = nn.Module
        HAVE_LIGHTNING = False


class _TaskHooks(_TaskBase):
    """What the two task modules share: hyper-parameter capture and the Lightning-named no-op plumbing that lets
    ``training_step`` run outside a Trainer."""

    def _capture_hparams(self, **hp):
        if HAVE_LIGHTNING:
            self.save_hyperparameters(hp)          # same ``self.hparams`` a Lightning ``.ckpt`` round-trips (:303)
        else:
            self.hparams = SimpleNamespace(**hp)
        self.logged: Dict[str, Any] = {}

    if not HAVE_LIGHTNING:
        def log_dict(self, d, **_kw):
            self.logged.update(d)

        def log(self, name, value, **_kw):
            self.logged[name] = value

    def _log_dict(self, d, **kw):
        """``self.log_dict`` (:446-455); outside a Trainer Lightning's own ``log_dict`` raises, so collect instead."""
        if HAVE_LIGHTNING and getattr(self, "_trainer", None) is None:
            self.logged.update(d)
        else:
            self.log_dict(d, **kw)

    def on_train_epoch_end(self, unused=None):     # :459-463
        lr = self.trainer.optimizers[0].param_groups[0]["lr"]
        self.log("lr", lr, on_epoch=True, logger=True)

    def on_save_checkpoint(self, checkpoint):      # :475-479
        checkpoint["ema_shadow"] = self.ema_shadow

    def on_load_checkpoint(self, checkpoint):      # :481-484
        self.ema_shadow = checkpoint["ema_shadow"]
        for m in self.modules():                   # checkpoint weights arrive through load_state_dict; be explicit
            if isinstance(m, Unet3D):
                m.mark_dirty()


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


class _FlowLossFn(torch.autograd.Function):
    """mse(VT, VT_hat) / mse(VT, 0) with its gradient w.r.t. VT_hat: ftb_mse_ratio_accumulate / ftb_mse_ratio_grad."""

    @staticmethod
    def forward(ctx, VT, VT_hat):
        v = VT.detach().float().contiguous()
        vh = VT_hat.detach().float().contiguous()
        acc = torch.zeros(2, dtype=torch.float64, device=v.device)
        with torch.cuda.device(v.device):
            _lib.check(_lib.lib.ftb_mse_ratio_accumulate(_lib.ptr(v), _lib.ptr(vh), v.numel(), _lib.ptr(acc),
                                                         _lib.stream_ptr()))
        ctx.save_for_backward(v, vh, acc)
        return (acc[0] / acc[1]).float()

    @staticmethod
    def backward(ctx, gout):
        v, vh, acc = ctx.saved_tensors
        dout = torch.empty_like(vh)
        with torch.cuda.device(v.device):
            _lib.check(_lib.lib.ftb_mse_ratio_grad(_lib.ptr(v), _lib.ptr(vh), v.numel(), _lib.ptr(acc), 1.0,
                                                   _lib.ptr(dout), _lib.stream_ptr()))
        return None, dout.mul_(gout.to(dout.dtype))   # gout stays on the device: no host sync


def flow_loss(VT: torch.Tensor, VT_hat: torch.Tensor) -> torch.Tensor:
    """mse(VT, VT_hat) / mse(VT, 0) (:443) as one reduction kernel; returns a 0-d fp32 tensor.  Differentiable
    w.r.t. ``VT_hat`` (one more kernel), so ``flow_loss(VT, net(XT, T)).backward()`` is the reference's line."""
    if not VT.is_cuda:
        raise RuntimeError("flow_loss runs on CUDA only (no CPU fallback)")
    return _FlowLossFn.apply(VT, VT_hat)


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


class Geo3DStochInterp(_TaskHooks):
    """The reference LightningModule (:264-484) on the B200 path: ``net`` (B200 Unet3D), frozen simplex
    ``embedding``, ``interpolator``, ``embed``, ``decode``, ``forward``, ``training_step``,
    ``configure_optimizers`` and the checkpoint hooks.  Attribute names match, so ``net.*`` /
    ``embedding.weight`` checkpoint keys load.  ``FlowTrainer`` (training.py) is the fused form of the same step."""

    def __init__(self, data_shape: Tuple[int, int, int] = (32, 32, 32),
                 time_range: List[float] = [0.0005, 0.9995], num_categories: int = 15,
                 embedding_dim: int = 20, lambda_angle: float = 0.1, learning_rate=None, lr_decay=None,
                 **model_params: Any):
        super().__init__()
        self._capture_hparams(data_shape=data_shape, time_range=time_range, num_categories=num_categories,
                              embedding_dim=embedding_dim, lambda_angle=lambda_angle, learning_rate=learning_rate,
                              lr_decay=lr_decay, **model_params)
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

    def flow_matching_loss(self, batch, noise1=None, X0=None, T=None):
        """The loss of training_step (:417-457) for given (or drawn: same order as the reference) noise.  With
        grad enabled and the module in ``train()`` mode the result carries the autograd graph through the B200
        network (``Unet3D``'s autograd bridge), so ``.backward()`` fills every ``p.grad``."""
        X1 = self.embed(batch)                                                   # :428
        n1 = torch.randn_like(X1) if noise1 is None else noise1
        with torch.cuda.device(X1.device):                                       # X1 + 1e-3 * randn (:429), in place
            _lib.check(_lib.lib.ftb_ode_axpy(_lib.ptr(X1), _lib.ptr(X1), _lib.ptr(n1.float().contiguous()), 1e-3,
                                             X1.numel(), None, 1, _lib.stream_ptr()))
        X0 = torch.randn_like(X1) if X0 is None else X0                          # :431
        if T is None:                                                            # :434-436
            T = torch.empty(X1.size(0), device=X1.device).uniform_(self.time_range[0], self.time_range[1])
        XT, VT = self.interpolator.flow_objective(T, X0, X1)                     # :439
        return flow_loss(VT, self.net(XT, T))                                    # :440-443

    def training_step(self, batch, batch_idx=None):
        """:417-457.  Lightning calls it with (batch, batch_idx); the reference's signature takes the batch only."""
        mse_loss = self.flow_matching_loss(batch)
        self._log_dict({"train_loss": mse_loss.detach()}, on_step=True, on_epoch=True, prog_bar=True, logger=True,
                       sync_dist=False)
        return mse_loss

    def configure_optimizers(self) -> Dict[str, Any]:
        """:465-473: Adam(lr) + ExponentialLR(gamma=lr_decay) over ``self.parameters()``."""
        optimizer = torch.optim.Adam(self.parameters(), lr=self.hparams.learning_rate)
        scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=self.hparams.lr_decay)
        return {"optimizer": optimizer, "lr_scheduler": scheduler}


class Geo3DStochInterpCond(_TaskHooks):
    """The conditional project's LightningModule
    (project/geodata-3d-conditional/model_train_sh_inference_cond.py:247-495) on the B200 path: ``net`` is the B200
    ``Unet3DCond``, the simplex ``embedding`` is frozen (:302), ``forward(x, ATb, t)``; ``conditioning(batch)`` is the
    fused embed + combined mask + ``ATb = X1 * mask`` front-end of training_step (:413-420); ``training_step`` /
    ``configure_optimizers`` / ``on_after_backward`` / checkpoint hooks as in :401-495.  ``CondFlowTrainer`` is the
    fused form of the same step."""

    def __init__(self, data_shape: Tuple[int, int, int] = (32, 32, 32),
                 time_range: List[float] = [0.0001, 0.9999], num_categories: int = 15, embedding_dim: int = 20,
                 lambda_reconstruct: float = 1.0, learning_rate: float = 2e-3, lr_decay: float = 0.997,
                 **model_params: Any):
        super().__init__()
        self._capture_hparams(data_shape=data_shape, time_range=time_range, num_categories=num_categories,
                              embedding_dim=embedding_dim, lambda_reconstruct=lambda_reconstruct,
                              learning_rate=learning_rate, lr_decay=lr_decay, **model_params)
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

    def cond_flow_loss(self, batch, noise1=None, X0=None, T=None, bores=None, n_bores=None, generator=None):
        """(loss, flow_loss, reconstruct_loss) of training_step (:413-452); differentiable through the B200 network
        in ``train()`` mode.  Draw order as the reference: boreholes, randn (X1 noise), randn (X0), uniform (T)."""
        from .training import cond_loss
        X1c, ATb, mask = self.conditioning(batch, bores, n_bores, generator)             # :413-420
        n1 = torch.randn_like(X1c) if noise1 is None else noise1
        X1 = torch.empty_like(X1c)
        with torch.cuda.device(X1c.device):                                              # :421
            _lib.check(_lib.lib.ftb_ode_axpy(_lib.ptr(X1), _lib.ptr(X1c), _lib.ptr(n1.float().contiguous()), 1e-4,
                                             X1.numel(), None, 1, _lib.stream_ptr()))
        X0 = torch.randn_like(X1) if X0 is None else X0                                  # :423
        if T is None:                                                                    # :426-428
            T = torch.empty(X1.size(0), device=X1.device).uniform_(self.time_range[0], self.time_range[1])
        XT, VT = self.interpolator.flow_objective(T, X0, X1)                             # :431
        VT_hat = self.net(XT, ATb, T)                                                    # :432
        return cond_loss(VT, VT_hat, XT, X1c, X1, mask, T, self.lambda_reconstruct)      # :434-452

    def training_step(self, batch, batch_idx=None):
        """:401-467."""
        loss, flow, rec = self.cond_flow_loss(batch)
        self._log_dict({"train_loss": loss.detach(), "flow_loss": flow, "reconstruct_loss": rec}, on_step=True,
                       on_epoch=True, prog_bar=True, logger=True, sync_dist=True)
        return loss

    def on_after_backward(self):
        """:476-485 without its per-parameter ``.item()`` host syncs: the total gradient norm stays a device scalar."""
        grads = [p.grad for p in self.parameters() if p.grad is not None]
        if grads:
            total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g.detach()) for g in grads]))
            self._log_dict({"grad_norm": total}, on_step=True, sync_dist=True)

    def configure_optimizers(self) -> Dict[str, Any]:
        """:487-495: AdamW(lr) + ExponentialLR(gamma=lr_decay)."""
        optimizer = torch.optim.AdamW(self.parameters(), lr=self.hparams.learning_rate)
        scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=self.hparams.lr_decay)
        return {"optimizer": optimizer, "lr_scheduler": scheduler}


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
                    param.copy_(shadow[name].to(param.device))   # in-place on the Parameter: bumps its version
    elif use_ema:
        print("WARNING: 'ema_shadow' not found in checkpoint. Using regular weights.")
    for m in module.modules():     # belt and braces: the engine re-reads every weight on its next forward
        if isinstance(m, Unet3D):
            m.mark_dirty()
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
