"""Stochastic interpolants with flowtrain's API (src/flowtrain/interpolation/interpolation.py).

``StochasticInterpolator(interp).flow_objective(T, X0, X1, Z=None) -> (XT, BT)`` and the
``get_XT / get_BT / denoising_objective / get_ST / get_VT / get_BT_from_score`` helpers keep the
reference signatures and error behaviour (:60-76: ValueError when a two-sided interpolant gets no Z,
UserWarning when a one-sided one gets a Z).  XT/BT for CUDA fp32 tensors are built by ONE fused
kernel (``ftb_interp_xt_bt``) instead of 4-6 elementwise passes; there is no CPU fallback for
that construction.  The scalar schedules alpha/beta/gamma(t) stay as tiny host-side torch
expressions because the solvers call them on 0-d time tensors (solvers.py:138-141).
"""
from __future__ import annotations

import math
import warnings
from abc import ABC, abstractmethod

import torch

from . import _lib

_KIND = {"linear": 0, "trig": 1, "encdec": 2, "sbdm": 3, "mirror": 4}


class BaseInterpolant(ABC):
    kind = None

    def __init__(self, one_sided=False):
        self.one_sided = one_sided
        self.gamma_a = 2.0

    def __repr__(self):
        return f"{type(self).__name__}(one_sided={self.one_sided})"

    __str__ = __repr__

    def is_one_sided(self):
        return self.one_sided

    @abstractmethod
    def alpha(self, t): ...
    @abstractmethod
    def beta(self, t): ...
    @abstractmethod
    def gamma(self, t): ...
    @abstractmethod
    def alpha_dot(self, t): ...
    @abstractmethod
    def beta_dot(self, t): ...
    @abstractmethod
    def gamma_dot(self, t): ...


def _bridge(a, t):
    return torch.sqrt(a * t * (1 - t))


def _bridge_dot(a, t):
    return 0.5 * a * (1 - 2 * t) / torch.sqrt(a * t * (1 - t))


class LinearInterpolant(BaseInterpolant):
    """alpha = 1-t, beta = t, gamma = sqrt(a t (1-t)) (zero when one-sided) — :379-412."""
    kind = "linear"

    def __init__(self, one_sided=False, gamma_a=2.0):
        super().__init__(one_sided)
        self.gamma_a = gamma_a

    def alpha(self, t): return 1 - t
    def beta(self, t): return t
    def gamma(self, t): return torch.zeros_like(t) if self.one_sided else _bridge(self.gamma_a, t)
    def alpha_dot(self, t): return -torch.ones_like(t)
    def beta_dot(self, t): return torch.ones_like(t)
    def gamma_dot(self, t): return torch.zeros_like(t) if self.one_sided else _bridge_dot(self.gamma_a, t)


class TrigInterpolant(BaseInterpolant):
    """alpha = cos(pi t/2), beta = sin(pi t/2) — :415-449."""
    kind = "trig"

    def __init__(self, one_sided=False, gamma_a=2.0):
        super().__init__(one_sided)
        self.gamma_a = gamma_a

    def alpha(self, t): return torch.cos(torch.pi * t / 2)
    def beta(self, t): return torch.sin(torch.pi * t / 2)
    def gamma(self, t): return torch.zeros_like(t) if self.one_sided else _bridge(self.gamma_a, t)
    def alpha_dot(self, t): return -torch.pi / 2 * torch.sin(torch.pi * t / 2)
    def beta_dot(self, t): return torch.pi / 2 * torch.cos(torch.pi * t / 2)
    def gamma_dot(self, t): return torch.zeros_like(t) if self.one_sided else _bridge_dot(self.gamma_a, t)


class EncDecInterpolant(BaseInterpolant):
    """cos^2(pi t) gated at t = 1/2, gamma = sin^2(pi t) — :452-484."""
    kind = "encdec"

    def __init__(self):
        super().__init__(one_sided=False)

    def alpha(self, t): return torch.where(t < 0.5, torch.cos(torch.pi * t) ** 2, torch.zeros_like(t))
    def beta(self, t): return torch.where(t > 0.5, torch.cos(torch.pi * t) ** 2, torch.zeros_like(t))
    def gamma(self, t): return torch.sin(torch.pi * t) ** 2
    def alpha_dot(self, t): return torch.where(t < 0.5, -torch.pi * torch.sin(2 * torch.pi * t), torch.zeros_like(t))
    def beta_dot(self, t): return torch.where(t > 0.5, -torch.pi * torch.sin(2 * torch.pi * t), torch.zeros_like(t))
    def gamma_dot(self, t): return torch.pi * torch.sin(2 * torch.pi * t)


class SBDMInterpolant(BaseInterpolant):
    """alpha = sqrt(1-t^2), beta = t, one-sided — :487-514."""
    kind = "sbdm"

    def __init__(self):
        super().__init__(one_sided=True)

    def alpha(self, t): return torch.sqrt(1 - t ** 2)
    def beta(self, t): return t
    def gamma(self, t): return torch.zeros_like(t)
    def alpha_dot(self, t): return -t / torch.sqrt(1 - t ** 2)
    def beta_dot(self, t): return torch.ones_like(t)
    def gamma_dot(self, t): return torch.zeros_like(t)


class MirrorInterpolant(BaseInterpolant):
    """alpha = 0, beta = 1, gamma = sqrt(a t (1-t)) — :517-546."""
    kind = "mirror"

    def __init__(self, gamma_a=2.0):
        super().__init__(one_sided=False)
        self.gamma_a = gamma_a

    def alpha(self, t): return torch.zeros_like(t)
    def beta(self, t): return torch.ones_like(t)
    def gamma(self, t): return _bridge(self.gamma_a, t)
    def alpha_dot(self, t): return torch.zeros_like(t)
    def beta_dot(self, t): return torch.zeros_like(t)
    def gamma_dot(self, t): return _bridge_dot(self.gamma_a, t)


def _fused_xt_bt(interp, T, X0, X1, Z, want_bt=True):
    """One kernel: XT = a X0 + b X1 (+ g Z), BT = a' X0 + b' X1 (+ g' Z)."""
    if interp.kind not in _KIND:
        raise NotImplementedError(f"no fused kernel for {type(interp).__name__}")
    if not (X0.is_cuda and X1.is_cuda):
        raise RuntimeError("interpolant construction runs on CUDA only (no CPU fallback)")
    assert X0.shape == X1.shape, "Shapes of X0 and X1 must match"
    if Z is not None:
        assert Z.shape == X0.shape, "Shape of Z must match X0 and X1"
    B = X0.shape[0]
    n = X0[0].numel()
    x0 = X0.detach().float().contiguous()
    x1 = X1.detach().float().contiguous()
    z = None if Z is None else Z.detach().float().contiguous()
    t = T.detach().reshape(-1).to(device=X0.device, dtype=torch.float32).contiguous()
    if t.numel() == 1 and B > 1:
        t = t.expand(B).contiguous()
    if t.numel() != B:
        raise ValueError(f"T must have one entry per sample ({B}), got {t.numel()}")
    xt = torch.empty_like(x0)
    bt = torch.empty_like(x0) if want_bt else None
    with torch.cuda.device(X0.device):
        if n % 4 == 0:
            _lib.check(_lib.lib.ftb_interp_xt_bt(
                _KIND[interp.kind], int(bool(interp.one_sided)), float(getattr(interp, "gamma_a", 2.0)),
                _lib.ptr(x0), _lib.ptr(x1), _lib.ptr(z), _lib.ptr(t), _lib.ptr(xt), _lib.ptr(bt), B, n,
                _lib.stream_ptr()))
        else:
            raise ValueError("per-sample element count must be a multiple of 4")
    return xt, bt


class StochasticInterpolator:
    """Drop-in for flowtrain.interpolation.StochasticInterpolator (:43-276)."""

    def __init__(self, interpolant):
        self.interp = interpolant

    def __repr__(self):
        return f"StochasticInterpolator({self.interp})"

    __str__ = __repr__

    def _check_z(self, Z):
        if not self.interp.one_sided and Z is None:
            raise ValueError("Z must be provided for two-sided interpolants")
        if self.interp.one_sided and Z is not None:
            warnings.warn("Z was provided for a one-sided interpolant which does not use it", UserWarning)

    def flow_objective(self, T, X0, X1, Z=None):
        self._check_z(Z)
        # get_XT/get_BT add gamma*Z whenever Z is passed (:181-184), even for one-sided (gamma = 0)
        return _fused_xt_bt(self.interp, T, X0, X1, Z, want_bt=True)

    def denoising_objective(self, T, X0, X1, Z=None):
        self._check_z(Z)
        XT, _ = _fused_xt_bt(self.interp, T, X0, X1, Z, want_bt=False)
        return XT, (X0 if self.interp.one_sided else Z)

    def get_XT(self, T, X0, X1, Z=None):
        self._check_z(Z)
        return _fused_xt_bt(self.interp, T, X0, X1, Z, want_bt=False)[0]

    def get_BT(self, T, X0, X1, Z=None):
        self._check_z(Z)
        return _fused_xt_bt(self.interp, T, X0, X1, Z, want_bt=True)[1]

    def get_VT(self, T, X0, X1):
        return _fused_xt_bt(self.interp, T, X0, X1, None, want_bt=True)[1]

    # the two helpers below are off the hot path (never called by the shipped scripts); they keep
    # the reference semantics with plain tensor expressions
    @staticmethod
    def _bt(T, X):
        return T.view(T.shape[0], *([1] * (X.dim() - 1))) if T.dim() == 1 else T

    def get_ST(self, T, Z):
        T = self._bt(T, Z)
        g = self.interp.alpha(T) if self.interp.one_sided else self.interp.gamma(T)
        return -(g ** (-1)) * Z

    def get_BT_from_score(self, T, VT, ST):
        T = self._bt(T, VT)
        return VT - self.interp.gamma_dot(T) * self.interp.gamma(T) * ST


__all__ = [
    "BaseInterpolant", "LinearInterpolant", "TrigInterpolant", "EncDecInterpolant", "SBDMInterpolant",
    "MirrorInterpolant", "StochasticInterpolator",
]
