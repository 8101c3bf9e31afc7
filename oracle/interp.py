"""Stochastic-interpolant coefficient schedules and objectives (TEST INFRASTRUCTURE).

Restates src/flowtrain/interpolation/interpolation.py in plain torch:
  * alpha/beta/gamma and time derivatives of the five interpolants (:379-546)
  * XT / BT construction (:156-216), flow and denoising objectives (:78-154)
  * score / velocity helpers (:218-276)
``coeffs(kind, t, ...)`` returns the six scalars the fused CUDA kernel consumes.
"""
from __future__ import annotations

import math

import torch

KINDS = ("linear", "trig", "encdec", "sbdm", "mirror")


def coeffs(kind, t, one_sided=False, gamma_a=2.0):
    """(alpha, beta, gamma, alpha_dot, beta_dot, gamma_dot) as tensors shaped like t."""
    t = torch.as_tensor(t)
    zero, one = torch.zeros_like(t), torch.ones_like(t)
    pi = math.pi
    if kind == "linear":  # LinearInterpolant :379-412
        a, b, ad, bd = 1 - t, t, -one, one
        if one_sided:
            g, gd = zero, zero
        else:
            g = torch.sqrt(gamma_a * t * (1 - t))
            gd = 0.5 * gamma_a * (1 - 2 * t) / torch.sqrt(gamma_a * t * (1 - t))
    elif kind == "trig":  # TrigInterpolant :415-449
        a, b = torch.cos(pi * t / 2), torch.sin(pi * t / 2)
        ad, bd = -pi / 2 * torch.sin(pi * t / 2), pi / 2 * torch.cos(pi * t / 2)
        if one_sided:
            g, gd = zero, zero
        else:
            g = torch.sqrt(gamma_a * t * (1 - t))
            gd = 0.5 * gamma_a * (1 - 2 * t) / torch.sqrt(gamma_a * t * (1 - t))
    elif kind == "encdec":  # EncDecInterpolant :452-484 (always two-sided)
        c2 = torch.cos(pi * t) ** 2
        a = torch.where(t < 0.5, c2, zero)
        b = torch.where(t > 0.5, c2, zero)
        g = torch.sin(pi * t) ** 2
        s2 = -pi * torch.sin(2 * pi * t)
        ad = torch.where(t < 0.5, s2, zero)
        bd = torch.where(t > 0.5, s2, zero)
        gd = pi * torch.sin(2 * pi * t)
    elif kind == "sbdm":  # SBDMInterpolant :487-514 (always one-sided)
        a, b, g = torch.sqrt(1 - t ** 2), t, zero
        ad, bd, gd = -t / torch.sqrt(1 - t ** 2), one, zero
    elif kind == "mirror":  # MirrorInterpolant :517-546 (always two-sided)
        a, b = zero, one
        g = torch.sqrt(gamma_a * t * (1 - t))
        ad, bd = zero, zero
        gd = 0.5 * gamma_a * (1 - 2 * t) / torch.sqrt(gamma_a * t * (1 - t))
    else:
        raise ValueError(kind)
    return a, b, g, ad, bd, gd


def is_one_sided(kind, one_sided=False):
    if kind == "sbdm":
        return True
    if kind in ("encdec", "mirror"):
        return False
    return bool(one_sided)


def _bt(T, X):  # reshape_time :27-40
    return T.view(T.shape[0], *([1] * (X.dim() - 1))) if T.dim() == 1 else T


def get_xt(kind, T, X0, X1, Z=None, one_sided=False, gamma_a=2.0):
    """get_XT :156-185."""
    T = _bt(T, X0)
    a, b, g, *_ = coeffs(kind, T, one_sided, gamma_a)
    XT = a * X0 + b * X1
    if Z is not None:
        XT = XT + g * Z
    return XT


def get_bt(kind, T, X0, X1, Z=None, one_sided=False, gamma_a=2.0):
    """get_BT :187-216."""
    T = _bt(T, X0)
    _, _, _, ad, bd, gd = coeffs(kind, T, one_sided, gamma_a)
    BT = ad * X0 + bd * X1
    if Z is not None:
        BT = BT + gd * Z
    return BT


def flow_objective(kind, T, X0, X1, Z=None, one_sided=False, gamma_a=2.0):
    """flow_objective :78-117 -> (XT, BT)."""
    if not is_one_sided(kind, one_sided) and Z is None:
        raise ValueError("Z must be provided for two-sided interpolants")
    return (get_xt(kind, T, X0, X1, Z, one_sided, gamma_a),
            get_bt(kind, T, X0, X1, Z, one_sided, gamma_a))


def denoising_objective(kind, T, X0, X1, Z=None, one_sided=False, gamma_a=2.0):
    """denoising_objective :119-154 -> (XT, target noise)."""
    XT = get_xt(kind, T, X0, X1, Z, one_sided, gamma_a)
    return XT, (X0 if is_one_sided(kind, one_sided) else Z)


def get_st(kind, T, Z, one_sided=False, gamma_a=2.0):
    """get_ST :226-251: score = -Z / gamma (alpha for one-sided)."""
    T = _bt(T, Z)
    a, _, g, *_ = coeffs(kind, T, one_sided, gamma_a)
    return -((a if is_one_sided(kind, one_sided) else g) ** (-1)) * Z


def get_bt_from_score(kind, T, VT, ST, one_sided=False, gamma_a=2.0):
    """get_BT_from_score :218-224."""
    T = _bt(T, VT)
    _, _, g, _, _, gd = coeffs(kind, T, one_sided, gamma_a)
    return VT - gd * g * ST
