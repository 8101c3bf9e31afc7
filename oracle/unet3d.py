"""Functional CPU restatement of the reference Unet3D velocity field (TEST INFRASTRUCTURE).

Plain torch ops on a ``{name: tensor}`` parameter dict (the reference ``state_dict()``
keys), fp32 by default.  Works on any torch device, so the same code is the CPU oracle in
unit tests and — run on the B200 with TF32 disabled — the 64^3 parity reference.

Reference: src/flowtrain/models/unet_attn_3d.py (line numbers cited per function).
An optional ``taps`` dict records every intermediate tensor so a failing kernel can be
localised layer by layer.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .synth import stage_plan


def _tap(taps, name, x):
    if taps is not None:
        taps[name] = x.detach()
    return x


def rms_norm(x, g):
    """RMSNorm.forward — unet_attn_3d.py:127-128: L2-normalise over C (eps 1e-12 clamp on
    the norm, F.normalize default), times g, times sqrt(C)."""
    return F.normalize(x, dim=1) * g * (x.shape[1] ** 0.5)


def fourier_time_embedding(t, freqs, phases):
    """RandomFourierEmbedding.forward — unet_attn_3d.py:203-208."""
    y = torch.outer(t, freqs)
    y = y + phases
    return y.cos() * math.sqrt(2)


def time_mlp(p, t):
    """time_mlp Sequential — unet_attn_3d.py:551-556 (Fourier -> Linear -> GELU(erf) -> Linear)."""
    y = fourier_time_embedding(t, p["time_mlp.0.freqs"], p["time_mlp.0.phases"])
    y = F.linear(y, p["time_mlp.1.weight"], p["time_mlp.1.bias"])
    y = F.gelu(y)
    return F.linear(y, p["time_mlp.3.weight"], p["time_mlp.3.bias"])


def block(p, prefix, x, scale_shift=None):
    """Block.forward — unet_attn_3d.py:232-244 (dropout is identity: eval / p=0)."""
    x = F.conv3d(x, p[f"{prefix}.proj.weight"], p[f"{prefix}.proj.bias"], padding=1)
    x = rms_norm(x, p[f"{prefix}.norm.g"])
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def resnet_block(p, prefix, x, temb, taps=None, mlp="mlp"):
    """ResnetBlock.forward — unet_attn_3d.py:265-278 (``mlp``); the conditional file names the
    same Sequential ``time_mlp`` (unet_attn_3d_cond_v3.py:337, :347-361)."""
    te = F.linear(F.silu(temb), p[f"{prefix}.{mlp}.1.weight"], p[f"{prefix}.{mlp}.1.bias"])
    te = te[:, :, None, None, None]
    scale_shift = te.chunk(2, dim=1)
    h = block(p, f"{prefix}.block1", x, scale_shift)
    _tap(taps, f"{prefix}.block1", h)
    h = block(p, f"{prefix}.block2", h)
    if f"{prefix}.res_conv.weight" in p:
        res = F.conv3d(x, p[f"{prefix}.res_conv.weight"], p[f"{prefix}.res_conv.bias"])
    else:
        res = x
    return _tap(taps, prefix, h + res)


def linear_attention(p, prefix, x, heads, dim_head, taps=None):
    """LinearAttention.forward — unet_attn_3d.py:308-341."""
    b, c, X, Y, Z = x.shape
    n = X * Y * Z
    xn = rms_norm(x, p[f"{prefix}.norm.g"])
    qkv = F.conv3d(xn, p[f"{prefix}.to_qkv.weight"])
    _tap(taps, f"{prefix}.qkv", qkv)
    q, k, v = (t.reshape(b, heads, dim_head, n) for t in qkv.chunk(3, dim=1))
    mk, mv = (m[None].expand(b, -1, -1, -1) for m in p[f"{prefix}.mem_kv"])
    k = torch.cat((mk, k), dim=-1)
    v = torch.cat((mv, v), dim=-1)
    q = q.softmax(dim=-2) * dim_head ** -0.5
    k = k.softmax(dim=-1)
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    _tap(taps, f"{prefix}.context", context)
    out = torch.einsum("bhde,bhdn->bhen", context, q)
    out = out.reshape(b, heads * dim_head, X, Y, Z)
    out = F.conv3d(out, p[f"{prefix}.to_out.0.weight"], p[f"{prefix}.to_out.0.bias"])
    return rms_norm(out, p[f"{prefix}.to_out.1.g"])


def full_attention(p, prefix, x, heads, dim_head, taps=None):
    """Attention.forward + Attend.forward math path — unet_attn_3d.py:357-373, :436-465."""
    b, c, X, Y, Z = x.shape
    n = X * Y * Z
    xn = rms_norm(x, p[f"{prefix}.norm.g"])
    qkv = F.conv3d(xn, p[f"{prefix}.to_qkv.weight"])
    q, k, v = (t.reshape(b, heads, dim_head, n).transpose(-1, -2) for t in qkv.chunk(3, dim=1))
    mk, mv = (m[None].expand(b, -1, -1, -1) for m in p[f"{prefix}.mem_kv"])
    k = torch.cat((mk, k), dim=-2)
    v = torch.cat((mv, v), dim=-2)
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * dim_head ** -0.5
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v)
    _tap(taps, f"{prefix}.attn_out", out)
    out = out.transpose(-1, -2).reshape(b, heads * dim_head, X, Y, Z)
    return F.conv3d(out, p[f"{prefix}.to_out.weight"], p[f"{prefix}.to_out.bias"])


def downsample(p, prefix, x):
    """Downsample.forward — unet_attn_3d.py:105-108."""
    x = F.interpolate(x, scale_factor=0.5, mode="trilinear", align_corners=True)
    return F.conv3d(x, p[f"{prefix}.conv.weight"], p[f"{prefix}.conv.bias"])


def upsample(p, prefix, x):
    """Upsample.forward — unet_attn_3d.py:85-88."""
    x = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True)
    return F.conv3d(x, p[f"{prefix}.conv.weight"], p[f"{prefix}.conv.bias"], padding=1)


def unet3d_forward(p, cfg, x, time, taps=None):
    """Unet3D.forward — unet_attn_3d.py:673-719."""
    heads, dh = cfg["attn_heads"], cfg["attn_dim_head"]
    dims, in_out, full_attn = stage_plan(cfg)
    n = len(in_out)

    def attn(prefix, x, full):
        f = full_attention if full else linear_attention
        return _tap(taps, prefix, f(p, prefix, x, heads, dh, taps) + x)

    x = F.conv3d(x, p["init_conv.weight"], p["init_conv.bias"], padding=3)
    _tap(taps, "init_conv", x)
    r = x.clone()
    t = _tap(taps, "time_mlp", time_mlp(p, time))
    h = []
    for i in range(n):
        x = resnet_block(p, f"downs.{i}.0", x, t, taps)
        h.append(x)
        x = resnet_block(p, f"downs.{i}.1", x, t, taps)
        x = attn(f"downs.{i}.2", x, full_attn[i])
        h.append(x)
        if i >= n - 1:
            x = F.conv3d(x, p[f"downs.{i}.3.weight"], p[f"downs.{i}.3.bias"], padding=1)
        else:
            x = downsample(p, f"downs.{i}.3", x)
        _tap(taps, f"downs.{i}.3", x)
    x = resnet_block(p, "mid_block1", x, t, taps)
    x = attn("mid_attn", x, True)
    x = resnet_block(p, "mid_block2", x, t, taps)
    for i in range(n):
        fa = full_attn[n - 1 - i]
        x = torch.cat((x, h.pop()), dim=1)
        x = resnet_block(p, f"ups.{i}.0", x, t, taps)
        x = torch.cat((x, h.pop()), dim=1)
        x = resnet_block(p, f"ups.{i}.1", x, t, taps)
        x = attn(f"ups.{i}.2", x, fa)
        if i == n - 1:
            x = F.conv3d(x, p[f"ups.{i}.3.weight"], p[f"ups.{i}.3.bias"], padding=1)
        else:
            x = upsample(p, f"ups.{i}.3", x)
        _tap(taps, f"ups.{i}.3", x)
    x = torch.cat((x, r), dim=1)
    x = resnet_block(p, "final_res_block", x, t, taps)
    return F.conv3d(x, p["final_conv.weight"], p["final_conv.bias"])


class OracleVelocity:
    """Callable ``(x[B,C,X,Y,Z], t[B]) -> v`` over the oracle forward; the CPU stand-in for
    ``model(XT, T)`` at the drop-in boundary (solvers.py:70)."""

    def __init__(self, params, cfg, device="cpu", dtype=torch.float32):
        self.cfg = cfg
        self.p = {k: v.to(device=device, dtype=dtype) for k, v in params.items()}

    @torch.no_grad()
    def __call__(self, x, t):
        return unet3d_forward(self.p, self.cfg, x, t)
