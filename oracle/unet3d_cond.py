"""Functional CPU restatement of the reference conditional UNet, Unet3DCond v3 (TEST INFRASTRUCTURE).

Reference: src/flowtrain/models/unet_attn_3d_cond_v3.py (line numbers cited per function);
shared pieces (ResnetBlock, attention, resampling, time MLP) come from oracle/unet3d.py, whose
reference code is identical in both files.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .synth import stage_plan
from .unet3d import (_tap, downsample, full_attention, linear_attention, resnet_block, rms_norm, time_mlp,
                     upsample)


def embed_atb(p, prefix, atb_opened, scale):
    """EmbedATb.forward — unet_attn_3d_cond_v3.py:131-139."""
    x = atb_opened
    if scale != 1.0:
        x = F.interpolate(x, scale_factor=scale, mode="trilinear", align_corners=True)
    x = F.conv3d(x, p[f"{prefix}.conv1.weight"], p[f"{prefix}.conv1.bias"], padding=2)
    x = F.silu(x)
    return F.conv3d(x, p[f"{prefix}.conv2.weight"], p[f"{prefix}.conv2.bias"], padding=2)


def mix_atb(p, prefix, x, atb, temb):
    """MixATb.forward — unet_attn_3d_cond_v3.py:175-190."""
    ax = torch.cat((x, atb), dim=1)
    t = F.linear(F.silu(temb), p[f"{prefix}.time_mlp.1.weight"], p[f"{prefix}.time_mlp.1.bias"])
    t = t[:, :, None, None, None]
    scale, shift = t.chunk(2, dim=1)
    ax = ax * (scale + 1) + shift
    h = F.conv3d(ax, p[f"{prefix}.conv1.weight"], p[f"{prefix}.conv1.bias"], padding=1)
    h = rms_norm(h, p[f"{prefix}.norm.g"])
    h = F.silu(h)
    h = F.conv3d(h, p[f"{prefix}.conv2.weight"], p[f"{prefix}.conv2.bias"], padding=1)
    return h + x


def unet3d_cond_forward(p, cfg, x, atb, time, taps=None):
    """Unet3DCond.forward — unet_attn_3d_cond_v3.py:769-828."""
    heads, dh = cfg["attn_heads"], cfg["attn_dim_head"]
    dims, in_out, full_attn = stage_plan(cfg)
    n = len(in_out)
    assert x.shape == atb.shape

    def attn(prefix, x, full):
        f = full_attention if full else linear_attention
        return _tap(taps, prefix, f(p, prefix, x, heads, dh, taps) + x)

    atb_opened = F.conv3d(atb, p["init_conv_ATb.weight"], p["init_conv_ATb.bias"], padding=3)
    _tap(taps, "init_conv_ATb", atb_opened)
    x = F.conv3d(x, p["init_conv_x.weight"], p["init_conv_x.bias"], padding=3)
    _tap(taps, "init_conv_x", x)
    r = x.clone()
    t = time_mlp(p, time)
    h = []
    for i in range(n):
        a = _tap(taps, f"downs.{i}.0", embed_atb(p, f"downs.{i}.0", atb_opened, 0.5 ** i))
        x = _tap(taps, f"downs.{i}.1", mix_atb(p, f"downs.{i}.1", x, a, t))
        x = resnet_block(p, f"downs.{i}.2", x, t, taps, mlp="time_mlp")
        h.append(x)
        x = resnet_block(p, f"downs.{i}.3", x, t, taps, mlp="time_mlp")
        x = attn(f"downs.{i}.4", x, full_attn[i])
        h.append(x)
        if i >= n - 1:
            x = F.conv3d(x, p[f"downs.{i}.5.weight"], p[f"downs.{i}.5.bias"], padding=1)
        else:
            x = downsample(p, f"downs.{i}.5", x)
        _tap(taps, f"downs.{i}.5", x)
    x = resnet_block(p, "mid_block1", x, t, taps, mlp="time_mlp")
    x = attn("mid_attn", x, True)
    x = resnet_block(p, "mid_block2", x, t, taps, mlp="time_mlp")
    for i in range(n):
        fa = full_attn[n - 1 - i]
        a = _tap(taps, f"ups.{i}.0", embed_atb(p, f"ups.{i}.0", atb_opened, 0.5 ** (n - i - 1)))
        x = _tap(taps, f"ups.{i}.1", mix_atb(p, f"ups.{i}.1", x, a, t))
        x = torch.cat((x, h.pop()), dim=1)
        x = resnet_block(p, f"ups.{i}.2", x, t, taps, mlp="time_mlp")
        x = torch.cat((x, h.pop()), dim=1)
        x = resnet_block(p, f"ups.{i}.3", x, t, taps, mlp="time_mlp")
        x = attn(f"ups.{i}.4", x, fa)
        if i == n - 1:
            x = F.conv3d(x, p[f"ups.{i}.5.weight"], p[f"ups.{i}.5.bias"], padding=1)
        else:
            x = upsample(p, f"ups.{i}.5", x)
        _tap(taps, f"ups.{i}.5", x)
    x = torch.cat((x, r), dim=1)
    x = resnet_block(p, "final_res_block", x, t, taps, mlp="time_mlp")
    return F.conv3d(x, p["final_conv.weight"], p["final_conv.bias"])
