"""Deterministic synthetic parameters for the Unet3D velocity network (TEST INFRASTRUCTURE).

The reference publishes no weights that are reachable offline (readme.md:27-33), so
parity runs use random-init weights.  torch's RNG stream depends on module construction
order, so instead every tensor is synthesised from a counter-based hash of
(seed, parameter name): the same call gives bit-identical weights in the build container
(where the real reference module is loaded with them to make the golden fixtures) and on
the GPU box (where /root/reference does not exist).

``unet3d_param_specs`` restates the parameter names and shapes of
``Unet3D.__init__`` (src/flowtrain/models/unet_attn_3d.py:509-667); the test-suite checks
the list against the real module's ``state_dict()`` in this container.
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict

import numpy as np
import torch

DEFAULT_CFG = dict(
    dim=48,
    dim_mults=(1, 2, 2, 3, 4),
    data_channels=18,
    dropout=0.0,
    self_condition=False,
    time_resolution=1024,
    time_sin_pos=False,
    time_bandwidth=1000.0,
    time_learned_emb=True,
    attn_enabled=True,
    attn_dim_head=32,
    attn_heads=4,
    full_attn=None,
    flash_attn=False,
)


def make_cfg(**kw):
    cfg = dict(DEFAULT_CFG)
    cfg.update(kw)
    cfg["dim_mults"] = tuple(cfg["dim_mults"])
    return cfg


def stage_plan(cfg):
    """dims / (dim_in, dim_out) pairs / full-attention flags — unet_attn_3d.py:538-567."""
    dim = cfg["dim"]
    mults = tuple(cfg["dim_mults"])
    dims = [dim] + [dim * m for m in mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    full_attn = cfg.get("full_attn")
    if not full_attn:
        full_attn = (False,) * (len(mults) - 1) + (True,)
    elif not isinstance(full_attn, tuple):
        full_attn = (full_attn,) * len(mults)
    return dims, in_out, tuple(full_attn)


def _resnet_specs(prefix, cin, cout, time_dim, mlp="mlp"):
    """ResnetBlock parameters; the time MLP attribute is ``mlp`` in unet_attn_3d.py:255 and
    ``time_mlp`` in unet_attn_3d_cond_v3.py:337."""
    s = OrderedDict()
    s[f"{prefix}.{mlp}.1.weight"] = (2 * cout, time_dim)
    s[f"{prefix}.{mlp}.1.bias"] = (2 * cout,)
    s[f"{prefix}.block1.proj.weight"] = (cout, cin, 3, 3, 3)
    s[f"{prefix}.block1.proj.bias"] = (cout,)
    s[f"{prefix}.block1.norm.g"] = (1, cout, 1, 1, 1)
    s[f"{prefix}.block2.proj.weight"] = (cout, cout, 3, 3, 3)
    s[f"{prefix}.block2.proj.bias"] = (cout,)
    s[f"{prefix}.block2.norm.g"] = (1, cout, 1, 1, 1)
    if cin != cout:
        s[f"{prefix}.res_conv.weight"] = (cout, cin, 1, 1, 1)
        s[f"{prefix}.res_conv.bias"] = (cout,)
    return s


def _attn_specs(prefix, dim, heads, dim_head, full, num_mem_kv=4):
    hidden = heads * dim_head
    s = OrderedDict()
    if full:  # Attention — unet_attn_3d.py:345-355
        s[f"{prefix}.mem_kv"] = (2, heads, num_mem_kv, dim_head)
        s[f"{prefix}.norm.g"] = (1, dim, 1, 1, 1)
        s[f"{prefix}.to_qkv.weight"] = (hidden * 3, dim, 1, 1, 1)
        s[f"{prefix}.to_out.weight"] = (dim, hidden, 1, 1, 1)
        s[f"{prefix}.to_out.bias"] = (dim,)
    else:  # LinearAttention — unet_attn_3d.py:285-306
        s[f"{prefix}.mem_kv"] = (2, heads, dim_head, num_mem_kv)
        s[f"{prefix}.norm.g"] = (1, dim, 1, 1, 1)
        s[f"{prefix}.to_qkv.weight"] = (hidden * 3, dim, 1, 1, 1)
        s[f"{prefix}.to_out.0.weight"] = (dim, hidden, 1, 1, 1)
        s[f"{prefix}.to_out.0.bias"] = (dim,)
        s[f"{prefix}.to_out.1.g"] = (1, dim, 1, 1, 1)
    return s


def unet3d_param_specs(cfg) -> "OrderedDict[str, tuple]":
    """name -> shape in ``state_dict()`` order of the reference Unet3D."""
    assert not cfg.get("self_condition", False), "self_condition is out of scope"
    assert not cfg.get("time_sin_pos", False), "only Fourier time embeddings are in scope"
    assert cfg.get("attn_enabled", True)
    dim = cfg["dim"]
    C = cfg["data_channels"]
    tr = cfg["time_resolution"]
    time_dim = dim * 4
    heads, dh = cfg["attn_heads"], cfg["attn_dim_head"]
    dims, in_out, full_attn = stage_plan(cfg)
    n = len(in_out)

    s = OrderedDict()
    s["init_conv.weight"] = (dim, C, 7, 7, 7)
    s["init_conv.bias"] = (dim,)
    s["time_mlp.0.freqs"] = (tr,)
    s["time_mlp.0.phases"] = (tr,)
    s["time_mlp.1.weight"] = (time_dim, tr)
    s["time_mlp.1.bias"] = (time_dim,)
    s["time_mlp.3.weight"] = (time_dim, time_dim)
    s["time_mlp.3.bias"] = (time_dim,)
    for i, ((din, dout), fa) in enumerate(zip(in_out, full_attn)):
        last = i >= n - 1
        s.update(_resnet_specs(f"downs.{i}.0", din, din, time_dim))
        s.update(_resnet_specs(f"downs.{i}.1", din, din, time_dim))
        s.update(_attn_specs(f"downs.{i}.2", din, heads, dh, fa))
        if last:
            s[f"downs.{i}.3.weight"] = (dout, din, 3, 3, 3)
            s[f"downs.{i}.3.bias"] = (dout,)
        else:
            s[f"downs.{i}.3.conv.weight"] = (dout, din, 1, 1, 1)
            s[f"downs.{i}.3.conv.bias"] = (dout,)
    # NB: state_dict order follows attribute registration order: downs, ups, mid_*, final_*
    ups = OrderedDict()
    for i, ((din, dout), fa) in enumerate(zip(reversed(in_out), reversed(full_attn))):
        last = i == n - 1
        ups.update(_resnet_specs(f"ups.{i}.0", dout + din, dout, time_dim))
        ups.update(_resnet_specs(f"ups.{i}.1", dout + din, dout, time_dim))
        ups.update(_attn_specs(f"ups.{i}.2", dout, heads, dh, fa))
        if last:
            ups[f"ups.{i}.3.weight"] = (din, dout, 3, 3, 3)
            ups[f"ups.{i}.3.bias"] = (din,)
        else:
            ups[f"ups.{i}.3.conv.weight"] = (din, dout, 3, 3, 3)
            ups[f"ups.{i}.3.conv.bias"] = (din,)
    s.update(ups)
    mid = dims[-1]
    s.update(_resnet_specs("mid_block1", mid, mid, time_dim))
    s.update(_attn_specs("mid_attn", mid, heads, dh, True))
    s.update(_resnet_specs("mid_block2", mid, mid, time_dim))
    s.update(_resnet_specs("final_res_block", dim * 2, dim, time_dim))
    s["final_conv.weight"] = (C, dim, 1, 1, 1)
    s["final_conv.bias"] = (C,)
    return s


def unet3d_cond_param_specs(cfg) -> "OrderedDict[str, tuple]":
    """name -> shape in ``state_dict()`` order of the reference Unet3DCond v3
    (src/flowtrain/models/unet_attn_3d_cond_v3.py:598-764): init_conv_x, init_conv_ATb, time_mlp,
    per stage [EmbedATb, MixATb, ResnetBlock, ResnetBlock, attention, down/up-sample]."""
    assert not cfg.get("self_condition", False) and not cfg.get("time_sin_pos", False)
    dim = cfg["dim"]
    C = cfg["data_channels"]
    tr = cfg["time_resolution"]
    time_dim = dim * 4
    heads, dh = cfg["attn_heads"], cfg["attn_dim_head"]
    dims, in_out, full_attn = stage_plan(cfg)
    n = len(in_out)

    def embed_mix(prefix, d):
        e = OrderedDict()
        e[f"{prefix}.0.conv1.weight"] = (d, C, 5, 5, 5)      # EmbedATb :127-129
        e[f"{prefix}.0.conv1.bias"] = (d,)
        e[f"{prefix}.0.conv2.weight"] = (d, d, 5, 5, 5)
        e[f"{prefix}.0.conv2.bias"] = (d,)
        e[f"{prefix}.1.time_mlp.1.weight"] = (4 * d, time_dim)  # MixATb :164-173
        e[f"{prefix}.1.time_mlp.1.bias"] = (4 * d,)
        e[f"{prefix}.1.conv1.weight"] = (d, 2 * d, 3, 3, 3)
        e[f"{prefix}.1.conv1.bias"] = (d,)
        e[f"{prefix}.1.norm.g"] = (1, d, 1, 1, 1)
        e[f"{prefix}.1.conv2.weight"] = (d, d, 3, 3, 3)
        e[f"{prefix}.1.conv2.bias"] = (d,)
        return e

    s = OrderedDict()
    s["init_conv_x.weight"] = (dim, C, 7, 7, 7)
    s["init_conv_x.bias"] = (dim,)
    s["init_conv_ATb.weight"] = (C, C, 7, 7, 7)
    s["init_conv_ATb.bias"] = (C,)
    s["time_mlp.0.freqs"] = (tr,)
    s["time_mlp.0.phases"] = (tr,)
    s["time_mlp.1.weight"] = (time_dim, tr)
    s["time_mlp.1.bias"] = (time_dim,)
    s["time_mlp.3.weight"] = (time_dim, time_dim)
    s["time_mlp.3.bias"] = (time_dim,)
    for i, ((din, dout), fa) in enumerate(zip(in_out, full_attn)):
        last = i >= n - 1
        s.update(embed_mix(f"downs.{i}", din))
        s.update(_resnet_specs(f"downs.{i}.2", din, din, time_dim, "time_mlp"))
        s.update(_resnet_specs(f"downs.{i}.3", din, din, time_dim, "time_mlp"))
        s.update(_attn_specs(f"downs.{i}.4", din, heads, dh, fa))
        if last:
            s[f"downs.{i}.5.weight"] = (dout, din, 3, 3, 3)
            s[f"downs.{i}.5.bias"] = (dout,)
        else:
            s[f"downs.{i}.5.conv.weight"] = (dout, din, 1, 1, 1)
            s[f"downs.{i}.5.conv.bias"] = (dout,)
    for i, ((din, dout), fa) in enumerate(zip(reversed(in_out), reversed(full_attn))):
        last = i == n - 1
        s.update(embed_mix(f"ups.{i}", dout))
        s.update(_resnet_specs(f"ups.{i}.2", dout + din, dout, time_dim, "time_mlp"))
        s.update(_resnet_specs(f"ups.{i}.3", dout + din, dout, time_dim, "time_mlp"))
        s.update(_attn_specs(f"ups.{i}.4", dout, heads, dh, fa))
        if last:
            s[f"ups.{i}.5.weight"] = (din, dout, 3, 3, 3)
            s[f"ups.{i}.5.bias"] = (din,)
        else:
            s[f"ups.{i}.5.conv.weight"] = (din, dout, 3, 3, 3)
            s[f"ups.{i}.5.conv.bias"] = (din,)
    mid = dims[-1]
    s.update(_resnet_specs("mid_block1", mid, mid, time_dim, "time_mlp"))
    s.update(_attn_specs("mid_attn", mid, heads, dh, True))
    s.update(_resnet_specs("mid_block2", mid, mid, time_dim, "time_mlp"))
    s.update(_resnet_specs("final_res_block", dim * 2, dim, time_dim, "time_mlp"))
    s["final_conv.weight"] = (C, dim, 1, 1, 1)
    s["final_conv.bias"] = (C,)
    return s


# ---------------------------------------------------------------------------------------
# counter-based value synthesis
# ---------------------------------------------------------------------------------------
def _uniform01(seed: int, name: str, n: int) -> np.ndarray:
    """n doubles in (0,1) from a Philox stream keyed by sha256(seed, name)."""
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    key = int.from_bytes(h[:16], "little")
    bg = np.random.Philox(key=key)
    raw = bg.random_raw(n).astype(np.uint64)
    return ((raw >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def _normal(seed: int, name: str, n: int) -> np.ndarray:
    u1 = _uniform01(seed, name + "#a", n)
    u2 = _uniform01(seed, name + "#b", n)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * math.pi * u2)


def synth_param(seed: int, name: str, shape, cfg) -> torch.Tensor:
    """Init statistics follow torch defaults used by the reference ctor (kaiming-uniform
    bound 1/sqrt(fan_in) for conv/linear weight and bias; mem_kv ~ N(0,1) :300,:353;
    freqs ~ N(0,1)*bandwidth, phases ~ U(0,1) :217-218).  RMSNorm gains are perturbed
    around 1 (reference init is exactly 1, :125) so that a kernel ignoring g fails."""
    n = int(np.prod(shape))
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "g":
        v = 1.0 + 0.2 * (_uniform01(seed, name, n) - 0.5)
    elif leaf == "mem_kv":
        v = _normal(seed, name, n)
    elif leaf == "freqs":
        v = _normal(seed, name, n) * float(cfg["time_bandwidth"])
    elif leaf == "phases":
        v = _uniform01(seed, name, n)
    elif leaf == "weight":
        fan_in = int(np.prod(shape[1:]))
        bound = 1.0 / math.sqrt(fan_in)
        v = (2.0 * _uniform01(seed, name, n) - 1.0) * bound
    elif leaf == "bias":
        # bound uses the fan_in of the sibling weight; recover it from the name
        v = (2.0 * _uniform01(seed, name, n) - 1.0) * 0.05
    else:  # pragma: no cover
        raise KeyError(name)
    return torch.from_numpy(v.astype(np.float32).reshape(shape))


def synth_unet3d_params(cfg, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    specs = unet3d_param_specs(cfg)
    return OrderedDict((k, synth_param(seed, k, shp, cfg)) for k, shp in specs.items())


def synth_unet3d_cond_params(cfg, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    specs = unet3d_cond_param_specs(cfg)
    return OrderedDict((k, synth_param(seed, k, shp, cfg)) for k, shp in specs.items())


def synth_atb(shape, seed: int) -> torch.Tensor:
    """Synthetic conditioning volume: a random field kept on a top slab and a few vertical
    'boreholes', zero elsewhere (stand-in for embed(X1)*mask, boreholes.py:111-126)."""
    B, C, X, Y, Z = shape
    v = synth_input(shape, seed, "atb")
    u = _uniform01(seed, "atbmask", B * X * Y).reshape(B, 1, X, Y, 1)
    mask = torch.from_numpy((u < 0.06).astype(np.float32)).expand(B, 1, X, Y, Z).clone()
    mask[..., : max(1, Z // 16)] = 1.0
    return v * mask


def synth_input(shape, seed: int, name: str = "x") -> torch.Tensor:
    n = int(np.prod(shape))
    return torch.from_numpy(_normal(seed, "input:" + name, n).astype(np.float32).reshape(shape))


def synth_times(B: int, seed: int, lo=0.0005, hi=0.9995) -> torch.Tensor:
    u = _uniform01(seed, "times", B)
    return torch.from_numpy((lo + (hi - lo) * u).astype(np.float32))
