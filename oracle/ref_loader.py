"""Load the REAL reference modules by file path (TEST INFRASTRUCTURE, build container only).

``import flowtrain`` fails here (torchdiffeq / matplotlib / pyvista / lightning are not
installed — SURVEY §8c), so the model, interpolation and solver files are loaded directly
with importlib.  ``solvers.py`` imports ``torchdiffeq.odeint`` at module import; a stub
module is injected whose ``odeint`` is a fixed-grid stepper, which lets the reference's own
``ode_func`` closures (t.item(), frozen_mask, eq-6.7 drift) run unmodified.

Nothing under ``-m gpu`` tests, smoke() or bench.py may call this: /root/reference does not
exist on the GPU box.  ``available()`` says whether the tree is present.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("FLOWTRAIN_REFERENCE", "/root/reference")
_cache = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src/flowtrain/models/unet_attn_3d.py"))


def _load(name, rel):
    if name in _cache:
        return _cache[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def unet3d_module():
    return _load("ref_unet_attn_3d", "src/flowtrain/models/unet_attn_3d.py")


def unet3d_cond_module():
    return _load("ref_unet_attn_3d_cond_v3", "src/flowtrain/models/unet_attn_3d_cond_v3.py")


def boreholes_module():
    return _load("ref_boreholes", "project/geodata-3d-conditional/boreholes.py")


def interpolation_module():
    return _load("ref_interpolation", "src/flowtrain/interpolation/interpolation.py")


def _fixed_grid_odeint(func, y0, t, method="euler", **_):
    """Stand-in for torchdiffeq.odeint: fixed-grid stepping on the output grid ``t``.
    'dopri5' (the reference default) is mapped to euler — only ode_func semantics are
    being exercised, not the adaptive controller."""
    import torch

    step = {"dopri5": "euler", "adaptive_heun": "heun"}.get(method, method)
    ys = [y0]
    y = y0
    for k in range(len(t) - 1):
        t0, t1 = t[k], t[k + 1]
        h = t1 - t0
        if step == "euler":
            y = y + h * func(t0, y)
        elif step == "heun":
            k1 = func(t0, y)
            k2 = func(t1, y + h * k1)
            y = y + (h / 2) * (k1 + k2)
        else:
            raise ValueError(step)
        ys.append(y)
    return torch.stack(ys)


def solvers_module():
    if "ref_solvers" in _cache:
        return _cache["ref_solvers"]
    interp = interpolation_module()
    if "torchdiffeq" not in sys.modules:
        stub = types.ModuleType("torchdiffeq")
        stub.odeint = _fixed_grid_odeint
        stub.__graft_stub__ = True
        sys.modules["torchdiffeq"] = stub
    if "flowtrain" not in sys.modules:
        pkg = types.ModuleType("flowtrain")
        pkg.__path__ = []
        sys.modules["flowtrain"] = pkg
        sys.modules["flowtrain.interpolation"] = interp
    return _load("ref_solvers", "src/flowtrain/solvers/solvers.py")


def build_reference_unet(cfg, params):
    """Instantiate the reference Unet3D and load the given state dict (strict)."""
    m = unet3d_module().Unet3D(**cfg)
    m.load_state_dict(params, strict=True)
    return m.eval()


def build_reference_unet_cond(cfg, params):
    """Instantiate the reference Unet3DCond v3 and load the given state dict (strict)."""
    m = unet3d_cond_module().Unet3DCond(**cfg)
    m.load_state_dict(params, strict=True)
    return m.eval()
