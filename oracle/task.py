"""Categorical embed / decode, training-step loss and EMA (TEST INFRASTRUCTURE).

Restates the parts of ``Geo3DStochInterp`` (a LightningModule, not importable here:
lightning / geogen are absent) that sit on the hot path:
  project/geodata-3d-unconditional/model_train_inference.py
    _initialize_embedding :330-356, embed :361-370, decode :373-404, training_step :417-457
  project/geodata-3d-conditional/callbacks.py  EMACallback.on_train_batch_end :238-268

``decode_numpy`` spells out the fp32 operation ORDER that torch's CPU kernels use for the
reference decode (verified bit-for-bit on logits in the build container,
tests/golden/make_golden.py): sequential sum of squares without FMA -> sqrt -> clamp 1e-12
-> divide; 15 sequential dot products without FMA; first-max argmax.  The CUDA decode
kernel follows the same order so the integer output is bit-exact.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import interp as _interp


def simplex_embedding(n_cats: int, n_dims: int) -> torch.Tensor:
    """_initialize_embedding :330-356 — centred simplex, unit rows."""
    m = torch.zeros(n_cats, n_dims)
    m[:, :n_cats] = torch.eye(n_cats)
    centroid = torch.ones(n_cats) / n_cats
    centroid = torch.cat([centroid, torch.zeros(n_dims - n_cats)])
    m[:, :n_cats] -= centroid[:n_cats].unsqueeze(0)
    return m / m.norm(dim=1, keepdim=True)


def embed(weight: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """embed :361-370 — [B,1,X,Y,Z] categories (-1..n-2) -> [B,E,X,Y,Z]."""
    idx = x.squeeze(1).long() + 1
    e = F.embedding(idx, weight)
    return e.permute(0, 4, 1, 2, 3).contiguous()


def decode_torch(weight: torch.Tensor, x: torch.Tensor, return_logits=False):
    """decode :373-404, same op sequence as the reference (incl. the broadcast temp)."""
    n_cat, E = weight.shape
    xn = F.normalize(x, dim=1)
    en = F.normalize(weight, dim=1)
    logits = (xn.unsqueeze(1) * en.view(1, n_cat, E, 1, 1, 1)).sum(dim=2)
    return logits if return_logits else torch.argmax(logits, dim=1)


def normalized_embedding(weight: torch.Tensor) -> torch.Tensor:
    """F.normalize(embedding.weight, dim=1) (:384) on CPU fp32 — 270 numbers, host side."""
    return F.normalize(weight.detach().float().cpu(), dim=1)


def decode_numpy(en: np.ndarray, x: np.ndarray, return_logits=False):
    """Explicit-order fp32 decode.  en: [n_cat,E] already normalised; x: [B,E,...]."""
    f32 = np.float32
    x = np.ascontiguousarray(x, dtype=f32)
    en = np.ascontiguousarray(en, dtype=f32)
    E = x.shape[1]
    ss = np.zeros_like(x[:, 0])
    for e in range(E):
        ss = (ss + (x[:, e] * x[:, e]).astype(f32)).astype(f32)
    nrm = np.maximum(np.sqrt(ss).astype(f32), f32(1e-12))
    xn = [(x[:, e] / nrm).astype(f32) for e in range(E)]
    n_cat = en.shape[0]
    logits = np.zeros((x.shape[0], n_cat) + x.shape[2:], f32)
    for c in range(n_cat):
        acc = np.zeros_like(ss)
        for e in range(E):
            acc = (acc + (xn[e] * en[c, e]).astype(f32)).astype(f32)
        logits[:, c] = acc
    if return_logits:
        return logits
    return np.argmax(logits, axis=1).astype(np.int64)  # first max, like torch.argmax


def flow_loss(VT: torch.Tensor, VT_hat: torch.Tensor) -> torch.Tensor:
    """training_step :443 — mse(VT, VT_hat) / mse(VT, 0)."""
    return F.mse_loss(VT, VT_hat) / F.mse_loss(VT, torch.zeros_like(VT))


def training_step_loss(net, weight, batch, noise1, X0, T, kind="linear", one_sided=True):
    """training_step :417-457 with all random draws passed in:
    X1 = embed(batch) + 1e-3*noise1 ; XT,VT = flow_objective(T,X0,X1) ; loss."""
    X1 = embed(weight, batch)
    X1 = X1 + 1e-3 * noise1
    XT, VT = _interp.flow_objective(kind, T, X0, X1, one_sided=one_sided)
    VT_hat = net(XT, T)
    return flow_loss(VT, VT_hat), (XT, VT, VT_hat)


def training_grads(params, cfg, XT, T, VT):
    """Autograd gradients of the training-step loss (:440-443) w.r.t. every parameter of the functional
    oracle Unet3D (oracle/unet3d.py), dropout 0.  Returns (loss, vhat, {name: grad})."""
    from . import unet3d
    p = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    vhat = unet3d.unet3d_forward(p, cfg, XT, T)
    loss = flow_loss(VT, vhat)
    names = list(p.keys())
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    return loss.detach(), vhat.detach(), {k: (g if g is not None else torch.zeros_like(p[k])) for k, g in zip(names, grads)}


def cond_training_grads(params, cfg, XT, ATb, T, VT):
    """Conditional counterpart of ``training_grads``: autograd gradients of the flow loss through the functional
    oracle Unet3DCond v3 (oracle/unet3d_cond.py; reference call site model_train_sh_inference_cond.py:431), dropout 0."""
    from . import unet3d_cond
    p = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    vhat = unet3d_cond.unet3d_cond_forward(p, cfg, XT, ATb, T)
    loss = flow_loss(VT, vhat)
    names = list(p.keys())
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    return loss.detach(), vhat.detach(), {k: (g if g is not None else torch.zeros_like(p[k])) for k, g in zip(names, grads)}


def cond_training_loss(VT, VT_hat, XT, X1_clean, T, mask, lambda_reconstruct, X1_noisy=None):
    """Loss of the conditional training_step (model_train_sh_inference_cond.py:432-452): flow loss with the 1e-6
    guard plus the T-weighted reconstruction of the observed voxels, b = X1[mask] taken BEFORE the 1e-4 noise is added
    (:418), b_hat = XT + (1 - T) VT_hat on the mask (:434-436), normalised by mse(X1, 0) of the NOISY X1 (:446)."""
    Tb = T.view(-1, 1, 1, 1, 1)
    X1n = X1_clean if X1_noisy is None else X1_noisy
    b = X1_clean[mask]
    b_hat = XT[mask] + ((1 - Tb) * VT_hat)[mask]
    mse = F.mse_loss(VT, VT_hat) / (F.mse_loss(VT, torch.zeros_like(VT)) + 1e-6)
    rec = (Tb.squeeze() * F.mse_loss(b, b_hat)) / (F.mse_loss(X1n, torch.zeros_like(X1n)) + 1e-6)
    return mse + lambda_reconstruct * rec.mean()


def adam_reference(p, g, m, v, step, lr=2e-4, b1=0.9, b2=0.999, eps=1e-8, max_norm=1.0, total_norm=None):
    """clip_grad_norm_(max_norm) + one torch.optim.Adam step (configure_optimizers :465-473, Lightning
    gradient_clip_val), restated on flat tensors.  Returns (p, m, v)."""
    if max_norm and total_norm is not None:
        g = g * min(1.0, max_norm / (float(total_norm) + 1e-6))
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    denom = v.sqrt() / (1 - b2 ** step) ** 0.5 + eps
    return p - (lr / (1 - b1 ** step)) * m / denom, m, v


def ema_update(shadow: torch.Tensor, param: torch.Tensor, decay: float) -> torch.Tensor:
    """EMACallback.on_train_batch_end :263-266 — shadow = a*shadow + (1-a)*param."""
    return decay * shadow + (1.0 - decay) * param


def vote_probabilities(decoded: torch.Tensor, n_cat: int) -> torch.Tensor:
    """inference_demo.ipynb cell 21 — one-hot vote over the ensemble axis (dim 0)."""
    oh = F.one_hot(decoded.long(), n_cat).float()
    return oh.mean(dim=0).movedim(-1, 0)


# ---------------------------------------------------------------------------------- conditioning masks
def make_boreholes_mask(X: torch.Tensor, bores: torch.Tensor, n_bores: torch.Tensor) -> torch.Tensor:
    """make_boreholes_mask — project/geodata-3d-conditional/boreholes.py:45-75 with the random borehole columns passed
    in (bores [B, max, 2] (x, y), n_bores [B]) so that parity is definable: the whole z column at each (x, y)."""
    B, C, sx, sy, sz = X.shape
    mask = torch.zeros((B, 1, sx, sy, sz), dtype=torch.bool)
    for b in range(B):
        pts = bores[b, : int(n_bores[b])].long()
        mask[b, 0, pts[:, 0], pts[:, 1], :] = True
    return mask


def make_surface_mask(X: torch.Tensor) -> torch.Tensor:
    """make_surface_mask — boreholes.py:77-111: top z slice; every air voxel (value -1) and the voxel at z-1 below it
    (clamped at 0)."""
    B, C, sx, sy, sz = X.shape
    mask = torch.zeros((B, 1, sx, sy, sz), dtype=torch.bool)
    mask[:, 0, :, :, sz - 1] = True
    for b in range(B):
        positions = (X[b, 0] == -1).nonzero(as_tuple=True)
        if positions[0].numel() > 0:
            xc, yc, zc = positions
            mask[b, 0, xc, yc, zc] = True
            mask[b, 0, xc, yc, torch.clamp(zc - 1, min=0)] = True
    return mask


def make_combined_mask(X: torch.Tensor, bores: torch.Tensor, n_bores: torch.Tensor) -> torch.Tensor:
    """make_combined_mask — boreholes.py:114-129."""
    return make_boreholes_mask(X, bores, n_bores) | make_surface_mask(X)


def jittered_grid_points(X: int, Y: int, n_bores: int, rand: torch.Tensor) -> torch.Tensor:
    """_jittered_grid_points — boreholes.py:9-42, with the 2 * n_x * n_y uniform draws passed in as ``rand`` [n_x*n_y, 2]
    (the reference calls torch.rand(1) twice per cell, x first, cells in (i, j) order)."""
    import math
    n_x = int(math.floor(math.sqrt(n_bores)))
    n_y = int(math.ceil(n_bores / n_x))
    cwx, cwy = X / n_x, Y / n_y
    points = []
    k = 0
    for i in range(n_x):
        for j in range(n_y):
            cx, cy = (i + 0.5) * cwx, (j + 0.5) * cwy
            rx = rand[k, 0:1] * cwx - cwx / 2
            ry = rand[k, 1:2] * cwy - cwy / 2
            k += 1
            px = torch.clamp(cx + rx, min=0, max=X - 1)
            py = torch.clamp(cy + ry, min=0, max=Y - 1)
            points.append((px.item(), py.item()))
    points = points[:n_bores]
    return torch.tensor(points, dtype=torch.long)


def replay_reference_borehole_draws(seed, B, X, Y):
    """The reference's call sequence on the global CPU generator (boreholes.py:66-68, :27-28): per sample one
    randint(8, 32), then rand(1) twice per grid cell."""
    import math
    torch.manual_seed(seed)
    bores = torch.zeros(B, 64, 2, dtype=torch.long)
    nb = torch.zeros(B, dtype=torch.long)
    for b in range(B):
        n = torch.randint(8, 32, (1,)).item()
        n_x = int(math.floor(math.sqrt(n)))
        n_y = int(math.ceil(n / n_x))
        rand = torch.stack([torch.cat((torch.rand(1), torch.rand(1))) for _ in range(n_x * n_y)])
        pts = jittered_grid_points(X, Y, n, rand)
        bores[b, : pts.shape[0]] = pts
        nb[b] = pts.shape[0]
    return bores, nb


# ---------------------------------------------------------------------------------- ensemble statistics
def ensemble_statistics(sols_decoded: torch.Tensor, num_categories: int = 15):
    """ensemble_analysis — project/geodata-3d-conditional/model_inference_experiments.py:442-459.  sols_decoded:
    [S, 1, X, Y, Z] categories in -1 .. num_categories-2.  Returns (probability_vector [1,C,X,Y,Z], entropy [X,Y,Z],
    most_probable [X,Y,Z] (air = -1), entropy_masked)."""
    oh = F.one_hot(sols_decoded.squeeze(1) + 1, num_categories).permute(0, 4, 1, 2, 3).float()
    pv = oh.mean(dim=0, keepdim=True)
    eps = 1e-8
    entropy = -torch.sum(pv * torch.log(pv + eps), dim=1).squeeze(0)
    most = torch.argmax(pv, dim=1).squeeze(0) - 1
    em = entropy.clone()
    em[most == -1] = -1
    return pv, entropy, most, em
