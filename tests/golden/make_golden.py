"""Generate the golden fixtures under tests/golden/ by running the REAL reference code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference ships no golden vectors, so these outputs of the reference itself
— on deterministic synthetic weights/inputs from oracle/synth.py — are what pins the oracle.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader, synth, task  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def toy_model(x, t):
    """Cheap analytic stand-in for model(XT, T) used to pin solver semantics."""
    tt = t.view(-1, *([1] * (x.dim() - 1)))
    return torch.sin(3.0 * x) * (1.0 + tt) - 0.5 * x * tt


def gen_unet():
    cfg = synth.make_cfg()
    p = synth.synth_unet3d_params(cfg, 0)
    m = ref_loader.build_reference_unet(cfg, p)
    cases = {
        "b1_32": ((1, 18, 32, 32, 32), 1, torch.tensor([0.3])),
        "b2_16": ((2, 18, 16, 16, 16), 2, torch.tensor([0.05, 0.9])),
    }
    out = {}
    for name, (shape, seed, t) in cases.items():
        x = synth.synth_input(shape, seed)
        with torch.no_grad():
            y = m(x, t)
        out[f"{name}.t"] = t.numpy()
        out[f"{name}.seed"] = np.int64(seed)
        out[f"{name}.shape"] = np.array(shape, np.int64)
        out[f"{name}.y"] = y.numpy()
    np.savez_compressed(os.path.join(OUT, "unet3d_full_seed0.npz"), **out)
    # a second, small architecture (2 stages, narrower) for breadth: dim 32, mults (1,2)
    cfg2 = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=18, time_resolution=64,
                          time_bandwidth=100.0, attn_heads=2, attn_dim_head=16)
    p2 = synth.synth_unet3d_params(cfg2, 3)
    m2 = ref_loader.build_reference_unet(cfg2, p2)
    x = synth.synth_input((2, 18, 16, 16, 16), 4)
    t = torch.tensor([0.2, 0.7])
    with torch.no_grad():
        y = m2(x, t)
    np.savez_compressed(os.path.join(OUT, "unet3d_small_seed3.npz"), y=y.numpy(), t=t.numpy())


TRAIN_CFGS = {
    # name -> (cfg overrides, param seed, input shape)
    "small": (dict(dim=32, dim_mults=(1, 2), data_channels=18, time_resolution=64, time_bandwidth=100.0,
                   attn_heads=2, attn_dim_head=16), 3, (2, 18, 16, 16, 16)),
    "full": (dict(), 0, (1, 18, 16, 16, 16)),
}
# parameters whose complete gradient is stored (the others: L2 norm + 16 sampled entries)
TRAIN_FULL_GRADS = ("init_conv.bias", "time_mlp.0.freqs", "time_mlp.3.bias", "downs.0.2.mem_kv", "mid_attn.mem_kv",
                    "downs.0.0.block1.norm.g", "downs.0.2.norm.g", "final_conv.weight", "final_conv.bias")


def gen_train():
    """Gradients of the training-step loss (model_train_inference.py:443, mse(V, Vhat)/mse(V, 0)) through the
    REFERENCE Unet3D (autograd), dropout 0, for fixed XT / T / VT."""
    for name, (over, pseed, shape) in TRAIN_CFGS.items():
        cfg = synth.make_cfg(**over)
        p = synth.synth_unet3d_params(cfg, pseed)
        m = ref_loader.build_reference_unet(cfg, p)   # eval(): dropout is the identity (p = 0 parity runs)
        xt = synth.synth_input(shape, 11, "xt")
        vt = synth.synth_input(shape, 12, "vt")
        t = synth.synth_times(shape[0], 13)
        vhat = m(xt, t)
        loss = task.flow_loss(vt, vhat)
        loss.backward()
        out = {"loss": np.float64(loss.item()), "t": t.numpy(), "vhat": vhat.detach().numpy()}
        for k, prm in m.named_parameters():
            g = prm.grad.detach().reshape(-1)
            out[f"norm/{k}"] = np.float64(g.double().norm().item())
            idx = torch.linspace(0, g.numel() - 1, min(16, g.numel())).long()
            out[f"sample/{k}"] = g[idx].numpy()
            if k in TRAIN_FULL_GRADS:
                out[f"full/{k}"] = prm.grad.detach().numpy()
        np.savez_compressed(os.path.join(OUT, f"train_{name}.npz"), **out)


TRAIN_COND_CFGS = {
    "small": (dict(dim=32, dim_mults=(1, 2), data_channels=15, time_resolution=64, time_bandwidth=100.0,
                   attn_heads=2, attn_dim_head=16), 8, (2, 15, 16, 16, 16)),
    "full": (dict(data_channels=15), 5, (1, 15, 16, 16, 16)),
}
TRAIN_COND_FULL_GRADS = ("init_conv_ATb.weight", "init_conv_ATb.bias", "init_conv_x.bias", "downs.0.0.conv1.bias",
                         "downs.0.0.conv2.bias", "downs.0.1.time_mlp.1.weight", "downs.0.1.time_mlp.1.bias",
                         "downs.0.1.norm.g", "downs.1.1.conv2.bias", "ups.0.0.conv1.bias", "ups.1.1.conv1.bias",
                         "final_conv.weight")


def gen_train_cond():
    """Gradients of the flow loss through the REFERENCE Unet3DCond v3 (autograd; call site
    model_train_sh_inference_cond.py:431), dropout 0, for fixed XT / ATb / T / VT."""
    for name, (over, pseed, shape) in TRAIN_COND_CFGS.items():
        cfg = synth.make_cfg(**over)
        p = synth.synth_unet3d_cond_params(cfg, pseed)
        m = ref_loader.build_reference_unet_cond(cfg, p)
        xt = synth.synth_input(shape, 11, "xt")
        vt = synth.synth_input(shape, 12, "vt")
        atb = synth.synth_atb(shape, 14)
        t = synth.synth_times(shape[0], 13)
        vhat = m(xt, atb, t)
        loss = task.flow_loss(vt, vhat)
        loss.backward()
        out = {"loss": np.float64(loss.item()), "t": t.numpy(), "vhat": vhat.detach().numpy()}
        for k, prm in m.named_parameters():
            g = prm.grad.detach().reshape(-1)
            out[f"norm/{k}"] = np.float64(g.double().norm().item())
            idx = torch.linspace(0, g.numel() - 1, min(16, g.numel())).long()
            out[f"sample/{k}"] = g[idx].numpy()
            if k in TRAIN_COND_FULL_GRADS:
                out[f"full/{k}"] = prm.grad.detach().numpy()
        np.savez_compressed(os.path.join(OUT, f"train_cond_{name}.npz"), **out)


def synth_categories(shape, seed):
    """Layered synthetic category volume [B,1,X,Y,Z] in -1..13 with air (-1) on top of an uneven surface."""
    g = torch.Generator().manual_seed(seed)
    B, _, X, Y, Z = shape
    cats = torch.randint(0, 14, shape, generator=g)
    height = torch.randint(Z // 2, Z + 1, (B, 1, X, Y, 1), generator=g)      # first air index per column (Z = no air)
    z = torch.arange(Z).view(1, 1, 1, 1, Z)
    return torch.where(z >= height, torch.full_like(cats, -1), cats)


def gen_cond_frontend():
    """Masks of the REFERENCE boreholes.py (make_surface_mask, make_boreholes_mask, make_combined_mask) on synthetic
    category volumes.  The borehole draw uses the global CPU generator: the fixture stores the seed, and the oracle
    test replays the same torch.randint / torch.rand(1) call sequence to recover the coordinates."""
    bm = ref_loader.boreholes_module()
    out = {}
    for name, shape, seed in (("a", (2, 1, 16, 16, 8), 40), ("b", (3, 1, 12, 20, 10), 41)):
        cats = synth_categories(shape, seed)
        out[f"{name}.cats"] = cats.to(torch.int8).numpy()
        out[f"{name}.surface"] = np.packbits(bm.make_surface_mask(cats).numpy())
        torch.manual_seed(seed + 100)
        out[f"{name}.boreholes"] = np.packbits(bm.make_boreholes_mask(cats).numpy())
        torch.manual_seed(seed + 100)
        out[f"{name}.combined"] = np.packbits(bm.make_combined_mask(cats).numpy())
        out[f"{name}.seed"] = np.int64(seed + 100)
    np.savez_compressed(os.path.join(OUT, "cond_frontend.npz"), **out)


def gen_unet_cond():
    """Unet3DCond v3 (the conditional project's model: 15-d embedding, mults 1,2,2,3,4)."""
    cfg = synth.make_cfg(data_channels=15)
    p = synth.synth_unet3d_cond_params(cfg, 5)
    m = ref_loader.build_reference_unet_cond(cfg, p)
    out = {}
    shape = (1, 15, 16, 16, 16)
    x, atb, t = synth.synth_input(shape, 6), synth.synth_atb(shape, 7), torch.tensor([0.4])
    with torch.no_grad():
        y = m(x, atb, t)
    out["full_b1_16.y"], out["full_b1_16.t"] = y.numpy(), t.numpy()
    cfg2 = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=15, time_resolution=64,
                          time_bandwidth=100.0, attn_heads=2, attn_dim_head=16)
    p2 = synth.synth_unet3d_cond_params(cfg2, 8)
    m2 = ref_loader.build_reference_unet_cond(cfg2, p2)
    shape = (2, 15, 16, 16, 16)
    x, atb, t = synth.synth_input(shape, 9), synth.synth_atb(shape, 10), torch.tensor([0.15, 0.8])
    with torch.no_grad():
        y = m2(x, atb, t)
    out["small_b2_16.y"], out["small_b2_16.t"] = y.numpy(), t.numpy()
    np.savez_compressed(os.path.join(OUT, "unet3d_cond.npz"), **out)


def gen_interp():
    im = ref_loader.interpolation_module()
    t = torch.linspace(0.01, 0.99, 50)
    out = {"t": t.numpy()}
    makers = {
        "linear_two": lambda: im.LinearInterpolant(one_sided=False),
        "linear_one": lambda: im.LinearInterpolant(one_sided=True),
        "trig_two": lambda: im.TrigInterpolant(one_sided=False),
        "trig_one": lambda: im.TrigInterpolant(one_sided=True),
        "encdec": lambda: im.EncDecInterpolant(),
        "sbdm": lambda: im.SBDMInterpolant(),
        "mirror": lambda: im.MirrorInterpolant(),
    }
    X0 = synth.synth_input((3, 4, 5, 6, 7), 10, "x0")
    X1 = synth.synth_input((3, 4, 5, 6, 7), 11, "x1")
    Z = synth.synth_input((3, 4, 5, 6, 7), 12, "z")
    T = torch.tensor([0.1, 0.45, 0.8])
    out["X0"], out["X1"], out["Z"], out["T"] = X0.numpy(), X1.numpy(), Z.numpy(), T.numpy()
    for name, mk in makers.items():
        ip = mk()
        tab = torch.stack([ip.alpha(t), ip.beta(t), ip.gamma(t), ip.alpha_dot(t),
                           ip.beta_dot(t), ip.gamma_dot(t)])
        out[f"{name}.coeffs"] = tab.numpy()
        si = im.StochasticInterpolator(ip)
        z = None if ip.one_sided else Z
        XT, BT = si.flow_objective(T, X0, X1, z)
        out[f"{name}.XT"], out[f"{name}.BT"] = XT.numpy(), BT.numpy()
        XTd, tgt = si.denoising_objective(T, X0, X1, z)
        out[f"{name}.denoise_target"] = tgt.numpy()
        out[f"{name}.ST"] = si.get_ST(T, Z).numpy()
    np.savez_compressed(os.path.join(OUT, "interpolants.npz"), **out)


def gen_solvers():
    sm = ref_loader.solvers_module()
    im = ref_loader.interpolation_module()
    x0 = synth.synth_input((2, 3, 4, 4, 4), 20, "ode")
    out = {"x0": x0.numpy()}
    # ODEFlowSolver.solve with the fixed-grid stub odeint (dopri5 -> euler)
    s = sm.ODEFlowSolver(model=toy_model)
    out["flow_euler_t0.001_tf1_n11"] = s.solve(x0, t0=0.001, tf=1.0, n_steps=11).numpy()
    mask = torch.zeros(4, dtype=torch.bool)
    mask[1] = True
    out["flow_euler_masked"] = s.solve(x0, frozen_mask=mask, t0=0.001, tf=1.0, n_steps=6).numpy()
    out["mask"] = mask.numpy()
    # odeSol_RK4 (pure reference code, no stub involved)
    out["rk4_n10"] = sm.odeSol_RK4(x0, toy_model, nsteps=10, Tf=1.0).numpy()
    # one-sided denoising ODE, eq 6.7
    ip = im.LinearInterpolant(one_sided=True)
    d = sm.ODEOneSidedDenoisingSolver(toy_model, ip)
    out["denoise_ode_n9"] = d.solve(x0, t0=0.05, tf=0.95, n_steps=9).numpy()
    # SDE variant: epsilon must be a tensor (scalar eps raises TypeError at :212-214)
    sde = sm.SDEOneSidedDenoisingSolver(toy_model, ip, epsilon=torch.tensor(0.1))
    torch.manual_seed(1234)
    out["denoise_sde_n7_seed1234"] = sde.solve(x0, t0=0.05, tf=0.95, n_steps=7).numpy()
    np.savez_compressed(os.path.join(OUT, "solvers.npz"), **out)


def gen_decode():
    W = task.simplex_embedding(15, 18)
    x = synth.synth_input((2, 18, 8, 8, 16), 30, "dec")
    # add exact embedding rows (perfect matches) and scaled/near-tie mixtures
    cats = torch.arange(2 * 8 * 8 * 16).reshape(2, 8, 8, 16) % 15
    emb = task.embed(W, (cats - 1).unsqueeze(1))
    x2 = emb + 0.3 * synth.synth_input(tuple(emb.shape), 31, "dec2")
    tie = 0.5 * (W[3] + W[7])  # exact two-way tie direction -> first index wins
    x3 = tie.view(1, 18, 1, 1, 1).expand(1, 18, 2, 2, 2).contiguous()
    out = {"W": W.numpy()}
    # reference decode op sequence (model_train_inference.py:373-404) — decode_torch is
    # that sequence verbatim on torch CPU ops; logits are stored to pin op ORDER bit-for-bit
    for name, xx in (("rand", x), ("noisy_emb", x2), ("tie", x3)):
        out[f"{name}.x"] = xx.numpy()
        out[f"{name}.logits"] = task.decode_torch(W, xx, return_logits=True).numpy()
        out[f"{name}.pred"] = task.decode_torch(W, xx).numpy()
    out["cats"] = cats.numpy()
    # 15-d conditional embedding too (model_train_sh_inference_cond.py uses embedding_dim 15)
    W15 = task.simplex_embedding(15, 15)
    x15 = synth.synth_input((1, 15, 4, 4, 8), 32, "dec15")
    out["W15"], out["x15"] = W15.numpy(), x15.numpy()
    out["pred15"] = task.decode_torch(W15, x15).numpy()
    np.savez_compressed(os.path.join(OUT, "decode.npz"), **out)


if __name__ == "__main__":
    assert ref_loader.available(), "reference tree not found"
    torch.set_num_threads(os.cpu_count())
    if len(sys.argv) > 1 and sys.argv[1] == "train":   # only the training fixtures
        gen_train()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cond_frontend":
        gen_cond_frontend()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "train_cond":
        gen_train_cond()
        sys.exit(0)
    gen_interp()
    gen_solvers()
    gen_decode()
    gen_unet()
    gen_unet_cond()
    gen_train()
    gen_train_cond()
    gen_cond_frontend()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
