"""CPU: the oracle restatement against golden vectors produced by the real reference
(tests/golden/make_golden.py), plus closed-form anchors."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import interp, ref_loader, solvers, synth, task, unet3d, unet3d_cond


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_param_specs_count_and_size():
    cfg = synth.make_cfg()
    specs = synth.unet3d_param_specs(cfg)
    assert len(specs) == 299  # 299 tensors
    assert sum(int(np.prod(s)) for s in specs.values()) == 25_193_410


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_param_specs_match_reference_state_dict():
    for cfg in (synth.make_cfg(), synth.make_cfg(dim=32, dim_mults=(1, 2), attn_heads=2,
                                                 attn_dim_head=16, time_resolution=64)):
        sd = ref_loader.unet3d_module().Unet3D(**cfg).state_dict()
        specs = synth.unet3d_param_specs(cfg)
        assert list(sd.keys()) == list(specs.keys())
        assert all(tuple(sd[k].shape) == tuple(specs[k]) for k in sd)


def test_unet_oracle_vs_reference_golden_16(golden_dir):
    g = _load(golden_dir, "unet3d_full_seed0.npz")
    cfg = synth.make_cfg()
    p = synth.synth_unet3d_params(cfg, 0)
    x = synth.synth_input(tuple(g["b2_16.shape"]), int(g["b2_16.seed"]))
    with torch.no_grad():
        y = unet3d.unet3d_forward(p, cfg, x, torch.from_numpy(g["b2_16.t"]))
    assert rel_l2(y, g["b2_16.y"]) < 1e-5  # same ops; differences = thread-count summation order


def test_unet_oracle_vs_reference_golden_32(golden_dir):
    g = _load(golden_dir, "unet3d_full_seed0.npz")
    cfg = synth.make_cfg()
    p = synth.synth_unet3d_params(cfg, 0)
    x = synth.synth_input(tuple(g["b1_32.shape"]), int(g["b1_32.seed"]))
    with torch.no_grad():
        y = unet3d.unet3d_forward(p, cfg, x, torch.from_numpy(g["b1_32.t"]))
    assert rel_l2(y, g["b1_32.y"]) < 1e-5


def test_unet_small_arch_golden(golden_dir):
    g = _load(golden_dir, "unet3d_small_seed3.npz")
    cfg = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=18, time_resolution=64,
                         time_bandwidth=100.0, attn_heads=2, attn_dim_head=16)
    p = synth.synth_unet3d_params(cfg, 3)
    x = synth.synth_input((2, 18, 16, 16, 16), 4)
    with torch.no_grad():
        y = unet3d.unet3d_forward(p, cfg, x, torch.from_numpy(g["t"]))
    assert rel_l2(y, g["y"]) < 1e-5


COND_SMALL = dict(dim=32, dim_mults=(1, 2), data_channels=15, time_resolution=64, time_bandwidth=100.0,
                  attn_heads=2, attn_dim_head=16)


def test_cond_param_specs_count_and_size():
    specs = synth.unet3d_cond_param_specs(synth.make_cfg(data_channels=15))
    assert len(specs) == 411
    assert sum(int(np.prod(s)) for s in specs.values()) == 53_049_349


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_cond_param_specs_match_reference_state_dict():
    for cfg in (synth.make_cfg(data_channels=15), synth.make_cfg(**COND_SMALL)):
        sd = ref_loader.unet3d_cond_module().Unet3DCond(**cfg).state_dict()
        specs = synth.unet3d_cond_param_specs(cfg)
        assert list(sd.keys()) == list(specs.keys())
        assert all(tuple(sd[k].shape) == tuple(specs[k]) for k in sd)


def test_unet_cond_oracle_vs_reference_golden(golden_dir):
    g = _load(golden_dir, "unet3d_cond.npz")
    cfg = synth.make_cfg(data_channels=15)
    p = synth.synth_unet3d_cond_params(cfg, 5)
    shape = (1, 15, 16, 16, 16)
    with torch.no_grad():
        y = unet3d_cond.unet3d_cond_forward(p, cfg, synth.synth_input(shape, 6), synth.synth_atb(shape, 7),
                                            torch.from_numpy(g["full_b1_16.t"]))
    assert rel_l2(y, g["full_b1_16.y"]) < 1e-5
    cfg2 = synth.make_cfg(**COND_SMALL)
    p2 = synth.synth_unet3d_cond_params(cfg2, 8)
    shape = (2, 15, 16, 16, 16)
    with torch.no_grad():
        y = unet3d_cond.unet3d_cond_forward(p2, cfg2, synth.synth_input(shape, 9), synth.synth_atb(shape, 10),
                                            torch.from_numpy(g["small_b2_16.t"]))
    assert rel_l2(y, g["small_b2_16.y"]) < 1e-5


KINDS = {
    "linear_two": ("linear", False), "linear_one": ("linear", True),
    "trig_two": ("trig", False), "trig_one": ("trig", True),
    "encdec": ("encdec", False), "sbdm": ("sbdm", True), "mirror": ("mirror", False),
}


@pytest.mark.parametrize("name", list(KINDS))
def test_interpolants_golden(golden_dir, name):
    g = _load(golden_dir, "interpolants.npz")
    kind, one = KINDS[name]
    t = torch.from_numpy(g["t"])
    tab = torch.stack(interp.coeffs(kind, t, one))
    np.testing.assert_array_equal(tab.numpy(), g[f"{name}.coeffs"])
    X0, X1, Z, T = (torch.from_numpy(g[k]) for k in ("X0", "X1", "Z", "T"))
    z = None if interp.is_one_sided(kind, one) else Z
    XT, BT = interp.flow_objective(kind, T, X0, X1, z, one_sided=one)
    np.testing.assert_array_equal(XT.numpy(), g[f"{name}.XT"])
    np.testing.assert_array_equal(BT.numpy(), g[f"{name}.BT"])
    _, tgt = interp.denoising_objective(kind, T, X0, X1, z, one_sided=one)
    np.testing.assert_array_equal(tgt.numpy(), g[f"{name}.denoise_target"])
    np.testing.assert_array_equal(interp.get_st(kind, T, Z, one).numpy(), g[f"{name}.ST"])


def test_interpolant_paper_anchors():
    # Fig. 5 of Albergo et al.: linear gamma = sqrt(2 t (1-t)), peak sqrt(1/2) at t = 1/2
    a, b, g, ad, bd, gd = interp.coeffs("linear", torch.tensor(0.5), False)
    assert abs(g.item() - 0.5 ** 0.5) < 1e-7 and abs(gd.item()) < 1e-7
    assert a.item() == 0.5 and b.item() == 0.5 and ad.item() == -1 and bd.item() == 1
    with pytest.raises(ValueError):
        interp.flow_objective("linear", torch.tensor([0.5]), torch.zeros(1, 2), torch.zeros(1, 2))


def _toy(x, t):
    tt = t.view(-1, *([1] * (x.dim() - 1)))
    return torch.sin(3.0 * x) * (1.0 + tt) - 0.5 * x * tt


def test_solvers_golden(golden_dir):
    g = _load(golden_dir, "solvers.npz")
    x0 = torch.from_numpy(g["x0"])
    tr = solvers.integrate(solvers.make_flow_func(_toy), x0, 0.001, 1.0, 11, "euler")
    np.testing.assert_allclose(tr.numpy(), g["flow_euler_t0.001_tf1_n11"], rtol=0, atol=1e-6)
    mask = torch.from_numpy(g["mask"])
    tr = solvers.integrate(solvers.make_flow_func(_toy, mask), x0, 0.001, 1.0, 6, "euler")
    np.testing.assert_allclose(tr.numpy(), g["flow_euler_masked"], rtol=0, atol=1e-6)
    assert torch.equal(tr[-1][..., mask], x0[..., mask])  # frozen voxels never move
    np.testing.assert_array_equal(solvers.ode_sol_rk4(x0, _toy, 10, 1.0).numpy(), g["rk4_n10"])
    f = solvers.make_denoise_func(_toy, "linear", True)
    tr = solvers.integrate(f, x0, 0.05, 0.95, 9, "euler")
    np.testing.assert_allclose(tr.numpy(), g["denoise_ode_n9"], rtol=2e-6, atol=2e-6)
    # SDE: replay the reference's randn_like draws (one per ode_func call, heun = 2 per step)
    torch.manual_seed(1234)
    f = solvers.make_denoise_func(_toy, "linear", True, epsilon=torch.tensor(0.1),
                                  noise=lambda i: torch.randn_like(x0))
    tr = solvers.integrate(f, x0, 0.05, 0.95, 7, "heun")
    np.testing.assert_allclose(tr.numpy(), g["denoise_sde_n7_seed1234"], rtol=5e-6, atol=5e-6)


def test_rk4_grid_quirk():
    # odeSol_RK4 does nsteps-1 updates and ends at t = Tf - h (solvers.py:233-243)
    seen = []

    def model(x, t):
        seen.append(float(t[0]))
        return torch.zeros_like(x)

    solvers.ode_sol_rk4(torch.zeros(1, 1), model, nsteps=5, Tf=1.0)
    assert len(seen) == 4 * 4 and abs(max(seen) - 0.8) < 1e-6


def test_decode_golden_bit_exact(golden_dir):
    g = _load(golden_dir, "decode.npz")
    W = torch.from_numpy(g["W"])
    en = task.normalized_embedding(W).numpy()
    for name in ("rand", "noisy_emb", "tie"):
        x = g[f"{name}.x"]
        logits = task.decode_numpy(en, x, return_logits=True)
        if name != "tie":  # torch picks another reduction strategy for 8-voxel tensors
            np.testing.assert_array_equal(logits, g[f"{name}.logits"])  # op order pinned bit-for-bit
        np.testing.assert_array_equal(task.decode_numpy(en, x), g[f"{name}.pred"])
    assert (g["tie.pred"] == 3).all()  # two-way tie -> first index
    en15 = task.normalized_embedding(torch.from_numpy(g["W15"])).numpy()
    np.testing.assert_array_equal(task.decode_numpy(en15, g["x15"]), g["pred15"])


def test_simplex_embedding_anchors():
    W = task.simplex_embedding(15, 18)
    gram = W @ W.T
    assert torch.allclose(gram.diagonal(), torch.ones(15), atol=1e-6)
    off = gram[~torch.eye(15, dtype=torch.bool)]
    assert torch.allclose(off, torch.full_like(off, -1 / 14), atol=1e-6)
    cats = torch.randint(-1, 14, (2, 1, 4, 4, 4))
    dec = task.decode_torch(W, task.embed(W, cats))
    assert torch.equal(dec, cats.squeeze(1) + 1)


def test_ema_and_loss():
    s, p = torch.ones(4), torch.zeros(4)
    assert torch.allclose(task.ema_update(s, p, 0.9995), torch.full((4,), 0.9995))
    v = torch.randn(2, 3, 4, 4, 4)
    assert task.flow_loss(v, torch.zeros_like(v)).item() == pytest.approx(1.0)
    assert task.flow_loss(v, v).item() == 0.0


# ------------------------------------------------------------------ training step (autograd of the oracle)
@pytest.mark.parametrize("name", ["small", "full"])
def test_training_grads_vs_reference_golden(golden_dir, name):
    """Gradients of the training loss through the oracle Unet3D == the reference module's autograd
    (tests/golden/make_golden.py gen_train): loss, output, every parameter's gradient norm and sampled
    entries, a handful of complete gradients."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    over, pseed, shape = mg.TRAIN_CFGS[name]
    g = _load(golden_dir, f"train_{name}.npz")
    cfg = synth.make_cfg(**over)
    params = synth.synth_unet3d_params(cfg, pseed)
    xt, vt = synth.synth_input(shape, 11, "xt"), synth.synth_input(shape, 12, "vt")
    t = torch.from_numpy(g["t"])
    loss, vhat, grads = task.training_grads(params, cfg, xt, t, vt)
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel_l2(vhat, g["vhat"]) <= 1e-5
    for k, gr in grads.items():
        want = float(g[f"norm/{k}"])
        got = gr.double().norm().item()
        assert abs(got - want) <= 2e-3 * want + 1e-9, (k, got, want)
        flat = gr.reshape(-1)
        idx = torch.linspace(0, flat.numel() - 1, min(16, flat.numel())).long()
        assert np.allclose(flat[idx].numpy(), g[f"sample/{k}"], rtol=5e-3, atol=2e-3 * want / max(1.0, flat.numel() ** 0.5))
    for k in mg.TRAIN_FULL_GRADS:
        assert rel_l2(grads[k], g[f"full/{k}"]) <= 1e-3, k


@pytest.mark.parametrize("name", ["small", "full"])
def test_cond_training_grads_vs_reference_golden(golden_dir, name):
    """Same for the conditional model: autograd through the oracle Unet3DCond v3 == the reference module's
    (tests/golden/make_golden.py gen_train_cond)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    over, pseed, shape = mg.TRAIN_COND_CFGS[name]
    g = _load(golden_dir, f"train_cond_{name}.npz")
    cfg = synth.make_cfg(**over)
    params = synth.synth_unet3d_cond_params(cfg, pseed)
    xt, vt = synth.synth_input(shape, 11, "xt"), synth.synth_input(shape, 12, "vt")
    atb = synth.synth_atb(shape, 14)
    t = torch.from_numpy(g["t"])
    loss, vhat, grads = task.cond_training_grads(params, cfg, xt, atb, t, vt)
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel_l2(vhat, g["vhat"]) <= 1e-5
    for k, gr in grads.items():
        want = float(g[f"norm/{k}"])
        got = gr.double().norm().item()
        assert abs(got - want) <= 2e-3 * want + 1e-9, (k, got, want)
        flat = gr.reshape(-1)
        idx = torch.linspace(0, flat.numel() - 1, min(16, flat.numel())).long()
        assert np.allclose(flat[idx].numpy(), g[f"sample/{k}"], rtol=5e-3, atol=2e-3 * want / max(1.0, flat.numel() ** 0.5))
    for k in mg.TRAIN_COND_FULL_GRADS:
        assert rel_l2(grads[k], g[f"full/{k}"]) <= 1e-3, k


@pytest.mark.parametrize("name", ["a", "b"])
def test_conditioning_masks_vs_reference_golden(golden_dir, name):
    """oracle make_surface_mask / make_boreholes_mask / make_combined_mask == the reference's boreholes.py
    (tests/golden/make_golden.py gen_cond_frontend), borehole draws replayed from the stored seed."""
    g = _load(golden_dir, "cond_frontend.npz")
    cats = torch.from_numpy(g[f"{name}.cats"]).long()
    B, _, X, Y, Z = cats.shape
    unpack = lambda k: torch.from_numpy(np.unpackbits(g[f"{name}.{k}"])[: cats.numel()].astype(bool)).view(cats.shape)
    assert torch.equal(task.make_surface_mask(cats), unpack("surface"))
    bores, nb = task.replay_reference_borehole_draws(int(g[f"{name}.seed"]), B, X, Y)
    assert torch.equal(task.make_boreholes_mask(cats, bores, nb), unpack("boreholes"))
    assert torch.equal(task.make_combined_mask(cats, bores, nb), unpack("combined"))


def test_ensemble_statistics_closed_form():
    """ensemble_statistics (model_inference_experiments.py:442-459) on a hand-made ensemble."""
    dec = torch.tensor([[-1, 0, 3, 3], [-1, 1, 3, 2], [-1, 0, 2, 2], [0, 0, 3, 13]]).view(4, 1, 1, 1, 4)
    pv, ent, most, em = task.ensemble_statistics(dec, 15)
    assert pv.shape == (1, 15, 1, 1, 4)
    assert torch.allclose(pv[0, :, 0, 0, 0], torch.tensor([0.75, 0.25] + [0.0] * 13))
    assert most.view(-1).tolist() == [-1, 0, 3, 2]          # ties -> first maximum
    assert abs(ent.view(-1)[1].item() - (-(0.75 * np.log(0.75) + 0.25 * np.log(0.25)))) < 1e-6
    assert em.view(-1)[0].item() == -1 and em.view(-1)[1].item() == ent.view(-1)[1].item()


def test_adaptive_rk_restatement_against_scipy_and_closed_form():
    """The restated torchdiffeq loop (oracle/solvers.odeint_adaptive; parity unpinned: torchdiffeq is not installed) on
    problems with known answers: linear decay (closed form), a nonlinear system against scipy's independent
    Dormand-Prince implementation, dense output on a non-uniform grid, and the adaptive_heun tableau."""
    from scipy.integrate import solve_ivp
    from oracle import solvers as osolv
    y0 = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    t = torch.tensor([0.0, 0.13, 0.5, 0.51, 1.0], dtype=torch.float64)
    st = {}
    y = osolv.odeint_adaptive(lambda tt, yy: -2.0 * yy, y0, t, "dopri5", rtol=1e-8, atol=1e-10, stats=st)
    want = y0[None] * torch.exp(-2.0 * t)[:, None]
    assert (y - want).abs().max().item() < 1e-7
    assert st["accepted"] >= 3

    def f(tt, yy):   # a stiff-ish nonlinear oscillator
        return torch.stack((yy[1], -yy[0] - 0.3 * yy[1] * (yy[0] ** 2 - 1.0), torch.sin(3.0 * tt + yy[2])))
    y = osolv.odeint_adaptive(f, y0, t, "dopri5", rtol=1e-8, atol=1e-10)
    ref = solve_ivp(lambda tt, yy: f(tt, torch.from_numpy(yy)).numpy(), (0.0, 1.0), y0.numpy(), method="DOP853",
                    t_eval=t.numpy(), rtol=1e-12, atol=1e-13).y.T
    assert np.abs(y.numpy() - ref).max() < 5e-7
    y2 = osolv.odeint_adaptive(f, y0, t, "adaptive_heun", rtol=1e-6, atol=1e-8)
    assert np.abs(y2.numpy() - ref).max() < 2e-4


def test_adam_reference_matches_torch():
    torch.manual_seed(0)
    p0 = torch.randn(1000)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=2e-4)
    m = torch.zeros(1000)
    v = torch.zeros(1000)
    q = p0.clone()
    for step in range(1, 4):
        gr = torch.randn(1000) * 3
        p.grad = gr.clone()
        tn = torch.nn.utils.clip_grad_norm_([p], 1.0)
        opt.step()
        q, m, v = task.adam_reference(q, gr, m, v, step, total_norm=tn)
        assert torch.allclose(q, p.detach(), rtol=1e-5, atol=1e-7)
