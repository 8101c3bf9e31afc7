"""GPU parity at the HEADLINE shapes (run with ``-m gpu``): what bench.py times is what is checked here.

* BASELINE configs[1]: one B=8 64^3 evaluation, and Heun / RK4 steps around the real network at B=8 64^3 (the planner
  picks different NZ / segment / ring plans at B=8 than at B=1, so the B=1 test does not cover it);
* configs[2]: the conditional model at 64^3, B=8, ONE conditioning volume shared by the batch, passed the way the
  reference passes it (``ATb.expand(n_samples, ...)``, model_inference_experiments.py:230-232);
* configs[4] / SURVEY a12: ``SDEOneSidedDenoisingSolver`` (solvers.py:180-222) and the denoise-ODE Heun leg against
  the reference-generated golden trajectory, the draws replayed through the ``noise=`` hook;
* the cache / sync hazards ADVICE r1 lists (recycled ATb address, ``.data.copy_`` weight swaps, a sampling forward
  between forward_train and backward, weight decay on frozen parameters).

The checker is the oracle run on the same GPU in true fp32 (TF32 off) or the committed golden fixtures.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BAR_BF16 = 2e-2     # velocity field, relative L2 (north_star)


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ftb():
    assert torch.cuda.is_available(), "-m gpu tests need a B200"
    import flowtrain_stochastic_interpolation_b200 as m
    return m


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _true_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.cuda.empty_cache()


def _golden(name):
    return np.load(os.path.join(os.path.dirname(__file__), "golden", name))


def _toy_model(x, t):
    tt = t.view(-1, *([1] * (x.dim() - 1)))
    return torch.sin(3.0 * x) * (1.0 + tt) - 0.5 * x * tt


# ------------------------------------------------------------------ SDE / denoise solvers (SURVEY a12, configs[4])
def test_sde_solver_vs_reference_golden(ftb, dev):
    """SDEOneSidedDenoisingSolver on the GPU vs the trajectory the REFERENCE class produced (golden: its adaptive_heun
    odeint replaced by fixed-grid Heun, two ``randn_like`` draws per step under seed 1234).  The draws are replayed in
    order through ``noise=`` so the comparison is deterministic."""
    g = _golden("solvers.npz")
    x0 = torch.from_numpy(g["x0"])
    torch.manual_seed(1234)
    draws = [torch.randn_like(x0) for _ in range(2 * 6)]       # 7 points = 6 Heun steps = 12 evaluations
    seen = []

    def noise(i):
        seen.append(i)
        return draws[i].to(dev)

    ip = ftb.LinearInterpolant(one_sided=True)
    sde = ftb.SDEOneSidedDenoisingSolver(_toy_model, ip, epsilon=torch.tensor(0.1), method="heun", noise=noise)
    got = sde.solve(x0.to(dev), t0=0.05, tf=0.95, n_steps=7)
    assert seen == list(range(12))
    assert got.shape == g["denoise_sde_n7_seed1234"].shape
    e = rel(got, g["denoise_sde_n7_seed1234"])
    print(f"SDE solver vs reference golden: rel-L2 {e:.3e}")
    assert e <= 1e-5
    want = torch.from_numpy(g["denoise_sde_n7_seed1234"])      # the trajectory grows to |x| ~ 80: relative per-element bar
    assert (got.cpu() - want).abs().max().item() <= 2e-6 * want.abs().max().item()
    # epsilon as a callable of t (the reference accepts both, :170-175), and eps = 0 == the denoising ODE
    sde0 = ftb.SDEOneSidedDenoisingSolver(_toy_model, ip, epsilon=lambda t: torch.tensor(0.0), method="heun",
                                          noise=lambda i: draws[i].to(dev))
    ode = ftb.ODEOneSidedDenoisingSolver(_toy_model, ip, method="heun")
    a = sde0.solve(x0.to(dev), t0=0.05, tf=0.95, n_steps=7)
    b = ode.solve(x0.to(dev), t0=0.05, tf=0.95, n_steps=7)
    assert rel(a, b) <= 2e-6      # fp32 vs float64 time grid (:187 vs :126)


def test_denoise_ode_heun_rk4_vs_oracle(ftb, dev):
    """Heun / RK4 legs of the eq-6.7 denoising ODE and the SDE drift at a second epsilon against the oracle loop."""
    from oracle import solvers as osol
    from oracle import synth
    x0 = synth.synth_input((2, 3, 4, 4, 4), 21, "ode")
    ip = ftb.LinearInterpolant(one_sided=True)
    for method in ("euler", "heun", "rk4"):
        got = ftb.ODEOneSidedDenoisingSolver(_toy_model, ip, method=method).solve(x0.to(dev), t0=0.05, tf=0.95, n_steps=9)
        want = osol.integrate(osol.make_denoise_func(_toy_model, "linear", True), x0, 0.05, 0.95, 9, method)
        assert rel(got, want) <= 1e-5, method
    g = torch.Generator().manual_seed(7)
    draws = [torch.randn(x0.shape, generator=g) for _ in range(4 * 4)]
    f = osol.make_denoise_func(_toy_model, "linear", True, epsilon=torch.tensor(0.03), noise=lambda i: draws[i])
    want = osol.integrate(f, x0, 0.1, 0.9, 5, "rk4")
    sde = ftb.SDEOneSidedDenoisingSolver(_toy_model, ip, epsilon=torch.tensor(0.03), method="rk4",
                                         noise=lambda i: draws[i].to(dev))
    assert rel(sde.solve(x0.to(dev), t0=0.1, tf=0.9, n_steps=5), want) <= 1e-5


def test_sde_around_the_network_vs_oracle(ftb, dev):
    """configs[4] in small: the SDE sampler around the real (full-architecture) network at 32^3 for 2 Heun steps vs the
    oracle loop around the oracle network with the same draws."""
    from oracle import solvers as osol
    from oracle import synth, unet3d
    cfg = synth.make_cfg()
    params = {k: v.to(dev) for k, v in synth.synth_unet3d_params(cfg, 0).items()}
    net = ftb.Unet3D(**cfg).to(dev)
    net.load_state_dict(synth.synth_unet3d_params(cfg, 0))
    x0 = synth.synth_input((1, 18, 32, 32, 32), 100).to(dev)
    g = torch.Generator().manual_seed(11)
    draws = [torch.randn(x0.shape, generator=g).to(dev) for _ in range(4)]
    ip = ftb.LinearInterpolant(one_sided=True)
    sde = ftb.SDEOneSidedDenoisingSolver(net, ip, epsilon=torch.tensor(0.1), method="heun", noise=lambda i: draws[i])
    got = sde.solve(x0, t0=0.3, tf=0.5, n_steps=3)
    with torch.no_grad():
        f = osol.make_denoise_func(lambda x, t: unet3d.unet3d_forward(params, cfg, x, t), "linear", True,
                                   epsilon=torch.tensor(0.1), noise=lambda i: draws[i])
        want = osol.integrate(f, x0, 0.3, 0.5, 3, "heun")
    e = rel(got[-1] - x0, want[-1] - x0)
    print(f"SDE around the network, displacement rel-L2 {e:.3e}")
    assert e <= BAR_BF16


# ------------------------------------------------------------------ configs[1]: B=8, 64^3
@pytest.fixture(scope="module")
def headline(ftb, dev):
    from oracle import synth
    cfg = synth.make_cfg()
    host = synth.synth_unet3d_params(cfg, 0)
    net = ftb.Unet3D(**cfg).to(dev)
    net.load_state_dict(host)
    return cfg, {k: v.to(dev) for k, v in host.items()}, net


def _oracle_eval(params, cfg, x, t, chunk=2):
    """The oracle at B=8 64^3 in true fp32, two samples at a time (samples are independent)."""
    from oracle import unet3d
    out = torch.empty_like(x)
    with torch.no_grad():
        for i in range(0, x.shape[0], chunk):
            out[i:i + chunk] = unet3d.unet3d_forward(params, cfg, x[i:i + chunk], t[i:i + chunk])
    return out


def test_headline_b8_64_evaluation_vs_oracle(ftb, dev, headline):
    """One evaluation at exactly the benchmark shape (B=8, 64^3, per-sample times) vs the oracle, per sample."""
    from oracle import synth
    cfg, params, net = headline
    x = synth.synth_input((8, 18, 64, 64, 64), 100).to(dev)
    t = torch.linspace(0.05, 0.95, 8, device=dev)
    with torch.no_grad():
        y = net(x, t)
    ref = _oracle_eval(params, cfg, x, t)
    per = [rel(y[i], ref[i]) for i in range(8)]
    print("B=8 64^3 per-sample rel-L2:", " ".join(f"{e:.2e}" for e in per))
    assert max(per) <= BAR_BF16
    # batch independence at the headline shape: sample 5 alone == sample 5 inside the batch (same kernels, other plan)
    with torch.no_grad():
        y5 = net(x[5:6].contiguous(), t[5:6].contiguous())
    assert rel(y5[0], ref[5]) <= BAR_BF16 and rel(y5[0], y[5]) <= 1e-2


def test_headline_heun_and_rk4_steps_around_the_network_64(ftb, dev, headline):
    """3 Heun steps and 1 RK4 step of the flow ODE around the REAL network at B=8 64^3 (bench.py's timed loop) vs the
    same integrators around the oracle network.  Tolerance on the displacement x_end - x_0 (the integral of the
    velocity): <= 2e-2 relative L2 per sample, the velocity-field bar; state max-abs <= 2e-3."""
    from oracle import solvers as osol
    from oracle import synth
    cfg, params, net = headline
    x0 = synth.synth_input((8, 18, 64, 64, 64), 100).to(dev)

    def omodel(x, t):
        return _oracle_eval(params, cfg, x, t)

    for method, n_pts, t0, tf in (("heun", 4, 0.2, 0.23), ("rk4", 2, 0.6, 0.61)):
        got = ftb.ODEFlowSolver(net, method=method).solve(x0, t0=t0, tf=tf, n_steps=n_pts, return_trajectory=False)
        with torch.no_grad():
            want = osol.integrate(osol.make_flow_func(omodel), x0, t0, tf, n_pts, method)[-1]
        per = [rel(got[i] - x0[i], want[i] - x0[i]) for i in range(8)]
        amax = (got - want).abs().max().item()
        print(f"{method}: displacement rel-L2 per sample max {max(per):.3e}, state max-abs {amax:.3e}")
        assert max(per) <= BAR_BF16 and amax <= 2e-3
        del got, want


# ------------------------------------------------------------------ configs[2]: conditional, 64^3, B=8, shared ATb
def test_cond_64_b8_shared_atb_vs_oracle(ftb, dev):
    from oracle import synth, unet3d_cond
    cfg = synth.make_cfg(data_channels=15)
    host = synth.synth_unet3d_cond_params(cfg, 5)
    params = {k: v.to(dev) for k, v in host.items()}
    net = ftb.Unet3DCond(**cfg).to(dev)
    net.load_state_dict(host)
    x = synth.synth_input((8, 15, 64, 64, 64), 11).to(dev)
    atb = synth.synth_atb((1, 15, 64, 64, 64), 12).to(dev)
    t = torch.linspace(0.1, 0.9, 8, device=dev)
    with torch.no_grad():
        y_exp = net(x, atb.expand(8, -1, -1, -1, -1), t)       # the reference's call shape (:230-232)
        n_cold = net.last_launches
        y_hot = net(x, atb.expand(8, -1, -1, -1, -1), t)       # a new view object of the same storage: cache hit
        assert net.last_launches < n_cold and torch.equal(y_exp, y_hot)
        y_one = net(x, atb, t)                                 # batch-1 ATb: same branch, same result
        assert rel(y_one, y_exp) <= 1e-6
        ref = torch.empty_like(x)
        for i in range(0, 8, 2):
            ref[i:i + 2] = unet3d_cond.unet3d_cond_forward(params, cfg, x[i:i + 2], atb.expand(2, -1, -1, -1, -1).contiguous(),
                                                           t[i:i + 2])
    per = [rel(y_exp[i], ref[i]) for i in range(8)]
    print("cond B=8 64^3 shared ATb per-sample rel-L2:", " ".join(f"{e:.2e}" for e in per))
    assert max(per) <= BAR_BF16


def test_cond_cache_survives_a_recycled_atb_address(ftb, dev):
    """``ATb.to(device).expand(n, ...)`` is a temporary; once it dies the caching allocator hands the
    same address to the NEXT conditioning volume (fresh tensor, version 0, same shape).  The cached ATb branch must
    not be served for it."""
    from oracle import synth, unet3d_cond
    cfg = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=15, time_resolution=64, time_bandwidth=100.0,
                         attn_heads=2, attn_dim_head=16)
    host = synth.synth_unet3d_cond_params(cfg, 8)
    params = {k: v.to(dev) for k, v in host.items()}
    net = ftb.Unet3DCond(**cfg).to(dev)
    net.load_state_dict(host)
    shape = (2, 15, 16, 16, 16)
    x = synth.synth_input(shape, 9).to(dev)
    t = torch.tensor([0.3, 0.7], device=dev)
    outs, ptrs = [], []
    for seed in (10, 20, 30):
        vol = synth.synth_atb((1,) + shape[1:], seed)                     # host volume, like the reference's loader
        with torch.no_grad():
            y = net(x, vol.to(dev).expand(2, -1, -1, -1, -1), t)          # temporary: freed after the call
            ref = unet3d_cond.unet3d_cond_forward(params, cfg, x, vol.to(dev).expand(2, -1, -1, -1, -1).contiguous(), t)
        assert rel(y, ref) <= BAR_BF16, seed
        outs.append(y)
        ptrs.append(net._atb_keepalive[0].data_ptr())
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    assert len(set(ptrs[:2])) == 2          # the keyed tensor is kept alive, so its address cannot be re-issued


def test_mark_dirty_after_a_data_copy_weight_swap(ftb, dev):
    """``param.data.copy_`` (reference EMACallback.apply_ema_weights) is invisible to the version counters:
    ``mark_dirty`` makes the engine pick the swapped weights up; ``EMAShadow.apply_to`` (in-place on the Parameter)
    is seen without it."""
    from oracle import synth, unet3d
    cfg = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=18, time_resolution=64, time_bandwidth=100.0,
                         attn_heads=2, attn_dim_head=16)
    a, b = synth.synth_unet3d_params(cfg, 3), synth.synth_unet3d_params(cfg, 4)
    net = ftb.Unet3D(**cfg).to(dev)
    net.load_state_dict(a)
    x = synth.synth_input((1, 18, 16, 16, 16), 1).to(dev)
    t = torch.tensor([0.4], device=dev)
    with torch.no_grad():
        ya = net(x, t)
        for n, p in net.named_parameters():
            p.data.copy_(b[n].to(dev))
        net.mark_dirty()
        yb = net(x, t)
        ref_b = unet3d.unet3d_forward({k: v.to(dev) for k, v in b.items()}, cfg, x, t)
    assert rel(yb, ref_b) <= BAR_BF16 and rel(ya, ref_b) > 0.1
    shadow = ftb.EMAShadow(decay=0.5, start_step=0)
    shadow.shadow = {n: a[n].to(dev).clone() for n, _ in net.named_parameters()}
    shadow.apply_to(net)
    with torch.no_grad():
        assert torch.equal(net(x, t), ya)


def test_sampling_forward_invalidates_the_training_tape(ftb, dev):
    """an eval forward between forward_train and backward frees the training workspace; the backward
    must refuse instead of differentiating clobbered activations."""
    from flowtrain_stochastic_interpolation_b200 import _lib
    from oracle import synth
    cfg = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=18, time_resolution=64, time_bandwidth=100.0,
                         attn_heads=2, attn_dim_head=16, dropout=0.0)
    net = ftb.Unet3D(**cfg).to(dev)
    net.load_state_dict(synth.synth_unet3d_params(cfg, 3))
    net.train()
    x = synth.synth_input((2, 18, 16, 16, 16), 11).to(dev)
    t = torch.tensor([0.2, 0.8], device=dev)
    out = net(x, t)
    with torch.no_grad():
        net(x, t)                                   # sampling forward on the same handle
    with pytest.raises(_lib.FtbError, match="no matching forward_train"):
        out.sum().backward()
    out = net(x, t)                                 # and the normal order still works
    out.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters() if p.requires_grad)


def test_adamw_leaves_frozen_parameters_alone(ftb, dev):
    """with ``time_learned_emb=False`` freqs / phases are frozen; torch's AdamW never touches a
    parameter without a gradient, so decoupled weight decay must not shrink them."""
    from oracle import synth
    small = dict(dim=32, dim_mults=(1, 2), time_resolution=64, time_bandwidth=100.0, attn_heads=2, attn_dim_head=16,
                 dropout=0.0, time_learned_emb=False)
    torch.manual_seed(0)
    mod = ftb.Geo3DStochInterpCond(embedding_dim=15, **small).to(dev)
    tr = ftb.CondFlowTrainer(mod, lr=1e-3, weight_decay=0.1)
    f0 = mod.net.time_mlp._modules["0"].freqs.detach().clone()
    w0 = mod.net.final_conv.weight.detach().clone()
    batch = torch.randint(-1, 14, (2, 1, 16, 16, 16), device=dev)
    for _ in range(2):
        loss = tr.step(batch)
    assert torch.isfinite(loss)
    assert torch.equal(mod.net.time_mlp._modules["0"].freqs, f0)
    assert not torch.equal(mod.net.final_conv.weight, w0)


# ------------------------------------------------------------------ Lightning-shaped training_step
def test_training_step_hook_matches_the_fused_trainer(ftb, dev):
    """``Geo3DStochInterp.training_step`` (autograd bridge + torch optimiser from ``configure_optimizers``, the way
    Lightning drives the reference, model_train_inference.py:417-473) against ``FlowTrainer.step`` on the same seed:
    same loss, same gradient, same Adam update."""
    from oracle import synth
    small = dict(dim=32, dim_mults=(1, 2), time_resolution=64, time_bandwidth=100.0, attn_heads=2, attn_dim_head=16,
                 dropout=0.0, time_learned_emb=True)
    sd = synth.synth_unet3d_params(synth.make_cfg(data_channels=18, **small), 3)
    batch = torch.randint(-1, 14, (2, 1, 16, 16, 16), device=dev)

    mod = ftb.Geo3DStochInterp(embedding_dim=18, learning_rate=2e-4, lr_decay=0.997, **small).to(dev)
    mod.net.load_state_dict(sd)
    mod.train()
    opt = mod.configure_optimizers()["optimizer"]
    torch.manual_seed(77)
    loss = mod.training_step(batch)
    assert loss.requires_grad and mod.logged["train_loss"].item() == loss.item()
    opt.zero_grad()
    loss.backward()
    g_hook = torch.cat([p.grad.reshape(-1) for p in mod.net.parameters()])
    opt.step()
    p_hook = torch.cat([p.detach().reshape(-1) for p in mod.net.parameters()])

    mod2 = ftb.Geo3DStochInterp(embedding_dim=18, **small).to(dev)
    mod2.net.load_state_dict(sd)
    tr = ftb.FlowTrainer(mod2, lr=2e-4, max_grad_norm=None, ema_decay=None)
    torch.manual_seed(77)
    loss2 = tr.step(batch)
    assert abs(loss.item() - loss2.item()) <= 1e-6 * abs(loss2.item())
    # same kernels on the same inputs; the bar is the run-to-run spread of the backward (fp32 atomic accumulation order
    # in the weight-gradient / norm-backward kernels, visible after bf16 rounding of the data gradients)
    assert rel(g_hook, tr.gflat) <= 1e-3
    assert rel(p_hook, tr.flat) <= 1e-5

    # conditional module: loss terms of the hook == the fused trainer's on the same draws
    cmod = ftb.Geo3DStochInterpCond(embedding_dim=15, **small).to(dev)
    cmod.train()
    ctr_mod = ftb.Geo3DStochInterpCond(embedding_dim=15, **small).to(dev)
    ctr_mod.net.load_state_dict(cmod.net.state_dict())
    ctr = ftb.CondFlowTrainer(ctr_mod, max_grad_norm=None, ema_decay=None, generator=torch.Generator().manual_seed(5))
    gen = torch.Generator().manual_seed(5)
    torch.manual_seed(78)
    closs, cflow, crec = cmod.cond_flow_loss(batch, generator=gen)
    closs.backward()
    torch.manual_seed(78)
    closs2 = ctr.step(batch)
    assert abs(closs.item() - closs2.item()) <= 1e-6 * abs(closs2.item())
    assert abs(cflow.item() - ctr.last_terms[0].item()) <= 1e-6 and abs(crec.item() - ctr.last_terms[1].item()) <= 1e-6
    g_c = torch.cat([p.grad.reshape(-1) if p.grad is not None else torch.zeros_like(p).reshape(-1)
                     for p in cmod.net.parameters()])
    assert rel(g_c, ctr.gflat) <= 1e-3
    cmod.on_after_backward()
    assert abs(cmod.logged["grad_norm"].item() - g_c.norm().item()) <= 1e-4 * g_c.norm().item()
