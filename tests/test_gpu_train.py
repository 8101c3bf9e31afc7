"""GPU parity tests of the training step (run with ``-m gpu`` on the B200 box): weight / data gradients of
the tcgen05 conv kernels against fp64 autograd, the parameter gradients of the whole Unet3D against the
oracle's autograd (pinned to the reference by tests/test_oracle_golden.py) and the committed reference
goldens, the fused clip + Adam + EMA kernels against their torch restatement.

Tolerances: single conv gradient with bf16-rounded operands <= 6e-3 rel-L2 (same bar as the forward conv);
whole-network parameter gradients in bf16 <= 5e-2 rel-L2 per tensor and <= 3e-2 over all parameters
(the bf16 velocity field itself is allowed 2e-2, north_star); optimiser kernels <= 1e-6.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BAR_CONV = 6e-3
BAR_GRAD_TENSOR = 5e-2
BAR_GRAD_ALL = 3e-2


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


def _golden(name):
    return np.load(os.path.join(os.path.dirname(__file__), "golden", name))


bf = lambda t: t.to(torch.bfloat16).float()


# ------------------------------------------------------------------ conv weight gradient (wgrad.cu)
def wgrad_case(B, c1, c2, cout, k, dims, unfold=0, seed=0):
    import torch.nn.functional as F
    from flowtrain_stochastic_interpolation_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator("cpu").manual_seed(seed)
    X, Y, Z = dims
    x = bf(torch.randn(B, c1, X, Y, Z, generator=g)).to(dev)
    x2 = bf(torch.randn(B, c2, X, Y, Z, generator=g)).to(dev) if c2 else None
    dy = bf(torch.randn(B, cout, X, Y, Z, generator=g)).to(dev)
    cin = c1 + c2
    xin = x if x2 is None else torch.cat((x, x2), 1)
    w = torch.zeros(cout, cin, k, k, k, dtype=torch.float64, device=dev, requires_grad=True)
    y = F.conv3d(xin.double(), w, None, padding=k // 2)
    (ref,) = torch.autograd.grad(y, w, dy.double())
    dw = torch.full((cout, cin, k, k, k), float("nan"), device=dev)
    _lib.check(_lib.lib.ftb_test_conv_wgrad(_lib.ptr(x), c1, _lib.ptr(x2), c2, _lib.ptr(dy), cout, k, _lib.ptr(dw),
                                            B, X, Y, Z, unfold, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert not torch.isnan(dw).any()
    return rel(dw, ref)


WGRAD_CASES = [
    dict(B=1, c1=48, c2=0, cout=48, k=3, dims=(4, 16, 16)),
    dict(B=2, c1=48, c2=0, cout=48, k=3, dims=(16, 32, 32)),
    dict(B=1, c1=48, c2=48, cout=48, k=3, dims=(8, 32, 16)),
    dict(B=1, c1=96, c2=0, cout=96, k=3, dims=(8, 16, 16)),
    dict(B=1, c1=144, c2=96, cout=144, k=3, dims=(4, 4, 4)),
    dict(B=2, c1=192, c2=144, cout=192, k=3, dims=(4, 4, 4)),
    dict(B=1, c1=144, c2=0, cout=192, k=3, dims=(2, 2, 2)),
    dict(B=1, c1=32, c2=0, cout=64, k=3, dims=(8, 8, 8)),
    dict(B=2, c1=18, c2=0, cout=48, k=7, dims=(16, 16, 16), unfold=1),
    dict(B=1, c1=18, c2=0, cout=32, k=7, dims=(8, 8, 24), unfold=1),
    dict(B=1, c1=48, c2=0, cout=384, k=1, dims=(8, 16, 16)),
    dict(B=1, c1=48, c2=0, cout=18, k=1, dims=(8, 16, 16)),
    dict(B=1, c1=96, c2=48, cout=48, k=1, dims=(4, 8, 8)),
    dict(B=1, c1=128, c2=0, cout=128, k=1, dims=(8, 8, 8)),
    dict(B=1, c1=48, c2=0, cout=48, k=3, dims=(3, 5, 7)),        # ragged
    dict(B=3, c1=48, c2=0, cout=96, k=3, dims=(5, 20, 12)),
    dict(B=1, c1=48, c2=0, cout=48, k=5, dims=(8, 8, 8)),
    # conditional model: EmbedATb 5^3 convs from the 15-channel opened ATb and at the widest stage, cubic 7^3 stems
    dict(B=2, c1=15, c2=0, cout=48, k=5, dims=(8, 16, 16)),
    dict(B=1, c1=192, c2=0, cout=192, k=5, dims=(4, 4, 4)),
    dict(B=1, c1=15, c2=0, cout=15, k=7, dims=(8, 16, 16)),
    dict(B=1, c1=15, c2=0, cout=48, k=7, dims=(8, 8, 16)),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()).replace(" ", ""))
def test_conv_wgrad_vs_autograd(ftb, case):
    assert wgrad_case(**case) <= BAR_CONV


def test_conv_wgrad_unstacked_layout(ftb, monkeypatch):
    """FTB_WGRAD_NOSTACK=1: natural [cg][h][w] halo layout, one MMA per (kh, kw) tap."""
    monkeypatch.setenv("FTB_WGRAD_NOSTACK", "1")
    assert wgrad_case(B=1, c1=48, c2=0, cout=48, k=3, dims=(4, 16, 16)) <= BAR_CONV
    assert wgrad_case(B=1, c1=96, c2=0, cout=96, k=3, dims=(8, 16, 16)) <= BAR_CONV


# ------------------------------------------------------------------ conv data gradient (forward kernel, flipped W^T)
def dgrad_case(B, cin, cout, k, dims, acc=False, seed=0):
    import torch.nn.functional as F
    from flowtrain_stochastic_interpolation_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator("cpu").manual_seed(seed)
    X, Y, Z = dims
    dy = bf(torch.randn(B, cout, X, Y, Z, generator=g)).to(dev)
    w = bf(torch.randn(cout, cin, k, k, k, generator=g) / (cout * k ** 3) ** 0.5).to(dev)
    a = bf(torch.randn(B, cin, X, Y, Z, generator=g)).to(dev) if acc else None
    x = torch.zeros(B, cin, X, Y, Z, dtype=torch.float64, device=dev, requires_grad=True)
    y = F.conv3d(x, w.double(), None, padding=k // 2)
    (ref,) = torch.autograd.grad(y, x, dy.double())
    if acc:
        ref = ref + a.double()
    dx = torch.full((B, cin, X, Y, Z), float("nan"), device=dev)
    _lib.check(_lib.lib.ftb_test_conv_dgrad(_lib.ptr(dy), _lib.ptr(w), cout, cin, k, _lib.ptr(a), _lib.ptr(dx),
                                            B, X, Y, Z, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert not torch.isnan(dx).any()
    return rel(dx, ref)


@pytest.mark.parametrize("case", [
    dict(B=1, cin=48, cout=48, k=3, dims=(8, 16, 16)),
    dict(B=2, cin=96, cout=48, k=3, dims=(8, 16, 16), acc=True),
    dict(B=1, cin=48, cout=384, k=1, dims=(8, 16, 16)),
    dict(B=1, cin=48, cout=18, k=1, dims=(4, 8, 8), acc=True),
    dict(B=1, cin=144, cout=192, k=3, dims=(4, 4, 4)),
    dict(B=1, cin=15, cout=48, k=5, dims=(8, 16, 16), acc=True),   # EmbedATb conv1 -> opened ATb
    dict(B=1, cin=96, cout=96, k=5, dims=(8, 8, 8)),
], ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()).replace(" ", ""))
def test_conv_dgrad_vs_autograd(ftb, case):
    assert dgrad_case(**case) <= BAR_CONV


# ------------------------------------------------------------------ trilinear adjoint
@pytest.mark.parametrize("di,do,acc", [((8, 8, 8), (4, 4, 4), False), ((4, 4, 4), (8, 8, 8), True),
                                       ((16, 16, 16), (32, 32, 32), False), ((32, 32, 32), (16, 16, 16), True),
                                       ((6, 10, 4), (3, 5, 2), False), ((3, 5, 7), (6, 10, 14), False),
                                       ((64, 64, 64), (4, 4, 4), False), ((16, 24, 8), (2, 3, 1), True)])
def test_trilinear_adjoint_vs_autograd(ftb, di, do, acc):
    """trilinear_resample_bwd == autograd of F.interpolate(mode="trilinear", align_corners=True) for 2x up / down,
    ragged sizes and the large down-scales of EmbedATb (64 -> 4); bf16 gradient storage."""
    import torch.nn.functional as F
    from flowtrain_stochastic_interpolation_b200 import _lib
    g = torch.Generator("cpu").manual_seed(1)
    B, C = 2, 24
    dout = bf(torch.randn(B, C, *do, generator=g)).cuda()
    a = bf(torch.randn(B, C, *di, generator=g)).cuda() if acc else None
    x = torch.zeros(B, C, *di, dtype=torch.float64, device="cuda", requires_grad=True)
    y = F.interpolate(x, size=do, mode="trilinear", align_corners=True)
    (ref,) = torch.autograd.grad(y, x, dout.double())
    if acc:
        ref = ref + a.double()
    din = torch.full((B, C) + tuple(di), float("nan"), device="cuda")
    _lib.check(_lib.lib.ftb_test_trilinear_bwd(_lib.ptr(dout), B, C, *di, *do, _lib.ptr(a), _lib.ptr(din),
                                               _lib.stream_ptr()))
    assert not torch.isnan(din).any()
    assert rel(din, ref) <= 4e-3   # bf16 output rounding


# ------------------------------------------------------------------ whole-network gradients
def _setup(ftb, dev, name):
    import importlib.util
    from oracle import synth
    gd = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(gd, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    over, pseed, shape = mg.TRAIN_CFGS[name]
    cfg = synth.make_cfg(**over)
    cfg["dropout"] = 0.0
    params = synth.synth_unet3d_params(cfg, pseed)
    net = ftb.Unet3D(**cfg).to(dev)
    net.load_state_dict(params)
    return cfg, params, net, shape, mg


def _unet_grads(ftb, dev, name, shape=None):
    from flowtrain_stochastic_interpolation_b200 import training
    from oracle import synth, task
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg, params, net, gshape, mg = _setup(ftb, dev, name)
    shape = shape or gshape
    xt = synth.synth_input(shape, 11, "xt").to(dev)
    vt = synth.synth_input(shape, 12, "vt").to(dev)
    t = synth.synth_times(shape[0], 13).to(dev)
    # oracle autograd on the same GPU, fp32, TF32 off
    loss_o, vhat_o, grads_o = task.training_grads({k: v.to(dev) for k, v in params.items()}, cfg, xt, t, vt)
    # product: train forward, loss gradient kernel, backward
    net.train()
    vhat = net(xt, t)
    loss = ftb.flow_loss(vt, vhat.detach())
    lo = torch.nn.functional.mse_loss(vt, vhat) / torch.nn.functional.mse_loss(vt, torch.zeros_like(vt))
    lo.backward()   # torch only differentiates the scalar loss; the network backward is ftb_unet3d_backward
    got = {k: p.grad.detach() for k, p in net.named_parameters()}
    return cfg, loss_o, vhat_o, grads_o, loss, vhat.detach(), got, mg


def _check_grads(grads_o, got, label):
    num = den = 0.0
    worst = ("", 0.0)
    total = sum(float(g.double().norm() ** 2) for g in grads_o.values()) ** 0.5
    for k, want in grads_o.items():
        e = rel(got[k], want)
        n = float(want.double().norm())
        num += float((got[k].double().cpu() - want.double().cpu()).norm() ** 2)
        den += n ** 2
        if n > 1e-3 * total / len(grads_o) ** 0.5 and e > worst[1]:
            worst = (k, e)
        if n > 1e-3 * total / len(grads_o) ** 0.5:   # tensors carrying a non-negligible share of the gradient
            assert e <= BAR_GRAD_TENSOR, f"{label}: grad {k}: rel-L2 {e:.3e} (norm {n:.3e})"
    allrel = (num / den) ** 0.5
    print(f"{label}: all-parameter grad rel-L2 {allrel:.3e}; worst tensor {worst[0]} {worst[1]:.3e}")
    assert allrel <= BAR_GRAD_ALL
    return allrel


def test_unet3d_small_arch_grads_vs_oracle(ftb, dev):
    cfg, loss_o, vhat_o, grads_o, loss, vhat, got, mg = _unet_grads(ftb, dev, "small")
    assert rel(vhat, vhat_o) <= 2e-2
    assert abs(loss.item() - loss_o.item()) <= 2e-2 * abs(loss_o.item())
    _check_grads(grads_o, got, "small arch 16^3")
    g = _golden("train_small.npz")   # and the REFERENCE's autograd (committed fixture)
    for k in mg.TRAIN_FULL_GRADS:
        assert rel(got[k], g[f"full/{k}"]) <= BAR_GRAD_TENSOR, k


def test_unet3d_full_arch_grads_vs_oracle_and_golden(ftb, dev):
    cfg, loss_o, vhat_o, grads_o, loss, vhat, got, mg = _unet_grads(ftb, dev, "full")
    assert rel(vhat, vhat_o) <= 2e-2
    _check_grads(grads_o, got, "full arch 16^3")
    g = _golden("train_full.npz")
    assert rel(vhat, g["vhat"]) <= 2e-2
    for k in mg.TRAIN_FULL_GRADS:
        assert rel(got[k], g[f"full/{k}"]) <= BAR_GRAD_TENSOR, k
    for k, gr in got.items():
        want = float(g[f"norm/{k}"])
        if want > 1e-6:
            assert abs(float(gr.double().norm()) - want) <= 5e-2 * want, k


def test_unet3d_full_arch_grads_32_batch2(ftb, dev):
    """Larger volume, batch 2 (multi-tile wgrad splits, linear attention at 32^3 / 16^3 / 8^3)."""
    cfg, loss_o, vhat_o, grads_o, loss, vhat, got, mg = _unet_grads(ftb, dev, "full", shape=(2, 18, 32, 32, 32))
    assert rel(vhat, vhat_o) <= 2e-2
    _check_grads(grads_o, got, "full arch 32^3 B=2")


# ------------------------------------------------------------------ conditional model (Unet3DCond v3) gradients
def _cond_grads(ftb, dev, name):
    import importlib.util
    from oracle import synth, task
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gd = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(gd, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    over, pseed, shape = mg.TRAIN_COND_CFGS[name]
    cfg = synth.make_cfg(**over)
    cfg["dropout"] = 0.0
    params = synth.synth_unet3d_cond_params(cfg, pseed)
    net = ftb.Unet3DCond(**cfg).to(dev)
    net.load_state_dict(params)
    xt = synth.synth_input(shape, 11, "xt").to(dev)
    vt = synth.synth_input(shape, 12, "vt").to(dev)
    atb = synth.synth_atb(shape, 14).to(dev)
    t = synth.synth_times(shape[0], 13).to(dev)
    loss_o, vhat_o, grads_o = task.cond_training_grads({k: v.to(dev) for k, v in params.items()}, cfg, xt, atb, t, vt)
    net.train()
    vhat = net(xt, atb, t)     # the reference call site: self.net(XT, ATb, T), model_train_sh_inference_cond.py:431
    lo = torch.nn.functional.mse_loss(vt, vhat) / torch.nn.functional.mse_loss(vt, torch.zeros_like(vt))
    lo.backward()
    got = {k: p.grad.detach() for k, p in net.named_parameters()}
    # the same module still samples (inference path, cached ATb branch) after a training step
    net.eval()
    with torch.no_grad():
        y = net(xt, atb, t)
    assert rel(y, vhat_o) <= 2e-2
    return loss_o, vhat_o, grads_o, lo.detach(), vhat.detach(), got, mg


@pytest.mark.parametrize("name", ["small", "full"])
def test_unet3d_cond_grads_vs_oracle_and_golden(ftb, dev, name):
    """Backward of the conditional model (EmbedATb 5^3 convs + trilinear of the opened ATb, MixATb FiLM of the concat,
    init_conv_ATb collecting from all ten embeddings) vs the oracle's autograd and the REFERENCE's (golden)."""
    loss_o, vhat_o, grads_o, loss, vhat, got, mg = _cond_grads(ftb, dev, name)
    assert rel(vhat, vhat_o) <= 2e-2
    assert abs(loss.item() - loss_o.item()) <= 2e-2 * abs(loss_o.item())
    _check_grads(grads_o, got, f"cond {name} arch 16^3")
    g = _golden(f"train_cond_{name}.npz")
    assert rel(vhat, g["vhat"]) <= 2e-2
    for k in mg.TRAIN_COND_FULL_GRADS:
        assert rel(got[k], g[f"full/{k}"]) <= BAR_GRAD_TENSOR, k
    for k, gr in got.items():
        want = float(g[f"norm/{k}"])
        if want > 1e-6:
            assert abs(float(gr.double().norm()) - want) <= 5e-2 * want, k


def test_cond_grads_ragged_volume_batch3(ftb, dev):
    """Conditional backward on a non-cubic volume (8 x 24 x 16, B = 3: EmbedATb trilinear to a ragged half scale, 5^3
    weight gradients over partial tiles) vs the oracle's autograd."""
    from oracle import synth, task
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=15, time_resolution=64, time_bandwidth=100.0,
                         attn_heads=2, attn_dim_head=16, dropout=0.0)
    params = synth.synth_unet3d_cond_params(cfg, 31)
    net = ftb.Unet3DCond(**cfg).to(dev)
    net.load_state_dict(params)
    shape = (3, 15, 8, 24, 16)
    xt, vt = synth.synth_input(shape, 41, "xt").to(dev), synth.synth_input(shape, 42, "vt").to(dev)
    atb = synth.synth_atb(shape, 44).to(dev)
    t = synth.synth_times(3, 43).to(dev)
    loss_o, vhat_o, grads_o = task.cond_training_grads({k: v.to(dev) for k, v in params.items()}, cfg, xt, atb, t, vt)
    net.train()
    vhat = net(xt, atb, t)
    lo = torch.nn.functional.mse_loss(vt, vhat) / torch.nn.functional.mse_loss(vt, torch.zeros_like(vt))
    lo.backward()
    assert rel(vhat.detach(), vhat_o) <= 2e-2
    _check_grads(grads_o, {k: p.grad.detach() for k, p in net.named_parameters()}, "cond small arch 8x24x16 B=3")


def test_train_forward_matches_inference_forward(ftb, dev):
    """The unfused train-mode forward and the fused inference forward are the same function."""
    cfg, params, net, shape, mg = _setup(ftb, dev, "full")
    from oracle import synth
    x = synth.synth_input(shape, 5).to(dev)
    t = synth.synth_times(shape[0], 6).to(dev)
    net.eval()
    with torch.no_grad():
        y_inf = net(x, t)
    net.train()
    y_tr = net(x, t).detach()
    net.eval()
    with torch.no_grad():
        y_inf2 = net(x, t)   # flat-bound parameters now; same result
    assert rel(y_tr, y_inf) <= 1e-2
    assert torch.equal(y_inf, y_inf2)


# ------------------------------------------------------------------ optimiser kernels + fused step
def test_adam_clip_ema_kernels_vs_torch(ftb, dev):
    from flowtrain_stochastic_interpolation_b200 import _lib
    from oracle import task
    g = torch.Generator("cpu").manual_seed(1)
    n = 1_000_003
    p = torch.randn(n, generator=g)
    m = torch.zeros(n)
    v = torch.zeros(n)
    pd, md, vd = p.to(dev), m.to(dev), v.to(dev)
    ss = torch.zeros(1, dtype=torch.float64, device=dev)
    for step in range(1, 4):
        gr = torch.randn(n, generator=g) * (0.01 if step == 2 else 3.0)   # step 2: norm < max_norm, no clipping
        gd = gr.to(dev)
        ss.zero_()
        _lib.check(_lib.lib.ftb_grad_sumsq(_lib.ptr(gd), n, _lib.ptr(ss), _lib.stream_ptr()))
        _lib.check(_lib.lib.ftb_adam_step(_lib.ptr(pd), _lib.ptr(gd), _lib.ptr(md), _lib.ptr(vd), n, 2e-4, 0.9, 0.999,
                                          1e-8, 0.0, 0, step, _lib.ptr(ss), 1.0, 1.0, _lib.stream_ptr()))
        tn = gr.double().norm().item()
        assert abs(ss.item() ** 0.5 - tn) <= 1e-9 * tn
        p, m, v = task.adam_reference(p, gr, m, v, step, total_norm=tn)
        assert torch.allclose(pd.cpu(), p, rtol=1e-5, atol=1e-7)
        assert torch.allclose(md.cpu(), m, rtol=1e-5, atol=1e-9)
    # AdamW, world-size scaling of a summed gradient
    p2 = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([p2], lr=1e-3, weight_decay=0.01)
    gr = torch.randn(n, generator=g)
    p2.grad = gr.clone()
    opt.step()
    pd2, md2, vd2 = p.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    gd = (gr * 4).to(dev)   # "sum over 4 ranks"
    _lib.check(_lib.lib.ftb_adam_step(_lib.ptr(pd2), _lib.ptr(gd), _lib.ptr(md2), _lib.ptr(vd2), n, 1e-3, 0.9, 0.999,
                                      1e-8, 0.01, 1, 1, None, 0.25, 0.0, _lib.stream_ptr()))
    assert torch.allclose(pd2.cpu(), p2.detach(), rtol=1e-5, atol=1e-7)


def test_flow_trainer_step_vs_oracle(ftb, dev):
    """FlowTrainer.step == training_step + clip + Adam + EMA of the reference, on fixed draws: the loss, the
    gradient the optimiser saw, and the parameter update given that gradient."""
    from oracle import synth, task
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=18, time_resolution=64, time_bandwidth=100.0,
                         attn_heads=2, attn_dim_head=16, dropout=0.0)
    params = synth.synth_unet3d_params(cfg, 3)
    kw = {k: v for k, v in cfg.items() if k != "data_channels"}
    mod = ftb.Geo3DStochInterp(data_shape=(16, 16, 16), embedding_dim=18, **kw).to(dev)
    mod.net.load_state_dict(params)
    tr = ftb.FlowTrainer(mod, lr=2e-4, max_grad_norm=1.0, ema_decay=0.9, ema_start_step=0)
    g = torch.Generator("cpu").manual_seed(5)
    batch = torch.randint(-1, 14, (2, 1, 16, 16, 16), generator=g).to(dev)
    n1 = synth.synth_input((2, 18, 16, 16, 16), 21, "n1").to(dev)
    x0 = synth.synth_input((2, 18, 16, 16, 16), 22, "x0").to(dev)
    T = synth.synth_times(2, 23).to(dev)
    p_before = tr.flat.clone()
    loss = tr.step(batch, noise1=n1, X0=x0, T=T)
    # oracle: same draws
    W = task.simplex_embedding(15, 18).to(dev)
    dparams = {k: v.to(dev) for k, v in params.items()}
    X1 = task.embed(W, batch) + 1e-3 * n1
    XT, VT = (1 - T.view(-1, 1, 1, 1, 1)) * x0 + T.view(-1, 1, 1, 1, 1) * X1, X1 - x0
    loss_o, _, grads_o = task.training_grads(dparams, cfg, XT, T, VT)
    assert abs(loss.item() - loss_o.item()) <= 2e-2 * abs(loss_o.item())
    gflat_o = torch.cat([grads_o[k].reshape(-1) for k in params.keys()])
    e = rel(tr.gflat, gflat_o)
    print(f"FlowTrainer gradient vs oracle rel-L2 {e:.3e}")
    assert e <= BAR_GRAD_ALL
    # the update, given the gradient the kernels produced
    tn = tr.gflat.double().norm().item()
    want_p, _, _ = task.adam_reference(p_before.cpu(), tr.gflat.cpu(), torch.zeros_like(p_before).cpu(),
                                       torch.zeros_like(p_before).cpu(), 1, total_norm=tn)
    assert torch.allclose(tr.flat.cpu(), want_p, rtol=1e-5, atol=1e-7)
    assert torch.equal(tr.ema_flat, tr.flat)           # first eligible step clones (callbacks.py:259-262)
    loss2 = tr.step(batch, noise1=n1, X0=x0, T=T)
    assert torch.isfinite(loss2)
    want_ema = 0.9 * want_p.to(dev) + 0.1 * tr.flat    # blend after the second step
    assert torch.allclose(tr.ema_flat, want_ema, rtol=1e-5, atol=1e-7)
    # the parameters the module exposes are the flat buffer (views), so sampling sees the new weights
    assert next(mod.net.parameters()).data_ptr() == tr.flat.data_ptr()


def test_cond_flow_trainer_step_vs_oracle(ftb, dev):
    """CondFlowTrainer.step == the conditional training_step (model_train_sh_inference_cond.py:401-467) + clip 0.3 +
    AdamW on fixed draws: loss (flow + T-weighted masked reconstruction), gradient seen by the optimiser, update."""
    from oracle import synth, task, unet3d_cond
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=15, time_resolution=64, time_bandwidth=100.0,
                         attn_heads=2, attn_dim_head=16, dropout=0.0)
    params = synth.synth_unet3d_cond_params(cfg, 8)
    kw = {k: v for k, v in cfg.items() if k != "data_channels"}
    mod = ftb.Geo3DStochInterpCond(data_shape=(16, 16, 16), embedding_dim=15, lambda_reconstruct=1.0, **kw).to(dev)
    mod.net.load_state_dict(params)
    tr = ftb.CondFlowTrainer(mod, lr=1e-3, max_grad_norm=0.3, ema_decay=0.9, ema_start_step=0)
    g = torch.Generator("cpu").manual_seed(5)
    shape = (2, 15, 16, 16, 16)
    batch = torch.randint(-1, 14, (2, 1, 16, 16, 16), generator=g)
    bores, nb = ftb.draw_boreholes(2, 16, 16, torch.Generator().manual_seed(9))
    n1 = synth.synth_input(shape, 21, "n1").to(dev)
    x0 = synth.synth_input(shape, 22, "x0").to(dev)
    T = synth.synth_times(2, 23, 1e-4, 0.9999).to(dev)
    p_before = tr.flat.clone()
    loss = tr.step(batch.to(dev), noise1=n1, X0=x0, T=T, bores=bores, n_bores=nb)
    # oracle: the reference's op sequence on the same draws, autograd through the functional Unet3DCond
    W = task.simplex_embedding(15, 15)
    mask = task.make_combined_mask(batch, bores, nb).expand(-1, 15, -1, -1, -1).to(dev)
    X1 = task.embed(W, batch).to(dev)
    ATb = X1 * mask
    X1 = X1 + 1e-4 * n1
    Tb = T.view(-1, 1, 1, 1, 1)
    XT, VT = (1 - Tb) * x0 + Tb * X1, X1 - x0
    p = {k: v.to(dev).clone().requires_grad_(True) for k, v in params.items()}
    vhat = unet3d_cond.unet3d_cond_forward(p, cfg, XT, ATb, T)
    X1_clean = task.embed(W, batch).to(dev)                          # b = X1[mask] is taken before the noise (:418)
    F = torch.nn.functional
    loss_o = task.cond_training_loss(VT, vhat, XT, X1_clean, T, mask, 1.0, X1_noisy=X1)   # oracle: :432-452
    rec = ((Tb.squeeze() * F.mse_loss(X1_clean[mask], XT[mask] + ((1 - Tb) * vhat)[mask]))
           / (F.mse_loss(X1, torch.zeros_like(X1)) + 1e-6)).mean()
    grads = torch.autograd.grad(loss_o, [p[k] for k in params.keys()], allow_unused=True)
    gflat_o = torch.cat([(gr if gr is not None else torch.zeros_like(p[k])).reshape(-1) for k, gr in zip(params.keys(), grads)])
    print(f"cond loss {loss.item():.6f} (oracle {loss_o.item():.6f}); flow {tr.last_terms[0].item():.6f} rec {tr.last_terms[1].item():.6f}")
    assert abs(loss.item() - loss_o.item()) <= 2e-2 * abs(loss_o.item())
    assert abs(tr.last_terms[1].item() - rec.item()) <= 3e-2 * abs(rec.item())
    e = rel(tr.gflat, gflat_o)
    print(f"CondFlowTrainer gradient vs oracle rel-L2 {e:.3e}")
    assert e <= BAR_GRAD_ALL
    # AdamW (decoupled weight decay 0.01) + clip 0.3, given the gradient the kernels produced
    gcl = tr.gflat.cpu() * min(1.0, 0.3 / (tr.gflat.double().norm().item() + 1e-6))
    pt = torch.nn.Parameter(p_before.cpu().clone())
    opt = torch.optim.AdamW([pt], lr=1e-3)
    pt.grad = gcl
    opt.step()
    assert torch.allclose(tr.flat.cpu(), pt.detach(), rtol=1e-5, atol=1e-7)
    loss2 = tr.step(batch.to(dev), noise1=n1, X0=x0, T=T, bores=bores, n_bores=nb)
    assert torch.isfinite(loss2)


def test_training_reduces_loss(ftb, dev):
    """A few fused steps on one fixed batch drive the loss down (end-to-end sanity of sign and scale)."""
    from oracle import synth
    cfg = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=18, time_resolution=64, time_bandwidth=100.0,
                         attn_heads=2, attn_dim_head=16, dropout=0.0)
    kw = {k: v for k, v in cfg.items() if k != "data_channels"}
    mod = ftb.Geo3DStochInterp(data_shape=(16, 16, 16), embedding_dim=18, **kw).to(dev)
    mod.net.load_state_dict(synth.synth_unet3d_params(cfg, 3))
    tr = ftb.FlowTrainer(mod, lr=1e-3, max_grad_norm=1.0, ema_decay=None)
    g = torch.Generator("cpu").manual_seed(9)
    batch = torch.randint(-1, 14, (2, 1, 16, 16, 16), generator=g).to(dev)
    n1 = synth.synth_input((2, 18, 16, 16, 16), 31, "n1").to(dev)
    x0 = synth.synth_input((2, 18, 16, 16, 16), 32, "x0").to(dev)
    T = synth.synth_times(2, 33).to(dev)
    losses = [tr.step(batch, noise1=n1, X0=x0, T=T).item() for _ in range(12)]
    print("losses", [f"{v:.4f}" for v in losses])
    assert losses[-1] < 0.9 * losses[0]


def test_flow_trainer_data_parallel_nccl(ftb):
    """N > 1: one rank per GPU, NCCL all-reduce bucketed from inside the backward (needs >= 2 GPUs; the
    single-GPU driver run skips it, `gpurun --gpus 2` runs it)."""
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(root, "tests", "_ddp_train_worker.py")],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DDP_OK" in r.stdout


def test_sharded_ensemble_votes_nccl(ftb):
    """N > 1: ensemble samples sharded over the ranks, one NCCL all-reduce of the vote histogram (needs >= 2 GPUs)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29641",
                        os.path.join(root, "tests", "_ensemble_worker.py")],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ENSEMBLE_OK" in r.stdout


def test_dropout_mask_is_consistent_between_forward_and_backward(ftb, dev):
    """dropout 0.1 (the reference's training default): the counter-based mask is regenerated by the backward.
    Pinned seed -> identical forwards; fresh seed -> different; and the directional derivative along the gradient,
    measured by central differences of the (mask-pinned) loss, matches |g| — it would not if forward and
    backward disagreed on which activations were dropped."""
    from flowtrain_stochastic_interpolation_b200 import training
    from oracle import synth
    cfg = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=18, time_resolution=64, time_bandwidth=100.0,
                         attn_heads=2, attn_dim_head=16, dropout=0.1)
    net = ftb.Unet3D(**cfg).to(dev)
    net.load_state_dict(synth.synth_unet3d_params(cfg, 3))
    net.train()
    shape = (2, 18, 16, 16, 16)
    xt, vt = synth.synth_input(shape, 11, "xt").to(dev), synth.synth_input(shape, 12, "vt").to(dev)
    t = synth.synth_times(2, 13).to(dev)

    def loss_of(out):
        return torch.nn.functional.mse_loss(vt, out) / torch.nn.functional.mse_loss(vt, torch.zeros_like(vt))

    net.set_dropout_seed(1234)
    y1 = net(xt, t)
    l1 = loss_of(y1)
    l1.backward()
    g = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
    with torch.no_grad():
        y1b = training.forward_train(net, xt, t)
        net.set_dropout_seed(99)
        y2 = training.forward_train(net, xt, t)
        net.eval()
        y0 = net(xt, t)
        net.train()
    assert torch.equal(y1.detach(), y1b)
    d12 = rel(y2, y1.detach())
    d10 = rel(y1.detach(), y0)
    print(f"dropout: other mask differs by {d12:.3f}, eval vs train {d10:.3f}")
    assert 1e-3 < d12 < 1.0 and 1e-3 < d10 < 1.0
    # central difference along the gradient direction with the mask pinned
    net.set_dropout_seed(1234)
    gn = g.double().norm().item()
    eps = 0.05
    flat = net._flat
    p0 = flat.clone()
    with torch.no_grad():
        vals = []
        for sgn in (+1, -1):
            flat.copy_(p0 + sgn * eps * g / gn)
            for p in net.parameters():
                p._version  # noqa: B018  (flat-bound views: mark dirty explicitly)
            from flowtrain_stochastic_interpolation_b200 import _lib
            _lib.check(_lib.lib.ftb_unet3d_mark_dirty(net._handle))
            vals.append(loss_of(training.forward_train(net, xt, t)).item())
        flat.copy_(p0)
    fd = (vals[0] - vals[1]) / (2 * eps)
    print(f"dropout: directional derivative {fd:.4f} vs |g| {gn:.4f}")
    assert abs(fd - gn) <= 0.15 * gn
