"""CPU: the C-ABI library loads and exports every symbol include/ftb.h declares, the host-side
mirror of the reference API (parameter tree, schedules, error behaviour) is right, and the
sample-sharding logic works under a world_size-2 gloo group.  No compute kernels run here."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import flowtrain_stochastic_interpolation_b200 as ftb
from flowtrain_stochastic_interpolation_b200 import _lib, sharding
from oracle import interp, ref_loader, synth, task

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    syms = _lib.header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(_lib.lib, s), f"libftb.so does not export {s}"
        assert s in _lib._SIGS, f"{s} has no ctypes signature"
    assert _lib.lib.ftb_version() >= 100


def test_no_cpu_fallback():
    net = ftb.Unet3D(**synth.make_cfg(dim=32, dim_mults=(1, 2), attn_heads=2, attn_dim_head=16, time_resolution=64))
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 18, 8, 8, 8), torch.zeros(1))
    with pytest.raises(RuntimeError, match="CUDA"):
        ftb.decode(task.simplex_embedding(15, 18), torch.zeros(1, 18, 2, 2, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        ftb.StochasticInterpolator(ftb.LinearInterpolant(True)).flow_objective(
            torch.rand(2), torch.zeros(2, 4, 2, 2, 2), torch.zeros(2, 4, 2, 2, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        ftb.ODEFlowSolver(lambda x, t: x).solve(torch.zeros(1, 4, 2, 2, 2))
    # the product package never imports the oracle
    for mod in list(sys.modules):
        if mod.startswith("flowtrain_stochastic_interpolation_b200"):
            src = getattr(sys.modules[mod], "__file__", None)
            if src and src.endswith(".py"):
                assert "oracle" not in open(src).read().replace("oracle/unet3d.py", ""), mod


@pytest.mark.parametrize("cfg", [synth.make_cfg(),
                                 synth.make_cfg(dim=32, dim_mults=(1, 2), attn_heads=2, attn_dim_head=16,
                                                time_resolution=64, time_learned_emb=False)])
def test_param_tree_matches_reference_keys(cfg):
    net = ftb.Unet3D(**cfg)
    specs = synth.unet3d_param_specs(cfg)
    sd = net.state_dict()
    assert list(sd.keys()) == list(specs.keys())
    assert all(tuple(sd[k].shape) == tuple(specs[k]) for k in sd)
    # learned vs random Fourier features: requires_grad of freqs/phases (unet_attn_3d.py:198-218)
    assert net.time_mlp._modules["0"].freqs.requires_grad == bool(cfg["time_learned_emb"])
    if ref_loader.available():
        ref = ref_loader.unet3d_module().Unet3D(**cfg)
        assert list(ref.state_dict().keys()) == list(sd.keys())
        assert [n for n, p in ref.named_parameters() if p.requires_grad] == \
               [n for n, p in net.named_parameters() if p.requires_grad]
        net.load_state_dict(ref.state_dict())  # strict


def test_ctor_rejects_unsupported_and_bad_dims():
    with pytest.raises(NotImplementedError):
        ftb.Unet3D(dim=32, self_condition=True)
    with pytest.raises(NotImplementedError):
        ftb.Unet3D(dim=32, time_sin_pos=True)
    with pytest.raises(_lib.FtbError):
        ftb.Unet3D(dim=20)  # not a multiple of 16


def test_init_statistics_like_reference():
    torch.manual_seed(0)
    net = ftb.Unet3D(**synth.make_cfg())
    sd = net.state_dict()
    w = sd["downs.0.0.block1.proj.weight"]
    bound = 1 / (48 * 27) ** 0.5
    assert w.abs().max() <= bound and w.abs().max() > 0.9 * bound
    assert torch.all(sd["downs.0.0.block1.norm.g"] == 1)
    assert 800 < sd["time_mlp.0.freqs"].std() < 1200
    assert 0 <= sd["time_mlp.0.phases"].min() and sd["time_mlp.0.phases"].max() <= 1


def test_schedules_match_oracle_and_reference():
    t = torch.linspace(0.01, 0.99, 50)
    cases = [(ftb.LinearInterpolant(False), "linear", False), (ftb.LinearInterpolant(True), "linear", True),
             (ftb.TrigInterpolant(False), "trig", False), (ftb.TrigInterpolant(True), "trig", True),
             (ftb.EncDecInterpolant(), "encdec", False), (ftb.SBDMInterpolant(), "sbdm", True),
             (ftb.MirrorInterpolant(), "mirror", False)]
    for ip, kind, one in cases:
        mine = torch.stack([ip.alpha(t), ip.beta(t), ip.gamma(t), ip.alpha_dot(t), ip.beta_dot(t), ip.gamma_dot(t)])
        want = torch.stack(interp.coeffs(kind, t, one))
        assert torch.equal(mine, want), kind
        assert ip.is_one_sided() == interp.is_one_sided(kind, one)
    si = ftb.StochasticInterpolator(ftb.LinearInterpolant(False))
    with pytest.raises(ValueError, match="Z must be provided"):
        si.flow_objective(torch.rand(2), torch.zeros(2, 4), torch.zeros(2, 4))
    assert repr(si) == "StochasticInterpolator(LinearInterpolant(one_sided=False))"


def test_simplex_embedding_matches_oracle():
    assert torch.equal(ftb.simplex_embedding(15, 18), task.simplex_embedding(15, 18))
    assert torch.equal(ftb.simplex_embedding(15, 15), task.simplex_embedding(15, 15))


def test_shard_indices_cover_everything_once():
    for n, world in ((64, 1), (64, 8), (10, 4), (3, 8)):
        seen = []
        for r in range(world):
            seen += list(sharding.shard_indices(n, r, world))
        assert sorted(seen) == list(range(n))
    assert list(sharding.shard_indices(10, 1, 4)) == [1, 5, 9]


def test_sharded_ensemble_gloo_world2():
    """world_size-2 gloo run of the N>1 host path: shard sample indices, run a (CPU stand-in)
    per-sample function, gather the decoded volumes and the vote histogram on rank 0."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", script],
                       capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_OK" in r.stdout


def test_bucketed_gradient_allreduce_gloo_world2():
    """Data-parallel training exchange: the gradient ranges the backward reports are coalesced and
    all-reduced exactly once each; world_size-2 gloo group on CPU tensors."""
    script = os.path.join(ROOT, "tests", "_gloo_bucket_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", script],
                       capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BUCKET_OK" in r.stdout


def test_training_plan_sizes_without_gpu():
    """The train-mode tape (forward + backward bookkeeping) is consistent: sizing it needs no GPU."""
    net = ftb.Unet3D(**synth.make_cfg(dropout=0.0))
    n = _lib.lib.ftb_unet3d_train_workspace_bytes(net._handle, 1, 16, 16, 16)
    assert n > _lib.lib.ftb_unet3d_workspace_bytes(net._handle, 1, 16, 16, 16) > 0
    total = _lib.lib.ftb_unet3d_param_offset(net._handle, _lib.lib.ftb_unet3d_num_params(net._handle))
    assert total == 25_193_410


def test_host_jittered_grid_points_match_oracle():
    """The host borehole draw (boreholes.jittered_grid_points, vectorised) equals the reference's per-cell loop
    (oracle restatement of boreholes.py:9-42) on the same uniform numbers."""
    import math
    for (X, Y, n) in ((64, 64, 8), (64, 64, 17), (32, 48, 31), (16, 16, 10)):
        g = torch.Generator().manual_seed(n)
        got = ftb.jittered_grid_points(X, Y, n, g)
        n_x = int(math.floor(math.sqrt(n)))
        n_y = int(math.ceil(n / n_x))
        rand = torch.rand(n_x * n_y, 2, generator=torch.Generator().manual_seed(n))
        want = task.jittered_grid_points(X, Y, n, rand)
        assert torch.equal(got, want), (X, Y, n)
        assert got.shape == (n, 2) and got.min() >= 0 and got[:, 0].max() <= X - 1 and got[:, 1].max() <= Y - 1
    bores, nb = ftb.draw_boreholes(3, 64, 64, torch.Generator().manual_seed(1))
    assert bores.shape == (3, 64, 2) and bores.dtype == torch.int32 and all(8 <= int(v) < 32 for v in nb)


def test_conditioning_frontend_has_no_cpu_path():
    cats = torch.zeros(1, 1, 4, 4, 4, dtype=torch.long)
    with pytest.raises(RuntimeError):
        ftb.make_surface_mask(cats)
    with pytest.raises(RuntimeError):
        ftb.EnsembleVotes(ftb.simplex_embedding(15, 18), (4, 4, 4), "cpu")


def test_lightning_checkpoint_round_trip(tmp_path):
    """A Lightning-style .ckpt (state_dict with net.* / embedding.weight keys + ema_shadow, callbacks.py:295-303) loads
    into the B200 module exactly as model_inference_experiments.py:387-403 does, with and without the EMA shadow."""
    cfg = dict(dim=32, dim_mults=(1, 2), time_resolution=64, time_bandwidth=100.0, time_learned_emb=True,
               attn_heads=2, attn_dim_head=16)
    src = ftb.Geo3DStochInterp(data_shape=(16, 16, 16), embedding_dim=18, **cfg)
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    shadow = {n: p.detach() + 1.0 for n, p in src.named_parameters() if p.requires_grad}
    path = tmp_path / "epoch=1.ckpt"
    torch.save({"state_dict": sd, "ema_shadow": shadow, "ema_update_on_cpu": True, "epoch": 1}, path)
    dst = ftb.Geo3DStochInterp(data_shape=(16, 16, 16), embedding_dim=18, **cfg)
    ftb.load_model_with_ema_option(dst, str(path), use_ema=False)
    for k, v in dst.state_dict().items():
        assert torch.equal(v, sd[k]), k
    ftb.load_model_with_ema_option(dst, str(path), use_ema=True)
    for n, p in dst.named_parameters():
        want = shadow[n] if n in shadow else sd[n]
        assert torch.equal(p.detach(), want), n
    assert "embedding.weight" not in shadow and torch.equal(dst.embedding.weight, sd["embedding.weight"])
    ck = ftb.lightning_checkpoint(dst)
    assert set(ck["state_dict"].keys()) == set(sd.keys())


def test_precision_switch_and_solver_method_validation():
    """Host-side argument checks of the additions that need no GPU: precision names, adaptive-method names, the
    conditional module's signature and the training bridge's refusal to produce input gradients."""
    net = ftb.Unet3D(dim=32, dim_mults=(1, 2), data_channels=18, time_resolution=64, time_learned_emb=True,
                     attn_heads=2, attn_dim_head=16)
    assert net.precision == "bf16" and net.set_precision("fp32") is net and net.precision == "fp32"
    with pytest.raises(ValueError):
        net.set_precision("fp16")
    with pytest.raises(RuntimeError):          # no CPU fallback, also in the fp32 mode
        net(torch.zeros(1, 18, 8, 8, 8), torch.zeros(1))
    from flowtrain_stochastic_interpolation_b200 import solvers
    assert set(solvers.ADAPTIVE_METHODS) == {"dopri5", "adaptive_heun"}
    with pytest.raises(ValueError):
        solvers.integrate_adaptive(lambda t, x, i: x, torch.zeros(4), 0.0, 1.0, 3, method="rk45")
    with pytest.raises(RuntimeError):
        solvers.integrate_adaptive(lambda t, x, i: x, torch.zeros(4), 0.0, 1.0, 3, method="dopri5")
    # tableau consistency: rows of beta sum to alpha, the 5th-order weights sum to 1, error weights to 0
    d = solvers._DOPRI5
    for a, row in zip(d["alpha"], d["beta"]):
        assert abs(sum(row) - a) < 1e-12
    assert abs(sum(d["c_sol"]) - 1.0) < 1e-12 and abs(sum(d["c_error"])) < 1e-12
    cnet = ftb.Unet3DCond(dim=32, dim_mults=(1, 2), data_channels=15, time_resolution=64, time_learned_emb=True,
                          attn_heads=2, attn_dim_head=16)
    assert cnet.set_precision("fp32").precision == "fp32"
    names = [n for n, _ in cnet.named_parameters()]
    assert names[0] == "init_conv_x.weight" and "downs.0.0.conv1.weight" in names and "downs.0.1.time_mlp.1.weight" in names


def test_solver_defaults_follow_the_reference_call():
    """``ODEFlowSolver(model, rtol=1e-6).solve(X0, t0, tf, n_steps=16)`` (model_train_inference.py:615-619) is an
    ADAPTIVE dopri5 solve with 16 output points in the reference (solvers.py:77); the drop-in must not silently turn it
    into 15 Euler steps.  The SDE solver defaults to adaptive_heun (solvers.py:220-222); fixed grids are explicit."""
    import warnings
    ip = ftb.LinearInterpolant(one_sided=True)
    toy = lambda x, t: x
    assert ftb.ODEFlowSolver(toy, rtol=1e-6).method == "dopri5"
    assert ftb.ODEOneSidedDenoisingSolver(toy, ip).method == "dopri5"
    assert ftb.SDEOneSidedDenoisingSolver(toy, ip, epsilon=torch.tensor(0.1)).method == "adaptive_heun"
    assert ftb.ODEFlowSolver(toy, method="heun").method == "heun"
    with pytest.raises(ValueError, match="fixed-grid"):      # tolerances + fixed grid: refuse, do not ignore
        ftb.ODEFlowSolver(toy, rtol=1e-4, method="euler")
    with pytest.raises(ValueError, match="method must be"):
        ftb.ODEFlowSolver(toy, method="rk45")
    with pytest.raises(AssertionError, match="one-sided"):
        ftb.ODEOneSidedDenoisingSolver(toy, ftb.LinearInterpolant(one_sided=False))
    net = ftb.Unet3D(**synth.make_cfg(dim=32, dim_mults=(1, 2), attn_heads=2, attn_dim_head=16, time_resolution=64))
    with warnings.catch_warnings(record=True) as w:          # bf16 field + 1e-6 tolerances: say so
        warnings.simplefilter("always")
        ftb.ODEFlowSolver(net)
        assert any("bf16" in str(x.message) for x in w)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        ftb.ODEFlowSolver(net.set_precision("fp32"))
        ftb.ODEFlowSolver(net.set_precision("bf16"), method="heun")
        assert not w


def test_task_modules_have_the_lightning_surface():
    """training_step / configure_optimizers / checkpoint hooks of the reference LightningModules
    (model_train_inference.py:417-484, model_train_sh_inference_cond.py:401-495): present, same optimiser classes,
    same hyper-parameter plumbing.  (The compute inside training_step is covered by the -m gpu tests.)"""
    small = dict(dim=32, dim_mults=(1, 2), attn_heads=2, attn_dim_head=16, time_resolution=64)
    m = ftb.Geo3DStochInterp(embedding_dim=18, learning_rate=2e-4, lr_decay=0.997, **small)
    for hook in ("training_step", "configure_optimizers", "on_save_checkpoint", "on_load_checkpoint",
                 "on_train_epoch_end", "embed", "decode", "forward"):
        assert callable(getattr(m, hook)), hook
    o = m.configure_optimizers()
    assert type(o["optimizer"]) is torch.optim.Adam and o["optimizer"].param_groups[0]["lr"] == 2e-4
    assert isinstance(o["lr_scheduler"], torch.optim.lr_scheduler.ExponentialLR) and o["lr_scheduler"].gamma == 0.997
    # the optimiser sees exactly the trainable parameters of net (the simplex embedding is frozen, :316)
    n_opt = sum(p.numel() for g in o["optimizer"].param_groups for p in g["params"] if p.requires_grad)
    assert n_opt == sum(p.numel() for p in m.net.parameters() if p.requires_grad)
    ck = {}
    m.ema_shadow = {"net.final_conv.bias": torch.ones(18)}
    m.on_save_checkpoint(ck)
    assert ck["ema_shadow"] is m.ema_shadow
    m2 = ftb.Geo3DStochInterp(embedding_dim=18, **small)
    m2.on_load_checkpoint(ck)
    assert m2.ema_shadow is ck["ema_shadow"]
    c = ftb.Geo3DStochInterpCond(embedding_dim=15, time_learned_emb=True, **small)
    assert c.hparams.learning_rate == 2e-3 and c.hparams.lr_decay == 0.997     # reference defaults :284-285
    oc = c.configure_optimizers()
    assert type(oc["optimizer"]) is torch.optim.AdamW and oc["optimizer"].param_groups[0]["lr"] == 2e-3
    assert callable(c.training_step) and callable(c.on_after_backward)
    with pytest.raises(RuntimeError, match="CUDA"):           # no CPU fallback inside the step either
        m.training_step(torch.zeros(1, 1, 8, 8, 8, dtype=torch.long))


def test_mark_dirty_and_ema_swap_plumbing():
    """``param.data.copy_`` (the reference EMACallback's weight swap) does not bump the version counter the weight
    sync watches; ``mark_dirty`` clears the sync state so the next forward re-reads every parameter, and
    ``load_model_with_ema_option`` calls it."""
    small = dict(dim=32, dim_mults=(1, 2), attn_heads=2, attn_dim_head=16, time_resolution=64)
    m = ftb.Geo3DStochInterp(embedding_dim=18, **small)
    net = m.net
    net._synced = {n: (p.data_ptr(), p._version) for n, p in net.named_parameters()}   # as after a forward
    p = net.final_conv.bias
    v0 = p._version
    p.data.copy_(torch.ones_like(p))
    assert p._version == v0                      # the hazard: invisible to the (data_ptr, _version) key
    net.mark_dirty()
    assert net._synced == {}
    shadow = {f"net.{n}": torch.full_like(q, 0.25) for n, q in net.named_parameters() if q.requires_grad}
    net._synced = {n: (q.data_ptr(), q._version) for n, q in net.named_parameters()}
    ftb.load_model_with_ema_option(m, {"state_dict": m.state_dict(), "ema_shadow": shadow}, use_ema=True)
    assert net._synced == {} and torch.all(net.final_conv.bias == 0.25)
    cnet = ftb.Unet3DCond(data_channels=15, time_learned_emb=True, **small)
    cnet._atb_key = ("stale",)
    cnet.mark_dirty()
    assert cnet._atb_key is None


def test_streaming_dataset_stand_in_contract():
    """the stand-in for geogen's GeoData3DStreamingDataset keeps the item contract the training step
    relies on (model_train_inference.py:249-260, :428): int64 [1, X, Y, Z], categories -1 .. 13, deterministic per
    index, fresh per epoch, usable behind a shuffling multi-worker DataLoader."""
    from torch.utils.data import DataLoader
    ds = ftb.SyntheticGeoStreamingDataset(model_resolution=[1, 16, 12, 20], model_bounds=((-1, 1), (-1, 1), (-1, 1)),
                                          dataset_size=6, device="cpu")
    assert len(ds) == 6
    a = ds[3]
    assert a.shape == (1, 16, 12, 20) and a.dtype == torch.int64
    assert -1 <= int(a.min()) and int(a.max()) <= 13
    assert (a == -1).any() and a.unique().numel() >= 3          # air + several rock units
    assert torch.equal(a, ds[3]) and not torch.equal(a, ds[4])
    ds.set_epoch(1)
    assert not torch.equal(a, ds[3])
    with pytest.raises(IndexError):
        ds[6]
    dl = DataLoader(ds, batch_size=4, shuffle=True, num_workers=2)
    shapes = [tuple(b.shape) for b in dl]
    assert shapes == [(4, 1, 16, 12, 20), (2, 1, 16, 12, 20)]
    loader = ftb.get_data_loader({"data": {"shape": [1, 8, 8, 8], "bounds": None, "epoch_size": 5, "batch_size": 2}})
    assert sum(b.shape[0] for b in loader) == 5
    # embed()'s index rule: cat + 1 must be a valid row of the 15-row embedding
    assert int((a + 1).min()) >= 0 and int((a + 1).max()) <= 14
    with pytest.raises(RuntimeError, match="CUDA"):
        ftb.DevicePrefetcher(loader, "cpu")


def test_bench_reads_roofline_traffic_from_the_ncu_summary():
    """bench.py computes roofline.traffic from the newest committed `ncu --set full` summary of the conv kernel
    (profiles/rNN_conv_igemm_ncu_full.csv): read + write DRAM bytes of every captured launch and their mean."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    t = bench.ncu_traffic()
    assert t is not None and t["source"].startswith("profiles/r") and t["source"].endswith("_conv_igemm_ncu_full.csv")
    assert len(t["launches"]) >= 1 and all(b > 1e8 for b in t["launches"])   # hundreds of MB per 64^3 launch
    assert abs(t["dram_bytes_per_launch"] - sum(t["launches"]) / len(t["launches"])) < 1.0
    # at or below the algorithmic bytes of a 48 -> 48 @64^3 B=8 launch with a residual (604 MB) plus 10 %
    assert t["dram_bytes_per_launch"] < 1.1 * 604e6
