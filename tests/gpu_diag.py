"""GPU bring-up diagnostics (run on the B200 box; not collected by pytest).

    python tests/gpu_diag.py --stage all        # every stage, each in its own subprocess
    python tests/gpu_diag.py --stage conv_igemm # one stage in-process

Each stage prints PASS/FAIL lines with error norms; a CUDA fault in one stage cannot poison the
next because the runner isolates stages in subprocesses with a timeout.
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["elementwise", "conv_naive", "conv_igemm", "conv_shapes", "attention_taps", "unet_naive",
          "unet_igemm", "timing"]


def rel(a, b):
    import torch
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def report(name, err, bar):
    print(f"{'PASS' if err <= bar else 'FAIL'} {name}: err={err:.3e} bar={bar:.1e}", flush=True)
    return err <= bar


def conv_case(B, c1, c2, cout, k, dims, impl, norm=False, film=False, silu=False, resid=False, seed=0):
    import torch
    import torch.nn.functional as F
    from flowtrain_stochastic_interpolation_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator("cpu").manual_seed(seed)
    X, Y, Z = dims
    bf = lambda t: t.to(torch.bfloat16).float()
    x = bf(torch.randn(B, c1, X, Y, Z, generator=g)).to(dev)
    x2 = bf(torch.randn(B, c2, X, Y, Z, generator=g)).to(dev) if c2 else None
    cin = c1 + c2
    w = bf(torch.randn(cout, cin, k, k, k, generator=g) / (cin * k ** 3) ** 0.5).to(dev)
    bias = (torch.randn(cout, generator=g) * 0.1).to(dev)
    gg = (1 + 0.2 * torch.randn(cout, generator=g)).to(dev) if norm else None
    sc = (0.3 * torch.randn(B, cout, generator=g)).to(dev) if film else None
    sh = (0.3 * torch.randn(B, cout, generator=g)).to(dev) if film else None
    rs = bf(torch.randn(B, cout, X, Y, Z, generator=g)).to(dev) if resid else None
    out = torch.full((B, cout, X, Y, Z), float("nan"), device=dev)
    xin = x if x2 is None else torch.cat((x, x2), 1)
    ref = F.conv3d(xin.double(), w.double(), bias.double(), padding=k // 2)
    if norm:
        ref = F.normalize(ref, dim=1) * gg.view(1, -1, 1, 1, 1).double() * cout ** 0.5
    if film:
        ref = ref * (sc.double()[:, :, None, None, None] + 1) + sh.double()[:, :, None, None, None]
    if silu:
        ref = F.silu(ref)
    if resid:
        ref = ref + rs.double()
    _lib.check(_lib.lib.ftb_test_conv3d(
        _lib.ptr(x), c1, _lib.ptr(x2), c2, _lib.ptr(w), _lib.ptr(bias), cout, k, _lib.ptr(gg), _lib.ptr(sc),
        _lib.ptr(sh), _lib.ptr(rs), 1 if silu else 0, _lib.ptr(out), B, X, Y, Z, impl, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return rel(out, ref), out, ref


def stage_elementwise():
    import numpy as np
    import torch
    import flowtrain_stochastic_interpolation_b200 as ftb
    from oracle import interp, synth, task
    dev = torch.device("cuda")
    ok = True
    X0, X1, Z = (synth.synth_input((3, 18, 8, 8, 8), s, n).to(dev) for s, n in ((1, "a"), (2, "b"), (3, "c")))
    T = torch.tensor([0.1, 0.45, 0.8], device=dev)
    for ip, kind, one in ((ftb.LinearInterpolant(True), "linear", True), (ftb.LinearInterpolant(False), "linear", False),
                          (ftb.TrigInterpolant(False), "trig", False), (ftb.EncDecInterpolant(), "encdec", False),
                          (ftb.SBDMInterpolant(), "sbdm", True), (ftb.MirrorInterpolant(), "mirror", False)):
        z = None if one else Z
        xt, bt = ftb.StochasticInterpolator(ip).flow_objective(T, X0, X1, z)
        rxt, rbt = interp.flow_objective(kind, T.cpu(), X0.cpu(), X1.cpu(), None if z is None else z.cpu(), one_sided=one)
        ok &= report(f"interp {kind} one_sided={one} XT", rel(xt.cpu(), rxt), 1e-6)
        ok &= report(f"interp {kind} one_sided={one} BT", rel(bt.cpu(), rbt), 1e-6)
    W = task.simplex_embedding(15, 18)
    x = synth.synth_input((2, 18, 8, 8, 16), 30, "dec").to(dev)
    d = ftb.decode(W.to(dev), x).cpu().numpy()
    want = task.decode_numpy(task.normalized_embedding(W).numpy(), x.cpu().numpy())
    ok &= report("decode mismatches", float((d != want).sum()), 0)
    cats = torch.randint(-1, 14, (2, 1, 4, 4, 8), device=dev)
    e = ftb.embed(W.to(dev), cats)
    ok &= report("embed", rel(e.cpu(), task.embed(W, cats.cpu())), 0)
    s = torch.randn(1000, device=dev); p = torch.randn(1000, device=dev)
    want = task.ema_update(s.cpu(), p.cpu(), 0.9995)
    ftb.ema_update_(s, p, 0.9995)
    ok &= report("ema", rel(s.cpu(), want), 1e-7)
    v = torch.randn(2, 18, 8, 8, 8, device=dev); vh = v + 0.1 * torch.randn_like(v)
    ok &= report("flow_loss", abs(ftb.flow_loss(v, vh).item() - task.flow_loss(v.cpu(), vh.cpu()).item()), 1e-6)
    return ok


def stage_conv_naive():
    ok = True
    for args in [dict(B=1, c1=48, c2=0, cout=48, k=3, dims=(4, 16, 16)),
                 dict(B=2, c1=18, c2=0, cout=48, k=7, dims=(8, 8, 8)),
                 dict(B=1, c1=48, c2=48, cout=48, k=3, dims=(4, 8, 8), norm=True, film=True, silu=True, resid=True)]:
        err, _, _ = conv_case(impl=1, **args)
        ok &= report(f"conv_naive {args}", err, 6e-3)
    return ok


def stage_conv_igemm():
    import torch
    ok = True
    err, out, ref = conv_case(B=1, c1=48, c2=0, cout=48, k=3, dims=(4, 16, 16), impl=0)
    ok &= report("conv_igemm 48->48 k3 4x16x16", err, 6e-3)
    if err > 6e-3:
        print("out[0,:4,0,0,:8]", out[0, :4, 0, 0, :8].cpu())
        print("ref[0,:4,0,0,:8]", ref[0, :4, 0, 0, :8].float().cpu())
        print("nan count", torch.isnan(out).sum().item(), "of", out.numel())
        # which (d,h,w) positions are wrong?
        bad = ((out - ref.float()).abs().amax(dim=1) > 0.05)[0]
        print("bad voxels", bad.sum().item(), "of", bad.numel())
        print("bad per d", bad.sum(dim=(1, 2)).tolist())
        print("bad per h", bad.sum(dim=(0, 2)).tolist())
        print("bad per w", bad.sum(dim=(0, 1)).tolist())
    err, _, _ = conv_case(B=1, c1=48, c2=0, cout=48, k=1, dims=(4, 16, 16), impl=0)
    ok &= report("conv_igemm 48->48 k1 4x16x16", err, 6e-3)
    return ok


def stage_conv_shapes():
    ok = True
    cases = [
        dict(B=2, c1=48, c2=0, cout=48, k=3, dims=(16, 32, 32)),
        dict(B=1, c1=48, c2=48, cout=48, k=3, dims=(8, 32, 16), norm=True, film=True, silu=True),
        dict(B=1, c1=96, c2=0, cout=96, k=3, dims=(8, 16, 16), norm=True, silu=True, resid=True),
        dict(B=1, c1=192, c2=144, cout=192, k=3, dims=(4, 4, 4), norm=True, film=True, silu=True),
        dict(B=1, c1=144, c2=0, cout=192, k=3, dims=(2, 2, 2)),
        dict(B=2, c1=18, c2=0, cout=48, k=7, dims=(16, 16, 16)),
        dict(B=1, c1=48, c2=0, cout=384, k=1, dims=(8, 16, 16)),
        dict(B=1, c1=48, c2=0, cout=18, k=1, dims=(8, 16, 16)),
        dict(B=1, c1=96, c2=48, cout=48, k=1, dims=(4, 8, 8)),
        dict(B=1, c1=15, c2=0, cout=48, k=5, dims=(8, 8, 8)),
        dict(B=1, c1=192, c2=0, cout=192, k=3, dims=(1, 1, 1)),
    ]
    for args in cases:
        try:
            err, _, _ = conv_case(impl=0, **args)
            ok &= report(f"conv_igemm {args}", err, 6e-3)
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"FAIL conv_igemm {args}: {type(e).__name__}: {e}", flush=True)
    import torch
    import torch.nn.functional as F
    from flowtrain_stochastic_interpolation_b200 import _lib
    for (di, do) in (((8, 8, 8), (4, 4, 4)), ((4, 4, 4), (8, 8, 8)), ((16, 16, 16), (32, 32, 32)), ((2, 2, 2), (1, 1, 1))):
        x = torch.randn(2, 48, *di, device="cuda").to(torch.bfloat16).float()
        out = torch.empty(2, 48, *do, device="cuda")
        _lib.check(_lib.lib.ftb_test_trilinear(_lib.ptr(x), 2, 48, *di, *do, _lib.ptr(out), _lib.stream_ptr()))
        ref = F.interpolate(x, size=do, mode="trilinear", align_corners=True)
        ok &= report(f"trilinear {di}->{do}", rel(out, ref), 4e-3)
    return ok


def _unet_compare(dims, B, impl_env, bar, seed_x=1):
    import torch
    if impl_env:
        os.environ["FTB_CONV_IMPL"] = impl_env
    import flowtrain_stochastic_interpolation_b200 as ftb
    from oracle import synth, unet3d
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    cfg = synth.make_cfg()
    params = synth.synth_unet3d_params(cfg, 0)
    net = ftb.Unet3D(**cfg).to(dev)
    net.load_state_dict(params)
    x = synth.synth_input((B, 18) + dims, seed_x).to(dev)
    t = torch.linspace(0.3, 0.8, B, device=dev)
    with torch.no_grad():
        y = net(x, t)
        torch.cuda.synchronize()
        taps = {}
        ref = unet3d.unet3d_forward({k: v.to(dev) for k, v in params.items()}, cfg, x, t, taps)
    ok = True
    for name in taps:
        try:
            got = net.get_tap(name)
        except Exception:  # noqa: BLE001
            continue
        c = taps[name].shape[1]
        e = rel(got[:, :c], taps[name])
        flag = "" if e < 3e-2 else "   <<<<<"
        print(f"  tap {name:28s} rel={e:.3e}{flag}", flush=True)
    ok &= report(f"unet {dims} B={B} impl={impl_env or 'igemm'} output", rel(y, ref), bar)
    print("launches", net.last_launches)
    return ok


def stage_attention_taps():
    return _unet_compare((16, 16, 16), 2, "naive", 2e-2)


def stage_unet_naive():
    return _unet_compare((32, 32, 32), 1, "naive", 2e-2)


def stage_unet_igemm():
    ok = _unet_compare((32, 32, 32), 1, "", 2e-2)
    return ok


def stage_timing():
    import torch
    import flowtrain_stochastic_interpolation_b200 as ftb
    from oracle import synth
    dev = torch.device("cuda")
    cfg = synth.make_cfg()
    net = ftb.Unet3D(**cfg).to(dev)
    net.load_state_dict(synth.synth_unet3d_params(cfg, 0))
    for B, dims in ((1, (64, 64, 64)), (4, (64, 64, 64))):
        x = synth.synth_input((B, 18) + dims, 1).to(dev)
        t = torch.full((B,), 0.5, device=dev)
        with torch.no_grad():
            for _ in range(3):
                net(x, t)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                net(x, t)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"TIMING unet fwd B={B} {dims}: {ms:.2f} ms/eval, {872.7 * B / ms:.1f} TFLOP/s algorithmic", flush=True)
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="all")
    ap.add_argument("--timeout", type=int, default=240)
    a = ap.parse_args()
    if a.stage != "all":
        ok = globals()["stage_" + a.stage]()
        sys.exit(0 if ok else 1)
    results = {}
    for s in STAGES:
        print(f"\n===== stage {s} =====", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", s], timeout=a.timeout)
            results[s] = "ok" if r.returncode == 0 else f"rc={r.returncode}"
        except subprocess.TimeoutExpired:
            results[s] = "timeout"
        print(f"===== stage {s}: {results[s]} ({time.time() - t0:.1f}s)", flush=True)
    print("\nSUMMARY", results)


if __name__ == "__main__":
    main()
