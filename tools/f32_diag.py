"""fp32-mode diagnostics (run on the GPU box with FTB_F32_KEEP=1): per-tap relative L2 of the fp32 accuracy mode
against the oracle (fp32, TF32 off) in execution order, and the cost of one fp32-mode evaluation at 64^3."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowtrain_stochastic_interpolation_b200 as ftb  # noqa: E402
from oracle import synth, unet3d  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


cfg = synth.make_cfg()
params = synth.synth_unet3d_params(cfg, 0)
net = ftb.Unet3D(**cfg).to(dev)
net.load_state_dict(params)
net.set_precision("fp32")
x = synth.synth_input((2, 18, 16, 16, 16), 7).to(dev)
t = torch.tensor([0.11, 0.77], device=dev)
with torch.no_grad():
    y = net(x, t)
    taps = {}
    ref = unet3d.unet3d_forward({k: v.to(dev) for k, v in params.items()}, cfg, x, t, taps)
for name, want in taps.items():
    try:
        got = net.get_tap_f32(name)
    except Exception:
        continue
    if got.shape == want.shape:
        print(f"{name:32s} {rel(got, want):.3e}")
print(f"output {rel(y, ref):.3e}  launches {net.last_launches}")
for B in (1, 2):
    x = synth.synth_input((B, 18, 64, 64, 64), 100).to(dev)
    t = torch.full((B,), 0.5, device=dev)
    with torch.no_grad():
        net.set_precision("fp32")
        net(x, t)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            net(x, t)
        torch.cuda.synchronize()
        f32_ms = (time.perf_counter() - t0) / 3 * 1e3
        net.set_precision("bf16")
        net(x, t)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            net(x, t)
        torch.cuda.synchronize()
        bf_ms = (time.perf_counter() - t0) / 3 * 1e3
    print(f"64^3 B={B}: fp32 mode {f32_ms:.2f} ms/eval, bf16 {bf_ms:.2f} ms/eval")
