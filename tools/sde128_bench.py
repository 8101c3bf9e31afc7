"""BASELINE configs[4]: stochastic-interpolant SDE sampling at 128^3 (same architecture, larger volume): the one-sided
denoising SDE of solvers.py:153-222 (eps = tensor(0.1), explicit noise draw per evaluation) with the fixed-step Heun
integrator, B = 1 and 2.  Side measurement (development tool)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowtrain_stochastic_interpolation_b200 as ftb  # noqa: E402
from oracle import synth  # noqa: E402  (synthetic weights only)

dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cfg = synth.make_cfg()
net = ftb.Unet3D(**cfg).to(dev).eval()
net.load_state_dict(synth.synth_unet3d_params(cfg, 0))
out = {}
for B in (1, 2):
    x0 = torch.randn(B, 18, 128, 128, 128, device=dev)
    sde = ftb.SDEOneSidedDenoisingSolver(net, ftb.LinearInterpolant(one_sided=True), epsilon=torch.tensor(0.1), method="heun")
    with torch.no_grad():
        sde.solve(x0, t0=0.05, tf=0.1, n_steps=3, return_trajectory=False)   # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        xe = sde.solve(x0, t0=0.05, tf=0.95, n_steps=steps + 1, return_trajectory=False)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    gf = 872.7 * 8 * B * 2 * steps
    out[f"B{B}"] = {"ms_per_heun_step": ms / steps, "ms_per_eval": ms / steps / 2, "tflops": gf / ms,
                    "samples_per_s_100_steps": B / (ms / steps * 100 / 1e3), "finite": bool(torch.isfinite(xe).all())}
print(json.dumps({"workload": f"configs[4]: SDE one-sided denoising sampler, 128^3, heun, {steps} steps timed", **out}, indent=1))
