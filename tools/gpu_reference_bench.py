"""Eager PyTorch on the same B200 (SURVEY §8d "GPU reference"): the oracle's functional restatement of Unet3D.forward
runs the reference's own ATen / cuDNN op sequence, so timing it on the GPU is what the unmodified reference would do on
this hardware.  fp32 with cuDNN TF32 on (torch's default, what the reference gets), TF32 off (true fp32) and bf16
autocast; B=8 and B=1 at 64^3; plus one training step (forward + backward + Adam) of the same module in fp32/TF32.
Development tool (not part of bench.py: the oracle is test infrastructure)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, task, unet3d  # noqa: E402

dev = torch.device("cuda:0")
cfg = synth.make_cfg()
params = {k: v.to(dev) for k, v in synth.synth_unet3d_params(cfg, 0).items()}


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
for B in (8, 1):
    x = synth.synth_input((B, 18, 64, 64, 64), 100).to(dev)
    t = torch.full((B,), 0.5, device=dev)
    for name, tf32, ac in (("fp32_tf32_on", True, False), ("fp32_tf32_off", False, False), ("bf16_autocast", True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            ms = timed(lambda: unet3d.unet3d_forward(params, cfg, x, t))
        out[f"eval_ms_B{B}_{name}"] = ms
        out[f"samples_per_s_heun100_B{B}_{name}"] = B / (200 * ms * 1e-3)
# training step, B=4 (B=8 fp32 activations of eager autograd do not leave much headroom for a fair cudnn workspace)
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
for B in (4,):
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in params.items()}
    opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=2e-4)
    xt = synth.synth_input((B, 18, 64, 64, 64), 11, "xt").to(dev)
    vt = synth.synth_input((B, 18, 64, 64, 64), 12, "vt").to(dev)
    t = synth.synth_times(B, 13).to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = task.flow_loss(vt, unet3d.unet3d_forward(p, cfg, xt, t))
        loss.backward()
        torch.nn.utils.clip_grad_norm_([v for v in p.values() if v.requires_grad], 1.0)
        opt.step()
    try:
        ms = timed(step, 3)
        out[f"train_ms_B{B}_fp32_tf32_on"] = ms
        out[f"train_voxels_per_s_B{B}_fp32_tf32_on"] = B * 64 ** 3 / (ms * 1e-3)
        out["train_peak_mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
    except torch.cuda.OutOfMemoryError as e:
        out[f"train_B{B}"] = "OOM"
print(json.dumps(out, indent=1))
