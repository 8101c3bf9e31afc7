"""One fp32-mode evaluation at 64^3, B=1 (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowtrain_stochastic_interpolation_b200 as ftb  # noqa: E402
from oracle import synth  # noqa: E402

dev = torch.device("cuda:0")
cfg = synth.make_cfg()
net = ftb.Unet3D(**cfg).to(dev)
net.load_state_dict(synth.synth_unet3d_params(cfg, 0))
net.set_precision("fp32")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = synth.synth_input((B, 18, 64, 64, 64), 100).to(dev)
t = torch.full((B,), 0.5, device=dev)
with torch.no_grad():
    for _ in range(2):
        y = net(x, t)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
