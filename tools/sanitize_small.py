"""Small forward passes (uncond 3-stage ragged + cond small) for compute-sanitizer memcheck."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowtrain_stochastic_interpolation_b200 as ftb
from oracle import synth
dev = torch.device("cuda:0")
cfg = synth.make_cfg(dim=48, dim_mults=(1, 2, 2), time_resolution=128)
net = ftb.Unet3D(**cfg).to(dev); net.load_state_dict(synth.synth_unet3d_params(cfg, 21))
x = synth.synth_input((1, 18, 12, 20, 28), 22).to(dev)
with torch.no_grad():
    y = net(x, torch.tensor([0.3], device=dev))
torch.cuda.synchronize(); print("uncond ok", float(y.abs().mean()))
cfg2 = synth.make_cfg(dim=48, dim_mults=(1, 2), data_channels=15, time_resolution=64)
net2 = ftb.Unet3DCond(**cfg2).to(dev); net2.load_state_dict(synth.synth_unet3d_cond_params(cfg2, 8))
shape = (1, 15, 10, 12, 18)
with torch.no_grad():
    y2 = net2(synth.synth_input(shape, 9).to(dev), synth.synth_atb(shape, 10).to(dev), torch.tensor([0.6], device=dev))
torch.cuda.synchronize(); print("cond ok", float(y2.abs().mean()))
W = ftb.simplex_embedding(15, 18).to(dev)
print("decode ok", int(ftb.decode(W, y).sum()))
