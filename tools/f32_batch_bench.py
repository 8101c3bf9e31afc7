import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowtrain_stochastic_interpolation_b200 as ftb
from oracle import synth
dev = torch.device("cuda:0")
cfg = synth.make_cfg()
net = ftb.Unet3D(**cfg).to(dev); net.load_state_dict(synth.synth_unet3d_params(cfg, 0)); net.set_precision("fp32")
for B in (4, 8):
    x = synth.synth_input((B, 18, 64, 64, 64), 100).to(dev); t = torch.full((B,), 0.5, device=dev)
    with torch.no_grad():
        net(x, t); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): net(x, t)
        e1.record(); torch.cuda.synchronize()
    print(f"fp32 mode 64^3 B={B}: {e0.elapsed_time(e1)/3:.2f} ms/eval, workspace {next(iter(net._workspace_f32.values())).numel()/1e9:.2f} GB")
