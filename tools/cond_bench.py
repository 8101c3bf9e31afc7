"""Configs[2] / [4] side measurements (not the headline bench): conditional Unet3DCond v3 at 64^3 with a
shared ATb (ensemble of B samples on one GPU), cold (ATb branch recomputed) vs cached, and 128^3."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowtrain_stochastic_interpolation_b200 as ftb
from oracle import synth
dev = torch.device("cuda:0")


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = synth.make_cfg(data_channels=15)
net = ftb.Unet3DCond(**cfg).to(dev); net.load_state_dict(synth.synth_unet3d_cond_params(cfg, 5))
x = synth.synth_input((B, 15, 64, 64, 64), 1).to(dev)
atb = synth.synth_atb((1, 15, 64, 64, 64), 2).to(dev)
t = torch.full((B,), 0.4, device=dev)
with torch.no_grad():
    cached = timeit(lambda: net(x, atb, t))
    def cold():
        atb.add_(0.0)   # version bump -> ATb branch recomputed
        return net(x, atb, t)
    cold_ms = timeit(cold)
gf_hoisted, gf_atb = 1109.1, 576.0   # SURVEY §8d per sample / per ATb volume
print(f"Unet3DCond 64^3 B={B} shared ATb: cached {cached:.2f} ms/eval ({gf_hoisted*B/cached:.0f} TFLOP/s-equivalent... "
      f"{gf_hoisted*B/cached/1e0:.0f} GF/ms), cold {cold_ms:.2f} ms/eval (ATb branch {cold_ms-cached:.2f} ms, {gf_atb/(cold_ms-cached):.0f} GF/ms)")
del net
cfg = synth.make_cfg()
net = ftb.Unet3D(**cfg).to(dev); net.load_state_dict(synth.synth_unet3d_params(cfg, 0))
for (b, s) in ((1, 128), (2, 128), (1, 64)):
    x = synth.synth_input((b, 18, s, s, s), 3).to(dev)
    t = torch.full((b,), 0.4, device=dev)
    with torch.no_grad():
        ms = timeit(lambda: net(x, t))
    gf = 872.7 * (s / 64) ** 3 * b
    print(f"Unet3D {s}^3 B={b}: {ms:.2f} ms/eval = {gf/ms:.0f} GF/ms (TFLOP/s)")
# B=1 latency: eager launch sequence vs one CUDA graph per evaluation
x = synth.synth_input((1, 18, 64, 64, 64), 3).to(dev)
t = torch.full((1,), 0.4, device=dev)
with torch.no_grad():
    eager = timeit(lambda: net(x, t), 20)
    net.enable_cuda_graph(True)
    graphed = timeit(lambda: net(x, t), 20)
    net.enable_cuda_graph(False)
print(f"Unet3D 64^3 B=1: eager {eager:.2f} ms/eval, CUDA graph {graphed:.2f} ms/eval")
