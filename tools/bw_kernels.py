"""One pass over every HBM-bound kernel of the path at the benchmark shapes (B=8, 64^3), bracketed by
cudaProfilerStart/Stop so `ncu --profile-from-start off --set full` captures exactly this pass:
  * one Heun step of the sampler   -> pack_unfold_w, trilinear (down / smem up), normact_fwd, axpy, heun_kernel
  * an RK4 combine, the SDE drift, the adaptive-solver passes (lincomb_dev, error_ratio_dev, advance)
  * decode, decode_logits, decode -> vote histogram + finalize, embed, interpolant
  * one optimiser step (FlowTrainer) -> normact_bwd, trilinear_bwd, chan_sum, mse, mse grad, grad sumsq, adam, ema
usage:  python tools/bw_kernels.py            (plain run, prints CUDA-event timings of the standalone kernels)
        ncu --profile-from-start off --set full --clock-control none -k regex:<bandwidth kernels> ... python tools/bw_kernels.py
Development / evidence tool; synthetic weights from oracle/synth.py only (no oracle compute)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowtrain_stochastic_interpolation_b200 as ftb  # noqa: E402
from flowtrain_stochastic_interpolation_b200 import _lib, solvers  # noqa: E402
from oracle import synth  # noqa: E402

dev = torch.device("cuda:0")
B, S = 8, 64
cfg = synth.make_cfg(dropout=0.1)
kw = {k: v for k, v in cfg.items() if k != "data_channels"}
mod = ftb.Geo3DStochInterp(data_shape=(S, S, S), embedding_dim=18, **kw).to(dev)
mod.net.load_state_dict(synth.synth_unet3d_params(cfg, 0))
net = mod.net.eval()
W = mod.embedding.weight
x = torch.randn(B, 18, S, S, S, device=dev)
cats = torch.randint(-1, 14, (B, 1, S, S, S), device=dev)
T = torch.rand(B, device=dev)
heun = ftb.ODEFlowSolver(net, method="heun")
ip = ftb.LinearInterpolant(one_sided=True)
sde = ftb.SDEOneSidedDenoisingSolver(lambda a, t: a, ip, epsilon=torch.tensor(0.1), method="euler")
toy = lambda a, t: a * 0.5
adaptive = ftb.ODEFlowSolver(toy, atol=1e-3, rtol=1e-3, method="dopri5")
votes = ftb.EnsembleVotes(W, (S, S, S), dev)
interp = ftb.StochasticInterpolator(ftb.LinearInterpolant(one_sided=False))


def standalone():
    with torch.no_grad():
        heun.solve(x, t0=0.1, tf=0.11, n_steps=2, return_trajectory=False)
        k = [torch.randn_like(x) for _ in range(4)]
        solvers._rk4(torch.empty_like(x), x, *k, 0.01)
        sde.solve(x, t0=0.3, tf=0.31, n_steps=2, return_trajectory=False)
        adaptive.solve(x, t0=0.0, tf=0.05, n_steps=2, return_trajectory=False)
        ftb.decode(W, x)
        ftb.decode(W, x, return_logits=True)
        votes.add(x)
        votes.finalize()
        mod.embed(cats)
        interp.flow_objective(T, x, k[0], k[1])


def train_step(tr):
    tr.step(cats)


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


PART = os.environ.get("FTB_BW_PART", "all")    # "standalone" | "train" | "all": what the profiler window covers
standalone()                                   # warm-up (weight packing, attribute setting)
torch.cuda.synchronize()
if PART in ("all", "standalone"):
    torch.cuda.profiler.start()
    standalone()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
if PART in ("all", "train"):
    net.train()
    tr = ftb.FlowTrainer(mod, lr=2e-4, max_grad_norm=1.0, ema_decay=0.9995)
    train_step(tr); train_step(tr)             # warm-up (EMA shadow created on the first step)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    train_step(tr)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()

if not os.environ.get("FTB_BW_NO_TIMING"):     # CUDA-event timings of the kernels that can be called alone (L2 > inputs)
    n = x.numel()
    out = {}
    with torch.no_grad():
        k = [torch.randn_like(x) for _ in range(4)]
        o = torch.empty_like(x)
        out["axpy_us"] = timed(lambda: solvers._axpy(o, x, k[0], 0.01)); out["axpy_bytes"] = 3 * n * 4
        out["heun_us"] = timed(lambda: solvers._heun(o, x, k[0], k[1], 0.01)); out["heun_bytes"] = 4 * n * 4
        out["rk4_us"] = timed(lambda: solvers._rk4(o, x, *k, 0.01)); out["rk4_bytes"] = 6 * n * 4
        out["decode_us"] = timed(lambda: ftb.decode(W, x)); out["decode_bytes"] = n * 4 + B * S ** 3 * 8
        out["decode_vote_us"] = timed(lambda: votes.add(x)); out["decode_vote_bytes"] = n * 4 + 2 * 15 * S ** 3 * 4
        out["interp_us"] = timed(lambda: interp.flow_objective(T, x, k[0], k[1])); out["interp_bytes"] = 5 * n * 4
        out["embed_us"] = timed(lambda: mod.embed(cats)); out["embed_bytes"] = B * S ** 3 * 8 + n * 4
    for name in [k_[:-3] for k_ in list(out) if k_.endswith("_us")]:
        out[name + "_gbs"] = out[name + "_bytes"] / out[name + "_us"] / 1e3
    print(json.dumps(out, indent=1))
