"""Side measurements for the conditional project (not the headline bench): CondFlowTrainer step at 64^3
(BASELINE configs[2]/[3] shapes: 15-d embedding, AdamW, clip 0.3, EMA), and the HBM-bound kernels either side of the
network (conditioning front-end, conditional loss, decode -> vote histogram, vote statistics) against their
algorithmic bytes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowtrain_stochastic_interpolation_b200 as ftb  # noqa: E402
from flowtrain_stochastic_interpolation_b200 import _lib  # noqa: E402
from oracle import synth  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = 64


def timeit(fn, n=10, flush=None):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if flush is not None:
            flush.zero_()   # > L2: the timed kernel starts from HBM
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
W = ftb.simplex_embedding(15, 15)
cats = torch.randint(-1, 14, (B, 1, S, S, S), generator=torch.Generator().manual_seed(1)).to(dev)
bores, nb = ftb.draw_boreholes(B, S, S, torch.Generator().manual_seed(2))
bores_d, nb_d = bores.to(dev), nb.to(dev)
n = S ** 3
ms = timeit(lambda: ftb.conditioning_frontend(cats, W, bores_d, nb_d), flush=flush)
byt = B * n * (8 + 1 + 2 * 15 * 4)
out["cond_frontend"] = {"ms": ms, "GBps": byt / ms / 1e6, "bytes": byt}
x = torch.randn(B, 15, S, S, S, device=dev)
votes = ftb.EnsembleVotes(W, (S, S, S), dev)
ms = timeit(lambda: votes.add(x), flush=flush)
byt = n * (B * 15 * 4 + 2 * 15 * 4)
out["decode_vote"] = {"ms": ms, "GBps": byt / ms / 1e6, "bytes": byt, "samples": B}
ms = timeit(lambda: votes.finalize(), flush=flush)
byt = n * (15 * 4 + 15 * 4 + 4 + 4 + 8)
out["vote_finalize"] = {"ms": ms, "GBps": byt / ms / 1e6, "bytes": byt}
vt, vh, xt, x1 = (torch.randn(B, 15, S, S, S, device=dev) for _ in range(4))
mask = (torch.rand(B, S, S, S, device=dev) < 0.1).to(torch.uint8)
T = torch.rand(B, device=dev)
acc = torch.zeros(6, dtype=torch.float64, device=dev)
dout = torch.empty_like(vh)
st = _lib.stream_ptr()
ms = timeit(lambda: _lib.check(_lib.lib.ftb_cond_loss_accumulate(_lib.ptr(vt), _lib.ptr(vh), _lib.ptr(xt), _lib.ptr(x1),
                                                                 _lib.ptr(x1), _lib.ptr(mask), _lib.ptr(T), B, 15, n,
                                                                 _lib.ptr(acc), st)), flush=flush)
byt = B * n * (15 * 4 * 3 + 1 + 0.1 * 15 * 4 * 2)
out["cond_loss_accumulate"] = {"ms": ms, "GBps": byt / ms / 1e6, "bytes": byt}
ms = timeit(lambda: _lib.check(_lib.lib.ftb_cond_loss_grad(_lib.ptr(vt), _lib.ptr(vh), _lib.ptr(xt), _lib.ptr(x1),
                                                           _lib.ptr(mask), _lib.ptr(T), B, 15, n, _lib.ptr(acc), 1.0, 1.0,
                                                           _lib.ptr(dout), st)), flush=flush)
byt = B * n * (15 * 4 * 3 + 1 + 0.1 * 15 * 4 * 2)
out["cond_loss_grad"] = {"ms": ms, "GBps": byt / ms / 1e6, "bytes": byt}
del x, vt, vh, xt, x1, dout, votes
torch.cuda.empty_cache()

# conditional training step (dropout 0.1, as the reference config)
cfg = synth.make_cfg(data_channels=15, dropout=0.1)
kw = {k: v for k, v in cfg.items() if k != "data_channels"}
mod = ftb.Geo3DStochInterpCond(data_shape=(S, S, S), embedding_dim=15, **kw).to(dev)
mod.net.load_state_dict(synth.synth_unet3d_cond_params(cfg, 5))
tr = ftb.CondFlowTrainer(mod, lr=1e-3, max_grad_norm=0.3, ema_decay=0.9995)
for _ in range(3):
    loss = tr.step(cats)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    loss = tr.step(cats)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
out["cond_train_step"] = {"batch": B, "ms_per_step": ms, "voxels_per_s": B * n / ms * 1e3, "loss": float(loss),
                          "workspace_GB": next(iter(mod.net._workspace.values())).numel() / 1e9,
                          "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
print(json.dumps(out, indent=1))
