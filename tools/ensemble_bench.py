"""BASELINE configs[2]: conditional 64^3 model (Unet3DCond v3, 15-d embedding) with surface + borehole ATb conditioning,
ensemble of 64 samples sharded over the ranks (sample i -> rank i % world, no data-path collective), 100-step ODE,
then the ensemble statistics: decode -> vote histogram per rank (one kernel), ONE all-reduce(sum) of the int32
histogram, probabilities / entropy / most-probable map on every rank.
  python tools/ensemble_bench.py [n_samples] [batch] [steps] [method]        (1 GPU)
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/ensemble_bench.py ...
Side measurement (development tool): prints one JSON object on rank 0."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowtrain_stochastic_interpolation_b200 as ftb  # noqa: E402
from flowtrain_stochastic_interpolation_b200 import sharding  # noqa: E402
from oracle import synth  # noqa: E402  (synthetic weights only)

n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 64
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
method = sys.argv[4] if len(sys.argv) > 4 else "euler"
rank, local, world = sharding.env_rank_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
S = 64
cfg = synth.make_cfg(data_channels=15)
kw = {k: v for k, v in cfg.items() if k != "data_channels"}
mod = ftb.Geo3DStochInterpCond(data_shape=(S, S, S), embedding_dim=15, **kw).to(dev).eval()
mod.net.load_state_dict(synth.synth_unet3d_cond_params(cfg, 5))
# one "true" model and its observations, shared by the whole ensemble (model_inference_experiments.py:228-232)
g = torch.Generator().manual_seed(3)
cats = torch.randint(-1, 14, (1, 1, S, S, S), generator=g).to(dev)
bores, nb = ftb.draw_boreholes(1, S, S, torch.Generator().manual_seed(4))
_, atb, mask = mod.conditioning(cats, bores, nb)
mine = list(sharding.shard_indices(n_samples, rank, world))
votes = ftb.EnsembleVotes(mod.embedding.weight, (S, S, S), dev)
solver = ftb.ODEFlowSolver(lambda x, t: mod.net(x, atb, t), method=method)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.no_grad():
    for i in range(0, len(mine), batch):
        idx = mine[i:i + batch]
        x0 = torch.stack([torch.randn(15, S, S, S, generator=torch.Generator().manual_seed(42 + j)) for j in idx]).to(dev)
        xe = solver.solve(x0, t0=0.001, tf=1.0, n_steps=steps + 1, return_trajectory=False)   # seeds 42 + i (:307)
        votes.add(xe)
    e1.record()
    torch.cuda.synchronize()
    solve_s = e0.elapsed_time(e1) / 1e3
    s0 = time.perf_counter()
    votes.all_reduce()
    stats = votes.finalize()
    torch.cuda.synchronize()
    stats_s = time.perf_counter() - s0
tot = torch.tensor([solve_s, stats_s, time.perf_counter() - t0], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
if rank == 0:
    evals = {"euler": 1, "heun": 2, "rk4": 4}[method] * steps
    print(json.dumps({
        "workload": f"configs[2]: Unet3DCond v3 64^3, ensemble of {n_samples}, {steps}-step {method}, batch {batch}/GPU, "
                    f"{world} rank(s), shared ATb (cached ATb branch)",
        "samples_per_s": n_samples / tot[2].item(), "solve_s": tot[0].item(), "ensemble_stats_s": tot[1].item(),
        "ms_per_eval": tot[0].item() * 1e3 / (evals * ((len(mine) + batch - 1) // batch)),
        "votes_total": int(votes.samples), "mean_entropy": float(stats["entropy"].mean()),
        "counts_sum_ok": bool(int(votes.counts.sum()) == n_samples * S ** 3)}))
if world > 1:
    dist.destroy_process_group()
