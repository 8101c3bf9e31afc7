"""torchrun worker (one rank per GPU, NCCL): FlowTrainer's data-parallel step.  Checks that the all-reduced
gradient equals the sum of the per-rank gradients (recomputed locally without communication) and that every
rank ends the step with identical parameters."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowtrain_stochastic_interpolation_b200 as ftb  # noqa: E402
from oracle import synth  # noqa: E402  (synthetic weights / inputs only)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = synth.make_cfg(dim=32, dim_mults=(1, 2), data_channels=18, time_resolution=64, time_bandwidth=100.0,
                         attn_heads=2, attn_dim_head=16, dropout=0.0)
    kw = {k: v for k, v in cfg.items() if k != "data_channels"}

    def make(distributed):
        mod = ftb.Geo3DStochInterp(data_shape=(16, 16, 16), embedding_dim=18, **kw).to(dev)
        mod.net.load_state_dict(synth.synth_unet3d_params(cfg, 3))
        return ftb.FlowTrainer(mod, lr=2e-4, max_grad_norm=1.0, ema_decay=None, distributed=distributed)

    def draws(r):
        g = torch.Generator("cpu").manual_seed(50 + r)
        batch = torch.randint(-1, 14, (2, 1, 16, 16, 16), generator=g).to(dev)
        return (batch, synth.synth_input((2, 18, 16, 16, 16), 60 + r, "n1").to(dev),
                synth.synth_input((2, 18, 16, 16, 16), 70 + r, "x0").to(dev), synth.synth_times(2, 80 + r).to(dev))

    tr = make(True)
    tr.broadcast_parameters(0)
    b, n1, x0, T = draws(rank)
    loss = tr.step(b, noise1=n1, X0=x0, T=T)
    gsum = tr.gflat.clone()
    # local recomputation of every rank's gradient, no communication
    want = torch.zeros_like(gsum)
    for r in range(world):
        loc = make(False)
        bb, nn1, xx0, TT = draws(r)
        loc.step(bb, noise1=nn1, X0=xx0, T=TT)
        want += loc.gflat
    err = ((gsum - want).double().norm() / want.double().norm()).item()
    assert err <= 2e-3, f"rank {rank}: all-reduced gradient differs from the sum of the local ones: {err:.3e}"
    flats = [torch.empty_like(tr.flat) for _ in range(world)]
    dist.all_gather(flats, tr.flat)
    for f in flats[1:]:
        assert torch.equal(f, flats[0]), "parameters diverged across ranks"
    assert torch.isfinite(loss)
    if rank == 0:
        print(f"DDP_OK world={world} grad_sum_rel_err={err:.2e}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
