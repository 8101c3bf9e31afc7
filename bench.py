#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 hot path (contract: task prompt, section 4).

Metric (BASELINE.json): 64^3 samples/s for a 100-step ODE solve of the unconditional Unet3D
(base 48, mults 1,2,2,3,4, 18-d embedding).  Workload at every N: BASELINE configs[1] —
batch 8 per GPU, Heun integrator (2 velocity evaluations per step), bf16 tensor-core convs with
fp32 state.  One bench "step" = ONE integrator step of the whole batch (the hot path once over one
batch: 2 x v_theta(x,t) + the fused stage kernels); the metric follows as

    samples/s = n_gpus * B / (100 * step_seconds + decode_seconds)

with the decode of the final state (one kernel, timed separately, in the same run) included.
Inputs larger than L2 (B=8 fp32 state 151 MB, every 64^3 activation 201 MB > 126 MB L2), so no
explicit L2 flush between iterations.

`value`  : device-resident state, CUDA events, barrier + synchronize on both sides, max over ranks.
`e2e`    : the same metric through the public API a flowtrain user calls
           (ODEFlowSolver(net).solve -> decode) starting from PINNED HOST noise and ending with the
           decoded int64 volume back on the host, for the FULL 100-step solve; "step" bytes are
           the solve's H2D/D2H bytes.
`roofline`: dominant kernel = conv_igemm (3^3/5^3/7^3 implicit-GEMM convs): algorithmic FLOPs of
           its launches / their CUDA-event durations, measured in a profile pass of the same step
           right after the timed region (events per launch on the launching stream).
`cpu_baseline` / `--impl reference`: the oracle port of the reference's CPU path (torch CPU ops,
           all host threads) on a bounded sample: single velocity evaluations at B=1.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "samples_per_sec_64cubed_100step_ode"
UNIT = "samples/s"
N_ODE_STEPS = 100
T0, TF = 0.001, 1.0                      # model_train_inference.py:618
GF_PER_EVAL = 872.7                      # algorithmic GFLOP per velocity evaluation per sample @64^3
EVALS_PER_STEP = {"euler": 1, "heun": 2, "rk4": 4}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--method", default="heun", choices=list(EVALS_PER_STEP))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-evals", type=int, default=2, help="timed CPU velocity evaluations in the cpu_baseline leg")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (train voxels/s)")
    ap.add_argument("--train-only", action="store_true", help="only the training-step leg (development)")
    ap.add_argument("--train-batch", type=int, default=8)
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the side measurements (fp32 accuracy mode, conditional training step)")
    return ap.parse_args()


def workload_name(a):
    return (f"configs[1]: unconditional {a.size}^3 sampling, batch {a.batch}/GPU, {a.method} integrator, "
            f"{N_ODE_STEPS}-step ODE, bf16 (Unet3D dim 48 mults 1,2,2,3,4, 18 ch)")


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []          # (arrival time, csv line)
        self.proc = None
        self.index = index
        self.t_begin = self.t_end = None

    def start(self):
        """Started BEFORE the warm-up: nvidia-smi needs a few hundred ms to deliver its first sample, longer than a short
        timed region; only samples that arrive between mark_begin() and mark_end() are reported."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def wait_ready(self, timeout=4.0):
        """Block until nvidia-smi has delivered its first sample (so the short timed region is covered)."""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        t1 = self.t_end if self.t_end is not None else float("inf")
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.02]
        window = "timed region"
        if len(rows) < 2:   # very short timed region: fall back to every sample taken under load (warm-up included)
            rows = [r for (_, r) in self.rows]
            window = "warm-up + timed region (fewer than 2 samples arrived inside the timed region)"
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "bf16_burst": d["bf16_tflops"], "src": "MEASURED_PEAKS.json (sustained bf16: kernel timed inside a long step)"}
    return {"hbm_gbs": 6650.0, "bf16": 1590.0, "bf16_burst": 1590.0, "src": "fallback (B200_PROFILING.md)"}


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel, read from the newest committed `ncu --set full` summary
    (profiles/rNN_conv_igemm_ncu_full.csv, written by tools/ncu_summary.py from the capture of the same round):
    dram__bytes_read.sum + dram__bytes_write.sum of every captured conv launch, and their mean."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9][0-9]_conv_igemm_ncu_full.csv")))
    if not files:
        return None
    with open(files[-1]) as f:
        rows = list(csv.reader(f))
    hdr = rows[0]

    def col(prefix):
        for i, h in enumerate(hdr):
            if h.startswith(prefix):
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[h.split("[")[1].rstrip("]")]
                return i, scale
        return None, 1.0

    (ri, rs), (wi, ws) = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    if ri is None or wi is None:
        return None
    per = [float(r[ri]) * rs + float(r[wi]) * ws for r in rows[1:] if len(r) > max(ri, wi)]
    if not per:
        return None
    return {"dram_bytes_per_launch": sum(per) / len(per), "launches": [int(x) for x in per],
            "source": os.path.relpath(files[-1], ROOT)}


# ---------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_velocity_evals(n_timed, size, threads=None):
    """Times the oracle's CPU restatement of Unet3D.forward (the reference's CPU path) at B=1."""
    import torch
    from oracle import synth, unet3d
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = synth.make_cfg()
    params = synth.synth_unet3d_params(cfg, 0)
    x = synth.synth_input((1, 18, size, size, size), 100)
    t = torch.tensor([0.5])
    times = []
    with torch.no_grad():
        unet3d.unet3d_forward(params, cfg, x, t)  # warm-up (oneDNN primitive creation)
        for _ in range(n_timed):
            t0 = time.perf_counter()
            unet3d.unet3d_forward(params, cfg, x, t)
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import synth, unet3d
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = synth.make_cfg()
    params = synth.synth_unet3d_params(cfg, 0)
    x = synth.synth_input((1, 18, a.size, a.size, a.size), 100)
    t = torch.tensor([0.5])
    per = EVALS_PER_STEP[a.method]
    with torch.no_grad():
        for _ in range(a.warmup):
            unet3d.unet3d_forward(params, cfg, x, t)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            unet3d.unet3d_forward(params, cfg, x, t)
        dt = (time.perf_counter() - t0) / a.steps
    # one bench step of the reference arm = ONE velocity evaluation at B=1 (bounded sample of the
    # 100 x `per` evaluations of a solve); integrator axpys and decode are < 0.1% on the CPU
    value = 1.0 / (N_ODE_STEPS * per * dt)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "reference_step": "one CPU velocity evaluation, B=1"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{a.steps} velocity evaluations at B=1 {a.size}^3 fp32 (oracle port of "
                                   f"Unet3D.forward on torch CPU ops); samples/s = 1/({N_ODE_STEPS}*{per}*s_per_eval)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------- training-step leg
TRAIN_GF_PER_SAMPLE = 3 * GF_PER_EVAL - 155.4   # fwd + dgrad + wgrad; the stem has no data gradient


def run_train_leg(a, ftb, _lib, dev, rank, world, dist):
    """BASELINE configs[3]: unconditional 64^3 interpolant training step, batch 8 per GPU, Adam lr 2e-4,
    clip 1.0, EMA 0.9995 every batch, data-parallel NCCL all-reduce overlapped with the backward.
    One step = FlowTrainer.step on a fresh synthetic category batch (embed, noise, interpolant, forward,
    loss, backward, all-reduce, clip + Adam, EMA).  value = world*B*V / step time (max over ranks)."""
    import ctypes as C
    import torch
    from oracle import synth
    cfg = synth.make_cfg(dropout=0.1)   # get_config() of the reference (model_train_inference.py:40-127)
    kw = {k: v for k, v in cfg.items() if k != "data_channels"}
    S, B = a.size, a.train_batch
    mod = ftb.Geo3DStochInterp(data_shape=(S, S, S), embedding_dim=18, **kw).to(dev)
    mod.net.load_state_dict(synth.synth_unet3d_params(cfg, 0))
    tr = ftb.FlowTrainer(mod, lr=2e-4, max_grad_norm=1.0, ema_decay=0.9995, ema_start_step=0)
    tr.broadcast_parameters(0)
    g = torch.Generator("cpu").manual_seed(1000 + rank)
    nb = 4
    host_batches = [torch.randint(-1, 14, (B, 1, S, S, S), generator=g, dtype=torch.int64).pin_memory() for _ in range(nb)]
    dev_batches = [b.to(dev) for b in host_batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(a.warmup, 1)):
        tr.step(dev_batches[i % nb])
    barrier()
    l0 = _lib.lib.ftb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        loss = tr.step(dev_batches[i % nb])
    e1.record()
    barrier()
    step_ms = e0.elapsed_time(e1) / a.steps
    launches = (_lib.lib.ftb_launch_count() - l0) / a.steps
    if os.environ.get("FTB_BENCH_MINIMAL"):   # profiling runs: only the warm-up and the timed steps
        return {"ms_per_step": step_ms, "gpu_launches_per_step": launches}
    # e2e: host batch in (pinned, H2D inside the timed region), loss scalar back on the host, every step
    barrier()
    loss_pinned = torch.empty(a.steps, dtype=torch.float32).pin_memory()
    w0 = time.perf_counter()
    for i in range(a.steps):
        loss_dev = tr.step(host_batches[i % nb].to(dev, non_blocking=True))
        loss_pinned[i].copy_(loss_dev, non_blocking=True)   # D2H of the step's loss, every step, without stalling the host
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - w0) * 1e3 / a.steps
    loss_host = float(loss_pinned[-1])
    # per-kernel-class times of one more step (CUDA events around every conv / wgrad launch)
    _lib.lib.ftb_profile_enable(1)
    tr.step(dev_batches[0])
    torch.cuda.synchronize()
    nk = 3
    fl, by, ms = (C.c_double * nk)(), (C.c_double * nk)(), (C.c_double * nk)()
    ln = (C.c_int * nk)()
    _lib.check(_lib.lib.ftb_profile_collect(fl, by, ms, ln, nk))
    _lib.lib.ftb_profile_enable(0)
    stats = torch.tensor([step_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms = stats.tolist()
    peaks = measured_peaks()
    V = S ** 3
    tf = TRAIN_GF_PER_SAMPLE * B / step_ms
    del tr, mod
    torch.cuda.empty_cache()
    return {
        "metric": "train_voxels_per_sec_64cubed", "value": world * B * V / (step_ms * 1e-3), "unit": "voxels/s",
        "ms_per_step": step_ms, "scaling": "weak", "dtype": "bf16 tensor cores, fp32 master weights / Adam / EMA",
        "config": {"workload": f"configs[3]: unconditional {S}^3 interpolant training step, batch {B}/GPU, Adam 2e-4, "
                               f"clip 1.0, EMA 0.9995 every batch, dropout 0.1, {world} rank(s)"
                               + (", NCCL all-reduce bucketed from inside the backward" if world > 1 else ""),
                   "global_batch": world * B},
        "loss": loss_host, "gpu_launches_per_step": launches,
        "e2e": {"value": world * B * V / (e2e_ms * 1e-3), "unit": "voxels/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": B * V * 8, "d2h_bytes_per_step": 4,
                "what": "FlowTrainer.step from a pinned host int64 batch (H2D every step) to the loss scalar in pinned host "
                        "memory (asynchronous D2H every step, one synchronisation at the end of the timed loop)"},
        "roofline": {"bound": "tensor", "algorithmic_gflop_per_step": TRAIN_GF_PER_SAMPLE * B,
                     "achieved": tf, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": tf / peaks["bf16"],
                     "conv_fwd_dgrad": {"ms": ms[0], "launches": int(ln[0]), "tflops": fl[0] / (ms[0] * 1e-3) / 1e12 if ms[0] > 0 else None},
                     "conv1x1": {"ms": ms[1], "launches": int(ln[1])},
                     "wgrad": {"ms": ms[2], "launches": int(ln[2]), "tflops": fl[2] / (ms[2] * 1e-3) / 1e12 if ms[2] > 0 else None}},
    }



def run_ensemble_leg(a, ftb, dev, rank, world, dist, n_samples=64):
    """BASELINE configs[2]: conditional 64^3 model (Unet3DCond v3, 15-d embedding, surface + borehole ATb at every
    resolution), ensemble of 64 samples sharded over the ranks (sample i -> rank i % world, no data-path collective),
    100-step Euler ODE per sample (seeds 42 + i, model_inference_experiments.py:307), then the ensemble statistics:
    decode -> vote histogram (one kernel per batch), ONE all-reduce(sum) of the int32 histogram, probabilities /
    entropy / most-probable map.  STRONG scaling: the ensemble size is fixed as N grows.  Timed with CUDA events from
    the first H2D copy of pinned host noise to the finalised statistics (all-reduce inside), max over ranks."""
    import torch
    from flowtrain_stochastic_interpolation_b200 import sharding
    from oracle import synth
    S, batch, steps = a.size, a.batch, N_ODE_STEPS
    cfg = synth.make_cfg(data_channels=15)
    kw = {k: v for k, v in cfg.items() if k != "data_channels"}
    mod = ftb.Geo3DStochInterpCond(data_shape=(S, S, S), embedding_dim=15, **kw).to(dev).eval()
    mod.net.load_state_dict(synth.synth_unet3d_cond_params(cfg, 5))
    cats = torch.randint(-1, 14, (1, 1, S, S, S), generator=torch.Generator().manual_seed(3)).to(dev)
    bores, nb = ftb.draw_boreholes(1, S, S, torch.Generator().manual_seed(4))
    with torch.no_grad():
        _, atb, _ = mod.conditioning(cats, bores, nb)     # one conditioning volume shared by the ensemble (:228-232)
        mine = list(sharding.shard_indices(n_samples, rank, world))
        host = [torch.stack([torch.randn(15, S, S, S, generator=torch.Generator().manual_seed(42 + j)) for j in mine[i:i + batch]]).pin_memory()
                for i in range(0, len(mine), batch)]
        solver = ftb.ODEFlowSolver(lambda x, t: mod.net(x, atb.expand(x.shape[0], -1, -1, -1, -1), t), method="euler")
        solver.solve(host[0].to(dev), t0=T0, tf=T0 + 3 * (TF - T0) / steps, n_steps=4, return_trajectory=False)   # warm-up
        votes = ftb.EnsembleVotes(mod.embedding.weight, (S, S, S), dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        for hb in host:
            xe = solver.solve(hb.to(dev, non_blocking=True), t0=T0, tf=TF, n_steps=steps + 1, return_trajectory=False)
            votes.add(xe)
        e1.record()
        votes.all_reduce()
        stats = votes.finalize()
        most_host = stats["most_probable"].cpu()            # D2H of the ensemble's result map
        e2.record()
        torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2), e0.elapsed_time(e2)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    solve_ms, stats_ms, tot_ms = t.tolist()
    ok = bool(int(votes.counts.sum()) == n_samples * S ** 3)
    n_local_batches = len(host)
    del mod, solver, votes
    torch.cuda.empty_cache()
    return {"metric": "ensemble_samples_per_sec_cond_64cubed_100step_ode", "value": n_samples / (tot_ms * 1e-3),
            "unit": "samples/s", "scaling": "strong", "n_gpus": world,
            "config": {"workload": f"configs[2]: Unet3DCond v3 {S}^3 (15 ch), ensemble of {n_samples} sharded over {world} "
                                   f"rank(s), batch {batch}/GPU, {steps}-step euler, one shared ATb (ATb-only branch "
                                   "computed once per trajectory), decode -> vote histogram, one all-reduce"},
            "solve_ms": solve_ms, "ensemble_stats_ms": stats_ms, "total_ms": tot_ms,
            "ms_per_eval": solve_ms / (steps * n_local_batches), "votes_ok": ok,
            "h2d_bytes": len(mine) * 15 * S ** 3 * 4, "d2h_bytes": int(most_host.numel()) * 8,
            "collective": "all_reduce(sum) of the [15, 64^3] int32 vote histogram" if world > 1 else "none (1 rank)"}

def run_extras(a, ftb, dev):
    """Side measurements carried in the JSON line under "extras" (never the headline): one velocity evaluation in
    the fp32 accuracy mode vs bf16 at B=1, and one optimiser step of the conditional project (BASELINE configs[2]'s
    model: Unet3DCond v3, 15-d embedding; CondFlowTrainer: AdamW 1e-3, clip 0.3, EMA, dropout 0.1)."""
    import torch
    from oracle import synth
    S = a.size
    out = {}

    def timed(fn, n):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    cfg = synth.make_cfg()
    net = ftb.Unet3D(**cfg).to(dev).eval()
    net.load_state_dict(synth.synth_unet3d_params(cfg, 0))
    x = synth.synth_input((1, 18, S, S, S), 100).to(dev)
    t = torch.full((1,), 0.5, device=dev)
    with torch.no_grad():
        bf = timed(lambda: net(x, t), 5)
        net.set_precision("fp32")
        f32 = timed(lambda: net(x, t), 5)
    out["fp32_mode"] = {"ms_per_eval_b1": f32, "bf16_ms_per_eval_b1": bf, "ratio": f32 / bf,
                        "what": "Unet3D.set_precision('fp32'): 3 x bf16 split convs, fp32 elsewhere (<= 1e-4 rel-L2)"}
    del net, x
    torch.cuda.empty_cache()
    B = a.train_batch
    cfg = synth.make_cfg(data_channels=15, dropout=0.1)
    kw = {k: v for k, v in cfg.items() if k != "data_channels"}
    mod = ftb.Geo3DStochInterpCond(data_shape=(S, S, S), embedding_dim=15, **kw).to(dev)
    mod.net.load_state_dict(synth.synth_unet3d_cond_params(cfg, 5))
    tr = ftb.CondFlowTrainer(mod, lr=1e-3, max_grad_norm=0.3, ema_decay=0.9995)
    cats = torch.randint(-1, 14, (B, 1, S, S, S), generator=torch.Generator().manual_seed(7)).to(dev)
    ms = timed(lambda: tr.step(cats), 3)
    out["cond_train"] = {"ms_per_step": ms, "voxels_per_s": B * S ** 3 / (ms * 1e-3), "batch": B,
                         "what": "CondFlowTrainer.step: conditioning front-end kernel, Unet3DCond v3 fwd/bwd, "
                                 "flow + reconstruction loss, clip 0.3 + AdamW, EMA (1 GPU)"}
    del tr, mod
    torch.cuda.empty_cache()

    # ---- north_star's target line names 100 EULER steps: the same sampler, one evaluation per step
    cfg = synth.make_cfg()
    net = ftb.Unet3D(**cfg).to(dev).eval()
    net.load_state_dict(synth.synth_unet3d_params(cfg, 0))
    xb = torch.randn(a.batch, 18, S, S, S, generator=torch.Generator().manual_seed(100)).to(dev)
    h = (TF - T0) / N_ODE_STEPS
    eul = ftb.ODEFlowSolver(net, method="euler")
    with torch.no_grad():
        k = 10
        ms = timed(lambda: eul.solve(xb, t0=T0, tf=T0 + k * h, n_steps=k + 1, return_trajectory=False), 2) / k
    out["euler"] = {"ms_per_step": ms, "samples_per_s": a.batch / (N_ODE_STEPS * ms * 1e-3), "batch": a.batch,
                    "tflops": GF_PER_EVAL * a.batch / ms, "what": "configs[1] with the Euler integrator (north_star target: "
                    "100 Euler steps): one velocity evaluation + one fused axpy per step"}
    # ---- configs[4]: one-sided denoising SDE at 128^3 (eps = 0.1, fresh noise per evaluation, fixed-grid Heun)
    del xb
    torch.cuda.empty_cache()
    x128 = torch.randn(1, 18, 128, 128, 128, generator=torch.Generator().manual_seed(101)).to(dev)
    sde = ftb.SDEOneSidedDenoisingSolver(net, ftb.LinearInterpolant(one_sided=True), epsilon=torch.tensor(0.1), method="heun")
    with torch.no_grad():
        k = 5
        ms = timed(lambda: sde.solve(x128, t0=0.05, tf=0.05 + k * 0.009, n_steps=k + 1, return_trajectory=False), 2) / k
    out["sde128"] = {"ms_per_heun_step": ms, "ms_per_eval": ms / 2, "batch": 1, "tflops": GF_PER_EVAL * 8 * 2 / ms,
                     "samples_per_s_100_steps": 1.0 / (100 * ms * 1e-3),
                     "what": "configs[4]: SDEOneSidedDenoisingSolver, 128^3, B=1, eps=0.1, Heun (2 evaluations + drift/noise "
                             "kernels + randn per step)"}
    del x128, sde, eul, net
    torch.cuda.empty_cache()
    # ---- the GPU bar to beat: the reference's own ATen / cuDNN op sequence (the oracle's functional
    # restatement of Unet3D.forward, eager PyTorch) on this same B200 -- a reported baseline, never the product path
    from oracle import unet3d as oracle_unet3d
    params = {k2: v.to(dev) for k2, v in synth.synth_unet3d_params(cfg, 0).items()}
    xg = synth.synth_input((a.batch, 18, S, S, S), 100).to(dev)
    tg = torch.full((a.batch,), 0.5, device=dev)
    ref = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    for name, tf32, n in (("tf32_on_default", True, 3), ("tf32_off_true_fp32", False, 1)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad():
            ms = timed(lambda: oracle_unet3d.unet3d_forward(params, cfg, xg, tg), n)
        ref[name] = {"ms_per_eval": ms, "samples_per_s_heun100": a.batch / (2 * N_ODE_STEPS * ms * 1e-3),
                     "samples_per_s_euler100": a.batch / (N_ODE_STEPS * ms * 1e-3)}
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    ref["what"] = (f"eager PyTorch (cuDNN / ATen) running the reference's op sequence on the same GPU, B={a.batch} {S}^3, one "
                   "velocity evaluation; tf32_on_default is what the unmodified reference gets on a GPU")
    out["gpu_reference"] = ref
    del params, xg
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------- B200 arm
def run_b200(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    import flowtrain_stochastic_interpolation_b200 as ftb
    from flowtrain_stochastic_interpolation_b200 import _lib
    from oracle import synth  # synthetic weights/inputs only (no oracle compute on this arm's timed path)

    if a.train_only:
        tl = run_train_leg(a, ftb, _lib, dev, rank, world, dist)
        if rank == 0:
            emit(tl)
        if world > 1:
            dist.destroy_process_group()
        return
    cfg = synth.make_cfg()
    net = ftb.Unet3D(**cfg).to(dev).eval()
    net.load_state_dict(synth.synth_unet3d_params(cfg, 0))
    B, S = a.batch, a.size
    per = EVALS_PER_STEP[a.method]
    W = ftb.simplex_embedding(15, 18).to(dev)
    # per-rank noise: sample index i -> seed 100 + i (independent samples, no communication)
    g = torch.Generator("cpu").manual_seed(100 + rank)
    x_host = torch.randn(B, 18, S, S, S, generator=g).pin_memory()
    x_dev = x_host.to(dev)
    solver = ftb.ODEFlowSolver(net, method=a.method)
    h = (TF - T0) / N_ODE_STEPS

    def run_steps(k, x):
        return solver.solve(x, t0=T0, tf=T0 + k * h, n_steps=k + 1, return_trajectory=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
            clocks.wait_ready()
        run_steps(max(a.warmup, 1), x_dev)
        barrier()
        l0 = _lib.lib.ftb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        clocks.mark_begin()
        e0.record()
        x_end = run_steps(a.steps, x_dev)
        e1.record()
        barrier()
        clocks.mark_end()
        step_ms = e0.elapsed_time(e1) / a.steps
        launches = _lib.lib.ftb_launch_count() - l0
        clk = clocks.stop() if rank == 0 else None
        # decode of the final state (part of the metric)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ftb.decode(W, x_end)
        torch.cuda.synchronize()
        d0.record()
        dec = ftb.decode(W, x_end)
        d1.record()
        torch.cuda.synchronize()
        decode_ms = d0.elapsed_time(d1)

        # ---- roofline leg: per-launch CUDA events around every conv_igemm launch of one more step
        _lib.lib.ftb_profile_enable(1)
        run_steps(1, x_dev)
        torch.cuda.synchronize()
        import ctypes as C
        nk = 2
        fl, by, ms = (C.c_double * nk)(), (C.c_double * nk)(), (C.c_double * nk)()
        ln = (C.c_int * nk)()
        _lib.check(_lib.lib.ftb_profile_collect(fl, by, ms, ln, nk))
        _lib.lib.ftb_profile_enable(0)
        t1, t2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t1.record()
        run_steps(1, x_dev)
        t2.record()
        torch.cuda.synchronize()
        one_step_ms = t1.elapsed_time(t2)

        # ---- e2e: full solve through the public API from pinned host noise to host categories
        e2e = None
        if not a.no_e2e:
            barrier()
            w0 = time.perf_counter()
            xd = x_host.to(dev, non_blocking=True)
            xe = solver.solve(xd, t0=T0, tf=TF, n_steps=N_ODE_STEPS + 1, return_trajectory=False)
            out_host = ftb.decode(W, xe).cpu()
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - w0
            e2e = (e2e_s, x_host.numel() * 4, out_host.numel() * 8)

    train_leg = None
    if not a.no_train:
        del solver, net
        torch.cuda.empty_cache()
        train_leg = run_train_leg(a, ftb, _lib, dev, rank, world, dist)

    ensemble = None
    if not a.no_extras and not os.environ.get("FTB_BENCH_MINIMAL"):
        try:
            if train_leg is None:
                del solver, net
                torch.cuda.empty_cache()
            ensemble = run_ensemble_leg(a, ftb, dev, rank, world, dist)
        except Exception as exc:   # side measurements never break the headline line
            ensemble = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    extras = None
    if not a.no_extras and world == 1 and not os.environ.get("FTB_BENCH_MINIMAL"):
        try:
            extras = run_extras(a, ftb, dev)
        except Exception as exc:   # side measurements never break the headline line
            extras = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    tot_s = (N_ODE_STEPS * step_ms + decode_ms) / 1e3
    stats = torch.tensor([step_ms, decode_ms, tot_s, e2e[0] if e2e else 0.0, float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    step_ms, decode_ms, tot_s, e2e_s, launches = stats.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    conv_tf = fl[0] / (ms[0] * 1e-3) / 1e12 if ms[0] > 0 else 0.0
    traffic = ncu_traffic()
    roof = {
        "bound": "tensor", "kernel": "conv_igemm_kernel (k>=3 convs)", "achieved": conv_tf, "peak": peaks["bf16"],
        "unit": "TFLOP/s", "frac": conv_tf / peaks["bf16"], "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
        "traffic_source": traffic if traffic else None,
        "peak_source": peaks["src"], "launches_per_step": int(ln[0]), "avg_launch_ms": ms[0] / max(ln[0], 1),
        "algorithmic_gflop_per_launch": fl[0] / max(ln[0], 1) / 1e9,
        "share_of_step": ms[0] / one_step_ms if one_step_ms > 0 else None,
        "conv1x1": {"ms": ms[1], "launches": int(ln[1]), "achieved_gbs": by[1] / (ms[1] * 1e-3) / 1e9 if ms[1] > 0 else None,
                    "hbm_peak_gbs": peaks["hbm_gbs"]},
        "whole_step_tflops": GF_PER_EVAL * per * B / step_ms,
        "whole_step_frac": GF_PER_EVAL * per * B / step_ms / peaks["bf16"],
    }
    line = {
        "metric": METRIC, "value": world * B / tot_s, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": world * B, "volume": [S, S, S],
                   "ode_steps": N_ODE_STEPS, "evals_per_step": per, "decode_ms": decode_ms,
                   "step": "one integrator step of the whole batch; value = n_gpus*B/(100*step+decode)",
                   "l2": "inputs larger than L2 (fp32 state 151 MB, activations 201 MB each at B=8); no flush",
                   "parallelism": f"independent samples, {world} rank(s), no data-path collective",
                   "weights": "synthetic random-init (oracle/synth.py seed 0)"},
        "clocks": clk, "gpu_launches": int(launches), "roofline": roof,
    }
    if train_leg:
        line["train"] = train_leg
    if ensemble:
        line["ensemble"] = ensemble
    if extras:
        line["extras"] = extras
    if e2e:
        line["e2e"] = {"value": world * B / e2e_s, "unit": UNIT, "h2d_bytes_per_step": e2e[1], "d2h_bytes_per_step": e2e[2],
                       "seconds_per_solve": e2e_s,
                       "what": "ODEFlowSolver.solve (full 100 steps) + decode, pinned-host X0 in, host int64 volume out; "
                               "bytes are per solve"}
    if not a.no_cpu_baseline:
        s_per_eval, cores = cpu_velocity_evals(a.cpu_evals, S)
        line["cpu_baseline"] = {
            "value": 1.0 / (N_ODE_STEPS * per * s_per_eval), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{a.cpu_evals} velocity evaluations at B=1 {S}^3 fp32 on the host ({s_per_eval:.2f} s each, oracle "
                      f"port of the reference CPU path); samples/s = 1/({N_ODE_STEPS}*{per}*s_per_eval)"}
    # the numbers of every leg once more, compact, as the LAST key (log tails keep the end of the line)
    line["summary"] = {
        "n_gpus": world, "sampling_samples_per_s": line["value"], "sampling_e2e_samples_per_s": line.get("e2e", {}).get("value"),
        "sampling_ms_per_heun_step": step_ms, "whole_step_frac_of_bf16_peak": roof["whole_step_frac"],
        "conv_kernel_frac_of_bf16_peak": roof["frac"],
        "train_voxels_per_s": train_leg.get("value") if train_leg else None,
        "train_ms_per_step": train_leg.get("ms_per_step") if train_leg else None,
        "ensemble64_cond_samples_per_s": ensemble.get("value") if ensemble else None,
        "ensemble64_total_ms": ensemble.get("total_ms") if ensemble else None,
        "euler_samples_per_s": (extras or {}).get("euler", {}).get("samples_per_s"),
        "sde128_ms_per_heun_step": (extras or {}).get("sde128", {}).get("ms_per_heun_step"),
        "gpu_reference_tf32_ms_per_eval": (extras or {}).get("gpu_reference", {}).get("tf32_on_default", {}).get("ms_per_eval"),
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    a = parse()
    # stdout carries exactly one JSON line: anything a library prints to file descriptor 1 (e.g. NCCL's version
    # banner under NCCL_DEBUG=VERSION) is sent to stderr instead
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
