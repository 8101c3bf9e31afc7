"""Streaming synthetic-data stand-in for the reference's training input.

The reference trains on ``geogen.dataset.GeoData3DStreamingDataset(model_resolution, model_bounds, dataset_size,
device)`` wrapped in a ``DataLoader(batch_size, shuffle=True, num_workers=16)``
(project/geodata-3d-unconditional/model_train_inference.py:249-260): every item is a freshly generated synthetic
geological model, an integer category volume ``[1, X, Y, Z]`` with values in ``-1 .. 13`` (-1 = air).  ``geogen`` is
an un-vendored dependency that is not installed here, so its generator cannot be restated; what the hot path needs
from it is the CONTRACT (shape, dtype, value range, an endless deterministic-per-index stream) and a way to get the
batches onto the GPU without stalling the training step.  This module provides both:

* ``SyntheticGeoStreamingDataset``: same constructor keywords and item contract; items are layered volumes (tilted,
  folded strata under a topographic surface, cut by a dike and a fault) generated on the host from the item index, so
  a run is reproducible and workers need no shared state.  Not a geological simulator, and not numerically comparable
  with GeoGen (stated in DESIGN.md): it stands in for the data SHAPE only.
* ``DevicePrefetcher``: wraps any iterable of host batches; copies batch k+1 to the GPU from pinned memory on a side
  stream while step k computes (double-buffered), so ``FlowTrainer.step`` never waits for a host->device copy.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset

N_ROCK = 14        # categories 0 .. 13; -1 is air (embed() maps cat + 1 -> row, model_train_inference.py:361-370)


def synthetic_geomodel(index: int, resolution: Sequence[int], bounds=None, seed: int = 0) -> np.ndarray:
    """One int64 category volume [X, Y, Z] in -1 .. 13, a pure function of (seed, index)."""
    X, Y, Z = (int(v) for v in resolution)
    if bounds is None:
        bounds = ((-1.0, 1.0),) * 3
    rng = np.random.default_rng(np.random.SeedSequence([int(seed), int(index)]))
    ax = [np.linspace(lo, hi, n, dtype=np.float32) for (lo, hi), n in zip(bounds, (X, Y, Z))]
    x, y, z = np.meshgrid(*ax, indexing="ij", sparse=True)
    span = [hi - lo for lo, hi in bounds]
    # stratigraphic coordinate: tilted planes + a gentle fold
    dip = rng.normal(0.0, 0.35, size=2).astype(np.float32)
    fold_amp = np.float32(rng.uniform(0.0, 0.15) * span[2])
    fold_k = np.float32(rng.uniform(1.0, 3.0) * 2 * np.pi / span[0])
    fold_dir = np.float32(rng.uniform(0, 2 * np.pi))
    u = x * np.cos(fold_dir) + y * np.sin(fold_dir)
    s = z - dip[0] * x - dip[1] * y - fold_amp * np.sin(fold_k * u)
    # a fault: everything on one side of a steep plane is shifted along the stratigraphic coordinate
    fn = rng.normal(size=3).astype(np.float32)
    fn[2] *= 0.3
    fn /= np.linalg.norm(fn)
    side = (fn[0] * x + fn[1] * y + fn[2] * z - np.float32(rng.uniform(-0.3, 0.3))) > 0
    s = s + side * np.float32(rng.uniform(-0.25, 0.25) * span[2])
    # layer boundaries: sorted random thicknesses over the stratigraphic range
    n_layers = int(rng.integers(4, 11))
    cuts = np.sort(rng.uniform(s.min(), s.max(), size=n_layers - 1).astype(np.float32))
    layer_cat = rng.permutation(N_ROCK - 1)[:n_layers]            # category 13 is kept for the dike
    vol = layer_cat[np.searchsorted(cuts, s)]
    # a dike: a thin slab of category 13 cutting the strata
    dn = rng.normal(size=3).astype(np.float32)
    dn[2] *= 0.2
    dn /= np.linalg.norm(dn)
    dist = dn[0] * x + dn[1] * y + dn[2] * z - np.float32(rng.uniform(-0.5, 0.5))
    vol = np.where(np.abs(dist) < np.float32(rng.uniform(0.02, 0.06)), N_ROCK - 1, vol)
    # topography: air (-1) above a smooth surface
    topo = (bounds[2][1] - 0.15 * span[2] * rng.uniform(0.2, 1.0)
            + 0.08 * span[2] * np.sin(np.float32(rng.uniform(1, 3)) * x + np.float32(rng.uniform(0, 6)))
            * np.cos(np.float32(rng.uniform(1, 3)) * y + np.float32(rng.uniform(0, 6))))
    vol = np.where(z > topo, -1, vol)
    return np.ascontiguousarray(np.broadcast_to(vol, (X, Y, Z))).astype(np.int64)


class SyntheticGeoStreamingDataset(Dataset):
    """Drop-in for the constructor / item contract of ``GeoData3DStreamingDataset`` (:249-254): ``len`` is
    ``dataset_size`` (the reference's "epoch size"), item ``i`` is an int64 tensor ``[1, X, Y, Z]`` with categories in
    ``-1 .. 13`` on ``device`` ("cpu" in the reference; workers generate on the host).  ``model_resolution`` may be
    ``[X, Y, Z]`` or the reference config's ``[C, X, Y, Z]`` (:243, leading channel entry ignored).  Each epoch
    re-seeds through ``set_epoch`` so the stream does not repeat (the reference generates fresh models forever)."""

    def __init__(self, model_resolution=(64, 64, 64), model_bounds=None, dataset_size: int = 1_000_000,
                 device: str = "cpu", seed: int = 0):
        res = tuple(int(v) for v in model_resolution)
        if len(res) == 4:
            res = res[1:]
        if len(res) != 3 or min(res) < 1:
            raise ValueError(f"model_resolution must be [X, Y, Z] or [C, X, Y, Z], got {model_resolution}")
        self.model_resolution = res
        self.model_bounds = tuple(tuple(float(v) for v in b) for b in model_bounds) if model_bounds is not None else None
        self.dataset_size = int(dataset_size)
        self.device = torch.device(device)
        self.seed = int(seed)
        self.epoch = 0

    def set_epoch(self, epoch: int):
        self.epoch = int(epoch)

    def __len__(self):
        return self.dataset_size

    def __getitem__(self, index: int) -> torch.Tensor:
        if not 0 <= index < self.dataset_size:
            raise IndexError(index)
        vol = synthetic_geomodel(self.epoch * self.dataset_size + index, self.model_resolution, self.model_bounds, self.seed)
        t = torch.from_numpy(vol).unsqueeze(0)
        return t if self.device.type == "cpu" else t.to(self.device)


class DevicePrefetcher:
    """Iterates device batches from an iterable of host batches, one batch ahead: batch k+1 is staged in pinned memory
    and copied on a side stream while the consumer works on batch k; ``__next__`` only makes the compute stream wait
    for the copy event (no host synchronisation).  Two pinned staging buffers and two device buffers are reused."""

    def __init__(self, batches: Iterable[torch.Tensor], device, depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher feeds the CUDA training step (no CPU path)")
        self.batches = batches
        self.depth = max(2, int(depth))
        self.stream = torch.cuda.Stream(device=self.device)

    def __iter__(self) -> Iterator[torch.Tensor]:
        it = iter(self.batches)
        pinned: list = [None] * self.depth
        devbuf: list = [None] * self.depth
        ready: list = [None] * self.depth      # copy-done events
        freed: list = [None] * self.depth      # consumer-done events (the device buffer may be overwritten)

        def stage(slot: int) -> bool:
            try:
                hb = next(it)
            except StopIteration:
                return False
            hb = hb if isinstance(hb, torch.Tensor) else hb[0]
            if pinned[slot] is None or pinned[slot].shape != hb.shape or pinned[slot].dtype != hb.dtype:
                pinned[slot] = torch.empty(hb.shape, dtype=hb.dtype).pin_memory()
                devbuf[slot] = torch.empty(hb.shape, dtype=hb.dtype, device=self.device)
            if ready[slot] is not None:
                ready[slot].synchronize()      # the previous H2D from this pinned buffer has finished
            pinned[slot].copy_(hb)
            with torch.cuda.stream(self.stream):
                if freed[slot] is not None:
                    self.stream.wait_event(freed[slot])
                devbuf[slot].copy_(pinned[slot], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                ready[slot] = ev
            return True

        from collections import deque
        queue = deque(s for s in range(self.depth - 1) if stage(s))    # staged, not yet consumed
        free = self.depth - 1                                          # the slot the next stage() may use
        while queue:
            slot = queue.popleft()
            if stage(free):                                            # overlap: the next copy flies during this step
                queue.append(free)
            free = slot
            torch.cuda.current_stream(self.device).wait_event(ready[slot])
            yield devbuf[slot]
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            freed[slot] = ev


def get_data_loader(config: dict, device: Optional[str] = None, num_workers: int = 0):
    """Counterpart of the reference ``get_data_loader`` (:240-260) on the stand-in dataset: same config keys
    (``config["data"]["shape" | "bounds" | "epoch_size" | "batch_size"]``), shuffled ``DataLoader``; pass the result to
    ``DevicePrefetcher`` to stream it onto the GPU."""
    from torch.utils.data import DataLoader
    data = config["data"]
    ds = SyntheticGeoStreamingDataset(model_resolution=data["shape"], model_bounds=data.get("bounds"),
                                      dataset_size=data["epoch_size"], device="cpu")
    return DataLoader(ds, batch_size=data["batch_size"], shuffle=True, num_workers=num_workers,
                      pin_memory=False, drop_last=False)


__all__ = ["SyntheticGeoStreamingDataset", "DevicePrefetcher", "get_data_loader", "synthetic_geomodel"]
