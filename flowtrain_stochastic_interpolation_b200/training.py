"""Training step on the B200 path.

Two entry points, both over the same C-ABI calls (``ftb_unet3d_forward_train`` / ``ftb_unet3d_backward``):

* drop-in autograd: a ``Unet3D`` in ``train()`` mode called with grad enabled behaves like the reference
  module inside ``Geo3DStochInterp.training_step``
  (project/geodata-3d-unconditional/model_train_inference.py:417-457): ``loss.backward()`` fills ``p.grad``
  of every parameter, any torch optimiser / Lightning loop works unchanged.
* ``FlowTrainer``: the whole step of the reference (embed -> noise -> interpolant -> net -> loss -> backward
  -> clip_grad_norm_ -> Adam -> EMA, :417-473 and callbacks.py:238-268) as kernels on flat fp32 buffers,
  with the data-parallel gradient all-reduce (NCCL through ``torch.distributed``) started bucket by bucket
  from inside the backward, as DDP does.

There is no PyTorch implementation of the math here; without libftb.so / a GPU everything raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib


# --------------------------------------------------------------------------- flat parameter storage
def flatten_parameters(net) -> torch.Tensor:
    """Move every parameter of ``net`` (a B200 ``Unet3D``) into ONE flat fp32 device buffer in state_dict
    order and bind the engine to it (no per-step copies).  The ``nn.Parameter`` objects stay the same
    (optimisers, EMA and checkpoints keep working); their storage becomes a view of the flat buffer."""
    params = list(net.named_parameters())
    dev = params[0][1].device
    if dev.type != "cuda":
        raise RuntimeError("flatten_parameters: move the module to a CUDA device first (no CPU path)")
    h = net._handle
    n = _lib.lib.ftb_unet3d_num_params(h)
    assert n == len(params)
    total = _lib.lib.ftb_unet3d_param_offset(h, n)
    flat = torch.empty(total, dtype=torch.float32, device=dev)
    offs = []
    with torch.no_grad():
        for i, (name, p) in enumerate(params):
            off = _lib.lib.ftb_unet3d_param_offset(h, i)
            view = flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.detach().float())
            p.data = view
            offs.append(off)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.ftb_unet3d_bind_params(h, _lib.ptr(flat), _lib.stream_ptr()))
    net._flat = flat
    net._flat_offsets = offs
    net._flat_versions = None
    net._synced.clear()
    net._graphs.clear()
    return flat


def _flat_still_bound(net) -> bool:
    flat = net._flat
    base = flat.data_ptr()
    for (name, p), off in zip(net.named_parameters(), net._flat_offsets):
        if p.data_ptr() != base + 4 * off or p.dtype != torch.float32:
            return False
    return True


def sync_flat(net):
    """Called before a forward when the parameters are flat-bound: a changed version counter (in-place
    optimiser step, load_state_dict) only needs the packed copies rebuilt."""
    if not _flat_still_bound(net):
        flatten_parameters(net)   # someone re-assigned .data (e.g. .to()): re-flatten
    vers = tuple(p._version for p in net.parameters())
    if vers != net._flat_versions:
        _lib.check(_lib.lib.ftb_unet3d_mark_dirty(net._handle))
        net._flat_versions = vers
        net._graphs.clear()


def train_workspace(net, device, B, X, Y, Z):
    key = ("train", str(device), B, X, Y, Z)
    ws = net._workspace.get(key)
    if ws is None:
        nbytes = _lib.lib.ftb_unet3d_train_workspace_bytes(net._handle, B, X, Y, Z)
        if nbytes == 0:
            raise _lib.FtbError(_lib.last_error() or "training workspace sizing failed")
        net._workspace.clear()
        net._graphs.clear()
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        net._workspace[key] = ws
    base = (ws.data_ptr() + 255) // 256 * 256
    return ws, base, ws.numel() - (base - ws.data_ptr())


def forward_train(net, xin, tin, ain=None):
    """Train-mode forward: keeps the intermediates in the workspace and records the backward.
    ``ain``: the conditioning volume ATb [B,C,X,Y,Z] of a ``Unet3DCond``."""
    B, _, X, Y, Z = xin.shape
    if getattr(net, "_flat", None) is None:
        flatten_parameters(net)
    sync_flat(net)
    ws, base, nbytes = train_workspace(net, xin.device, B, X, Y, Z)
    out = torch.empty_like(xin)
    # nn.Dropout(p) of every ResnetBlock.block1 (unet_attn_3d.py:244, :261): fresh counter-based mask per forward
    p_drop = net.dropout_p if net.training else 0.0
    if net._drop_seed is not None:
        seed = net._drop_seed
    else:
        net._drop_count += 1
        seed = (torch.initial_seed() * 0x9E3779B1 + net._drop_count) & 0xFFFFFFFFFFFFFFFF
    _lib.check(_lib.lib.ftb_unet3d_set_dropout(net._handle, float(p_drop), seed))
    if ain is not None:
        _lib.check(_lib.lib.ftb_unet3d_cond_forward_train(net._handle, _lib.ptr(xin), _lib.ptr(ain), _lib.ptr(tin),
                                                          _lib.ptr(out), B, X, Y, Z, C.c_void_p(base), nbytes,
                                                          _lib.stream_ptr()))
    else:
        _lib.check(_lib.lib.ftb_unet3d_forward_train(net._handle, _lib.ptr(xin), _lib.ptr(tin), _lib.ptr(out), B, X, Y,
                                                     Z, C.c_void_p(base), nbytes, _lib.stream_ptr()))
    return out


def backward_into(net, dout, gflat, bucket_cb=None):
    """Adds the parameter gradients of the last ``forward_train`` into ``gflat`` (flat, state_dict order)."""
    B, _, X, Y, Z = dout.shape
    ws, base, nbytes = train_workspace(net, dout.device, B, X, Y, Z)
    cb = _lib.BUCKET_CB(bucket_cb) if bucket_cb is not None else C.cast(None, _lib.BUCKET_CB)
    _lib.check(_lib.lib.ftb_unet3d_backward(net._handle, _lib.ptr(dout), _lib.ptr(gflat), C.c_void_p(base), nbytes,
                                            cb, None, _lib.stream_ptr()))


class UnetTrainFn(torch.autograd.Function):
    """autograd bridge: forward = ftb_unet3d_(cond_)forward_train, backward = ftb_unet3d_backward.  The gradient
    w.r.t. the network INPUTS (x, ATb) is not computed: the training steps never need it, XT and ATb are data
    (the conditional project freezes its embedding, model_train_sh_inference_cond.py:302)."""

    @staticmethod
    def forward(ctx, net, x, t, atb, *params):
        ctx.net = net
        with torch.cuda.device(x.device):
            out = forward_train(net, x, t, atb)
        return out

    @staticmethod
    def backward(ctx, dout):
        net = ctx.net
        dout = dout.detach().float().contiguous()
        gflat = torch.zeros_like(net._flat)
        with torch.cuda.device(dout.device):
            backward_into(net, dout, gflat)
        grads = []
        for (name, p), off in zip(net.named_parameters(), net._flat_offsets):
            grads.append(gflat[off:off + p.numel()].view(p.shape) if p.requires_grad else None)
        return (None, None, None, None, *grads)


class _CondLossFn(torch.autograd.Function):
    """Flow + T-weighted masked reconstruction loss of the conditional training_step
    (model_train_sh_inference_cond.py:434-452) and its gradient w.r.t. VT_hat: ftb_cond_loss_accumulate / _grad."""

    @staticmethod
    def forward(ctx, VT, VT_hat, XT, X1c, X1, mask, T, lam):
        f = lambda t: t.detach().float().contiguous()
        VT, vhat, XT, X1c, X1, T = f(VT), f(VT_hat), f(XT), f(X1c), f(X1), f(T)
        m8 = mask.contiguous().view(torch.uint8)
        B, E = X1.shape[0], X1.shape[1]
        n = X1[0, 0].numel()
        acc6 = torch.zeros(6, dtype=torch.float64, device=VT.device)
        with torch.cuda.device(VT.device):
            _lib.check(_lib.lib.ftb_cond_loss_accumulate(_lib.ptr(VT), _lib.ptr(vhat), _lib.ptr(XT), _lib.ptr(X1c),
                                                         _lib.ptr(X1), _lib.ptr(m8), _lib.ptr(T), B, E, n,
                                                         _lib.ptr(acc6), _lib.stream_ptr()))
        ctx.save_for_backward(VT, vhat, XT, X1c, m8, T, acc6)
        ctx.lam = float(lam)
        N = float(VT.numel())
        flow = (acc6[0] / N) / (acc6[1] / N + 1e-6)
        rec = (acc6[5] / B) * (acc6[2] / acc6[3]) / (acc6[4] / N + 1e-6)
        ctx.mark_non_differentiable(flow, rec)
        return (flow + ctx.lam * rec).float(), flow.float(), rec.float()

    @staticmethod
    def backward(ctx, gout, _gf, _gr):
        VT, vhat, XT, X1c, m8, T, acc6 = ctx.saved_tensors
        B, E = X1c.shape[0], X1c.shape[1]
        n = X1c[0, 0].numel()
        dout = torch.empty_like(vhat)
        with torch.cuda.device(VT.device):
            _lib.check(_lib.lib.ftb_cond_loss_grad(_lib.ptr(VT), _lib.ptr(vhat), _lib.ptr(XT), _lib.ptr(X1c),
                                                   _lib.ptr(m8), _lib.ptr(T), B, E, n, _lib.ptr(acc6), ctx.lam, 1.0,
                                                   _lib.ptr(dout), _lib.stream_ptr()))
        return None, dout.mul_(gout.to(dout.dtype)), None, None, None, None, None, None


def cond_loss(VT, VT_hat, XT, X1_clean, X1, mask, T, lambda_reconstruct=1.0):
    """(loss, flow_loss, reconstruct_loss) of the conditional training_step (:434-452); ``loss`` is differentiable
    w.r.t. ``VT_hat``.  ``mask``: the bool [B,1,X,Y,Z] conditioning mask, ``X1_clean`` the embedding before the 1e-4
    noise (``b`` is gathered from it, :417)."""
    if not VT.is_cuda:
        raise RuntimeError("cond_loss runs on CUDA only (no CPU fallback)")
    return _CondLossFn.apply(VT, VT_hat, XT, X1_clean, X1, mask, T, lambda_reconstruct)


# --------------------------------------------------------------------------- fused training step
class BucketAllReduce:
    """Starts ``all_reduce(sum)`` of gradient ranges on a side stream as the backward completes them
    (the engine calls ``__call__(user, offset, count)`` from inside ``ftb_unet3d_backward``), so the
    NCCL traffic overlaps the rest of the backward like DDP's bucketed reducer.  ``min_elems`` coalesces
    adjacent ranges.  Works on any backend (gloo on CPU tensors in the tests)."""

    def __init__(self, gflat: torch.Tensor, group=None, min_elems: int = 1 << 20):
        import torch.distributed as dist
        self.dist = dist
        self.gflat = gflat
        self.group = group
        self.min_elems = min_elems
        self.cuda = gflat.is_cuda
        self.side = torch.cuda.Stream(device=gflat.device) if self.cuda else None
        self.pending = None     # (lo, hi) contiguous range not yet sent
        self.works = []
        self.launched = []      # ranges actually reduced (for tests)

    def _launch(self, lo, hi):
        view = self.gflat[lo:hi]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.gflat.device))
            self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                self.works.append(self.dist.all_reduce(view, group=self.group, async_op=True))
        else:
            self.works.append(self.dist.all_reduce(view, group=self.group, async_op=True))
        self.launched.append((lo, hi))

    def __call__(self, user, offset, count):
        lo, hi = int(offset), int(offset) + int(count)
        if self.pending is not None and self.pending[0] == hi:        # ranges arrive back to front
            self.pending = (lo, self.pending[1])
        elif self.pending is not None and self.pending[1] == lo:
            self.pending = (self.pending[0], hi)
        else:
            if self.pending is not None:
                self._launch(*self.pending)
            self.pending = (lo, hi)
        if self.pending[1] - self.pending[0] >= self.min_elems:
            self._launch(*self.pending)
            self.pending = None

    def finish(self):
        if self.pending is not None:
            self._launch(*self.pending)
            self.pending = None
        for w in self.works:
            w.wait()
        self.works = []
        if self.cuda:
            torch.cuda.current_stream(self.gflat.device).wait_stream(self.side)


class FlowTrainer:
    """One optimiser step of ``Geo3DStochInterp`` (training_step :417-457, configure_optimizers :465-473,
    Lightning's ``gradient_clip_val``, EMACallback.on_train_batch_end callbacks.py:238-268) entirely in
    kernels: no autograd graph, flat fp32 parameter / gradient / Adam-moment / EMA buffers.

    ``module`` is the B200 ``Geo3DStochInterp``.  Data parallelism: one process per GPU, identical initial
    weights (``broadcast_parameters``), per-rank batches; gradients are summed with NCCL bucket by bucket
    while the backward is still running and scaled by 1/world inside the Adam kernel."""

    def __init__(self, module, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False,
                 max_grad_norm: Optional[float] = 1.0, ema_decay: Optional[float] = 0.9995, ema_start_step=0,
                 lr_decay: Optional[float] = None, process_group=None, distributed: Optional[bool] = None):
        import torch.distributed as dist
        self.module = module
        self.net = module.net
        self.net.train()   # Block1 dropout (if the module was built with dropout > 0) is active, as in Lightning's fit
        self.flat = flatten_parameters(self.net)
        self.gflat = torch.zeros_like(self.flat)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.weight_decay, self.decoupled = float(weight_decay), bool(decoupled)
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self.ema_decay, self.ema_start_step = ema_decay, ema_start_step
        self.ema_flat = None
        self.lr_decay = lr_decay
        self.step_count = 0
        self.group = process_group
        self.distributed = dist.is_available() and dist.is_initialized() if distributed is None else distributed
        self.world = dist.get_world_size(process_group) if self.distributed else 1
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.flat.device)
        self.acc = torch.zeros(2, dtype=torch.float64, device=self.flat.device)
        self.last_grad_norm = None
        # parameters the reference keeps frozen (freqs / phases when not learned) get a zero gradient and are left
        # out of the optimiser launches: torch's Adam / AdamW skip parameters without a gradient, so decoupled weight
        # decay must not touch them either
        self.frozen = [(off, p.numel()) for (n, p), off in zip(self.net.named_parameters(), self.net._flat_offsets)
                       if not p.requires_grad]
        self.trainable = []
        for (n, p), off in zip(self.net.named_parameters(), self.net._flat_offsets):
            if not p.requires_grad:
                continue
            if self.trainable and self.trainable[-1][0] + self.trainable[-1][1] == off:
                self.trainable[-1] = (self.trainable[-1][0], self.trainable[-1][1] + p.numel())
            else:
                self.trainable.append((off, p.numel()))

    def broadcast_parameters(self, src=0):
        if self.distributed:
            import torch.distributed as dist
            dist.broadcast(self.flat, src=src, group=self.group)
            _lib.check(_lib.lib.ftb_unet3d_mark_dirty(self.net._handle))

    def step(self, batch, noise1=None, X0=None, T=None):
        """batch: [B,1,X,Y,Z] integer categories.  Returns the (per-rank) loss as a 0-d device tensor."""
        mod, net, lib = self.module, self.net, _lib.lib
        dev = self.flat.device
        with torch.cuda.device(dev), torch.no_grad():
            st = _lib.stream_ptr()
            X1 = mod.embed(batch)                                            # :428
            if noise1 is None:
                noise1 = torch.randn_like(X1)
            _lib.check(lib.ftb_ode_axpy(_lib.ptr(X1), _lib.ptr(X1), _lib.ptr(noise1), 1e-3, X1.numel(), None, 1, st))  # :429
            if X0 is None:
                X0 = torch.randn_like(X1)                                    # :431
            if T is None:
                T = torch.empty(X1.size(0), device=dev).uniform_(mod.time_range[0], mod.time_range[1])  # :434-436
            XT, VT = mod.interpolator.flow_objective(T, X0, X1)              # :439
            XT = XT.contiguous()
            VT = VT.contiguous()
            Tin = T.to(device=dev, dtype=torch.float32).contiguous()
            vhat = forward_train(net, XT, Tin)                               # :440
            self.acc.zero_()
            _lib.check(lib.ftb_mse_ratio_accumulate(_lib.ptr(VT), _lib.ptr(vhat), VT.numel(), _lib.ptr(self.acc), st))  # :443
            dout = torch.empty_like(vhat)
            _lib.check(lib.ftb_mse_ratio_grad(_lib.ptr(VT), _lib.ptr(vhat), VT.numel(), _lib.ptr(self.acc), 1.0,
                                              _lib.ptr(dout), st))
            self._optimizer_tail(dout)
            return (self.acc[0] / self.acc[1]).float()

    def _optimizer_tail(self, dout):
        """backward -> (bucketed all-reduce) -> clip + Adam(W) -> EMA, shared by both training steps."""
        net, lib, st = self.net, _lib.lib, _lib.stream_ptr()
        self.gflat.zero_()
        reducer = BucketAllReduce(self.gflat, self.group) if self.world > 1 else None
        backward_into(net, dout, self.gflat, reducer)
        if reducer is not None:
            reducer.finish()
        for off, n in self.frozen:
            self.gflat[off:off + n].zero_()
        self.step_count += 1
        self.sumsq.zero_()
        _lib.check(lib.ftb_grad_sumsq(_lib.ptr(self.gflat), self.gflat.numel(), _lib.ptr(self.sumsq), st))
        for off, cnt in self.trainable:     # one launch when nothing is frozen (contiguous ranges are merged)
            _lib.check(lib.ftb_adam_step(_lib.ptr(self.flat[off:]), _lib.ptr(self.gflat[off:]), _lib.ptr(self.m[off:]),
                                         _lib.ptr(self.v[off:]), cnt, self.lr, self.betas[0], self.betas[1], self.eps,
                                         self.weight_decay, 1 if self.decoupled else 0, self.step_count,
                                         _lib.ptr(self.sumsq), 1.0 / self.world, self.max_grad_norm, st))
        _lib.check(lib.ftb_unet3d_mark_dirty(net._handle))
        net._flat_versions = None
        if self.ema_decay is not None and self.step_count >= self.ema_start_step:   # callbacks.py:243-266
            if self.ema_flat is None:
                self.ema_flat = self.flat.clone()
            else:
                _lib.check(lib.ftb_ema_update(_lib.ptr(self.ema_flat), _lib.ptr(self.flat), self.flat.numel(),
                                              float(self.ema_decay), st))
        self.last_grad_norm = self.sumsq   # squared norm of the summed gradient (device)

    def epoch_end(self):
        """ExponentialLR(gamma=lr_decay) per epoch (:465-473)."""
        if self.lr_decay:
            self.lr *= self.lr_decay

    def ema_state_dict(self):
        """EMA shadow under the parameter names (checkpoint['ema_shadow'], callbacks.py:295-317)."""
        if self.ema_flat is None:
            return {}
        return {n: self.ema_flat[off:off + p.numel()].view(p.shape).clone()
                for (n, p), off in zip(self.net.named_parameters(), self.net._flat_offsets) if p.requires_grad}


class CondFlowTrainer(FlowTrainer):
    """One optimiser step of the conditional project (training_step
    project/geodata-3d-conditional/model_train_sh_inference_cond.py:401-467; AdamW lr 1e-3 :487-495, Lightning
    ``gradient_clip_val`` 0.3, EMA 0.9995 every batch) in kernels: fused conditioning front-end (embed + surface /
    borehole mask + ATb), interpolant, ``Unet3DCond`` forward/backward, the flow + T-weighted reconstruction loss
    and its gradient without a host sync, clip + AdamW, EMA.  ``module`` is the B200 ``Geo3DStochInterpCond``."""

    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, decoupled=True,
                 max_grad_norm: Optional[float] = 0.3, ema_decay: Optional[float] = 0.9995, ema_start_step=0,
                 lr_decay: Optional[float] = None, process_group=None, distributed: Optional[bool] = None,
                 generator: Optional[torch.Generator] = None):
        super().__init__(module, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled,
                         max_grad_norm=max_grad_norm, ema_decay=ema_decay, ema_start_step=ema_start_step,
                         lr_decay=lr_decay, process_group=process_group, distributed=distributed)
        self.acc6 = torch.zeros(6, dtype=torch.float64, device=self.flat.device)
        self.generator = generator   # CPU generator of the borehole draw
        self.last_terms = None

    def step(self, batch, noise1=None, X0=None, T=None, bores=None, n_bores=None):
        """batch: [B,1,X,Y,Z] integer categories.  Returns the (per-rank) loss as a 0-d device tensor; the random
        draws (noise1 = randn for X1, X0, T, borehole columns) can be passed in for parity runs."""
        mod, net, lib = self.module, self.net, _lib.lib
        dev = self.flat.device
        with torch.cuda.device(dev), torch.no_grad():
            st = _lib.stream_ptr()
            X1c, ATb, mask = mod.conditioning(batch, bores, n_bores, self.generator)       # :413-420
            if noise1 is None:
                noise1 = torch.randn_like(X1c)
            X1 = torch.empty_like(X1c)
            _lib.check(lib.ftb_ode_axpy(_lib.ptr(X1), _lib.ptr(X1c), _lib.ptr(noise1), 1e-4, X1.numel(), None, 1, st))  # :421
            if X0 is None:
                X0 = torch.randn_like(X1)                                                  # :423
            if T is None:
                T = torch.empty(X1.size(0), device=dev).uniform_(mod.time_range[0], mod.time_range[1])   # :426-428
            XT, VT = mod.interpolator.flow_objective(T, X0, X1)                            # :431
            XT, VT = XT.contiguous(), VT.contiguous()
            Tin = T.to(device=dev, dtype=torch.float32).contiguous()
            vhat = forward_train(net, XT, Tin, ATb)                                        # :432
            B, E = X1.shape[0], X1.shape[1]
            n = X1[0, 0].numel()
            m8 = mask.view(torch.uint8)
            self.acc6.zero_()
            _lib.check(lib.ftb_cond_loss_accumulate(_lib.ptr(VT), _lib.ptr(vhat), _lib.ptr(XT), _lib.ptr(X1c),
                                                    _lib.ptr(X1), _lib.ptr(m8), _lib.ptr(Tin), B, E, n,
                                                    _lib.ptr(self.acc6), st))              # :434-451
            dout = torch.empty_like(vhat)
            _lib.check(lib.ftb_cond_loss_grad(_lib.ptr(VT), _lib.ptr(vhat), _lib.ptr(XT), _lib.ptr(X1c), _lib.ptr(m8),
                                              _lib.ptr(Tin), B, E, n, _lib.ptr(self.acc6),
                                              float(mod.lambda_reconstruct), 1.0, _lib.ptr(dout), st))
            self._optimizer_tail(dout)
            a = self.acc6
            N = float(VT.numel())
            flow = (a[0] / N) / (a[1] / N + 1e-6)
            rec = (a[5] / B) * (a[2] / a[3]) / (a[4] / N + 1e-6)
            self.last_terms = (flow.float(), rec.float())          # "flow_loss", "reconstruct_loss" of log_dict :454-465
            return (flow + mod.lambda_reconstruct * rec).float()
