"""Samplers with flowtrain's solver API (src/flowtrain/solvers/solvers.py) on fused CUDA updates.

``Solver(model, ...).solve(X0, t0, tf, n_steps) -> [n_steps, B, C, X, Y, Z]`` keeps the reference
meaning of every argument: the time grid is ``linspace(t0, tf, n_steps)`` (n_steps POINTS, :59), the
model is called as ``model(XT, T)`` with ``T = full((B,), t_k)`` fp32 (:68-70), ``frozen_mask``
zeroes the velocity on the masked trailing dims (:71-73), and the eq-6.7 drift / SDE term follow
:138-143 / :205-216.

The reference hands ``ode_func`` to torchdiffeq's ADAPTIVE dopri5 (ODE solvers, :77, :148) /
adaptive_heun (SDE solver, :220-222) with ``atol`` / ``rtol``, and ``n_steps`` is the number of OUTPUT
points.  Called with the reference signature — ``ODEFlowSolver(model, rtol=1e-6).solve(X0, t0=t0, tf=tf,
n_steps=16)``, model_train_inference.py:615-619 — the classes here do the same: ``method=None`` selects
the adaptive stepper the reference uses (``integrate_adaptive``, a restatement of torchdiffeq's loop;
the step controller lives on the device, see ``ode_adaptive.cu``).  The fixed-grid integrators are an
explicit opt-in, ``method`` in {"euler", "heun", "rk4"} (torchdiffeq's fixed-grid convention: one step
per grid interval, so ``n_steps`` points = ``n_steps - 1`` steps); ``atol`` / ``rtol`` have no meaning
there and passing non-default values with a fixed-grid method raises.  The stage combinations run as
single fused kernels (ftb_ode_axpy / heun_combine / rk4_combine / ftb_ode_lincomb) on the fp32 state,
and the fixed-grid loop has no host synchronisation (the reference does one ``t.item()`` device->host
sync per evaluation).  ``return_trajectory=False`` keeps only the end state (the 16-point trajectory at
64^3 is 1.2 GB per sample batch of 4).
"""
from __future__ import annotations

import ctypes as C
import warnings

import torch

from . import _lib
from .interpolation import BaseInterpolant

METHODS = ("euler", "heun", "rk4")
ADAPTIVE_METHODS = ("dopri5", "adaptive_heun")
_DEFAULT_TOL = 1e-6   # reference defaults (:35, :105, :170)


def _resolve_method(method, default, atol, rtol, model):
    """``method=None`` -> the reference's adaptive stepper.  Fixed-grid methods ignore atol / rtol, so a caller who
    passes tolerances AND a fixed grid is told instead of silently losing the accuracy they asked for."""
    if method is None:
        method = default
    if method not in METHODS + ADAPTIVE_METHODS:
        raise ValueError(f"method must be one of {METHODS + ADAPTIVE_METHODS}, got {method!r}")
    if method in METHODS and (atol != _DEFAULT_TOL or rtol != _DEFAULT_TOL):
        raise ValueError(f"atol / rtol were given together with the fixed-grid method {method!r}, which ignores them; "
                         f"drop them or use method={default!r}")
    if method in ADAPTIVE_METHODS and getattr(model, "precision", None) == "bf16" and min(atol, rtol) < 1e-3:
        warnings.warn(
            f"adaptive {method} at atol={atol:g} / rtol={rtol:g} around a bf16 velocity field: the field's own rounding "
            "noise (~5e-3 relative) is far above the tolerance, so the controller will shrink the step until the "
            "error estimate resolves that noise.  Use model.set_precision('fp32') for the reference's accuracy, loosen "
            "the tolerances, or pass an explicit fixed-grid method ('euler' | 'heun' | 'rk4').", RuntimeWarning,
            stacklevel=3)
    return method

# Butcher tableaus as torchdiffeq defines them (dopri5.py, adaptive_heun.py): alpha, beta rows, c_sol, c_error, c_mid
_DOPRI5 = dict(
    order=5,
    alpha=[1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0],
    beta=[[1 / 5],
          [3 / 40, 9 / 40],
          [44 / 45, -56 / 15, 32 / 9],
          [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
          [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
          [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]],
    c_sol=[35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0],
    c_error=[35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
             -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0],
    c_mid=[6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
           187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2],
)
_ADAPTIVE_HEUN = dict(order=2, alpha=[1.0], beta=[[1.0]], c_sol=[0.5, 0.5], c_error=[0.5, -0.5], c_mid=[0.5, 0.0])


def _flat(x):
    return x.detach().float().contiguous()


def _axpy(out, x, k, h, frozen=None):
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.ftb_ode_axpy(_lib.ptr(out), _lib.ptr(x), _lib.ptr(k), float(h), x.numel(),
                                         _lib.ptr(frozen), 0 if frozen is None else frozen.numel(),
                                         _lib.stream_ptr()))
    return out


def _zero_frozen(k, frozen):
    """dxdt[..., frozen_mask] = 0 (:73) as an in-place kernel: k = 0*k where frozen."""
    if frozen is None:
        return k
    zeros = torch.zeros_like(k)
    return _axpy(k, zeros, k, 1.0, frozen)


def _heun(out, x, k1, k2, h):
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.ftb_ode_heun_combine(_lib.ptr(out), _lib.ptr(x), _lib.ptr(k1), _lib.ptr(k2),
                                                 float(h), x.numel(), _lib.stream_ptr()))
    return out


def _rk4(out, x, k1, k2, k3, k4, h):
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.ftb_ode_rk4_combine(_lib.ptr(out), _lib.ptr(x), _lib.ptr(k1), _lib.ptr(k2),
                                                _lib.ptr(k3), _lib.ptr(k4), float(h), x.numel(),
                                                _lib.stream_ptr()))
    return out


def _prep_mask(frozen_mask, x):
    if frozen_mask is None:
        return None
    m = frozen_mask.to(device=x.device, dtype=torch.bool)
    if tuple(x.shape[x.dim() - m.dim():]) != tuple(m.shape):
        raise IndexError(f"frozen_mask of shape {tuple(m.shape)} does not index the trailing dims of {tuple(x.shape)}")
    return m.to(torch.uint8).contiguous()


def integrate_fixed(func, X0, t0, tf, n_steps, method="euler", return_trajectory=True, grid_dtype=torch.float32):
    """Fixed-grid integration of dx/dt = func(t_k (python float), x, eval_index).  ``grid_dtype``: dtype of the
    reference's ``linspace`` (fp32 at :59 / :187, float64 at :126)."""
    if method not in METHODS:
        raise ValueError(f"method must be one of {METHODS}")
    if not X0.is_cuda:
        raise RuntimeError("the samplers run on CUDA only (no CPU fallback)")
    if len(X0.shape) == 3:  # solvers.py:62-63
        X0 = X0.unsqueeze(0)
    grid = torch.linspace(t0, tf, n_steps, dtype=grid_dtype)  # on the host
    x = _flat(X0).clone()
    traj = torch.empty((n_steps,) + tuple(x.shape), dtype=torch.float32, device=x.device) if return_trajectory else None
    if traj is not None:
        traj[0].copy_(x)
    n_eval = 0
    for k in range(n_steps - 1):
        tk, tk1 = grid[k], grid[k + 1]
        h = (tk1 - tk).item()
        nxt = traj[k + 1] if traj is not None else torch.empty_like(x)
        if method == "euler":
            k1 = func(tk.item(), x, n_eval); n_eval += 1
            _axpy(nxt, x, k1, h)
        elif method == "heun":
            k1 = func(tk.item(), x, n_eval); n_eval += 1
            xe = _axpy(torch.empty_like(x), x, k1, h)
            k2 = func(tk1.item(), xe, n_eval); n_eval += 1
            _heun(nxt, x, k1, k2, h)
        else:
            tm = (tk + (tk1 - tk) / 2).item()
            k1 = func(tk.item(), x, n_eval); n_eval += 1
            xs = _axpy(torch.empty_like(x), x, k1, h / 2)
            k2 = func(tm, xs, n_eval); n_eval += 1
            xs = _axpy(torch.empty_like(x), x, k2, h / 2)
            k3 = func(tm, xs, n_eval); n_eval += 1
            xs = _axpy(torch.empty_like(x), x, k3, h)
            k4 = func(tk1.item(), xs, n_eval); n_eval += 1
            _rk4(nxt, x, k1, k2, k3, k4, h)
        x = nxt
    return traj if traj is not None else x


def _ptr_array(tensors):
    import ctypes as C
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def _dbl_array(vals):
    import ctypes as C
    return (C.c_double * len(vals))(*[float(v) for v in vals])


def integrate_adaptive(func, X0, t0, tf, n_steps, method="dopri5", rtol=1e-6, atol=1e-6, return_trajectory=True,
                       max_num_steps=100000, stats=None, grid_dtype=torch.float32, lag=1):
    """torchdiffeq's adaptive Runge-Kutta loop (rk_common.RKAdaptiveStepsizeODESolver of torchdiffeq 0.2.x, the
    un-vendored dependency behind solvers.py:77, :148, :220-222) restated: initial step from ``_select_initial_step``,
    steps accepted when the RMS of error / (atol + rtol max(|y0|, |y1|)) is <= 1, step factor
    ``min(10, max(0.9 / ratio^(1/order), 0.2))`` (dfactor 1 after an accepted step), first-same-as-last reuse of the last
    stage, and the output grid ``linspace(t0, tf, n_steps)`` evaluated with the quartic dense-output interpolant.

    The step controller is DEVICE-RESIDENT (ode_adaptive.cu): time, step size, the accept / reject decision, the output
    cursor and the counters live in a 16-double device struct; stage times, stage inputs, the error ratio, the dense
    output and the in-place state advance are kernels that read it.  The host enqueues attempted steps and never reads
    the step being computed: after enqueueing step ``s`` it looks at an asynchronous pinned copy of the controller
    from step ``s - lag`` to learn when the solve has finished (steps enqueued after that are no-ops on the state, at
    most ``lag`` of them).  No ``.item()`` per step - the reference syncs once per EVALUATION (``t.item()``, :68).
    Time-like values are doubles (torchdiffeq keeps them in float64).  ``func(T, x, eval_index) -> dx/dt`` receives the
    evaluation's time as a device fp32 tensor ``T`` of shape [B].  Parity unpinned (torchdiffeq is not installed)."""
    if method not in ADAPTIVE_METHODS:
        raise ValueError(f"method must be one of {ADAPTIVE_METHODS}, got {method!r}")
    if not X0.is_cuda:
        raise RuntimeError("the samplers run on CUDA only (no CPU fallback)")
    tab = _DOPRI5 if method == "dopri5" else _ADAPTIVE_HEUN
    order = tab["order"]
    y0 = _flat(X0).clone()
    n, B, dev, lib = y0.numel(), y0.shape[0], y0.device, _lib.lib
    # the reference builds the output grid in fp32 (solvers.py:59; float64 at :126); torchdiffeq carries time in float64
    ghost = torch.linspace(t0, tf, n_steps, dtype=grid_dtype).double() if n_steps > 1 else torch.tensor([float(t0)], dtype=torch.float64)
    n_out = ghost.numel()
    ghost = ghost.pin_memory()
    grid = ghost.to(dev, non_blocking=True)    # pinned + asynchronous: setting a solve up does not synchronise either
    ctl = torch.empty(16, dtype=torch.float64, device=dev)
    tbuf = torch.empty(B, dtype=torch.float32, device=dev)
    cp = lambda k: C.c_void_p(ctl.data_ptr() + 8 * k)
    n_eval = 0

    def stage_time(alpha):
        _lib.check(lib.ftb_ode_ctl_stage_time(_lib.ptr(tbuf), _lib.ptr(ctl), float(alpha), B, _lib.stream_ptr()))
        return tbuf

    def lincomb(ks, coefs):
        out = torch.empty_like(y0)
        _lib.check(lib.ftb_ode_lincomb_dev(_lib.ptr(out), _lib.ptr(y0), _ptr_array(ks), _dbl_array(coefs), len(ks), n,
                                           _lib.ptr(ctl), _lib.stream_ptr()))
        return out

    def sumsq(a1, a2, slot):   # ctl[slot] += sum(((a1 - a2) / (atol + rtol |y0|))^2)
        _lib.check(lib.ftb_ode_scaled_sumsq(_lib.ptr(a1), _lib.ptr(a2), _lib.ptr(y0), float(rtol), float(atol), n, cp(slot),
                                            _lib.stream_ptr()))

    with torch.cuda.device(dev):
        _lib.check(lib.ftb_ode_ctl_init(_lib.ptr(ctl), float(ghost[0]), _lib.stream_ptr()))
        # ---- _before_integrate: f0 and the first step (_select_initial_step), all on the device
        f0 = func(stage_time(0.0), y0, n_eval).clone(); n_eval += 1      # owned: advanced in place below
        sumsq(y0, None, 3)
        sumsq(f0, None, 4)
        _lib.check(lib.ftb_ode_ctl_first_step(_lib.ptr(ctl), 0, n, order, _lib.stream_ptr()))
        y1 = lincomb([f0], [1.0])
        f1 = func(stage_time(1.0), y1, n_eval); n_eval += 1
        sumsq(f1, f0, 5)
        _lib.check(lib.ftb_ode_ctl_first_step(_lib.ptr(ctl), 1, n, order, _lib.stream_ptr()))
        traj = torch.empty((n_out,) + tuple(y0.shape), dtype=torch.float32, device=dev) if return_trajectory else None
        last = None if return_trajectory else torch.empty_like(y0)
        if traj is not None:
            traj[0].copy_(y0)
        if n_out == 1:
            if stats is not None:
                stats.update(accepted=0, rejected=0, evals=n_eval)
            return traj if traj is not None else y0
        nstage = len(tab["alpha"])
        fsal = tab["c_sol"][-1] == 0 and list(tab["c_sol"][:-1]) == list(tab["beta"][-1])
        lag = max(1, int(lag))
        pinned = torch.zeros((lag + 1, 16), dtype=torch.float64).pin_memory()
        events = [torch.cuda.Event() for _ in range(lag + 1)]
        step, final = 0, None
        while final is None:
            # ---- _runge_kutta_step, enqueued blind: every time-like quantity is read from ctl by the kernels
            ks = [f0]
            for i in range(nstage):
                yi = lincomb(ks, tab["beta"][i])
                ks.append(func(stage_time(tab["alpha"][i]), yi, n_eval)); n_eval += 1
            y1 = yi if fsal else lincomb(ks, tab["c_sol"])
            f1 = ks[-1]
            _lib.check(lib.ftb_ode_error_ratio_dev(_lib.ptr(y0), _lib.ptr(y1), _ptr_array(ks), _dbl_array(tab["c_error"]),
                                                   len(ks), float(rtol), float(atol), n, _lib.ptr(ctl), _lib.stream_ptr()))
            _lib.check(lib.ftb_ode_ctl_step(_lib.ptr(ctl), _lib.ptr(grid), n_out, n, order, int(max_num_steps),
                                            _lib.stream_ptr()))
            _lib.check(lib.ftb_ode_advance(_lib.ptr(y0), _lib.ptr(f0), _lib.ptr(y1), _lib.ptr(f1), _ptr_array(ks),
                                           _dbl_array(tab["c_mid"]), len(ks), _lib.ptr(ctl), _lib.ptr(grid), n_out,
                                           _lib.ptr(traj), _lib.ptr(last), n, _lib.stream_ptr()))
            slot = step % (lag + 1)
            pinned[slot].copy_(ctl, non_blocking=True)
            events[slot].record()
            if step >= lag:   # look at a step the GPU finished while this one was being enqueued
                old = (step - lag) % (lag + 1)
                events[old].synchronize()
                flags = int(pinned[old, 13])
                if flags & 4:
                    raise FloatingPointError("adaptive solver: non-finite error estimate")
                if flags & 8:
                    raise RuntimeError(f"max_num_steps exceeded ({max_num_steps})")
                if flags & 2:
                    final = pinned[old].clone()
            step += 1
    if stats is not None:
        stats.update(accepted=int(final[9]), rejected=int(final[10]), evals=n_eval, enqueued_steps=step)
    return traj if traj is not None else last


def _integrate(func, X0, t0, tf, n_steps, method, return_trajectory, rtol, atol, stats=None, grid_dtype=torch.float32):
    if method in ADAPTIVE_METHODS:
        return integrate_adaptive(func, X0, t0, tf, n_steps, method, rtol, atol, return_trajectory, stats=stats,
                                  grid_dtype=grid_dtype)
    return integrate_fixed(func, X0, t0, tf, n_steps, method, return_trajectory, grid_dtype)


class ODEFlowSolver:
    """Flow ODE dx/dt = model(x, t) — reference ODEFlowSolver (:14-77).  Default (``method=None``): the reference's
    adaptive dopri5 with ``atol`` / ``rtol`` (:35, :77) — use it with ``model.set_precision("fp32")``: the bf16
    velocity field's own noise (~5e-3) is far above the reference's 1e-6 tolerances (a RuntimeWarning says so).
    ``method`` "euler" | "heun" | "rk4" integrate on the fixed output grid instead (see module doc)."""

    def __init__(self, model, atol=1e-6, rtol=1e-6, method=None):
        self.model = model
        self.atol = atol
        self.rtol = rtol
        self.method = _resolve_method(method, "dopri5", atol, rtol, model)

    def solve(self, X0, frozen_mask=None, t0=0.0, tf=1.0, n_steps=32, return_trajectory=True):
        if len(X0.shape) == 3:
            X0 = X0.unsqueeze(0)
        mask = _prep_mask(frozen_mask, X0)
        tbuf = torch.empty(X0.shape[0], dtype=torch.float32, device=X0.device)

        def ode_func(t, XT, _i):
            with torch.no_grad():
                # T = full((B,), t) without a device->host sync; the adaptive stepper hands the device tensor itself
                T = t if isinstance(t, torch.Tensor) else tbuf.fill_(t)
                dxdt = _flat(self.model(XT, T))
                return _zero_frozen(dxdt, mask)

        return _integrate(ode_func, X0, t0, tf, n_steps, self.method, return_trajectory, self.rtol, self.atol,
                          getattr(self, "stats", None))


class ODEOneSidedDenoisingSolver:
    """Eq. (6.7) ODE from a learned denoiser — reference :80-148."""

    def __init__(self, model, interpolant: BaseInterpolant, atol=1e-6, rtol=1e-6, method=None):
        self.model = model
        self.interp = interpolant
        self.atol = atol
        self.rtol = rtol
        self.method = _resolve_method(method, "dopri5", atol, rtol, model)   # :148
        assert isinstance(interpolant, BaseInterpolant), "ODEOneSidedDenoisingSolver requires a BaseInterpolant"
        assert self.interp.is_one_sided(), "ODEOneSidedDenoisingSolver requires a one-sided interpolant"

    def _drift(self, t, XT, eta, noise=None, eps=None):
        ip = self.interp
        out = torch.empty_like(XT)
        if isinstance(t, torch.Tensor):
            # adaptive stepper: the evaluation's time lives on the device, so do the five schedule values
            # (alpha, beta, alpha_dot, beta_dot of :138-141, eps of :207) - computed there, no host round trip
            tt = t[0]
            vals = [f(tt) for f in (ip.alpha, ip.beta, ip.alpha_dot, ip.beta_dot)]
            e = self.epsilon(tt) if eps is not None else 0.0
            vals.append(torch.as_tensor(e, dtype=torch.float32, device=XT.device))
            coef = torch.stack([v.to(device=XT.device, dtype=torch.float32).reshape(()) for v in vals]).contiguous()
            with torch.cuda.device(XT.device):
                _lib.check(_lib.lib.ftb_denoise_drift_dev(_lib.ptr(out), _lib.ptr(XT), _lib.ptr(eta), _lib.ptr(noise),
                                                          _lib.ptr(coef), 0 if eps is None else 1, XT.numel(),
                                                          _lib.stream_ptr()))
            return out
        tt = torch.tensor(t, dtype=torch.float32)
        a, b, ad, bd = (float(f(tt)) for f in (ip.alpha, ip.beta, ip.alpha_dot, ip.beta_dot))
        with torch.cuda.device(XT.device):
            _lib.check(_lib.lib.ftb_denoise_drift(
                _lib.ptr(out), _lib.ptr(XT), _lib.ptr(eta), _lib.ptr(noise), a, b, ad, bd,
                0.0 if eps is None else float(eps), 0 if eps is None else 1, XT.numel(), _lib.stream_ptr()))
        return out

    def solve(self, X0, t0=0.0, tf=1.0, n_steps=32, return_trajectory=True):
        if len(X0.shape) == 3:
            X0 = X0.unsqueeze(0)
        tbuf = torch.empty(X0.shape[0], dtype=torch.float32, device=X0.device)

        def ode_func(t, XT, _i):
            with torch.no_grad():
                T = t if isinstance(t, torch.Tensor) else tbuf.fill_(t)
                eta = _flat(self.model(XT, T))
                return self._drift(t, XT, eta)

        return _integrate(ode_func, X0, t0, tf, n_steps, self.method, return_trajectory, self.rtol, self.atol,
                          getattr(self, "stats", None), grid_dtype=torch.float64)   # float64 linspace, :126


class SDEOneSidedDenoisingSolver(ODEOneSidedDenoisingSolver):
    """Eq. (6.7) with the epsilon * score + sqrt(2 eps) * noise term — reference :153-222.
    ``epsilon`` may be a number, a 0-d tensor or a callable of t.  ``noise`` (optional callable
    ``eval_index -> tensor``) makes the draws explicit for parity tests; by default
    ``torch.randn_like`` is drawn once per evaluation like the reference (:212)."""

    def __init__(self, model, interpolant, epsilon, atol=1e-6, rtol=1e-6, method=None, noise=None):
        self.model = model
        self.interp = interpolant
        self.epsilon = epsilon if callable(epsilon) else (lambda t: epsilon)
        self.atol = atol
        self.rtol = rtol
        self.method = _resolve_method(method, "adaptive_heun", atol, rtol, model)   # :220-222
        self.noise = noise

    def solve(self, X0, t0=0.0, tf=1.0, n_steps=32, return_trajectory=True):
        assert self.interp.one_sided, "ODEOneSidedDenoisingSolver requires a one-sided interpolant"
        if self.method in ADAPTIVE_METHODS and self.noise is None:
            warnings.warn("SDEOneSidedDenoisingSolver with the reference's adaptive_heun draws fresh noise in every "
                          "evaluation, so the embedded error estimate never falls below atol / rtol and the step size "
                          "collapses (the reference behaves the same, solvers.py:212-222); pass method='heun' for a "
                          "fixed-grid solve.", RuntimeWarning, stacklevel=2)
        if len(X0.shape) == 3:
            X0 = X0.unsqueeze(0)
        tbuf = torch.empty(X0.shape[0], dtype=torch.float32, device=X0.device)

        def ode_func(t, XT, i):
            with torch.no_grad():
                dev_time = isinstance(t, torch.Tensor)
                T = t if dev_time else tbuf.fill_(t)
                eta = _flat(self.model(XT, T))
                eps = True if dev_time else float(self.epsilon(torch.tensor(t, dtype=torch.float32)))
                z = self.noise(i) if self.noise is not None else torch.randn_like(XT)
                return self._drift(t, XT, eta, _flat(z).to(XT.device), eps)

        return _integrate(ode_func, X0, t0, tf, n_steps, self.method, return_trajectory, self.rtol, self.atol,
                          getattr(self, "stats", None))


def odeSol_RK4(x0, model, nsteps=100, Tf=1.0):
    """Reference odeSol_RK4 (:225-245) with its quirks kept: starts at t = 0, performs nsteps-1
    updates of size h = Tf/nsteps (ends at t = Tf - h), t accumulated per sample in fp32."""
    if not x0.is_cuda:
        raise RuntimeError("the samplers run on CUDA only (no CPU fallback)")
    x0 = _flat(x0)
    traj = torch.zeros((nsteps,) + tuple(x0.shape), device=x0.device)
    traj[0].copy_(x0)
    t = torch.zeros(x0.shape[0], device=x0.device)
    with torch.no_grad():
        h = Tf / nsteps
        for i in range(nsteps - 1):
            xt = traj[i]
            k1 = _flat(model(xt, t))
            k2 = _flat(model(_axpy(torch.empty_like(xt), xt, k1, h / 2), t + h / 2))
            k3 = _flat(model(_axpy(torch.empty_like(xt), xt, k2, h / 2), t + h / 2))
            k4 = _flat(model(_axpy(torch.empty_like(xt), xt, k3, h), t + h))
            _rk4(traj[i + 1], xt, k1, k2, k3, k4, h)
            t = t + h
    return traj


__all__ = ["ODEFlowSolver", "ODEOneSidedDenoisingSolver", "SDEOneSidedDenoisingSolver", "odeSol_RK4",
           "integrate_fixed", "integrate_adaptive"]
