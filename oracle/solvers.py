"""Fixed-grid ODE / SDE integrators around a velocity (or denoiser) callable
(TEST INFRASTRUCTURE).

The reference samplers (src/flowtrain/solvers/solvers.py:40-77, :120-148, :180-222) hand
``ode_func`` to torchdiffeq's adaptive dopri5 / adaptive_heun.  torchdiffeq (>=0.2.5,<0.3,
pyproject.toml:19) is not vendored and not installed, so the adaptive controller is
"parity unpinned".  What is restated here:

  * the time grid ``t = linspace(t0, tf, n_steps)`` (:59) — n_steps is the number of grid
    POINTS, so a "100-step" solve is n_steps=101; the returned trajectory has n_steps
    entries like ``odeint`` output;
  * ``ode_func``: ``T = full((B,), t.item())`` in fp32, ``dxdt = model(XT, T)``, optional
    ``dxdt[..., frozen_mask] = 0`` (:66-74);
  * eq-6.7 drift of the one-sided denoising solvers (:130-143, :193-216) with the SDE noise
    passed in explicitly (the reference draws ``randn_like`` inside ode_func);
  * fixed-grid steppers in torchdiffeq's fixed-grid convention (this build's contract):
      euler : x += h f(t_k, x)
      heun  : k1 = f(t_k, x); k2 = f(t_k+h, x + h k1); x += h/2 (k1 + k2)
      rk4   : classic 4-stage, combine exactly as odeSol_RK4 (:235-240)
  * ``odeSol_RK4`` itself (:225-245), including its quirks: starts at t=0, does nsteps-1
    updates with h = Tf/nsteps (ends at t = Tf - h).
"""
from __future__ import annotations

import torch

from . import interp as _interp


def time_grid(t0, tf, n_steps):
    return torch.linspace(t0, tf, n_steps)  # solvers.py:59 (fp32, CPU)


def _full_t(x, tval):
    return torch.full((x.shape[0],), float(tval), device=x.device, dtype=torch.float32)


def make_flow_func(model, frozen_mask=None):
    """ode_func of ODEFlowSolver.solve — solvers.py:66-74."""

    def f(tval, x, step=None):
        with torch.no_grad():
            d = model(x, _full_t(x, tval))
            if frozen_mask is not None:
                d[..., frozen_mask] = 0
            return d

    return f


def make_denoise_func(model, kind="linear", one_sided=True, gamma_a=2.0, epsilon=None,
                      noise=None):
    """ode_func of ODEOneSidedDenoisingSolver (:130-143) and, when ``epsilon`` is given,
    SDEOneSidedDenoisingSolver (:193-216).  ``noise(step_index, stage_index)`` supplies the
    standard-normal tensor the reference draws with randn_like (:212)."""
    counter = {"n": 0}

    def f(tval, x, step=None):
        with torch.no_grad():
            eta = model(x, _full_t(x, tval))
            tt = torch.tensor(float(tval), dtype=torch.float32, device=x.device)
            a, b, _, ad, bd, _ = _interp.coeffs(kind, tt, one_sided, gamma_a)
            d = ad * eta + (bd / b) * (x - a * eta)
            if epsilon is not None:
                eps = epsilon(tt) if callable(epsilon) else epsilon
                eps = torch.as_tensor(eps, dtype=torch.float32, device=x.device)
                score = -eta / a
                z = noise(counter["n"])
                counter["n"] += 1
                d = d + (eps * score + z * torch.sqrt(2 * eps))
            return d

    return f


def integrate(f, x0, t0=0.0, tf=1.0, n_steps=32, method="euler"):
    """Fixed-grid integration on linspace(t0, tf, n_steps); returns [n_steps, *x0.shape]."""
    t = time_grid(t0, tf, n_steps)
    x = x0.clone()
    traj = [x.clone()]
    for k in range(n_steps - 1):
        tk, tk1 = t[k], t[k + 1]
        h = (tk1 - tk).item()  # fp32 difference of fp32 grid points
        if method == "euler":
            x = x + h * f(tk.item(), x)
        elif method == "heun":
            k1 = f(tk.item(), x)
            k2 = f(tk1.item(), x + h * k1)
            x = x + (h / 2) * (k1 + k2)
        elif method == "rk4":
            tm = (tk + (tk1 - tk) / 2).item()
            k1 = f(tk.item(), x)
            k2 = f(tm, x + h * k1 / 2)
            k3 = f(tm, x + h * k2 / 2)
            k4 = f(tk1.item(), x + h * k3)
            x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        else:
            raise ValueError(method)
        traj.append(x.clone())
    return torch.stack(traj)


def ode_sol_rk4(x0, model, nsteps=100, Tf=1.0):
    """odeSol_RK4 — solvers.py:225-245 (t is a per-sample fp32 vector accumulated by +h)."""
    traj = torch.zeros(nsteps, *x0.shape, device=x0.device)
    traj[0] = x0
    t = torch.zeros(x0.shape[0], device=x0.device)
    with torch.no_grad():
        h = Tf / nsteps
        for i in range(nsteps - 1):
            xt = traj[i]
            k1 = model(xt, t)
            k2 = model(xt + h * k1 / 2, t + h / 2)
            k3 = model(xt + h * k2 / 2, t + h / 2)
            k4 = model(xt + h * k3, t + h)
            traj[i + 1] = traj[i] + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
            t = t + h
    return traj


# ---------------------------------------------------------------------------------- adaptive Runge-Kutta
# torchdiffeq (>=0.2.5,<0.3, pyproject.toml:19) is the un-vendored dependency behind ODEFlowSolver's default
# method="dopri5" (solvers.py:77, :148) and the SDE solver's "adaptive_heun" (:220-222); it is not installed here, so
# this restates its published algorithm (torchdiffeq/_impl/rk_common.py, dopri5.py, adaptive_heun.py, misc.py) with
# plain torch ops.  PARITY UNPINNED: no reference test or fixture holds an adaptive-solver result.
DOPRI5 = dict(
    order=5,
    alpha=[1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0],
    beta=[[1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9],
          [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
          [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
          [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]],
    c_sol=[35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0],
    c_error=[35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
             -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0],
    c_mid=[6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
           187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2])
ADAPTIVE_HEUN = dict(order=2, alpha=[1.0], beta=[[1.0]], c_sol=[0.5, 0.5], c_error=[0.5, -0.5], c_mid=[0.5, 0.0])


def _rms(x):
    return float(x.double().pow(2).mean().sqrt())


def odeint_adaptive(func, y0, t, method="dopri5", rtol=1e-6, atol=1e-6, stats=None):
    """odeint(func, y0, t, method=...) of torchdiffeq restated.  func(t: float, y) -> dy/dt; t: 1-d tensor of output
    times (ascending).  Returns [len(t), *y0.shape]."""
    tab = DOPRI5 if method == "dopri5" else ADAPTIVE_HEUN
    order = tab["order"]
    tt = [float(v) for v in t]
    y0 = y0.clone()
    f0 = func(tt[0], y0)
    # misc._select_initial_step(func, t0, y0, order - 1, rtol, atol, norm, f0)
    scale = atol + y0.abs() * rtol
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = func(tt[0] + h0, y0 + h0 * f0)
    d2 = _rms((f1 - f0) / scale) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / float(order))
    dt = min(100 * h0, h1)
    out = [y0.clone()]
    t_cur, interp, acc_n, rej_n = tt[0], None, 0, 0
    for t_out in tt[1:]:
        while t_out > t_cur:
            t1 = t_cur + dt
            k = [f0]
            for a, brow in zip(tab["alpha"], tab["beta"]):     # rk_common._runge_kutta_step
                ti = t1 if a == 1.0 else t_cur + a * dt
                yi = y0.clone()
                for kj, b in zip(k, brow):
                    if b != 0:
                        yi = yi + kj * (b * dt)
                k.append(func(ti, yi))
            if not (tab["c_sol"][-1] == 0 and list(tab["c_sol"][:-1]) == list(tab["beta"][-1])):
                yi = y0.clone()
                for kj, c in zip(k, tab["c_sol"]):
                    if c != 0:
                        yi = yi + kj * (c * dt)
            y1, f1 = yi, k[-1]
            err = torch.zeros_like(y0)
            for kj, c in zip(k, tab["c_error"]):
                if c != 0:
                    err = err + kj * (c * dt)
            ratio = _rms(err / (atol + rtol * torch.max(y0.abs(), y1.abs())))   # _compute_error_ratio
            if ratio <= 1:
                ymid = y0.clone()
                for kj, c in zip(k, tab["c_mid"]):
                    if c != 0:
                        ymid = ymid + kj * (c * dt)
                interp = (t_cur, t1, y0, y1, ymid, f0, f1, dt)
                y0, f0, t_cur = y1, f1, t1
                acc_n += 1
            else:
                rej_n += 1
            if ratio == 0:                                     # _optimal_step_size
                dt = dt * 10.0
            else:
                dfactor = 1.0 if ratio < 1 else 0.2
                dt = dt * min(10.0, max(0.9 / ratio ** (1.0 / order), dfactor))
        ta, tb, ya, yb, ym, fa, fb, h = interp                 # _interp_fit / _interp_evaluate
        a = 2 * h * (fb - fa) - 8 * (yb + ya) + 16 * ym
        b = h * (5 * fa - 3 * fb) + 18 * ya + 14 * yb - 32 * ym
        c = h * (fb - 4 * fa) - 11 * ya - 5 * yb + 16 * ym
        d = h * fa
        x = (t_out - ta) / (tb - ta)
        out.append(ya + x * (d + x * (c + x * (b + x * a))))
    if stats is not None:
        stats.update(accepted=acc_n, rejected=rej_n)
    return torch.stack(out)
