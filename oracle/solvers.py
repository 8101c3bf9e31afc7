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
