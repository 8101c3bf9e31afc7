"""oracle/ — TEST INFRASTRUCTURE, not product code.

CPU (plain torch/numpy) restatement of the flowtrain hot path: the Unet3D velocity
field, the stochastic interpolants, the fixed-grid ODE/SDE integrators and the
categorical embed/decode + training-step loss.  Every function cites the reference
file:line it follows (paths relative to the upstream repo root).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package, and only as the checker / CPU baseline.
The product package (``flowtrain_stochastic_interpolation_b200``) never imports it and has
no CPU fallback.

Pinning: the reference ships no golden vectors (its single test is a plot).
The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the
build container by importing the reference modules by file path
(``tests/golden/make_golden.py``) and committed under ``tests/golden/``.  The one part
that stays "parity unpinned" is the adaptive dopri5 / adaptive_heun stepping of the
un-vendored ``torchdiffeq`` dependency (>=0.2.5,<0.3): ``solvers.odeint_adaptive`` restates
its published algorithm and is checked against closed forms and scipy's independent
Dormand-Prince integrator, but no reference test or fixture holds an adaptive-solver result.
The fixed-grid Euler / Heun / RK4 integrators are this build's own contract (torchdiffeq
fixed-grid convention); ``odeSol_RK4`` is a restatement of reference code (solvers.py:225-245).
Also restated here: the conditioning masks (boreholes.py, pinned by the reference's own masks in
``tests/golden/cond_frontend.npz``), the conditional training loss and the ensemble statistics
(project/geodata-3d-conditional scripts), and autograd through both functional networks (pinned by
the reference modules' gradients in ``tests/golden/train*.npz``).
"""
