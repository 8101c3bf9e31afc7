#!/usr/bin/env python
"""Time conv_igemm alone (CUDA events around the launch, via ftb_profile_*) on the Unet3D layer
shapes.  usage: python tools/conv_bench.py [B]   (env FTB_NZ / FTB_WSLOT override the planner)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowtrain_stochastic_interpolation_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
SHAPES = [  # (c1, c2, cout, k, size, norm/silu/resid flags)
    (48, 0, 48, 3, 64, 1), (48, 48, 48, 3, 64, 1), (96, 0, 96, 3, 32, 1), (96, 96, 96, 3, 32, 1),
    (18, 0, 48, 7, 64, 0), (48, 0, 384, 1, 64, 0), (96, 0, 48, 3, 64, 0), (144, 0, 144, 3, 16, 1),
    (48, 48, 48, 1, 64, 0), (96, 96, 96, 1, 32, 0), (192, 0, 192, 3, 8, 1), (192, 192, 192, 3, 4, 1),
]


def run(c1, c2, cout, k, n, full):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, c1, n, n, n, generator=g).to(dev)
    x2 = torch.randn(B, c2, n, n, n, generator=g).to(dev) if c2 else None
    w = (torch.randn(cout, c1 + c2, k, k, k, generator=g) * 0.05).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    gg = torch.ones(cout, device=dev) if full else None
    rs = torch.randn(B, cout, n, n, n, generator=g).to(dev) if full else None
    out = torch.empty(B, cout, n, n, n, device=dev)
    best = 1e9
    for it in range(3):
        _lib.lib.ftb_profile_enable(1)
        _lib.check(_lib.lib.ftb_test_conv3d(
            _lib.ptr(x), c1, _lib.ptr(x2), c2, _lib.ptr(w), _lib.ptr(bias), cout, k, _lib.ptr(gg), None, None,
            _lib.ptr(rs), 1 if full else 0, _lib.ptr(out), B, n, n, n, 0, _lib.stream_ptr()))
        torch.cuda.synchronize()
        nk = 2
        fl, by, ms, ln = (C.c_double * nk)(), (C.c_double * nk)(), (C.c_double * nk)(), (C.c_int * nk)()
        _lib.check(_lib.lib.ftb_profile_collect(fl, by, ms, ln, nk))
        _lib.lib.ftb_profile_enable(0)
        t = ms[0] + ms[1]
        best = min(best, t)
    flops = 2.0 * B * n ** 3 * (c1 + c2) * cout * k ** 3
    byts = B * n ** 3 * (c1 + c2 + cout) * 2.0
    print(f"B{B} {c1}+{c2}->{cout} k{k} @{n}^3 full={full}: {best*1e3:8.1f} us  {flops/best/1e9:7.1f} TF/s  "
          f"{byts/best/1e6:7.1f} GB/s  [{os.environ.get('FTB_NZ','-')},{os.environ.get('FTB_WSLOT','-')}]", flush=True)


def dbg():
    try:
        f = _lib.lib.ftb_test_conv_debug
    except AttributeError:
        return
    buf = (C.c_longlong * (148 * 16))()
    if f(buf, 148 * 16) != 0:
        return
    import statistics
    rows = [buf[i * 16:(i + 1) * 16] for i in range(148) if buf[i * 16] > 0]
    if rows:
        med = [statistics.median(r[k] for r in rows) for k in range(10)]
        print(f"    issuer0 cycles: total {med[0]:.0f}  wait acc {med[1]:.0f}  planes {med[2]:.0f}  weights {med[3]:.0f}  "
              f"chunks {med[4]:.0f}  -> busy {med[0]-med[1]-med[2]-med[3]:.0f} (issue {med[5]:.0f}, table {med[6]:.0f}); epilogue warp4: total {med[8]:.0f} wait {med[9]:.0f}", flush=True)


for s in SHAPES:
    run(*s)
    if os.environ.get("FTB_CONV_DBG"):
        dbg()
