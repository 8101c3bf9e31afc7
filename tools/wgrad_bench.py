#!/usr/bin/env python
"""Time the tcgen05 weight-gradient kernel alone (CUDA events around the launch, via ftb_profile_*) on the
Unet3D layer shapes.  usage: python tools/wgrad_bench.py [B]   (env FTB_WGRAD_NOSTACK=1: one MMA per tap)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowtrain_stochastic_interpolation_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ONLY = int(sys.argv[2]) if len(sys.argv) > 2 else -1
dev = torch.device("cuda:0")
SHAPES = [  # (c1, c2, cout, k, size, unfold)
    (48, 0, 48, 3, 64, 0), (48, 48, 48, 3, 64, 0), (96, 0, 96, 3, 32, 0), (96, 96, 96, 3, 32, 0),
    (18, 0, 48, 7, 64, 1), (48, 0, 384, 1, 64, 0), (128, 0, 48, 1, 64, 0), (144, 0, 144, 3, 16, 0),
]


def run(c1, c2, cout, k, n, unfold):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, c1, n, n, n, generator=g).to(dev)
    x2 = torch.randn(B, c2, n, n, n, generator=g).to(dev) if c2 else None
    dy = torch.randn(B, cout, n, n, n, generator=g).to(dev)
    dw = torch.empty(cout, c1 + c2, k, k, k, device=dev)
    best = 1e9
    for it in range(3):
        _lib.lib.ftb_profile_enable(1)
        _lib.check(_lib.lib.ftb_test_conv_wgrad(_lib.ptr(x), c1, _lib.ptr(x2), c2, _lib.ptr(dy), cout, k, _lib.ptr(dw),
                                                B, n, n, n, unfold, _lib.stream_ptr()))
        torch.cuda.synchronize()
        nk = 3
        fl, by, ms, ln = (C.c_double * nk)(), (C.c_double * nk)(), (C.c_double * nk)(), (C.c_int * nk)()
        _lib.check(_lib.lib.ftb_profile_collect(fl, by, ms, ln, nk))
        _lib.lib.ftb_profile_enable(0)
        best = min(best, ms[2])
    flops = 2.0 * B * n ** 3 * (c1 + c2) * cout * k ** 3
    print(f"wgrad B{B} {c1}+{c2}->{cout} k{k} @{n}^3: {best*1e3:8.1f} us  {flops/best/1e9:7.1f} TF/s "
          f"[nostack={os.environ.get('FTB_WGRAD_NOSTACK','0')}]", flush=True)


for i, s in enumerate(SHAPES):
    if ONLY < 0 or i == ONLY:
        run(*s)
