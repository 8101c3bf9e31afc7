import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowtrain_stochastic_interpolation_b200 import _lib
dev = torch.device("cuda:0")
B, n, c1, cout, k = 8, 64, 18, 48, 7
g = torch.Generator().manual_seed(0)
x = torch.randn(B, c1, n, n, n, generator=g).to(dev)
w = (torch.randn(cout, c1, k, k, k, generator=g) * 0.05).to(dev)
bv = torch.randn(cout, generator=g).to(dev)
out = torch.empty(B, cout, n, n, n, device=dev)
best = 1e9
for it in range(3):
    _lib.lib.ftb_profile_enable(1)
    _lib.check(_lib.lib.ftb_test_conv3d(_lib.ptr(x), c1, None, 0, _lib.ptr(w), _lib.ptr(bv), cout, k, None, None, None, None, 0, _lib.ptr(out), B, n, n, n, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    nk = 2
    fl, by, ms, ln = (C.c_double * nk)(), (C.c_double * nk)(), (C.c_double * nk)(), (C.c_int * nk)()
    _lib.check(_lib.lib.ftb_profile_collect(fl, by, ms, ln, nk)); _lib.lib.ftb_profile_enable(0)
    best = min(best, ms[0] + ms[1])
print(f"stem: {best*1e3:8.1f} us  cap={os.environ.get('FTB_CGCAP')} nz={os.environ.get('FTB_NZ')} iss={os.environ.get('FTB_NISS')} ws={os.environ.get('FTB_WSLOT')}", flush=True)
