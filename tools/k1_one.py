import sys, torch
sys.path.insert(0, '/root/repo')
from flowtrain_stochastic_interpolation_b200 import _lib
dev = torch.device("cuda:0")
B, n, c1, cout = 8, 64, 48, 128
g = torch.Generator().manual_seed(0)
x = torch.randn(B, c1, n, n, n, generator=g).to(dev)
w = (torch.randn(cout, c1, 1, 1, 1, generator=g) * 0.05).to(dev)
out = torch.empty(B, cout, n, n, n, device=dev)
for it in range(2):
    _lib.check(_lib.lib.ftb_test_conv3d(_lib.ptr(x), c1, None, 0, _lib.ptr(w), None, cout, 1, None, None, None, None, 0, _lib.ptr(out), B, n, n, n, 0, _lib.stream_ptr()))
torch.cuda.synchronize()
