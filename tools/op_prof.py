"""One call of each op-level kernel of interest (for ncu): python tools/op_prof.py"""
import sys, torch
sys.path.insert(0, '.')
from stylemc_b200.ops import bias_act, upfirdn2d
f = upfirdn2d.setup_filter([1, 3, 3, 1], device='cuda')
t = torch.randn(16, 32, 1025, 1025, device='cuda')
th = t.half()
for _ in range(2):
    upfirdn2d.upfirdn2d(t, f, padding=[1, 1, 1, 1], gain=4)
    upfirdn2d.upfirdn2d(th, f, padding=[1, 1, 1, 1], gain=4)
torch.cuda.synchronize()
