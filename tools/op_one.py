import sys, torch
sys.path.insert(0, '.')
from stylemc_b200.ops import bias_act, upfirdn2d
f = upfirdn2d.setup_filter([1, 3, 3, 1], device='cuda')
t = torch.randn(16, 32, 1025, 1025, device='cuda')
x = torch.randn(16, 32, 1024, 1024, device='cuda', dtype=torch.float16)
b = torch.randn(32, device='cuda', dtype=torch.float16)
for _ in range(2):
    upfirdn2d.upfirdn2d(t, f, padding=[1, 1, 1, 1], gain=4)
    bias_act.bias_act(x, b, act='lrelu', clamp=256)
torch.cuda.synchronize()
