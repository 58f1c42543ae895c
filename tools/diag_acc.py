"""Dev diagnostic: how much of the x3 error is tensor-core accumulation (long K) vs operand split."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from stylemc_b200 import gemm
torch.manual_seed(0)
n, c, o, h = 2, 512, 512, 16
x = torch.randn(n, c, h, h, device='cuda').abs() * 1.0      # positive-ish activations like lrelu outputs
x = torch.where(torch.rand_like(x) < 0.5, x, -0.2 * x)
w = torch.randn(o, c, 3, 3, device='cuda')
ref = F.conv2d(x.double(), w.double(), padding=1).permute(0, 2, 3, 1)
A = gemm.split_planes(x.permute(0, 2, 3, 1).contiguous(), True).reshape(2 * n, h, h, c)
B = gemm.split_planes(w.permute(2, 3, 0, 1).reshape(9 * o, c).contiguous(), True).reshape(2 * 9 * o, c)
def rel(a): return ((a.double() - ref).abs().max() / ref.abs().max()).item(), ((a.double() - ref).norm() / ref.norm()).item()
out = torch.empty(n, h, h, o, device='cuda')
gemm.igemm(A, B, n, h, h, o, gemm.TAPS_3X3, precision='x3', a_plane_stride_imgs=n, b_rows_per_tap=9 * o, out_f32=out)
print('x3 one launch (K=13824): max-rel %.3e l2-rel %.3e' % rel(out))
gemm.igemm(A, B, n, h, h, o, gemm.TAPS_3X3, precision='x1', out_f32=out)
print('x1 one launch: max-rel %.3e l2-rel %.3e' % rel(out))
acc = torch.zeros_like(out); tmp = torch.empty_like(out)
for t in gemm.TAPS_3X3:
    gemm.igemm(A, B, n, h, h, o, [t], precision='x3', a_plane_stride_imgs=n, b_rows_per_tap=9 * o, out_f32=tmp)
    acc += tmp
print('x3 per-tap launches summed in fp32 (K=1536 each): max-rel %.3e l2-rel %.3e' % rel(acc))
acc = torch.zeros(n, h, h, o, device='cuda', dtype=torch.float64)
for t in gemm.TAPS_3X3:
    gemm.igemm(A, B, n, h, h, o, [t], precision='x3', a_plane_stride_imgs=n, b_rows_per_tap=9 * o, out_f32=tmp)
    acc += tmp.double()
print('  ... summed in fp64: max-rel %.3e l2-rel %.3e' % rel(acc))
for ck in (2048, 1024, 512, 256):
    gemm.igemm(A, B, n, h, h, o, gemm.TAPS_3X3, precision='x3', a_plane_stride_imgs=n, b_rows_per_tap=9 * o, out_f32=out, acc_chunk_k=ck)
    print('x3 promoted accumulation chunk K=%d: max-rel %.3e l2-rel %.3e' % ((ck,) + rel(out)))
# pure fp32 conv on the GPU (cuDNN) for scale
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
print('torch fp32 conv2d: max-rel %.3e l2-rel %.3e' % rel(F.conv2d(x, w, padding=1).permute(0, 2, 3, 1)))
# signed error: is the tensor-core accumulate biased toward zero?
gemm.igemm(A, B, n, h, h, o, gemm.TAPS_3X3, precision='x3', a_plane_stride_imgs=n, b_rows_per_tap=9 * o, out_f32=out)
e = (out.double() - ref) * ref.sign()
print('mean signed error along sign(ref): %.3e (rms err %.3e)' % (e.mean().item(), (out.double() - ref).pow(2).mean().sqrt().item()))
