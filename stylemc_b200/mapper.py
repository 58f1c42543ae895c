"""The latent mapper and its training step (latent_mappers.py:12-93, train_latent_mapper.py:120-196) on repo kernels.

``Mapper`` turns the 8 trainable S rows of an image into that image's own delta (two ``SubMapperModulation`` halves: PixelNorm over the 4 rows,
then 5 x [Linear(512, 512) -> LayerNorm([4, 512], no affine) -> LeakyReLU]); ``train_latent_mapper.py`` optimises its weights with Adam under the
find_direction objective.  Here the synthesis + CLIP (+ identity) passes are the fused engines of ``DirectionFinder`` asked for the PER-SAMPLE style
gradient (``SynthesisEngine.backward(per_sample=True)``), and the mapper itself is an op-level network with hand-written backward passes: the
linears (forward, input gradient AND weight gradient -- the only trainable weights on this path) on ``smc_matmul_nt_f64acc``, PixelNorm on ``smc_pixelnorm``,
LayerNorm on ``smc_layernorm_fwd / bwd``, LeakyReLU on ``smc_bias_act``, Adam on ``smc_adam_step``.  Parameter names are the reference's
``state_dict`` keys (``torch.save(mapper.state_dict(), ...)``, train_latent_mapper.py:183,206), so checkpoints move both ways.
"""
import math

import torch

from . import _lib, direction
from .ops import bias_act

ROWS, WIDTH, LAYERS = 4, 512, 5            # SubMapperModulation(layernum=4), five ModulationModules (latent_mappers.py:34-39)


def _matmul_nt(a, sa, b, sb, m, n, k, bias=None):
    """y[m, n] = sum_k a[m * sa[0] + k * sa[1]] * b[n * sb[0] + k * sb[1]] (+ bias[n]) on ``smc_matmul_nt_f64acc`` (fp32 operands, float64
    accumulation).  Not the tensor-core GEMM: the output of these layers becomes the per-image delta S, and the synthesis gradient moves by 8e-4
    when delta is perturbed by 5e-6 (leaky-ReLU slope flips; measured with the float64 oracle) -- the split-fp16 GEMM's 1e-6 is too coarse."""
    y = torch.empty([m, n], dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call('smc_matmul_nt_f64acc', _lib.ptr(a), sa[0], sa[1], _lib.ptr(b), sb[0], sb[1], _lib.ptr(bias), _lib.ptr(y), m, n, k, _lib.stream())
    return y


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b with gradients w.r.t. x, W and b (torch.nn.Linear, latent_mappers.py:16)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = x.contiguous(), w.contiguous()
        ctx.save_for_backward(x, w)
        r, i = x.shape
        return _matmul_nt(x, (i, 1), w, (i, 1), r, w.shape[0], i, b.contiguous())

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        (r, i), o = x.shape, w.shape[0]
        dx = _matmul_nt(dy, (o, 1), w, (1, i), r, i, o)              # dx[r, i] = sum_o dy[r, o] W[o, i]
        dw = _matmul_nt(dy, (1, o), x, (1, i), o, i, r)              # dW[o, i] = sum_r dy[r, o] x[r, i]
        from .ops.fma import _reduce_to
        db = _reduce_to(dy, None, torch.Size([1, o])).reshape(-1)
        return dx, dw, db


class _PixelNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.call('smc_pixelnorm', _lib.ptr(x), None, _lib.ptr(y), x.shape[0], x.shape[1], x.shape[2], _lib.stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        x, = ctx.saved_tensors
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.call('smc_pixelnorm', _lib.ptr(x), _lib.ptr(dy.contiguous()), _lib.ptr(dx), x.shape[0], x.shape[1], x.shape[2], _lib.stream())
        return dx


class _LayerNormFn(torch.autograd.Function):
    """LayerNorm([4, 512], elementwise_affine=False) (latent_mappers.py:17): rows of 2048 values."""

    @staticmethod
    def forward(ctx, x):
        b = x.shape[0]
        x2 = x.reshape(b, -1).contiguous()
        wd = x2.shape[1]
        ones, zeros = torch.ones(wd, device=x.device), torch.zeros(wd, device=x.device)
        y, mean, rstd = torch.empty_like(x2), torch.empty(b, device=x.device), torch.empty(b, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call('smc_layernorm_fwd', _lib.ptr(x2), 1, 0, _lib.ptr(ones), _lib.ptr(zeros), _lib.ptr(y), None, None, _lib.ptr(mean), _lib.ptr(rstd), b, wd,
                      _lib.stream())
        ctx.save_for_backward(x2, mean, rstd, ones)
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, ones = ctx.saved_tensors
        dx = torch.empty_like(x2)
        with torch.cuda.device(x2.device):
            _lib.call('smc_layernorm_bwd', _lib.ptr(dy.reshape(x2.shape).contiguous()), _lib.ptr(x2), 1, 0, _lib.ptr(ones), _lib.ptr(mean), _lib.ptr(rstd),
                      _lib.ptr(dx), x2.shape[0], x2.shape[1], 0, _lib.stream())
        return dx.reshape(dy.shape)


class Mapper:
    """``latent_mappers.Mapper(neg_slope)`` (:68-93) with trainable parameters held as leaf tensors under the reference's state_dict keys."""

    def __init__(self, neg_slope=0.01, device='cuda', seed=0):
        self.device, self.neg_slope = torch.device(device), float(neg_slope)
        g = torch.Generator().manual_seed(seed)
        bound = 1.0 / math.sqrt(WIDTH)                                # torch.nn.Linear's default init range
        self.params = {}
        for half in ('course_mapping', 'medium_mapping'):             # (sic: latent_mappers.py:71)
            for i in range(LAYERS):
                k = f'{half}.modulation_module_list.{i}.fc.'
                self.params[k + 'weight'] = ((torch.rand(WIDTH, WIDTH, generator=g) * 2 - 1) * bound).to(self.device).requires_grad_(True)
                self.params[k + 'bias'] = ((torch.rand(WIDTH, generator=g) * 2 - 1) * bound).to(self.device).requires_grad_(True)

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.params.items()}

    def load_state_dict(self, sd):
        if set(sd) != set(self.params):
            raise RuntimeError(f'mapper state_dict keys differ: {sorted(set(sd) ^ set(self.params))[:4]} ...')
        with torch.no_grad():
            for k, v in sd.items():
                self.params[k].copy_(v.to(self.device, torch.float32))

    def _half(self, x, name):
        """SubMapperModulation.forward (:41-45) with embedding=None."""
        b = x.shape[0]
        x = _PixelNormFn.apply(x)
        for i in range(LAYERS):
            k = f'{name}.modulation_module_list.{i}.fc.'
            y = _LinearFn.apply(x.reshape(b * ROWS, WIDTH), self.params[k + 'weight'], self.params[k + 'bias']).reshape(b, ROWS, WIDTH)
            x = bias_act.bias_act(_LayerNormFn.apply(y), act='lrelu', alpha=self.neg_slope, gain=1, dim=2)      # LeakyReLU(neg_slope), :22,31
        return x

    def __call__(self, x):
        """x [B, 8, 512] (the trainable S rows of each image) -> delta [B, 8, 512] (:75-93)."""
        _lib.require_cuda(x, 'x')
        x = x.to(torch.float32)
        return torch.cat([self._half(x[:, :ROWS].contiguous(), 'course_mapping'), self._half(x[:, ROWS:2 * ROWS].contiguous(), 'medium_mapping')], dim=1)


class MapperTrainer:
    """One optimisation step of train_latent_mapper.py:136-196: delta = mapper(styles[:, rows]); the find_direction objective on
    (styles + delta, styles) with a per-image delta; Adam on the mapper's weights with the cosine learning-rate rule (:143-147)."""

    def __init__(self, G, clip_model, pos_tokens, neg_tokens, resolution, mapper, learning_rate=0.001, betas=(0.9, 0.999), **finder_kw):
        self.finder = direction.DirectionFinder(G, clip_model, pos_tokens, neg_tokens, resolution, device=mapper.device, **finder_kw)
        self.mapper, self.lr, self.betas, self.t = mapper, learning_rate, betas, 0
        self.m = {k: torch.zeros_like(v) for k, v in mapper.params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in mapper.params.items()}

    def loss_and_grads(self, styles):
        """-> (dict of losses, {parameter name: gradient}).  styles [n, 26, 512] on the device."""
        f = self.finder
        styles = styles.to(f.device, torch.float32)
        n = styles.shape[0]
        rows = f._rows_idx if getattr(f, '_rows_idx', None) is not None else torch.tensor(f.rows, dtype=torch.int64, device=f.device)
        f._rows_idx = rows
        for p in self.mapper.params.values():
            p.grad = None
        delta = self.mapper(styles.index_select(1, rows))                              # train_latent_mapper.py:155-156
        styles2 = styles.clone().index_add_(1, rows, delta.detach())                   # :157-158
        g_delta, parts = f.loss_and_grad(styles, n, styles_edit=styles2, per_sample=True)
        l2 = f.l2_reg_coef * delta.detach().square().mean()                            # find_direction.py:190-191 on the per-image delta
        g_total = g_delta + (2.0 * f.l2_reg_coef / delta.numel()) * delta.detach()
        delta.backward(g_total)
        clip_loss = f.clip_loss_coef * sum(w for _, _, w in f.clips) + parts[:1]
        out = dict(clip_loss=clip_loss, l2_loss=l2, loss=clip_loss + l2)
        if f._use_id():
            out['identity_loss'] = f.identity_loss_coef + parts[1:2]
            out['loss'] = out['loss'] + out['identity_loss']
        return out, {k: p.grad for k, p in self.mapper.params.items()}

    def step(self, styles, lr=None):
        lr = self.lr if lr is None else lr
        out, grads = self.loss_and_grads(styles)
        self.t += 1
        b1, b2 = self.betas
        with torch.no_grad(), torch.cuda.device(self.mapper.device):
            for k, p in self.mapper.params.items():
                _lib.call('smc_adam_step', _lib.ptr(p), _lib.ptr(grads[k].contiguous()), _lib.ptr(self.m[k]), _lib.ptr(self.v[k]), p.numel(), float(lr),
                          b1, b2, 1e-8, 1.0 - b1 ** self.t, math.sqrt(1.0 - b2 ** self.t), _lib.stream())
        return out
