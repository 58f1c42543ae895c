"""The find_direction optimisation step (find_direction.py:256-353) on the fused engines, and its data-parallel form.

One step = two synthesis forwards (edited + original S), ``unprocess`` of both, two CLIP ViT-B/32 forwards, the directional
CLIP loss (clip_loss.py:24-34), the backward pass CLIP -> unprocess -> synthesis down to the trainable S rows, the analytic
gradient of the L2 term (find_direction.py:190-191) and one SGD update (:285,339).  Across GPUs the seed batch is sharded by
rows; the only exchange is an all-reduce of the ``[8, 512]`` gradient (16 KiB) and of the loss scalars.

``clip_loss_type='default'`` with ``clip_type='small'`` (one ViT-B/32 tower, the benchmark configuration) or ``'double'`` (ViT-B/32 +
0.5 * ViT-B/16, find_direction.py:117-119,160-164: pass both models); identity and landmark terms are outside the accelerated path
(SURVEY.md section 2).
"""
import contextlib
import math
import os

import torch

from . import _lib, resample, synthesis, utils

S_TRAINABLE_SPACE_CHANNELS = [2, 3, 5, 6, 8, 9, 11, 12]     # find_direction.py:41
N_STYLE_CHANNELS = 26                                       # find_direction.py:38
RESOLUTION_DICT = {256: 6, 512: 7, 1024: 8}                 # find_direction.py:263
DOUBLE_CLIP_WEIGHTS = (1.0, 0.5)                            # find_direction.py:164: clip1 + 0.5 * clip2 (ViT-B/32, ViT-B/16)


_NVTX = os.environ.get('STYLEMC_NVTX', '0') != '0'


def _phase(name):
    """NVTX range around a phase of the step when STYLEMC_NVTX=1 (the reference's only tracing hook is ``misc.profiled_function`` ->
    ``torch.autograd.profiler.record_function``, misc.py:98-103); a no-op otherwise."""
    return torch.cuda.nvtx.range(name) if _NVTX else contextlib.nullcontext()


class CLIPLoss:
    """clip_loss.CLIPLoss (clip_loss.py:8-34) with the model and the token ids injected (``clip.load`` / ``clip.tokenize`` are
    external).  ``__call__`` is the differentiable drop-in; ``loss_and_grad`` is the fused kernel used by the step."""

    normalize_features = False          # clip_loss.py:25-27: the difference of the raw embeddings
    preprocess = 'unprocess'            # find_direction.py:49-52 ahead of the towers
    needs_source = True                 # the original image's embedding enters the loss

    def __init__(self, model, pos_tokens, neg_tokens):
        self.model = model
        t = model.encode_text(pos_tokens) - model.encode_text(neg_tokens)            # clip_loss.py:15-17
        self.text_features = (t / t.norm(dim=1, keepdim=True)).contiguous()          # :18, [1, 512]

    def __call__(self, src_image, tgt_image):
        e = self.model.encode_image(tgt_image) - self.model.encode_image(src_image)  # :25-27
        e = e / e.norm(dim=1, keepdim=True)                                          # :28
        cos = torch.nn.functional.cosine_similarity(e, self.text_features)           # :29-32
        return (len(src_image) - cos.sum()) / len(src_image)                         # :34

    def loss_and_grad(self, e_src, e_tgt, coef, inv_count, want_grad=True, gscale_target=64.0):
        """Returns (loss_part, d_tgt, gscale): loss_part = -coef * inv_count * sum_n cos_n (add ``coef * n * inv_count`` for
        the full term); d_tgt = S * d(loss)/d(e_tgt) with the power-of-two loss scale S stored in the device scalar gscale."""
        n, e = e_tgt.shape
        dev = e_tgt.device
        part = torch.empty(1, dtype=torch.float32, device=dev)
        d_tgt = torch.empty_like(e_tgt) if want_grad else None
        gscale = torch.ones(1, dtype=torch.float32, device=dev) if want_grad else None
        with torch.cuda.device(dev):
            _lib.call('smc_clip_loss', _lib.ptr(e_src), _lib.ptr(e_tgt), _lib.ptr(self.text_features), _lib.ptr(part), _lib.ptr(d_tgt), n, e,
                      float(coef), float(inv_count), _lib.ptr(gscale), float(gscale_target), int(self.normalize_features),
                      e if self.text_features.shape[0] > 1 else 0, _lib.stream())
        return part, d_tgt, gscale


# clip_loss_nada.py:12-40: every class string is embedded through these 27 prompt templates (data of the loss definition)
NADA_TEMPLATES = [
    'a photo of a {}.', 'a rendering of a {}.', 'a cropped photo of the {}.', 'the photo of a {}.', 'a photo of a clean {}.', 'a photo of a dirty {}.',
    'a dark photo of the {}.', 'a photo of my {}.', 'a photo of the cool {}.', 'a close-up photo of a {}.', 'a bright photo of the {}.',
    'a cropped photo of a {}.', 'a photo of the {}.', 'a good photo of the {}.', 'a photo of one {}.', 'a close-up photo of the {}.',
    'a rendition of the {}.', 'a photo of the clean {}.', 'a rendition of a {}.', 'a photo of a nice {}.', 'a good photo of a {}.',
    'a photo of the nice {}.', 'a photo of the small {}.', 'a photo of the weird {}.', 'a photo of the large {}.', 'a photo of a cool {}.',
    'a photo of a small {}.']


def nada_template_texts(class_str):
    """clip_loss_nada.py:203-204 ``compose_text_with_templates``: the strings to hand to the tokenizer."""
    return [t.format(class_str) for t in NADA_TEMPLATES]


class CLIPLossNADA(CLIPLoss):
    """``clip_loss_type='nada'`` (find_direction.py:101-107; clip_loss_nada.py:206-218 ``clip_directional_loss``): the text direction is the
    normalised mean over the prompt templates of (target - source) of the NORMALISED text embeddings (:150-157), both image embeddings are
    normalised before the difference (:142-148), and the images go through the NADA preprocessing (:86-89: no clamp) instead of
    ``unprocess``.  ``source_tokens`` / ``target_tokens``: the tokenised ``nada_template_texts`` of the negative / positive prompt, [27, 77]."""
    normalize_features = True
    preprocess = 'nada'

    def __init__(self, model, source_tokens, target_tokens):
        self.model = model
        s, t = model.encode_text(source_tokens), model.encode_text(target_tokens)
        s, t = s / s.norm(dim=-1, keepdim=True), t / t.norm(dim=-1, keepdim=True)          # get_text_features(norm=True), :129-140
        d = (t - s).mean(dim=0, keepdim=True)                                               # :154
        self.text_features = (d / d.norm(dim=-1, keepdim=True)).contiguous()                # :155, [1, 512]

    def __call__(self, src_image, tgt_image):
        """Differentiable drop-in on raw GAN outputs (what find_direction.py:151-158 passes)."""
        a = self.model.encode_image(resample.nada_preprocess(tgt_image))
        b = self.model.encode_image(resample.nada_preprocess(src_image))
        e = a / a.norm(dim=-1, keepdim=True) - b / b.norm(dim=-1, keepdim=True)
        e = e / e.norm(dim=-1, keepdim=True)
        return (1.0 - torch.nn.functional.cosine_similarity(e, self.text_features)).mean()


class CLIPLossNADAGlobal(CLIPLoss):
    """``clip_loss_type='nada_global'`` (find_direction.py:108-114; clip_loss_nada.py:220-229 ``global_clip_loss``):
    ``mean(1 - logits_per_image / 100)`` with ``logits = exp(logit_scale) * cos(image, text)`` for the single prompt ``f'a {target_class}'``
    (:326-327).  The original image does not enter this loss.  ``text_tokens`` [1, 77]."""
    preprocess = 'nada'
    needs_source = False

    def __init__(self, model, text_tokens):
        self.model = model
        t = model.encode_text(text_tokens)
        if t.shape[0] != 1:
            raise RuntimeError('nada_global: one prompt (find_direction.py passes [f"a {target_class}"])')
        self.text_features = (t / t.norm(dim=-1, keepdim=True)).contiguous()
        self.logit_gain = math.exp(model.logit_scale) / 100.0

    def __call__(self, src_image, tgt_image):
        a = self.model.encode_image(resample.nada_preprocess(tgt_image))
        return (1.0 - self.logit_gain * torch.nn.functional.cosine_similarity(a, self.text_features)).mean()

    def loss_and_grad(self, e_src, e_tgt, coef, inv_count, want_grad=True, gscale_target=64.0):
        """cos(e_tgt, text) is the directional kernel with a zero source embedding; the logit gain rides in the coefficient."""
        return super().loss_and_grad(torch.zeros_like(e_tgt), e_tgt, coef * self.logit_gain, inv_count, want_grad, gscale_target)


def shard_rows(n_total, rank, world):
    """Rows [lo, hi) of a global batch that rank ``rank`` of ``world`` processes (contiguous, sizes differ by at most one)."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_step(grad, part, group, local_rows=None):
    """Sum the shard gradients [rows, 512] and the loss partial sums over the ranks in ONE collective (16 KiB + 4 B).  Because
    every rank divides by the GLOBAL seed count (loss_and_grad's ``global_count``), the sum is exactly the full-batch mean
    gradient of clip_loss.py:34, also for ragged shards.  When the caller does not know the global count, ``local_rows`` (this
    rank's row count) rides in the same buffer: ``grad`` / ``part`` must then be UN-normalised sums (count 1) and the result is
    divided by the all-reduced row count on the device -- ragged or empty shards included, no guess, no host sync."""
    tail = [] if local_rows is None else [torch.full([1], float(local_rows), dtype=grad.dtype, device=grad.device)]
    buf = torch.cat([grad.reshape(-1), part.reshape(-1)] + tail)
    torch.distributed.all_reduce(buf, op=torch.distributed.ReduceOp.SUM, group=group)
    if local_rows is not None:
        buf = buf[:-1] / buf[-1]
    return buf[:grad.numel()].reshape(grad.shape), buf[grad.numel():]


def cosine_lr(base_lr, it, total):
    """find_direction.py:298-299 (``it`` is 1-based)."""
    return math.cos(math.pi * it / total) * base_lr * 0.5 + base_lr * 0.5


class DirectionFinder:
    """State and step function of the S-space direction search for one prompt pair.

    G: frozen generator (attribute structure of the unpickled network); clip_model: stylemc_b200.clip.CLIPModel (clip_type='small'),
    or the pair (ViT-B/32 model, ViT-B/16 model) for clip_type='double' (weights 1 and 0.5, find_direction.py:164);
    resolution: 256 / 512 / 1024 -> until_k per find_direction.py:263 (the network is truncated after that block).
    precision: synthesis engine mode; 'x3p' reproduces the fp32 reference's gradient to <= 1e-3, 'x1' is the fast fp16-operand
    mode (images <= 1e-2, loss <= 1e-3, gradient ~1e-2: the lrelu kinks amplify forward rounding, DESIGN.md "Numerics").
    process_group: torch.distributed group for the data-parallel run (None = single process)."""

    def __init__(self, G, clip_model, pos_tokens, neg_tokens, resolution, device='cuda', learning_rate=1.5, clip_loss_coef=1.0,
                 l2_reg_coef=0.1, noise_mode='const', precision='x3p', micro_batch=16, process_group=None,
                 trainable_rows=S_TRAINABLE_SPACE_CHANNELS, original_precision=None, clip_loss_type='default', id_loss=None,
                 identity_loss_coef=0.0):
        self.device = torch.device(device)
        self._src_cache = {}                 # (source_key, first row, rows) -> CLIP embeddings of the un-edited images (loss_and_grad)
        self.engine = utils.engine_for(G, self.device, precision)
        # the original-image branch carries no gradient (find_direction.py:312); it may run in another engine mode (diagnostics:
        # tests/diag/diag_original_branch.py measures what that does to the loss and the gradient).  Default: the same engine.
        self.engine_original = self.engine if original_precision in (None, precision) else utils.engine_for(G, self.device, original_precision)
        self.until_k = RESOLUTION_DICT[resolution] if resolution in RESOLUTION_DICT else int(math.log2(resolution)) - 2
        self.until_k = min(self.until_k, len(self.engine.blocks) - 1)
        models = list(clip_model) if isinstance(clip_model, (tuple, list)) else [clip_model]
        if len(models) > len(DOUBLE_CLIP_WEIGHTS):
            raise ValueError('clip_model: one model (clip_type small) or two (clip_type double)')
        # find_direction.py:100-122 ``init_clip_loss``.  'default': pos / neg = the tokenised prompt and negative prompt; 'nada': the tokenised
        # ``nada_template_texts`` of the prompt / negative prompt ([27, 77] each); 'nada_global': pos = tokenised f'a {prompt}' (neg unused)
        if clip_loss_type == 'default':
            make = lambda m: CLIPLoss(m, pos_tokens, neg_tokens)
        elif clip_loss_type == 'nada':
            make = lambda m: CLIPLossNADA(m, neg_tokens, pos_tokens)
        elif clip_loss_type == 'nada_global':
            make = lambda m: CLIPLossNADAGlobal(m, pos_tokens)
        else:
            raise ValueError("clip_loss_type must be 'default', 'nada' or 'nada_global'")
        self.clip_loss_type = clip_loss_type
        # identity term (find_direction.py:179-180,193): ``id_loss`` = stylemc_b200.idloss.IDLoss; its coefficient defaults to 0 here (the
        # benchmark objective is the CLIP + L2 terms), the reference CLI's default is 0.6 (find_direction.py:225)
        self.id_loss, self.identity_loss_coef = id_loss, float(identity_loss_coef)
        if self.identity_loss_coef != 0.0 and id_loss is None:
            raise ValueError('identity_loss_coef != 0 needs id_loss (stylemc_b200.idloss.IDLoss)')
        self.clips = [(m, make(m), w) for m, w in zip(models, DOUBLE_CLIP_WEIGHTS)]
        self.clip, self.loss_fn = self.clips[0][0], self.clips[0][1]
        self.lr, self.clip_loss_coef, self.l2_reg_coef = learning_rate, clip_loss_coef, l2_reg_coef
        self.noise_mode, self.micro_batch = noise_mode, micro_batch
        self.rows = list(trainable_rows)
        self.delta = torch.zeros([1, len(self.rows), synthesis.STYLE_WIDTH], dtype=torch.float32, device=self.device)   # :270-273
        self.group = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.kernel_launches = 0
        # The original-image branch (find_direction.py:312: no gradient) is independent of the edited one until the loss: it runs on a
        # second CUDA stream, so its under-filled launches (low-resolution layers, ViT GEMMs with a few dozen tiles) share the SMs
        self.overlap = os.environ.get('STYLEMC_OVERLAP', '1') != '0'
        self._side = None

    # ---- pieces ----------------------------------------------------------------------------------
    def direction(self):
        """styles_direction [1, 26, 512] (find_direction.py:306-307; the tensor saved as direction_*.npz, :349-351)."""
        d = torch.zeros([1, N_STYLE_CHANNELS, synthesis.STYLE_WIDTH], dtype=torch.float32, device=self.device)
        if getattr(self, '_rows_idx', None) is None:         # device-side row indices: no host-to-device copy inside a step (graph capture)
            self._rows_idx = torch.tensor(self.rows, dtype=torch.int64, device=self.device)
        return d.index_copy_(1, self._rows_idx, self.delta)

    def load_direction(self, styles_direction):
        """Resume from a saved direction [1, 26, 512] (find_direction.py:266-270: the trainable rows are selected out of it)."""
        d = torch.as_tensor(styles_direction, dtype=torch.float32)
        if tuple(d.shape) != (1, N_STYLE_CHANNELS, synthesis.STYLE_WIDTH):
            raise RuntimeError(f'direction must be [1, {N_STYLE_CHANNELS}, {synthesis.STYLE_WIDTH}], got {tuple(d.shape)}')
        self.delta.copy_(d[:, self.rows].to(self.device))

    def seed_delta(self, scale=1e-2, seed=0):
        """Replace an all-zero delta by ``scale * N(0, 1)`` drawn from a seeded HOST generator (identical on every rank).

        At delta == 0 (find_direction.py:270, the default start) the edited and the original image are the same tensor, so
        ``tgt - src`` is exactly 0 and the reference's loss is 0/0 = NaN (clip_loss.py:27-28); it only gets off the ground where
        cuDNN's non-determinism makes the two passes differ in the last bit.  ``smc_clip_loss`` gives such a sample cos = 0 and a
        zero gradient, and the L2 gradient is 0 too, so SGD would stay at 0 forever: the loop driver calls this instead."""
        g = torch.Generator().manual_seed(seed)
        self.delta.copy_((scale * torch.randn(self.delta.shape, generator=g)).to(self.device))

    def _encode_original(self, s):
        """CLIP embeddings of the un-edited images (find_direction.py:312: no gradient), one per tower."""
        with _phase('original_branch'):
            _, original, _ = getattr(self, 'engine_original', self.engine).forward(s, self.until_k, self.noise_mode, save=False)
            u_s = resample.unprocess_fwd(original, mode=getattr(self.clips[0][1], 'preprocess', 'unprocess'))
            e_s = [model.encode_image_fwd(u_s, save=False)[0] for model, _, _ in self.clips]
            return e_s + [original] if self._use_id() else e_s

    def clear_source_cache(self):
        """Forget the cached original-image embeddings (``source_key``): call it when the styles behind a key change (a new styles_array)."""
        self._src_cache.clear()

    def _use_id(self):
        return getattr(self, 'id_loss', None) is not None and self.identity_loss_coef != 0.0

    def loss_and_grad(self, styles, global_count=None, styles_edit=None, per_sample=False, source_key=None):
        """styles [n, 26, 512] (this rank's shard, device) -> (grad [8, 512] summed over the shard, clip-loss partial sum).
        ``global_count`` is the number of seeds in the whole step (all ranks); the CLIP loss is their mean (clip_loss.py:34).
        ``styles_edit`` (optional [n, 26, 512]): the edited S of every image instead of ``styles + direction()`` (the latent mapper's delta
        differs per image); ``per_sample=True`` returns the gradient w.r.t. each image's own trainable rows, [n, 8, 512].
        ``source_key`` (hashable, optional): names this batch of styles.  The CLIP embedding of the un-edited images depends on the styles
        only, never on delta, and the reference loop draws the same batches again and again (find_direction.py:303-304 over n_epochs): the
        first step that sees a key keeps the embeddings (512 floats per image and tower), later steps with that key skip the whole
        original-image branch (a third of a step).  The caller guarantees that a key always comes with the same styles.  Not used with the
        identity term (it needs the original IMAGE) and never by bench.py (a step there does all of its work)."""
        n_total = styles.shape[0]
        count = n_total if global_count is None else global_count
        grad = torch.zeros([len(self.rows), synthesis.STYLE_WIDTH], dtype=torch.float32, device=self.device)
        part_sum = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._id_part = torch.zeros(1, dtype=torch.float32, device=self.device)
        samples = []
        direction = self.direction()
        eng = self.engine
        loss0 = self.clips[0][1]
        pre = getattr(loss0, 'preprocess', 'unprocess')    # find_direction.py:49-52, or the NADA preprocessing (clip_loss_nada.py:86-89)
        need_src = getattr(loss0, 'needs_source', True) or self._use_id()     # nada_global alone never looks at the original image
        for lo in range(0, n_total, self.micro_batch):
            s = styles[lo:lo + self.micro_batch].to(self.device, torch.float32)
            s2 = s + direction if styles_edit is None else styles_edit[lo:lo + self.micro_batch].to(self.device, torch.float32)   # find_direction.py:308
            e_s = [None] * len(self.clips)
            cache_key = (source_key, lo, s.shape[0]) if (source_key is not None and need_src and not self._use_id()) else None
            cached = self._src_cache.get(cache_key) if cache_key is not None else None
            need_run = need_src and cached is None
            if cached is not None:
                e_s = cached
            if self.overlap and need_run:
                cur = torch.cuda.current_stream(self.device)
                if self._side is None:
                    self._side = torch.cuda.Stream(self.device)
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    e_s = self._encode_original(s)
                s.record_stream(self._side)
            with _phase('synthesis_fwd'):
                _, img, saved = eng.forward(s2, self.until_k, self.noise_mode, save=True, grad_rows=self.rows)   # :309
            with _phase('unprocess_fwd'):
                u_t = resample.unprocess_fwd(img, mode=pre)                                        # :159-160
            if not self.overlap and need_run:
                e_s = self._encode_original(s)
            g224 = gscale = None
            for i, (model, loss_fn, weight) in enumerate(self.clips):
                with _phase('clip_fwd'):
                    e_t, csaved = model.encode_image_fwd(u_t, save=True)
                if self.overlap and need_run and i == 0:
                    cur.wait_stream(self._side)
                    for e in e_s:
                        e.record_stream(cur)
                if cache_key is not None and cached is None and i == 0:
                    self._src_cache[cache_key] = e_s
                part, d_t, gs = loss_fn.loss_and_grad(e_s[i], e_t, self.clip_loss_coef * weight, 1.0 / count)
                with _phase('clip_bwd'):
                    g = model.encode_image_bwd(csaved, d_t)
                del csaved
                # every tower's pixel gradient carries its own power-of-two loss scale: bring the later ones to the first one's
                g224, gscale = (g, gs) if g224 is None else (g224 + g * (gscale / gs), gscale)
                part_sum += part
            with _phase('unprocess_bwd'):
                g_img = resample.unprocess_bwd(g224, img, unscale=gscale, mode=pre)
            if self._use_id():
                with _phase('identity_loss'):       # id_loss.py:26-39 on (generated, original): adds its image gradient and its partial sum
                    part_id, g_id = self.id_loss.loss_and_grad(img, e_s[len(self.clips)], self.identity_loss_coef, 1.0 / count)
                    g_img = g_img + g_id
                    self._id_part += part_id
            with _phase('synthesis_bwd'):
                if per_sample:
                    g_sum, g_each = eng.backward(saved, g_img, self.rows, self.noise_mode, per_sample=True,
                                                  grad_lo=os.environ.get('STYLEMC_MAPPER_GRAD_LO', '1') != '0')
                    grad += g_sum
                    samples.append(g_each)
                else:
                    grad += eng.backward(saved, g_img, self.rows, self.noise_mode)
        # with the identity term the partial sums travel together: [clip, identity] (one all-reduce either way)
        parts = torch.cat([part_sum, self._id_part]) if self._use_id() else part_sum
        return (torch.cat(samples), parts) if per_sample else (grad, parts)

    def step_graph(self, styles, lr=None, global_count=None):
        """``step`` replayed from a CUDA graph: the ~500 launches of a step (two streams, the NCCL all-reduce included) are captured once per
        (shard size, global count) and replayed with one ``cudaGraphLaunch``; the S batch and the learning rate are copied into static
        device buffers first.  Worth it where a step is short (BASELINE configs[2]: 17 seeds per GPU at 256 px = 25 ms, half of it launch
        gaps); the returned tensors are the graph's static outputs, overwritten by the next replay."""
        lr = self.lr if lr is None else lr
        styles = styles.to(self.device, torch.float32)
        key = (tuple(styles.shape), global_count)
        if not hasattr(self, '_graphs'):
            self._graphs = {}
        if key not in self._graphs:
            static_s, static_lr = styles.clone(), torch.zeros(1, dtype=torch.float32, device=self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                  # warm-up outside the capture (allocator, smem opt-ins, side stream); lr 0: delta unchanged
                for _ in range(2):
                    self.step(static_s, lr=static_lr, global_count=global_count)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count
            with torch.cuda.graph(graph):
                out = self.step(static_s, lr=static_lr, global_count=global_count)
            self._graphs[key] = (graph, static_s, static_lr, out, _lib.launch_count - n0)
        graph, static_s, static_lr, out, launches = self._graphs[key]
        static_s.copy_(styles, non_blocking=True)
        static_lr.fill_(float(lr))
        graph.replay()
        _lib.launch_count += launches            # kernels of this library inside the replayed graph (bench.py's gpu_launches claim)
        return out

    def step(self, styles, lr=None, global_count=None, source_key=None):
        """One optimisation step on this rank's shard.  Returns a dict of device scalars (loss, clip_loss, l2_loss, grad_norm).
        ``source_key``: see ``loss_and_grad`` (skips the original-image branch for a batch of styles seen before)."""
        lr = self.lr if lr is None else lr
        if self.world > 1 and global_count is None:
            # shard sizes of the other ranks are unknown (ragged / empty shards): sum un-normalised, divide by the all-reduced count
            grad, part = self.loss_and_grad(styles, 1, source_key=source_key)
            grad, part = allreduce_step(grad, part, self.group, local_rows=styles.shape[0])
        else:
            grad, part = self.loss_and_grad(styles, styles.shape[0] if global_count is None else global_count, source_key=source_key)
            if self.world > 1:
                grad, part = allreduce_step(grad, part, self.group)
        numel = self.delta.numel()
        l2 = self.l2_reg_coef * self.delta.square().mean()                            # find_direction.py:190-191 (batch independent)
        clip_loss = self.clip_loss_coef * sum(w for _, _, w in self.clips) + part[:1]     # sum over towers of w * coef * (count - sum cos) / count
        id_loss = (self.identity_loss_coef + part[1:2]) if self._use_id() else None      # coef * (count - sum <f_hat, f>) / count, id_loss.py:33-39
        l2_scale = 2.0 * self.l2_reg_coef / numel
        grad_total = grad + l2_scale * self.delta[0]
        with torch.cuda.device(self.device):
            if torch.is_tensor(lr):     # device scalar (graph replay)
                _lib.call('smc_sgd_step_dev', _lib.ptr(self.delta), _lib.ptr(grad), numel, _lib.ptr(lr), 1.0, float(l2_scale), _lib.stream())
            else:
                _lib.call('smc_sgd_step', _lib.ptr(self.delta), _lib.ptr(grad), numel, float(lr), 1.0, float(l2_scale), _lib.stream())   # :339
        out = dict(loss=clip_loss + l2, clip_loss=clip_loss, l2_loss=l2, grad=grad_total, grad_norm=grad_total.norm())
        if id_loss is not None:
            out['identity_loss'] = id_loss
            out['loss'] = out['loss'] + id_loss                                       # find_direction.py:193
        return out
