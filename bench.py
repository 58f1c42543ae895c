#!/usr/bin/env python
"""Headline benchmark: find_direction synth + CLIP fwd/bwd images/sec @1024px (BASELINE.json metric, configs[3]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on the host cores (CPU oracle)

One "step" = one find_direction optimisation step on a batch of seeds: two synthesis forwards, two unprocess, two CLIP ViT-B/32
forwards, the directional loss, the backward pass to the trainable S rows, (all-reduce,) SGD.  One image = one seed through one
step.  ``value`` is measured with the S batch resident in HBM; ``e2e`` runs the same step through the public API with the S batch
coming from pinned host memory and the loss read back every step.  Prints ONE JSON line on rank 0.
"""
import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'find_direction synth+CLIP fwd/bwd images/sec @1024px'
POS_BODY = [320, 1125, 539, 320, 1710, 539, 320, 14387, 2308]       # placeholder token ids (no BPE vocabulary offline)
NEG_BODY = [320, 1125, 539, 320, 1710, 539, 320, 25173, 786, 1237]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--resolution', type=int, default=1024)
    ap.add_argument('--batch', type=int, default=64, help='seeds per GPU per step (BASELINE configs[3]: 64/GPU)')
    ap.add_argument('--micro-batch', type=int, default=64, help='seeds per pass through the network inside a step')
    ap.add_argument('--precision', default='x3p', choices=['x1', 'mixed', 'x3', 'x3p'])
    ap.add_argument('--lr', type=float, default=0.05, help='SGD learning rate (reference default 1.5 is tuned for trained weights; random-init nets diverge with it)')
    ap.add_argument('--cpu-sample', type=int, default=1, help='seeds per step of the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='find_direction', choices=['find_direction', 'generate_fromS'],
                    help="generate_fromS: forward-only render (BASELINE configs[1], batch 32 at 1024 px); not the headline metric")
    ap.add_argument('--clip-type', default='small', choices=['small', 'double'],
                    help="double: ViT-B/32 + 0.5 * ViT-B/16 (the reference CLI's default, find_direction.py:216); not the headline metric (BASELINE fixes ViT-B/32)")
    ap.add_argument('--global-seeds', type=int, default=0,
                    help='STRONG scaling: total seeds per step, sharded over the ranks with direction.shard_rows (ragged shards allowed; BASELINE configs[2]: '
                         '--resolution 256 --global-seeds 129).  0 = weak scaling with --batch seeds per GPU')
    ap.add_argument('--gpu-library-baseline', action='store_true',
                    help="also time the reference's GPU formulation (oracle modules = ATen / cuDNN / cuBLAS fp32 on cuda:0, TF32 off and on) for one step; "
                         'reported as gpu_library_baseline next to cpu_baseline (BASELINE.md section 3, second bar)')
    ap.add_argument('--cuda-graph', type=int, default=-1, help='replay the step from a CUDA graph (DirectionFinder.step_graph): 1 on, 0 off, '
                    '-1 = on for single-GPU --global-seeds runs, off otherwise')
    ap.add_argument('--profile-step', action='store_true', help='run warm-up, then ONE step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=p.get('bf16_tflops_sustained', p.get('bf16_tflops')), hbm=p.get('hbm_gbs'), src='measured (MEASURED_PEAKS.json, sustained bf16)')
    return dict(tflops=1400.0, hbm=6650.0, src='fallback (B200_PROFILING.md)')


def synth_flops_per_image(blocks, until_k):
    """Algorithmic forward FLOPs of the synthesis network per image (2*9*Cin*Cout*Hin^2 for conv0,
    2*9*C^2*H^2 for conv1, 2*C*3*H^2 for ToRGB); the dgrad backward costs the same minus b4."""
    total = 0
    for k, b in enumerate(blocks[:until_k + 1]):
        r = b.resolution
        if b.conv0 is not None:
            total += 2 * 9 * b.conv0.cin * b.conv0.cout * (r // 2) ** 2
        total += 2 * 9 * b.conv1.cin * b.conv1.cout * r * r + 2 * 3 * b.torgb.cin * r * r
    return total


# DRAM bytes per launch of the dominant shape, from the committed ncu capture (key = n_img, H, W, Cin, Cout, taps); None when not captured
TOP_KERNEL_DRAM_BYTES = {(64, 1024, 1024, 32, 32, 9): 12.99e9}   # mean of the 3 launches per step: 10.49 (no-grad fwd), 15.42 (grad fwd), 13.07 GB (dgrad, hi-only gradient planes)
TRAFFIC_SOURCE = 'ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r02g_hconv_launches.md (rows 14, 78, 177) and r02f_top_kernel.md'

VIT_FLOPS_FWD = 2 * (49 * 3072 * 768 + 12 * 50 * (768 * 2304 + 768 * 768 + 2 * 768 * 3072) + 12 * 12 * 2 * 50 * 50 * 64)   # per image
VIT_B16_FLOPS_FWD = 2 * (196 * 768 * 768 + 12 * 197 * (768 * 2304 + 768 * 768 + 2 * 768 * 3072) + 12 * 12 * 2 * 197 * 197 * 64)   # 35.2 GFLOP


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.path = tempfile.mktemp(suffix='.csv')
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], None, set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        busy = sorted(sm)[len(sm) // 4:] if sm else []       # drop the idle tail of the sampling window
        return dict(sm_mhz=statistics.median(busy) if busy else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def hooked_step(fn):
    """Run ``fn`` once with every smc_igemm launch bracketed by CUDA events on the launching stream.  Returns (records, step ms);
    a record = (event0, event1, algorithmic FLOPs, shape key, algorithmic HBM bytes)."""
    from stylemc_b200 import _lib
    records = []

    @contextlib.contextmanager
    def hook(d, alg_taps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        planes_a = max(1, d.ntaps // max(1, alg_taps) - 1)                       # x3: hi + lo planes of A are read
        e = d.epi
        out_b = (4 if e.out_f32 else 0) + sum(2 for q in (e.out_hi, e.out_lo, e.out_raw, e.out_raw_lo) if q)
        nbytes = float(d.n_img) * d.H * d.W * (2 * planes_a * d.C + out_b * d.n_out * max(1, d.nprob))   # algorithmic HBM bytes: A once, outputs once
        terms = max(1, d.ntaps // max(1, alg_taps))            # fp16 MMAs per algorithmic product: 3 (hi/lo split), 2 (hi-only A operand), 1
        records.append((e0, e1, 2.0 * d.n_img * d.H * d.W * d.n_out * d.C * alg_taps, (d.n_img, d.H, d.W, d.C, d.n_out, alg_taps), nbytes, terms))

    _lib.igemm_hook = hook
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    try:
        fn()
    finally:
        _lib.igemm_hook = None
    t1.record()
    torch.cuda.synchronize()
    return records, t0.elapsed_time(t1)


def summarize_records(records):
    """(summed ms, summed FLOPs, key of the shape with the largest summed time, [ms, flops, launches, bytes, MMA flops] of that shape,
    summed MMA flops).  MMA flops = algorithmic flops x the fp16 MMAs issued per product in that launch."""
    ig_ms = sum(r[0].elapsed_time(r[1]) for r in records)
    ig_flops = sum(r[2] for r in records)
    ig_mma = sum(r[2] * r[5] for r in records)
    shapes = {}
    for e0, e1, fl, key, nb, terms in records:
        t = shapes.setdefault(key, [0.0, 0.0, 0, 0.0, 0.0])
        t[0] += e0.elapsed_time(e1); t[1] += fl; t[2] += 1; t[3] += nb; t[4] += fl * terms
    top_key, top = max(shapes.items(), key=lambda kv: kv[1][0])
    return ig_ms, ig_flops, top_key, top, ig_mma


def top_kernel_roofline(pk, top_key, top, step_ms):
    """roofline object of the heaviest smc_igemm shape: SURVEY.md 8d's max(flops / peak_flops, bytes / peak_bw) of the ALGORITHMIC work."""
    top_tf = top[1] / (top[0] / 1e3) / 1e12
    top_gbs = top[3] / (top[0] / 1e3) / 1e9
    bound = 'hbm' if top[3] / (pk['hbm'] * 1e9) >= top[1] / (pk['tflops'] * 1e12) else 'tensor'
    return {'bound': bound,
            'kernel': f'hconv_kernel (tcgen05 halo-tile implicit GEMM), heaviest shape of the step: n={top_key[0]} {top_key[1]}x{top_key[2]} '
                      f'Cin={top_key[3]} Cout={top_key[4]} taps={top_key[5]}',
            'achieved': round(top_gbs if bound == 'hbm' else top_tf, 2), 'peak': pk['hbm'] if bound == 'hbm' else pk['tflops'],
            'unit': 'GB/s' if bound == 'hbm' else 'TFLOP/s',
            'frac': round(top_gbs / pk['hbm'] if bound == 'hbm' else top_tf / pk['tflops'], 4),
            'algorithmic_tflops': round(top_tf, 2), 'algorithmic_gbs': round(top_gbs, 1),
            'hbm_frac': round(top_gbs / pk['hbm'], 4), 'tensor_frac': round(top_tf / pk['tflops'], 4),
            'algorithmic_flop_per_byte': round(top[1] / top[3], 1), 'ridge_flop_per_byte': round(pk['tflops'] * 1e3 / pk['hbm'], 1),
            'algorithmic_gbyte_per_launch': round(top[3] / top[2] / 1e9, 2),
            'launches_per_step': top[2], 'avg_launch_ms': round(top[0] / top[2], 4),
            'algorithmic_gflop_per_launch': round(top[1] / top[2] / 1e9, 2),
            'mma_per_product': round(top[4] / top[1], 2),                 # mean over the shape's launches: 3 forward (hi/lo split), 2 backward
            'tensor_pipe_achieved': round(top[4] / (top[0] / 1e3) / 1e12, 2),
            'tensor_pipe_frac': round(top[4] / (top[0] / 1e3) / 1e12 / pk['tflops'], 4),
            'traffic': TOP_KERNEL_DRAM_BYTES.get(top_key), 'traffic_source': TRAFFIC_SOURCE if top_key in TOP_KERNEL_DRAM_BYTES else None,
            'peak_source': pk['src'], 'share_of_step': round(top[0] / step_ms, 3)}


# ------------------------------------------------------------------------------------------------------------------
def run_ours(a):
    from stylemc_b200 import _lib, clip, direction, networks, utils
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl ours needs a CUDA device (there is no CPU path)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    group = None
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=dev)
        group = torch.distributed.group.WORLD
    assert world == a.gpus, f'--gpus {a.gpus} but WORLD_SIZE={world} (launch N > 1 under torch.distributed.run)'

    G = networks.make_generator(a.resolution, seed=0)
    # S pool: W -> S through the affine layers once (w_s_converter.py:75-82); two batches per rank alternate between steps
    strong = a.global_seeds > 0
    if strong:
        # every rank builds the same global pool and takes its rows of each batch (contiguous, sizes differ by at most one)
        lo, hi = direction.shard_rows(a.global_seeds, rank, world)
        local_n, global_n = hi - lo, a.global_seeds
        ws = torch.randn(2 * global_n, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(1000))
        styles_host, shapes = utils.get_styles(G, ws, utils.split_ws(G, ws), 'cpu')
        styles_host = torch.cat([styles_host[lo:hi], styles_host[global_n + lo:global_n + hi]]).contiguous()
    else:
        local_n, global_n = a.batch, a.batch * world
        ws = torch.randn(2 * a.batch, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(1000 + rank))
        styles_host, shapes = utils.get_styles(G, ws, utils.split_ws(G, ws), 'cpu')
    styles_host = styles_host.pin_memory()
    clip_precision = 'x3p' if a.precision == 'x3p' else ('x3' if a.precision != 'x1' else 'x1')
    model = clip.CLIPModel(clip.random_params(seed=0), dev, precision=clip_precision)
    if a.clip_type == 'double':
        model = (model, clip.CLIPModel(clip.random_params(seed=1, cfg=clip.VIT_B16), dev, precision=clip_precision, cfg=clip.VIT_B16))
    finder = direction.DirectionFinder(G, model, clip.placeholder_tokens(POS_BODY), clip.placeholder_tokens(NEG_BODY), a.resolution, device=dev,
                                       precision=a.precision, micro_batch=a.micro_batch, process_group=group, learning_rate=a.lr)
    # delta == 0 makes edited == original and the directional loss 0/0 (NaN in the reference): start from a small random direction
    finder.delta.copy_(0.05 * torch.randn(finder.delta.shape, generator=torch.Generator().manual_seed(7)).to(dev))
    styles_dev = styles_host.to(dev)
    total_steps = a.steps + a.warmup

    def sync():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item()

    # graph replay pays where a step is short and there is no per-step host sync around an NCCL collective: on one GPU it is +3.7 % at 17 seeds;
    # on 8 GPUs +1.9 % resident, but the e2e loop (loss read back every step) ran at HALF speed with the all-reduce inside the graph
    use_graph = a.cuda_graph == 1 or (a.cuda_graph == -1 and strong and world == 1)
    do_step = finder.step_graph if use_graph else finder.step

    def step_resident(i):
        lr = direction.cosine_lr(finder.lr, i + 1, total_steps)
        do_step(styles_dev[(i % 2) * local_n:(i % 2 + 1) * local_n], lr=lr, global_count=global_n)

    losses = []

    def step_e2e(i):
        lr = direction.cosine_lr(finder.lr, i + 1, total_steps)
        s = styles_host[(i % 2) * local_n:(i % 2 + 1) * local_n].to(dev, non_blocking=True)      # H2D from pinned memory
        losses.append(do_step(s, lr=lr, global_count=global_n)['loss'].item())                    # D2H of the step's loss

    for i in range(a.warmup):
        step_resident(i)
    if a.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident(0)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = _lib.launch_count
    ms = timed(step_resident, a.steps)
    launches = _lib.launch_count - n0
    ms_e2e = timed(step_e2e, a.steps)
    clocks = sampler.stop() if sampler else None

    # ---- roofline of the dominant kernel family (smc_igemm): every launch of ONE extra step bracketed by CUDA events
    overlap, finder.overlap = finder.overlap, False       # per-launch timing: keep the two image branches on one stream for this step
    records, step_ms_hooked = hooked_step(lambda: finder.step(styles_dev[:local_n], lr=0.0, global_count=global_n))
    finder.overlap = overlap
    ig_ms, ig_flops, top_key, top, ig_mma = summarize_records(records)

    if rank != 0:
        return
    pk = peaks()
    eng = finder.engine
    sf = synth_flops_per_image(eng.blocks, finder.until_k)
    alg_flops_img = 3 * sf + 3 * VIT_FLOPS_FWD                # 2 fwd + dgrad bwd of synthesis, 2 fwd + input-grad bwd of the ViT
    if a.clip_type == 'double':
        alg_flops_img += 3 * VIT_B16_FLOPS_FWD
    imgs = global_n * a.steps
    value = imgs / (ms / 1e3)
    e2e = imgs / (ms_e2e / 1e3)
    achieved = ig_flops / (ig_ms / 1e3) / 1e12
    # every layer is reported against max(flops / peak_flops, bytes / peak_bw) of its ALGORITHMIC work; the 32/64-channel convs
    # sit below the ridge (b1024.conv1: ~100-144 FLOP/B against ~216 FLOP/B): their bound is HBM.  tensor_pipe_frac is the same launch
    # seen from the tensor pipe, which executes mma_per_product fp16 MMAs per algorithmic product in split precision.
    roof = top_kernel_roofline(pk, top_key, top, step_ms_hooked)
    roof.update({'family': {'kernel': 'every smc_igemm launch of one step (hconv_kernel + igemm_kernel: convs, dgrads, CLIP linears)',
                            'achieved': round(achieved, 2), 'frac': round(achieved / pk['tflops'], 4),
                            'mma_per_product': round(ig_mma / ig_flops, 2),
                            'tensor_pipe_frac': round(ig_mma / (ig_ms / 1e3) / 1e12 / pk['tflops'], 4),
                            'launches_per_step': len(records), 'share_of_step': round(ig_ms / step_ms_hooked, 3)},
                 'algorithmic_gflop_per_image': round(alg_flops_img / 1e9, 1),
                 'step_algorithmic_tflops': round(alg_flops_img * global_n / world / (ms / a.steps / 1e3) / 1e12, 2)})
    out = {
        'metric': METRIC, 'value': round(value, 3), 'unit': 'images/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
        'ms_per_step': round(ms / a.steps, 3), 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak', 'vs_baseline': None,
        'dtype': 'f16 hi+lo split operands (3 MMAs per product forward, 2 backward), f32 accumulate' if a.precision != 'x1' else 'f16 operands, f32 accumulate',
        'data': 'synthetic (random-init StyleGAN2 config-f + random-init CLIP ViT-B/32, S from randn W)',
        'config': {'workload': (f'find_direction {a.resolution}px, {global_n} seeds per step sharded over {world} GPU(s) (BASELINE configs[2] when 256px / 129), '
                                f'fwd+bwd to delta-S [1,8,512]') if strong else
                               f'find_direction {a.resolution}px, batch {a.batch}/GPU (BASELINE configs[3]), fwd+bwd to delta-S [1,8,512]',
                   'resolution': a.resolution, 'batch_per_gpu': local_n, 'global_batch': global_n, 'micro_batch': a.micro_batch,
                   'clip_type': a.clip_type + (' (ViT-B/32 + 0.5 * ViT-B/16: NOT the headline configuration)' if a.clip_type == 'double' else ' (ViT-B/32)'),
                   'precision': a.precision, 'learning_rate': a.lr, 'cuda_graph': bool(use_graph), 'parallelism': f'dp{world} (seed shards; all-reduce of the 16 KiB gradient)',
                   'l2_flush': 'none needed: each step streams >10 GB of activations, far larger than the 126 MB L2'},
        'clocks': clocks,
        'e2e': {'value': round(e2e, 3), 'unit': 'images/s', 'ms_per_step': round(ms_e2e / a.steps, 3),
                'h2d_bytes_per_step': local_n * 26 * 512 * 4, 'd2h_bytes_per_step': 4, 'losses': [round(v, 6) for v in losses]},
        'gpu_launches': launches,
        # dominant kernel = the smc_igemm shape with the largest summed time in one step (a conv on csrc/hconv.cu).  `achieved` counts
        # ALGORITHMIC FLOPs (2 * pixels * taps * Cin * Cout per launch); in split precision every algorithmic product
        # costs three fp16 MMAs, so the tensor pipe does mma_per_product x that work (tensor_pipe_*).
        'roofline': roof,
    }
    if world == 1 and not a.no_cpu_baseline:
        out['cpu_baseline'] = cpu_reference(a, steps=2, warmup=1)
    if world == 1 and a.gpu_library_baseline:
        del finder, model, styles_dev
        torch.cuda.empty_cache()
        out['gpu_library_baseline'] = gpu_library_reference(a, dev)
    print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference(a, steps, warmup):
    """The reference algorithm (CPU oracle: restated network modules + the reference's op semantics, fp32, oneDNN) on a bounded
    sample of the same workload: ``cpu_sample`` seeds per step at the benchmark resolution, all host threads."""
    from oracle import direction as o_dir
    from oracle import synthesis as o_syn
    from oracle import vit as o_vit
    torch.set_num_threads(os.cpu_count())
    G = o_syn.make_generator(a.resolution, seed=0)
    ws = torch.randn(a.cpu_sample, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(1000))
    S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
    model = o_vit.CLIP(seed=0)
    loss_fn = o_dir.CLIPLoss(model, o_vit.synthetic_tokens('pos'), o_vit.synthetic_tokens('neg'))
    delta = 0.05 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(7))
    until_k = o_dir.RESOLUTION_TO_K.get(a.resolution, len(shapes) - 1)
    times = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        r = o_dir.direction_step(G, shapes, loss_fn, S, delta, until_k)
        delta = o_dir.sgd_update(delta, r['grad'], a.lr)
        times.append(time.perf_counter() - t)
    sec = sum(times[warmup:]) / steps
    return {'value': round(a.cpu_sample / sec, 4), 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{a.cpu_sample} seed(s) per step at {a.resolution}px, {steps} timed steps after {warmup} warm-up, fp32 oneDNN',
            'seconds_per_step': round(sec, 3)}


def gpu_library_reference(a, dev, batches=(8, 4, 2, 1)):
    """BASELINE.md section 3 "second bar": the reference's GPU FORMULATION -- the oracle's modules (same op sequence as the reference:
    F.conv2d / F.conv_transpose2d, upfirdn2d / bias_act as ATen ops, autograd, CLIP as torch matmuls) moved to the device, i.e.
    cuDNN / cuBLAS / ATen fp32 -- one find_direction step at the bench resolution, TF32 off and on, at the largest batch of
    ``batches`` that fits.  The reference's own JIT-compiled plugins cannot be built on the GPU box (no reference tree there); their
    ATen equivalents are what ``impl='ref'`` runs.  A reported baseline (never the --impl reference arm)."""
    from oracle import direction as o_dir
    from oracle import synthesis as o_syn
    from oracle import vit as o_vit
    G = o_syn.make_generator(a.resolution, seed=0).to(dev)
    params = {k: v.to(dev) for k, v in o_vit.random_clip_params(seed=0).items()}
    model = o_vit.CLIP(params)
    loss_fn = o_dir.CLIPLoss(model, o_vit.synthetic_tokens('pos').to(dev), o_vit.synthetic_tokens('neg').to(dev))
    until_k = o_dir.RESOLUTION_TO_K.get(a.resolution, 100)
    delta = (0.05 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(7))).to(dev)
    res = {'kind': 'oracle modules on cuda:0 (ATen / cuDNN / cuBLAS fp32 + autograd): the reference formulation, not its JIT plugins', 'unit': 'images/s'}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    ws = torch.randn(max(batches), G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(1000)).to(dev)
    S_all, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))        # once: get_styles turns the affine layers into Identity
    try:
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = tf32
            for n in batches:
                try:
                    S = S_all[:n]
                    times = []
                    for i in range(4):
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        r = o_dir.direction_step(G, shapes, loss_fn, S, delta, until_k)
                        e1.record()
                        torch.cuda.synchronize()
                        times.append(e0.elapsed_time(e1))
                    ms = statistics.median(times[1:])
                    res['tf32' if tf32 else 'fp32'] = {'value': round(n / (ms / 1e3), 2), 'ms_per_step': round(ms, 2), 'batch': n,
                                                        'loss': round(float(r['loss']), 6)}
                    break
                except torch.cuda.OutOfMemoryError:
                    torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return res


def run_reference(a):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    b = cpu_reference(a, steps=a.steps, warmup=a.warmup)
    out = {'impl': 'reference', 'metric': METRIC, 'value': b['value'], 'unit': 'images/s', 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
           'ms_per_step': round(b['seconds_per_step'] * 1e3, 1), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
           'data': 'synthetic (same random-init networks and S as the CUDA arm)',
           'config': {'workload': f'find_direction {a.resolution}px, {a.cpu_sample} seed(s) per step on the host cores (bounded sample of configs[3])',
                      'resolution': a.resolution, 'batch_per_step': a.cpu_sample},
           'cpu_baseline': b, 'e2e': {'value': b['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out), flush=True)


def run_generate(a):
    """BASELINE configs[1]: generate_fromS forward-only synthesis from S at 1024 px, batch 32, one B200, no CLIP.  A step renders the
    original and the edited image of 32 style vectors (generate_fromS.py:137-207) into the uint8 canvas; images/s counts both.
    ``value``: styles resident in HBM, canvas left on the device.  ``e2e``: styles from pinned host memory every step and the uint8 canvas
    copied back to pinned host memory (what the reference hands to PIL, generate_fromS.py:175)."""
    from stylemc_b200 import _lib, generate, networks, utils
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    batch = 32
    G = networks.make_generator(a.resolution, seed=0)
    ws = torch.randn(batch, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(1000))
    styles, _ = utils.get_styles(G, ws, utils.split_ws(G, ws), 'cpu')
    styles = styles.pin_memory()
    direction = torch.zeros(1, 26, 512)
    direction[:, [2, 3, 5, 6, 8, 9, 11, 12]] = 0.1 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(7))
    direction = direction.to(dev)
    styles_dev = styles.to(dev)
    canvas_host = torch.empty([batch, a.resolution, 2 * a.resolution, 3], dtype=torch.uint8).pin_memory()
    render = lambda s: generate.generate_fromS(G, s, direction, 1.0, device=dev, precision=a.precision)

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def step_e2e():
        out = render(styles.to(dev, non_blocking=True))
        canvas_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()           # the canvas is on the host before the next step starts

    for _ in range(a.warmup):
        out = render(styles_dev)
    sampler = ClockSampler(dev.index)
    n0 = _lib.launch_count
    ms = timed(lambda: render(styles_dev))
    launches = _lib.launch_count - n0
    ms_e2e = timed(step_e2e)
    clocks = sampler.stop()
    records, step_ms = hooked_step(lambda: render(styles_dev))
    ig_ms, ig_flops, top_key, top, ig_mma = summarize_records(records)
    pk = peaks()
    roof = top_kernel_roofline(pk, top_key, top, step_ms)
    roof['family'] = {'kernel': 'every smc_igemm launch of one step', 'achieved': round(ig_flops / (ig_ms / 1e3) / 1e12, 2),
                      'tensor_pipe_frac': round(ig_mma / (ig_ms / 1e3) / 1e12 / pk['tflops'], 4), 'launches_per_step': len(records),
                      'share_of_step': round(ig_ms / step_ms, 3)}
    imgs = 2 * batch * a.steps
    print(json.dumps({'metric': 'generate_fromS forward-only synthesis images/sec @1024px', 'value': round(imgs / (ms / 1e3), 2),
                      'unit': 'images/s', 'n_gpus': 1, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': round(ms / a.steps, 3),
                      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                      'dtype': 'f16x3 operands (hi+lo split), f32 accumulate' if a.precision != 'x1' else 'f16 operands, f32 accumulate', 'data': 'synthetic',
                      'config': {'workload': f'generate_fromS {a.resolution}px, 32 styles per step, original + edited -> uint8 canvas (BASELINE configs[1])',
                                 'precision': a.precision, 'l2_flush': 'none needed: a step streams several GB of activations'},
                      'clocks': clocks,
                      'e2e': {'value': round(imgs / (ms_e2e / 1e3), 2), 'unit': 'images/s', 'ms_per_step': round(ms_e2e / a.steps, 3),
                              'h2d_bytes_per_step': batch * 26 * 512 * 4, 'd2h_bytes_per_step': canvas_host.numel()},
                      'gpu_launches': launches, 'roofline': roof, 'canvas_shape': list(out.shape)}), flush=True)


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'generate_fromS':
        run_generate(args)
    else:
        run_ours(args)
