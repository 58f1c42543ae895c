"""CPU, world_size 2 over gloo: the data-parallel exchange of the step (seed shards -> one all-reduce of the [8, 512] gradient and
the loss partial sums) reproduces the full-batch result, also for ragged shards.  The per-shard loss/gradient here comes from
the CPU oracle (the CUDA path needs a GPU); what is under test is direction.shard_rows + direction.allreduce_step."""
import os
import socket

import torch
import torch.multiprocessing as mp

from oracle import direction as o_dir
from oracle import synthesis as o_syn


class TinyClip:
    """Cheap stand-in for the CLIP towers (a fixed random linear map of a 16x16 pooled image)."""

    def __init__(self):
        g = torch.Generator().manual_seed(0)
        self.w = torch.randn(3 * 16 * 16, 512, generator=g) * 0.05
        self.t = torch.randn(2, 512, generator=g)

    def encode_image(self, x):
        return torch.nn.functional.adaptive_avg_pool2d(x, 16).flatten(1) @ self.w

    def encode_text(self, tok):
        return self.t[int(tok[0, 1]) % 2][None]


def shard_loss_and_grad(G, shapes, loss_fn, S, delta, count):
    """What DirectionFinder.loss_and_grad returns on a shard: gradient and -coef/count * sum cos, both divided by the GLOBAL count."""
    delta = delta.clone().requires_grad_(True)
    direction = torch.zeros(1, 26, 512).index_put((torch.tensor([0]).view(1, 1), torch.tensor(o_dir.S_TRAINABLE_ROWS).view(1, -1)), delta)
    _, img = o_syn.generate_image(G, 100, S + direction, shapes)
    with torch.no_grad():
        _, orig = o_syn.generate_image(G, 100, S, shapes)
    n = S.shape[0]
    part = (loss_fn(o_dir.unprocess(orig), o_dir.unprocess(img)) - 1.0) * n / count      # = -sum cos / count
    grad, = torch.autograd.grad(part, delta)
    return grad[0], part.detach().reshape(1)


def make():
    G = o_syn.make_generator(64, seed=1, channel_base=1024, channel_max=512)
    ws = torch.randn(5, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(2))
    S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
    tok = torch.zeros(1, 77, dtype=torch.int64)
    pos, neg = tok.clone(), tok.clone()
    pos[0, 1], neg[0, 1] = 0, 1
    loss_fn = o_dir.CLIPLoss(TinyClip(), pos, neg)
    delta = 0.1 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(4))
    return G, shapes, loss_fn, S, delta


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.distributed.init_process_group('gloo', rank=rank, world_size=world)
    from stylemc_b200 import direction
    torch.set_num_threads(2)
    G, shapes, loss_fn, S, delta = make()
    lo, hi = direction.shard_rows(S.shape[0], rank, world)              # 5 seeds over 2 ranks: 3 + 2 (ragged)
    grad, part = shard_loss_and_grad(G, shapes, loss_fn, S[lo:hi], delta, S.shape[0])
    # global count unknown to the ranks (DirectionFinder.step without global_count): un-normalised sums + the row count in the same buffer
    grad_u, part_u = direction.allreduce_step(grad * S.shape[0], part * S.shape[0], torch.distributed.group.WORLD, local_rows=hi - lo)
    grad, part = direction.allreduce_step(grad, part, torch.distributed.group.WORLD)
    if rank == 0:
        torch.save(dict(grad=grad, part=part, span=(lo, hi), grad_u=grad_u, part_u=part_u), out)
    torch.distributed.destroy_process_group()


def test_two_rank_allreduce_equals_full_batch(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'rank0.pt')
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    G, shapes, loss_fn, S, delta = make()
    grad, part = shard_loss_and_grad(G, shapes, loss_fn, S, delta, S.shape[0])
    assert got['span'] == (0, 3)
    assert ((got['grad'] - grad).norm() / grad.norm()).item() <= 1e-5
    assert abs(got['part'].item() - part.item()) <= 1e-6
    assert ((got['grad_u'] - grad).norm() / grad.norm()).item() <= 1e-5        # ragged 3 + 2 shards, count all-reduced with the data
    assert abs(got['part_u'].item() - part.item()) <= 1e-6


def loop_worker(rank, world, port, outdir):
    """npzio.find_direction on two ranks with a stand-in step: every rank must draw the same batch and take disjoint rows of it."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.distributed.init_process_group('gloo', rank=rank, world_size=world)
    from stylemc_b200 import direction, npzio as io
    group = torch.distributed.group.WORLD
    S = torch.arange(11, dtype=torch.float32).view(11, 1, 1).expand(11, 26, 512).contiguous()      # row id in every element
    f = object.__new__(direction.DirectionFinder)
    f.device, f.rows, f.lr, f.world, f.group = torch.device('cpu'), direction.S_TRAINABLE_SPACE_CHANNELS, 1.0, world, group
    f.delta = torch.zeros(1, 8, 512)
    log = []

    def step(styles, lr=None, global_count=None, source_key=None):
        assert source_key is not None and source_key[2] - source_key[1] == styles.shape[0]     # (batch index, first row, last row) of this shard
        ids = styles[:, 0, 0].tolist()
        rows = torch.zeros(11)
        rows[[int(i) for i in ids]] = 1
        torch.distributed.all_reduce(rows, group=group)                     # union of the shards of this iteration
        assert rows.max().item() == 1, 'two ranks processed the same seed'
        assert rows.sum().item() == global_count, 'the shards do not cover the batch every rank was told about'
        log.append(rows.nonzero().flatten().tolist())
        f.delta += rows.sum()
        return dict(loss=torch.tensor(0.0))
    f.step = step
    final = io.find_direction(f, S, batch_size=4, n_epochs=2, outdir=outdir, text_prompt='p', seed=5, zero_init='keep')
    if rank == 0:
        torch.save(dict(log=log, final=final), os.path.join(outdir, 'log.pt'))
    else:
        assert not os.path.exists(os.path.join(outdir, 'rank1_wrote_something'))
    torch.distributed.destroy_process_group()


def test_two_rank_loop_draws_one_batch_and_shards_it(tmp_path):
    import numpy as np
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(loop_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(tmp_path / 'log.pt')
    rng = np.random.RandomState(5)
    want = [list(range(4 * i, min(4 * i + 4, 11))) for i in (rng.randint(0, 3) for _ in range(6))]
    assert got['log'] == want                                                 # contiguous batches (find_direction.py:303-304), incl. the ragged one
    assert sorted(os.listdir(tmp_path)) == ['direction_p.npz', 'log.pt']      # only rank 0 writes
    assert got['final'][0, 2, 0].item() == sum(len(w) for w in want)
