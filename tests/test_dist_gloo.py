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
    grad, part = direction.allreduce_step(grad, part, torch.distributed.group.WORLD)
    if rank == 0:
        torch.save(dict(grad=grad, part=part, span=(lo, hi)), out)
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
