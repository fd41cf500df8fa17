"""world_size-2 gloo test of the sharding contract the multi-GPU learner relies on:
SUM over ranks of (shard gradient bucket + shard loss sums) == full-batch values."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class TinyNet(torch.nn.Module):
    """No BatchNorm: per-shard statistics would (legitimately) differ from full-batch ones."""

    def __init__(self):
        super().__init__()
        self.body = torch.nn.Linear(27, 32)
        self.p = torch.nn.Linear(32, 9)
        self.v = torch.nn.Linear(32, 1)

    def forward(self, x, hidden=None):
        h = torch.relu(self.body(x.flatten(1)))
        return {'policy': self.p(h), 'value': torch.tanh(self.v(h))}


ARGS = {'turn_based_training': True, 'observation': False, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0,
        'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1, 'policy_target': 'UPGO', 'value_target': 'VTRACE'}


def bucket_for(batch):
    from oracle.torch_learner import loss_from_raw
    torch.manual_seed(0)
    net = TinyNet()
    B, T, Pa = batch['action'].shape[:3]
    outs = net(batch['observation'].flatten(0, 2))
    raw = {k: v.unflatten(0, (B, T, Pa)) for k, v in outs.items()}
    losses, dcnt = loss_from_raw(raw, batch, ARGS)
    losses['total'].backward()
    grads = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    tail = torch.stack([losses['p'], losses['v'], torch.zeros(()), losses['ent'], losses['total'], dcnt]).detach()
    return torch.cat([grads, tail])


def worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from handyrl_b200 import multigpu as hdist
    from handyrl_b200.synthetic import synthetic_batch
    r, w, _ = hdist.init_from_env('gloo')
    assert (r, w) == (rank, world)
    full = synthetic_batch(11, 8, 2, 9, seed=3)          # 11 windows: uneven split 6 + 5
    shard = hdist.shard_batch(full, rank, world)
    lo, hi = hdist.shard_bounds(11, rank, world)
    assert shard['action'].shape[0] == hi - lo
    flat = bucket_for(shard)
    hdist.allreduce_sum_(flat)
    if rank == 0:
        ref = bucket_for(full)
        out.put((flat.numpy(), ref.numpy()))
    torch.distributed.destroy_process_group()


def test_sum_allreduce_of_shards_equals_full_batch():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got, ref = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_allclose(got[:-6], ref[:-6], rtol=1e-4, atol=1e-5)      # gradients
    np.testing.assert_allclose(got[-6:], ref[-6:], rtol=1e-5, atol=1e-4)      # p, v, r, ent, total, dcnt
    assert got[-1] == ref[-1]                                                  # dcnt is an exact count


def test_shard_bounds_cover_the_batch():
    from handyrl_b200.multigpu import shard_bounds
    for B in (1, 7, 512, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
