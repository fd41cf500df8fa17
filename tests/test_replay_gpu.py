"""GPU parity of the replay gather/pad kernel (hrl_gather_pad) against the reference's make_batch golden."""
import os
import pickle
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN, 'batch_cases.pkl'), 'rb') as f:
    BATCH_CASES = pickle.load(f)


def windows_for(case, replay, handles):
    from handyrl_b200.replay import WINDOW_DTYPE
    eps = case['episodes']
    win = np.zeros(len(case['selected']), WINDOW_DTYPE)
    solo = not case['args']['turn_based_training']
    random.seed(9)   # the golden generator seeded `random` like this before the reference's make_batch
    for b, sel in enumerate(case['selected']):
        cs = case['args']['compress_steps']
        idx = next(i for i, ep in enumerate(eps)
                   if ep['steps'] == sel['total'] and ep['moment'][sel['base'] // cs:sel['base'] // cs + len(sel['moment'])] == sel['moment'])
        h = handles[idx]
        player = random.choice(range(replay.Ps)) if solo else 0
        win[b] = (h.first_step, sel['start'], sel['end'], sel['train_start'], sel['total'], h.outcome_row, player)
    return win


@pytest.mark.parametrize('name', sorted(BATCH_CASES))
def test_gather_pad_matches_reference_make_batch(name):
    from handyrl_b200.replay import DeviceReplay
    from handyrl_b200.batch import tree_leaves
    case = BATCH_CASES[name]
    replay = DeviceReplay(capacity_steps=4096, max_episodes=64)
    handles = [replay.add(ep) for ep in case['episodes']]
    win = windows_for(case, replay, handles)
    out = replay.gather(win, case['args'])
    torch.cuda.synchronize()
    ref = case['batch']
    obs = replay.split_observation(out['observation'])
    for g, r in zip(tree_leaves(obs), tree_leaves(ref['observation'])):
        assert tuple(g.shape) == r.shape
        assert np.array_equal(g.cpu().numpy(), r.astype(np.float32))
    for k, r in ref.items():
        if k == 'observation':
            continue
        g = out[k].cpu().numpy()
        assert g.shape == r.shape, (k, g.shape, r.shape)
        assert np.array_equal(g, r.astype(g.dtype)), k          # bit-exact: values, masks and indices


def test_sampler_and_ring_eviction():
    from handyrl_b200.replay import DeviceReplay
    case = BATCH_CASES['tictactoe']
    eps = case['episodes']
    steps = [e['steps'] for e in eps]
    replay = DeviceReplay(capacity_steps=sum(steps[:3]) + 1, max_episodes=64)
    for ep in eps:
        replay.add(ep)
        used = sorted((h.first_step, h.first_step + h.steps) for h in replay.handles)
        assert all(a[1] <= b[0] for a, b in zip(used, used[1:]))          # stored episodes never overlap
        assert used[-1][1] <= replay.capacity
    assert 1 <= len(replay) <= 3
    args = dict(case['args'], batch_size=16)
    random.seed(0)
    win = replay.sample_windows(16, args)
    out = replay.gather(win, args)
    torch.cuda.synchronize()
    em = out['episode_mask'][:, :, 0, 0].cpu().numpy()
    assert np.array_equal(em.sum(1), (win['end'] - win['start']).astype(np.float32))
    assert torch.all(out['action_mask'][out['episode_mask'][:, :, 0, 0] == 0] == 1e32)


def test_gathered_batch_feeds_the_loss_kernel():
    """gather/pad output -> fused loss == host make_batch -> fused loss, bit for bit."""
    from handyrl_b200 import ops
    from handyrl_b200.batch import make_batch
    from handyrl_b200.replay import DeviceReplay
    from handyrl_b200.synthetic import synthetic_outputs
    case = BATCH_CASES['tictactoe']
    args = dict(case['args'], **{'lambda': 0.7}, entropy_regularization=0.1, entropy_regularization_decay=0.1,
                policy_target='UPGO', value_target='VTRACE')
    replay = DeviceReplay(4096, 64)
    handles = [replay.add(ep) for ep in case['episodes']]
    dev = replay.gather(windows_for(case, replay, handles), args)
    host = {k: v.cuda() for k, v in make_batch(case['selected'], args).items() if k != 'observation'}
    outs = {k: v.cuda() for k, v in synthetic_outputs({'action_mask': host['action_mask']}, seed=5).items()}
    a = ops.loss_fwd_bwd(outs, dev, args)
    b = ops.loss_fwd_bwd(outs, host, args)
    torch.cuda.synchronize()
    assert torch.equal(a.losses, b.losses) and torch.equal(a.dpolicy, b.dpolicy) and torch.equal(a.dvalue, b.dvalue)
