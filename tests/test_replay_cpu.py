"""Host logic of the GPU replay store (ring bookkeeping, window sampling, episode tap) on CPU tensors."""
import os
import pickle
import random

import numpy as np
import pytest

from conftest import GOLDEN
from handyrl_b200.replay import DeviceReplay, WINDOW_DTYPE
from handyrl_b200.train import EpisodeDeque

with open(os.path.join(GOLDEN, 'batch_cases.pkl'), 'rb') as f:
    CASE = pickle.load(f)['tictactoe']


def test_ring_never_overlaps_and_drops_oldest():
    eps = CASE['episodes'] * 4
    for cap in (12, 17, 22, 30, 400):
        rp = DeviceReplay(capacity_steps=cap, max_episodes=5, device='cpu')
        for ep in eps:
            h = rp.add(ep)
            spans = sorted((x.first_step, x.first_step + x.steps) for x in rp.handles)
            assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
            assert spans[-1][1] <= cap and len(rp) <= 5
            assert rp.handles[-1] == h                       # newest last, like the reference deque
            # stored rows equal the decoded episode
            fe = __import__('handyrl_b200.batch', fromlist=['x']).flatten_moments(
                __import__('handyrl_b200.batch', fromlist=['x']).decode_moments(ep['moment']), ep['outcome'])
            assert np.array_equal(rp.st_prob[h.first_step:h.first_step + h.steps].numpy(), fe.prob)
            assert np.array_equal(rp.st_outcome[h.outcome_row].numpy(), fe.outcome)


def test_too_long_episode_is_refused():
    rp = DeviceReplay(capacity_steps=3, max_episodes=5, device='cpu')
    with pytest.raises(ValueError):
        rp.add(CASE['episodes'][0])


def test_sample_windows_follow_the_reference_sampler():
    """Same `random` stream -> the same (start, end, train_start) as Batcher.select_episode picked for the golden batch."""
    rp = DeviceReplay(capacity_steps=4096, max_episodes=1000, device='cpu')
    for ep in CASE['episodes']:
        rp.add(ep)
    random.seed(5)
    win = rp.sample_windows(len(CASE['selected']), CASE['args'])
    assert win.dtype == WINDOW_DTYPE
    for w, sel in zip(win, CASE['selected']):
        assert (w['start'], w['end'], w['train_start'], w['total']) == (sel['start'], sel['end'], sel['train_start'], sel['total'])


def test_episode_deque_taps_every_append_once():
    seen = []
    d = EpisodeDeque()
    d.append(1)
    d.listener = seen.append
    d.extend([2, 3])
    d.append(4)
    d.popleft()
    assert list(d) == [2, 3, 4] and seen == [2, 3, 4]


def test_vectorised_sampler_follows_the_reference_law():
    """numpy-Generator path of sample_windows: episode idx accepted with probability (idx+1)/count (train.py:294-297),
    train_start uniform over 1 + max(0, steps - forward_steps) candidates, window bounds as train.py:305-306."""
    rp = DeviceReplay(capacity_steps=4096, max_episodes=1000, device='cpu')
    for ep in CASE['episodes']:
        rp.add(ep)
    args = dict(CASE['args'], forward_steps=4, burn_in_steps=2)
    n = len(rp)
    win = rp.sample_windows(60000, args, np.random.default_rng(3))
    first = {h.first_step: i for i, h in enumerate(rp.handles)}
    idx = np.array([first[f] for f in win['first_step']])
    freq = np.bincount(idx, minlength=n) / len(idx)
    want = (np.arange(n) + 1) / (n * (n + 1) / 2)
    assert np.abs(freq - want).max() < 0.01
    steps = np.array([h.steps for h in rp.handles])[idx]
    assert np.array_equal(win['total'], steps)
    assert np.all(win['train_start'] >= 0) and np.all(win['train_start'] <= np.maximum(0, steps - 4))
    assert np.array_equal(win['start'], np.maximum(0, win['train_start'] - 2))
    assert np.array_equal(win['end'], np.minimum(win['train_start'] + 4, steps))
    longest = idx == int(np.argmax([h.steps for h in rp.handles]))
    ts = win['train_start'][longest]
    cand = 1 + max(0, max(h.steps for h in rp.handles) - 4)
    assert np.abs(np.bincount(ts, minlength=cand) / len(ts) - 1 / cand).max() < 0.02


def test_staged_upload_of_many_episodes_equals_one_by_one():
    from handyrl_b200.wire import episode_to_flat
    eps = CASE['episodes'] * 3
    for cap in (25, 60, 4096):
        a = DeviceReplay(capacity_steps=cap, max_episodes=7, device='cpu')
        b = DeviceReplay(capacity_steps=cap, max_episodes=7, device='cpu')
        for ep in eps:
            a.add(ep)
        hb = b.add_flat_many([episode_to_flat(ep) for ep in eps])
        assert a.handles == b.handles and hb[-len(b):] == b.handles
        for h in a.handles:
            rows = slice(h.first_step, h.first_step + h.steps)
            for col in ('st_obs', 'st_prob', 'st_action', 'st_amask', 'st_value', 'st_reward', 'st_return', 'st_flags', 'st_turn'):
                assert np.array_equal(getattr(a, col)[rows].numpy(), getattr(b, col)[rows].numpy()), col
            assert np.array_equal(a.st_outcome[h.outcome_row].numpy(), b.st_outcome[h.outcome_row].numpy())


def test_host_batcher_fetches_the_episode_once_and_survives_a_shifting_deque():
    """the window must fit the episode object it was drawn for even when the Learner pops episodes
    concurrently; an exception inside a batcher thread must not kill it."""
    from handyrl_b200.train import Batcher
    args = dict(CASE['args'], batch_size=4, num_batchers=1)
    eps = EpisodeDeque(CASE['episodes'])

    class Shifting(Batcher):
        def _fetch(self, idx):
            got = super()._fetch(idx)
            if len(self.episodes) > 3:
                self.episodes.popleft()        # the Learner trims the deque right after the read
            return got

    b = Shifting(args, eps)
    for _ in range(20):
        sel = b.select_episode()
        assert sel['total'] == sel['_episode']['steps'] and sel['end'] <= sel['total']
        eps.append(CASE['episodes'][0])
    batch = b._make()
    assert batch['action'].shape[0] == 4


def test_flat_cache_is_bounded():
    from handyrl_b200.train import _FlatCache
    c = _FlatCache(budget_bytes=3000)
    eps = [dict(e) for e in CASE['episodes']]
    fes = [c.get(e) for e in eps]
    assert c.used <= 3000 + max(_FlatCache._size(f) for f in fes) and len(c.items) < len(eps)
    assert c.get(eps[-1]) is fes[-1]            # most recent entry is still cached
