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
            assert rp.handles[-1] is h                       # newest last, like the reference deque
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
