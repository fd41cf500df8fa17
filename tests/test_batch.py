"""Host batching vs the reference's make_batch on real self-play episodes (golden: batch_cases.pkl)."""
import os
import pickle
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from handyrl_b200.batch import make_batch, tree_leaves, sample_window

with open(os.path.join(GOLDEN, 'batch_cases.pkl'), 'rb') as f:
    BATCH_CASES = pickle.load(f)


@pytest.mark.parametrize('name', sorted(BATCH_CASES))
def test_make_batch_matches_reference(name):
    case = BATCH_CASES[name]
    random.seed(9)   # the generator seeded `random` the same way before the reference call
    got = make_batch(case['selected'], case['args'])
    ref = case['batch']
    assert set(got) == set(ref)
    for k in ref:
        if k == 'observation':
            for g, r in zip(tree_leaves(got[k]), tree_leaves(ref[k])):
                assert g.shape == r.shape and g.dtype == torch.from_numpy(r).dtype
                assert np.array_equal(g.numpy(), r)
            continue
        g = got[k].numpy()
        assert g.shape == ref[k].shape, (k, g.shape, ref[k].shape)
        if k == 'selected_prob':
            # the reference yields float64 when a None was replaced by python 1.0 (SURVEY hard part 4)
            assert g.dtype == np.float32
            assert np.array_equal(g, ref[k].astype(np.float32))
        else:
            assert g.dtype == ref[k].dtype, (k, g.dtype, ref[k].dtype)
            assert np.array_equal(g, ref[k]), k      # bit-exact, masks and indices included


def test_sample_window_replays_reference_sampler():
    """Same `random` stream -> same windows as Batcher.select_episode (train.py:291-315)."""
    case = BATCH_CASES['geister_burnin']
    eps, args = case['episodes'], case['args']
    random.seed(5)
    for sel in case['selected']:
        idx, st, ed, tst, ep = sample_window(lambda: len(eps), lambda i: (eps[i]['steps'], eps[i]), args)
        assert ep is eps[idx]
        assert (st, ed, tst, eps[idx]['steps']) == (sel['start'], sel['end'], sel['train_start'], sel['total'])
