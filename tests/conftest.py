import ast
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped, not failed, on a machine without a CUDA device (the product has no CPU path)."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason='needs a CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_cases(fname):
    """Golden .npz -> {case: {key: array}} (+ parsed 'meta' dict when present)."""
    z = np.load(os.path.join(GOLDEN, fname), allow_pickle=False)
    cases = {}
    for k in z.files:
        case, key = k.split('/', 1)
        cases.setdefault(case, {})[key] = z[k]
    for c in cases.values():
        if 'meta' in c:
            c['meta'] = ast.literal_eval(str(c['meta']))
    return cases


def case_args(meta):
    """The reference's train_args for a golden loss case."""
    return {
        'turn_based_training': meta['turn_based'], 'observation': meta['observation'],
        'gamma': meta['gamma'], 'lambda': meta['lmb'], 'burn_in_steps': meta['burn_in'],
        'forward_steps': meta['T'] - meta['burn_in'],
        'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1,
        'policy_target': meta['policy_target'], 'value_target': meta['value_target'],
    }


@pytest.fixture(scope='session')
def loss_cases():
    return load_cases('loss_cases.npz')


@pytest.fixture(scope='session')
def target_cases():
    return load_cases('target_cases.npz')


def net_case_setup(case):
    """(stand-in net loaded with the reference's initial weights, [three seeded batches]) for a net_step_cases.pkl case."""
    from handyrl_b200 import nets
    from handyrl_b200.synthetic import synthetic_geese_batch, synthetic_geister_batch
    B, T, P, A = case['dims']
    args = case['args']
    if case['net'] == 'geister':
        net = nets.load_state_by_order(nets.geister_net(), case['state0'])
        batches = [synthetic_geister_batch(B, T, P, A, turn_based=args['turn_based_training'], observation=args['observation'],
                                           burn_in=args['burn_in_steps'], seed=s) for s in case['seeds']]
    else:
        net = nets.load_state_by_order(nets.geese_net(), case['state0'])
        batches = [synthetic_geese_batch(B, T, P, A, seed=s) for s in case['seeds']]
    return net, batches


def noise_driven(case, key):
    """Weights whose gradient is analytically ZERO -- the bias of a convolution that feeds straight into BatchNorm
    (GeeseNet's TorusConv2d, hungry_geese.py:23-35) -- receive Adam updates of ~lr*sign(rounding noise): two equally valid
    fp32 runs disagree on them by up to 2*lr per step, and nothing downstream depends on them except that BatchNorm's
    running mean, which absorbs the bias.  Skipped in weight checks."""
    return case['net'] == 'geese' and (key.endswith('conv.bias') or key.endswith('bn.running_mean'))
