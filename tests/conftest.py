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
