"""Flat episode wire format (section 8 f-2): lossless w.r.t. the per-moment decode, backward compatible."""
import os
import pickle
import sys

import numpy as np
import pytest

from conftest import GOLDEN
from handyrl_b200 import wire
from handyrl_b200.batch import decode_moments, flatten_moments, tree_leaves, make_batch
from handyrl_b200.replay import DeviceReplay

with open(os.path.join(GOLDEN, 'batch_cases.pkl'), 'rb') as f:
    CASES = pickle.load(f)
REF = os.environ.get('HANDYRL_REFERENCE', '/root/reference')


def same_flat(a, b):
    assert a.steps == b.steps and list(a.players) == list(b.players)
    for k in wire._FIELDS:
        x, y = getattr(a, k), getattr(b, k)
        assert x.dtype == y.dtype and np.array_equal(x, y), k
    for x, y in zip(tree_leaves(a.obs), tree_leaves(b.obs)):
        assert x.dtype == y.dtype and np.array_equal(x, y)


@pytest.mark.parametrize('name', sorted(CASES))
def test_pack_unpack_is_lossless(name):
    for ep in CASES[name]['episodes']:
        want = flatten_moments(decode_moments(ep['moment']), ep['outcome'])
        packed = wire.pack_episode(ep)
        assert packed['moment'] == ep['moment'] and packed['steps'] == ep['steps'] and packed['outcome'] == ep['outcome']
        same_flat(wire.episode_to_flat(packed), want)
        slim = wire.pack_episode(ep, drop_moments=True)
        assert slim['moment'] == [] and 'flat' in slim
        same_flat(wire.episode_to_flat(slim), want)
        assert wire.pack_episode(packed) is packed           # idempotent


def test_replay_accepts_both_formats():
    eps = CASES['geister_burnin']['episodes']
    a = DeviceReplay(4096, 64, device='cpu')
    b = DeviceReplay(4096, 64, device='cpu')
    for ep in eps:
        a.add(ep)
        b.add(wire.pack_episode(ep, drop_moments=True))
    for k in ('st_obs', 'st_prob', 'st_action', 'st_amask', 'st_value', 'st_reward', 'st_return', 'st_flags', 'st_turn', 'st_outcome'):
        assert (getattr(a, k) == getattr(b, k)).all(), k


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'handyrl')), reason='reference checkout not mounted')
def test_worker_hook_makes_the_reference_generator_ship_flat_episodes():
    sys.path.insert(0, REF)
    try:
        import random
        from handyrl.environment import make_env, prepare_env
        from handyrl.generation import Generator
        import handyrl.generation as gen
        from handyrl.model import ModelWrapper
        original = wire.install_worker_hook()
        try:
            env_args = {'env': 'TicTacToe'}
            prepare_env(env_args)
            env = make_env(env_args)
            model = ModelWrapper(env.net())
            random.seed(3)
            ep = Generator(env, {'gamma': 0.8, 'compress_steps': 4}).generate({p: model for p in env.players()},
                                                                              {'player': env.players(), 'model_id': {}})
            assert ep is not None and 'flat' in ep and ep['steps'] == len(decode_moments(ep['moment']))
            same_flat(wire.unpack_flat(ep['flat']), flatten_moments(decode_moments(ep['moment']), ep['outcome']))
        finally:
            gen.Generator.generate = original
    finally:
        sys.path.remove(REF)
        for m in [m for m in sys.modules if m == 'handyrl' or m.startswith('handyrl.')]:
            del sys.modules[m]
