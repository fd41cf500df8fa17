"""Flat episode wire format.

The reference ships an episode as bz2-pickled blocks of per-step, per-player nested dicts
(handyrl/generation.py:84-91); the learner then walks those dicts in Python for every episode
(batch.flatten_moments here, make_batch in the reference).  `pack_episode` does that walk ONCE on the worker and
attaches the result -- step-major numpy arrays, exactly batch.FlatEpisode / the replay-store row layout -- under the
key 'flat' (one bz2-pickled dict).  Everything the reference's Learner reads stays in place ('args', 'steps',
'outcome', and 'moment' unless drop_moments=True), so Learner.feed_episodes (train.py:456-483) and the host Batcher
keep working; the GPU replay feeder uses 'flat' when present and skips the per-moment decode.
"""
import bz2
import pickle

import numpy as np

from .batch import FlatEpisode, decode_moments, flatten_moments, tree_leaves, tree_map

_FIELDS = ('prob', 'action', 'amask', 'value', 'reward', 'ret', 'flags', 'turn', 'outcome')


def pack_flat(fe):
    payload = {k: getattr(fe, k) for k in _FIELDS}
    payload['steps'], payload['players'], payload['obs'] = fe.steps, list(fe.players), fe.obs
    return bz2.compress(pickle.dumps(payload, protocol=pickle.HIGHEST_PROTOCOL))


def unpack_flat(blob):
    payload = pickle.loads(bz2.decompress(blob))
    fe = FlatEpisode()
    for k in _FIELDS + ('steps', 'players', 'obs'):
        setattr(fe, k, payload[k])
    return fe


def pack_episode(episode, drop_moments=False):
    """Reference episode dict -> the same dict plus 'flat' (and without 'moment' if drop_moments)."""
    if episode is None or 'flat' in episode:
        return episode
    fe = flatten_moments(decode_moments(episode['moment']), episode['outcome'])
    out = dict(episode)
    out['flat'] = pack_flat(fe)
    if drop_moments:
        out['moment'] = []
    return out


def episode_to_flat(episode):
    """FlatEpisode of an episode in either format."""
    if 'flat' in episode:
        return unpack_flat(episode['flat'])
    return flatten_moments(decode_moments(episode['moment']), episode['outcome'])


def install_worker_hook(drop_moments=False):
    """Wrap the reference's Generator.generate (generation.py:20-93) so that workers forked after this call ship
    flat episodes.  Returns the original method (to undo)."""
    import handyrl.generation as gen
    original = gen.Generator.generate
    if getattr(original, '_hrl_flat', False):
        return original

    def generate(self, models, args):
        return pack_episode(original(self, models, args), drop_moments=drop_moments)

    generate._hrl_flat = True
    gen.Generator.generate = generate
    return original
