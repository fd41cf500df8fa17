"""GPU-resident replay: episodes are decoded ONCE on arrival into flat device arrays; a training batch is
B window descriptors + one gather/pad kernel (csrc/gather_kernel.cu, C ABI hrl_gather_pad).

Replaces, for the learner, the reference's per-sample bz2+pickle decode and Python collation
(make_batch, handyrl/train.py:33-124) and the batcher process pool that ships pickled batches through
pipes (train.py:270-289, connection.py:133-173).  Sampling semantics (recency-biased episode choice,
window placement, burn-in) are those of Batcher.select_episode (train.py:291-315), shared with the
host path through batch.sample_window.

Store layout = include/hrl_b200.h HrlGatherArgs: one row per episode step, columns per player slot.
Nested observations are stored as the concatenation of their flattened leaves.
"""
import ctypes as C
import random
from collections import deque

import numpy as np
import torch

from . import _capi
from ._capi import HrlGatherArgs, HrlWindow, check, lib
from .batch import flatten_moments, decode_moments, sample_window, tree_leaves, tree_map

WINDOW_DTYPE = np.dtype([('first_step', '<i8'), ('start', '<i4'), ('end', '<i4'), ('train_start', '<i4'),
                         ('total', '<i4'), ('outcome_row', '<i4'), ('player', '<i4')])
assert WINDOW_DTYPE.itemsize == C.sizeof(HrlWindow)


class EpisodeHandle:
    __slots__ = ('first_step', 'steps', 'outcome_row')

    def __init__(self, first_step, steps, outcome_row):
        self.first_step, self.steps, self.outcome_row = first_step, steps, outcome_row


class DeviceReplay:
    """Ring of decoded episodes in HBM.

    capacity_steps bounds the stored steps, max_episodes the stored episodes (the reference trims its
    deque to `maximum_episodes`, train.py:474-483); when either is exceeded the oldest episodes go.
    """

    def __init__(self, capacity_steps, max_episodes, device='cuda'):
        self.device = torch.device(device)
        self.capacity = int(capacity_steps)
        self.max_episodes = int(max_episodes)
        self.handles = deque()
        self.write = 0
        self.ready = False
        self.next_outcome_row = 0

    def _allocate(self, fe):
        S, dev = self.capacity, self.device
        self.Ps = len(fe.players)
        self.A = fe.amask.shape[-1]
        self.obs_template = tree_map(lambda a: a[0, 0], fe.obs)                  # one observation, nested
        self.leaf_shapes = [tuple(l.shape[2:]) for l in tree_leaves(fe.obs)]
        self.leaf_sizes = [int(np.prod(s)) if len(s) else 1 for s in self.leaf_shapes]
        self.OE = int(sum(self.leaf_sizes))
        f = dict(dtype=torch.float32, device=dev)
        self.st_obs = torch.zeros((S, self.Ps, self.OE), **f)
        self.st_prob = torch.ones((S, self.Ps), **f)
        self.st_action = torch.zeros((S, self.Ps), dtype=torch.int32, device=dev)
        self.st_amask = torch.zeros((S, self.Ps, self.A), **f)
        self.st_value = torch.zeros((S, self.Ps), **f)
        self.st_reward = torch.zeros((S, self.Ps), **f)
        self.st_return = torch.zeros((S, self.Ps), **f)
        self.st_flags = torch.zeros((S, self.Ps), dtype=torch.uint8, device=dev)
        self.st_turn = torch.zeros((S,), dtype=torch.int32, device=dev)
        self.n_outcome_rows = self.max_episodes + 64
        self.st_outcome = torch.zeros((self.n_outcome_rows, self.Ps), **f)
        self.ready = True

    def __len__(self):
        return len(self.handles)

    def add(self, episode):
        """Decode one episode dict (the reference's wire format, generation.py:84-91, or the flat format of
        wire.py) and upload it."""
        from .wire import episode_to_flat
        return self.add_flat(episode_to_flat(episode))

    def add_flat(self, fe):
        if not self.ready:
            self._allocate(fe)
        n = fe.steps
        if n > self.capacity:
            raise ValueError('episode of %d steps exceeds the replay capacity of %d steps' % (n, self.capacity))
        if self.write + n > self.capacity:          # episodes are stored contiguously: wrap
            # the previous lap's episodes beyond the write pointer are the oldest ones: drop them so
            # that deque order == ring order again
            while self.handles and self.handles[0].first_step >= self.write:
                self.handles.popleft()
            self.write = 0
        lo, hi = self.write, self.write + n
        # evict whatever the new rows overwrite, and the oldest episode beyond max_episodes
        while self.handles and (len(self.handles) >= self.max_episodes or
                                (self.handles[0].first_step < hi and self.handles[0].first_step + self.handles[0].steps > lo)):
            self.handles.popleft()
        dev = self.device
        obs = np.concatenate([l.reshape(n, self.Ps, -1).astype(np.float32) for l in tree_leaves(fe.obs)], axis=2)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
        self.st_obs[lo:hi] = up(obs)
        self.st_prob[lo:hi] = up(fe.prob)
        self.st_action[lo:hi] = up(fe.action)
        self.st_amask[lo:hi] = up(fe.amask)
        self.st_value[lo:hi] = up(fe.value[..., 0])
        self.st_reward[lo:hi] = up(fe.reward)
        self.st_return[lo:hi] = up(fe.ret)
        self.st_flags[lo:hi] = up(fe.flags)
        self.st_turn[lo:hi] = up(fe.turn)
        row = self.next_outcome_row
        self.next_outcome_row = (row + 1) % self.n_outcome_rows
        self.st_outcome[row] = up(fe.outcome)
        h = EpisodeHandle(lo, n, row)
        self.handles.append(h)
        self.write = hi
        return h

    # ------------------------------------------------------------------ batches
    def batch_shapes(self, args):
        T = args['burn_in_steps'] + args['forward_steps']
        alternating = bool(args['turn_based_training'] and not args['observation'])
        P = 1 if not args['turn_based_training'] else self.Ps
        Pa = 1 if alternating else P
        return T, P, Pa, alternating

    def empty_batch(self, B, args):
        """Allocate the output tensors of one batch in the reference layout (train.py:114-124)."""
        T, P, Pa, _ = self.batch_shapes(args)
        f = dict(dtype=torch.float32, device=self.device)
        return {
            'observation': torch.empty((B, T, Pa, self.OE), **f),
            'selected_prob': torch.empty((B, T, Pa, 1), **f), 'value': torch.empty((B, T, P, 1), **f),
            'action': torch.empty((B, T, Pa, 1), dtype=torch.int64, device=self.device),
            'outcome': torch.empty((B, 1, P, 1), **f), 'reward': torch.empty((B, T, P, 1), **f),
            'return': torch.empty((B, T, P, 1), **f), 'episode_mask': torch.empty((B, T, 1, 1), **f),
            'turn_mask': torch.empty((B, T, P, 1), **f), 'observation_mask': torch.empty((B, T, P, 1), **f),
            'action_mask': torch.empty((B, T, Pa, self.A), **f), 'progress': torch.empty((B, T, 1), **f),
        }

    def sample_windows(self, B, args, rng=random):
        """B window descriptors drawn like Batcher.select_episode (train.py:291-315)."""
        win = np.zeros(B, WINDOW_DTYPE)
        solo = not args['turn_based_training']
        for b in range(B):
            idx, st, ed, tst = sample_window(lambda: len(self.handles), lambda i: self.handles[i].steps, args, rng)
            h = self.handles[idx]
            win[b] = (h.first_step, st, ed, tst, h.steps, h.outcome_row, 0)
        if solo:
            for b in range(B):          # make_batch draws the solo player per window, in order (train.py:57-58)
                win[b]['player'] = rng.choice(range(self.Ps))
        return win

    def gather(self, windows, args, out=None):
        """Run the gather/pad kernel for an array of WINDOW_DTYPE descriptors; returns the batch dict
        (observation as the concatenated-leaf tensor; see split_observation)."""
        B = len(windows)
        T, P, Pa, alternating = self.batch_shapes(args)
        if out is None:
            out = self.empty_batch(B, args)
        wdev = torch.from_numpy(windows.view(np.uint8).reshape(B, -1)).to(self.device, non_blocking=True)
        g = HrlGatherArgs()
        g.B, g.T, g.P, g.Pa, g.A, g.Ps = B, T, P, Pa, self.A, self.Ps
        g.burn_in = args['burn_in_steps']
        g.obs_elems = self.OE
        g.turn_alternating = int(alternating)
        ptr = lambda t: C.c_void_p(t.data_ptr())
        g.windows = ptr(wdev)
        g.st_obs, g.st_prob, g.st_action, g.st_amask = ptr(self.st_obs), ptr(self.st_prob), ptr(self.st_action), ptr(self.st_amask)
        g.st_value, g.st_reward, g.st_return = ptr(self.st_value), ptr(self.st_reward), ptr(self.st_return)
        g.st_flags, g.st_turn, g.st_outcome = ptr(self.st_flags), ptr(self.st_turn), ptr(self.st_outcome)
        g.observation, g.selected_prob, g.value, g.action = (ptr(out['observation']), ptr(out['selected_prob']),
                                                             ptr(out['value']), ptr(out['action']))
        g.outcome, g.reward, g.ret = ptr(out['outcome']), ptr(out['reward']), ptr(out['return'])
        g.episode_mask, g.turn_mask, g.observation_mask = ptr(out['episode_mask']), ptr(out['turn_mask']), ptr(out['observation_mask'])
        g.action_mask, g.progress = ptr(out['action_mask']), ptr(out['progress'])
        check(lib().hrl_gather_pad(C.byref(g), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        out['_windows'] = wdev      # keep the descriptor buffer alive until the kernel has run
        return out

    def split_observation(self, flat_obs):
        """(B,T,Pa,OE) concatenated leaves -> the env's nested observation structure."""
        pieces, off = [], 0
        for shape, size in zip(self.leaf_shapes, self.leaf_sizes):
            pieces.append(flat_obs[..., off:off + size].reshape(*flat_obs.shape[:3], *shape))
            off += size
        it = iter(pieces)
        return tree_map(lambda _: next(it), self.obs_template)
