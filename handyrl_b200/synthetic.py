"""Seeded synthetic replay batches in the reference's batch-dict format.

The layout is the one `make_batch` emits (reference handyrl/train.py:114-124):
every tensor is (B, T, P-or-Pa, ...) batch-major, fp32 except `action` (int64).
The recipe is the one written down in SURVEY.md section 8(d) so that the bench, the
parity tests and the golden-vector generator all draw the same data.

Nothing here touches the GPU; callers move the dict with `.to(device)`.
"""

import torch


def synthetic_batch(B, T, P, A, *, turn_based=True, observation=False, reward_kind='zero',
                    gamma=0.8, seed=0, obs_shape=(3, 3, 3), burn_in=0, with_obs=True):
    """Build one replay batch.

    turn_based & not observation -> Pa = 1 (turn-alternating layout, train.py:65-66)
    otherwise                    -> Pa = P
    reward_kind: 'zero' (TicTacToe-like) or 'step' (-0.01 per step, Geister-like geister.py:436-438)
    burn_in > 0 also draws a random number of front-padded steps (train.py:94).
    """
    g = torch.Generator().manual_seed(seed)
    Pa = 1 if (turn_based and not observation) else P

    def randint(lo, hi, shape):  # inclusive bounds
        return torch.randint(lo, hi + 1, shape, generator=g)

    length = randint(max(1, T // 2), T, (B,))
    front = randint(0, burn_in, (B,)) if burn_in > 0 else torch.zeros(B, dtype=torch.long)
    length = torch.minimum(length, T - front)
    t_idx = torch.arange(T).unsqueeze(0)
    live = (t_idx >= front.unsqueeze(1)) & (t_idx < (front + length).unsqueeze(1))     # (B,T)
    emask = live.float()

    if turn_based:
        first = randint(0, P - 1, (B,))
        turn = (t_idx + first.unsqueeze(1)) % P                                          # (B,T)
        tmask = torch.nn.functional.one_hot(turn, P).float() * emask.unsqueeze(-1)       # (B,T,P)
    else:
        tmask = emask.unsqueeze(-1).expand(B, T, P).clone()
    omask = emask.unsqueeze(-1).expand(B, T, P).clone() if observation else tmask.clone()

    # which policy rows carry a real decision
    if Pa == 1:
        row_live = live.unsqueeze(-1)                                                    # (B,T,1)
    else:
        row_live = tmask > 0                                                             # (B,T,P)

    legal = torch.rand((B, T, Pa, A), generator=g) < 0.7
    legal[..., 0] = True
    legal = legal & row_live.unsqueeze(-1)
    amask = (~legal).float() * 1e32

    # uniform draw among the legal actions; 0 on dead rows
    score = torch.rand((B, T, Pa, A), generator=g) - (~legal).float() * 2.0
    action = score.argmax(-1)
    action = torch.where(row_live, action, torch.zeros_like(action))

    prob = torch.rand((B, T, Pa), generator=g).clamp(0.05, 1.0)
    prob = torch.where(row_live, prob, torch.ones_like(prob))

    if P == 2:
        o = randint(-1, 1, (B,)).float()
        outcome = torch.stack([o, -o], dim=1)
    else:
        outcome = randint(-1, 1, (B, P)).float()

    value = torch.tanh(torch.randn((B, T, P), generator=g)) * omask
    ended = t_idx >= (front + length).unsqueeze(1)
    value = torch.where(ended.unsqueeze(-1), outcome.unsqueeze(1).expand(B, T, P), value)

    if reward_kind == 'zero':
        reward = torch.zeros(B, T, P)
        ret = torch.zeros(B, T, P)
    else:
        reward = -0.01 * emask.unsqueeze(-1).expand(B, T, P).clone()
        ret = torch.zeros(B, T, P)
        acc = torch.zeros(B, P)
        for t in range(T - 1, -1, -1):
            acc = reward[:, t] + gamma * acc
            ret[:, t] = acc
        ret = ret * emask.unsqueeze(-1)

    step = (t_idx - front.unsqueeze(1)).clamp(min=0).float()
    progress = torch.where(live, step / length.unsqueeze(1).float(), torch.ones(B, T))

    batch = {
        'selected_prob': prob.unsqueeze(-1).contiguous(),
        'value': value.unsqueeze(-1).contiguous(),
        'action': action.unsqueeze(-1).contiguous(),
        'outcome': outcome.view(B, 1, P, 1).contiguous(),
        'reward': reward.unsqueeze(-1).contiguous(),
        'return': ret.unsqueeze(-1).contiguous(),
        'episode_mask': emask.view(B, T, 1, 1).contiguous(),
        'turn_mask': tmask.unsqueeze(-1).contiguous(),
        'observation_mask': omask.unsqueeze(-1).contiguous(),
        'action_mask': amask.contiguous(),
        'progress': progress.unsqueeze(-1).contiguous(),
    }
    if with_obs:
        obs = (torch.rand((B, T, Pa) + tuple(obs_shape), generator=g) < 0.3).float()
        batch['observation'] = obs * row_live.view(B, T, Pa, *([1] * len(obs_shape))).float()
    return batch


def synthetic_geister_batch(B, T, P, A, *, turn_based=True, observation=True, burn_in=0, seed=0):
    """Geister-shaped batch: step rewards/returns and a dict observation {'scalar': (18,), 'board': (7,6,6)}
    (the observation structure of geister.py:131-167)."""
    batch = synthetic_batch(B, T, P, A, turn_based=turn_based, observation=observation, reward_kind='step',
                            seed=seed, burn_in=burn_in, with_obs=False)
    Pa = batch['action'].shape[2]
    g = torch.Generator().manual_seed(seed + 7)
    batch['observation'] = {'scalar': torch.rand((B, T, Pa, 18), generator=g),
                            'board': (torch.rand((B, T, Pa, 7, 6, 6), generator=g) < 0.3).float()}
    return batch


def synthetic_geese_batch(B, T, P, A=4, *, seed=0, board=(7, 11), planes=17):
    """Hungry-Geese-shaped batch (simultaneous layout, P = Pa players): 17 x 7 x 11 observations whose plane 0 marks
    exactly one cell (the player's head, which the net's policy head reads, hungry_geese.py:50) and whose other
    planes are sparse occupancy maps."""
    batch = synthetic_batch(B, T, P, A, turn_based=False, observation=False, seed=seed, with_obs=False)
    g = torch.Generator().manual_seed(seed + 11)
    cells = board[0] * board[1]
    obs = (torch.rand((B, T, P, planes, cells), generator=g) < 0.08).float()
    head = torch.randint(0, cells, (B, T, P), generator=g)
    obs[..., 0, :] = torch.nn.functional.one_hot(head, cells).float()
    live = (batch['turn_mask'] > 0).view(B, T, P, 1, 1).float()
    batch['observation'] = (obs * live).view(B, T, P, planes, *board)
    return batch


def tictactoe_episodes(n, *, seed=0, compress_steps=4, gamma=0.8):
    """n self-play games of 3x3 noughts-and-crosses between uniformly random players, in the reference's episode wire
    format (generation.py:84-91: {'args', 'steps', 'outcome', 'moment': [bz2(pickle(list of moments))...]}, a moment
    holding per-player observation / selected_prob / action_mask / action / value / reward / return and 'turn';
    only the turn player acts and observes, generation.py:35-40).  Stands in for the worker processes as the episode
    source of bench.py's trainer leg: the learner consumes these exactly as it consumes the workers' episodes."""
    import bz2
    import pickle
    import random
    import numpy as np
    rng = random.Random(seed)
    lines = [(0, 1, 2), (3, 4, 5), (6, 7, 8), (0, 3, 6), (1, 4, 7), (2, 5, 8), (0, 4, 8), (2, 4, 6)]
    keys = ('observation', 'selected_prob', 'action_mask', 'action', 'value', 'reward', 'return')
    episodes = []
    for _ in range(n):
        board = [0] * 9                       # +1 first player's stones, -1 second player's
        moments, winner = [], None
        for ply in range(9):
            player = ply % 2
            colour = 1 if player == 0 else -1
            legal = [a for a in range(9) if board[a] == 0]
            grid = np.array(board, np.float32).reshape(3, 3)
            obs = np.stack([np.ones((3, 3), np.float32), (grid == colour).astype(np.float32), (grid == -colour).astype(np.float32)])
            mask = np.full(9, 1e32, np.float32)
            mask[legal] = 0
            action = rng.choice(legal)
            m = {k: {0: None, 1: None} for k in keys}
            m['observation'][player] = obs
            m['selected_prob'][player] = np.float32(1.0 / len(legal))
            m['action_mask'][player] = mask
            m['action'][player] = action
            m['value'][player] = np.array([rng.uniform(-0.5, 0.5)], np.float32)
            m['reward'] = {0: 0, 1: 0}
            m['return'] = {0: 0.0, 1: 0.0}
            m['turn'] = [player]
            moments.append(m)
            board[action] = colour
            if any(board[a] == board[b] == board[c] == colour for a, b, c in lines):
                winner = player
                break
        outcome = {0: 0, 1: 0} if winner is None else {winner: 1, 1 - winner: -1}
        episodes.append({'args': {'role': 'g', 'player': [0, 1], 'model_id': {0: 0, 1: 0}}, 'steps': len(moments), 'outcome': outcome,
                         'moment': [bz2.compress(pickle.dumps(moments[i:i + compress_steps]))
                                    for i in range(0, len(moments), compress_steps)]})
    return episodes


def synthetic_outputs(batch, *, has_value=True, has_return=False, seed=1):
    """Raw net outputs for kernel-only tests: policy ~ N(0,1), value = tanh(N(0,1))."""
    g = torch.Generator().manual_seed(seed)
    B, T, Pa, A = batch['action_mask'].shape
    out = {'policy': torch.randn((B, T, Pa, A), generator=g)}
    if has_value:
        out['value'] = torch.tanh(torch.randn((B, T, Pa, 1), generator=g))
    if has_return:
        out['return'] = 0.5 * torch.randn((B, T, Pa, 1), generator=g)
    return out


def bytes_per_cell(P, Pa, A, T, R):
    """Algorithmic bytes of the fused loss pass per (b,t) cell, SURVEY.md 8(d)."""
    return 12 * Pa * A + Pa * (20 + 8 * R) + 16 * P + 8 + 4.0 * P / T
