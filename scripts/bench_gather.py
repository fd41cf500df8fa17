"""Micro-benchmark of the replay gather/pad kernel (K2): synthetic replay store, cfg2-shaped batches."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from handyrl_b200.batch import FlatEpisode
from handyrl_b200.replay import DeviceReplay

def fake_episode(steps, Ps, A, obs_shape, rng):
    fe = FlatEpisode()
    fe.steps, fe.players = steps, list(range(Ps))
    fe.obs = rng.random((steps, Ps) + obs_shape, dtype=np.float32)
    fe.prob = rng.random((steps, Ps), dtype=np.float32)
    fe.action = rng.integers(0, A, (steps, Ps)).astype(np.int32)
    fe.amask = np.where(rng.random((steps, Ps, A)) < 0.3, 1e32, 0).astype(np.float32)
    fe.value = rng.random((steps, Ps, 1), dtype=np.float32)
    fe.reward = np.zeros((steps, Ps), np.float32); fe.ret = np.zeros((steps, Ps), np.float32)
    fe.flags = np.full((steps, Ps), 3, np.uint8); fe.turn = (np.arange(steps) % Ps).astype(np.int32)
    fe.outcome = np.array([1, -1][:Ps], np.float32)
    return fe

for name, (B, T, Ps, A, obs_shape, alt) in {'cfg2 (TicTacToe-like)': (512, 32, 2, 9, (3, 3, 3), True),
                                            'cfg5 shard (64x64 obs, 512 actions)': (512, 64, 2, 512, (1, 64, 64), True)}.items():
    rng = np.random.default_rng(0)
    replay = DeviceReplay(capacity_steps=200_000 if A < 100 else 40_000, max_episodes=4000)
    n_eps = 2000 if A < 100 else 300
    for _ in range(n_eps):
        replay.add_flat(fake_episode(int(rng.integers(T // 2, 2 * T)), Ps, A, obs_shape, rng))
    args = {'turn_based_training': True, 'observation': not alt, 'burn_in_steps': 0, 'forward_steps': T, 'maximum_episodes': 4000,
            'compress_steps': 4}
    random.seed(0)
    wins = [replay.sample_windows(B, args) for _ in range(16)]
    outs = [replay.empty_batch(B, args) for _ in range(4)]
    for i in range(4):
        replay.gather(wins[i], args, out=outs[i])
    torch.cuda.synchronize()
    wdev = [torch.from_numpy(w.view(np.uint8).reshape(B, -1)).cuda() for w in wins]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 64
    e0.record()
    for i in range(reps):
        replay.gather(wins[i % 16], args, out=outs[i % 4])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out_bytes = sum(v.numel() * v.element_size() for k, v in outs[0].items() if torch.is_tensor(v) and k != '_windows')
    live = float(outs[0]['episode_mask'].mean())
    alg = out_bytes * (1 + live)        # every batch byte written once + the live fraction read once from the store
    print('%s: %.1f us/batch (incl. host launch), batch %.2f MB, live fraction %.2f -> %.0f GB/s algorithmic' % (name, ms * 1e3, out_bytes / 1e6, live, alg / ms / 1e6))
