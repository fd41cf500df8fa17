"""Multi-rank parity on real GPUs (skipped below 2 visible devices; run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).

 * an N-rank sharded LearnerStep equals the single-GPU full-batch step (BatchNorm-free net: the loss is a SUM over the
   batch, train.py:202-213, so SUM of shard gradients is the full-batch gradient; clip and Adam see global quantities),
 * the fused NVLink peer-memory all-reduce (hrl_peer_allreduce_sumsq, world >= 2) agrees with the NCCL path,
 * all ranks hold bit-identical weights after every step,
 * the Trainer's own multi-GPU mode (helper processes, no torchrun) keeps ranks identical through epochs.
"""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
needs2 = pytest.mark.skipif(NGPU < 2, reason='needs at least 2 GPUs')


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


ARGS = {'turn_based_training': True, 'observation': False, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0, 'forward_steps': 8,
        'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1, 'policy_target': 'UPGO', 'value_target': 'VTRACE'}
DIMS = (16, 8, 2, 9)


def _net():
    from handyrl_b200.nets import BoardNet
    torch.manual_seed(11)
    return BoardNet(norm=False)


def _batch(s):
    from handyrl_b200.synthetic import synthetic_batch
    B, T, P, A = DIMS
    return synthetic_batch(B, T, P, A, turn_based=True, observation=False, seed=900 + s)


def _run_steps(stepper, batches):
    losses = []
    for b in batches:
        stepper.step(stepper.new_packed().fill(b))
        losses.append(stepper.read_losses())
    w = stepper.cpu_state_dict()
    return losses, torch.cat([v.reshape(-1).float() for v in w.values()])


def _rank_main(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from handyrl_b200.multigpu import shard_batch
    from handyrl_b200.train import LearnerStep
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    res = {}
    full = [_batch(s) for s in range(3)]
    for mode in ('peer', 'nccl'):
        stepper = LearnerStep(_net(), ARGS, shard_batch(full[0], rank, world), lr=1e-3, device=torch.device('cuda', rank),
                              process_group=dist.group.WORLD, peer_allreduce=(mode == 'peer'))
        assert (stepper.peer is not None) == (mode == 'peer')
        losses, w = _run_steps(stepper, [shard_batch(b, rank, world) for b in full])
        gathered = [torch.empty_like(w).cuda() for _ in range(world)]
        dist.all_gather(gathered, w.cuda())
        res[mode] = {'losses': losses, 'weights': w.numpy(),
                     'ranks_identical': all(torch.equal(g, gathered[0]) for g in gathered)}
        stepper.close()
    if rank == 0:
        single = LearnerStep(_net(), ARGS, full[0], lr=1e-3, device=torch.device('cuda', 0))
        losses, w = _run_steps(single, full)
        res['single'] = {'losses': losses, 'weights': w.numpy()}
    with open(os.path.join(out_dir, 'rank%d.pkl' % rank), 'wb') as f:
        pickle.dump(res, f)
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


@needs2
@pytest.mark.parametrize('world', [w for w in (2, 4, 8) if w <= max(NGPU, 2)])
def test_sharded_step_equals_full_batch_step(world):
    import torch.multiprocessing as mp
    out_dir = tempfile.mkdtemp(prefix='hrl_mgpu_')
    mp.spawn(_rank_main, args=(world, _free_port(), out_dir), nprocs=world, join=True)
    res = [pickle.load(open(os.path.join(out_dir, 'rank%d.pkl' % r), 'rb')) for r in range(world)]
    single = res[0]['single']
    for mode in ('peer', 'nccl'):
        for r in range(world):
            assert res[r][mode]['ranks_identical'], (mode, r)
            assert np.array_equal(res[r][mode]['weights'], res[0][mode]['weights']), (mode, r)     # bit-identical ranks
        for s in range(3):
            got, ref = res[0][mode]['losses'][s], single['losses'][s]
            scale = max(abs(v) for v in ref.values())
            for k, v in ref.items():           # the reduced bucket carries the GLOBAL loss sums and data count
                assert abs(got[k] - v) <= 1e-5 * scale + 1e-6, (mode, s, k, got[k], v)
            assert got['dcnt'] == ref['dcnt']
        # weights: 1e-6 everywhere except the rare element whose gradient sits at rounding-noise level, where Adam's
        # sign-like first steps may differ by up to ~lr per step between two summation orders
        diff = np.abs(res[0][mode]['weights'] - single['weights'])
        assert (diff > 1e-6).mean() <= 1e-3 and diff.max() <= 1e-4, (mode, (diff > 1e-6).sum(), diff.max())
    diff = np.abs(res[0]['peer']['weights'] - res[0]['nccl']['weights'])
    assert (diff > 1e-6).mean() <= 1e-3 and diff.max() <= 1e-4
    for s in range(3):
        for k, v in res[0]['nccl']['losses'][s].items():
            assert abs(res[0]['peer']['losses'][s][k] - v) <= 1e-5 * abs(v) + 1e-6


@needs2
def test_trainer_spawns_helper_ranks_and_keeps_them_identical():
    """Trainer(num_gpus=2) behind the reference's thread protocol: helper process on GPU 1, whole replay on both ranks,
    the weights (checksum) and the device-side learning rate of the helper equal rank 0's after every epoch."""
    import threading
    from handyrl_b200.nets import tictactoe_net
    from handyrl_b200.train import Trainer
    with open(os.path.join(GOLDEN, 'batch_cases.pkl'), 'rb') as f:
        case = pickle.load(f)['tictactoe']
    args = dict(case['args'], batch_size=8, minimum_episodes=4, num_batchers=1, **{'lambda': 0.7}, seed=3,
                entropy_regularization=0.1, entropy_regularization_decay=0.1, policy_target='UPGO', value_target='VTRACE',
                gpu_replay=True, num_gpus=2, multi_gpu_probe=True, multi_gpu_chunk=4)
    tr = Trainer(args, tictactoe_net())
    assert tr.world == 2
    tr.episodes.extend(case['episodes'])
    th = threading.Thread(target=tr.run, daemon=True)
    th.start()
    last = 0
    for epoch in range(3):
        if epoch == 1:
            tr.episodes.extend(case['episodes'][:3])          # episodes arriving mid-training reach every rank
        model, steps = tr.update()
        assert steps > last and steps % 4 == 0
        last = steps
        (helper_sum, helper_lr), = tr.fleet.collect_reports()
        mine = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).double()
        pad = torch.zeros(tr.stepper.state.n_pad - mine.numel(), dtype=torch.float64)
        assert abs(float(torch.cat([mine, pad]).sum()) - helper_sum) <= 1e-9 * max(1.0, abs(helper_sum))
        assert helper_lr == float(tr.stepper.opt.lr.item())
    tr.stop()
    th.join(timeout=30)
    assert not th.is_alive()
